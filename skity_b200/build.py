"""Build recipes for the B200 backend (all outputs in-tree, git-ignored, shipped by gpurun).

  build_cuda()  nvcc  skity_b200/csrc/*.cu              -> skity_b200/lib/libskb.so        (C ABI, include/skb.h)
  build_host()  g++   skity_b200/host/*.cc + skity core -> skity_b200/lib/libskb_skity.so  (skity::Canvas plug-in)

The host library is a skity plug-in, so it links skity's own core classes
(Path, Paint, Matrix, Stroke, Canvas, shaders ...), compiled from the reference
sources where they lie (never copied).  The reference's software rasteriser
(src/render/sw/*) is NOT linked: the few symbols of it that core TUs mention
resolve to stubs that abort — there is no CPU fallback in the product.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
LIBDIR = os.path.join(PKG, "lib")
BUILD = os.path.join(REPO, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact parity with the CPU path needs one rounding per float op: no FMA contraction,
    # IEEE division and square root (the defaults, stated explicitly)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]

# skity core translation units the plug-in links against (no src/render/sw/*).
SKITY_CORE_TUS = """
src/render/canvas.cc src/render/canvas_state.cc
src/geometry/stroke.cc src/geometry/matrix.cc src/geometry/conic.cc src/geometry/cubic.cc
src/geometry/geometry.cc src/geometry/rect.cc src/geometry/rrect.cc
src/graphic/path.cc src/graphic/path_priv.cc src/graphic/path_visitor.cc src/graphic/path_scanner.cc
src/graphic/paint.cc src/graphic/bitmap.cc src/graphic/bitmap_sampler.cc src/graphic/blend_mode.cc
src/graphic/color.cc src/graphic/color_priv.cc src/graphic/image.cc src/graphic/contour_measure.cc
src/graphic/path_measure.cc
src/effect/image_filter.cc src/effect/mask_filter.cc src/effect/shader.cc src/effect/gradient_shader.cc
src/effect/pixmap_shader.cc src/effect/color_filter.cc src/effect/path_effect.cc
src/effect/dash_path_effect.cc src/effect/discrete_path_effect.cc
src/io/data.cc src/io/pixmap.cc
src/recorder/picture_recorder.cc src/recorder/recording_canvas.cc src/recorder/display_list.cc
src/recorder/display_list_region.cc src/recorder/display_list_rtree.cc
src/gpu/gpu_texture.cc src/utils/arena_allocator.cc src/logging.cc src/tracing.cc
src/base/mapping.cc src/base/unique_fd.cc
src/base/platform/posix/file_posix.cc src/base/platform/posix/mapping_posix.cc
""".split()

# The reference's .skp reader / writer (module/io/CMakeLists.txt:5-43): linked so that serialized pictures can be played
# onto the canvas (skp_player.hpp).  Its text and image-codec references resolve to the aborting stubs like the core's.
SKITY_IO_TUS = """
module/io/src/io/flat/blender_flat.cc module/io/src/io/flat/blob_flat.cc module/io/src/io/flat/color_filter_flat.cc
module/io/src/io/flat/font_desc_flat.cc module/io/src/io/flat/font_flat.cc module/io/src/io/flat/image_filter_flat.cc
module/io/src/io/flat/local_matrix_flat.cc module/io/src/io/flat/mask_filter_flat.cc module/io/src/io/flat/matrix_flat.cc
module/io/src/io/flat/paint_flat.cc module/io/src/io/flat/path_flat.cc module/io/src/io/flat/path_effect_flat.cc
module/io/src/io/flat/rrect_flat.cc module/io/src/io/flat/shader_flat.cc module/io/src/io/flat/vertices_flat.cc
module/io/src/io/read/read_typeface.cc module/io/src/io/memory_read.cc module/io/src/io/memory_writer.cc
module/io/src/record/record_playback.cc module/io/src/stream/file_read_stream.cc module/io/src/stream/file_write_stream.cc
module/io/src/stream/stream.cc module/io/src/utils/parse_path.cc module/io/src/picture.cc
""".split()
IO_FLAGS = ["-include", "cstring"]   # module/io/src/picture.cc uses std::memcpy without including <cstring>


def io_include_flags(ref):
    return [f"-I{ref}/module/io/include", f"-I{ref}/module/io", f"-I{ref}/module/io/src", f"-I{ref}/module/codec/include"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-8000:] + "\n")
        raise RuntimeError(f"command failed: {cmd[0]}")
    return r


def _compile(job):
    src, obj, cmd = job
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return
    _run(cmd + ["-c", src, "-o", obj])


def cuda_sources():
    return sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))


def cuda_lib_path():
    return os.path.join(LIBDIR, "libskb.so")


def host_lib_path():
    return os.path.join(LIBDIR, "libskb_skity.so")


def build_cuda(verbose=True, extra_flags=()):
    """Compile every CUDA source for sm_100a into the C-ABI library."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(BUILD, "cuda")
    os.makedirs(objdir, exist_ok=True)
    srcs = cuda_sources()
    hdrs = glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + glob.glob(os.path.join(PKG, "csrc", "*.h")) + \
        glob.glob(os.path.join(REPO, "include", "*.h"))
    newest_hdr = max([os.path.getmtime(h) for h in hdrs] + [0])
    jobs = []
    for s in srcs:
        obj = os.path.join(objdir, os.path.basename(s) + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) < newest_hdr:
            os.remove(obj)
        jobs.append((s, obj, ["nvcc", *NVCC_FLAGS, *extra_flags, f"-I{REPO}", f"-I{REPO}/include"]))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        list(ex.map(_compile, jobs))
    lib = cuda_lib_path()
    _run(["nvcc", "-shared", "-o", lib, *[j[1] for j in jobs], "-cudart", "static", "-Xlinker", "--no-undefined"])
    if verbose:
        print(f"built {lib} from {len(srcs)} CUDA sources")
    return lib


def _gen_stubs(objs, out_c):
    undefined, defined = set(), set()
    for o in objs:
        for line in subprocess.run(["nm", o], capture_output=True, text=True).stdout.splitlines():
            parts = line.split()
            if len(parts) == 2 and parts[0] == "U":
                undefined.add(parts[1])
            elif len(parts) == 3 and parts[1] in "TWVBDRtdbr":
                defined.add(parts[2])
    missing = sorted(s for s in undefined - defined if s.startswith("_ZN5skity") or s.startswith("_ZNK5skity"))
    with open(out_c, "w") as f:
        f.write("/* generated by skity_b200/build.py: text and CPU-raster entry points of skity that the CUDA\n"
                "   plug-in never reaches; they abort — the product has no CPU fallback. */\n")
        f.write("#include <stdio.h>\n#include <stdlib.h>\n")
        for s in missing:
            f.write(f'void {s}(void) {{ fprintf(stderr, "skb: unreachable skity entry point %s\\n", "{s}"); abort(); }}\n')
    return missing


def build_host(ref="/root/reference", verbose=True):
    """Compile the skity::Canvas plug-in.  Needs the reference tree (headers + core TUs)."""
    if not os.path.isdir(ref):
        raise FileNotFoundError(f"{ref} (the prebuilt {host_lib_path()} is used where the reference tree is absent)")
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(BUILD, "skity_host")
    os.makedirs(objdir, exist_ok=True)
    flags = ["g++", "-std=c++17", "-O2", "-fPIC", "-w", "-DSKITY_CPU", "-DSKITY_RELEASE", "-DNDEBUG",
             "-fno-exceptions", "-fno-rtti", f"-I{ref}", f"-I{ref}/include", f"-I{ref}/module/wgx/include",
             f"-I{REPO}/third_party/glm_shim", f"-I{REPO}", f"-I{REPO}/include"]
    flags = flags + io_include_flags(ref)
    jobs = [(os.path.join(ref, tu), os.path.join(objdir, tu.replace("/", "_") + ".o"), flags) for tu in SKITY_CORE_TUS]
    jobs += [(os.path.join(ref, tu), os.path.join(objdir, tu.replace("/", "_") + ".o"), flags + IO_FLAGS) for tu in SKITY_IO_TUS]
    own = sorted(glob.glob(os.path.join(PKG, "host", "*.cc")))
    hdr_time = max(os.path.getmtime(h) for h in glob.glob(os.path.join(PKG, "host", "*.hpp")) +
                   glob.glob(os.path.join(REPO, "include", "*.h")))
    for s in own:
        obj = os.path.join(objdir, "skb_" + os.path.basename(s) + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) < hdr_time:
            os.remove(obj)
        jobs.append((s, obj, flags))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        list(ex.map(_compile, jobs))
    objs = [j[1] for j in jobs]
    stubs_c = os.path.join(objdir, "unreachable_stubs.c")
    missing = _gen_stubs(objs, stubs_c)
    stubs_o = stubs_c[:-2] + ".o"
    _run(["gcc", "-O1", "-fPIC", "-w", "-c", stubs_c, "-o", stubs_o])
    lib = host_lib_path()
    if not os.path.exists(cuda_lib_path()):
        raise RuntimeError("build the CUDA library first (build_cuda): the plug-in links its C ABI")
    link = ["g++", "-shared", "-o", lib, *objs, stubs_o, f"-L{LIBDIR}", "-l:libskb.so", "-Wl,-rpath,$ORIGIN",
            "-Wl,--no-undefined", "-lpthread", "-ldl"]
    _run(link)
    if verbose:
        print(f"built {lib} ({len(SKITY_CORE_TUS)} skity core + {len(SKITY_IO_TUS)} module/io TUs + {len(own)} plug-in sources, {len(missing)} stubs)")
    return lib


if __name__ == "__main__":
    what = sys.argv[1:] or ["cuda", "host"]
    if "cuda" in what:
        build_cuda()
    if "host" in what:
        build_host()
