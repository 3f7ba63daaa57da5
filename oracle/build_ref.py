#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds the reference's own software backend as the oracle.

Compiles the UNMODIFIED translation units of lynx-family/skity's CPU raster path
from where they lie under /root/reference (never copied into this repo) with
plain g++, plus oracle/ref_driver.cc, into oracle/_ref/libskity_ref.so.

The reference's stock CMake build cannot run here: it needs un-vendored
third-party trees (glm, freetype, libpng, ... — cmake/ThirdPartyDep.cmake:8).
The raster path itself only needs glm, for which third_party/glm_shim/ is a
stand-in (see its header).  Text entry points (Font/Typeface/TextBlob ...),
never reached by path/gradient/blur/clip draws, are left undefined by those TUs
and are satisfied by generated abort() stubs.

Usage: python oracle/build_ref.py [--ref /root/reference] [--opt -O2] [--out oracle/_ref]
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

# Translation units of the reference compiled into the oracle.
TUS = """
src/render/sw/sw_raster.cc src/render/sw/sw_edge.cc src/render/sw/sw_span_brush.cc
src/render/sw/sw_render_target.cc src/render/sw/sw_stack_blur.cc src/render/sw/sw_canvas.cc
src/render/sw/sw_a8_drawable.cc
src/render/hw/coverage/coverage_aa_tiler.cc src/render/hw/coverage/coverage_aa_line_encoder.cc
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


IO_TUS = """
module/io/src/io/flat/blender_flat.cc module/io/src/io/flat/blob_flat.cc module/io/src/io/flat/color_filter_flat.cc
module/io/src/io/flat/font_desc_flat.cc module/io/src/io/flat/font_flat.cc module/io/src/io/flat/image_filter_flat.cc
module/io/src/io/flat/local_matrix_flat.cc module/io/src/io/flat/mask_filter_flat.cc module/io/src/io/flat/matrix_flat.cc
module/io/src/io/flat/paint_flat.cc module/io/src/io/flat/path_flat.cc module/io/src/io/flat/path_effect_flat.cc
module/io/src/io/flat/rrect_flat.cc module/io/src/io/flat/shader_flat.cc module/io/src/io/flat/vertices_flat.cc
module/io/src/io/read/read_typeface.cc module/io/src/io/memory_read.cc module/io/src/io/memory_writer.cc
module/io/src/record/record_playback.cc module/io/src/stream/file_read_stream.cc module/io/src/stream/file_write_stream.cc
module/io/src/stream/stream.cc module/io/src/utils/parse_path.cc module/io/src/picture.cc
""".split()   # module/io (.skp), as in skity_b200/build.py


def cxx_flags(ref, opt, march=None):
    return [
        "-std=c++17", opt, *([f"-march={march}", "-ffp-contract=off"] if march else []), "-fPIC", "-w", "-DSKITY_CPU", "-DSKITY_RELEASE", "-DNDEBUG",
        "-fno-exceptions", "-fno-rtti",
        # default build: no -march / -ffast-math, x86-64 baseline float semantics.  The timing build (build_fast:
        # -O3 -march=x86-64-v3, SURVEY 8d's "-O3 -march=native" made portable to the GPU box's host) keeps ISO C++
        # (-std=c++17 => -ffp-contract=off), so it renders the same bytes — checked by tests/test_oracle_pinning.py
        f"-I{ref}", f"-I{ref}/include", f"-I{ref}/module/wgx/include",
        f"-I{REPO}/third_party/glm_shim", f"-I{REPO}",
        f"-I{ref}/module/io/include", f"-I{ref}/module/io", f"-I{ref}/module/io/src", f"-I{ref}/module/codec/include",
    ]


def compile_one(args):
    src, obj, flags = args
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return None
    r = subprocess.run(["g++", *flags, "-c", src, "-o", obj], capture_output=True, text=True)
    return (src, r.stderr) if r.returncode else None


def gen_stubs(objs, out_c):
    """abort() stubs for the text symbols the compiled TUs leave undefined."""
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
        f.write("/* generated by oracle/build_ref.py: reference text entry points that the raster path never reaches */\n")
        f.write("#include <stdio.h>\n#include <stdlib.h>\n")
        for s in missing:
            f.write(f'void {s}(void) {{ fprintf(stderr, "oracle stub reached: {s}\\n"); abort(); }}\n')
    return missing


def build(ref="/root/reference", out=None, opt="-O2", jobs=None, name="libskity_ref.so",
          driver=None, verbose=True, march=None):
    out = out or os.path.join(HERE, "_ref")
    driver = driver or os.path.join(HERE, "ref_driver.cc")
    objdir = os.path.join(out, "obj" + opt.replace("-", "_") + ("_" + march.replace("-", "_") if march else ""))
    os.makedirs(objdir, exist_ok=True)
    flags = cxx_flags(ref, opt, march)
    work = []
    for tu in TUS:
        src = os.path.join(ref, tu)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        work.append((src, os.path.join(objdir, tu.replace("/", "_") + ".o"), flags))
    for tu in IO_TUS:   # picture.cc uses std::memcpy without including <cstring>
        work.append((os.path.join(ref, tu), os.path.join(objdir, tu.replace("/", "_") + ".o"), flags + ["-include", "cstring"]))
    drv_obj = os.path.join(objdir, os.path.basename(driver) + ".o")
    # the driver replays scenes with the repo's scene player: rebuild it when that header changes
    player = os.path.join(REPO, "skity_b200", "host", "scene_player.hpp")
    if os.path.exists(drv_obj) and os.path.exists(player) and os.path.getmtime(drv_obj) < os.path.getmtime(player):
        os.remove(drv_obj)
    work.append((driver, drv_obj, flags))
    with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count()) as ex:
        errs = [e for e in ex.map(compile_one, work) if e]
    if errs:
        for src, err in errs:
            sys.stderr.write(f"FAILED {src}\n{err[:2000]}\n")
        raise RuntimeError("reference oracle build failed")
    objs = [w[1] for w in work]
    stubs_c = os.path.join(objdir, "text_stubs.c")
    missing = gen_stubs(objs, stubs_c)
    stubs_o = stubs_c[:-2] + ".o"
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-w", "-c", stubs_c, "-o", stubs_o])
    lib = os.path.join(out, name)
    subprocess.check_call(["g++", "-shared", "-o", lib, *objs, stubs_o, "-Wl,--no-undefined", "-lpthread"])
    if verbose:
        print(f"built {lib} ({len(TUS)} reference + {len(IO_TUS)} module/io TUs, {len(missing)} text stubs)")
    return lib


FAST_NAME = "libskity_ref_O3.so"


def build_fast(ref="/root/reference", out=None, verbose=True):
    """The CPU-baseline build of the reference: -O3 -march=x86-64-v3 (AVX2 + FMA + BMI2: what every host of a B200 box
    has; -march=native of the build container would not be portable to it)."""
    return build(ref, out, opt="-O3", name=FAST_NAME, verbose=verbose, march="x86-64-v3")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=None)
    ap.add_argument("--opt", default="-O2")
    ap.add_argument("--fast", action="store_true", help="also build the -O3 -march=x86-64-v3 timing variant")
    a = ap.parse_args()
    build(a.ref, a.out, a.opt)
    if a.fast:
        build_fast(a.ref, a.out)
