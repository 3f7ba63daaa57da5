"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libskity_ref.so (the
reference's own software backend, built by oracle/build_ref.py).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libskity_ref.so")           # -O2, baseline x86-64: the parity authority
FAST_LIB_PATH = os.path.join(_HERE, "_ref", "libskity_ref_O3.so")   # -O3 -march=x86-64-v3: the CPU-baseline timing build
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def fast_available():
    """The -O3 -march=x86-64-v3 build exists and this host can run it (AVX2, FMA, BMI2)."""
    if not os.path.exists(FAST_LIB_PATH):
        return False
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return all(f in flags for f in ("avx2", "fma", "bmi2", "movbe", "f16c"))


def use_fast_build():
    """Switch this process to the timing build (before the first call; bench.py's CPU-baseline legs only)."""
    global LIB_PATH, _lib
    if _lib is not None:
        raise RuntimeError("refsw: library already loaded")
    LIB_PATH = FAST_LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ref_render_scene.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_double)]
        _lib.ref_render_scene.restype = ctypes.c_int
        _lib.ref_render_skp.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float),
                                        ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        _lib.ref_render_skp.restype = ctypes.c_int
        _lib.ref_raster_path.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        _lib.ref_raster_path.restype = ctypes.c_long
        _lib.ref_stack_blur.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                        ctypes.c_void_p]
        _lib.ref_stack_blur.restype = ctypes.c_int
        _lib.ref_coverage_aa_tile.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                              ctypes.POINTER(ctypes.c_long)]
        _lib.ref_coverage_aa_tile.restype = ctypes.c_long
    return _lib


def render_scene(blob, return_seconds=False):
    """Render an SKSC blob with the reference SW canvas -> (H, W, 4) uint8 premul RGBA."""
    _, _, w, h, _, _ = struct.unpack_from("<6I", blob, 0)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    sec = ctypes.c_double(0.0)
    rc = lib().ref_render_scene(blob, len(blob), out.ctypes.data, ctypes.byref(sec))
    if rc != 0:
        raise RuntimeError(f"ref_render_scene failed: {rc}")
    return (out, sec.value) if return_seconds else out


def render_skp(skp, width, height, matrix6=(1, 0, 0, 0, 1, 0), return_seconds=False):
    """A serialized picture (.skp bytes) played back onto the reference SW canvas -> (H, W, 4) uint8 premul RGBA."""
    out = np.zeros((height, width, 4), dtype=np.uint8)
    sec = ctypes.c_double(0.0)
    m = (ctypes.c_float * 6)(*matrix6)
    rc = lib().ref_render_skp(skp, len(skp), int(width), int(height), m, out.ctypes.data, ctypes.byref(sec))
    if rc != 0:
        raise RuntimeError(f"ref_render_skp failed: {rc}")
    return (out, sec.value) if return_seconds else out


def raster_path(path, matrix6=(1, 0, 0, 0, 1, 0), clip=(-1e9, -1e9, 1e9, 1e9), cap=1 << 22):
    """SWRaster::RastePath on one PathData -> (spans int32[n,4] = x,y,len,cover ; bounds float32[4])."""
    rec = path.encode()
    m = np.asarray(matrix6, dtype=np.float32)
    c = np.asarray(clip, dtype=np.float32)
    spans = np.zeros((cap, 4), dtype=np.int32)
    bounds = np.zeros(4, dtype=np.float32)
    n = lib().ref_raster_path(rec, len(rec), m.ctypes.data, c.ctypes.data, spans.ctypes.data, cap,
                              bounds.ctypes.data)
    if n < 0:
        raise RuntimeError(f"ref_raster_path failed: {n}")
    if n > cap:
        return raster_path(path, matrix6, clip, cap=int(n))
    return spans[:n].copy(), bounds


def stack_blur(rgba, radius):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, _ = rgba.shape
    out = np.zeros_like(rgba)
    rc = lib().ref_stack_blur(rgba.ctypes.data, w, h, int(radius), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_stack_blur failed")
    return out


def coverage_aa_tile(path, matrix6=(1, 0, 0, 0, 1, 0), scissor=None, tile_cap=1 << 16, line_cap=1 << 20):
    """The reference's CoverageAAPathTiler + line encoder on one PathData -> (tiles (n, 5) int32: tile_x, tile_y, first
    line or -1, line count, backdrop; lines (m, 4) uint16 8.8 in encoded order)."""
    rec = path.encode()
    m = np.asarray(matrix6, dtype=np.float32)
    sc = None if scissor is None else np.asarray(scissor, dtype=np.float32)
    tiles = np.zeros((tile_cap, 5), dtype=np.int32)
    lines = np.zeros((line_cap, 4), dtype=np.uint16)
    nl = ctypes.c_long(0)
    n = lib().ref_coverage_aa_tile(rec, len(rec), m.ctypes.data, None if sc is None else sc.ctypes.data, tiles.ctypes.data,
                                   tile_cap, lines.ctypes.data, line_cap, ctypes.byref(nl))
    if n == -1:
        return coverage_aa_tile(path, matrix6, scissor, tile_cap * 8, line_cap * 8)
    if n < 0:
        raise RuntimeError(f"ref_coverage_aa_tile failed: {n}")
    return tiles[:n].copy(), lines[:nl.value].copy()
