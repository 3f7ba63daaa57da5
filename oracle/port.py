"""TEST INFRASTRUCTURE — ctypes binding of the plain-C oracle port (oracle/skb_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskb_oracle.so")
SRC_PATH = os.path.join(_HERE, "skb_oracle.c")
_lib = None

SEG_DTYPE = np.dtype([("type_flags", "<u4"), ("w", "<f4"), ("p", "<f4", (8,)), ("start", "<f4", (2,))])


def build(force=False):
    hdr = os.path.join(os.path.dirname(_HERE), "include", "skb_dl.h")
    area = os.path.join(_HERE, "skb_area_oracle.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(SRC_PATH), os.path.getmtime(hdr), os.path.getmtime(area))):
        return LIB_PATH
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-Wall", "-Wno-unused-function", "-o", LIB_PATH, SRC_PATH, "-lm"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.skbo_render.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        _lib.skbo_render.restype = ctypes.c_int
        _lib.skbo_raster_path.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        _lib.skbo_raster_path.restype = ctypes.c_long
        _lib.skbo_set_coord_mode.argtypes = [ctypes.c_int]
        _lib.skbo_set_row_band.argtypes = [ctypes.c_int, ctypes.c_int]
        _lib.skbo_set_coverage_mode.argtypes = [ctypes.c_int]
        _lib.skbo_area_tile_path.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long,
                                             ctypes.POINTER(ctypes.c_long)]
        _lib.skbo_area_tile_path.restype = ctypes.c_long
        _lib.skbo_area_raster_path.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        _lib.skbo_area_raster_path.restype = ctypes.c_long
        _lib.skbo_stack_blur.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


def dl_header(dl):
    names = ["magic", "version", "total_bytes", "flags", "n_surfaces", "n_ops", "n_paths", "n_segs", "n_paints",
             "n_stop_floats", "n_clip_states", "reserved0", "off_surfaces", "off_ops", "off_paths", "off_segs",
             "off_paints", "off_stops"]
    return dict(zip(names, struct.unpack_from("<18I", dl, 0)))


def set_coord_mode(mode):
    """0 auto (wide above 8192 px, the device's default), 1 the reference's int32 arithmetic, 2 wide."""
    lib().skbo_set_coord_mode(int(mode))


def set_row_band(y0, y1):
    """Render only the canvas rows [y0, y1) (y1 <= y0: all rows) — lets several processes share a very large frame."""
    lib().skbo_set_row_band(int(y0), int(y1))


def render_parallel(dl, procs=None):
    """port.render() of a large frame by `procs` forked processes, each producing a band of rows."""
    import multiprocessing as mp
    h = dl_header(dl)
    w, hh = struct.unpack_from("<2I", dl, h["off_surfaces"])
    procs = procs or os.cpu_count() or 1
    bands = [(hh * i // procs, hh * (i + 1) // procs) for i in range(procs)]
    out = np.zeros((hh, w, 4), dtype=np.uint8)
    global _pr_dl
    _pr_dl = dl
    with mp.get_context("fork").Pool(procs) as pool:
        for (y0, y1), rows in zip(bands, pool.imap(_render_band, bands)):
            out[y0:y1] = rows
    _pr_dl = None
    return out


_pr_dl = None


def _render_band(band):
    y0, y1 = band
    set_row_band(y0, y1)
    try:
        return render(_pr_dl)[y0:y1].copy()
    finally:
        set_row_band(0, 0)


def render(dl, initial=None):
    """Execute a display list on the CPU port -> (H, W, 4) uint8 premultiplied RGBA of surface 0."""
    h = dl_header(dl)
    w, hh = struct.unpack_from("<2I", dl, h["off_surfaces"])
    out = np.zeros((hh, w, 4), dtype=np.uint8)
    init_ptr = None
    if initial is not None:
        initial = np.ascontiguousarray(initial, dtype=np.uint8)
        init_ptr = initial.ctypes.data
    rc = lib().skbo_render(dl, len(dl), init_ptr, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"skbo_render failed: {rc}")
    return out


def raster_path(segs, ctm=(1, 0, 0, 0, 1, 0), clip=(-1e9, -1e9, 1e9, 1e9), even_odd=False, cap=1 << 22):
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    m = np.asarray(ctm, dtype=np.float32)
    c = np.asarray(clip, dtype=np.float32)
    spans = np.zeros((cap, 4), dtype=np.int32)
    bounds = np.zeros(4, dtype=np.float32)
    n = lib().skbo_raster_path(segs.ctypes.data, len(segs), m.ctypes.data, c.ctypes.data, int(even_odd),
                               spans.ctypes.data, cap, bounds.ctypes.data)
    if n > cap:
        return raster_path(segs, ctm, clip, even_odd, int(n))
    return spans[:n].copy(), bounds


def set_coverage_mode(mode):
    """0 = the software backend's analytic AA (default); 1 = the reference's coverage-AA path for unclipped fills
    (what the CUDA backend computes under SKB_COVERAGE_AREA)."""
    lib().skbo_set_coverage_mode(int(mode))


def render_area(dl, initial=None):
    """render() with coverage-AA ("AREA") coverage."""
    set_coverage_mode(1)
    try:
        return render(dl, initial)
    finally:
        set_coverage_mode(0)


def area_tile_path(segs, ctm=(1, 0, 0, 0, 1, 0), scissor=None, even_odd=False, tile_cap=1 << 16, line_cap=1 << 20):
    """The port's CoverageAAPathTiler restatement on one path -> (tiles (n, 5) int32: tile_x, tile_y, first line or -1,
    line count, backdrop; lines (m, 4) uint16 in range order)."""
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    m = np.asarray(ctm, dtype=np.float32)
    sc = None if scissor is None else np.asarray(scissor, dtype=np.float32)
    tiles = np.zeros((tile_cap, 5), dtype=np.int32)
    lines = np.zeros((line_cap, 4), dtype=np.uint16)
    nl = ctypes.c_long(0)
    n = lib().skbo_area_tile_path(segs.ctypes.data, len(segs), m.ctypes.data, None if sc is None else sc.ctypes.data,
                                  int(even_odd), tiles.ctypes.data, tile_cap, lines.ctypes.data, line_cap, ctypes.byref(nl))
    if n < 0:
        return area_tile_path(segs, ctm, scissor, even_odd, tile_cap * 8, line_cap * 8)
    return tiles[:n].copy(), lines[:nl.value].copy()


def area_coverage(segs, w, h, ctm=(1, 0, 0, 0, 1, 0), clip=None, even_odd=False, cap=1 << 22):
    """Coverage-AA ("AREA") coverage of one path on a w x h surface -> (h, w) uint8."""
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    m = np.asarray(ctm, dtype=np.float32)
    c = np.asarray(clip if clip is not None else (0, 0, w, h), dtype=np.float32)
    spans = np.zeros((cap, 4), dtype=np.int32)
    n = lib().skbo_area_raster_path(segs.ctypes.data, len(segs), m.ctypes.data, c.ctypes.data, int(even_odd), int(w), int(h),
                                    spans.ctypes.data, cap)
    if n > cap:
        return area_coverage(segs, w, h, ctm, clip, even_odd, int(n))
    out = np.zeros((h, w), dtype=np.uint8)
    for x, y, ln, cv in spans[:n]:
        out[y, x:x + ln] = cv
    return out


def stack_blur(rgba, radius):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, _ = rgba.shape
    out = np.zeros_like(rgba)
    lib().skbo_stack_blur(rgba.ctypes.data, out.ctypes.data, w, h, int(radius))
    return out


def dl_segments(dl, path_index):
    """Segments of one path of a display list as a SEG_DTYPE array."""
    h = dl_header(dl)
    seg_off, n_segs = struct.unpack_from("<2I", dl, h["off_paths"] + 16 * path_index)
    return np.frombuffer(dl, dtype=SEG_DTYPE, count=n_segs, offset=h["off_segs"] + 48 * seg_off).copy()
