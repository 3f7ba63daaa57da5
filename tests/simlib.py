"""Builds and binds tests/sim/sim_core.cc — the CPU simulation of the CUDA stages' per-thread code."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "sim", "sim_core.cc")
_LIB = os.path.join(_HERE, "sim", "libskb_sim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [_SRC] + [os.path.join(_REPO, "skity_b200", "csrc", f) for f in
                         ("skb_core.cuh", "skb_walk.cuh", "skb_stages.cuh", "skb_clip.cuh", "skb_rowwalk.cuh", "skb_area.cuh")] + [os.path.join(_HERE, "sim", "walk_nested.hpp")] + [os.path.join(_REPO, "include", "skb_dl.h")]
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                                   f"-I{_REPO}", _SRC, "-o", _LIB])
        _lib = ctypes.CDLL(_LIB)
        _lib.sim_path_cover.restype = ctypes.c_long
        _lib.sim_path_cover.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.sim_rowwalk_check.restype = ctypes.c_int
        _lib.sim_rowwalk_check.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.sim_area_cover.restype = ctypes.c_int
        _lib.sim_area_cover.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.sim_render_dl.restype = ctypes.c_int
        _lib.sim_render_dl.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def render_dl(dl):
    """Whole-frame CPU simulation of the device stages (fills, gradients, path clips; no blur)."""
    import struct
    off_surfaces = struct.unpack_from("<18I", dl, 0)[12]
    w, h = struct.unpack_from("<2I", dl, off_surfaces)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.int64)
    rc = lib().sim_render_dl(dl, len(dl), out.ctypes.data, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"sim_render_dl failed: {rc}")
    return out, stats


def path_cover(segs, ctm, clip, even_odd, w, h):
    segs = np.ascontiguousarray(segs)
    m = np.asarray(ctm, dtype=np.float32)
    c = np.asarray(clip, dtype=np.float32)
    d = np.zeros((h, w), dtype=np.uint8)
    a = np.zeros((h, w), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.int64)
    n = lib().sim_path_cover(segs.ctypes.data, len(segs), m.ctypes.data, c.ctypes.data, int(even_odd), w, h,
                             d.ctypes.data, a.ctypes.data, stats.ctypes.data)
    return d, a, int(n), stats


def rowwalk_check(segs, ctm, clip, even_odd, w, h):
    """Row-parallel walk vs the sequential sweep on one path -> (status, stats): 0 identical records, 1 flagged for the
    sequential fallback, 2 MISMATCH, 3 empty path."""
    segs = np.ascontiguousarray(segs)
    m = np.asarray(ctm, dtype=np.float32)
    c = np.asarray(clip, dtype=np.float32)
    stats = np.zeros(4, dtype=np.int64)
    rc = lib().sim_rowwalk_check(segs.ctypes.data, len(segs), m.ctypes.data, c.ctypes.data, int(even_odd), w, h, stats.ctypes.data)
    return int(rc), stats


def area_cover(segs, ctm, clip, even_odd, w, h):
    """Coverage mode AREA (skb_area.cuh) of one path, thread by thread on the CPU -> ((h, w) uint8 coverage, stats)."""
    segs = np.ascontiguousarray(segs)
    m = np.asarray(ctm, dtype=np.float32)
    c = np.asarray(clip, dtype=np.float32)
    out = np.zeros((h, w), dtype=np.uint8)
    stats = np.zeros(2, dtype=np.int64)
    rc = lib().sim_area_cover(segs.ctypes.data, len(segs), m.ctypes.data, c.ctypes.data, int(even_odd), w, h, out.ctypes.data,
                              stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"sim_area_cover failed: {rc}")
    return out, stats
