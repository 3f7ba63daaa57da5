"""ctypes binding of the CUDA backend's C ABI (include/skb.h, skity_b200/lib/libskb.so).

Fails loudly when the library or a B200-class device is missing: there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKB_LIB", os.path.join(_PKG, "lib", "libskb.so"))
_lib = None


class FrameStats(ctypes.Structure):
    _fields_ = [("n_ops", ctypes.c_uint32), ("n_segs", ctypes.c_uint32), ("n_prims", ctypes.c_uint32),
                ("n_edges_slots", ctypes.c_uint32),
                ("n_rows", ctypes.c_uint64), ("n_records", ctypes.c_uint64), ("n_items", ctypes.c_uint64),
                ("n_items_nonempty", ctypes.c_uint64), ("n_cmds", ctypes.c_uint64), ("n_tiles", ctypes.c_uint64),
                ("pool_capacity", ctypes.c_uint64),
                ("n_launches", ctypes.c_uint32), ("n_retries", ctypes.c_uint32),
                ("ms_total", ctypes.c_float), ("ms_stage", ctypes.c_float * 8),
                ("bytes_fine", ctypes.c_uint64), ("bytes_cover", ctypes.c_uint64), ("bytes_walk", ctypes.c_uint64),
                ("bytes_blur", ctypes.c_uint64), ("n_rw_retried", ctypes.c_uint32), ("n_rw_sequential", ctypes.c_uint32),
                ("n_area_lines", ctypes.c_uint32), ("n_area_tile_lines", ctypes.c_uint32), ("bytes_area", ctypes.c_uint64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "ms_stage" else v
        return d


STAGE_NAMES = ["flatten", "setup", "walk", "coverage", "bin", "fine", "blur", "clip"]


class SkbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkbError(f"{LIB_PATH} is missing: build it with `python -m skity_b200.build cuda` "
                           "(nvcc, sm_100a) — there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        vp, u32, i32, sz = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_size_t
        L.skb_get_last_error_string.restype = ctypes.c_char_p
        L.skb_version_string.restype = ctypes.c_char_p
        L.skb_device_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        L.skb_device_destroy.argtypes = [vp]
        L.skb_device_destroy.restype = None
        L.skb_device_sm_count.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
        L.skb_surface_create.argtypes = [vp, u32, u32, ctypes.POINTER(vp)]
        L.skb_surface_destroy.argtypes = [vp]
        L.skb_surface_destroy.restype = None
        L.skb_surface_set_band.argtypes = [vp, u32, u32]
        L.skb_surface_set_coord_mode.argtypes = [vp, ctypes.c_int]
        L.skb_surface_set_walk_mode.argtypes = [vp, ctypes.c_int]
        L.skb_surface_set_coverage_mode.argtypes = [vp, ctypes.c_int]
        L.skb_frame_begin.argtypes = [vp, ctypes.c_int]
        L.skb_frame_encode.argtypes = [vp, vp, sz]
        L.skb_frame_flush.argtypes = [vp]
        L.skb_display_list_validate.argtypes = [vp, sz]
        L.skb_display_list_cull_rows.argtypes = [vp, sz, i32, i32, vp, sz, ctypes.POINTER(sz)]
        L.skb_surface_sync.argtypes = [vp]
        L.skb_surface_read_pixels.argtypes = [vp, u32, u32, u32, u32, vp, sz]
        L.skb_surface_read_pixels_async.argtypes = [vp, u32, u32, u32, u32, vp, sz]
        L.skb_surface_write_pixels.argtypes = [vp, u32, u32, u32, u32, vp, sz]
        L.skb_frame_read_surface.argtypes = [vp, u32, vp, sz]
        L.skb_surface_device_ptr.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(sz)]
        L.skb_surface_export_canvas.argtypes = [vp, ctypes.c_char_p]
        L.skb_surface_set_remote_canvas.argtypes = [vp, ctypes.c_char_p]
        L.skb_surface_stream.argtypes = [vp, ctypes.POINTER(vp)]
        L.skb_frame_get_stats.argtypes = [vp, ctypes.POINTER(FrameStats)]
        L.skb_debug_read_coverage.argtypes = [vp, u32, i32, i32, u32, u32, vp, vp]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = lib().skb_get_last_error_string().decode(errors="replace")
        raise SkbError(f"{what} failed ({rc}): {msg}")


class Device:
    def __init__(self, ordinal=0):
        self._h = ctypes.c_void_p()
        _check(lib().skb_device_create(ordinal, ctypes.byref(self._h)), "skb_device_create")
        self.ordinal = ordinal

    @property
    def sm_count(self):
        n = ctypes.c_int()
        _check(lib().skb_device_sm_count(self._h, ctypes.byref(n)), "skb_device_sm_count")
        return n.value

    def create_surface(self, width, height):
        return Surface(self, width, height)

    def close(self):
        if self._h:
            lib().skb_device_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Surface:
    def __init__(self, device, width, height):
        self.device = device
        self.width, self.height = int(width), int(height)
        self._h = ctypes.c_void_p()
        _check(lib().skb_surface_create(device._h, self.width, self.height, ctypes.byref(self._h)), "skb_surface_create")

    def set_band(self, y0, y1):
        _check(lib().skb_surface_set_band(self._h, y0, y1), "skb_surface_set_band")

    def set_coord_mode(self, mode):
        """0 auto (wide above 8192 px), 1 the reference's int32 arithmetic (wraps at 8192 px), 2 wide."""
        _check(lib().skb_surface_set_coord_mode(self._h, int(mode)), "skb_surface_set_coord_mode")

    def set_walk_mode(self, mode):
        """0 the sequential sweep (one thread per path), 1 the row-parallel sweep (same records)."""
        _check(lib().skb_surface_set_walk_mode(self._h, int(mode)), "skb_surface_set_walk_mode")

    def set_coverage_mode(self, mode):
        """0 exact (the software backend's analytic AA, default), 1 AREA (tile-binned signed-area coverage)."""
        _check(lib().skb_surface_set_coverage_mode(self._h, int(mode)), "skb_surface_set_coverage_mode")

    def begin(self, clear=True):
        _check(lib().skb_frame_begin(self._h, 1 if clear else 0), "skb_frame_begin")

    def encode(self, display_list):
        """display_list: bytes, or (address, nbytes) of pinned host memory."""
        if isinstance(display_list, tuple):
            addr, n = display_list
            _check(lib().skb_frame_encode(self._h, ctypes.c_void_p(addr), n), "skb_frame_encode")
        else:
            buf = (ctypes.c_char * len(display_list)).from_buffer_copy(display_list)
            _check(lib().skb_frame_encode(self._h, ctypes.cast(buf, ctypes.c_void_p), len(display_list)),
                   "skb_frame_encode")

    def flush(self):
        _check(lib().skb_frame_flush(self._h), "skb_frame_flush")

    def sync(self):
        _check(lib().skb_surface_sync(self._h), "skb_surface_sync")

    def read_pixels(self, x=0, y=0, width=None, height=None, out=None):
        w = self.width - x if width is None else width
        h = self.height - y if height is None else height
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        _check(lib().skb_surface_read_pixels(self._h, x, y, w, h, out.ctypes.data, w * 4), "skb_surface_read_pixels")
        return out

    def read_pixels_async(self, out, x=0, y=0):
        """Enqueue the read-back on the surface's stream; `out` (h, w, 4 uint8, ideally pinned) is valid
        after the next sync()."""
        h, w, _ = out.shape
        _check(lib().skb_surface_read_pixels_async(self._h, x, y, w, h, out.ctypes.data, w * 4), "skb_surface_read_pixels_async")

    def write_pixels(self, rgba, x=0, y=0):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w, _ = rgba.shape
        _check(lib().skb_surface_write_pixels(self._h, x, y, w, h, rgba.ctypes.data, w * 4), "skb_surface_write_pixels")

    def read_batch_canvas(self, index, width, height, out=None):
        """Read back canvas `index` (1-based) of the last flushed batch display list."""
        if out is None:
            out = np.empty((height, width, 4), dtype=np.uint8)
        _check(lib().skb_frame_read_surface(self._h, index, out.ctypes.data, width * 4), "skb_frame_read_surface")
        return out

    def export_canvas(self):
        """64-byte CUDA IPC handle of the canvas, for the other GPUs' fine passes to store their bands into."""
        buf = ctypes.create_string_buffer(64)
        _check(lib().skb_surface_export_canvas(self._h, buf), "skb_surface_export_canvas")
        return buf.raw

    def set_remote_canvas(self, handle):
        """Store this surface's finished band into the canvas `handle` names (None: back to local stores)."""
        _check(lib().skb_surface_set_remote_canvas(self._h, handle), "skb_surface_set_remote_canvas")

    def device_ptr(self):
        p, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        _check(lib().skb_surface_device_ptr(self._h, ctypes.byref(p), ctypes.byref(pitch)), "skb_surface_device_ptr")
        return p.value, pitch.value

    def stream(self):
        p = ctypes.c_void_p()
        _check(lib().skb_surface_stream(self._h, ctypes.byref(p)), "skb_surface_stream")
        return p.value or 0

    def stats(self):
        st = FrameStats()
        _check(lib().skb_frame_get_stats(self._h, ctypes.byref(st)), "skb_frame_get_stats")
        return st.as_dict()

    def read_coverage(self, op_index, x, y, width, height):
        d = np.zeros((height, width), dtype=np.uint8)
        a = np.zeros((height, width), dtype=np.uint8)
        _check(lib().skb_debug_read_coverage(self._h, op_index, x, y, width, height, d.ctypes.data, a.ctypes.data),
               "skb_debug_read_coverage")
        return d, a

    def render(self, display_list, clear=True):
        """One frame: begin, encode (H2D), flush, read back -> (H, W, 4) uint8 premultiplied RGBA."""
        self.begin(clear)
        self.encode(display_list)
        self.flush()
        return self.read_pixels()

    def close(self):
        if self._h:
            lib().skb_surface_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def validate_display_list(dl):
    """skb_display_list_validate: raises SkbError with the library's reason when the list is malformed (no GPU needed)."""
    buf = bytes(dl)
    rc = lib().skb_display_list_validate(buf, len(buf))
    if rc != 0:
        raise SkbError(f"skb_display_list_validate failed ({rc}): {lib().skb_get_last_error_string().decode(errors='replace')}")


def cull_display_list_rows(dl, row0, row1, out=None, size_only=False):
    """skb_display_list_cull_rows: the part of display list `dl` (bytes, or a (pointer, size) pair) that can reach canvas
    rows [row0, row1).  Returns bytes — or, with `out` = (pointer, capacity), the number of bytes written there; with
    `size_only` just the size the result needs."""
    if isinstance(dl, tuple):
        src, n = ctypes.c_void_p(dl[0]), dl[1]
    else:
        buf = bytes(dl)
        src, n = buf, len(buf)
    need = ctypes.c_size_t(0)
    L = lib()

    def check(rc):
        if rc != 0:
            raise SkbError(f"skb_display_list_cull_rows failed ({rc}): {L.skb_get_last_error_string().decode(errors='replace')}")
    if out is not None:
        check(L.skb_display_list_cull_rows(src, n, int(row0), int(row1), ctypes.c_void_p(out[0]), out[1], ctypes.byref(need)))
        return need.value
    check(L.skb_display_list_cull_rows(src, n, int(row0), int(row1), None, 0, ctypes.byref(need)))
    if size_only:
        return need.value
    res = ctypes.create_string_buffer(need.value)
    check(L.skb_display_list_cull_rows(src, n, int(row0), int(row1), res, need.value, ctypes.byref(need)))
    return res.raw[:need.value]
