"""ctypes binding of the host plug-in (skity_b200/lib/libskb_skity.so): scene blob -> display list."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libskb_skity.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m skity_b200.build host` "
                               "(needs the reference tree) — there is no fallback encoder")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.skbh_encode_scene.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p),
                                           ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
        _lib.skbh_encode_scene.restype = ctypes.c_int
        _lib.skbh_encode_recorded_scene.argtypes = _lib.skbh_encode_scene.argtypes
        _lib.skbh_encode_recorded_scene.restype = ctypes.c_int
        _lib.skbh_free.argtypes = [ctypes.c_void_p]
        _lib.skbh_encode_scene_batch.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_uint32,
                                                 ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                                 ctypes.c_char_p, ctypes.c_size_t]
        _lib.skbh_encode_scene_batch.restype = ctypes.c_int
        _lib.skbh_encode_skp.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]
        _lib.skbh_encode_skp.restype = ctypes.c_int
        _lib.skbh_render_scene_cuda.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_char_p, ctypes.c_size_t]
        _lib.skbh_render_scene_cuda.restype = ctypes.c_int
        _lib.skbh_render_scene_cuda_frames.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                       ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_size_t]
        _lib.skbh_render_scene_cuda_frames.restype = ctypes.c_int
    return _lib


def encode_scene(blob, allow_unsupported=False, recorded=False):
    """Replay an SKSC scene through CudaCanvas and return the SKDL display list bytes.  With `recorded` the
    scene is first recorded into a skity::DisplayList (PictureRecorder) and that picture is replayed."""
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    msg = ctypes.create_string_buffer(256)
    fn = lib().skbh_encode_recorded_scene if recorded else lib().skbh_encode_scene
    rc = fn(blob, len(blob), ctypes.byref(out), ctypes.byref(n), msg, 256)
    if rc != 0:
        raise RuntimeError(f"skbh_encode_scene failed: {rc}")
    try:
        data = ctypes.string_at(out, n.value)
    finally:
        lib().skbh_free(out)
    if msg.value and not allow_unsupported:
        raise RuntimeError(f"scene uses a feature outside the CUDA backend's scope: {msg.value.decode()}")
    return data


def encode_skp(skp, width, height, matrix6=(1, 0, 0, 0, 1, 0), allow_unsupported=False):
    """A serialized picture (.skp bytes, the reference's module/io) played back through CudaCanvas onto a width x height
    canvas under the affine matrix (sx kx tx ky sy ty) -> SKDL display list bytes."""
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    msg = ctypes.create_string_buffer(256)
    m = (ctypes.c_float * 6)(*matrix6)
    rc = lib().skbh_encode_skp(skp, len(skp), int(width), int(height), m, ctypes.byref(out), ctypes.byref(n), msg, 256)
    if rc != 0:
        raise RuntimeError(f"skbh_encode_skp failed: {rc}")
    try:
        data = ctypes.string_at(out, n.value)
    finally:
        lib().skbh_free(out)
    if msg.value and not allow_unsupported:
        raise RuntimeError(f"picture uses a feature outside the CUDA backend's scope: {msg.value.decode()}")
    return data


def encode_scene_batch(blobs, allow_unsupported=False):
    """Replay a batch of independent scenes into ONE display list.  Returns (display list bytes, canvas surface
    index of every scene) — blur temporaries take surface indices in between."""
    joined = b"".join(blobs)
    sizes = (ctypes.c_size_t * len(blobs))(*[len(b) for b in blobs])
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    msg = ctypes.create_string_buffer(256)
    ids = (ctypes.c_uint32 * len(blobs))()
    rc = lib().skbh_encode_scene_batch(joined, sizes, len(blobs), ids, ctypes.byref(out), ctypes.byref(n), msg, 256)
    if rc != 0:
        raise RuntimeError(f"skbh_encode_scene_batch failed: {rc}")
    try:
        data = ctypes.string_at(out, n.value)
    finally:
        lib().skbh_free(out)
    if msg.value and not allow_unsupported:
        raise RuntimeError(f"scene uses a feature outside the CUDA backend's scope: {msg.value.decode()}")
    return data, list(ids)


def render_scene_cuda(blob, device_ordinal=0):
    """Replay an SKSC scene through the complete skity plug-in path on the GPU:
    CudaContextCreate -> GPUContext::CreateSurface -> LockCanvas -> Canvas calls -> Flush -> ReadPixels."""
    import struct

    import numpy as np
    _, _, w, h, _, _ = struct.unpack_from("<6I", blob, 0)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    msg = ctypes.create_string_buffer(512)
    rc = lib().skbh_render_scene_cuda(blob, len(blob), device_ordinal, out.ctypes.data, msg, 512)
    if rc != 0:
        raise RuntimeError(f"skbh_render_scene_cuda failed ({rc}): {msg.value.decode(errors='replace')}")
    return out


def render_scene_cuda_frames(blob, frames, device_ordinal=0, want_pixels=False):
    """`frames` frames of the scene through the skity plug-in path on ONE context and surface
    (LockCanvas -> Canvas calls -> Flush -> ReadPixels per frame).  -> (mean ms per frame of the frames after the first:
    dict(total, canvas_calls, flush, read_pixels), pixels of the last frame or None)."""
    import struct

    import numpy as np
    _, _, w, h, _, _ = struct.unpack_from("<6I", blob, 0)
    out = np.zeros((h, w, 4), dtype=np.uint8) if want_pixels else None
    ms = (ctypes.c_double * 4)()
    msg = ctypes.create_string_buffer(512)
    rc = lib().skbh_render_scene_cuda_frames(blob, len(blob), device_ordinal, int(frames), out.ctypes.data if want_pixels else None,
                                             ms, msg, 512)
    if rc != 0:
        raise RuntimeError(f"skbh_render_scene_cuda_frames failed ({rc}): {msg.value.decode(errors='replace')}")
    return dict(total=ms[0], canvas_calls=ms[1], flush=ms[2], read_pixels=ms[3]), out
