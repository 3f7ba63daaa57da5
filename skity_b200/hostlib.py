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
        _lib.skbh_free.argtypes = [ctypes.c_void_p]
    return _lib


def encode_scene(blob, allow_unsupported=False):
    """Replay an SKSC scene through CudaCanvas and return the SKDL display list bytes."""
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    msg = ctypes.create_string_buffer(256)
    rc = lib().skbh_encode_scene(blob, len(blob), ctypes.byref(out), ctypes.byref(n), msg, 256)
    if rc != 0:
        raise RuntimeError(f"skbh_encode_scene failed: {rc}")
    try:
        data = ctypes.string_at(out, n.value)
    finally:
        lib().skbh_free(out)
    if msg.value and not allow_unsupported:
        raise RuntimeError(f"scene uses a feature outside the CUDA backend's scope: {msg.value.decode()}")
    return data
