"""TEST / BASELINE INFRASTRUCTURE — the reference's software backend on canvases beyond its numeric range.

SWFDot6ToFixed is `x << 10` in int32 (src/render/sw/sw_subpixel.hpp:43): coordinates >= 8192 px wrap, so the compiled
reference (oracle/_ref) cannot render the 16384^2 config directly.  A random-fills scene is therefore cut into windows
of 4096^2, each rendered under Translate(-Tx, -Ty) with only the paths that touch it; T = 0 for the first window of an
axis and 4096*i - 2048 otherwise, so that every coordinate of a window's paths lies in [0, 8192) (no wrap, no sign
change of the float -> fixed truncation).  Used by bench.py's CPU arms (timing) and by
tests/golden/make_config_digests.py (a second opinion beside the wide-mode port; the two are not bit-identical, see
there).  Only tests/, bench.py's CPU legs and the golden generators may import this.
"""
import struct

import numpy as np

WINDOW = 4096
MARGIN = 2048
_MAGIC = 0x43534B53
_OP_TRANSLATE = 3


def _gather_f32(body, starts, npts):
    k = np.arange(4 * npts)
    raw = body[starts[:, None] + k[None, :]]
    return np.ascontiguousarray(raw).view(np.float32).reshape(len(starts), npts)


def fills_records(blob):
    """(offset, length, bbox) of every DrawPath record of a scene.scene_random_fills_fast blob."""
    n_ops = struct.unpack_from("<6I", blob, 0)[4]
    body = np.frombuffer(blob, np.uint8, offset=24)
    off = np.zeros(n_ops, np.int64)
    ln = np.where(np.arange(n_ops) % 2 == 0, 168, 200)
    off[1:] = np.cumsum(ln)[:-1]
    bbox = np.zeros((n_ops, 4), np.float32)
    for parity, npts in ((0, 18), (1, 26)):
        idx = np.arange(parity, n_ops, 2)
        if len(idx) == 0:
            continue
        pts = _gather_f32(body, off[idx] + 32, npts)
        xs, ys = pts[:, 0::2], pts[:, 1::2]
        bbox[idx] = np.stack([xs.min(1), ys.min(1), xs.max(1), ys.max(1)], axis=1)
    return off + 24, ln, bbox


def window_origin(i):
    """(global start of window i, translation T, start of the window in the translated frame)"""
    g0 = WINDOW * i
    t = 0 if i == 0 else g0 - MARGIN
    return g0, t, g0 - t


def window_scene(blob, records, ix, iy, size):
    """-> (sub-scene blob, (crop_x, crop_y) of the window inside its frame, number of paths)"""
    off, ln, bbox = records
    gx0, tx, lx0 = window_origin(ix)
    gy0, ty, ly0 = window_origin(iy)
    gx1, gy1 = min(gx0 + WINDOW, size), min(gy0 + WINDOW, size)
    keep = (bbox[:, 2] >= gx0 - 2) & (bbox[:, 0] <= gx1 + 2) & (bbox[:, 3] >= gy0 - 2) & (bbox[:, 1] <= gy1 + 2)
    idx = np.nonzero(keep)[0]
    w, h = lx0 + (gx1 - gx0), ly0 + (gy1 - gy0)
    ops = [struct.pack("<2I2f", _OP_TRANSLATE, 8, float(-tx), float(-ty))]
    ops += [blob[off[i]:off[i] + ln[i]] for i in idx]
    sub = struct.pack("<6I", _MAGIC, 1, w, h, len(ops), 0) + b"".join(ops)
    return sub, (lx0, ly0), len(idx)


def n_windows(size):
    return (size + WINDOW - 1) // WINDOW
