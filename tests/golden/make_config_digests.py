#!/usr/bin/env python3
"""Digests of the BASELINE.json configs at their NAMED sizes, from the reference's own software
backend (oracle/_ref/libskity_ref.so, the unmodified sources compiled by oracle/build_ref.py).

    python tests/golden/make_config_digests.py [c2 c3 c4b c4a]      (build container; minutes of CPU)

The full frames are far too large to commit (C4a is 1 GiB), so each config stores what
tests/config_digest.py computes from a frame: SHA-256 of the bytes, their sum, and two checksums per
64x64-pixel block (sum of the bytes, and a position-weighted sum) — enough to say WHERE a frame
differs.  tests/test_gpu_parity.py renders the same scenes through the C ABI and compares.

C4a (1M paths, 16384^2) lies outside the reference's numeric range: SWFDot6ToFixed is `x << 10` in
int32 (src/render/sw/sw_subpixel.hpp:43), so coordinates >= 8192 px wrap and the compiled reference
cannot render the scene.  Its digest therefore comes from the pinned C port (oracle/skb_oracle.c) in
wide-coordinate mode — the port's one deviation from the reference, the same 24.8 -> 16.16 conversion
without the overflow, which is what include/skb.h's SKB_COORD_WIDE computes on the device.

Beside it the compiled reference renders the scene as 4x4 windows of 4096^2 under Translate(-Tx, -Ty)
(T = 0 for the first window of an axis, 4096*i - 2048 otherwise, so that every coordinate of a window's
paths lies in [0, 8192): no wrap, no sign change).  Whole-pixel translation is exact in fp32 at these
magnitudes and commutes with the quarter-pixel y snapping, but the reference's result is still not
translation invariant: ChopQuadAtYExtrema (src/geometry/geometry.cc:323-349) interpolates the already
transformed control points in fp32, so a chop point lands one ulp elsewhere under another translation, which now
and then moves a coordinate across a 1/256-pixel truncation or a y across a quarter-pixel snap.  The windowed
frame is thus a second opinion with a tolerance, not the bit-exact target; how far the two agree is stored
as c4a.windowed_agreement (pixels equal, within 1/255, within 2/255, maximum difference).
"""
import multiprocessing as mp
import os
import struct
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import config_digest  # noqa: E402
from skity_b200 import scene  # noqa: E402

from oracle import windows  # noqa: E402
WINDOW = windows.WINDOW


def _render(blob):
    from oracle import refsw
    return refsw.render_scene(blob)


def _window_job(args):
    blob, records, ix, iy, size = args
    sub, (lx0, ly0), cnt = windows.window_scene(blob, records, ix, iy, size)
    img = _render(sub)
    return ix, iy, img[ly0:, lx0:].copy(), cnt


def reference_windowed(blob, size, procs):
    """The reference's frame of a fills scene on a canvas larger than its 8192-px numeric range."""
    records = windows.fills_records(blob)
    n = windows.n_windows(size)
    out = np.zeros((size, size, 4), np.uint8)
    jobs = [(blob, records, ix, iy, size) for iy in range(n) for ix in range(n)]
    with mp.get_context("fork").Pool(procs) as pool:
        for ix, iy, img, cnt in pool.imap_unordered(_window_job, jobs):
            out[iy * WINDOW:iy * WINDOW + img.shape[0], ix * WINDOW:ix * WINDOW + img.shape[1]] = img
            print(f"  window ({ix},{iy}): {cnt} paths", flush=True)
    return out


def config_scene(name):
    if name == "c2":
        return scene.scene_c2(20000, 4096, 2)
    if name == "c3":
        return scene.scene_c3(2000, 8192, 3)
    if name == "c4a":
        return scene.scene_c4a()
    if name.startswith("c4b_"):
        return scene.scene_c4b(int(name[4:]))
    raise KeyError(name)


C4B_CANVASES = 64


def main():
    from oracle import refsw
    assert refsw.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    which = sys.argv[1:] or ["c3", "c4b", "c2", "c4a"]
    procs = os.cpu_count() or 1
    path = os.path.join(HERE, "config_digests.npz")
    store = dict(np.load(path)) if os.path.exists(path) else {}
    for name in which:
        t0 = time.time()
        if name == "c4a":
            from oracle import port
            from skity_b200 import hostlib
            s = config_scene("c4a")
            blob = s.encode()
            port.set_coord_mode(0)   # auto: wide above 8192 px, the device's default
            img = port.render_parallel(hostlib.encode_scene(blob), procs)
            config_digest.put(store, "c4a", img)
            print(f"  port (wide): {time.time() - t0:.1f} s", flush=True)
            win = reference_windowed(blob, s.width, procs)
            agree = np.zeros(5, np.float64)   # equal, <= 1, <= 2 (fractions of pixels), max difference, pixels
            n_px = 0
            for y in range(0, s.height, 1024):
                d = np.abs(img[y:y + 1024].astype(np.int16) - win[y:y + 1024].astype(np.int16)).max(axis=2)
                agree[0] += (d == 0).sum()
                agree[1] += (d <= 1).sum()
                agree[2] += (d <= 2).sum()
                agree[3] = max(agree[3], d.max())
                n_px += d.size
            agree[:3] /= n_px
            agree[4] = n_px
            store["c4a.windowed_agreement"] = agree
            print(f"  vs windowed reference: equal {agree[0]:.6f}, <=1 {agree[1]:.6f}, <=2 {agree[2]:.6f}, max {int(agree[3])}", flush=True)
        elif name == "c4b":
            with mp.get_context("fork").Pool(procs) as pool:
                imgs = pool.map(_render, [config_scene(f"c4b_{i}").encode() for i in range(C4B_CANVASES)])
            for i, img in enumerate(imgs):
                config_digest.put(store, f"c4b_{i}", img)
        else:
            img = _render(config_scene(name).encode())
            config_digest.put(store, name, img)
        print(f"{name}: {time.time() - t0:.1f} s", flush=True)
        np.savez_compressed(path, **store)
    print(f"{path}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
