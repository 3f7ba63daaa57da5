"""Digest of a rendered frame: SHA-256, byte sum and two checksums per 64x64-pixel block.

Used for the BASELINE.json configs at their named sizes, whose reference frames (up to 1 GiB) cannot be
committed: tests/golden/make_config_digests.py stores the digests of the reference's frames in
tests/golden/config_digests.npz, the GPU parity tests compute the same digest of the CUDA frame and
compare — equal digests mean equal frames, and the block grids say where two frames differ.
"""
import hashlib

import numpy as np

BLOCK = 64


def block_grids(img):
    """img (H, W, 4) uint8 -> (sums uint32[by, bx], weighted uint32[by, bx])."""
    h, w, c = img.shape
    by, bx = (h + BLOCK - 1) // BLOCK, (w + BLOCK - 1) // BLOCK
    sums = np.zeros((by, bx), np.uint32)
    wsum = np.zeros((by, bx), np.uint32)
    # weights depend on the position inside the block, so moved or swapped pixels change the second grid
    wy = (np.arange(BLOCK, dtype=np.uint32) * 2654435761 + 12345) & 0xFFFF
    wx = (np.arange(BLOCK * 4, dtype=np.uint32) * 40503 + 77) & 0xFFFF
    for j in range(by):
        rows = img[j * BLOCK:(j + 1) * BLOCK]
        rh = rows.shape[0]
        flat = rows.reshape(rh, w * c).astype(np.uint32)
        pad = bx * BLOCK * c - w * c
        if pad:
            flat = np.pad(flat, ((0, 0), (0, pad)))
        blk = flat.reshape(rh, bx, BLOCK * c)
        sums[j] = blk.sum(axis=(0, 2), dtype=np.uint32)
        t = (blk * wx[None, None, :]).sum(axis=2, dtype=np.uint32)          # (rh, bx)
        wsum[j] = (t * wy[:rh, None]).sum(axis=0, dtype=np.uint32)
    return sums, wsum


def digest(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    sums, wsum = block_grids(img)
    sha = hashlib.sha256(img.tobytes() if img.nbytes < (1 << 28) else memoryview(img.reshape(-1))).digest()
    return {"sha256": np.frombuffer(sha, np.uint8).copy(), "byte_sum": np.array([int(sums.sum(dtype=np.uint64))], np.uint64),
            "shape": np.array(img.shape[:2], np.uint32), "block_sum": sums, "block_wsum": wsum}


def put(store, name, img):
    for k, v in digest(img).items():
        store[f"{name}.{k}"] = v


def compare(store, name, img):
    """-> (equal, description of the differences)"""
    d = digest(img)
    if tuple(store[f"{name}.shape"]) != tuple(d["shape"]):
        return False, f"{name}: shape {tuple(d['shape'])} != {tuple(store[name + '.shape'])}"
    if np.array_equal(store[f"{name}.sha256"], d["sha256"]):
        return True, ""
    bad = (store[f"{name}.block_sum"] != d["block_sum"]) | (store[f"{name}.block_wsum"] != d["block_wsum"])
    ys, xs = np.nonzero(bad)
    where = ", ".join(f"({x * BLOCK},{y * BLOCK})" for y, x in list(zip(ys, xs))[:8])
    return False, (f"{name}: frame differs from the reference's: {int(bad.sum())} of {bad.size} 64x64 blocks, byte sum "
                   f"{int(d['byte_sum'][0])} vs {int(store[name + '.byte_sum'][0])}; first blocks at {where}")
