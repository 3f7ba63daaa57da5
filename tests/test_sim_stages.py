"""CPU simulation of the CUDA stages' per-thread code (flatten, setup, walk, per-pixel trapezoid
coverage) against the oracle's span lists: the kernels' logic is checked here without a GPU."""
import struct

import numpy as np
import pytest

import simlib
from oracle import port
from skity_b200 import hostlib, scene


def planes_from_spans(spans, w, h):
    p0 = np.zeros((h, w), np.uint8)
    p1 = np.zeros((h, w), np.uint8)
    cnt = np.zeros((h, w), np.int32)
    for x, y, ln, c in spans:
        if y < 0 or y >= h or c == 0:
            continue
        for xx in range(max(x, 0), min(x + ln, w)):
            k = cnt[y, xx]
            if k == 0:
                p0[y, xx] = c
            elif k == 1:
                p1[y, xx] = c
            cnt[y, xx] = k + 1
    return p0, p1, cnt


def check_scene(s):
    W, H = s.width, s.height
    dl = hostlib.encode_scene(s.encode())
    hd = port.dl_header(dl)
    both = 0
    for i in range(hd["n_ops"]):
        op = struct.unpack_from("<8I10f", dl, hd["off_ops"] + 72 * i)
        if op[0] != 1:
            continue
        ctm, clip = op[8:14], op[14:18]
        segs = port.dl_segments(dl, op[2])
        spans, _ = port.raster_path(segs, ctm, clip, op[6])
        d, a, nrec, stats = simlib.path_cover(segs, ctm, clip, op[6], W, H)
        p0, p1, cnt = planes_from_spans(spans, W, H)
        s0 = np.where(d > 0, d, a)
        s1 = np.where(d > 0, a, 0)
        assert cnt.max() <= 2, f"op {i}: a pixel is covered by more than two spans"
        assert np.array_equal(s0, p0) and np.array_equal(s1, p1), f"op {i}"
        both += int(stats[2])
    return both


def test_sampler_round_trip_is_identity():
    """BitmapSampler's nearest path converts a texel to floats and back (Color4fFromColor / Color4fToColor,
    src/graphic/color.cc:44-59): with IEEE single precision that is the identity on all 256 byte values, which the
    device sampler relies on (it uses the texel bytes as they are)."""
    assert simlib.lib().sim_requant_changes() == 0
    c = np.arange(256, dtype=np.float32)
    back = np.clip((c / np.float32(255.0)).astype(np.float32) * np.float32(255.0), 0, 255).astype(np.uint8)
    assert np.array_equal(back, np.arange(256, dtype=np.uint8))


def test_sim_star_and_fills():
    check_scene(scene.scene_c0(blur=False))
    check_scene(scene.scene_random_fills(40, 320, 1, box=200.0))


def test_sim_strokes_and_transforms():
    check_scene(scene.scene_c2(30, 320, 2, clip_every=0))
    z = np.load(__import__("os").path.join(__import__("conftest").ROOT, "tests", "golden", "wrap_8192_256.npz"))
    # replay the stored scene blob (coordinates beyond 8192 px wrap in the 16.16 conversion)
    from skity_b200.scene import Scene
    s = Scene(256, 256)
    blob = z["scene"].tobytes()
    n_ops = struct.unpack_from("<6I", blob, 0)[4]
    off = 24
    for _ in range(n_ops):
        code, nbytes = struct.unpack_from("<2I", blob, off)
        s.ops.append(blob[off:off + 8 + nbytes])
        off += 8 + nbytes
    check_scene(s)


def test_sim_whole_frame_with_nested_clips():
    """Clip stage logic (skb_clip.cuh) on the CPU: whole frames with nested ClipPath vs the oracle port."""
    import os
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", "c2_clips_90_512.npz"))
    got, stats = simlib.render_dl(z["dl"].tobytes())
    assert stats[0] == 0
    assert stats[3] == 0, "clip_row_seek must reproduce the sequential sweep state at any pixel of a row"
    assert np.array_equal(got, z["rgba"])
    s = scene.scene_c2(60, 384, 6, clip_every=12, clip_box=240.0, max_depth=3)
    dl = hostlib.encode_scene(s.encode())
    got, stats = simlib.render_dl(dl)
    assert stats[0] == 0
    assert np.array_equal(got, port.render(dl))


@pytest.mark.parametrize("mode", ["0", "1"])
def test_sim_sweep_other_forms_agree(monkeypatch, mode):
    """The sweep exists in more forms than the flat single loop covered above (skb_walk.cuh): the reference-shaped
    nested loops (0), and the flat loop run on the edges where the flatten stage left them (1; the default, like k_walk,
    runs it on a compact copy in sweep order, walk_prologue E2); all
    must produce the oracle's coverage."""
    monkeypatch.setenv("SKB_SIM_WALK_MODE", mode)
    check_scene(scene.scene_c0(blur=False))
    check_scene(scene.scene_c2(16, 256, 5, clip_every=0))
    check_scene(scene.scene_random_fills(24, 256, 9, box=160.0))


@pytest.mark.parametrize("seed", [61, 62])
def test_sim_clip_rows_shared_between_threads(seed):
    """More clipped frames through the CPU build of the clip stage: pixels equal to the oracle port, and the
    state a thread reconstructs when it enters a row in the middle (clip_row_seek) equal to the sequential one."""
    s = scene.scene_c2(40, 384, seed, clip_every=6, clip_box=260.0)
    dl = hostlib.encode_scene(s.encode())
    got, stats = simlib.render_dl(dl)
    assert stats[0] == 0 and stats[3] == 0
    assert np.array_equal(got, port.render(dl))


@pytest.mark.parametrize("name", ["clip_spans_off_surface_75", "clip_zero_length_span_532", "clip_inherited_ghost_span_354"])
def test_sim_clip_fixtures_found_by_fuzzing(name):
    """The clip-stack corner cases the GPU fuzz found (spans off the surface, spans that cover nothing), through the
    CPU build of the same stage code, against the compiled reference's pixels."""
    import os
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    got, stats = simlib.render_dl(z["dl"].tobytes())
    assert stats[0] == 0 and stats[3] == 0
    assert np.array_equal(got, z["rgba"])


def test_sweep_gradient_angle_is_the_c_librarys_atan2f():
    """The device evaluates the sweep gradient's angle with its own atan2f (skb_core.cuh: skb_atan2f), which must be
    glibc's bit for bit: the last bit of the angle decides the colour at a hard stop."""
    import ctypes
    lib = simlib.lib()
    lib.sim_atan2f_mismatches.restype = ctypes.c_long
    lib.sim_atan2f_mismatches.argtypes = [ctypes.c_long, ctypes.c_ulonglong]
    assert lib.sim_atan2f_mismatches(3_000_000, 12345) == 0


def test_fixed_point_division_in_fp64_is_the_integer_division():
    """fx_div computes SWFixedDiv's truncating 64-bit quotient in FP64 with one exact correction step (skb_core.cuh); the
    host build runs the same IEEE operations as the device: no operand pair may differ from the integer form."""
    import ctypes
    lib = simlib.lib()
    lib.sim_fx_div_mismatches.restype = ctypes.c_long
    lib.sim_fx_div_mismatches.argtypes = [ctypes.c_long, ctypes.c_ulonglong]
    assert lib.sim_fx_div_mismatches(20_000_000, 777) == 0


# ---- row-parallel walk (skb_rowwalk.cuh) against the sequential sweep, record by record ---------------------------
def _rowwalk_scene(s, w=None, h=None):
    import collections
    dl = hostlib.encode_scene(s.encode(), allow_unsupported=True)
    hd = port.dl_header(dl)
    res = collections.Counter()
    for i in range(hd["n_ops"]):
        op = struct.unpack_from("<8I10f", dl, hd["off_ops"] + 72 * i)
        if op[0] not in (1, 2):
            continue
        segs = port.dl_segments(dl, op[2])
        rc, _ = simlib.rowwalk_check(segs, op[8:14], op[14:18], op[6], w or s.width, h or s.height)
        assert rc != 2, f"op {i}: the row-parallel walk's records differ from the sequential sweep's"
        res[rc] += 1
    return res


def test_rowwalk_records_equal_the_sequential_sweep():
    """The row-parallel form of the sweep (one thread per path row: chords chained through per-path band tables,
    forcing bits composed over rows, every table entry re-derived by the final pass) must emit, row by row and in
    order, exactly the trapezoid records of the sequential sweep — or flag the path for it.  Fills, strokes with
    every join/cap, nested clip paths with their clip bounds (edges culled in y), star."""
    total = _rowwalk_scene(scene.scene_c0(blur=False))
    total += _rowwalk_scene(scene.scene_random_fills_fast(1500, 2048, 1, 256.0))
    total += _rowwalk_scene(scene.scene_random_fills_fast(1500, 1024, 4, 128.0))
    total += _rowwalk_scene(scene.scene_c2(400, 1024, 2, clip_every=0))
    total += _rowwalk_scene(scene.scene_c2(600, 1024, 5, clip_every=40, clip_box=300.0))
    done, flagged = total[0], total[1]
    assert done > 3500
    assert flagged <= 0.005 * done, f"{flagged} of {done + flagged} paths fell back to the sequential sweep"


@pytest.mark.parametrize("seed0", [0, 40])
def test_rowwalk_fuzz_scenes(seed0):
    """... and on the feature fuzz scenes (transforms, conics, layers, filters' temporaries, clips)."""
    total = {0: 0, 1: 0, 3: 0}
    for seed in range(seed0, seed0 + 40):
        for k, v in _rowwalk_scene(scene.scene_fuzz(seed)[0], 4096, 4096).items():
            total[k] = total.get(k, 0) + v
    assert total[0] > 500 and total[1] <= 0.01 * total[0]


def test_rowwalk_wide_coordinates():
    lib = simlib.lib()
    lib.sim_set_wide(1)
    try:
        s = scene.scene_random_fills_fast(1200, 16384, 4, 128.0)
        res = _rowwalk_scene(s)
    finally:
        lib.sim_set_wide(0)
    assert res[0] > 1100 and res[1] <= 3
