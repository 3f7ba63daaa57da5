"""Coverage mode AREA (north star stages 2-3: tile-binned lines, signed-area accumulation, backdrop prefix sums) —
the CPU side: the oracle's restatement is pinned against the reference's own compiled CoverageAAPathTiler and against
the reference's exact-match golden image; the device's per-thread code (skb_area.cuh, run thread by thread by the CPU
simulation) is compared with the oracle.  The GPU parity tests of the mode are in test_gpu_parity.py."""
import os
import struct

import numpy as np
import pytest

from conftest import ROOT
from oracle import port, refsw
from skity_b200 import hostlib, scene
from skity_b200.scene import Paint, PathData, Scene

GOLDEN = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not refsw.available(), reason="oracle/_ref not built (needs /root/reference)")
needs_host = pytest.mark.skipif(not os.path.exists(hostlib.LIB_PATH), reason="host plug-in not built")


def _segs(path):
    s = Scene(64, 64)
    s.draw_path(path, Paint())
    dl = hostlib.encode_scene(s.encode())
    return port.dl_segments(dl, 0)


def _shapes():
    """Shapes in the spirit of the reference's tiler unit tests (test/ut/render/hw/coverage_aa_path_tiler_test.cc):
    tile-aligned and unaligned rectangles, lines through tile corners in all four directions, shapes left of / above
    the scissor, conics, a self-intersecting star, an open contour, degenerate contours."""
    out = []
    def rect(l, t, r, b, ft=scene.WINDING):
        return PathData(ft).move_to(l, t).line_to(r, t).line_to(r, b).line_to(l, b).close()
    out.append(rect(0, 0, 16, 16))
    out.append(rect(16, 16, 48, 32))
    out.append(rect(3.25, 5.5, 60.75, 41.125))
    out.append(rect(-20, -20, 40, 40))
    out.append(PathData().move_to(0, 0).line_to(32, 32).line_to(0, 32).close())          # through tile corners, +x +y
    out.append(PathData().move_to(32, 32).line_to(0, 0).line_to(32, 0).close())          # -x -y
    out.append(PathData().move_to(0, 32).line_to(32, 0).line_to(32, 32).close())         # +x -y
    out.append(PathData().move_to(32, 0).line_to(0, 32).line_to(0, 0).close())           # -x +y
    out.append(PathData().move_to(16, 0).line_to(16, 48).line_to(40, 48).close())        # vertical on a tile edge
    out.append(PathData().move_to(0, 16).line_to(48, 16).line_to(48, 40).close())        # horizontal on a tile edge
    out.append(PathData(scene.EVEN_ODD).move_to(8, 8).line_to(56, 8).line_to(56, 56).line_to(8, 56).close()
               .move_to(20, 20).line_to(44, 20).line_to(44, 44).line_to(20, 44).close())
    out.append(scene.star_path())
    out.append(PathData().move_to(10, 40).conic_to(40, -30, 70, 40, 0.7071).line_to(40, 60).close())
    out.append(PathData().move_to(10, 40).conic_to(40, -30, 70, 40, 3.5).close())
    out.append(PathData().move_to(5, 5).quad_to(90, 10, 50, 80).cubic_to(20, 30, -40, 120, 5, 5))
    out.append(PathData().move_to(1, 1).line_to(9, 1).line_to(9, 9))                      # open: force-closed
    out.append(PathData().move_to(7, 7).line_to(7, 7).close())                            # degenerate
    out.append(PathData().move_to(5, 9).line_to(50, 9).close())                           # zero height
    return out


@needs_ref
@needs_host
def test_port_tiler_equals_the_reference_tiler_on_shapes():
    mats = [(1, 0, 0, 0, 1, 0), (1.5, 0.25, -7.5, -0.3, 0.8, 11.25), (0.5, 0, 100.5, 0, 0.5, -3.75)]
    scissors = [None, (8, 8, 40, 40), (20.5, 3.25, 200, 47.75), (-100, -100, 4, 4)]
    n = 0
    for p in _shapes():
        segs = _segs(p)
        for m in mats:
            for sc in scissors:
                rt, rl = refsw.coverage_aa_tile(p, m, sc)
                pt, pl = port.area_tile_path(segs, m, sc, p.fill_type == scene.EVEN_ODD)
                assert np.array_equal(rt, pt), (n, "tiles")
                assert np.array_equal(rl, pl), (n, "lines")
                n += 1
    assert n == len(_shapes()) * 12


@needs_ref
@needs_host
def test_port_tiler_equals_the_reference_tiler_on_random_paths():
    rng = np.random.RandomState(5)
    for i in range(150):
        p = scene._random_closed_path(rng, rng.uniform(0, 300), rng.uniform(0, 300), rng.uniform(10, 400), i)
        segs = _segs(p)
        ang, sc = rng.uniform(0, 6.28), rng.uniform(0.3, 2.5)
        m = ((np.cos(ang) * sc, -np.sin(ang) * sc, rng.uniform(-50, 50), np.sin(ang) * sc, np.cos(ang) * sc, rng.uniform(-50, 50))
             if i % 3 else (1, 0, 0, 0, 1, 0))
        scis = None if i % 2 else (20.5, 30.25, 250.75, 200.5)
        rt, rl = refsw.coverage_aa_tile(p, m, scis)
        pt, pl = port.area_tile_path(segs, m, scis, p.fill_type == scene.EVEN_ODD)
        assert np.array_equal(rt, pt) and np.array_equal(rl, pl), i


def test_port_area_coverage_equals_the_reference_golden_image():
    """ShapeGolden.CanonicalEdgesExact (test/golden/cases/shape/shape.cc:624-672): white paths on black compared with
    coverage_aa_images/canonical_edges_exact.png by the reference's exact-match rule — the PNG's grey level IS the A8
    coverage round(alpha * 255) of the reference's coverage-AA path."""
    z = np.load(os.path.join(GOLDEN, "golden_canonical_edges_192x144.npz"))
    dl, png = z["dl"].tobytes(), z["reference_png"]
    hd = port.dl_header(dl)
    cov = np.zeros((144, 192), dtype=np.uint8)
    for i in range(1, hd["n_ops"]):          # op 0 is the black clear
        op = struct.unpack_from("<8I10f", dl, hd["off_ops"] + 72 * i)
        segs = port.dl_segments(dl, op[2])
        cov = np.maximum(cov, port.area_coverage(segs, 192, 144, op[8:14], op[14:18], op[6] == 1))
    assert np.array_equal(cov, png[..., 0])
    assert set(np.unique(cov)) == {0, 64, 128, 191, 255}
    # through the software brush (AlphaMulQ truncates) the frame is within 1/255 of the image everywhere
    got = port.render_area(dl)
    assert np.abs(got.astype(int) - png.astype(int)).max() <= 1


@needs_host
def test_device_code_equals_the_port_thread_by_thread():
    """skb_area.cuh run by the CPU simulation (lines binned in reversed order) against the oracle's restatement."""
    import simlib
    rng = np.random.RandomState(7)
    paths = [(p, None) for p in _shapes()]
    for i in range(120):
        w = int(rng.randint(100, 400))
        paths.append((scene._random_closed_path(rng, rng.uniform(-20, w + 20), rng.uniform(-20, w + 20), rng.uniform(10, 500), i), w))
    for i, (p, w) in enumerate(paths):
        w = w or 96
        segs = _segs(p)
        ang, sc = rng.uniform(0, 6.28), rng.uniform(0.3, 2.5)
        m = ((np.cos(ang) * sc, -np.sin(ang) * sc, rng.uniform(-50, 50), np.sin(ang) * sc, np.cos(ang) * sc, rng.uniform(-50, 50))
             if i % 3 else (1, 0, 0, 0, 1, 0))
        clip = (0, 0, w, w) if i % 2 else (20.5, 30.25, w - 40.25, w - 17.5)
        eo = p.fill_type == scene.EVEN_ODD
        a = port.area_coverage(segs, w, w, m, clip, eo)
        b, _ = simlib.area_cover(segs, m, clip, eo, w, w)
        assert np.array_equal(a, b), i


def test_area_versus_software_coverage_histogram():
    """AREA is a different algorithm from the software backend's analytic AA: record how far apart the two are on a
    curved scene (the number DESIGN.md quotes), and that interiors agree."""
    z = np.load(os.path.join(GOLDEN, "c1_fills_120_512.npz"))
    dl, sw = z["dl"].tobytes(), z["rgba"]
    ar = port.render_area(dl)
    d = np.abs(ar.astype(np.int16) - sw.astype(np.int16)).max(axis=2)
    frac1, frac2 = float((d <= 1).mean()), float((d <= 2).mean())
    print(f"AREA vs software backend on c1_fills_120_512: <=1/255 on {frac1:.4%}, <=2/255 on {frac2:.4%}, max {int(d.max())}")
    assert frac2 > 0.90           # the bulk (interiors, background) agrees; edge pixels do not
