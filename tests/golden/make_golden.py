#!/usr/bin/env python3
"""Generates the committed golden vectors from the reference's OWN software backend
(oracle/_ref/libskity_ref.so, compiled from the unmodified sources by oracle/build_ref.py).

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py
Each fixture stores the SKSC scene blob, the encoded SKDL display list and the reference's
premultiplied RGBA8 output (or span list), so tests can replay them anywhere.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refsw  # noqa: E402
from skity_b200 import hostlib, scene  # noqa: E402
from skity_b200.scene import Paint, PathData, Scene  # noqa: E402


def golden_scenes():
    out = {}
    out["c0_star_blur_800x600"] = scene.scene_c0()
    out["c0_star_plain_800x600"] = scene.scene_c0(blur=False)
    out["c1_fills_120_512"] = scene.scene_random_fills(120, 512, 1, box=200.0)
    out["c2_gradients_90_512"] = scene.scene_c2(90, 512, 2, clip_every=0)
    out["c2_clips_90_512"] = scene.scene_c2(90, 512, 9, clip_every=30, clip_box=300.0)
    out["c3_blur_12_640"] = scene.scene_c3(12, 640, 3, box=160.0)
    # transforms, conics, even-odd, rect clip, big coordinates (int32 wrap of the 16.16 conversion at >= 8192 px)
    s = Scene(400, 300)
    s.save()
    s.translate(200, 150)
    s.rotate(30)
    s.scale(1.5, 0.75)
    p = PathData(scene.EVEN_ODD)
    p.move_to(-80, -60).conic_to(0, -120, 80, -60, 0.7071).line_to(60, 70).cubic_to(20, 10, -20, 130, -60, 70).close()
    p.move_to(-30, -20).line_to(30, -20).line_to(30, 30).line_to(-30, 30).close()
    s.draw_path(p, Paint(fill=(0.9, 0.3, 0.1, 0.8)))
    s.restore()
    s.save()
    s.clip_rect(20.5, 30.25, 250.75, 200.5)
    s.draw_rect(0, 0, 400, 300, Paint(fill=(0.1, 0.4, 0.9, 0.5)))
    s.draw_path(scene.star_path(), Paint(style=scene.STROKE, stroke=(0, 0.6, 0.2, 1), stroke_width=6.0,
                                         join=scene.ROUND_JOIN, cap=scene.ROUND_CAP))
    s.restore()
    out["mixed_transform_clip_400x300"] = s
    s = Scene(256, 256)
    p = PathData()
    p.move_to(8180, 10).line_to(8300, 40).line_to(8200, 200).close()   # x crosses 8192: reference wraps
    s.save()
    s.translate(-8100, 0)
    s.draw_path(p, Paint(fill=(0.2, 0.8, 0.3, 1)))
    s.restore()
    p2 = PathData()
    p2.move_to(10, 10).quad_to(250, 20, 120, 240).close()
    s.draw_path(p2, Paint(fill=(0.7, 0.1, 0.6, 0.6)))
    out["wrap_8192_256"] = s
    # the reference's own unit test SWCanvas.StrokeThenFillDrawsFillAfterStroke (test/ut/render/sw_canvas_test.cc:67-86)
    s = Scene(48, 48)
    s.draw_rect(10, 10, 34, 34, Paint(style=3, stroke_width=10.0, stroke=(1, 0, 0, 1), fill=(1, 1, 1, 1)))
    out["ut_stroke_then_fill_48"] = s
    # the reference's golden test ShapeGolden.CanonicalEdgesExact (test/golden/cases/shape/shape.cc:624-672)
    s = Scene(192, 144)
    s.draw_rect(0, 0, 192, 144, Paint(fill=(0, 0, 0, 1)))          # canvas->Clear(Color_BLACK)
    white = Paint(fill=(1, 1, 1, 1))

    def add_rect(p, l, t, r, b):                                    # Path::AddRect, kCW from the top-left corner
        return p.move_to(l, t).line_to(r, t).line_to(r, b).line_to(l, b).close()
    w = PathData(scene.WINDING)
    add_rect(w, 8.5, 8.5, 31.5, 31.5)
    w.move_to(40, 8).line_to(72, 24).line_to(40, 40).close()
    w.move_to(88, 8).line_to(88, 40).line_to(120, 24).close()
    w.move_to(16, 48).line_to(48, 64).line_to(16, 80).line_to(0, 64).close()
    add_rect(w, 128.5, 8.5, 151.5, 87.5)
    s.draw_path(w, white)
    e = PathData(scene.EVEN_ODD)
    add_rect(e, 72.5, 48.5, 119.5, 87.5)
    add_rect(e, 88.5, 60.5, 103.5, 76.5)
    s.draw_path(e, white)
    out["golden_canonical_edges_192x144"] = s
    out["blend_modes_480"] = scene.scene_blend_modes()
    out["filters_512"] = scene.scene_filters()
    out["filters_morphology_512"] = scene.scene_filters(34)
    out["layers_512"] = scene.scene_layers()
    out["filtered_layers_384"] = scene.scene_filtered_layers()
    out["conical_512"] = scene.scene_conical()
    out["color_filters_512"] = scene.scene_color_filters()
    out["images_512"] = scene.scene_images()
    out["clipped_blends_400"] = scene.scene_clipped_blends()
    # found by the GPU fuzz: StackBlur's seeding quirk yields pixels that are not valid premultiplied colours, whose
    # SrcOver sum overflows a channel and carries into the next one up in the reference's A|R|G|B registers
    out["filters_channel_carry_283"] = scene.scene_filters(5107, size=283)
    # found by the GPU fuzz: a pixel with a zero-coverage direct span AND an accumulated span under kSrcIn / kScreen
    out["blend_zero_then_accum_418"] = scene.scene_blend_modes(5016, size=418)
    # found by the GPU fuzz: a clip path whose spans all lie off the surface still clips (HasClip() is a non-empty
    # span list), and so does a nested clip whose only span is the zero-length one FindSpan makes when a parent
    # span ends exactly where an own span starts
    out["clip_spans_off_surface_75"] = scene.scene_fuzz(9002)[0]
    out["clip_zero_length_span_532"] = scene.scene_fuzz(11012)[0]
    # ... and such spans are handed down: here a third-level clip consists of one zero-length span inherited from
    # the zero-length / zero-coverage spans of the second level
    out["clip_inherited_ghost_span_354"] = scene.scene_fuzz(22152)[0]
    # ClipOp::kDifference (sw_canvas.cc:56-133,158-217): the reference's own golden case, and seeded scenes of the three
    # ways the op combines (one difference clip per Save level; difference and intersect nested; difference on
    # difference = PerformMerge)
    out["ref_clip_path_difference_400"] = scene.scene_ref_clip_path_difference()
    out["clip_difference_flat_8"] = scene.scene_difference_clips(8, "flat")
    out["clip_difference_flat_31"] = scene.scene_difference_clips(31, "flat")
    out["clip_difference_refined_12"] = scene.scene_difference_clips(12, "refined")
    out["clip_difference_carved_9"] = scene.scene_difference_clips(9, "carved")
    out["clip_difference_single_317"] = scene.scene_difference_clips(317, "single")
    out["clip_difference_single_44"] = scene.scene_difference_clips(44, "single")
    out["clip_difference_mixed_23"] = scene.scene_difference_clips(23, "mixed")
    out["clip_difference_merge_5"] = scene.scene_difference_clips(5, "merge")
    # two same-size images of different content, both temporaries (the encoder's image table must not confuse them)
    out["images_same_size_256"] = scene.scene_images_same_size()
    return out


SKP_TIGER = "/root/reference/resources/skp/tiger.skp"
SKP_TIGER_MATRIX = (1.0, 0.0, -130.0, 0.0, 1.0, 20.0)   # canvas->Translate(-130, 20), test/golden/cases/skp/skp.cc:60


def skp_fixture():
    """The reference's own SKP golden case (SKP_Golden.Tiger, test/golden/cases/skp/skp.cc:51-68): tiger.skp read by the
    reference's module/io and played back onto a 1000x1000 canvas under Translate(-130, 20) — onto the reference's
    software canvas (rgba) and onto the CUDA canvas (dl).  The .skp itself stays in the reference tree; its SHA-256 is
    stored so that a test can tell which file the vector came from."""
    import hashlib
    skp = open(SKP_TIGER, "rb").read()
    rgba = refsw.render_skp(skp, 1000, 1000, SKP_TIGER_MATRIX)
    dl = hostlib.encode_skp(skp, 1000, 1000, SKP_TIGER_MATRIX)
    path = os.path.join(HERE, "skp_tiger_1000.npz")
    np.savez_compressed(path, dl=np.frombuffer(dl, dtype=np.uint8), rgba=rgba,
                        skp_sha256=np.frombuffer(hashlib.sha256(skp).digest(), dtype=np.uint8))
    print(f"skp_tiger_1000: sum={int(rgba.astype(np.int64).sum())} -> {os.path.getsize(path)} bytes")


def main():
    assert refsw.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    only = set(sys.argv[1:])        # optional: regenerate just the named fixtures
    if not only or "skp_tiger_1000" in only:
        skp_fixture()
    for name, s in golden_scenes().items():
        if only and name not in only:
            continue
        blob = s.encode()
        rgba = refsw.render_scene(blob)
        dl = hostlib.encode_scene(blob)
        path = os.path.join(HERE, name + ".npz")
        extra = {}
        if name == "golden_canonical_edges_192x144":
            # the reference's checked-in golden image of this test (coverage-AA GPU backend, exact-match rule):
            # decoded here so the vector travels without the reference tree
            from PIL import Image
            png = "/root/reference/test/golden/cases/shape/coverage_aa_images/canonical_edges_exact.png"
            extra["reference_png"] = np.array(Image.open(png).convert("RGBA"))
        np.savez_compressed(path, scene=np.frombuffer(blob, dtype=np.uint8), dl=np.frombuffer(dl, dtype=np.uint8),
                            rgba=rgba, **extra)
        print(f"{name}: {rgba.shape[1]}x{rgba.shape[0]} sum={int(rgba.astype(np.int64).sum())} "
              f"-> {os.path.getsize(path)} bytes")
    if only and "raster_spans" not in only:
        return
    # span-level vectors: SWRaster::RastePath of a few paths
    rng = np.random.RandomState(11)
    spans_out = {}
    for i in range(12):
        p = scene._random_closed_path(rng, 100, 100, 180.0, i)
        sp, b = refsw.raster_path(p, clip=(0, 0, 200, 200))
        dlp = hostlib.encode_scene(_single(p).encode())
        spans_out[f"spans_{i}"] = sp
        spans_out[f"bounds_{i}"] = b
        spans_out[f"dl_{i}"] = np.frombuffer(dlp, dtype=np.uint8)
    sp, b = refsw.raster_path(scene.star_path())
    spans_out["spans_star"] = sp
    spans_out["bounds_star"] = b
    spans_out["dl_star"] = np.frombuffer(hostlib.encode_scene(_single(scene.star_path(), 400, 400).encode()), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "raster_spans.npz"), **spans_out)
    print("raster_spans.npz written")


def _single(path, w=200, h=200):
    s = Scene(w, h)
    s.draw_path(path, Paint(fill=(0, 0, 0, 1)))
    return s


if __name__ == "__main__":
    main()
