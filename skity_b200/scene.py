"""SKSC scene blobs: a serialised sequence of skity::Canvas calls.

One blob drives both the reference software canvas (oracle/_ref, through
skity_b200/host/scene_player.hpp) and the CUDA canvas, the way the reference's
golden harness replays one DisplayList on every backend
(test/golden/common/golden_test_env.hpp:45-47).  The synthetic generators
implement the seeded workloads of BASELINE.json / SURVEY.md §8(d).
"""
import struct

import numpy as np

MAGIC = 0x43534B53  # "SKSC"

# skity::Path::Verb (include/skity/graphic/path.hpp:45-60)
MOVE, LINE, QUAD, CONIC, CUBIC, CLOSE = 0, 1, 2, 3, 4, 5
# skity::Paint::Style / Cap / Join (include/skity/graphic/paint.hpp:46-104)
FILL, STROKE, STROKE_AND_FILL = 0, 1, 2
BUTT, ROUND_CAP, SQUARE = 0, 1, 2
MITER, ROUND_JOIN, BEVEL = 0, 1, 2
# skity::TileMode (include/skity/graphic/tile_mode.hpp:13-30)
CLAMP, REPEAT, MIRROR, DECAL = 0, 1, 2, 3
WINDING, EVEN_ODD = 0, 1

OP_SAVE, OP_RESTORE, OP_TRANSLATE, OP_SCALE, OP_ROTATE, OP_CONCAT = 1, 2, 3, 4, 5, 6
OP_CLIP_RECT, OP_CLIP_PATH, OP_DRAW_PATH, OP_DRAW_RECT, OP_SAVE_LAYER, OP_DRAW_IMAGE_RECT = 7, 8, 9, 10, 11, 12


class PathData:
    """Verbs + points in the reference's Path layout (xy only)."""

    def __init__(self, fill_type=WINDING):
        self.fill_type = fill_type
        self.verbs = []
        self.pts = []
        self.weights = []

    def move_to(self, x, y):
        self.verbs.append(MOVE)
        self.pts += [x, y]
        return self

    def line_to(self, x, y):
        self.verbs.append(LINE)
        self.pts += [x, y]
        return self

    def quad_to(self, x1, y1, x2, y2):
        self.verbs.append(QUAD)
        self.pts += [x1, y1, x2, y2]
        return self

    def conic_to(self, x1, y1, x2, y2, w):
        self.verbs.append(CONIC)
        self.pts += [x1, y1, x2, y2]
        self.weights.append(w)
        return self

    def cubic_to(self, x1, y1, x2, y2, x3, y3):
        self.verbs.append(CUBIC)
        self.pts += [x1, y1, x2, y2, x3, y3]
        return self

    def close(self):
        self.verbs.append(CLOSE)
        return self

    def encode(self):
        nv = len(self.verbs)
        pts = np.asarray(self.pts, dtype=np.float32)
        ws = np.asarray(self.weights, dtype=np.float32)
        verbs = bytes(self.verbs) + b"\0" * ((-nv) % 4)
        return struct.pack("<4I", self.fill_type, nv, len(pts) // 2, len(ws)) + verbs + pts.tobytes() + ws.tobytes()


class Paint:
    def __init__(self, style=FILL, fill=(0, 0, 0, 1), stroke=(0, 0, 0, 1), stroke_width=1.0, miter=4.0,
                 cap=BUTT, join=MITER, blur_radius=0.0, blur_style=1, shader=None, blend=None,
                 image_filter=None, color_filter=None):
        self.style, self.fill, self.stroke = style, fill, stroke
        self.stroke_width, self.miter, self.cap, self.join = stroke_width, miter, cap, join
        self.blur_radius, self.blur_style = blur_radius, blur_style
        self.blend = blend    # skity::BlendMode value, None = default (kSrcOver)
        # None | dict(type=1, sigma=(sx, sy)) ImageFilters::Blur | dict(type=2, offset=(dx, dy), sigma=(sx, sy), color=0xAARRGGBB)
        self.image_filter = image_filter
        # None | dict(type=1, color=0xAARRGGBB, mode=BlendMode) | dict(type=2, matrix=[20 floats]) | dict(type=3) | dict(type=4)
        self.color_filter = color_filter
        self.shader = shader  # dict(type=1|2|3, p=(..4), tile=, colors=[(r,g,b,a)..], stops=[..]|None, local=None|6)

    def encode(self):
        extras = self.blend is not None or self.image_filter is not None or self.color_filter is not None
        out = struct.pack("<I2f2I", self.style | (0x100 if extras else 0), self.stroke_width, self.miter, self.cap, self.join)
        out += np.asarray(self.fill, dtype=np.float32).tobytes()
        out += np.asarray(self.stroke, dtype=np.float32).tobytes()
        if self.blur_radius > 0:
            out += struct.pack("<If", self.blur_style, self.blur_radius)
        else:
            out += struct.pack("<If", 0, 0.0)
        out += self._encode_shader()
        if extras:
            f = self.image_filter or {}
            dx, dy = f.get("offset", (0.0, 0.0))
            sx, sy = f.get("sigma", (0.0, 0.0))
            out += struct.pack("<2I4fI", 3 if self.blend is None else self.blend, f.get("type", 0), dx, dy, sx, sy,
                               f.get("color", 0))
            c = self.color_filter or {}
            out += struct.pack("<3I", c.get("type", 0), c.get("color", 0), c.get("mode", 3))
            out += np.asarray(c.get("matrix", [0.0] * 20), dtype=np.float32).tobytes()
        return out

    def _encode_shader(self):
        out = b""
        sh = self.shader
        if not sh:
            return struct.pack("<I", 0)
        colors = np.asarray(sh["colors"], dtype=np.float32).reshape(-1, 4)
        stops = sh.get("stops")
        stops = np.asarray(stops if stops is not None else [], dtype=np.float32)
        local = sh.get("local")
        out += struct.pack("<I", sh["type"])
        out += np.asarray(sh["p"], dtype=np.float32).tobytes()
        out += struct.pack("<4I", sh.get("tile", CLAMP), len(colors), len(stops), 1 if local is not None else 0)
        out += np.asarray(local if local is not None else [1, 0, 0, 0, 1, 0], dtype=np.float32).tobytes()
        out += colors.tobytes() + stops.tobytes()
        if sh["type"] == 4:
            out += struct.pack("<2f", *sh["radii"])
        if sh["type"] == 5:   # image shader: p = (w, h, seed, unpremul); tile = x mode
            out += struct.pack("<2I", sh.get("tile_y", sh.get("tile", CLAMP)), sh.get("filter", 0))
        return out


class Scene:
    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)
        self.ops = []
        self.n_draws = 0

    def _op(self, code, payload=b""):
        assert len(payload) % 4 == 0
        self.ops.append(struct.pack("<2I", code, len(payload)) + payload)

    def save(self):
        self._op(OP_SAVE)

    def restore(self):
        self._op(OP_RESTORE)

    def draw_image_rect(self, image, src, dst, paint, filter=0):
        """Canvas::DrawImageRect of the procedural test image `image` = (w, h, seed, unpremul)."""
        self._op(OP_DRAW_IMAGE_RECT, struct.pack("<4I", *image) + struct.pack("<8f", *src, *dst) + struct.pack("<I", filter) +
                 paint.encode())
        self.n_draws += 1

    def save_layer(self, l, t, r, b, paint):
        """Canvas::SaveLayer(bounds, paint); the matching restore() composites the layer with `paint`."""
        self._op(OP_SAVE_LAYER, struct.pack("<4f", l, t, r, b) + paint.encode())

    def translate(self, dx, dy):
        self._op(OP_TRANSLATE, struct.pack("<2f", dx, dy))

    def scale(self, sx, sy):
        self._op(OP_SCALE, struct.pack("<2f", sx, sy))

    def rotate(self, deg):
        self._op(OP_ROTATE, struct.pack("<f", deg))

    def concat(self, m6):
        self._op(OP_CONCAT, np.asarray(m6, dtype=np.float32).tobytes())

    def clip_rect(self, l, t, r, b, intersect=True):
        self._op(OP_CLIP_RECT, struct.pack("<4fI", l, t, r, b, 1 if intersect else 0))

    def clip_path(self, path, intersect=True):
        self._op(OP_CLIP_PATH, path.encode() + struct.pack("<I", 1 if intersect else 0))

    def draw_path(self, path, paint):
        self._op(OP_DRAW_PATH, path.encode() + paint.encode())
        self.n_draws += 1

    def draw_rect(self, l, t, r, b, paint):
        self._op(OP_DRAW_RECT, struct.pack("<4f", l, t, r, b) + paint.encode())
        self.n_draws += 1

    def encode(self):
        return struct.pack("<6I", MAGIC, 1, self.width, self.height, len(self.ops), 0) + b"".join(self.ops)


# ---------------------------------------------------------------------------
# Workloads.  numpy RandomState is MT19937, seeded per config.
# ---------------------------------------------------------------------------

def star_path(fill_type=WINDING):
    """README star (README.md:105-126; also test/golden/cases/clip/clip.cc:51-65)."""
    p = PathData(fill_type)
    p.move_to(199, 34)
    for x, y in [(253, 143), (374, 160), (287, 244), (307, 365), (199, 309), (97, 365), (112, 245), (26, 161),
                 (146, 143)]:
        p.line_to(x, y)
    return p.close()


README_BLUE = (0x42 / 255.0, 0x85 / 255.0, 0xF4 / 255.0, 1.0)


def scene_c0(blur=True, plain=True):
    """Config 0: README star, kFill, nonzero, 0x4285F4, plus MakeBlur(kNormal,10) copy at +400."""
    s = Scene(800, 600)
    if plain:
        s.draw_path(star_path(), Paint(fill=README_BLUE))
    if blur:
        s.save()
        s.translate(400, 0)
        s.draw_path(star_path(), Paint(fill=README_BLUE, blur_radius=10.0))
        s.restore()
    return s


def _random_closed_path(rng, cx, cy, box, index, n_seg=4):
    """4 segments, even index quads / odd index cubics, control points uniform in a box."""
    half = box * 0.5

    def pt():
        return (np.float32(cx + rng.uniform(-half, half)), np.float32(cy + rng.uniform(-half, half)))

    p = PathData(EVEN_ODD if index % 3 == 0 else WINDING)
    p.move_to(*pt())
    for _ in range(n_seg):
        if index % 2 == 0:
            a, b = pt(), pt()
            p.quad_to(a[0], a[1], b[0], b[1])
        else:
            a, b, c = pt(), pt(), pt()
            p.cubic_to(a[0], a[1], b[0], b[1], c[0], c[1])
    return p.close()


def scene_random_fills(n_paths, size, seed, box=256.0, width=None, height=None):
    """C1 / C4a / C4b family: random quad/cubic closed paths, solid colour, alpha in [0.5,1]."""
    w = width or size
    h = height or size
    rng = np.random.RandomState(seed)
    s = Scene(w, h)
    for i in range(n_paths):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        path = _random_closed_path(rng, cx, cy, box, i)
        col = (rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.5, 1.0))
        s.draw_path(path, Paint(fill=tuple(np.float32(c) for c in col)))
    return s


def scene_c1(n_paths=10000, size=4096, seed=1):
    return scene_random_fills_fast(n_paths, size, seed, 256.0)


def scene_c4a(n_paths=1000000, size=16384, seed=4):
    return scene_random_fills_fast(n_paths, size, seed, 128.0)


def scene_c4b(index, n_paths=1000):
    return scene_random_fills_fast(n_paths, 0, 5 + index, 256.0, width=1920, height=1080)


def _random_gradient(rng, cx, cy, box, kind):
    n = int(rng.randint(3, 6))
    colors = [(rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.5, 1.0)) for _ in range(n)]
    stops = None
    if rng.uniform() < 0.5:
        stops = np.sort(rng.uniform(0, 1, n)).astype(np.float32)
        stops[0], stops[-1] = 0.0, 1.0
    tile = [CLAMP, REPEAT, MIRROR][int(rng.randint(0, 3))]
    half = box * 0.5
    if kind == 1:
        p = (cx - rng.uniform(0, half), cy - rng.uniform(0, half), cx + rng.uniform(0, half), cy + rng.uniform(0, half))
    elif kind == 2:
        p = (cx, cy, rng.uniform(box * 0.1, half), 0.0)
    else:
        p = (cx, cy, 0.0, 360.0)
    return dict(type=kind, p=tuple(np.float32(v) for v in p), tile=tile, colors=colors, stops=stops)


def scene_c2(n_paths=20000, size=4096, seed=2, clip_every=500, clip_box=1024.0, max_depth=3):
    """Gradient-heavy: fill/stroke/stroke+fill cycle, linear/radial/sweep shaders, nested ClipPath."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    depth = 0
    for i in range(n_paths):
        if clip_every and i % clip_every == 0:
            if depth >= max_depth:
                while depth > 0:
                    s.restore()
                    depth -= 1
            s.save()
            depth += 1
            ccx, ccy = rng.uniform(size * 0.25, size * 0.75), rng.uniform(size * 0.25, size * 0.75)
            blob = _random_closed_path(rng, ccx, ccy, clip_box * max(1.0, size / 2048.0), 1)
            blob.fill_type = WINDING
            s.clip_path(blob, True)
        cx, cy = rng.uniform(0, size), rng.uniform(0, size)
        path = _random_closed_path(rng, cx, cy, 256.0, i)
        style = i % 3
        kind = 1 + (i % 3)
        paint = Paint(style=style, stroke_width=float(np.float32(rng.uniform(1, 8))), cap=(i // 3) % 3,
                      join=(i // 9) % 3, fill=(0, 0, 0, 1), stroke=(0, 0, 0, 1),
                      shader=_random_gradient(rng, cx, cy, 256.0, kind))
        s.draw_path(path, paint)
    while depth > 0:
        s.restore()
        depth -= 1
    return s


def scene_c3(n_paths=2000, size=8192, seed=3, box=512.0):
    """Blur stress: MaskFilter::MakeBlur(kNormal, r), sigma U[4,64] => r=(sigma-0.5)/0.57735."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    for i in range(n_paths):
        cx, cy = rng.uniform(0, size), rng.uniform(0, size)
        path = _random_closed_path(rng, cx, cy, box, i)
        col = (rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.5, 1.0))
        sigma = rng.uniform(4.0, 64.0)
        radius = float(np.float32((sigma - 0.5) / 0.57735))
        s.draw_path(path, Paint(fill=tuple(np.float32(c) for c in col), blur_radius=radius))
    return s


# ---------------------------------------------------------------------------
# Vectorised generator for the random-fill family (identical bytes to scene_random_fills, ~100x faster)
# ---------------------------------------------------------------------------
class SceneBlob:
    """A pre-encoded SKSC scene (same interface as Scene for the harnesses)."""

    def __init__(self, width, height, n_draws, blob):
        self.width, self.height, self.n_draws, self._blob = int(width), int(height), int(n_draws), blob

    def encode(self):
        return self._blob


def scene_random_fills_fast(n_paths, size, seed, box=256.0, width=None, height=None):
    w = width or size
    h = height or size
    rng = np.random.RandomState(seed)
    n_even = (n_paths + 1) // 2          # quads
    n_odd = n_paths // 2                 # cubics
    per_even, per_odd = 2 + 2 + 16 + 4, 2 + 2 + 24 + 4
    # the scalar generator interleaves even/odd paths: reproduce its exact draw order
    counts = np.where(np.arange(n_paths) % 2 == 0, per_even, per_odd)
    starts = np.concatenate([[0], np.cumsum(counts)])
    u = rng.random_sample(int(starts[-1]))
    half = box * 0.5

    def build(idx, n_pts_xy, verbs, op_bytes):
        m = len(idx)
        if m == 0:
            return np.zeros((0, op_bytes), np.uint8)
        base = starts[idx]
        cx = 0.0 + (w - 0.0) * u[base]
        cy = 0.0 + (h - 0.0) * u[base + 1]
        k = np.arange(n_pts_xy)
        raw = u[base[:, None] + 2 + k[None, :]]
        off = -half + (half - -half) * raw
        centre = np.where(k[None, :] % 2 == 0, cx[:, None], cy[:, None])
        pts = (centre + off).astype(np.float32)
        cb = base + 2 + n_pts_xy
        col = np.stack([0.0 + 1.0 * u[cb], 0.0 + 1.0 * u[cb + 1], 0.0 + 1.0 * u[cb + 2], 0.5 + (1.0 - 0.5) * u[cb + 3]],
                       axis=1).astype(np.float32)
        rec = np.zeros((m, op_bytes), np.uint8)
        nv = len(verbs)
        vpad = (nv + 3) & ~3
        path_bytes = 16 + vpad + 4 * n_pts_xy
        hdr = np.zeros((m, 6), np.uint32)
        hdr[:, 0] = OP_DRAW_PATH
        hdr[:, 1] = op_bytes - 8
        hdr[:, 2] = np.where(idx % 3 == 0, EVEN_ODD, WINDING)
        hdr[:, 3] = nv
        hdr[:, 4] = n_pts_xy // 2
        hdr[:, 5] = 0
        rec[:, :24] = hdr.view(np.uint8).reshape(m, 24)
        rec[:, 24:24 + nv] = np.asarray(verbs, np.uint8)[None, :]
        rec[:, 24 + vpad:24 + vpad + 4 * n_pts_xy] = pts.view(np.uint8).reshape(m, 4 * n_pts_xy)
        po = 8 + path_bytes
        paint = np.zeros((m, 16), np.float32)
        pu = paint.view(np.uint32)
        pu[:, 0] = FILL
        paint[:, 1] = 1.0
        paint[:, 2] = 4.0
        pu[:, 3] = BUTT
        pu[:, 4] = MITER
        paint[:, 5:9] = col
        paint[:, 9:13] = np.asarray([0, 0, 0, 1], np.float32)[None, :]
        rec[:, po:po + 64] = paint.view(np.uint8).reshape(m, 64)
        return rec

    even_idx = np.arange(0, n_paths, 2)
    odd_idx = np.arange(1, n_paths, 2)
    ev = build(even_idx, 18, [MOVE, QUAD, QUAD, QUAD, QUAD, CLOSE], 168)
    od = build(odd_idx, 26, [MOVE, CUBIC, CUBIC, CUBIC, CUBIC, CLOSE], 200)
    pair = 168 + 200
    body = np.zeros(n_even * 168 + n_odd * 200, np.uint8)
    if n_odd:
        both = body[:n_odd * pair].reshape(n_odd, pair)
        both[:, :168] = ev[:n_odd]
        both[:, 168:] = od
    if n_even > n_odd:
        body[n_odd * pair:] = ev[-1]
    return SceneBlob(w, h, n_paths, struct.pack("<6I", MAGIC, 1, w, h, n_paths, 0) + body.tobytes())


def scene_blend_modes(seed=21, size=480):
    """Every blend mode SWRenderTarget implements (kClear..kScreen, kSoftLight; one unsupported mode,
    kOverlay, which falls back to kSrcOver: src/graphic/blend_mode.cc:125-192) over a busy backdrop,
    with translucent and opaque sources, a gradient source and anti-aliased edges."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size * 0.5, Paint(fill=(0.9, 0.8, 0.2, 1.0)))
    s.draw_rect(0, size * 0.25, size * 0.6, size, Paint(fill=(0.1, 0.5, 0.9, 0.6)))
    for i in range(10):
        p = _random_closed_path(rng, rng.uniform(0, size), rng.uniform(0, size), 200.0, i)
        s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (rng.uniform(0.3, 1.0),)))
    modes = list(range(0, 15)) + [21, 15]
    cols = 5
    cell = size / cols
    for k, mode in enumerate(modes):
        cx, cy = (k % cols + 0.5) * cell, (k // cols + 0.5) * cell
        p = _random_closed_path(rng, cx, cy, cell * 1.1, k)
        alpha = 1.0 if k % 3 == 0 else float(rng.uniform(0.3, 0.9))
        if k % 4 == 3:
            sh = dict(type=1, p=(cx - cell / 2, cy, cx + cell / 2, cy), tile=CLAMP,
                      colors=[(1, 0, 0, 1), (0, 1, 0, 0.5), (0, 0, 1, 1)], stops=None)
            s.draw_path(p, Paint(shader=sh, blend=mode))
        else:
            s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (alpha,), blend=mode))
    return s


def scene_filters(seed=33, size=512):
    """Mask-filter blur styles (kNormal/kSolid/kOuter/kInner, src/effect/mask_filter.cc:51-103) and the image
    filters the SW backend implements by blurring (ImageFilters::Blur / DropShadow,
    src/effect/image_filter.cc:196-238), on fills and strokes, over a non-empty backdrop."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.95, 0.95, 0.9, 1.0)))
    s.draw_rect(size * 0.3, 0, size * 0.6, size, Paint(fill=(0.2, 0.3, 0.5, 0.5)))
    cell = size / 4
    k = 0
    for style in (1, 2, 3, 4):
        for j in range(2):
            cx, cy = (k % 4 + 0.5) * cell, (k // 4 + 0.5) * cell
            p = _random_closed_path(rng, cx, cy, cell * 0.7, k)
            col = tuple(rng.uniform(0, 1, 3)) + ((1.0,) if j == 0 else (0.6,))
            if j == 0:
                s.draw_path(p, Paint(fill=col, blur_radius=float(rng.uniform(2, 14)), blur_style=style))
            else:
                s.draw_path(p, Paint(style=STROKE, stroke=col, stroke_width=6.0, blur_radius=float(rng.uniform(2, 9)),
                                     blur_style=style))
            k += 1
    for j in range(4):
        cx, cy = (k % 4 + 0.5) * cell, (k // 4 + 0.5) * cell
        p = _random_closed_path(rng, cx, cy, cell * 0.7, k)
        col = tuple(rng.uniform(0, 1, 3)) + (float(rng.uniform(0.5, 1.0)),)
        if j < 2:
            f = dict(type=1, sigma=(float(rng.uniform(1, 6)), float(rng.uniform(1, 6))))
        else:
            f = dict(type=2, offset=(float(rng.uniform(-9, 9)), float(rng.uniform(3, 9))),
                     sigma=(float(rng.uniform(1, 5)), float(rng.uniform(1, 5))), color=0xC0102060 if j == 2 else 0xFF203010)
        s.draw_path(p, Paint(fill=col, image_filter=f))
        k += 1
    if seed == 34:  # morphology image filters (a later addition: other seeds keep their committed fixtures)
        for j, f in enumerate([dict(type=3, sigma=(3.0, 2.0)), dict(type=4, sigma=(2.0, 3.0)), dict(type=3, sigma=(4.0, 0.0)),
                               dict(type=4, sigma=(0.0, 2.5)), dict(type=3, sigma=(0.5, 0.5))]):
            p = _random_closed_path(rng, (j + 0.5) * size / 5, size * 0.5, size / 5.5, j)
            s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (float(rng.uniform(0.5, 1.0)),), image_filter=f))
            s.draw_path(p, Paint(style=STROKE, stroke=(0.1, 0.1, 0.1, 0.8), stroke_width=3.0, image_filter=f))
    s.save()
    s.translate(size * 0.5, size * 0.85)
    s.rotate(-12)
    s.scale(1.4, 0.8)
    s.draw_path(_random_closed_path(rng, 0, 0, 90.0, 3),
                Paint(fill=(0.8, 0.1, 0.2, 0.9), image_filter=dict(type=2, offset=(6.0, 5.0), sigma=(3.0, 3.0), color=0x80000000)))
    s.restore()
    return s


def scene_layers(seed=44, size=512):
    """SaveLayer (SWCanvas::OnSaveLayer / OnLayerRestore, src/render/sw/sw_canvas.cc:441-484,891-902):
    translucent and blend-mode layers, a layer under a rotated CTM (fractional device-space origin), Save /
    clips inside a layer, a blurred draw inside a layer, and a layer nested in a layer."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.92, 0.9, 0.85, 1.0)))
    for i in range(6):
        p = _random_closed_path(rng, rng.uniform(0, size), rng.uniform(0, size), 220.0, i)
        s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (rng.uniform(0.4, 1.0),)))
    # 1. translucent group
    s.save_layer(40.5, 30.25, 260.75, 240.5, Paint(fill=(0, 0, 0, 0.55)))
    s.draw_rect(60, 50, 200, 180, Paint(fill=(0.9, 0.1, 0.1, 1.0)))
    s.draw_path(_random_closed_path(rng, 170, 150, 160.0, 1), Paint(fill=(0.1, 0.7, 0.2, 0.8)))
    s.restore()
    # 2. layer under a rotated, scaled CTM, with a Save + rect clip + path clip inside
    s.save()
    s.translate(330, 140)
    s.rotate(17)
    s.scale(1.2, 0.9)
    s.save_layer(-110, -90, 120, 100, Paint(fill=(0, 0, 0, 0.8), blend=14))
    s.draw_rect(-100, -80, 100, 90, Paint(fill=(0.2, 0.3, 0.9, 0.9)))
    s.save()
    s.clip_rect(-60, -50, 70, 60)
    s.clip_path(star_path_small(70.0))
    s.draw_rect(-100, -80, 100, 90, Paint(fill=(1.0, 0.9, 0.1, 1.0)))
    s.restore()
    s.draw_path(_random_closed_path(rng, 30, 20, 120.0, 2), Paint(style=STROKE, stroke=(0, 0, 0, 0.7), stroke_width=5.0))
    s.restore()
    s.restore()
    # 3. nested layers with a blurred draw inside
    s.save_layer(60, 280, 470, 500, Paint(fill=(0, 0, 0, 0.9)))
    s.draw_rect(80, 300, 300, 480, Paint(fill=(0.1, 0.6, 0.7, 1.0)))
    s.draw_path(_random_closed_path(rng, 330, 390, 150.0, 3), Paint(fill=(0.8, 0.2, 0.6, 1.0), blur_radius=6.0, blur_style=1))
    s.save_layer(200.5, 320.5, 450, 470, Paint(fill=(0, 0, 0, 0.5), blend=12))
    s.draw_path(_random_closed_path(rng, 320, 400, 170.0, 4), Paint(fill=(0.9, 0.8, 0.1, 1.0)))
    s.restore()
    s.restore()
    return s


def star_path_small(r):
    """Five-pointed star centred on the origin (winding fill)."""
    p = PathData(WINDING)
    for k in range(5):
        a = -np.pi / 2 + k * 4 * np.pi / 5
        x, y = float(r * np.cos(a)), float(r * np.sin(a))
        if k == 0:
            p.move_to(x, y)
        else:
            p.line_to(x, y)
    return p.close()


def scene_conical(seed=55, size=512):
    """Two-point conical gradients (ConicalGradientColorBrush, src/render/sw/sw_span_brush.cc:385-530): the
    general case with r1 above, below and at the focal unit radius, swapped circles (r1 ~ 0), concentric
    circles, equal radii, a negative radius; all tile modes, with stops and a local matrix."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(1, 1, 1, 1)))
    cell = size / 4
    cases = [
        ((0.3, 0.5), 0.1, (0.7, 0.5), 0.45),   # general, growing
        ((0.3, 0.5), 0.45, (0.7, 0.5), 0.1),   # general, shrinking
        ((0.3, 0.3), 0.2, (0.6, 0.6), 0.0),    # end radius 0: circles swapped
        ((0.5, 0.5), 0.05, (0.5, 0.5), 0.5),   # concentric
        ((0.2, 0.5), 0.25, (0.8, 0.5), 0.25),  # equal radii: strip
        ((0.5, 0.5), 0.3, (0.5, 0.5), 0.3),    # concentric and equal: transparent
        ((0.3, 0.5), -0.1, (0.7, 0.5), 0.4),   # negative radius: transparent
        ((0.2, 0.2), 0.0, (0.5, 0.5), 0.6),    # start radius 0 (focal on centre)
        ((0.4, 0.5), 0.2, (0.6, 0.5), 0.2 + 0.2),  # r1 relative to focal distance near 1
        ((0.45, 0.5), 0.3, (0.6, 0.5), 0.35),
        ((0.1, 0.9), 0.05, (0.9, 0.1), 0.3),
        ((0.5, 0.2), 0.15, (0.5, 0.8), 0.5),
    ]
    for k, (a, ra, b, rb) in enumerate(cases):
        x0, y0 = (k % 4) * cell, (k // 4) * cell
        colors = [tuple(rng.uniform(0, 1, 3)) + (float(rng.uniform(0.5, 1)),) for _ in range(3 + k % 3)]
        stops = sorted(float(v) for v in rng.uniform(0, 1, len(colors))) if k % 2 else None
        if stops:
            stops[0], stops[-1] = 0.0, 1.0
        sh = dict(type=4, p=(x0 + a[0] * cell, y0 + a[1] * cell, x0 + b[0] * cell, y0 + b[1] * cell),
                  radii=(ra * cell, rb * cell), tile=k % 4, colors=colors, stops=stops,
                  local=(1.0, 0.1, 3.0, -0.05, 0.95, -2.0) if k % 5 == 4 else None)
        s.draw_rect(x0 + 4, y0 + 4, x0 + cell - 4, y0 + cell - 4, Paint(shader=sh))
    s.save()
    s.translate(size * 0.5, size * 0.88)
    s.rotate(20)
    sh = dict(type=4, p=(-60, 0, 40, 10), radii=(10.0, 70.0), tile=MIRROR, colors=[(1, 0, 0, 1), (0, 0, 1, 0.6), (0, 1, 0, 1)], stops=None)
    s.draw_path(_random_closed_path(rng, 0, 0, 200.0, 2), Paint(shader=sh))
    s.restore()
    return s


def scene_fuzz(seed):
    """Seeded random scene of one of ten feature classes (seed % 10); returns (scene, bit_exact_expected).
    Used by tests/gpu_fuzz.py; scenes that exposed bugs are kept as golden fixtures."""
    rng = np.random.RandomState(seed)
    kind = seed % 10
    w, h = int(rng.randint(40, 700)), int(rng.randint(40, 700))
    if seed >= 1000000:  # later additions (seeds below keep their meaning: some are golden fixtures)
        if seed % 3 == 1:
            return scene_images(seed, size=int(rng.randint(200, 600))), False
        if seed % 3 == 2:
            return scene_clipped_blends(seed, size=int(rng.randint(150, 500))), True
        return scene_color_filters(seed, size=int(rng.randint(200, 600))), False
    if kind == 0:
        return scene_random_fills(int(rng.randint(5, 150)), 0, seed, box=float(rng.uniform(20, 500)), width=w, height=h), True
    if kind == 1:
        return scene_c2(int(rng.randint(5, 80)), max(w, h), seed, clip_every=0), False
    if kind == 2:
        return scene_c2(int(rng.randint(10, 80)), max(w, h), seed, clip_every=int(rng.randint(4, 20)),
                              clip_box=float(rng.uniform(60, 400))), False
    if kind == 3:
        return scene_c3(int(rng.randint(1, 8)), max(w, h), seed, box=float(rng.uniform(30, 200))), True
    if kind == 4:  # transforms + conics + tiny/huge shapes
        s = Scene(w, h)
        for i in range(int(rng.randint(3, 40))):
            s.save()
            s.translate(float(rng.uniform(0, w)), float(rng.uniform(0, h)))
            s.rotate(float(rng.uniform(0, 360)))
            s.scale(float(rng.uniform(0.2, 3)), float(rng.uniform(0.2, 3)))
            p = PathData(int(rng.randint(0, 2)))
            r = float(rng.uniform(0.3, 120))
            p.move_to(r, 0).conic_to(r, r, 0, r, 0.7071).conic_to(-r, r, -r, 0, 0.7071)
            p.conic_to(-r, -r, 0, -r, float(rng.uniform(0.1, 3))).conic_to(r, -r, r, 0, 0.7071).close()
            if rng.uniform() < 0.5:
                p.move_to(-r / 3, -r / 3).line_to(r / 3, -r / 3).line_to(0, r / 2).close()
            col = tuple(np.float32(v) for v in rng.uniform(0, 1, 4))
            style = int(rng.randint(0, 4))
            s.draw_path(p, Paint(style=style, fill=col, stroke=col[::-1], stroke_width=float(rng.uniform(0.2, 12)),
                                 cap=int(rng.randint(0, 3)), join=int(rng.randint(0, 3))))
            s.restore()
        return s, True
    if kind == 6:
        return scene_blend_modes(seed, size=int(rng.randint(160, 600))), False
    if kind == 7:
        return scene_filters(seed, size=int(rng.randint(200, 600))), True
    if kind == 8:
        return scene_layers(seed, size=int(rng.randint(300, 640))), True
    if kind == 9:
        return scene_conical(seed, size=int(rng.randint(200, 600))), False
    s = Scene(w, h)  # solid draws under nested clips and rect clips
    depth = 0
    for i in range(int(rng.randint(5, 60))):
        if rng.uniform() < 0.2 and depth < 3:
            s.save(); depth += 1
            if rng.uniform() < 0.5:
                s.clip_path(_random_closed_path(rng, rng.uniform(0, w), rng.uniform(0, h), float(rng.uniform(50, 400)), int(rng.randint(0, 6))), True)
            else:
                x, y = rng.uniform(0, w), rng.uniform(0, h)
                s.clip_rect(float(x), float(y), float(x + rng.uniform(10, 300)), float(y + rng.uniform(10, 300)), True)
        elif rng.uniform() < 0.1 and depth > 0:
            s.restore(); depth -= 1
        col = tuple(np.float32(v) for v in rng.uniform(0, 1, 4))
        s.draw_path(_random_closed_path(rng, rng.uniform(0, w), rng.uniform(0, h), float(rng.uniform(20, 400)), i),
                    Paint(style=int(rng.randint(0, 3)), fill=col, stroke=col, stroke_width=float(rng.uniform(0.5, 9))))
    while depth:
        s.restore(); depth -= 1
    return s, True


def scene_color_filters(seed=66, size=512):
    """Paint colour filters (src/effect/color_filter.cc): ColorFilters::Blend with several modes, Matrix (saturation,
    channel swap, alpha-changing), both sRGB gamma tables; on solid, gradient, translucent and blurred draws, combined
    with non-default blend modes."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.85, 0.9, 0.95, 1.0)))
    for i in range(5):
        p = _random_closed_path(rng, rng.uniform(0, size), rng.uniform(0, size), 260.0, i)
        s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (rng.uniform(0.4, 1.0),)))
    sat = [0.3, 0.6, 0.1, 0, 0, 0.3, 0.6, 0.1, 0, 0, 0.3, 0.6, 0.1, 0, 0, 0, 0, 0, 1, 0]
    swap = [0, 0, 1, 0, 0, 0, 1, 0, 0, 0.1, 1, 0, 0, 0, -0.1, 0, 0, 0, 0.8, 0.1]
    wild = [1.5, -0.5, 0, 0, 0.2, 0, 1.2, 0, 0.3, -0.2, 0.4, 0.4, 0.4, 0, 0, 0.2, 0.2, 0.2, 0.5, 0]
    filters = [dict(type=1, color=0x80FF2010, mode=3), dict(type=1, color=0xFF2040C0, mode=13), dict(type=1, color=0x6010A020, mode=5),
               dict(type=1, color=0xC0C0C000, mode=14), dict(type=1, color=0xFF808080, mode=1), dict(type=1, color=0x40FFFFFF, mode=12),
               dict(type=2, matrix=sat), dict(type=2, matrix=swap), dict(type=2, matrix=wild), dict(type=3), dict(type=4),
               dict(type=1, color=0x90003366, mode=9)]
    cell = size / 4
    for k, cf in enumerate(filters):
        cx, cy = (k % 4 + 0.5) * cell, (k // 4 + 0.5) * cell
        p = _random_closed_path(rng, cx, cy, cell * 1.05, k)
        alpha = 1.0 if k % 2 == 0 else float(rng.uniform(0.3, 0.9))
        if k % 3 == 1:
            sh = dict(type=2, p=(cx, cy, cell * 0.6, 0), tile=MIRROR, colors=[(1, 0.2, 0, 1), (0, 0.8, 0.3, 0.4), (0.1, 0, 1, 1)], stops=None)
            s.draw_path(p, Paint(shader=sh, color_filter=cf, blend=14 if k % 2 else None))
        elif k % 3 == 2:
            s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (alpha,), color_filter=cf, blur_radius=4.0, blur_style=1))
        else:
            s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (alpha,), color_filter=cf))
    return s


def scene_images_same_size(seed=91, size=256):
    """Two application images of the SAME size and different content, each created as a temporary for its draw (the
    scene player destroys the Image as soon as the Canvas call returns, so the second pixmap may sit at the first
    one's address): every draw must show its own pixels."""
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.2, 0.2, 0.25, 1.0)))
    img_a, img_b = (24, 24, seed, 0), (24, 24, seed + 5, 0)
    s.draw_image_rect(img_a, (0, 0, 24, 24), (8, 8, 120, 120), Paint(), filter=0)
    s.draw_image_rect(img_b, (0, 0, 24, 24), (136, 8, 248, 120), Paint(), filter=0)
    s.draw_image_rect(img_a, (0, 0, 24, 24), (8, 136, 120, 248), Paint(), filter=1)
    s.draw_image_rect(img_b, (0, 0, 24, 24), (136, 136, 248, 248), Paint(), filter=1)
    return s


def scene_images(seed=77, size=512):
    """Application images (Canvas::DrawImageRect and image shaders, src/render/sw/sw_canvas.cc:641-677,755-787;
    BitmapSampler, src/graphic/bitmap_sampler.cc): premultiplied and unpremultiplied pixels, nearest / bilinear /
    cubic-as-bilinear sampling, all tile modes per axis, scaled and rotated, sub-rectangles, translucent paints,
    a blend mode and a path clip."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.9, 0.9, 0.85, 1.0)))
    img_a, img_b = (37, 29, seed, 0), (16, 23, seed + 1, 1)
    s.draw_image_rect(img_a, (0, 0, 37, 29), (10, 10, 158, 126), Paint(), filter=0)
    s.draw_image_rect(img_a, (0, 0, 37, 29), (170, 10.5, 318.25, 126), Paint(fill=(0, 0, 0, 0.7)), filter=1)
    s.draw_image_rect(img_b, (2, 3, 14, 20), (330, 10, 500, 126), Paint(), filter=2)
    s.draw_image_rect(img_b, (0, 0, 16, 23), (340.5, 20.5, 356.5, 43.5), Paint(), filter=0)      # 1:1 on half pixels
    s.save()
    s.translate(100, 230)
    s.rotate(25)
    s.scale(1.3, 0.8)
    s.draw_image_rect(img_a, (5, 4, 30, 25), (-60, -50, 70, 60), Paint(fill=(0, 0, 0, 0.9), blend=14), filter=1)
    s.restore()
    k = 0
    for tx in (CLAMP, REPEAT, MIRROR, DECAL):
        for flt in (0, 1):
            x0, y0 = 200 + (k % 4) * 78, 150 + (k // 4) * 90
            sh = dict(type=5, p=(img_a[0], img_a[1], img_a[2], img_a[3]) if k % 2 else (img_b[0], img_b[1], img_b[2], img_b[3]),
                      tile=tx, tile_y=(tx + 1) % 4, filter=flt, colors=[(0, 0, 0, 1), (1, 1, 1, 1)], stops=None,
                      local=(1.5, 0.2, x0 + 10.0, -0.1, 1.2, y0 + 8.0))
            s.draw_rect(x0, y0, x0 + 72, y0 + 84, Paint(shader=sh))
            k += 1
    s.save()
    s.clip_path(star_path_small(90.0).translate(110, 410) if hasattr(PathData, "translate") else _random_closed_path(rng, 110, 410, 200.0, 1))
    sh = dict(type=5, p=img_a, tile=MIRROR, tile_y=REPEAT, filter=1, colors=[(0, 0, 0, 1), (1, 1, 1, 1)], stops=None,
              local=(2.0, 0, 0, 0, 2.0, 0))
    s.draw_rect(0, 300, 230, 512, Paint(shader=sh, fill=(0, 0, 0, 0.8)))
    s.restore()
    sh = dict(type=5, p=img_b, tile=REPEAT, tile_y=MIRROR, filter=0, colors=[(0, 0, 0, 1), (1, 1, 1, 1)], stops=None,
              local=(3.0, 0.5, 250.0, -0.5, 3.0, 340.0))
    s.draw_path(_random_closed_path(rng, 380, 420, 230.0, 2), Paint(shader=sh))
    s.draw_path(_random_closed_path(rng, 380, 420, 200.0, 3), Paint(style=STROKE, shader=sh, stroke=(0, 0, 0, 0.5), stroke_width=9.0))
    return s


def scene_clipped_blends(seed=88, size=400):
    """Blend modes that act on zero-coverage pixels (kClear, kSrc, kSrcIn, kDstIn, kSrcOut, kDstATop, kModulate,
    kSoftLight) and colour filters on draws UNDER nested path clips: the clipped sub-spans of coverage 0 still
    blend (FindSpan keeps min(clip, span) = 0, sw_canvas.cc:219-265)."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.8, 0.85, 0.6, 1.0)))
    for i in range(6):
        p = _random_closed_path(rng, rng.uniform(0, size), rng.uniform(0, size), 240.0, i)
        s.draw_path(p, Paint(fill=tuple(rng.uniform(0, 1, 3)) + (rng.uniform(0.4, 1.0),)))
    modes = [0, 1, 5, 6, 7, 10, 13, 21, 3, 14]
    s.save()
    s.clip_path(_random_closed_path(rng, size * 0.5, size * 0.5, size * 0.95, 1))
    for k, mode in enumerate(modes):
        if k == 5:
            s.save()
            s.clip_path(_random_closed_path(rng, size * 0.55, size * 0.5, size * 0.7, 2))
        p = _random_closed_path(rng, rng.uniform(0.2, 0.8) * size, rng.uniform(0.2, 0.8) * size, 170.0, k)
        cf = None
        if k % 3 == 2:
            cf = dict(type=1, color=0x9020C040, mode=1 if k % 2 else 5)
        elif k % 3 == 1:
            cf = dict(type=3)
        col = tuple(rng.uniform(0, 1, 3)) + ((1.0,) if k % 2 else (float(rng.uniform(0.4, 0.9)),))
        s.draw_path(p, Paint(fill=col, blend=mode, color_filter=cf))
        if k % 4 == 3:
            s.draw_path(p, Paint(style=STROKE, stroke=col[::-1][1:] + (0.8,), stroke_width=5.0, blend=mode, color_filter=cf))
    s.restore()
    s.restore()
    return s


def scene_filtered_layers(seed=99, size=384):
    """SaveLayer whose paint carries a mask filter or an image filter: the layer is drawn back through HandleFilter
    (resampled into a temporary, blurred / shadowed, composited), src/render/sw/sw_canvas.cc:891-902,365-369."""
    rng = np.random.RandomState(seed)
    s = Scene(size, size)
    s.draw_rect(0, 0, size, size, Paint(fill=(0.93, 0.93, 0.9, 1.0)))
    s.save_layer(30.5, 20.25, 200, 180, Paint(fill=(0, 0, 0, 0.8), blur_radius=5.0, blur_style=1))
    s.draw_rect(50, 40, 170, 150, Paint(fill=(0.9, 0.2, 0.1, 1.0)))
    s.draw_path(_random_closed_path(rng, 120, 100, 130.0, 1), Paint(fill=(0.1, 0.6, 0.9, 0.9)))
    s.restore()
    s.save()
    s.translate(280, 110)
    s.rotate(-15)
    s.save_layer(-80, -70, 90, 80, Paint(image_filter=dict(type=2, offset=(7.0, 6.0), sigma=(3.0, 3.0), color=0xA0000000)))
    s.draw_path(_random_closed_path(rng, 0, 0, 140.0, 2), Paint(fill=(0.2, 0.8, 0.3, 1.0)))
    s.draw_rect(-40, -30, 30, 20, Paint(style=STROKE, stroke=(0.1, 0.1, 0.5, 1.0), stroke_width=6.0))
    s.restore()
    s.restore()
    s.save_layer(60, 220, 340, 370, Paint(fill=(0, 0, 0, 0.9), blur_radius=3.0, blur_style=2, blend=14))
    s.draw_path(_random_closed_path(rng, 200, 295, 190.0, 3), Paint(fill=(0.8, 0.7, 0.1, 1.0)))
    s.save_layer(120, 240, 300, 350, Paint(image_filter=dict(type=1, sigma=(2.5, 4.0))))
    s.draw_rect(140, 260, 280, 330, Paint(fill=(0.5, 0.1, 0.7, 0.8)))
    s.restore()
    s.restore()
    return s


def scene_difference_clips(seed, mode="single", size=None):
    """Solid (and a few stroked / translucent) draws under ClipOp::kDifference clips (sw_canvas.cc:56-133,158-217).
    mode "single": every Save level holds ONE difference clip (path, rounded shape or rotated rectangle) — how the
    reference's own goldens and examples use the op; Save levels nest, so an inner difference clip lands on the outer
    one (PerformMerge).  "flat": the same without nesting, sometimes behind an axis-aligned intersecting ClipRect (which
    the software backend keeps as clip bounds, not as a span list); "refined": a difference clip followed by one or two
    intersecting clips in the same Save level; "carved": a difference clip applied while intersecting path clips are in
    force — the forms the CUDA backend implements; "mixed": difference and intersect clips nested in one Save level
    (difference after intersect, intersect after difference); "merge": difference on difference (PerformMerge)."""
    rng = np.random.RandomState(seed)
    w = h = size or int(rng.randint(60, 520))
    if size is None:
        h = int(rng.randint(60, 520))
    s = Scene(w, h)

    def clip_shape(intersect):
        k = rng.uniform()
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        if k < 0.6:
            s.clip_path(_random_closed_path(rng, cx, cy, float(rng.uniform(40, 400)), int(rng.randint(0, 6))), intersect)
        elif k < 0.8:
            s.clip_rect(float(cx), float(cy), float(cx + rng.uniform(10, 300)), float(cy + rng.uniform(10, 300)), intersect)
        else:
            s.rotate(float(rng.uniform(-40, 40)))
            s.clip_rect(float(cx), float(cy), float(cx + rng.uniform(10, 300)), float(cy + rng.uniform(10, 300)), intersect)

    depth = 0
    for i in range(int(rng.randint(4, 40))):
        if rng.uniform() < 0.3 and depth < (1 if mode in ("flat", "refined", "carved") else 2):
            s.save(); depth += 1
            if mode == "carved":    # a difference clip applied while intersecting path clips are in force (and clips after it)
                clip_shape(True)
                if rng.uniform() < 0.3:
                    clip_shape(True)
                clip_shape(False)
                if rng.uniform() < 0.4:
                    clip_shape(True)
            elif mode == "refined":   # a difference clip refined by intersecting clips in the same Save level
                clip_shape(False)
                clip_shape(True)
                if rng.uniform() < 0.4:
                    clip_shape(True)
            elif mode == "flat":
                if rng.uniform() < 0.4:
                    x, y = rng.uniform(0, w * 0.6), rng.uniform(0, h * 0.6)
                    s.clip_rect(float(x), float(y), float(x + rng.uniform(40, 400)), float(y + rng.uniform(40, 400)), True)
                clip_shape(False)
            elif mode == "single":
                clip_shape(False)
            elif mode == "mixed":
                first = rng.uniform() < 0.5
                clip_shape(first)
                clip_shape(not first)
                if rng.uniform() < 0.3:
                    clip_shape(True)
            else:
                clip_shape(False)
                clip_shape(False)
        elif rng.uniform() < 0.15 and depth > 0:
            s.restore(); depth -= 1
        col = tuple(np.float32(v) for v in rng.uniform(0, 1, 4))
        if rng.uniform() < 0.5:
            col = col[:3] + (np.float32(1.0),)
        s.draw_path(_random_closed_path(rng, rng.uniform(0, w), rng.uniform(0, h), float(rng.uniform(20, 400)), i),
                    Paint(style=int(rng.randint(0, 3)), fill=col, stroke=col, stroke_width=float(rng.uniform(0.5, 9))))
    while depth:
        s.restore(); depth -= 1
    return s


def scene_ref_clip_path_difference():
    """The reference's own golden case ClipGolden.ClipPathDifference (test/golden/cases/clip/clip.cc:171-203): the star
    stroked green, a two-quad clip path stroked red, ClipPath(kDifference), the star filled blue; 400x400."""
    s = Scene(400, 400)
    star = star_path()
    green, red, blue = (0.0, 1.0, 0.0, 1.0), (1.0, 0.0, 0.0, 1.0), (0.0, 0.0, 1.0, 1.0)
    s.draw_path(star, Paint(style=STROKE, stroke=green, stroke_width=1.0))
    clip = PathData().move_to(10.0, 10.0).quad_to(300.0, 10.0, 150.0, 150.0).quad_to(10.0, 300.0, 300.0, 300.0).close()
    s.draw_path(clip, Paint(style=STROKE, stroke=red, stroke_width=1.0))
    s.clip_path(clip, False)
    s.draw_path(star, Paint(style=FILL, fill=blue))
    return s
