/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * skb_oracle.c — plain-C CPU restatement ("port") of the reference's software
 * raster path, operating on the flat display list of include/skb_dl.h.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load it.  It is PINNED: tests/test_oracle_pinning.py checks it bit-for-bit
 * against the reference's own compiled backend (oracle/_ref, built from the
 * unmodified sources by oracle/build_ref.py) on spans, pixels and blur, and
 * against the golden vectors committed under tests/golden/.  That includes ClipOp::kDifference (spans_subtract, the
 * merge in clip_refine), whose result in the reference depends on how std::sort orders equal keys: the sorts are
 * restated as libstdc++'s algorithm (sort_spans), and 350 random difference-clip scenes (single, nested with
 * intersect, difference on difference) equal the compiled reference's bit for bit.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference root).  Arithmetic notes that matter for bit-exactness:
 *   - all fixed-point maths is int32 two's-complement with WRAPPING shifts, as
 *     g++ compiles the reference (sw_subpixel.hpp:39-69);
 *   - float maths is IEEE fp32 with one rounding per operation (the reference
 *     is built for baseline x86-64: no FMA contraction) — compile this file
 *     with -ffp-contract=off and without -ffast-math / -march=native.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/skb_dl.h"

typedef int32_t fx; /* SWFixed 16.16 / SWFDot6 26.6 (sw_subpixel.hpp:19-20) */
#define FX1 (1 << 16)
#define FX_MAX 0x7FFFFFFF
#define FX_MIN (-0x7FFFFFFF)

/* ------------------------------------------------------------------ helpers */
static inline fx shl(fx v, int s) { return (fx)((uint32_t)v << s); }       /* sw_subpixel.hpp:45-47 */
static inline fx fx_mul(fx a, fx b) { return (fx)(((int64_t)a * b) >> 16); } /* :39-41 */
static inline fx fx_div(fx n, fx d) {                                        /* :64-69 SWFixedDiv/SWFDot6Div */
  int64_t q = (int64_t)((uint64_t)(int64_t)n << 16) / d;
  if (q < FX_MIN) q = FX_MIN;
  if (q > FX_MAX) q = FX_MAX;
  return (fx)q;
}
static inline fx snap_y(fx y) { /* sw_edge.hpp:36-41 — quarter-pixel round-to-nearest */
  return (fx)((((uint32_t)y + (FX1 >> 3)) >> 14) << 14);
}
static inline int fx_floor_i(fx x) { return x >> 16; }
static inline int fx_ceil_i(fx x) { return (fx)((uint32_t)x + FX1 - 1) >> 16; }
static inline int fx_round_i(fx x) { return (fx)((uint32_t)x + (FX1 >> 1)) >> 16; }
static inline fx fx_round_fx(fx x) { return (fx)(((uint32_t)x + (FX1 >> 1)) & 0xFFFF0000u); }
static inline fx fx_ceil_fx(fx x) { return (fx)(((uint32_t)x + FX1 - 1) & 0xFFFF0000u); }
static inline fx fx_floor_fx(fx x) { return (fx)((uint32_t)x & 0xFFFF0000u); }
static inline fx i_to_fx(int n) { return (fx)((uint32_t)n << 16); }
static inline fx fx_abs(fx v) { return v < 0 ? (fx)(0u - (uint32_t)v) : v; }
static inline fx fx_add(fx a, fx b) { return (fx)((uint32_t)a + (uint32_t)b); }
static inline fx fx_sub(fx a, fx b) { return (fx)((uint32_t)a - (uint32_t)b); }
static int clz32(uint32_t x) { /* src/geometry/math.hpp:61-90 */
  int n = 0;
  if (x == 0) return 32;
  while (!(x & 0x80000000u)) {
    x <<= 1;
    n++;
  }
  return n;
}
/* float -> int as the reference's static_cast<int>/(SWFDot6) does on x86-64 (cvttss2si) */
static inline int32_t f2i(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT32_MIN;
  return (int32_t)v;
}

/* ------------------------------------------------------------------- edges */
typedef struct edge {
  int prev, next; /* indices; -1 = none */
  fx x, y, dx, dy, upper_x, upper_y, lower_y;
  int curve_count; /* int8 in the reference; values 0..64 */
  int curve_shift;
  int winding;
  /* quadratic forward-difference state (SWQuadEdge, sw_edge.hpp:65-81) */
  fx qx, qy, qdx, qdy, qddx, qddy, q_first_y, q_last_x, q_last_y, snapped_x, snapped_y;
} edge;

/* SWEdge::UpdateLine — sw_edge.cc:44-70 */
static int update_line(edge* e, fx x0, fx y0, fx x1, fx y1, fx slope) {
  if (y0 > y1) {
    fx t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
    e->winding = -e->winding;
  }
  fx x0x1 = fx_sub(x1, x0) >> 10;
  fx y0y1 = fx_sub(y1, y0) >> 10;
  if (y0y1 == 0) return 0;
  e->x = x0;
  e->y = y0;
  e->dx = slope;
  e->dy = (x0x1 == 0 || slope == 0) ? FX_MAX : fx_abs(fx_div(y0y1, x0x1));
  e->upper_x = x0;
  e->upper_y = y0;
  e->lower_y = y1;
  return 1;
}

/* SWEdge::SetLine — sw_edge.cc:18-42.  trunc(v*4*64) << 10 >> 2 with int32 wrap.
 * g_wide: the backend's wide-coordinate mode (include/skb.h SKB_COORD_WIDE) — the same 24.8 -> 16.16 conversion
 * without the int32 overflow of SWFDot6ToFixed (sw_subpixel.hpp:43), i.e. `<< 8`.  The ONE deliberate deviation from
 * the reference in this file; identical to it wherever the reference does not wrap (|coordinate| < 8192 px).  The
 * compiled reference cannot serve as the checker there: it has no such mode, and rendering a large canvas as
 * translated windows is not the same function either (ChopQuadAtYExtrema interpolates in device space in fp32,
 * geometry.cc:323-349, so the reference's own output changes by up to a quarter-pixel snap under a whole-pixel
 * translation — measured by tests/test_oracle_pinning.py::test_reference_is_not_translation_invariant). */
static int g_wide = 0;
static inline fx fx_from_24_8(fx t) { return g_wide ? shl(t, 8) : (shl(t, 10) >> 2); }
static inline fx line_coord(float v) { return fx_from_24_8(f2i((v * 4) * 64)); }
static int set_line(edge* e, float x0f, float y0f, float x1f, float y1f) {
  fx x0 = line_coord(x0f), y0 = snap_y(line_coord(y0f));
  fx x1 = line_coord(x1f), y1 = snap_y(line_coord(y1f));
  e->winding = 1;
  fx y0y1 = fx_sub(y1, y0) >> 10;
  if (y0y1 == 0) return 0;
  fx x0x1 = fx_sub(x1, x0) >> 10;
  fx slope = fx_div(x0x1, y0y1);
  e->curve_count = 0;
  e->curve_shift = 0;
  return update_line(e, x0, y0, x1, y1, slope);
}

/* SWQuadEdge::UpdateQuad — sw_edge.cc:233-292 */
static int update_quad(edge* e) {
  int success = 0;
  int count = e->curve_count;
  fx oldx = e->qx, oldy = e->qy, dx = e->qdx, dy = e->qdy;
  fx newx, newy, nsx, nsy;
  int shift = e->curve_shift;
  do {
    fx slope;
    if (--count > 0) {
      newx = fx_add(oldx, dx >> shift);
      newy = fx_add(oldy, dy >> shift);
      if (fx_abs(dy >> shift) >= FX1 * 2) {
        fx diffy = fx_sub(newy, e->snapped_y) >> 10;
        slope = diffy ? fx_div(fx_sub(newx, e->snapped_x) >> 10, diffy) : FX_MAX;
        fx r = fx_round_fx(newy);
        nsy = e->q_last_y < r ? e->q_last_y : r;
        nsx = fx_sub(newx, fx_mul(slope, fx_sub(newy, nsy)));
      } else {
        fx s = snap_y(newy);
        nsy = e->q_last_y < s ? e->q_last_y : s;
        nsx = newx;
        fx diffy = fx_sub(nsy, e->snapped_y) >> 10;
        slope = diffy ? fx_div(fx_sub(newx, e->snapped_x) >> 10, diffy) : FX_MAX;
      }
      dx = fx_add(dx, e->qddx);
      dy = fx_add(dy, e->qddy);
    } else {
      newx = e->q_last_x;
      newy = e->q_last_y;
      nsy = newy;
      nsx = newx;
      fx diffy = fx_sub(newy, e->snapped_y) >> 10;
      slope = diffy ? fx_div(fx_sub(newx, e->snapped_x) >> 10, diffy) : FX_MAX;
    }
    if (slope < FX_MAX) success = update_line(e, e->snapped_x, e->snapped_y, nsx, nsy, slope);
    oldx = newx;
    oldy = newy;
  } while (count > 0 && !success);
  e->qx = newx;
  e->qy = newy;
  e->qdx = dx;
  e->qdy = dy;
  e->snapped_x = nsx;
  e->snapped_y = nsy;
  e->curve_count = count;
  return success;
}

/* cheap_distance / diff_to_shift — sw_edge.cc:97-119 (shiftAA = kDefaultAccuracy = 2) */
static int diff_to_shift(fx dx, fx dy) {
  dx = fx_abs(dx);
  dy = fx_abs(dy);
  fx dist = (dx > dy ? dx : dy) + ((dx < dy ? dx : dy) >> 1);
  dist = fx_add(dist, 1 << 4) >> 5;
  return (32 - clz32((uint32_t)dist)) >> 1;
}

/* SWQuadEdge::SetQuad — sw_edge.cc:121-231.  pts = x0 y0 x1 y1 x2 y2 (y-monotone). */
static int set_quad(edge* e, const float* p) {
  const float scale = 256.0f;
  fx x0 = f2i(p[0] * scale), y0 = f2i(p[1] * scale);
  fx x1 = f2i(p[2] * scale), y1 = f2i(p[3] * scale);
  fx x2 = f2i(p[4] * scale), y2 = f2i(p[5] * scale);
  int w = 1;
  if (y0 > y2) {
    fx t = x0; x0 = x2; x2 = t;
    t = y0; y0 = y2; y2 = t;
    w = -1;
  }
  int top = fx_add(y0, 32) >> 6, bottom = fx_add(y2, 32) >> 6;
  if (top == bottom) return 0;
  fx ddx = fx_sub(fx_sub(shl(x1, 1), x0), x2) >> 2;
  fx ddy = fx_sub(fx_sub(shl(y1, 1), y0), y2) >> 2;
  int shift = diff_to_shift(ddx, ddy);
  if (shift == 0) shift = 1;
  else if (shift > 6) shift = 6;
  e->winding = w;
  e->curve_count = 1 << shift;
  e->curve_shift = shift - 1;
  fx A = shl(fx_add(fx_sub(fx_sub(x0, x1), x1), x2), 9);
  fx B = shl(fx_sub(x1, x0), 10);
  e->qx = shl(x0, 10);
  e->qdx = fx_add(B, A >> shift);
  e->qddx = A >> (shift - 1);
  A = shl(fx_add(fx_sub(fx_sub(y0, y1), y1), y2), 9);
  B = shl(fx_sub(y1, y0), 10);
  e->qy = shl(y0, 10);
  e->qdy = fx_add(B, A >> shift);
  e->qddy = A >> (shift - 1);
  e->q_last_x = shl(x2, 10);
  e->q_last_y = shl(y2, 10);
  e->qx >>= 2; e->qy >>= 2; e->qdx >>= 2; e->qdy >>= 2;
  e->qddx >>= 2; e->qddy >>= 2; e->q_last_x >>= 2; e->q_last_y >>= 2;
  if (g_wide) {  /* the end points without the overflow of `<< 10` (see line_coord) */
    e->qx = shl(x0, 8); e->qy = shl(y0, 8); e->q_last_x = shl(x2, 8); e->q_last_y = shl(y2, 8);
  }
  e->qy = snap_y(e->qy);
  e->q_last_y = snap_y(e->q_last_y);
  e->q_first_y = e->qy;
  e->snapped_x = e->qx;
  e->snapped_y = e->qy;
  /* fields UpdateLine may leave untouched if every chord is degenerate */
  e->x = e->y = e->dx = e->dy = e->upper_x = e->upper_y = e->lower_y = 0;
  update_quad(e);
  return 1;
}

/* SWEdge::CanBeIgnored — sw_edge.cc:72-88 */
static int can_be_ignored(float top, float bottom, fx y0, fx y1) {
  fx start_y = snap_y(line_coord(top));
  fx stop_y = snap_y(line_coord(bottom));
  return (y0 >= stop_y || y1 <= start_y);
}

/* -------------------------------------------------- path lowering (floats) */
typedef struct { float x, y; } v2;
static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }

/* Matrix * Vec4 with z=0,w=1 in glm's operation order:
 * (m0*x + m1*y) + (m2*0 + m3*1)  — src/geometry/matrix.cc:463, Path::CopyWithMatrix path.cc:1274-1296 */
static inline v2 xform(const float* m, v2 p) {
  float x = (m[0] * p.x + m[1] * p.y) + (0.0f * 0.0f + m[2] * 1.0f);
  float y = (m[3] * p.x + m[4] * p.y) + (0.0f * 0.0f + m[5] * 1.0f);
  return V(x, y);
}

typedef struct prim { int n; v2 p[3]; } prim; /* n = 2 line, 3 quad — already transformed */
typedef struct primbuf { prim* v; size_t n, cap; float l, t, r, b; int have; } primbuf;

static void pb_bound(primbuf* pb, v2 p) {
  if (!pb->have) { pb->l = pb->r = p.x; pb->t = pb->b = p.y; pb->have = 1; return; }
  if (p.x < pb->l) pb->l = p.x;
  if (p.x > pb->r) pb->r = p.x;
  if (p.y < pb->t) pb->t = p.y;
  if (p.y > pb->b) pb->b = p.y;
}
static void pb_push(primbuf* pb, int n, v2 a, v2 b, v2 c) {
  if (pb->n == pb->cap) {
    pb->cap = pb->cap ? pb->cap * 2 : 64;
    pb->v = (prim*)realloc(pb->v, pb->cap * sizeof(prim));
  }
  prim* q = &pb->v[pb->n++];
  q->n = n; q->p[0] = a; q->p[1] = b; q->p[2] = c;
}

/* CubicCoeff — src/geometry/geometry.cc:135-169 */
typedef struct { v2 A, B, C, D; } cubic_coeff;
static cubic_coeff cubic_coeff_make(v2 p0, v2 p1, v2 p2, v2 p3) {
  cubic_coeff c;
  c.A = V((p3.x + 3.0f * (p1.x - p2.x)) - p0.x, (p3.y + 3.0f * (p1.y - p2.y)) - p0.y);
  c.B = V(3.0f * ((p2.x - (p1.x + p1.x)) + p0.x), 3.0f * ((p2.y - (p1.y + p1.y)) + p0.y));
  c.C = V(3.0f * (p1.x - p0.x), 3.0f * (p1.y - p0.y));
  c.D = p0;
  return c;
}
static v2 cubic_eval(const cubic_coeff* c, float t) {
  return V(((c->A.x * t + c->B.x) * t + c->C.x) * t + c->D.x, ((c->A.y * t + c->B.y) * t + c->C.y) * t + c->D.y);
}
/* QuadCoeff of the tangent control polygon — geometry.cc:44-67, cubic.cc:13-27 */
typedef struct { v2 A, B, C; } quad_coeff;
static quad_coeff quad_coeff_make(v2 q0, v2 q1, v2 q2) {
  quad_coeff c;
  c.C = q0;
  c.B = V((q1.x - q0.x) + (q1.x - q0.x), (q1.y - q0.y) + (q1.y - q0.y));
  c.A = V((q2.x - (q1.x + q1.x)) + q0.x, (q2.y - (q1.y + q1.y)) + q0.y);
  return c;
}
static v2 quad_eval(const quad_coeff* c, float t) {
  return V((c->A.x * t + c->B.x) * t + c->C.x, (c->A.y * t + c->B.y) * t + c->C.y);
}

/* Cubic::ToQuads quad count — src/geometry/cubic.cc:29-38 */
static int cubic_quad_count(v2 p1, v2 c1, v2 c2, v2 p2) {
  float accuracy = 0.1f;
  double max_hypot2 = 432.0 * accuracy * accuracy;
  v2 a = V(c1.x * 3.0f - p1.x, c1.y * 3.0f - p1.y);
  v2 b = V(c2.x * 3.0f - p2.x, c2.y * 3.0f - p2.y);
  v2 p = V(b.x - a.x, b.y - a.y);
  float err = p.x * p.x + p.y * p.y;
  double n = ceil(pow(err / max_hypot2, 1. / 6.0));
  if (!(n > 1.)) n = 1.;
  if (n > 1e6) n = 1e6; /* guard only; the reference would allocate unboundedly */
  return (int)n;
}

/* Conic::Chop + subdivided(level 1) — src/geometry/conic.cc:26-65,169-199 */
static int between(float a, float b, float c) { return (a - b) * (c - b) <= 0; }
static void conic_to_quads(v2 p0, v2 p1, v2 p2, float w, v2 out[5]) {
  float scale = 1.0f / (w + 1.0f);
  v2 wp1 = V(w * p1.x, w * p1.y);
  v2 m = V(((p0.x + (wp1.x + wp1.x)) + p2.x) * scale * 0.5f, ((p0.y + (wp1.y + wp1.y)) + p2.y) * scale * 0.5f);
  if (!(isfinite(m.x) && isfinite(m.y))) {
    double w_d = w, w_2 = w_d * 2, scale_half = 1 / (1 + w_d) * 0.5;
    m.x = (float)((p0.x + w_2 * p1.x + p2.x) * scale_half);
    m.y = (float)((p0.y + w_2 * p1.y + p2.y) * scale_half);
  }
  v2 d0p1 = V((p0.x + wp1.x) * scale, (p0.y + wp1.y) * scale);
  v2 d1p1 = V((wp1.x + p2.x) * scale, (wp1.y + p2.y) * scale);
  v2 d0p2 = m, d1p0 = m;
  float startY = p0.y, endY = p2.y;
  if (between(startY, p1.y, endY)) {
    float midY = d0p2.y;
    if (!between(startY, midY, endY)) {
      float closerY = fabsf(midY - startY) < fabsf(midY - endY) ? startY : endY;
      d0p2.y = d1p0.y = closerY;
    }
    if (!between(startY, d0p1.y, d0p2.y)) d0p1.y = startY;
    if (!between(d1p0.y, d1p1.y, endY)) d1p1.y = endY;
  }
  out[0] = p0; out[1] = d0p1; out[2] = d0p2; out[3] = d1p1; out[4] = p2;
  /* PointAreFinite fallback — conic.cc:318-324, point_priv.hpp:27-35 */
  float prod = 0;
  for (int i = 0; i < 5; i++) prod *= (out[i].x * out[i].y);
  if (!(prod == 0)) {
    for (int i = 1; i < 4; i++) out[i] = p1;
  }
}

/* Lower one path: Stroke::QuadPath (stroke.cc:914-962) + CopyWithMatrix + bounds
 * (Path::ComputePtBounds path.cc:1347-1360, Rect::SetBoundsCheck rect.cc:12-50). */
static void lower_path(const skb_dl_seg* segs, uint32_t n, const float* ctm, primbuf* pb) {
  v2 prev_cubic_end = V(0, 0);
  for (uint32_t i = 0; i < n; i++) {
    const skb_dl_seg* s = &segs[i];
    uint32_t type = s->type_flags & SKB_SEG_TYPE_MASK;
    v2 start = (s->type_flags & SKB_SEG_P0_FROM_PREV_CUBIC) ? prev_cubic_end : V(s->start[0], s->start[1]);
    v2 p0 = V(s->p[0], s->p[1]), p1 = V(s->p[2], s->p[3]), p2 = V(s->p[4], s->p[5]), p3 = V(s->p[6], s->p[7]);
    v2 ts = xform(ctm, start);
    pb_bound(pb, ts);
    switch (type) {
      case SKB_SEG_POINT:
        break;
      case SKB_SEG_LINE:
      case SKB_SEG_CLOSE: {
        v2 e = xform(ctm, p1);
        pb_bound(pb, e);
        pb_push(pb, 2, ts, e, e);
      } break;
      case SKB_SEG_QUAD: {
        v2 c = xform(ctm, p1), e = xform(ctm, p2);
        pb_bound(pb, c);
        pb_bound(pb, e);
        pb_push(pb, 3, ts, c, e);
      } break;
      case SKB_SEG_CONIC: {
        v2 q[5];
        conic_to_quads(p0, p1, p2, s->w, q);
        v2 t1 = xform(ctm, q[1]), t2 = xform(ctm, q[2]), t3 = xform(ctm, q[3]), t4 = xform(ctm, q[4]);
        pb_bound(pb, t1); pb_bound(pb, t2); pb_bound(pb, t3); pb_bound(pb, t4);
        pb_push(pb, 3, ts, t1, t2);
        pb_push(pb, 3, t2, t3, t4);
      } break;
      case SKB_SEG_CUBIC: {
        int cnt = cubic_quad_count(p0, p1, p2, p3);
        double quad_count = (double)cnt;
        cubic_coeff cc = cubic_coeff_make(p0, p1, p2, p3);
        quad_coeff qc = quad_coeff_make(V(3.f * (p1.x - p0.x), 3.f * (p1.y - p0.y)),
                                        V(3.f * (p2.x - p1.x), 3.f * (p2.y - p1.y)),
                                        V(3.f * (p3.x - p2.x), 3.f * (p3.y - p2.y)));
        v2 cur = ts;
        v2 last = start;
        for (int k = 0; k < cnt; k++) {
          float t0 = (float)((double)k / quad_count), t1 = (float)((double)(k + 1) / quad_count);
          v2 a = cubic_eval(&cc, t0), b = cubic_eval(&cc, t1); /* Subsegment — cubic.cc:54-61 */
          float sc = (t1 - t0) * (1.f / 3.f);
          v2 ta = quad_eval(&qc, t0), tb = quad_eval(&qc, t1);
          v2 c1 = V(a.x + ta.x * sc, a.y + ta.y * sc);
          v2 c2 = V(b.x - tb.x * sc, b.y - tb.y * sc);
          v2 ctrl = V(((c1.x * 3.f - a.x) + (c2.x * 3.f - b.x)) / 4.f, ((c1.y * 3.f - a.y) + (c2.y * 3.f - b.y)) / 4.f);
          v2 tc = xform(ctm, ctrl), te = xform(ctm, b);
          pb_bound(pb, tc);
          pb_bound(pb, te);
          pb_push(pb, 3, cur, tc, te);
          cur = te;
          last = b;
        }
        prev_cubic_end = last;
      } break;
      default:
        break;
    }
  }
}

/* ChopQuadAtYExtrema — src/geometry/geometry.cc:311-349.  Returns number of output quads (1 or 2). */
static int chop_quad_y(const v2 src[3], v2 dst[5]) {
  float a = src[0].y, b = src[1].y, c = src[2].y;
  float ab = a - b, bc = b - c;
  if (ab < 0) bc = -bc;
  if (ab == 0 || bc < 0) {
    float number = a - b, denom = a - b - b + c; /* valid_unit_divide — geometry.hpp:102-123 */
    if (number < 0) { number = -number; denom = -denom; }
    if (!(denom == 0 || number == 0 || number >= denom)) {
      float r = number / denom;
      if (!isnan(r) && r != 0) {
        /* QuadCoeff::ChopQuadAt — geometry.cc:117-133 ; Interp = v0 + (v1-v0)*t */
        v2 p01 = V(src[0].x + (src[1].x - src[0].x) * r, src[0].y + (src[1].y - src[0].y) * r);
        v2 p12 = V(src[1].x + (src[2].x - src[1].x) * r, src[1].y + (src[2].y - src[1].y) * r);
        dst[0] = src[0];
        dst[1] = p01;
        dst[2] = V(p01.x + (p12.x - p01.x) * r, p01.y + (p12.y - p01.y) * r);
        dst[3] = p12;
        dst[4] = src[2];
        dst[1].y = dst[3].y = dst[2].y;
        return 2;
      }
    }
    b = fabsf(a - b) < fabsf(b - c) ? a : c;
  }
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = src[2];
  dst[1].y = b;
  return 1;
}

/* ------------------------------------------------------------ span builder */
typedef struct skbo_span { int32_t x, y, len, cover; } skbo_span;
typedef struct spanvec { skbo_span* v; size_t n, cap; } spanvec;
static void sv_push(spanvec* s, int x, int y, int len, int cover) {
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 256;
    s->v = (skbo_span*)realloc(s->v, s->cap * sizeof(skbo_span));
  }
  skbo_span* q = &s->v[s->n++];
  q->x = x; q->y = y; q->len = len; q->cover = cover;
}

/* SpanBuilder + RealSpanBuilder — sw_raster.cc:19-136, sw_raster.hpp:27-76 */
typedef struct builder {
  spanvec* out;
  float scan_top;
  int left, width;
  uint8_t* alphas;
  int32_t curr_y; /* INT32_MIN = none */
} builder;
static void real_span(builder* b, int x, int y, int w, uint8_t a) {
  if ((float)y < b->scan_top) return;
  sv_push(b->out, x, y, w, a);
}
static void real_spans(builder* b, int x, int y, const uint8_t* aa, int len) {
  if ((float)y < b->scan_top) return;
  for (int i = 0; i < len; i++) sv_push(b->out, x + i, y, 1, aa[i]);
}
static void acc_flush(builder* b) {
  if (b->curr_y == INT32_MIN) return;
  int curr = 0, n = b->width;
  while (curr < n) {
    if (b->alphas[curr] > 0) {
      int start = curr, w = 1;
      uint8_t a = b->alphas[curr];
      do {
        curr++;
        if (curr >= n) break;
        if (b->alphas[curr] == a) w++;
        else break;
      } while (b->alphas[curr]);
      real_span(b, b->left + start, b->curr_y, w, a);
    } else {
      curr++;
    }
  }
}
static void acc_row(builder* b, int y) {
  if (b->curr_y == y) return;
  if (b->curr_y == INT32_MIN) { b->curr_y = y; return; }
  acc_flush(b);
  memset(b->alphas, 0, (size_t)b->width);
  b->curr_y = y;
}
static inline void acc_add(builder* b, int x, uint8_t a) {
  int o = x - b->left;
  if (o < 0 || o >= b->width) return; /* the reference would write out of bounds here */
  unsigned s = (unsigned)b->alphas[o] + a;
  b->alphas[o] = s > 255 ? 255 : (uint8_t)s;
}
static void acc_span(builder* b, int x, int y, int w, uint8_t a) {
  if ((float)y < b->scan_top) return;
  acc_row(b, y);
  for (int i = 0; i < w; i++) acc_add(b, x + i, a);
}
static void acc_spans(builder* b, int x, int y, const uint8_t* aa, int len) {
  if ((float)y < b->scan_top) return;
  acc_row(b, y);
  for (int i = 0; i < len; i++) acc_add(b, x + i, aa[i]);
}

/* --------------------------------------------------- trapezoid row blitting */
static inline uint8_t partial_alpha_mul(uint8_t alpha, uint8_t full) { return (uint8_t)((alpha * full) >> 8); } /* :155-157 */
static inline uint8_t trapezoid_to_alpha(fx l1, fx l2) { return (uint8_t)((fx_add(l1, l2) / 2) >> 8); }         /* :265-269 */
static inline uint8_t partial_triangle_to_alpha(fx a, fx b) {                                                   /* :272-278 */
  fx area = (fx)((uint32_t)(a >> 11) * (uint32_t)(a >> 11) * (uint32_t)(b >> 11));
  return (uint8_t)((area >> 8) & 0xFF);
}
static void blit_single(builder* b, int y, int x, uint8_t alpha, uint8_t full, int no_real) { /* :281-289 */
  if (full == 0xFF && !no_real) real_span(b, x, y, 1, alpha);
  else acc_span(b, x, y, 1, partial_alpha_mul(alpha, full));
}
static void blit_two(builder* b, int y, int x, uint8_t a1, uint8_t a2, uint8_t full, int no_real) { /* :291-301 */
  if (full == 0xFF && !no_real) {
    real_span(b, x, y, 1, a1);
    real_span(b, x + 1, y, 1, a2);
  } else {
    acc_span(b, x, y, 1, a1);
    acc_span(b, x + 1, y, 1, a2);
  }
}
static void blit_full(builder* b, int y, int x, int len, uint8_t full, int no_real) { /* :303-310 */
  if (full == 0xFF && !no_real) real_span(b, x, y, len, full);
  else acc_span(b, x, y, len, full);
}
/* compute_alpha_above_line — :314-339 */
static void alpha_above(uint8_t* al, fx l, fx r, fx dY, uint8_t full) {
  int R = fx_ceil_i(r);
  if (R == 0) return;
  if (R == 1) {
    al[0] = partial_alpha_mul((uint8_t)(fx_sub(fx_sub(shl(R, 17), l), r) >> 9), full);
  } else {
    fx first = fx_sub(FX1, l);
    fx last = fx_sub(r, shl(R - 1, 16));
    fx firstH = fx_mul(first, dY);
    al[0] = (uint8_t)(fx_mul(first, firstH) >> 9);
    fx a16 = fx_add(firstH, dY >> 1);
    for (int i = 1; i < R - 1; ++i) {
      al[i] = (uint8_t)(a16 >> 8);
      a16 = fx_add(a16, dY);
    }
    al[R - 1] = (uint8_t)(full - partial_triangle_to_alpha(last, dY));
  }
}
/* compute_alpha_below_line — :343-368 */
static void alpha_below(uint8_t* al, fx l, fx r, fx dY, uint8_t full) {
  int R = fx_ceil_i(r);
  if (R == 0) return;
  if (R == 1) {
    al[0] = partial_alpha_mul(trapezoid_to_alpha(l, r), full);
  } else {
    fx first = fx_sub(FX1, l);
    fx last = fx_sub(r, shl(R - 1, 16));
    fx lastH = fx_mul(last, dY);
    al[R - 1] = (uint8_t)(fx_mul(last, lastH) >> 9);
    fx a16 = fx_add(lastH, dY >> 1);
    for (int i = R - 2; i > 0; i--) {
      al[i] = (uint8_t)((a16 >> 8) & 0xFF);
      a16 = fx_add(a16, dY);
    }
    al[0] = (uint8_t)(full - partial_triangle_to_alpha(first, dY));
  }
}
/* blit_aaa_trapezoid_row — :370-455 */
static void blit_aaa_row(builder* b, int y, fx ul, fx ur, fx ll, fx lr, fx lDY, fx rDY, uint8_t full, int no_real) {
  int L = fx_floor_i(ul), R = fx_ceil_i(lr);
  int len = R - L;
  if (len == 1) {
    blit_single(b, y, L, trapezoid_to_alpha(fx_sub(ur, ul), fx_sub(lr, ll)), full, no_real);
    return;
  }
  if (len <= 0) return; /* the reference has no such guard; nothing is emitted for len<=0 either */
  uint8_t* al = (uint8_t*)malloc((size_t)(len + 1) * 2);
  uint8_t* tmp = al + len + 1;
  memset(al, full, (size_t)len);
  memset(tmp, 0, (size_t)len + 1);
  int uL = fx_floor_i(ul), lL = fx_ceil_i(ll);
  if (uL + 2 == lL) {
    fx first = fx_sub(fx_add(i_to_fx(uL), FX1), ul);
    fx second = fx_sub(fx_sub(ll, ul), first);
    uint8_t a1 = (uint8_t)(full - partial_triangle_to_alpha(first, lDY));
    uint8_t a2 = partial_triangle_to_alpha(second, lDY);
    al[0] = al[0] > a1 ? al[0] - a1 : 0;
    al[1] = al[1] > a2 ? al[1] - a2 : 0;
  } else {
    alpha_below(tmp + uL - L, fx_sub(ul, i_to_fx(uL)), fx_sub(ll, i_to_fx(uL)), lDY, full);
    for (int i = uL; i < lL; ++i) {
      if (al[i - L] > tmp[i - L]) al[i - L] -= tmp[i - L];
      else al[i - L] = 0;
    }
  }
  int uR = fx_floor_i(ur), lR = fx_ceil_i(lr);
  if (uR + 2 == lR) {
    fx first = fx_sub(fx_add(i_to_fx(uR), FX1), ur);
    fx second = fx_sub(fx_sub(lr, ur), first);
    uint8_t a1 = partial_triangle_to_alpha(first, rDY);
    uint8_t a2 = (uint8_t)(full - partial_triangle_to_alpha(second, rDY));
    al[len - 2] = al[len - 2] > a1 ? al[len - 2] - a1 : 0;
    al[len - 1] = al[len - 1] > a2 ? al[len - 1] - a2 : 0;
  } else {
    alpha_above(tmp + uR - L, fx_sub(ur, i_to_fx(uR)), fx_sub(lr, i_to_fx(uR)), rDY, full);
    for (int i = uR; i < lR; ++i) {
      if (al[i - L] > tmp[i - L]) al[i - L] -= tmp[i - L];
      else al[i - L] = 0;
    }
  }
  if (full == 0xFF && !no_real) real_spans(b, L, y, al, len);
  else acc_spans(b, L, y, al, len);
  free(al);
}
/* blit_trapezoid_row — :457-544 */
static void blit_trapezoid_row(builder* b, int y, fx ul, fx ur, fx ll, fx lr, fx lDY, fx rDY, uint8_t full, int no_real) {
  if (ul > ur) return;
  if (ll > lr) { /* approximate_intersection — :253-262 */
    fx l1 = ul, r1 = ll, l2 = ur, r2 = lr;
    if (l1 > r1) { fx t = l1; l1 = r1; r1 = t; }
    if (l2 > r2) { fx t = l2; l2 = r2; r2 = t; }
    /* 64-bit sum: equal to the reference's int32 sum wherever that does not overflow (every coordinate below 8192 px);
     * at the right clip of a 16384-px canvas (wide mode) the int32 sum would reach 2^31 and flip sign */
    ll = lr = (fx)(((int64_t)(l1 > l2 ? l1 : l2) + (int64_t)(r1 < r2 ? r1 : r2)) / 2);
  }
  if (ul == ur && ll == lr) return;
  if (ul > ll) { fx t = ul; ul = ll; ll = t; }
  if (ur > lr) { fx t = ur; ur = lr; lr = t; }
  fx joinLeft = fx_ceil_fx(ll);
  fx joinRite = fx_floor_fx(ur);
  if (joinLeft <= joinRite) {
    if (ul < joinLeft) {
      int len = fx_ceil_i(fx_sub(joinLeft, ul));
      if (len == 1) {
        blit_single(b, y, ul >> 16, trapezoid_to_alpha(fx_sub(joinLeft, ul), fx_sub(joinLeft, ll)), full, no_real);
      } else if (len == 2) {
        fx first = fx_sub(fx_sub(joinLeft, FX1), ul);
        fx second = fx_sub(fx_sub(ll, ul), first);
        uint8_t a1 = partial_triangle_to_alpha(first, lDY);
        uint8_t a2 = (uint8_t)(full - partial_triangle_to_alpha(second, lDY));
        blit_two(b, y, ul >> 16, a1, a2, full, no_real);
      } else {
        blit_aaa_row(b, y, ul, joinLeft, ll, joinLeft, lDY, FX_MAX, full, no_real);
      }
    }
    if (joinLeft < joinRite) {
      blit_full(b, y, fx_floor_i(joinLeft), fx_floor_i(fx_sub(joinRite, joinLeft)), full, no_real);
    }
    if (lr > joinRite) {
      int len = fx_ceil_i(fx_sub(lr, joinRite));
      if (len == 1) {
        blit_single(b, y, joinRite >> 16, trapezoid_to_alpha(fx_sub(ur, joinRite), fx_sub(lr, joinRite)), full, no_real);
      } else if (len == 2) {
        fx first = fx_sub(fx_add(joinRite, FX1), ur);
        fx second = fx_sub(fx_sub(lr, ur), first);
        uint8_t a1 = (uint8_t)(full - partial_triangle_to_alpha(first, rDY));
        uint8_t a2 = partial_triangle_to_alpha(second, rDY);
        blit_two(b, y, joinRite >> 16, a1, a2, full, no_real);
      } else {
        blit_aaa_row(b, y, joinRite, ur, joinRite, lr, FX_MAX, rDY, full, no_real);
      }
    }
  } else {
    blit_aaa_row(b, y, ul, ur, ll, lr, lDY, rDY, full, no_real);
  }
}

/* ------------------------------------------------------------- edge walker */
#define HEAD 0
#define TAIL 1
static inline void upd_nny(fx y, fx next_y, fx* nny) { if (y > next_y && y < *nny) *nny = y; } /* :159-162 */
static inline void check_intersection(edge* E, int e, fx next_y, fx* nny) {                        /* :164-169 */
  int p = E[e].prev;
  if (E[p].prev >= 0 && fx_add(E[p].x, E[p].dx) > fx_add(E[e].x, E[e].dx)) *nny = fx_add(next_y, FX1 >> 2);
}
static inline void remove_edge(edge* E, int e) { E[E[e].prev].next = E[e].next; E[E[e].next].prev = E[e].prev; }
static inline void insert_after(edge* E, int e, int after) {
  E[e].prev = after;
  E[e].next = E[after].next;
  E[E[after].next].prev = e;
  E[after].next = e;
}
static void backward_insert_on_x(edge* E, int e) { /* :183-193 */
  fx x = E[e].x;
  int prev = E[e].prev;
  while (E[prev].prev >= 0 && E[prev].x > x) prev = E[prev].prev;
  if (E[prev].next != e) {
    remove_edge(E, e);
    insert_after(E, e, prev);
  }
}
static void insert_new_edges(edge* E, int ne, fx y, fx* nny) { /* :207-247 */
  if (E[ne].upper_y > y) { upd_nny(E[ne].upper_y, y, nny); return; }
  int prev = E[ne].prev;
  if (E[prev].x <= E[ne].x) {
    while (E[ne].upper_y <= y) {
      check_intersection(E, ne, y, nny);
      upd_nny(E[ne].lower_y, y, nny);
      ne = E[ne].next;
    }
    upd_nny(E[ne].upper_y, y, nny);
    return;
  }
  int start = prev; /* backward_insert_start — :200-205 */
  while (E[start].prev >= 0 && E[start].x > E[ne].x) start = E[start].prev;
  do {
    int next = E[ne].next;
    int placed = 0;
    for (;;) {
      if (E[start].next == ne) { placed = 1; break; }
      int after = E[start].next;
      if (E[after].x >= E[ne].x) break;
      start = after;
    }
    if (!placed) {
      remove_edge(E, ne);
      insert_after(E, ne, start);
    }
    check_intersection(E, ne, y, nny);
    upd_nny(E[ne].lower_y, y, nny);
    start = ne;
    ne = next;
  } while (E[ne].upper_y <= y);
  upd_nny(E[ne].upper_y, y, nny);
}
static inline void go_y_shift(edge* e, fx dst_y, int y_shift) { e->y = dst_y; e->x = fx_add(e->x, e->dx >> y_shift); } /* sw_edge.hpp:55-58 */
static inline int too_close_edges(edge* E, int prev, int next, fx lowerY) { /* :139-143 */
  return next >= 0 && prev >= 0 && E[next].upper_y < lowerY &&
         fx_add(E[prev].x, FX1) >= fx_sub(E[next].x, fx_abs(E[next].dx));
}
static inline int too_close_rite(int prevRite, fx ul, fx ll) { return prevRite > fx_floor_i(ul) || prevRite > fx_floor_i(ll); }

/* WalkEdges — sw_raster.cc:546-677.  E[0] = head sentinel, E[1] = tail sentinel. */
static void walk_edges(edge* E, int even_odd, builder* sb, int start_y, int stop_y, fx left_clip, fx right_clip) {
  E[HEAD].x = E[HEAD].upper_x = left_clip;
  E[TAIL].x = E[TAIL].upper_x = right_clip;
  fx y = E[E[HEAD].next].upper_y > i_to_fx(start_y) ? E[E[HEAD].next].upper_y : i_to_fx(start_y);
  fx nny = FX_MAX;
  {
    int e;
    for (e = E[HEAD].next; E[e].upper_y <= y; e = E[e].next) {
      edge* q = &E[e]; /* SWEdge::GoY(dst) — sw_edge.hpp:43-51 */
      if (y == fx_add(q->y, FX1)) { q->x = fx_add(q->x, q->dx); q->y = y; }
      else if (q->y != y) { q->x = fx_add(q->upper_x, fx_mul(q->dx, fx_sub(q->y, q->upper_y))); q->y = y; }
      upd_nny(q->lower_y, y, &nny);
    }
    upd_nny(E[e].upper_y, y, &nny);
  }
  int mask = even_odd ? 1 : -1;
  for (;;) {
    int w = 0, in_interval = 0;
    fx prev_x = E[HEAD].x;
    fx c1 = fx_ceil_fx(fx_add(y, 1));
    fx next_y = nny < c1 ? nny : c1;
    int cur = E[HEAD].next, left_edge = HEAD;
    fx left = left_clip, left_dy = 0;
    int prev_right = fx_floor_i(left_clip);
    nny = FX_MAX;
    int y_shift = 0;
    if (fx_sub(next_y, y) & (FX1 >> 2)) { y_shift = 2; next_y = fx_add(y, FX1 >> 2); }
    else if (fx_sub(next_y, y) & (FX1 >> 1)) { y_shift = 1; }
    uint8_t full = (uint8_t)fx_round_i((fx)(0xFF * (int64_t)fx_sub(next_y, y))); /* fixed_to_alpha — :249,151-153 */
    while (E[cur].upper_y <= y) {
      edge* c = &E[cur];
      w += c->winding;
      int prev_in = in_interval;
      in_interval = (w & mask) != 0;
      int is_left = in_interval && !prev_in, is_right = !in_interval && prev_in;
      if (is_left) {
        left = c->x > left_clip ? c->x : left_clip;
        left_dy = c->dy;
        left_edge = cur;
        go_y_shift(c, next_y, y_shift);
      } else if (is_right) {
        fx right = right_clip < c->x ? right_clip : c->x;
        go_y_shift(c, next_y, y_shift);
        fx nl = left_clip > E[left_edge].x ? left_clip : E[left_edge].x;
        fx nr = right_clip < c->x ? right_clip : c->x;
        fx right_dy = c->dy;
        int no_real = (full == 0xFF && (too_close_rite(prev_right, left, E[left_edge].x) ||
                                        too_close_edges(E, cur, c->next, next_y)));
        blit_trapezoid_row(sb, y >> 16, left, right, nl, nr, left_dy, right_dy, full, no_real);
        prev_right = fx_ceil_i(right > c->x ? right : c->x);
      } else {
        go_y_shift(c, next_y, y_shift);
      }
      int next = c->next;
      while (c->lower_y <= next_y) {
        if (c->curve_count > 0) {
          c->snapped_x = c->x; /* KeepContinuous — sw_edge.cc:294-297 */
          c->snapped_y = c->y;
          if (!update_quad(c)) break;
        } else {
          break;
        }
      }
      if (c->lower_y <= next_y) {
        remove_edge(E, cur);
      } else {
        upd_nny(c->lower_y, next_y, &nny);
        fx new_x = c->x;
        if (new_x < prev_x) backward_insert_on_x(E, cur);
        else prev_x = new_x;
        check_intersection(E, cur, next_y, &nny);
      }
      cur = next;
    }
    if (in_interval) {
      fx nl = left_clip > E[left_edge].x ? left_clip : E[left_edge].x;
      int no_real = full == 0xFF && too_close_edges(E, E[left_edge].prev, left_edge, next_y);
      blit_trapezoid_row(sb, y >> 16, left, right_clip, nl, right_clip, left_dy, 0, full, no_real);
    }
    y = next_y;
    if (y >= i_to_fx(stop_y)) break;
    insert_new_edges(E, cur, y, &nny);
  }
}

/* SortEdges — sw_raster.cc:679-697: std::sort by (upper_y, x, dx).  std::sort is not
 * stable and stroke outlines are full of edges with identical keys, so the result
 * the reference produces depends on libstdc++'s algorithm.  It is restated here
 * (GCC 13 bits/stl_algo.h: introsort, threshold 16, median-of-three to first,
 * unguarded partition, final insertion sort; heapsort when the depth limit hits). */
static inline int edge_less(const edge* a, const edge* b) {
  int va = a->upper_y, vb = b->upper_y;
  if (va == vb) { va = a->x; vb = b->x; }
  if (va == vb) { va = a->dx; vb = b->dx; }
  return va < vb;
}
static inline void edge_swap(edge* a, edge* b) { edge t = *a; *a = *b; *b = t; }
static void ss_unguarded_linear_insert(edge* last) {
  edge val = *last;
  edge* next = last - 1;
  while (edge_less(&val, next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void ss_insertion_sort(edge* first, edge* last) {
  if (first == last) return;
  for (edge* i = first + 1; i != last; ++i) {
    if (edge_less(i, first)) {
      edge val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(edge));
      *first = val;
    } else {
      ss_unguarded_linear_insert(i);
    }
  }
}
static void ss_adjust_heap(edge* first, long hole, long len, edge value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (edge_less(first + child, first + (child - 1))) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  long parent = (hole - 1) / 2; /* __push_heap */
  while (hole > top && edge_less(first + parent, &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void ss_heap_sort(edge* first, edge* last) { /* __partial_sort(first,last,last) */
  long len = last - first;
  if (len >= 2) { /* __make_heap */
    long parent = (len - 2) / 2;
    for (;;) {
      edge v = first[parent];
      ss_adjust_heap(first, parent, len, v);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) { /* __sort_heap */
    --last;
    edge v = *last;
    *last = *first;
    ss_adjust_heap(first, 0, last - first, v);
  }
}
static void ss_introsort_loop(edge* first, edge* last, long depth) {
  while (last - first > 16) {
    if (depth == 0) { ss_heap_sort(first, last); return; }
    --depth;
    edge* mid = first + (last - first) / 2;
    edge *a = first + 1, *b = mid, *c = last - 1; /* __move_median_to_first */
    if (edge_less(a, b)) {
      if (edge_less(b, c)) edge_swap(first, b);
      else if (edge_less(a, c)) edge_swap(first, c);
      else edge_swap(first, a);
    } else if (edge_less(a, c)) edge_swap(first, a);
    else if (edge_less(b, c)) edge_swap(first, c);
    else edge_swap(first, b);
    edge *lo = first + 1, *hi = last; /* __unguarded_partition */
    for (;;) {
      while (edge_less(lo, first)) ++lo;
      --hi;
      while (edge_less(first, hi)) --hi;
      if (!(lo < hi)) break;
      edge_swap(lo, hi);
      ++lo;
    }
    ss_introsort_loop(lo, last, depth);
    last = lo;
  }
}
static void sort_edges(edge* first, size_t n) {
  if (n == 0) return;
  long lg = 0;
  for (size_t v = n; v > 1; v >>= 1) lg++;
  ss_introsort_loop(first, first + n, lg * 2);
  if (n > 16) {
    ss_insertion_sort(first, first + 16);
    for (edge* i = first + 16; i != first + n; ++i) ss_unguarded_linear_insert(i);
  } else {
    ss_insertion_sort(first, first + n);
  }
}

/* Test-harness aid, not part of the reference: restrict the canvas (surface 0) to the pixel rows [g_band_y0, g_band_y1)
 * so that a very large frame can be produced by several processes.  Spans of different rows are independent, so the
 * rows a process keeps are exactly those of the whole frame: paths whose bounds miss the band are skipped before the
 * sweep (g_band_skip, set per op by skbo_render), spans outside it are dropped before the brush. */
static int g_band_y0 = 0, g_band_y1 = 0, g_band_skip = 0;

/* SWRaster::RastePath — sw_raster.cc:731-786.  Appends spans to `out`; bounds4 = raster bounds_. */
static void raster_path(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd,
                        spanvec* out, float* bounds4) {
  primbuf pb;
  memset(&pb, 0, sizeof(pb));
  lower_path(segs, n_segs, ctm, &pb);
  float sl = pb.have ? pb.l : 0, st = pb.have ? pb.t : 0, sr = pb.have ? pb.r : 0, sbm = pb.have ? pb.b : 0;
  float bl = floorf(sl), bt = floorf(st), br = ceilf(sr), bb = ceilf(sbm);
  if (bounds4) { bounds4[0] = bl; bounds4[1] = bt; bounds4[2] = br; bounds4[3] = bb; }
  if (g_band_skip && (bb <= (float)g_band_y0 || bt >= (float)g_band_y1)) { free(pb.v); return; }
  { /* Rect::Intersect — rect.cc:160-172 */
    float l = sl > clip[0] ? sl : clip[0], r = sr < clip[2] ? sr : clip[2];
    float t = st > clip[1] ? st : clip[1], b = sbm < clip[3] ? sbm : clip[3];
    if (!(l < r && t < b)) { l = t = r = b = 0; }
    sl = floorf(l); st = floorf(t); sr = ceilf(r); sbm = ceilf(b);
  }
  if (!(sl < sr && st < sbm)) { free(pb.v); return; }
  /* SWEdgeBuilder::BuildEdges — sw_edge.cc:299-336 */
  size_t cap = pb.n * 2 + 2;
  edge* E = (edge*)calloc(cap, sizeof(edge));
  size_t ne = 2;
  for (size_t i = 0; i < pb.n; i++) {
    prim* q = &pb.v[i];
    if (q->n == 2) {
      edge* e = &E[ne];
      memset(e, 0, sizeof(*e));
      if (set_line(e, q->p[0].x, q->p[0].y, q->p[1].x, q->p[1].y) && !can_be_ignored(st, sbm, e->upper_y, e->lower_y)) ne++;
    } else {
      v2 mono[5];
      int k = chop_quad_y(q->p, mono);
      for (int j = 0; j < k; j++) {
        edge* e = &E[ne];
        memset(e, 0, sizeof(*e));
        float pts[6] = {mono[2 * j].x, mono[2 * j].y, mono[2 * j + 1].x, mono[2 * j + 1].y, mono[2 * j + 2].x, mono[2 * j + 2].y};
        if (set_quad(e, pts) && !can_be_ignored(st, sbm, e->q_first_y, e->q_last_y)) ne++;
      }
    }
  }
  free(pb.v);
  if (ne == 2) { free(E); return; }
  sort_edges(E + 2, ne - 2);
  /* ProcessEdges — :706-729 */
  for (size_t i = 2; i < ne; i++) { E[i].prev = (int)i - 1; E[i].next = (int)i + 1; }
  E[2].prev = HEAD;
  E[ne - 1].next = TAIL;
  E[HEAD].prev = -1; E[HEAD].next = 2;
  E[HEAD].upper_y = E[HEAD].lower_y = FX_MIN; E[HEAD].x = FX_MIN; E[HEAD].dx = 0; E[HEAD].dy = FX_MAX; E[HEAD].upper_x = FX_MIN;
  E[TAIL].prev = (int)ne - 1; E[TAIL].next = -1;
  E[TAIL].upper_y = E[TAIL].lower_y = FX_MAX; E[TAIL].x = FX_MAX; E[TAIL].dx = 0; E[TAIL].dy = FX_MAX; E[TAIL].upper_x = FX_MAX;

  builder sb;
  sb.out = out;
  sb.scan_top = st;
  sb.left = (int)bl;
  sb.width = (int)(br - bl);
  if (sb.width < 0) sb.width = 0;
  sb.alphas = (uint8_t*)calloc((size_t)sb.width + 1, 1);
  sb.curr_y = INT32_MIN;
  int start_y = (int)bt, stop_y = (int)sbm;
  fx left_bound = (fx)((uint32_t)sl << 16), right_bound = (fx)((uint32_t)sr << 16);
  walk_edges(E, even_odd, &sb, start_y, stop_y, left_bound, right_bound);
  acc_flush(&sb);
  free(sb.alphas);
  free(E);
}

/* Coverage mode of unclipped fills: 0 = the software backend's analytic-AA scan converter (raster_path above), 1 = the
 * reference's coverage-AA path (tile-binned lines + signed-area accumulation, skb_area_oracle.h) — what the CUDA
 * backend computes under SKB_COVERAGE_AREA. */
static int g_coverage_mode = 0;
#include "skb_area_oracle.h"

/* ------------------------------------------------------- colour and blend */
/* Colours are kept in the reference's register layout A<<24|R<<16|G<<8|B
 * (include/skity/graphic/color.hpp:34-63); pixels in memory are R,G,B,A bytes. */
static inline uint32_t mul_div_255_round(uint32_t a, uint32_t b) { /* color_priv.hpp:27-41 */
  uint32_t prod = a * b + 128;
  return (prod + (prod >> 8)) >> 8;
}
static inline uint32_t color4f_to_color(const float* c) { /* color.cc:53-59 — truncation */
  float v[4] = {c[3] * 255, c[0] * 255, c[1] * 255, c[2] * 255};
  uint32_t o = 0;
  for (int i = 0; i < 4; i++) {
    float f = v[i] < 0.f ? 0.f : v[i];
    f = f > 255.f ? 255.f : f;
    if (f != f) f = 0.f; /* glm::clamp(NaN) = min(max(NaN,0),255): std::max(NaN,0)=NaN... guard */
    o = (o << 8) | (uint32_t)(uint8_t)f;
  }
  return o;
}
static inline uint32_t color_to_pm(uint32_t c) { /* color_priv.hpp:43-58, color_priv.cc:105-108 */
  uint32_t a = c >> 24, r = (c >> 16) & 0xFF, g = (c >> 8) & 0xFF, b = c & 0xFF;
  if (a != 255) { r = mul_div_255_round(r, a); g = mul_div_255_round(g, a); b = mul_div_255_round(b, a); }
  return (a << 24) | (r << 16) | (g << 8) | b;
}
static inline uint32_t alpha_mul_q(uint32_t c, uint32_t scale) { /* color_priv.hpp:62-68 */
  uint32_t mask = 0xFF00FF;
  uint32_t rb = ((c & mask) * scale) >> 8;
  uint32_t ag = ((c >> 8) & mask) * scale;
  return (rb & mask) | (ag & ~mask);
}
/* SWRenderTarget::BlendPixel for kSrcOver on a premultiplied RGBA target —
 * sw_render_target.cc:12-35,105-113 ; blend_mode.cc:141-145 ; color_priv.hpp:70-72 */
static inline void blend_src_over(uint8_t* px, uint32_t src) {
  uint32_t a = src >> 24;
  if (a == 0) return;
  if (a != 255) {
    uint32_t dst = ((uint32_t)px[3] << 24) | ((uint32_t)px[0] << 16) | ((uint32_t)px[1] << 8) | px[2];
    src = src + alpha_mul_q(dst, 256 - a);
  }
  px[0] = (uint8_t)(src >> 16);
  px[1] = (uint8_t)(src >> 8);
  px[2] = (uint8_t)src;
  px[3] = (uint8_t)(src >> 24);
}

/* PorterDuffBlend — src/graphic/blend_mode.cc:92-192 (colours A<<24|R<<16|G<<8|B, premultiplied) */
static inline uint32_t pm_color_mul(uint32_t s, uint32_t d) { /* color_priv.hpp:74-79 */
  return (mul_div_255_round(s >> 24, d >> 24) << 24) | (mul_div_255_round((s >> 16) & 0xFF, (d >> 16) & 0xFF) << 16) |
         (mul_div_255_round((s >> 8) & 0xFF, (d >> 8) & 0xFF) << 8) | mul_div_255_round(s & 0xFF, d & 0xFF);
}
static float soft_light_component(float sx, float sy, float dx, float dy) { /* blend_mode.cc:92-108 */
  if (2.f * sx <= sy) {
    return dx * dx * (sy - 2 * sx) / dy + (1 - dy) * sx + dx * (-sy + 2 * sx + 1);
  } else if (4.f * dx <= dy) {
    float DSqd = dx * dx, DCub = DSqd * dx, DaSqd = dy * dy, DaCub = DaSqd * dy;
    return (DaSqd * (sx - dx * (3 * sy - 6 * sx - 1)) + 12 * dy * DSqd * (sy - 2 * sx) - 16 * DCub * (sy - 2 * sx) -
            DaCub * sx) / DaSqd;
  }
  return dx * (sy - 2 * sx + 1) + sx - sqrtf(dy * dx) * (sy - 2 * sx) - dy * sx;
}
static uint32_t porter_duff(uint32_t src, uint32_t dst, uint32_t mode) {
  uint32_t sa = src >> 24, da = dst >> 24;
  if (mode > 14 && mode != 21) mode = 3; /* :129-133 */
  switch (mode) {
    case 0: return 0;
    case 1: return src;
    case 2: return dst;
    case 3: return sa == 0 ? dst : src + alpha_mul_q(dst, 256 - sa);
    case 4: return da == 255 ? dst : dst + alpha_mul_q(src, 256 - da);
    case 5: return da == 255 ? src : alpha_mul_q(src, da + 1);
    case 6: return sa == 255 ? dst : alpha_mul_q(dst, sa + 1);
    case 7: return da == 0 ? src : alpha_mul_q(src, 256 - da);
    case 8: return sa == 0 ? dst : alpha_mul_q(dst, 256 - sa);
    case 9: return alpha_mul_q(src, da + 1) + alpha_mul_q(dst, 256 - sa);
    case 10: return alpha_mul_q(dst, sa + 1) + alpha_mul_q(src, 256 - da);
    case 11: return alpha_mul_q(src, 256 - da) + alpha_mul_q(dst, 256 - sa);
    case 12: {
      uint32_t r = 0;
      for (int sh = 0; sh < 32; sh += 8) {
        uint32_t v = ((src >> sh) & 0xFF) + ((dst >> sh) & 0xFF);
        r |= (v > 255u ? 255u : v) << sh;
      }
      return r;
    }
    case 13: return pm_color_mul(src, dst);
    case 14: return src + dst - pm_color_mul(src, dst);
    case 21: {
      if (da == 0) return src;
      float s4[4], d4[4], o[4]; /* r g b a, Color4fFromColor color.cc:44-51 */
      s4[0] = ((src >> 16) & 0xFF) / 255.f; s4[1] = ((src >> 8) & 0xFF) / 255.f; s4[2] = (src & 0xFF) / 255.f; s4[3] = sa / 255.f;
      d4[0] = ((dst >> 16) & 0xFF) / 255.f; d4[1] = ((dst >> 8) & 0xFF) / 255.f; d4[2] = (dst & 0xFF) / 255.f; d4[3] = da / 255.f;
      for (int k = 0; k < 3; k++) o[k] = soft_light_component(s4[k], s4[3], d4[k], d4[3]);
      o[3] = s4[3] + (1 - s4[3]) * d4[3];
      return color4f_to_color(o);
    }
    default: return dst;
  }
}
/* SWRenderTarget::BlendPixel on a premultiplied RGBA target, any mode — sw_render_target.cc:12-35.
 * FastBlend (:97-141) short-cuts give the values the formulas give. */
static inline void blend_pixel(uint8_t* px, uint32_t src, uint32_t mode) {
  if (mode == 3) { blend_src_over(px, src); return; }
  uint32_t dst = ((uint32_t)px[3] << 24) | ((uint32_t)px[0] << 16) | ((uint32_t)px[1] << 8) | px[2];
  uint32_t r = porter_duff(src, dst, mode);
  px[0] = (uint8_t)(r >> 16);
  px[1] = (uint8_t)(r >> 8);
  px[2] = (uint8_t)r;
  px[3] = (uint8_t)(r >> 24);
}

/* ColorFilter::FilterColor — src/effect/color_filter.cc:123-197 ; PMColorToColor / ColorToPMColor —
 * src/graphic/color_priv.cc:85-108 (UnPreMultiply scale = round(255 * 2^24 / alpha)) */
static uint32_t apply_color_filter(const uint32_t* blk, uint32_t c) {
  uint32_t type = blk[0];
  if (type == SKB_CF_BLEND) return porter_duff(blk[2], c, blk[1]);
  uint32_t a = c >> 24;
  uint32_t scale = a ? (uint32_t)((0xFF000000u + a / 2) / a) : 0u;
  uint32_t ch[4], o[4];
  ch[0] = (uint32_t)(((uint64_t)scale * ((c >> 16) & 0xFF) + (1u << 23)) >> 24);
  ch[1] = (uint32_t)(((uint64_t)scale * ((c >> 8) & 0xFF) + (1u << 23)) >> 24);
  ch[2] = (uint32_t)(((uint64_t)scale * (c & 0xFF) + (1u << 23)) >> 24);
  ch[3] = a;
  if (type == SKB_CF_MATRIX) {
    const int16_t* m = (const int16_t*)(blk + 4);
    for (int i = 0; i < 4; i++) {
      int32_t mul = (int32_t)ch[0] * m[5 * i] + (int32_t)ch[1] * m[5 * i + 1] + (int32_t)ch[2] * m[5 * i + 2] + (int32_t)ch[3] * m[5 * i + 3];
      int32_t v = mul / 255 + m[5 * i + 4];
      o[i] = (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  } else {
    const uint8_t* t = (const uint8_t*)(blk + 4);
    for (int i = 0; i < 3; i++) o[i] = t[ch[i]];
    o[3] = a;
  }
  return color_to_pm((o[3] << 24) | (o[0] << 16) | (o[1] << 8) | o[2]);
}

typedef struct surface { uint32_t w, h; uint8_t* px; } surface;

/* GradientColorBrush::LerpColor — sw_span_brush.cc:21-32,239-299 */
static void lerp_color(const skb_dl_paint* p, const float* pool, float t, float out[4]) {
  const float* colors = pool + p->stop_off;
  const float* stops = colors + 4 * (size_t)p->n_colors;
  if (fabsf(t) <= (1.0f / 4096)) t = 0.0f;
  else if (fabsf(t - 1.0f) <= (1.0f / 4096)) t = 1.0f;
  if (p->tile_mode == 3 && (t < 0.0 || t >= 1.0)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  if (p->tile_mode == 0) { t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t); }
  else if (p->tile_mode == 1) { t = t - floorf(t); }
  else if (p->tile_mode == 2) {
    float t1 = t - 1;
    float t2 = (float)(t1 - 2 * floor(t1 * 0.5) - 1); /* evaluated in double, rounded once (t1 * 0.5 promotes) */
    t = fabsf(t2);
  }
  int n = (int)p->n_colors;
  float step = 1.f / (n - 1);
  if ((p->has_stops & 1u) && t <= stops[0]) { memcpy(out, colors, 16); return; }
  int i, si = 0, ei = 1;
  float start = 0.f, end = 0.f;
  for (i = 0; i < n - 1; i++) {
    if (p->has_stops & 1u) { start = stops[i]; end = stops[i + 1]; }
    else { start = step * i; end = step * (i + 1); }
    if (t >= start && t <= end) { si = i; ei = i + 1; break; }
  }
  if (i == n - 1 && n > 0) { memcpy(out, colors + 4 * (n - 1), 16); return; }
  float total = end - start, value = t - start, mix = 0.5f;
  if (total > 0) mix = value / total;
  for (int k = 0; k < 4; k++) out[k] = colors[4 * si + k] * (1 - mix) + colors[4 * ei + k] * mix;
}

/* sampled-texel quantisation: Color4fFromColor then Color4fToColor — color.cc:44-59 */
static inline uint32_t requant(uint32_t c) {
  float f = (float)c / 255.f;
  float v = f * 255;
  v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
  return (uint32_t)(uint8_t)v;
}

/* SWSpanBrush colour for one pixel (premultiplied, register layout) —
 * Solid :140-152, Linear :312-320, Sweep :337-354, Radial :371-379, Pixmap :569-579 + bitmap_sampler.cc:26-40,85-108 */
/* BitmapSampler::GetColor — src/graphic/bitmap_sampler.cc:11-108 ; PixmapBrush::CalculateColor —
 * sw_span_brush.cc:569-579 */
static float remap_tile(float t, uint32_t mode) { /* RemapFloatTile :12-23 */
  if (mode == 0) t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  else if (mode == 1) t = t - floorf(t);
  else if (mode == 2) {
    float t1 = t - 1;
    float t2 = t1 - 2 * floor(t1 * 0.5) - 1;
    t = fabsf(t2);
  }
  return t;
}
static const uint8_t* image_texel(const surface* s, float fx_, float fy_) { /* SampleXY :26-34 */
  uint32_t ix = (uint32_t)(long long)fx_, iy = (uint32_t)(long long)fy_; /* glm::clamp<uint32_t>(float, ...) on x86-64 */
  if (ix > s->w - 1) ix = s->w - 1;
  if (iy > s->h - 1) iy = s->h - 1;
  return s->px + ((size_t)iy * s->w + ix) * 4;
}
static uint32_t sample_image(const skb_dl_paint* p, const surface* s, float u, float v) {
  uint32_t xmode = p->tile_mode & 0xF;
  uint32_t ymode = (p->tile_mode & SKB_PAINT_IMAGE_YMODE) ? ((p->tile_mode >> 4) & 0xF) : xmode;
  if ((xmode == 3 && (u < 0.0 || u >= 1.0)) || (ymode == 3 && (v < 0.0 || v >= 1.0))) return 0;
  u = remap_tile(u, xmode);
  v = remap_tile(v, ymode);
  float c4[4]; /* r g b a */
  if (!(p->tile_mode & SKB_PAINT_IMAGE_LINEAR)) {
    const uint8_t* t = image_texel(s, u * s->w, v * s->h);
    for (int k = 0; k < 4; k++) c4[k] = t[k] / 255.f;
  } else { /* SampleUnitLinear :44-83 */
    float w = (float)s->w, h = (float)s->h;
    float x = u * w, y = v * h;
    float i0 = floorf(x - 0.5f), j0 = floorf(y - 0.5f);
    if (xmode == 1) i0 = i0 - w * floorf(i0 / w);
    if (ymode == 1) j0 = j0 - h * floorf(j0 / h);
    float i1 = i0 + 1.0f, j1 = j0 + 1.0f;
    if (xmode == 1) i1 = i1 - w * floorf(i1 / w);
    if (ymode == 1) j1 = j1 - h * floorf(j1 / h);
    float a = (x - 0.5f) - floorf(x - 0.5f), b = (y - 0.5f) - floorf(y - 0.5f);
    const uint8_t *t00 = image_texel(s, i0, j0), *t10 = image_texel(s, i1, j0), *t01 = image_texel(s, i0, j1),
                  *t11 = image_texel(s, i1, j1);
    float w00 = (1 - a) * (1 - b), w10 = a * (1 - b), w01 = (1 - a) * b, w11 = a * b;
    for (int k = 0; k < 4; k++)
      c4[k] = ((w00 * (t00[k] / 255.f) + w10 * (t10[k] / 255.f)) + w01 * (t01[k] / 255.f)) + w11 * (t11[k] / 255.f);
  }
  uint32_t c = color4f_to_color(c4);
  return (p->tile_mode & SKB_PAINT_IMAGE_UNPREMUL) ? color_to_pm(c) : c;
}

static uint32_t paint_color(const skb_dl_paint* p, const float* pool, const surface* surfs, int x, int y) {
  float fx_ = x + 0.5f, fy_ = y + 0.5f;
  float u = fx_ * p->m[0] + fy_ * p->m[1] + p->m[2];
  float v = fx_ * p->m[3] + fy_ * p->m[4] + p->m[5];
  float c[4];
  switch (p->type) {
    case SKB_PAINT_SOLID:
      return color_to_pm(color4f_to_color(p->color));
    case SKB_PAINT_LINEAR:
      lerp_color(p, pool, u, c);
      return color_to_pm(color4f_to_color(c));
    case SKB_PAINT_RADIAL:
      lerp_color(p, pool, sqrtf(u * u + v * v), c);
      return color_to_pm(color4f_to_color(c));
    case SKB_PAINT_SWEEP: {
      float angle = atan2f(-v, -u);
      const float k1Over2Pi = 0.1591549430918f;
      float t = (float)((angle * k1Over2Pi + 0.5 + p->bias) * p->scale);
      lerp_color(p, pool, t, c);
      return color_to_pm(color4f_to_color(c));
    }
    case SKB_PAINT_CONICAL: { /* ConicalGradientColorBrush::CalculateConical — sw_span_brush.cc:450-513 */
      const float* e = pool + p->stop_off + 5 * (size_t)p->n_colors;
      int kind = (int)e[0];
      float t;
      if (kind == 1) {
        float qx = (u - e[1]) * e[3], qy = (v - e[2]) * e[3];
        t = sqrtf(qx * qx + qy * qy) * e[4] - e[5];
      } else if (kind == 2) {
        float r = e[6], r_2 = r * r;
        float x2 = u * e[7] + v * e[8] + e[9], y2 = u * e[10] + v * e[11] + e[12];
        t = r_2 - y2 * y2;
        if (t < 0.0) return 0;
        t = x2 + sqrtf(t);
      } else if (kind == 3 || kind == 5) {
        float x2 = u * e[7] + v * e[8] + e[9], y2 = u * e[10] + v * e[11] + e[12];
        float r1 = e[13], r1sq = e[14], f = e[15], xt = -1.f;
        if (fabsf(r1 - 1.f) < (1.0f / 4096)) {
          xt = (x2 * x2 + y2 * y2) / 2;
        } else if (r1 > 1.f) {
          float m = r1sq - 1.f, delta = m * y2 * y2 + r1sq * x2 * x2;
          xt = (sqrtf(delta) - x2) / m;
        } else {
          float m = r1sq - 1.f, delta = m * y2 * y2 + r1sq * x2 * x2;
          if (delta > 0) {
            float xt1 = (sqrtf(delta) - x2) / m, xt2 = (-sqrtf(delta) - x2) / m;
            xt = 1.f - f < 0 ? (xt2 < xt1 ? xt2 : xt1) : (xt1 < xt2 ? xt2 : xt1);
          }
        }
        if (xt < 0) return 0;
        t = f + (1.f - f) * xt;
        if (kind == 5) t = (float)(1.0 - t);
      } else {
        return 0;
      }
      lerp_color(p, pool, t, c);
      return color_to_pm(color4f_to_color(c));
    }
    case SKB_PAINT_IMAGE:
      return sample_image(p, &surfs[p->image_surface], u, v);
  }
  return 0;
}

/* SWSpanBrush::Brush + BrushH — sw_span_brush.cc:66-138 */
static void brush_spans(surface* dst, const skbo_span* spans, size_t n, const skb_dl_paint* p, const float* pool,
                        const surface* surfs) {
  int iw = (int)dst->w, ih = (int)dst->h;
  uint32_t galpha = p->type == SKB_PAINT_IMAGE ? (p->global_alpha & 0xFF) : 255u;
  uint32_t mode = p->blend ? p->blend - 1 : 3u;
  for (size_t i = 0; i < n; i++) {
    int x = spans[i].x, y = spans[i].y, len = spans[i].len;
    if (y < 0 || y >= ih) continue;
    if (x >= iw || x + len < 0) continue;
    if (x < 0) { len += x; x = 0; }
    if (x + len >= iw) len = iw - x;
    if (len <= 0) continue;
    uint32_t alpha = (uint8_t)(spans[i].cover & galpha);
    for (int l = 0; l < len; l++) {
      uint32_t color = paint_color(p, pool, surfs, x + l, y);
      if (alpha != 255) color = alpha_mul_q(color, alpha);
      if (p->has_stops >> 8) color = apply_color_filter((const uint32_t*)pool + ((p->has_stops >> 8) - 1), color);
      blend_pixel(dst->px + ((size_t)y * dst->w + (x + l)) * 4, color, mode);
    }
  }
}

/* ---------------------------------------------------------------- clip ops */
/* SWCanvas::State::FindSpan — sw_canvas.cc:219-265 (note the `+ 1` at :257) */
static void find_span(const spanvec* clip, const skbo_span* s, spanvec* out) {
  for (size_t i = 0; i < clip->n; i++) {
    const skbo_span* c = &clip->v[i];
    if (c->y != s->y) continue;
    if (c->x < s->x) {
      if (c->x + c->len < s->x) continue;
      int last = c->x + c->len < s->x + s->len ? c->x + c->len : s->x + s->len;
      sv_push(out, s->x, s->y, last - s->x, c->cover < s->cover ? c->cover : s->cover);
    } else if (c->x == s->x) {
      sv_push(out, s->x, s->y, s->len < c->len ? s->len : c->len, c->cover < s->cover ? c->cover : s->cover);
    } else {
      if (s->x + s->len < c->x) continue;
      int l = s->x + s->len - c->x + 1;
      sv_push(out, c->x, s->y, l < c->len ? l : c->len, c->cover < s->cover ? c->cover : s->cover);
    }
  }
}
/* std::sort on a span array with a strict-weak "less" — the same libstdc++ algorithm as sort_edges above (GCC 13
 * bits/stl_algo.h: introsort, threshold 16, median-of-three to first, unguarded partition, final insertion sort,
 * heapsort at the depth limit).  spans_subtraction and PerformMerge sort with comparators that leave ties
 * (equal x; equal y, x and cover), and what the reference does next depends on how those ties come out. */
typedef int (*span_less_fn)(const skbo_span*, const skbo_span*);
static inline void span_swap(skbo_span* a, skbo_span* b) { skbo_span t = *a; *a = *b; *b = t; }
static void sp_unguarded_linear_insert(skbo_span* last, span_less_fn less) {
  skbo_span val = *last;
  skbo_span* next = last - 1;
  while (less(&val, next)) { *last = *next; last = next; --next; }
  *last = val;
}
static void sp_insertion_sort(skbo_span* first, skbo_span* last, span_less_fn less) {
  if (first == last) return;
  for (skbo_span* i = first + 1; i != last; ++i) {
    if (less(i, first)) {
      skbo_span val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(skbo_span));
      *first = val;
    } else {
      sp_unguarded_linear_insert(i, less);
    }
  }
}
static void sp_adjust_heap(skbo_span* first, long hole, long len, skbo_span value, span_less_fn less) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(first + child, first + (child - 1))) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  long parent = (hole - 1) / 2;
  while (hole > top && less(first + parent, &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
static void sp_heap_sort(skbo_span* first, skbo_span* last, span_less_fn less) {
  long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      skbo_span v = first[parent];
      sp_adjust_heap(first, parent, len, v, less);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    skbo_span v = *last;
    *last = *first;
    sp_adjust_heap(first, 0, last - first, v, less);
  }
}
static void sp_introsort_loop(skbo_span* first, skbo_span* last, long depth, span_less_fn less) {
  while (last - first > 16) {
    if (depth == 0) { sp_heap_sort(first, last, less); return; }
    --depth;
    skbo_span* mid = first + (last - first) / 2;
    skbo_span *a = first + 1, *b = mid, *c = last - 1;
    if (less(a, b)) {
      if (less(b, c)) span_swap(first, b);
      else if (less(a, c)) span_swap(first, c);
      else span_swap(first, a);
    } else if (less(a, c)) span_swap(first, a);
    else if (less(b, c)) span_swap(first, c);
    else span_swap(first, b);
    skbo_span *lo = first + 1, *hi = last;
    for (;;) {
      while (less(lo, first)) ++lo;
      --hi;
      while (less(first, hi)) --hi;
      if (!(lo < hi)) break;
      span_swap(lo, hi);
      ++lo;
    }
    sp_introsort_loop(lo, last, depth, less);
    last = lo;
  }
}
static void sort_spans(skbo_span* first, size_t n, span_less_fn less) {
  if (n == 0) return;
  long lg = 0;
  for (size_t v = n; v > 1; v >>= 1) lg++;
  sp_introsort_loop(first, first + n, lg * 2, less);
  if (n > 16) {
    sp_insertion_sort(first, first + 16, less);
    for (skbo_span* i = first + 16; i != first + n; ++i) sp_unguarded_linear_insert(i, less);
  } else {
    sp_insertion_sort(first, first + n, less);
  }
}
static int span_x_less(const skbo_span* a, const skbo_span* b) { return a->x < b->x; }
/* spans_subtraction — sw_canvas.cc:43-133 */
static void spans_subtract(const spanvec* sub, const spanvec* min, spanvec* out) {
  spanvec ms = {0, 0, 0};
  for (size_t i = 0; i < sub->n; i++) {
    const skbo_span* s = &sub->v[i];
    ms.n = 0;
    for (size_t j = 0; j < min->n; j++)
      if (min->v[j].y == s->y) sv_push(&ms, min->v[j].x, min->v[j].y, min->v[j].len, min->v[j].cover);
    if (ms.n == 0) { sv_push(out, s->x, s->y, s->len, s->cover); continue; }
    sort_spans(ms.v, ms.n, span_x_less);   /* std::sort by x alone (sw_canvas.cc:71-72): ties as libstdc++ leaves them */
    int cx = s->x, cl = s->len;
    for (size_t j = 0; j < ms.n; j++) {
      const skbo_span* m = &ms.v[j];
      if (m->x + m->len < cx || m->x > cx + cl) continue;
      if (m->x < cx) {
        if (m->x + m->len > cx + cl) { cl = 0; break; }
        int last = cx + cl, len = m->x + m->len - cx;
        if (len == 0) continue;
        sv_push(out, cx, s->y, len, s->cover);
        cx += len;
        cl = last - cx;
      } else {
        if (m->x + m->len < cx + cl) {
          int last = cx + cl;
          sv_push(out, cx, s->y, m->x - cx, s->cover);
          cx = m->x + m->len;
          cl = last - cx;
        } else {
          sv_push(out, cx, s->y, m->x - cx, s->cover);
          cl = 0;
        }
      }
      if (cl <= 0) break;
    }
    if (cl > 0) sv_push(out, cx, s->y, cl, s->cover);
  }
  free(ms.v);
}
/* SWCanvas::State — sw_canvas.hpp:26-45.  HasClip() is `!clip_spans_.empty()`: a clip whose span
 * list came out EMPTY (e.g. two disjoint nested clips) behaves as NO clip at all. */
typedef struct clip_state { spanvec spans; int op; } clip_state;
static inline int clip_has(const clip_state* c) { return c && c->spans.n > 0; }
/* State::PerformClip — sw_canvas.cc:158-176 */
static void perform_clip(const clip_state* st, const spanvec* in, spanvec* out) {
  if (st->op == 0) { spans_subtract(in, &st->spans, out); return; }
  for (size_t i = 0; i < in->n; i++) find_span(&st->spans, &in->v[i], out);
}
static int merge_less(const skbo_span* a, const skbo_span* b) {   /* PerformMerge's comparator, sw_canvas.cc:201-213 */
  if (a->y < b->y) return 1;
  if (a->y == b->y) return a->x != b->x ? a->x < b->x : a->cover > b->cover;
  return 0;
}
/* SWCanvas::OnClipPath + State::RecursiveClip/PerformMerge — sw_canvas.cc:178-217,315-336 */
static void clip_refine(const clip_state* parent, const spanvec* fresh, int op, clip_state* out) {
  memset(out, 0, sizeof(*out));
  if (!clip_has(parent)) {
    for (size_t i = 0; i < fresh->n; i++) sv_push(&out->spans, fresh->v[i].x, fresh->v[i].y, fresh->v[i].len, fresh->v[i].cover);
    out->op = op;
    return;
  }
  if (parent->op == op) {
    if (op == 1) {
      perform_clip(parent, fresh, &out->spans);
    } else {
      for (size_t i = 0; i < fresh->n; i++) sv_push(&out->spans, fresh->v[i].x, fresh->v[i].y, fresh->v[i].len, fresh->v[i].cover);
      for (size_t i = 0; i < parent->spans.n; i++) {
        const skbo_span* s = &parent->spans.v[i];
        sv_push(&out->spans, s->x, s->y, s->len, s->cover);
      }
      sort_spans(out->spans.v, out->spans.n, merge_less);
    }
    out->op = parent->op;
  } else {
    if (parent->op == 0) spans_subtract(fresh, &parent->spans, &out->spans);
    else spans_subtract(&parent->spans, fresh, &out->spans);
    out->op = 1;
  }
}

/* morph<type, direction> — image_filter.cc:294-340: per channel max (dilate) / min (erode) over the window
 * [i - radius, i + radius] clamped to the line, along x (dir 0) or y (dir 1); radius = min(radius, n - 1) */
static void morph_pass(const uint8_t* src, uint8_t* dst, int w, int h, int radius, int dir, int erode) {
  int n = dir == 0 ? w : h;
  if (radius > n - 1) radius = n - 1;
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      int i = dir == 0 ? x : y;
      int lo = i - radius < 0 ? 0 : i - radius, hi = i + radius > n - 1 ? n - 1 : i + radius;
      int v[4] = {erode ? 255 : 0, erode ? 255 : 0, erode ? 255 : 0, erode ? 255 : 0};
      for (int k = lo; k <= hi; k++) {
        const uint8_t* p = src + ((size_t)(dir == 0 ? y : k) * w + (dir == 0 ? k : x)) * 4;
        for (int c = 0; c < 4; c++) v[c] = erode ? (p[c] < v[c] ? p[c] : v[c]) : (p[c] > v[c] ? p[c] : v[c]);
      }
      uint8_t* o = dst + ((size_t)y * w + x) * 4;
      for (int c = 0; c < 4; c++) o[c] = (uint8_t)v[c];
    }
  }
}

/* --------------------------------------------------------------- stack blur */
/* SWStackBlur::GetMulSum/GetShrSum — sw_stack_blur.cc:286-338.  The two 255-entry
 * tables are the classic StackBlur reciprocal tables: shr = the largest s with
 * 2^s/(r+1)^2 <= 512 and mul = ceil(2^s/(r+1)^2)  (checked entry-by-entry
 * against the reference's arrays in tests/test_oracle_pinning.py). */
static void blur_mul_shr(int radius, uint64_t* mul, int* shr) {
  uint64_t d = (uint64_t)(radius + 1) * (uint64_t)(radius + 1);
  int s = 0;
  while ((1ull << (s + 1)) <= 512ull * d) s++;
  *shr = s;
  *mul = ((1ull << s) + d - 1) / d;
}
/* SWStackBlur::Blur — sw_stack_blur.cc:18-284.  The sliding "stack" is restated
 * as the triangular kernel it computes: sum(x) = SUM_{i=-r..r} (r+1-|i|) * p[clamp(x+i)],
 * out = uint8((sum * mul[r]) >> shr[r]).  The vertical pass reproduces the
 * reference's seeding of out_sum.b/.a from the G channel (:178-179): that
 * error is constant, so it drifts the B and A sums by -y*(r+1)*(g0-b0|a0). */
static void stack_blur(const uint8_t* src, uint8_t* dst, int w, int h, int radius) {
  if (radius > 254) radius = 254;
  if (radius <= 1) { memcpy(dst, src, (size_t)w * h * 4); return; }
  uint64_t mul;
  int shr;
  blur_mul_shr(radius, &mul, &shr);
  int r1 = radius + 1;
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      uint64_t s[4] = {0, 0, 0, 0};
      for (int i = -radius; i <= radius; i++) {
        int xx = x + i;
        xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
        const uint8_t* p = src + ((size_t)y * w + xx) * 4;
        uint64_t wt = (uint64_t)(r1 - (i < 0 ? -i : i));
        for (int c = 0; c < 4; c++) s[c] += wt * p[c];
      }
      uint8_t* o = dst + ((size_t)y * w + x) * 4;
      for (int c = 0; c < 4; c++) o[c] = (uint8_t)((s[c] * mul) >> shr);
    }
  }
  uint8_t* col = (uint8_t*)malloc((size_t)h * 4);
  for (int x = 0; x < w; x++) {
    for (int y = 0; y < h; y++) memcpy(col + 4 * y, dst + ((size_t)y * w + x) * 4, 4);
    /* memory order here is what the reference names b,g,r,a = bytes 0,1,2,3 */
    uint64_t drift0 = (uint64_t)r1 * ((uint64_t)col[1] - (uint64_t)col[0]);
    uint64_t drift3 = (uint64_t)r1 * ((uint64_t)col[1] - (uint64_t)col[3]);
    for (int y = 0; y < h; y++) {
      uint64_t s[4] = {0, 0, 0, 0};
      for (int i = -radius; i <= radius; i++) {
        int yy = y + i;
        yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
        uint64_t wt = (uint64_t)(r1 - (i < 0 ? -i : i));
        for (int c = 0; c < 4; c++) s[c] += wt * col[4 * yy + c];
      }
      s[0] -= (uint64_t)y * drift0;
      s[3] -= (uint64_t)y * drift3;
      uint8_t* o = dst + ((size_t)y * w + x) * 4;
      for (int c = 0; c < 4; c++) o[c] = (uint8_t)((s[c] * mul) >> shr);
    }
  }
  free(col);
}

/* ------------------------------------------------------------- public API */
#define SKBO_API __attribute__((visibility("default")))

static const void* dl_section(const uint8_t* dl, uint32_t off) { return dl + off; }

/* Executes a whole display list.  Surfaces are allocated zeroed (Bitmap calloc,
 * src/io/pixmap.cc:77); surface 0 is copied to canvas_rgba (w*h*4).  If
 * `initial` is non-NULL surface 0 starts from those pixels.  Returns 0 / <0. */
/* 0 = auto (wide for canvases wider or taller than 8192 px, like skb_surface_set_coord_mode's default),
 * 1 = the reference's int32 arithmetic, 2 = wide.  See line_coord. */
static int g_coord_mode = 0;
SKBO_API void skbo_set_coord_mode(int mode) { g_coord_mode = mode; g_wide = mode == 2; }
/* rows [y0, y1) of the canvas only (y1 <= y0: the whole canvas); see g_band_y0 */
SKBO_API void skbo_set_row_band(int y0, int y1) { g_band_y0 = y0; g_band_y1 = y1; }
/* 0 = software analytic AA (default), 1 = coverage-AA ("AREA") for unclipped fills */
SKBO_API void skbo_set_coverage_mode(int mode) { g_coverage_mode = mode; }

/* The tiled form of one path in coverage-AA mode, for pinning against the reference's own CoverageAAPathTiler:
 * tiles_out[i] = tile_x, tile_y, first line, line count, backdrop; lines_out[j] = from_x, from_y, to_x, to_y (8.8), in
 * range order.  Returns the tile count (negative: capacity exceeded); *n_lines_out = lines written. */
SKBO_API long skbo_area_tile_path(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm6, const float* scissor4, int even_odd,
                                  int32_t* tiles_out, long tile_cap, uint16_t* lines_out, long line_cap, long* n_lines_out) {
  area_tiler t;
  if (!area_tile_path(&t, segs, n_segs, ctm6, scissor4)) { area_tiler_free(&t); *n_lines_out = 0; return 0; }
  area_resolve_backdrops(&t, even_odd);
  uint32_t* off = (uint32_t*)calloc(t.n_ranges + 1, sizeof(uint32_t));
  for (size_t i = 0; i < t.n_ranges; i++) off[i + 1] = off[i] + t.range_counts[i];
  uint32_t* cursor = (uint32_t*)malloc((t.n_ranges + 1) * sizeof(uint32_t));
  memcpy(cursor, off, (t.n_ranges + 1) * sizeof(uint32_t));
  long rc = (long)t.n_tiles;
  if ((long)t.n_tiles > tile_cap || (long)t.n_lines > line_cap) rc = -1;
  if (rc >= 0) {
    for (size_t i = 0; i < t.n_lines; i++) {
      uint16_t* o = lines_out + 4 * (size_t)cursor[t.lines[i].range]++;
      o[0] = t.lines[i].from_x; o[1] = t.lines[i].from_y; o[2] = t.lines[i].to_x; o[3] = t.lines[i].to_y;
    }
    for (size_t i = 0; i < t.n_tiles; i++) {
      int32_t* o = tiles_out + 5 * i;
      uint32_t r = t.tiles[i].range;
      o[0] = t.tiles[i].tile_x; o[1] = t.tiles[i].tile_y;
      o[2] = r == AREA_NO_RANGE ? -1 : (int32_t)off[r];
      o[3] = r == AREA_NO_RANGE ? 0 : (int32_t)t.range_counts[r];
      o[4] = t.tiles[i].backdrop;
    }
    *n_lines_out = (long)t.n_lines;
  }
  free(off); free(cursor);
  area_tiler_free(&t);
  return rc;
}

SKBO_API int skbo_render(const uint8_t* dl, size_t bytes, const uint8_t* initial, uint8_t* canvas_rgba) {
  if (bytes < sizeof(skb_dl_header)) return -1;
  const skb_dl_header* h = (const skb_dl_header*)dl;
  if (h->magic != SKB_DL_MAGIC || h->version != SKB_DL_VERSION || h->total_bytes > bytes) return -1;
  const skb_dl_surface* sdesc = (const skb_dl_surface*)dl_section(dl, h->off_surfaces);
  g_wide = g_coord_mode == 2 || (g_coord_mode == 0 && h->n_surfaces && (sdesc[0].width > 8192 || sdesc[0].height > 8192));
  const skb_dl_op* ops = (const skb_dl_op*)dl_section(dl, h->off_ops);
  const skb_dl_path* paths = (const skb_dl_path*)dl_section(dl, h->off_paths);
  const skb_dl_seg* segs = (const skb_dl_seg*)dl_section(dl, h->off_segs);
  const skb_dl_paint* paints = (const skb_dl_paint*)dl_section(dl, h->off_paints);
  const float* pool = (const float*)dl_section(dl, h->off_stops);
  surface* surfs = (surface*)calloc(h->n_surfaces ? h->n_surfaces : 1, sizeof(surface));
  int rc_images = 0;
  for (uint32_t i = 0; i < h->n_surfaces; i++) {
    surfs[i].w = sdesc[i].width;
    surfs[i].h = sdesc[i].height;
    surfs[i].px = (uint8_t*)calloc((size_t)surfs[i].w * surfs[i].h * 4 + 4, 1);
    if (sdesc[i].flags & SKB_SURFACE_IMAGE) { /* an application image: its pixels come with the display list */
      size_t nb = (size_t)surfs[i].w * surfs[i].h * 4;
      if ((size_t)sdesc[i].reserved + nb > bytes) { rc_images = -6; continue; }
      memcpy(surfs[i].px, dl + sdesc[i].reserved, nb);
    }
  }
  if (initial && h->n_surfaces) memcpy(surfs[0].px, initial, (size_t)surfs[0].w * surfs[0].h * 4);
  clip_state* clips = (clip_state*)calloc((size_t)h->n_clip_states + 1, sizeof(clip_state));
  spanvec spans = {0, 0, 0}, clipped = {0, 0, 0};
  int rc = rc_images;
  for (uint32_t i = 0; i < h->n_ops && rc == 0; i++) {
    const skb_dl_op* op = &ops[i];
    switch (op->kind) {
      case SKB_OP_FILL: {
        const skb_dl_path* p = &paths[op->path];
        spans.n = 0;
        const int banded = g_band_y1 > g_band_y0 && op->surface == 0 && !op->clip_in;
        g_band_skip = banded;
        if (g_coverage_mode == 1 && !op->clip_in)
          area_raster_path(segs + p->seg_off, p->n_segs, op->ctm, op->clip_bounds, (int)op->fill_type, (int)surfs[op->surface].w,
                           (int)surfs[op->surface].h, &spans);
        else
          raster_path(segs + p->seg_off, p->n_segs, op->ctm, op->clip_bounds, (int)op->fill_type, &spans, NULL);
        g_band_skip = 0;
        if (banded) {
          size_t k = 0;
          for (size_t j = 0; j < spans.n; j++)
            if (spans.v[j].y >= g_band_y0 && spans.v[j].y < g_band_y1) spans.v[k++] = spans.v[j];
          spans.n = k;
        }
        const spanvec* use = &spans;
        if (op->clip_in && clip_has(&clips[op->clip_in])) {
          clipped.n = 0;
          perform_clip(&clips[op->clip_in], &spans, &clipped);
          use = &clipped;
        }
        brush_spans(&surfs[op->surface], use->v, use->n, &paints[op->paint], pool, surfs);
      } break;
      case SKB_OP_CLIP: {
        const skb_dl_path* p = &paths[op->path];
        spans.n = 0;
        raster_path(segs + p->seg_off, p->n_segs, op->ctm, op->clip_bounds, (int)op->fill_type, &spans, NULL);
        if (op->clip_out == 0 || op->clip_out > h->n_clip_states) { rc = -2; break; }
        clip_refine(op->clip_in ? &clips[op->clip_in] : NULL, &spans, (int)op->aux, &clips[op->clip_out]);
      } break;
      case SKB_OP_BLUR: {
        surface* d = &surfs[op->surface];
        surface* s = &surfs[op->aux];
        if (d->w != s->w || d->h != s->h) { rc = -3; break; }
        if (op->fill_type == 6 || op->fill_type == 7) { /* MorphologyImageFilter::OnFilter — image_filter.cc:294-385 */
          float rxf = op->clip_bounds[0], ryf = op->clip_bounds[1];
          int erode = op->fill_type == 7, w_ = (int)s->w, h_ = (int)s->h;
          if (rxf > 0 && ryf > 0) {
            uint8_t* tmp = (uint8_t*)calloc((size_t)w_ * h_ * 4 + 4, 1);
            morph_pass(s->px, tmp, w_, h_, (int)rxf, 0, erode);
            morph_pass(tmp, d->px, w_, h_, (int)ryf, 1, erode);
            free(tmp);
          } else if (rxf > 0) {
            morph_pass(s->px, d->px, w_, h_, (int)rxf, 0, erode);
          } else if (ryf > 0) {
            morph_pass(s->px, d->px, w_, h_, (int)ryf, 1, erode);
          }
          break;
        }
        stack_blur(s->px, d->px, (int)s->w, (int)s->h, (int)op->clip_bounds[0]);
        /* MaskFilterOnFilter styles — src/effect/mask_filter.cc:64-100 ; DropShadowImageFilter::OnFilter —
         * src/effect/image_filter.cc:222-233 (pixels are R,G,B,A bytes) */
        if (op->fill_type) {
          for (size_t i = 0; i < (size_t)s->w * s->h; i++) {
            const uint8_t* raw = s->px + 4 * i;
            uint8_t* out = d->px + 4 * i;
            uint32_t raw_a = raw[3], blur_a = out[3];
            if (op->fill_type == 2) {
              if (raw_a > 0) memcpy(out, raw, 4);
            } else if (op->fill_type == 3) {
              if (raw_a > 0 && raw_a >= blur_a) memset(out, 0, 4);
            } else if (op->fill_type == 4) {
              if (raw_a > 0) {
                const float a_factor = 1.f / 255.f;
                uint32_t c = ((uint32_t)out[3] << 24) | ((uint32_t)out[0] << 16) | ((uint32_t)out[1] << 8) | out[2];
                c = alpha_mul_q(c, (unsigned)(a_factor * raw_a * blur_a));
                out[0] = (uint8_t)(c >> 16); out[1] = (uint8_t)(c >> 8); out[2] = (uint8_t)c; out[3] = (uint8_t)(c >> 24);
              } else {
                memset(out, 0, 4);
              }
            } else if (op->fill_type == 5) {
              uint32_t col = op->paint; /* the filter's Color, unpremultiplied A<<24|R<<16|G<<8|B */
              if (blur_a > 0) { out[0] = (uint8_t)(col >> 16); out[1] = (uint8_t)(col >> 8); out[2] = (uint8_t)col; out[3] = (uint8_t)blur_a; }
              else memset(out, 0, 4);
            }
          }
        }
      } break;
      default:
        rc = -4;
    }
  }
  if (rc == 0 && h->n_surfaces) memcpy(canvas_rgba, surfs[0].px, (size_t)surfs[0].w * surfs[0].h * 4);
  for (uint32_t i = 0; i < h->n_surfaces; i++) free(surfs[i].px);
  for (uint32_t i = 0; i <= h->n_clip_states; i++) free(clips[i].spans.v);
  free(clips);
  free(surfs);
  free(spans.v);
  free(clipped.v);
  return rc;
}

/* Coverage of one lowered path: spans in emission order (x,y,len,cover) + raster bounds. */
SKBO_API long skbo_raster_path(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm6, const float* clip4,
                               int even_odd, int32_t* spans_out, long cap, float* bounds4) {
  spanvec sv = {0, 0, 0};
  raster_path(segs, n_segs, ctm6, clip4, even_odd, &sv, bounds4);
  long n = (long)sv.n;
  for (long i = 0; i < n && i < cap; i++) memcpy(spans_out + 4 * i, &sv.v[i], 16);
  free(sv.v);
  return n;
}

/* Coverage-AA ("AREA") coverage of one path as spans (x, y, len, A8 cover), on a surf_w x surf_h surface. */
SKBO_API long skbo_area_raster_path(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm6, const float* clip4, int even_odd,
                                    int surf_w, int surf_h, int32_t* spans_out, long cap) {
  spanvec sv = {0, 0, 0};
  area_raster_path(segs, n_segs, ctm6, clip4, even_odd, surf_w, surf_h, &sv);
  long n = (long)sv.n;
  for (long i = 0; i < n && i < cap; i++) memcpy(spans_out + 4 * i, &sv.v[i], 16);
  free(sv.v);
  return n;
}

SKBO_API void skbo_stack_blur(const uint8_t* src, uint8_t* dst, int w, int h, int radius) {
  stack_blur(src, dst, w, h, radius);
}

SKBO_API const char* skbo_version(void) { return "skb-oracle-port 1 (pinned against oracle/_ref)"; }
