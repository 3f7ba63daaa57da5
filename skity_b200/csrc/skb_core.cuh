// skb_core.cuh — per-thread building blocks of the CUDA raster pipeline.
//
// Everything here is a pure function of its arguments (no warp collectives, no
// globals), marked SKB_HD so the same code is compiled by nvcc into the kernels
// of skb_kernels.cu and by g++ into the CPU simulation the tests use to check
// kernel logic where no GPU is present (tests/sim/).
//
// The arithmetic reproduces Skity's software rasteriser bit for bit; each block
// cites the reference file:line whose behaviour it has to match:
//   * 16.16 / 26.6 fixed point with wrapping int32 shifts  (src/render/sw/sw_subpixel.hpp:19-90)
//   * edges, quadratic forward differencing                 (src/render/sw/sw_edge.cc:18-297)
//   * the analytic-AA edge walk                             (src/render/sw/sw_raster.cc:138-247,546-677)
//   * per-pixel trapezoid coverage                          (src/render/sw/sw_raster.cc:249-544)
//   * colour / blend integer maths                          (src/graphic/color_priv.hpp:20-80)
// Float maths must be compiled with one rounding per operation (-fmad=false).
#ifndef SKB_CORE_CUH
#define SKB_CORE_CUH

#include <stdint.h>
#include <math.h>

#include "include/skb_dl.h"

#if defined(__CUDACC__)
#define SKB_HD __host__ __device__ __forceinline__
#define SKB_HDN __host__ __device__
#else
#define SKB_HD inline
#define SKB_HDN inline
#endif

#if !defined(__CUDACC__)
struct uint2 { unsigned int x, y; };  // host stand-in for the CUDA vector type (CPU simulation builds)
#endif

namespace skb {

typedef int32_t fx;
#define SKB_FX1 65536
#define SKB_FX_MAX 0x7FFFFFFF
#define SKB_FX_MIN (-0x7FFFFFFF)
#define SKB_TILE 16

// ------------------------------------------------------------------ fixed point
SKB_HD fx shl(fx v, int s) { return (fx)((uint32_t)v << s); }
SKB_HD fx fx_add(fx a, fx b) { return (fx)((uint32_t)a + (uint32_t)b); }
SKB_HD fx fx_sub(fx a, fx b) { return (fx)((uint32_t)a - (uint32_t)b); }
SKB_HD fx fx_abs(fx v) { return v < 0 ? (fx)(0u - (uint32_t)v) : v; }
SKB_HD fx fx_mul(fx a, fx b) { return (fx)(((int64_t)a * (int64_t)b) >> 16); }
// SWFixedDiv: the truncating 64-bit quotient (n << 16) / d clamped to +-0x7FFFFFFF.
// Computed in FP64 (the emulated 64-bit integer division costs several times as much on the device): with a = |n| * 2^16
// < 2^47 and b = |d| < 2^31 both exact doubles, a * (1 / b) with a correctly rounded reciprocal is within 2^-5 of a / b,
// so its truncation is the exact quotient or one off; the remainder a - q * b (an integer below 2^48: the fused
// multiply-add is exact) tells which, and one step corrects it.  Every operation is a single IEEE operation, the same on
// the host and on the device (checked against the integer form by tests/test_sim_stages.py).
SKB_HD fx fx_div_int(fx n, fx d) {
  int64_t q = (int64_t)((uint64_t)(int64_t)n << 16) / (int64_t)d;
  if (q < (int64_t)SKB_FX_MIN) q = SKB_FX_MIN;
  if (q > (int64_t)SKB_FX_MAX) q = SKB_FX_MAX;
  return (fx)q;
}
SKB_HD fx fx_div(fx n, fx d) {
#if !defined(SKB_NO_FAST_DIV)
  const bool neg = (n < 0) != (d < 0);
  const double a = (double)(n < 0 ? 0u - (uint32_t)n : (uint32_t)n) * 65536.0;
  const double b = (double)(d < 0 ? 0u - (uint32_t)d : (uint32_t)d);
#if defined(__CUDA_ARCH__)
  double q = trunc(__dmul_rn(a, __drcp_rn(b)));
  const double r = __fma_rn(-q, b, a);
#else
  double q = trunc(a * (1.0 / b));
  const double r = fma(-q, b, a);
#endif
  if (r < 0.0) q -= 1.0;
  else if (r >= b) q += 1.0;
#if defined(__CUDA_ARCH__)
  const int32_t qi = __double2int_rz(q);   // cvt.rzi.s32.f64 saturates: a quotient of 2^31 or more gives 0x7FFFFFFF
#else
  q = q < 2147483647.0 ? q : 2147483647.0;
  const int32_t qi = (int32_t)q;
#endif
  return neg ? -qi : qi;
#else
  return fx_div_int(n, d);
#endif
}
SKB_HD fx snap_y(fx y) { return (fx)((((uint32_t)y + (SKB_FX1 >> 3)) >> 14) << 14); }
SKB_HD int fx_floor_i(fx x) { return x >> 16; }
SKB_HD int fx_ceil_i(fx x) { return (fx)((uint32_t)x + SKB_FX1 - 1) >> 16; }
SKB_HD int fx_round_i(fx x) { return (fx)((uint32_t)x + (SKB_FX1 >> 1)) >> 16; }
SKB_HD fx fx_round_fx(fx x) { return (fx)(((uint32_t)x + (SKB_FX1 >> 1)) & 0xFFFF0000u); }
SKB_HD fx fx_ceil_fx(fx x) { return (fx)(((uint32_t)x + SKB_FX1 - 1) & 0xFFFF0000u); }
SKB_HD fx fx_floor_fx(fx x) { return (fx)((uint32_t)x & 0xFFFF0000u); }
SKB_HD fx i_to_fx(int n) { return (fx)((uint32_t)n << 16); }
SKB_HD fx fx_min(fx a, fx b) { return a < b ? a : b; }
SKB_HD fx fx_max(fx a, fx b) { return a > b ? a : b; }
SKB_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
// static_cast<int>(float) as x86-64 performs it (cvttss2si: out of range / NaN -> INT_MIN)
SKB_HD int32_t f2i(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int32_t)0x80000000;
  return (int32_t)v;
}

// ------------------------------------------------------------------------ edges
// One edge of the scan converter, split by access frequency:
//   Edge      (SWEdge, sw_edge.hpp:18-63)  — touched for every band; 32 bytes, small enough to keep
//                                            a whole path's active list in shared memory
//   QuadState (SWQuadEdge, sw_edge.hpp:65-81) — forward-difference state, touched only when an edge
//                                            steps to its next chord; stays in global memory
// SWEdge's y and upper_x are not kept: until the sweep reaches an edge they equal upper_y and x (UpdateLine sets the
// four together) — which is all SWEdge::GoY reads of them (sw_edge.hpp:43-51) — and afterwards the sweep itself knows
// the y it has brought the edge to.  32 bytes: one sector, two 128-bit accesses.
struct alignas(16) Edge {
  fx x, dx, dy, upper_y;
  fx lower_y;
  int32_t curve;  // curve_count | curve_shift << 8 | (winding & 0xFF) << 16 | valid << 24 | quadratic << 25
  int32_t prev, next;
};
// SWQuadEdge's snapped_x / snapped_y are arguments of update_quad instead of state: every caller sets them right
// before the call (KeepContinuous) and nothing reads what UpdateQuad leaves in them.  32 bytes: one sector.
struct alignas(16) QuadState {
  fx qx, qy, qdx, qdy, qddx, qddy, q_last_x, q_last_y;
};
SKB_HD int edge_count(const Edge& e) { return e.curve & 0xFF; }
SKB_HD int edge_shift(const Edge& e) { return (e.curve >> 8) & 0xFF; }
SKB_HD int edge_winding(const Edge& e) { return (int)(int8_t)((e.curve >> 16) & 0xFF); }
SKB_HD void edge_set_curve(Edge& e, int count, int shift, int winding) {
  e.curve = (count & 0xFF) | ((shift & 0xFF) << 8) | ((winding & 0xFF) << 16) | (e.curve & 0xFF000000);
}
SKB_HD void edge_set_count(Edge& e, int count) { e.curve = (e.curve & ~0xFF) | (count & 0xFF); }
SKB_HD void edge_negate_winding(Edge& e) {
  int w = -edge_winding(e);
  e.curve = (e.curve & ~0xFF0000) | ((w & 0xFF) << 16);
}

// SWEdge::UpdateLine (sw_edge.cc:44-70)
SKB_HD int update_line(Edge& e, fx x0, fx y0, fx x1, fx y1, fx slope) {
  if (y0 > y1) {
    fx t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
    edge_negate_winding(e);
  }
  fx x0x1 = fx_sub(x1, x0) >> 10;
  fx y0y1 = fx_sub(y1, y0) >> 10;
  if (y0y1 == 0) return 0;
  e.x = x0;
  e.dx = slope;
  e.dy = (x0x1 == 0 || slope == 0) ? SKB_FX_MAX : fx_abs(fx_div(y0y1, x0x1));
  e.upper_y = y0;
  e.lower_y = y1;
  return 1;
}

// float pixel coordinate -> 16.16 as SetLine does: trunc(v*4*64) << 10 >> 2 (sw_edge.cc:22-29).
// SWFDot6ToFixed is `x << 10` in int32 (sw_subpixel.hpp:43): at 8192 px the shift leaves the int32 range and the
// coordinate wraps.  `wide` selects the wide-coordinate mode of this backend (include/skb.h, skb_surface_set_coord_mode):
// the same conversion carried out without the overflow (24.8 -> 16.16 is `<< 8`), equal to the reference wherever the
// reference does not wrap and valid up to +-32767 px (what 16.16 holds).
SKB_HD fx fx_from_24_8(int32_t t, int wide) { return wide ? (fx)((uint32_t)t << 8) : (shl(t, 10) >> 2); }
SKB_HD fx line_coord(float v, int wide = 0) { return fx_from_24_8(f2i((v * 4.0f) * 64.0f), wide); }

// SWEdge::SetLine (sw_edge.cc:18-42)
SKB_HD int set_line(Edge& e, float x0f, float y0f, float x1f, float y1f, int wide = 0) {
  fx x0 = line_coord(x0f, wide), y0 = snap_y(line_coord(y0f, wide));
  fx x1 = line_coord(x1f, wide), y1 = snap_y(line_coord(y1f, wide));
  edge_set_curve(e, 0, 0, 1);
  fx y0y1 = fx_sub(y1, y0) >> 10;
  if (y0y1 == 0) return 0;
  fx x0x1 = fx_sub(x1, x0) >> 10;
  fx slope = fx_div(x0x1, y0y1);
  return update_line(e, x0, y0, x1, y1, slope);
}

// SWQuadEdge::UpdateQuad (sw_edge.cc:233-292)
SKB_HDN int update_quad(Edge& e, QuadState& q, const fx snapped_x, const fx snapped_y) {
  int success = 0;
  int count = edge_count(e);
  fx oldx = q.qx, oldy = q.qy, dx = q.qdx, dy = q.qdy;
  fx newx = 0, newy = 0, nsx = 0, nsy = 0;
  const int shift = edge_shift(e);
  do {
    // The three cases of the reference (a chord stepping >= 2 px in y snaps to whole pixels and moves x along
    // the chord, a shorter one snaps to quarter pixels, the last chord ends on the curve's end point) differ
    // only in which y the slope is taken to and in how the snapped end is formed, so they share ONE division:
    // lanes of a warp that are in different cases stay converged through the expensive part.
    bool steep = false;
    fx slope_y;
    if (--count > 0) {
      const fx sdy = dy >> shift;
      newx = fx_add(oldx, dx >> shift);
      newy = fx_add(oldy, sdy);
      steep = fx_abs(sdy) >= SKB_FX1 * 2;
      nsy = fx_min(q.q_last_y, steep ? fx_round_fx(newy) : snap_y(newy));
      slope_y = steep ? newy : nsy;
      dx = fx_add(dx, q.qddx);
      dy = fx_add(dy, q.qddy);
    } else {
      newx = q.q_last_x;
      newy = q.q_last_y;
      nsy = newy;
      slope_y = newy;
    }
    const fx diffy = fx_sub(slope_y, snapped_y) >> 10;
    const fx slope = diffy ? fx_div(fx_sub(newx, snapped_x) >> 10, diffy) : SKB_FX_MAX;
    nsx = steep ? fx_sub(newx, fx_mul(slope, fx_sub(newy, nsy))) : newx;
    if (slope < SKB_FX_MAX) success = update_line(e, snapped_x, snapped_y, nsx, nsy, slope);
    oldx = newx;
    oldy = newy;
  } while (count > 0 && !success);
  q.qx = newx;
  q.qy = newy;
  q.qdx = dx;
  q.qdy = dy;
  edge_set_count(e, count);
  return success;
}

// diff_to_shift with shiftAA = 2 (sw_edge.cc:97-119)
SKB_HD int diff_to_shift(fx dx, fx dy) {
  dx = fx_abs(dx);
  dy = fx_abs(dy);
  fx dist = fx_max(dx, dy) + (fx_min(dx, dy) >> 1);
  dist = fx_add(dist, 1 << 4) >> 5;
  return (32 - clz32((uint32_t)dist)) >> 1;
}

// SWQuadEdge::SetQuad (sw_edge.cc:121-231). p = x0 y0 x1 y1 x2 y2 of a y-monotone quad.
// On success *first_y / *last_y receive q_first_y / q_last_y (used by CanBeIgnored).
SKB_HDN int set_quad(Edge& e, QuadState& q, const float* p, fx* first_y, fx* last_y, int wide = 0) {
  fx x0 = f2i(p[0] * 256.0f), y0 = f2i(p[1] * 256.0f);
  fx x1 = f2i(p[2] * 256.0f), y1 = f2i(p[3] * 256.0f);
  fx x2 = f2i(p[4] * 256.0f), y2 = f2i(p[5] * 256.0f);
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
  edge_set_curve(e, 1 << shift, shift - 1, w);
  fx A = shl(fx_add(fx_sub(fx_sub(x0, x1), x1), x2), 9);
  fx B = shl(fx_sub(x1, x0), 10);
  q.qx = fx_from_24_8(x0, wide);
  q.qdx = fx_add(B, A >> shift) >> 2;
  q.qddx = (A >> (shift - 1)) >> 2;
  A = shl(fx_add(fx_sub(fx_sub(y0, y1), y1), y2), 9);
  B = shl(fx_sub(y1, y0), 10);
  q.qy = snap_y(fx_from_24_8(y0, wide));
  q.qdy = fx_add(B, A >> shift) >> 2;
  q.qddy = (A >> (shift - 1)) >> 2;
  q.q_last_x = fx_from_24_8(x2, wide);
  q.q_last_y = snap_y(fx_from_24_8(y2, wide));
  *first_y = q.qy;
  *last_y = q.q_last_y;
  e.x = e.dx = e.dy = e.upper_y = e.lower_y = 0;
  update_quad(e, q, q.qx, q.qy);
  return 1;
}

// SWEdge::CanBeIgnored (sw_edge.cc:72-88)
SKB_HD int can_be_ignored(float scan_top, float scan_bottom, fx y0, fx y1, int wide = 0) {
  fx start_y = snap_y(line_coord(scan_top, wide));
  fx stop_y = snap_y(line_coord(scan_bottom, wide));
  return (y0 >= stop_y || y1 <= start_y);
}

// ----------------------------------------------------------- curve lowering (fp32)
struct V2 { float x, y; };
SKB_HD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }

// Matrix * (x, y, 0, 1) in glm's order: (m0*x + m1*y) + (m2*0 + m3*1)
// (Path::CopyWithMatrix src/graphic/path.cc:1274-1296 -> src/geometry/matrix.cc:463)
SKB_HD V2 xform(const float* m, V2 p) {
  float zx = 0.0f * 0.0f + m[2] * 1.0f;
  float zy = 0.0f * 0.0f + m[5] * 1.0f;
  return v2((m[0] * p.x + m[1] * p.y) + zx, (m[3] * p.x + m[4] * p.y) + zy);
}

struct CubicCoeff { V2 A, B, C, D; };  // src/geometry/geometry.cc:135-169
SKB_HD CubicCoeff cubic_coeff(V2 p0, V2 p1, V2 p2, V2 p3) {
  CubicCoeff c;
  c.A = v2((p3.x + 3.0f * (p1.x - p2.x)) - p0.x, (p3.y + 3.0f * (p1.y - p2.y)) - p0.y);
  c.B = v2(3.0f * ((p2.x - (p1.x + p1.x)) + p0.x), 3.0f * ((p2.y - (p1.y + p1.y)) + p0.y));
  c.C = v2(3.0f * (p1.x - p0.x), 3.0f * (p1.y - p0.y));
  c.D = p0;
  return c;
}
SKB_HD V2 cubic_eval(const CubicCoeff& c, float t) {
  return v2(((c.A.x * t + c.B.x) * t + c.C.x) * t + c.D.x, ((c.A.y * t + c.B.y) * t + c.C.y) * t + c.D.y);
}
struct QuadCoeff { V2 A, B, C; };  // geometry.cc:44-67
SKB_HD QuadCoeff quad_coeff(V2 q0, V2 q1, V2 q2) {
  QuadCoeff c;
  c.C = q0;
  c.B = v2((q1.x - q0.x) + (q1.x - q0.x), (q1.y - q0.y) + (q1.y - q0.y));
  c.A = v2((q2.x - (q1.x + q1.x)) + q0.x, (q2.y - (q1.y + q1.y)) + q0.y);
  return c;
}
SKB_HD V2 quad_eval(const QuadCoeff& c, float t) {
  return v2((c.A.x * t + c.B.x) * t + c.C.x, (c.A.y * t + c.B.y) * t + c.C.y);
}

// Cubic::ToQuads subdivision count (src/geometry/cubic.cc:29-38): double pow, ceil, >= 1.
SKB_HDN int cubic_quad_count(V2 p1, V2 c1, V2 c2, V2 p2) {
  const float accuracy = 0.1f;
  const double max_hypot2 = 432.0 * (double)accuracy * (double)accuracy;
  V2 a = v2(c1.x * 3.0f - p1.x, c1.y * 3.0f - p1.y);
  V2 b = v2(c2.x * 3.0f - p2.x, c2.y * 3.0f - p2.y);
  V2 p = v2(b.x - a.x, b.y - a.y);
  float err = p.x * p.x + p.y * p.y;
  double n = ceil(pow((double)err / max_hypot2, 1. / 6.0));
  if (!(n > 1.)) n = 1.;
  if (n > 4096.) n = 4096.;  // guard; the reference allocates without bound
  return (int)n;
}

// k-th quad of Cubic::ToQuads (cubic.cc:39-61): returns control and end point (source space).
SKB_HD void cubic_sub_quad(const CubicCoeff& cc, const QuadCoeff& qc, int k, int cnt, V2* ctrl, V2* end) {
  double quad_count = (double)cnt;
  float t0 = (float)((double)k / quad_count), t1 = (float)((double)(k + 1) / quad_count);
  V2 a = cubic_eval(cc, t0), b = cubic_eval(cc, t1);
  float sc = (t1 - t0) * (1.f / 3.f);
  V2 ta = quad_eval(qc, t0), tb = quad_eval(qc, t1);
  V2 c1 = v2(a.x + ta.x * sc, a.y + ta.y * sc);
  V2 c2 = v2(b.x - tb.x * sc, b.y - tb.y * sc);
  *ctrl = v2(((c1.x * 3.f - a.x) + (c2.x * 3.f - b.x)) / 4.f, ((c1.y * 3.f - a.y) + (c2.y * 3.f - b.y)) / 4.f);
  *end = b;
}

// Conic -> exactly two quads (Stroke::QuadPath stroke.cc:929-939, Conic::Chop + subdivided conic.cc:26-65,169-199)
SKB_HD int between_f(float a, float b, float c) { return (a - b) * (c - b) <= 0; }
SKB_HD int finite_f(float v) { return (v - v) == 0.0f; }
SKB_HDN void conic_to_quads(V2 p0, V2 p1, V2 p2, float w, V2 out[5]) {
  float scale = 1.0f / (w + 1.0f);
  V2 wp1 = v2(w * p1.x, w * p1.y);
  V2 m = v2(((p0.x + (wp1.x + wp1.x)) + p2.x) * scale * 0.5f, ((p0.y + (wp1.y + wp1.y)) + p2.y) * scale * 0.5f);
  if (!(finite_f(m.x) && finite_f(m.y))) {
    double w_d = w, w_2 = w_d * 2, scale_half = 1 / (1 + w_d) * 0.5;
    m.x = (float)((p0.x + w_2 * p1.x + p2.x) * scale_half);
    m.y = (float)((p0.y + w_2 * p1.y + p2.y) * scale_half);
  }
  V2 d0p1 = v2((p0.x + wp1.x) * scale, (p0.y + wp1.y) * scale);
  V2 d1p1 = v2((wp1.x + p2.x) * scale, (wp1.y + p2.y) * scale);
  V2 d0p2 = m, d1p0 = m;
  float startY = p0.y, endY = p2.y;
  if (between_f(startY, p1.y, endY)) {
    float midY = d0p2.y;
    if (!between_f(startY, midY, endY)) {
      float closerY = fabsf(midY - startY) < fabsf(midY - endY) ? startY : endY;
      d0p2.y = d1p0.y = closerY;
    }
    if (!between_f(startY, d0p1.y, d0p2.y)) d0p1.y = startY;
    if (!between_f(d1p0.y, d1p1.y, endY)) d1p1.y = endY;
  }
  out[0] = p0; out[1] = d0p1; out[2] = d0p2; out[3] = d1p1; out[4] = p2;
  float prod = 0;
  for (int i = 0; i < 5; i++) prod *= (out[i].x * out[i].y);
  if (!(prod == 0)) {
    for (int i = 1; i < 4; i++) out[i] = p1;
  }
}

// ChopQuadAtYExtrema (src/geometry/geometry.cc:311-349). Returns 1 or 2 monotone quads in dst[5].
SKB_HDN int chop_quad_y(const V2 src[3], V2 dst[5]) {
  float a = src[0].y, b = src[1].y, c = src[2].y;
  float ab = a - b, bc = b - c;
  if (ab < 0) bc = -bc;
  if (ab == 0 || bc < 0) {
    float number = a - b, denom = a - b - b + c;
    if (number < 0) { number = -number; denom = -denom; }
    if (!(denom == 0 || number == 0 || number >= denom)) {
      float r = number / denom;
      if (r == r && r != 0) {
        V2 p01 = v2(src[0].x + (src[1].x - src[0].x) * r, src[0].y + (src[1].y - src[0].y) * r);
        V2 p12 = v2(src[1].x + (src[2].x - src[1].x) * r, src[1].y + (src[2].y - src[1].y) * r);
        dst[0] = src[0];
        dst[1] = p01;
        dst[2] = v2(p01.x + (p12.x - p01.x) * r, p01.y + (p12.y - p01.y) * r);
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

// Number of lowered primitives (lines / quads) a display-list segment expands to.
SKB_HDN int seg_prim_count(const skb_dl_seg& s) {
  switch (s.type_flags & SKB_SEG_TYPE_MASK) {
    case SKB_SEG_LINE:
    case SKB_SEG_CLOSE:
    case SKB_SEG_QUAD:
      return 1;
    case SKB_SEG_CONIC:
      return 2;
    case SKB_SEG_CUBIC:
      return cubic_quad_count(v2(s.p[0], s.p[1]), v2(s.p[2], s.p[3]), v2(s.p[4], s.p[5]), v2(s.p[6], s.p[7]));
    default:
      return 0;
  }
}

// The point a segment really starts at in the lowered path (see SKB_SEG_P0_FROM_PREV_CUBIC).
SKB_HDN V2 seg_start_point(const skb_dl_seg* segs, uint32_t i) {
  const skb_dl_seg& s = segs[i];
  if (s.type_flags & SKB_SEG_P0_FROM_PREV_CUBIC) {
    const skb_dl_seg& c = segs[i - 1];
    CubicCoeff cc = cubic_coeff(v2(c.p[0], c.p[1]), v2(c.p[2], c.p[3]), v2(c.p[4], c.p[5]), v2(c.p[6], c.p[7]));
    return cubic_eval(cc, 1.0f);
  }
  return v2(s.start[0], s.start[1]);
}

// k-th lowered primitive of segment i, transformed by the CTM.  Returns 2 (line) or 3 (quad).
SKB_HDN int seg_prim(const skb_dl_seg* segs, uint32_t i, int k, int n_prims, const float* ctm, V2 out[3]) {
  const skb_dl_seg& s = segs[i];
  uint32_t type = s.type_flags & SKB_SEG_TYPE_MASK;
  V2 p0 = v2(s.p[0], s.p[1]), p1 = v2(s.p[2], s.p[3]), p2 = v2(s.p[4], s.p[5]), p3 = v2(s.p[6], s.p[7]);
  switch (type) {
    case SKB_SEG_LINE:
    case SKB_SEG_CLOSE:
      out[0] = xform(ctm, seg_start_point(segs, i));
      out[1] = xform(ctm, p1);
      out[2] = out[1];
      return 2;
    case SKB_SEG_QUAD:
      out[0] = xform(ctm, seg_start_point(segs, i));
      out[1] = xform(ctm, p1);
      out[2] = xform(ctm, p2);
      return 3;
    case SKB_SEG_CONIC: {
      V2 q[5];
      conic_to_quads(p0, p1, p2, s.w, q);
      if (k == 0) {
        out[0] = xform(ctm, seg_start_point(segs, i));
        out[1] = xform(ctm, q[1]);
        out[2] = xform(ctm, q[2]);
      } else {
        out[0] = xform(ctm, q[2]);
        out[1] = xform(ctm, q[3]);
        out[2] = xform(ctm, q[4]);
      }
      return 3;
    }
    case SKB_SEG_CUBIC: {
      CubicCoeff cc = cubic_coeff(p0, p1, p2, p3);
      QuadCoeff qc = quad_coeff(v2(3.f * (p1.x - p0.x), 3.f * (p1.y - p0.y)), v2(3.f * (p2.x - p1.x), 3.f * (p2.y - p1.y)),
                                v2(3.f * (p3.x - p2.x), 3.f * (p3.y - p2.y)));
      V2 ctrl, end;
      cubic_sub_quad(cc, qc, k, n_prims, &ctrl, &end);
      if (k == 0) {
        out[0] = xform(ctm, seg_start_point(segs, i));
      } else {
        double quad_count = (double)n_prims;
        out[0] = xform(ctm, cubic_eval(cc, (float)((double)k / quad_count)));
      }
      out[1] = xform(ctm, ctrl);
      out[2] = xform(ctm, end);
      return 3;
    }
    default:
      return 0;
  }
}

// ------------------------------------------------- trapezoid rows (walker output)
// One call of blit_trapezoid_row (sw_raster.cc:457-544) as the walker would make it.
struct alignas(16) TrapRec {  // 32 bytes: moved as two 128-bit words
  int32_t y;           // pixel row
  fx ul, ur, ll, lr;   // upper-left/right, lower-left/right x of the band's interval
  fx ldy, rdy;         // |dy/dx| of the left / right edge
  uint32_t flags;      // full alpha (bits 0-7) | no_real_span_builder << 8
};
#define SKB_REC_LINK 0x80000000u  // flags of the chunk-link pseudo record (y = next record index)

SKB_HD uint8_t partial_alpha_mul(uint32_t alpha, uint32_t full) { return (uint8_t)((alpha * full) >> 8); }
SKB_HD uint8_t trapezoid_to_alpha(fx l1, fx l2) { return (uint8_t)((fx_add(l1, l2) / 2) >> 8); }
SKB_HD uint8_t partial_triangle_to_alpha(fx a, fx b) {
  uint32_t area = (uint32_t)(a >> 11) * (uint32_t)(a >> 11) * (uint32_t)(b >> 11);
  return (uint8_t)((((fx)area) >> 8) & 0xFF);
}

// compute_alpha_below_line evaluated at index j (sw_raster.cc:343-368)
SKB_HD uint8_t alpha_below_at(int j, fx l, fx r, fx dY, uint32_t full) {
  int R = fx_ceil_i(r);
  if (R == 1) {
    return partial_alpha_mul(trapezoid_to_alpha(l, r), full);
  }
  fx first = fx_sub(SKB_FX1, l);
  fx last = fx_sub(r, shl(R - 1, 16));
  fx lastH = fx_mul(last, dY);
  if (j == R - 1) return (uint8_t)(fx_mul(last, lastH) >> 9);
  if (j == 0) return (uint8_t)(full - partial_triangle_to_alpha(first, dY));
  fx a16 = fx_add(fx_add(lastH, dY >> 1), (fx)((uint32_t)(R - 2 - j) * (uint32_t)dY));
  return (uint8_t)((a16 >> 8) & 0xFF);
}
// compute_alpha_above_line evaluated at index j (sw_raster.cc:314-339)
SKB_HD uint8_t alpha_above_at(int j, fx l, fx r, fx dY, uint32_t full) {
  int R = fx_ceil_i(r);
  if (R == 1) {
    return partial_alpha_mul((uint8_t)(fx_sub(fx_sub(shl(R, 17), l), r) >> 9), full);
  }
  fx first = fx_sub(SKB_FX1, l);
  fx last = fx_sub(r, shl(R - 1, 16));
  fx firstH = fx_mul(first, dY);
  if (j == 0) return (uint8_t)(fx_mul(first, firstH) >> 9);
  if (j == R - 1) return (uint8_t)(full - partial_triangle_to_alpha(last, dY));
  fx a16 = fx_add(fx_add(firstH, dY >> 1), (fx)((uint32_t)(j - 1) * (uint32_t)dY));
  return (uint8_t)(a16 >> 8);
}

// approximate_intersection's mean (sw_raster.cc:253-262): `(max(l1,l2) + min(r1,r2)) / 2`.  The reference adds in int32; the sum
// stays below 2^31 for every coordinate the reference can represent without wrapping (|x| < 8192 px -> |sum| < 2^30), so
// the 64-bit sum used here is identical to it there.  In the wide-coordinate mode two x at the right clip of a 16384-px
// canvas add up to exactly 2^31: the int32 sum would flip sign and turn the band into a 32768-pixel trapezoid (and in
// the reference's SpanBuilder into a write before its row buffer).
SKB_HD fx mean_no_overflow(fx a, fx b) { return (fx)(((int64_t)a + (int64_t)b) / 2); }

// blit_aaa_trapezoid_row at pixel x (sw_raster.cc:370-455). `accum` = the row goes through the
// accumulating SpanBuilder (partial-height band or "too close"), which scales single alphas.
// `side` (a compile-time constant at every call): 0 = both slanted sides; 1 / 2 = only the left / right one, for the
// calls whose other side is the vertical join line — there its two tests (`u + 2 == l`, `x` beyond it) cannot hold for
// a pixel of the zone, so leaving them out changes nothing but the code the compiler has to schedule.
SKB_HD bool aaa_row_at(int x, fx ul, fx ur, fx ll, fx lr, fx lDY, fx rDY, uint32_t full, bool accum, uint8_t* out,
                       const int side = 0) {
  int L = fx_floor_i(ul), R = fx_ceil_i(lr);
  int len = R - L;
  if (x < L || x >= R) return false;
  if (len == 1) {
    uint8_t a = trapezoid_to_alpha(fx_sub(ur, ul), fx_sub(lr, ll));
    *out = accum ? partial_alpha_mul(a, full) : a;
    return true;
  }
  int i = x - L;
  uint32_t a = full;
  int uL = L, lL = fx_ceil_i(ll);
  if (side == 2) {
  } else if (uL + 2 == lL) {
    fx first = fx_sub(fx_add(i_to_fx(uL), SKB_FX1), ul);
    fx second = fx_sub(fx_sub(ll, ul), first);
    if (i == 0) {
      uint8_t a1 = (uint8_t)(full - partial_triangle_to_alpha(first, lDY));
      a = a > a1 ? a - a1 : 0;
    } else if (i == 1) {
      uint8_t a2 = partial_triangle_to_alpha(second, lDY);
      a = a > a2 ? a - a2 : 0;
    }
  } else if (x < lL) {
    uint8_t t = alpha_below_at(x - uL, fx_sub(ul, i_to_fx(uL)), fx_sub(ll, i_to_fx(uL)), lDY, full);
    a = a > t ? a - t : 0;
  }
  int uR = fx_floor_i(ur), lR = R;
  if (side == 1) {
  } else if (uR + 2 == lR) {
    fx first = fx_sub(fx_add(i_to_fx(uR), SKB_FX1), ur);
    fx second = fx_sub(fx_sub(lr, ur), first);
    if (i == len - 2) {
      uint8_t a1 = partial_triangle_to_alpha(first, rDY);
      a = a > a1 ? a - a1 : 0;
    } else if (i == len - 1) {
      uint8_t a2 = (uint8_t)(full - partial_triangle_to_alpha(second, rDY));
      a = a > a2 ? a - a2 : 0;
    }
  } else if (x >= uR) {
    uint8_t t = alpha_above_at(x - uR, fx_sub(ur, i_to_fx(uR)), fx_sub(lr, i_to_fx(uR)), rDY, full);
    a = a > t ? a - t : 0;
  }
  *out = (uint8_t)a;
  return true;
}

// Pixel extent [x0, x1) a trapezoid row can touch (conservative; used for culling).
SKB_HD void trap_extent(const TrapRec& r, int* x0, int* x1) {
  fx lo = fx_min(fx_min(r.ul, r.ll), fx_min(r.ur, r.lr));
  fx hi = fx_max(fx_max(r.ul, r.ll), fx_max(r.ur, r.lr));
  *x0 = fx_floor_i(lo);
  *x1 = fx_ceil_i(hi);
}

// blit_trapezoid_row evaluated at one pixel (sw_raster.cc:457-544): returns whether the
// reference would emit a coverage value for pixel x, and that value.
SKB_HDN bool trap_alpha_at(const TrapRec& r, int x, uint8_t* out) {
  fx ul = r.ul, ur = r.ur, ll = r.ll, lr = r.lr;
  const uint32_t full = r.flags & 0xFF;
  const bool accum = !(full == 0xFF && !((r.flags >> 8) & 1));
  if (ul > ur) return false;
  if (ll > lr) {  // approximate_intersection (sw_raster.cc:253-262)
    fx l1 = ul, r1 = ll, l2 = ur, r2 = lr;
    if (l1 > r1) { fx t = l1; l1 = r1; r1 = t; }
    if (l2 > r2) { fx t = l2; l2 = r2; r2 = t; }
    ll = lr = mean_no_overflow(fx_max(l1, l2), fx_min(r1, r2));
  }
  if (ul == ur && ll == lr) return false;
  if (ul > ll) { fx t = ul; ul = ll; ll = t; }
  if (ur > lr) { fx t = ur; ur = lr; lr = t; }
  fx joinLeft = fx_ceil_fx(ll);
  fx joinRite = fx_floor_fx(ur);
  if (joinLeft <= joinRite) {
    if (ul < joinLeft) {
      int len = fx_ceil_i(fx_sub(joinLeft, ul));
      int x0 = ul >> 16;
      if (len == 1) {
        if (x == x0) {
          uint8_t a = trapezoid_to_alpha(fx_sub(joinLeft, ul), fx_sub(joinLeft, ll));
          *out = accum ? partial_alpha_mul(a, full) : a;
          return true;
        }
      } else if (len == 2) {
        if (x == x0 || x == x0 + 1) {
          fx first = fx_sub(fx_sub(joinLeft, SKB_FX1), ul);
          fx second = fx_sub(fx_sub(ll, ul), first);
          *out = x == x0 ? partial_triangle_to_alpha(first, r.ldy)
                         : (uint8_t)(full - partial_triangle_to_alpha(second, r.ldy));
          return true;
        }
      } else {
        if (aaa_row_at(x, ul, joinLeft, ll, joinLeft, r.ldy, SKB_FX_MAX, full, accum, out)) return true;
      }
    }
    if (joinLeft < joinRite) {
      int xs = fx_floor_i(joinLeft);
      int n = fx_floor_i(fx_sub(joinRite, joinLeft));
      if (x >= xs && x < xs + n) {
        *out = (uint8_t)full;
        return true;
      }
    }
    if (lr > joinRite) {
      int len = fx_ceil_i(fx_sub(lr, joinRite));
      int x0 = joinRite >> 16;
      if (len == 1) {
        if (x == x0) {
          uint8_t a = trapezoid_to_alpha(fx_sub(ur, joinRite), fx_sub(lr, joinRite));
          *out = accum ? partial_alpha_mul(a, full) : a;
          return true;
        }
      } else if (len == 2) {
        if (x == x0 || x == x0 + 1) {
          fx first = fx_sub(fx_add(joinRite, SKB_FX1), ur);
          fx second = fx_sub(fx_sub(lr, ur), first);
          *out = x == x0 ? (uint8_t)(full - partial_triangle_to_alpha(first, r.rdy))
                         : partial_triangle_to_alpha(second, r.rdy);
          return true;
        }
      } else {
        if (aaa_row_at(x, joinRite, ur, joinRite, lr, SKB_FX_MAX, r.rdy, full, accum, out)) return true;
      }
    }
    return false;
  }
  return aaa_row_at(x, ul, ur, ll, lr, r.ldy, r.rdy, full, accum, out);
}

// ---- prepared form: normalise a record once, then evaluate many pixels cheaply ---------------
// Most pixels of a trapezoid row are either outside it or in its fully covered interior; only the
// few pixels under the two slanted edges need the triangle/ramp formulas.  trap_prepare() does the
// per-record work of blit_trapezoid_row (swaps, join points) once; trap_prep_alpha() is then a
// range test for interior/outside pixels and the edge formulas otherwise.  Results are identical
// to trap_alpha_at() (checked exhaustively by the tests).
struct TrapPrep {
  fx ul, ur, ll, lr, ldy, rdy, join_left, join_rite;
  int L, R;    // pixels [L, R) receive a value
  int jl, jr;  // pixels [jl, jr) receive `full`
  uint32_t full;
  int mode;    // 0 nothing, 1 left part / interior / right part, 2 one anti-aliased row
  bool accum;
};

SKB_HDN TrapPrep trap_prepare(const TrapRec& r) {
  TrapPrep p;
  fx ul = r.ul, ur = r.ur, ll = r.ll, lr = r.lr;
  p.full = r.flags & 0xFF;
  p.accum = !(p.full == 0xFF && !((r.flags >> 8) & 1));
  p.ldy = r.ldy;
  p.rdy = r.rdy;
  p.mode = 0;
  p.L = p.R = p.jl = p.jr = 0;
  p.ul = p.ur = p.ll = p.lr = p.join_left = p.join_rite = 0;
  if (ul > ur) return p;
  if (ll > lr) {
    fx l1 = ul, r1 = ll, l2 = ur, r2 = lr;
    if (l1 > r1) { fx t = l1; l1 = r1; r1 = t; }
    if (l2 > r2) { fx t = l2; l2 = r2; r2 = t; }
    ll = lr = mean_no_overflow(fx_max(l1, l2), fx_min(r1, r2));
  }
  if (ul == ur && ll == lr) return p;
  if (ul > ll) { fx t = ul; ul = ll; ll = t; }
  if (ur > lr) { fx t = ur; ur = lr; lr = t; }
  p.ul = ul; p.ur = ur; p.ll = ll; p.lr = lr;
  p.join_left = fx_ceil_fx(ll);
  p.join_rite = fx_floor_fx(ur);
  if (p.join_left <= p.join_rite) {
    p.mode = 1;
    p.jl = p.join_left >> 16;
    p.jr = p.join_rite >> 16;
    p.L = ul < p.join_left ? (ul >> 16) : p.jl;
    p.R = lr > p.join_rite ? p.jr + fx_ceil_i(fx_sub(lr, p.join_rite)) : p.jr;
  } else {
    p.mode = 2;
    p.L = fx_floor_i(ul);
    p.R = fx_ceil_i(lr);
    int lL = fx_ceil_i(ll), uR = fx_floor_i(ur);
    if (p.R - p.L > 1 && lL < uR) {
      p.jl = lL;
      p.jr = uR;
    } else {
      p.jl = p.jr = p.L;
    }
  }
  return p;
}

SKB_HDN bool trap_prep_alpha(const TrapPrep& p, int x, uint8_t* out) {
  if (p.mode == 0 || x < p.L || x >= p.R) return false;
  if (x >= p.jl && x < p.jr) {
    *out = (uint8_t)p.full;
    return true;
  }
  if (p.mode == 2) return aaa_row_at(x, p.ul, p.ur, p.ll, p.lr, p.ldy, p.rdy, p.full, p.accum, out);
  if (x < p.jl) {
    int len = p.jl - p.L;
    if (len == 1) {
      uint8_t a = trapezoid_to_alpha(fx_sub(p.join_left, p.ul), fx_sub(p.join_left, p.ll));
      *out = p.accum ? partial_alpha_mul(a, p.full) : a;
      return true;
    }
    if (len == 2) {
      fx first = fx_sub(fx_sub(p.join_left, SKB_FX1), p.ul);
      fx second = fx_sub(fx_sub(p.ll, p.ul), first);
      *out = x == p.L ? partial_triangle_to_alpha(first, p.ldy) : (uint8_t)(p.full - partial_triangle_to_alpha(second, p.ldy));
      return true;
    }
    return aaa_row_at(x, p.ul, p.join_left, p.ll, p.join_left, p.ldy, SKB_FX_MAX, p.full, p.accum, out, 1);
  }
  int len = p.R - p.jr;
  if (len == 1) {
    uint8_t a = trapezoid_to_alpha(fx_sub(p.ur, p.join_rite), fx_sub(p.lr, p.join_rite));
    *out = p.accum ? partial_alpha_mul(a, p.full) : a;
    return true;
  }
  if (len == 2) {
    fx first = fx_sub(fx_add(p.join_rite, SKB_FX1), p.ur);
    fx second = fx_sub(fx_sub(p.lr, p.ur), first);
    *out = x == p.jr ? (uint8_t)(p.full - partial_triangle_to_alpha(first, p.rdy)) : partial_triangle_to_alpha(second, p.rdy);
    return true;
  }
  return aaa_row_at(x, p.join_rite, p.ur, p.join_rite, p.lr, SKB_FX_MAX, p.rdy, p.full, p.accum, out, 2);
}

// ------------------------------------------------------------- colour and blend
// Pixels are handled as little-endian words of the R,G,B,A bytes in memory:
// word = R | G<<8 | B<<16 | A<<24.  AlphaMulQ / SrcOver treat the four bytes alike
// (alpha sits in the top byte in both layouts), so no swizzle is needed.
// The reference blends in registers laid out A<<24|R<<16|G<<8|B.  Its packed adds (PMSrcOver and the
// other Porter-Duff sums) let a channel that overflows carry into the next one up — B into G, G into R, R
// into A — which happens for real when a source is not a valid premultiplied colour (StackBlur's seeding
// quirk produces such pixels).  To carry the same way, destination and source are swapped into that
// order around every blend (one PRMT each way).
SKB_HD uint32_t swap_rb(uint32_t c) { return (c & 0xFF00FF00u) | ((c >> 16) & 0xFFu) | ((c & 0xFFu) << 16); }
SKB_HD uint32_t mul_div_255_round(uint32_t a, uint32_t b) {
  uint32_t prod = a * b + 128;
  return (prod + (prod >> 8)) >> 8;
}
SKB_HD uint32_t unit_to_byte(float f) {  // Color4fToColor per channel (src/graphic/color.cc:53-59): clamp, truncate
  float v = f * 255.0f;
  v = v < 0.f ? 0.f : v;
  v = v > 255.f ? 255.f : v;
  if (!(v == v)) v = 0.f;
  return (uint32_t)(uint8_t)v;
}
// Color4f (unpremultiplied r,g,b,a) -> premultiplied pixel word (Color4fToColor + ColorToPMColor)
SKB_HD uint32_t color4f_to_pm_word(float r, float g, float b, float a) {
  uint32_t A = unit_to_byte(a), R = unit_to_byte(r), G = unit_to_byte(g), B = unit_to_byte(b);
  if (A != 255) {
    R = mul_div_255_round(R, A);
    G = mul_div_255_round(G, A);
    B = mul_div_255_round(B, A);
  }
  return R | (G << 8) | (B << 16) | (A << 24);
}
SKB_HD uint32_t alpha_mul_q(uint32_t c, uint32_t scale) {  // color_priv.hpp:62-68
  const uint32_t mask = 0xFF00FF;
  uint32_t rb = ((c & mask) * scale) >> 8;
  uint32_t ag = ((c >> 8) & mask) * scale;
  return (rb & mask) | (ag & ~mask);
}
// One SWSpanBrush::BrushH + SWRenderTarget::BlendPixel(kSrcOver) step on a premultiplied target
// (sw_span_brush.cc:108-119, sw_render_target.cc:88-113, color_priv.hpp:70-72).
SKB_HD uint32_t blend_cover(uint32_t dst, uint32_t color, uint32_t cover) {
  if (cover != 255) color = alpha_mul_q(color, cover);
  uint32_t a = color >> 24;
  if (a == 0) return dst;
  if (a == 255) return color;
  return color + alpha_mul_q(dst, 256 - a);
}

// PorterDuffBlend (src/graphic/blend_mode.cc:92-192) on premultiplied pixel words.  Every mode treats
// the colour channels alike, so the R/B order of the word does not matter; alpha is the top byte.
// `mode` is skity::BlendMode's value; modes the reference does not implement (above kScreen, other
// than kSoftLight) fall back to kSrcOver there (:129-133) — the host canonicalises them.
#define SKB_BLEND_SRC_OVER 3u
SKB_HD uint32_t pm_color_mul(uint32_t s, uint32_t d) {  // PMColorMul (color_priv.hpp:74-79)
  return (mul_div_255_round(s >> 24, d >> 24) << 24) | (mul_div_255_round((s >> 16) & 0xFF, (d >> 16) & 0xFF) << 16) |
         (mul_div_255_round((s >> 8) & 0xFF, (d >> 8) & 0xFF) << 8) | mul_div_255_round(s & 0xFF, d & 0xFF);
}
SKB_HD float soft_light_component(float sx, float sy, float dx, float dy) {  // blend_mode.cc:92-108
  if (2.f * sx <= sy) {
    return dx * dx * (sy - 2 * sx) / dy + (1 - dy) * sx + dx * (-sy + 2 * sx + 1);
  } else if (4.f * dx <= dy) {
    float DSqd = dx * dx;
    float DCub = DSqd * dx;
    float DaSqd = dy * dy;
    float DaCub = DaSqd * dy;
    return (DaSqd * (sx - dx * (3 * sy - 6 * sx - 1)) + 12 * dy * DSqd * (sy - 2 * sx) - 16 * DCub * (sy - 2 * sx) -
            DaCub * sx) /
           DaSqd;
  } else {
    return dx * (sy - 2 * sx + 1) + sx - sqrtf(dy * dx) * (sy - 2 * sx) - dy * sx;
  }
}
SKB_HDN uint32_t porter_duff(uint32_t src, uint32_t dst, uint32_t mode) {
  const uint32_t sa = src >> 24, da = dst >> 24;
  switch (mode) {
    case 0: return 0;                                                 // kClear
    case 1: return src;                                               // kSrc
    case 2: return dst;                                               // kDst
    case 3: return sa == 0 ? dst : src + alpha_mul_q(dst, 256 - sa);  // kSrcOver (PMSrcOver)
    case 4: return da == 255 ? dst : dst + alpha_mul_q(src, 256 - da);                // kDstOver
    case 5: return da == 255 ? src : alpha_mul_q(src, da + 1);                        // kSrcIn
    case 6: return sa == 255 ? dst : alpha_mul_q(dst, sa + 1);                        // kDstIn
    case 7: return da == 0 ? src : alpha_mul_q(src, 256 - da);                        // kSrcOut
    case 8: return sa == 0 ? dst : alpha_mul_q(dst, 256 - sa);                        // kDstOut
    case 9: return alpha_mul_q(src, da + 1) + alpha_mul_q(dst, 256 - sa);             // kSrcATop
    case 10: return alpha_mul_q(dst, sa + 1) + alpha_mul_q(src, 256 - da);            // kDstATop
    case 11: return alpha_mul_q(src, 256 - da) + alpha_mul_q(dst, 256 - sa);          // kXor
    case 12: {                                                                        // kPlus
      uint32_t r = 0;
      for (int sh = 0; sh < 32; sh += 8) {
        uint32_t v = ((src >> sh) & 0xFF) + ((dst >> sh) & 0xFF);
        r |= (v > 255u ? 255u : v) << sh;
      }
      return r;
    }
    case 13: return pm_color_mul(src, dst);                                           // kModulate
    case 14: return src + dst - pm_color_mul(src, dst);                               // kScreen
    case 21: {                                                                        // kSoftLight
      if (da == 0) return src;
      float s[4], d[4];
      for (int k = 0; k < 4; k++) {  // Color4fFromColor (color.cc:44-51)
        s[k] = (float)((src >> (8 * k)) & 0xFF) / 255.f;
        d[k] = (float)((dst >> (8 * k)) & 0xFF) / 255.f;
      }
      uint32_t r = 0;
      for (int k = 0; k < 3; k++) r |= unit_to_byte(soft_light_component(s[k], s[3], d[k], d[3])) << (8 * k);
      r |= unit_to_byte(s[3] + (1 - s[3]) * d[3]) << 24;
      return r;
    }
    default: return dst;
  }
}
// SWSpanBrush::BrushH + SWRenderTarget::BlendPixel for any blend mode (sw_span_brush.cc:108-119,
// sw_render_target.cc:12-35; FastBlend's shortcuts :97-141 give the same values as the formulas).
// Modes whose result changes the destination even when the (coverage-scaled) source is zero.  The
// reference's directly emitted spans include pixels whose coverage came out as 0
// (RealSpanBuilder::BuildSpans, sw_raster.cc:45-53), so for these modes such pixels are blended too;
// accumulated spans never carry 0 (SpanBuilder::Flush :111-134).
SKB_HD bool blend_zero_src_matters(uint32_t mode) {
  return mode == 0 || mode == 1 || mode == 5 || mode == 6 || mode == 7 || mode == 10 || mode == 13 || mode == 21;
}
SKB_HD uint32_t paint_blend_mode(const skb_dl_paint& p) { return p.blend ? p.blend - 1 : SKB_BLEND_SRC_OVER; }
SKB_HD uint32_t blend_cover_mode(uint32_t dst, uint32_t color, uint32_t cover, uint32_t mode) {
  if (mode == SKB_BLEND_SRC_OVER) return blend_cover(dst, color, cover);
  if (cover != 255) color = alpha_mul_q(color, cover);
  return porter_duff(color, dst, mode);
}

// ColorFilter::FilterColor on a premultiplied colour in the reference's register order A<<24|R<<16|G<<8|B
// (src/effect/color_filter.cc:123-197; PMColorToColor / ColorToPMColor src/graphic/color_priv.cc:85-108, the
// unpremultiply scale table is round(255 * 2^24 / alpha)).
SKB_HDN uint32_t apply_color_filter(const uint32_t* blk, uint32_t c) {
  const uint32_t type = blk[0];
  if (type == SKB_CF_BLEND) return porter_duff(blk[2], c, blk[1]);  // PorterDuffBlend(filter colour, src, mode)
  // unpremultiply
  const uint32_t a = c >> 24;
  const uint32_t scale = a ? (uint32_t)((0xFF000000u + a / 2) / a) : 0u;
  uint32_t ch[4];  // r g b a
  ch[0] = (uint32_t)(((uint64_t)scale * ((c >> 16) & 0xFF) + (1u << 23)) >> 24);
  ch[1] = (uint32_t)(((uint64_t)scale * ((c >> 8) & 0xFF) + (1u << 23)) >> 24);
  ch[2] = (uint32_t)(((uint64_t)scale * (c & 0xFF) + (1u << 23)) >> 24);
  ch[3] = a;
  uint32_t o[4];
  if (type == SKB_CF_MATRIX) {
    for (int i = 0; i < 4; i++) {
      int32_t m[5];
      for (int j = 0; j < 5; j++) {
        const int k = 5 * i + j;
        m[j] = (int32_t)(int16_t)((blk[4 + (k >> 1)] >> (16 * (k & 1))) & 0xFFFF);
      }
      int32_t mul = (int32_t)ch[0] * m[0] + (int32_t)ch[1] * m[1] + (int32_t)ch[2] * m[2] + (int32_t)ch[3] * m[3];
      int32_t v = mul / 255 + m[4];
      o[i] = (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  } else {  // SKB_CF_TABLE
    for (int i = 0; i < 3; i++) o[i] = (blk[4 + (ch[i] >> 2)] >> (8 * (ch[i] & 3))) & 0xFF;
    o[3] = a;
  }
  // ColorToPMColor
  if (o[3] != 255) {
    o[0] = mul_div_255_round(o[0], o[3]);
    o[1] = mul_div_255_round(o[1], o[3]);
    o[2] = mul_div_255_round(o[2], o[3]);
  }
  return (o[3] << 24) | (o[0] << 16) | (o[1] << 8) | o[2];
}

// atan2f as glibc 2.39 computes it (the fdlibm single-precision algorithm: e_atan2f.c / s_atanf.c, plain float
// operations, no FMA): the sweep gradient's angle has to come out bit-identical to the reference's, because at a
// hard colour stop or at the wrap of a repeating gradient the last bit of the angle decides which colour a pixel
// takes.  CUDA's own atan2f differs in the last place.  Verified against the C library on 2*10^7 random inputs spanning 80 binades
// (tests/test_sim_stages.py checks a sample on every run).  Special values follow the same branches.
SKB_HD uint32_t f_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
SKB_HDN float skb_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                        9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                        4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  const int32_t hx = (int32_t)f_bits(x), ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c800000) {  // |x| >= 2^26
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {  // |x| < 1.1875
      if (ix < 0x3f300000) {
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else if (ix < 0x401c0000) {  // |x| < 2.4375
      id = 2;
      x = (x - 1.5f) / (1.0f + 1.5f * x);
    } else {
      id = 3;
      x = -1.0f / x;
    }
  }
  const float z = x * x, w = z * z;
  const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return hx < 0 ? -r : r;
}
SKB_HDN float skb_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = (int32_t)f_bits(x), ix = hx & 0x7fffffff, hy = (int32_t)f_bits(y), iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return skb_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : (m == 1 ? -pi_o_4 - tiny : (m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny));
    return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? pi + tiny : -pi - tiny));
  }
  if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 24) z = pi_o_2 + 0.5f * pi_lo;       // |y/x| > 2^24
  else if (hx < 0 && k < -26) z = 0.0f;        // |y|/x < -2^26
  else z = skb_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// GradientColorBrush::LerpColor (sw_span_brush.cc:21-32,239-299) -> premultiplied pixel word
SKB_HDN uint32_t gradient_color(const skb_dl_paint& p, const float* pool, float t) {
  const float* colors = pool + p.stop_off;
  const float* stops = colors + 4 * (size_t)p.n_colors;
  if (fabsf(t) <= (1.0f / 4096)) t = 0.0f;
  else if (fabsf(t - 1.0f) <= (1.0f / 4096)) t = 1.0f;
  if (p.tile_mode == 3 && (t < 0.0f || t >= 1.0f)) return 0;
  if (p.tile_mode == 0) {
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  } else if (p.tile_mode == 1) {
    t = t - floorf(t);
  } else if (p.tile_mode == 2) {
    double t1 = (double)(t - 1);
    t = fabsf((float)(t1 - 2 * floor((double)(t - 1) * 0.5) - 1));
  }
  int n = (int)p.n_colors;
  float step = 1.f / (n - 1);
  int si = 0, ei = 1;
  float start = 0.f, end = 0.f;
  int i = 0;
  const bool has_stops = SKB_PAINT_HAS_STOPS(p) != 0;
  bool first = has_stops && t <= stops[0];
  if (!first) {
    for (i = 0; i < n - 1; i++) {
      if (has_stops) { start = stops[i]; end = stops[i + 1]; }
      else { start = step * i; end = step * (i + 1); }
      if (t >= start && t <= end) { si = i; ei = i + 1; break; }
    }
  }
  float c[4];
  if (first) {
    for (int k = 0; k < 4; k++) c[k] = colors[k];
  } else if (i == n - 1 && n > 0) {
    for (int k = 0; k < 4; k++) c[k] = colors[4 * (n - 1) + k];
  } else {
    float total = end - start, value = t - start, mix = 0.5f;
    if (total > 0) mix = value / total;
    for (int k = 0; k < 4; k++) c[k] = colors[4 * si + k] * (1 - mix) + colors[4 * ei + k] * mix;
  }
  return color4f_to_pm_word(c[0], c[1], c[2], c[3]);
}

// u8 -> float -> u8 round trip of the nearest sampler (Color4fFromColor, Color4fToColor; color.cc:44-59)
SKB_HD uint32_t requant(uint32_t c) {
  float f = (float)c / 255.f;
  return unit_to_byte(f);
}

struct SurfaceView { const uint8_t* px; uint32_t w, h, pitch; };  // pitch in bytes

// float -> int the way x86-64 does for (int)f and static_cast<uint32_t>(f) (via 64-bit truncation)
SKB_HD int32_t f2i_trunc(float f) { return f2i(f); }
SKB_HD uint32_t f2u_wrap(float f) {
  if (!(f > -9.2233720e18f && f < 9.2233720e18f)) return 0u;
  return (uint32_t)(long long)f;
}

// BitmapSampler::GetColor (src/graphic/bitmap_sampler.cc:11-108) + PixmapBrush::CalculateColor
// (sw_span_brush.cc:569-579): decal test, tile remap, nearest or bilinear sample, byte round trip, premultiply.
SKB_HD float remap_tile(float t, uint32_t mode) {  // RemapFloatTile (:12-23)
  if (mode == 0) {
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  } else if (mode == 1) {
    t = t - floorf(t);
  } else if (mode == 2) {
    float t1 = t - 1;
    float t2 = (float)((double)t1 - 2 * floor((double)t1 * 0.5) - 1);
    t = fabsf(t2);
  }
  return t;
}
// f2u_wrap for an argument known to be >= 0 (or NaN) and below 2^31 — a remapped unit coordinate times the image
// size: one saturating conversion (NaN -> 0, like f2u_wrap) instead of range tests and a 64-bit one.
SKB_HD uint32_t f2u_nonneg(float f) {
#if defined(__CUDA_ARCH__)
  return __float2uint_rz(f);
#else
  return f2u_wrap(f);
#endif
}
SKB_HD uint32_t image_texel(const SurfaceView& s, float fx_, float fy_, const bool nonneg = false) {  // SampleXY: glm::clamp<uint32_t>(float, 0, n-1)
  uint32_t ix = nonneg ? f2u_nonneg(fx_) : f2u_wrap(fx_), iy = nonneg ? f2u_nonneg(fy_) : f2u_wrap(fy_);
  if (ix > s.w - 1) ix = s.w - 1;
  if (iy > s.h - 1) iy = s.h - 1;
  return *reinterpret_cast<const uint32_t*>(s.px + (size_t)iy * s.pitch + (size_t)ix * 4);  // R | G<<8 | B<<16 | A<<24
}
// The nearest sampler on its own (forced inline: the fine pass runs it in a loop specialised for image paints, with
// `tile_mode` in a register).
SKB_HD uint32_t sample_image_nearest(uint32_t tile_mode, const SurfaceView& s, float u, float v, const uint8_t* requant_lut) {
  const uint32_t xmode = tile_mode & 0xF;
  const uint32_t ymode = (tile_mode & SKB_PAINT_IMAGE_YMODE) ? ((tile_mode >> 4) & 0xF) : xmode;
  if ((xmode == 3 && (u < 0.0f || u >= 1.0f)) || (ymode == 3 && (v < 0.0f || v >= 1.0f))) return 0;
  u = remap_tile(u, xmode);
  v = remap_tile(v, ymode);
  uint32_t r, g, b, a;
  // after the decal test and the remap u and v are in [0, 1] (or NaN)
  const uint32_t t = image_texel(s, u * (float)s.w, v * (float)s.h, true);
  const uint32_t t0 = t & 0xFF, t1 = (t >> 8) & 0xFF, t2 = (t >> 16) & 0xFF, t3 = t >> 24;
  // The sampler's u8 -> float -> u8 round trip (requant(): Color4fFromColor, Color4fToColor) gives every byte back
  // unchanged under IEEE single-precision division and multiplication — all 256 values are checked by
  // tests/test_sim_stages.py::test_sampler_round_trip_is_identity — so the bytes are used as they are.
  (void)requant_lut;
  r = t0; g = t1; b = t2; a = t3;
  // an unpremultiplied texture is premultiplied after sampling (sw_span_brush.cc:573-576)
  if ((tile_mode & SKB_PAINT_IMAGE_UNPREMUL) && a != 255) {
    r = mul_div_255_round(r, a);
    g = mul_div_255_round(g, a);
    b = mul_div_255_round(b, a);
  }
  return r | (g << 8) | (b << 16) | (a << 24);
}
SKB_HDN uint32_t sample_image(const skb_dl_paint& p, const SurfaceView& s, float u, float v, const uint8_t* requant_lut) {
  if (!(p.tile_mode & SKB_PAINT_IMAGE_LINEAR)) return sample_image_nearest(p.tile_mode, s, u, v, requant_lut);
  const uint32_t xmode = p.tile_mode & 0xF;
  const uint32_t ymode = (p.tile_mode & SKB_PAINT_IMAGE_YMODE) ? ((p.tile_mode >> 4) & 0xF) : xmode;
  if ((xmode == 3 && (u < 0.0f || u >= 1.0f)) || (ymode == 3 && (v < 0.0f || v >= 1.0f))) return 0;
  u = remap_tile(u, xmode);
  v = remap_tile(v, ymode);
  uint32_t r, g, b, a;
  {  // SampleUnitLinear (:44-83)
    const float w = (float)s.w, h = (float)s.h;
    float x = u * w, y = v * h;
    float i0 = floorf(x - 0.5f), j0 = floorf(y - 0.5f);
    if (xmode == 1) i0 = i0 - w * floorf(i0 / w);  // glm::mod
    if (ymode == 1) j0 = j0 - h * floorf(j0 / h);
    float i1 = i0 + 1.0f, j1 = j0 + 1.0f;
    if (xmode == 1) i1 = i1 - w * floorf(i1 / w);
    if (ymode == 1) j1 = j1 - h * floorf(j1 / h);
    const float fa = (x - 0.5f) - floorf(x - 0.5f), fb = (y - 0.5f) - floorf(y - 0.5f);  // glm::fract
    const uint32_t t00 = image_texel(s, i0, j0), t10 = image_texel(s, i1, j0), t01 = image_texel(s, i0, j1),
                   t11 = image_texel(s, i1, j1);
    const float w00 = (1 - fa) * (1 - fb), w10 = fa * (1 - fb), w01 = (1 - fa) * fb, w11 = fa * fb;
    uint32_t o[4];
    for (int c = 0; c < 4; c++) {
      const float c00 = (float)((t00 >> (8 * c)) & 0xFF) / 255.f, c10 = (float)((t10 >> (8 * c)) & 0xFF) / 255.f;
      const float c01 = (float)((t01 >> (8 * c)) & 0xFF) / 255.f, c11 = (float)((t11 >> (8 * c)) & 0xFF) / 255.f;
      o[c] = unit_to_byte(((w00 * c00 + w10 * c10) + w01 * c01) + w11 * c11);
    }
    r = o[0]; g = o[1]; b = o[2]; a = o[3];
  }
  // an unpremultiplied texture is premultiplied after sampling (sw_span_brush.cc:573-576)
  if ((p.tile_mode & SKB_PAINT_IMAGE_UNPREMUL) && a != 255) {
    r = mul_div_255_round(r, a);
    g = mul_div_255_round(g, a);
    b = mul_div_255_round(b, a);
  }
  return r | (g << 8) | (b << 16) | (a << 24);
}

// Source colour of paint `p` at pixel centre (x+.5, y+.5): premultiplied pixel word.
// Solid sw_span_brush.cc:140-152, Linear :312-320, Sweep :337-354, Radial :371-379,
// Pixmap :569-579 + bitmap_sampler.cc:26-40,85-108.
// `img` is the surface an IMAGE paint samples (ignored by the other paint types).
// `requant_lut` (optional): requant() of every byte, tabulated by the caller.
SKB_HDN uint32_t paint_color(const skb_dl_paint& p, const float* pool, const SurfaceView& img, int x, int y,
                             const uint8_t* requant_lut = nullptr) {
  if (p.type == SKB_PAINT_SOLID) return color4f_to_pm_word(p.color[0], p.color[1], p.color[2], p.color[3]);
  float fxc = x + 0.5f, fyc = y + 0.5f;
  float u = fxc * p.m[0] + fyc * p.m[1] + p.m[2];
  float v = fxc * p.m[3] + fyc * p.m[4] + p.m[5];
  switch (p.type) {
    case SKB_PAINT_LINEAR:
      return gradient_color(p, pool, u);
    case SKB_PAINT_RADIAL:
      return gradient_color(p, pool, sqrtf(u * u + v * v));
    case SKB_PAINT_SWEEP: {
      float angle = skb_atan2f(-v, -u);  // glibc's, bit for bit
      const float k1Over2Pi = 0.1591549430918f;
      float t = (float)(((double)(angle * k1Over2Pi) + 0.5 + (double)p.bias) * (double)p.scale);
      return gradient_color(p, pool, t);
    }
    case SKB_PAINT_CONICAL: {  // ConicalGradientColorBrush::CalculateConical (sw_span_brush.cc:450-513)
      const float* e = pool + p.stop_off + 5 * (size_t)p.n_colors;
      const int kind = (int)e[0];
      float t;
      if (kind == 1) {
        float qx = (u - e[1]) * e[3], qy = (v - e[2]) * e[3];
        t = sqrtf(qx * qx + qy * qy) * e[4] - e[5];
      } else if (kind == 2) {
        float r = e[6];
        float r_2 = r * r;
        float x2 = u * e[7] + v * e[8] + e[9];
        float y2 = u * e[10] + v * e[11] + e[12];
        t = r_2 - y2 * y2;
        if (t < 0.0f) return 0;
        t = x2 + sqrtf(t);
      } else if (kind == 3 || kind == 5) {
        float x2 = u * e[7] + v * e[8] + e[9];
        float y2 = u * e[10] + v * e[11] + e[12];
        const float r1 = e[13], r1sq = e[14], f = e[15];
        float xt = -1.f;
        if (fabsf(r1 - 1.f) < (1.0f / 4096)) {
          xt = (x2 * x2 + y2 * y2) / 2;
        } else if (r1 > 1.f) {
          float m = r1sq - 1.f;
          float delta = m * y2 * y2 + r1sq * x2 * x2;
          xt = (sqrtf(delta) - x2) / m;
        } else {
          float m = r1sq - 1.f;
          float delta = m * y2 * y2 + r1sq * x2 * x2;
          if (delta > 0) {
            float xt1 = (sqrtf(delta) - x2) / m;
            float xt2 = (-sqrtf(delta) - x2) / m;
            xt = 1.f - f < 0 ? (xt2 < xt1 ? xt2 : xt1) : (xt1 < xt2 ? xt2 : xt1);  // std::min / std::max
          }
        }
        if (xt < 0) return 0;
        t = f + (1.f - f) * xt;
        if (kind == 5) t = 1.0f - t;
      } else {
        return 0;  // negative radius, or concentric circles of equal radius: transparent
      }
      return gradient_color(p, pool, t);
    }
    case SKB_PAINT_IMAGE:
      return sample_image(p, img, u, v, requant_lut);
    default:
      return 0;
  }
}

// SWStackBlur reciprocal (sw_stack_blur.cc:286-338): shr = max s with 2^s/(r+1)^2 <= 512, mul = ceil(2^s/(r+1)^2)
SKB_HD void blur_mul_shr(int radius, uint32_t* mul, int* shr) {
  uint64_t d = (uint64_t)(radius + 1) * (uint64_t)(radius + 1);
  int s = 0;
  while ((1ull << (s + 1)) <= 512ull * d) s++;
  *shr = s;
  *mul = (uint32_t)(((1ull << s) + d - 1) / d);
}

}  // namespace skb

#endif  // SKB_CORE_CUH
