// skb_area.cuh — coverage mode SKB_COVERAGE_AREA: tile-binned lines and per-tile signed-area accumulation with
// backdrop prefix sums (north star stages 2-3), i.e. the algorithm of the reference's own GPU coverage-AA path,
// evaluated in parallel:
//   flattening   PathVisitor::HandleQuadTo / HandleConicTo / HandleCubicTo (src/graphic/path_visitor.cc:112-209) with
//                Wang's formula at precision 4 (src/geometry/wangs_formula.hpp:114-161) on the TRANSFORMED control points
//                (CoverageAAPathTiler::Tile transforms the path first, coverage_aa_tiler.cc:73-94) — one thread per line,
//                the per-segment counts resolved by a device prefix scan;
//   binning      CoverageAAPathTiler::ProcessGlobalLine / AddTileLine / AddLeftBoundaryLine / AddBackdropDelta
//                (coverage_aa_tiler.cc:136-324): a DDA over 16-px tiles per line; clipped lines in unsigned 8.8 per tile,
//                an auxiliary vertical line where a line crosses a tile's left edge, +-1 backdrop deltas where it crosses
//                a horizontal tile boundary, full-height left-edge lines folded into the tile's local backdrop;
//   backdrops    ResolveBackdrops (:111-134): prefix sum of the deltas along each tile row;
//   per pixel    coverage_aa_edge_contribution / coverage_aa_resolve_alpha
//                (src/render/hw/coverage/wgsl_coverage_aa_common.hpp:10-65), fp32, one rounding per operation, the lines
//                of a tile summed in the order the reference's tiler emits them (kept by a per-tile key sort, so the
//                result does not depend on the order the threads binned them in).
// Coverage leaves this stage as the same A8 tile masks the exact mode produces (uint8(alpha * 255 + 0.5)); everything
// after it (bin, fine pass, blur) is shared.  Pure per-thread code (SKB_HD), compiled into the kernels of skb_backend.cu
// and into the CPU simulation (tests/sim/).
#ifndef SKB_AREA_CUH
#define SKB_AREA_CUH

#include "skity_b200/csrc/skb_core.cuh"

namespace skb {

SKB_HD float skb_inf() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7f800000);
#else
  union { uint32_t u; float f; } v;
  v.u = 0x7f800000u;
  return v.f;
#endif
}

#define SKB_AREA_TILE 16
#define SKB_AREA_FIXED_LIMIT 4096   // kCoverageAATileWidth * kCoverageAASubpixelScale

// ---- flattening ---------------------------------------------------------------------------------------------------
// VectorXform() is the identity: fC0 * v.x + fC1 * v.y with fC0 = (1, 0), fC1 = (0, 1) — written out so that the
// arithmetic (signed zeros, NaNs) is the reference's.
SKB_HD float area_idx(float a, float b) { return 1.0f * a + 0.0f * b; }
SKB_HD float area_idy(float a, float b) { return 0.0f * a + 1.0f * b; }
SKB_HD int area_count_of(float num) {   // PathVisitor: num <= 1 -> one line; DEBUG_CHECK(num < 1 << 10)
  if (num <= 1.0f) return 1;
  if (!(num < 1024.f)) num = 1023.f;
  return (int)num;
}
SKB_HDN int area_quad_count(V2 p0, V2 p1, V2 p2) {   // ceil(wangs_formula::Quadratic(4, pts)); LengthTermP2<2>(4) = 1
  const float vx = (-2.0f * p1.x + p0.x) + p2.x, vy = (-2.0f * p1.y + p0.y) + p2.y;
  const float wx = area_idx(vx, vy), wy = area_idy(vx, vy);
  return area_count_of(ceilf(sqrtf(sqrtf((wx * wx + wy * wy) * 1.0f))));
}
SKB_HDN int area_cubic_count(V2 p0, V2 p1, V2 p2, V2 p3) {   // ceil(wangs_formula::Cubic(4, pts)); LengthTermP2<3>(4) = 9
  const float ax = (-2.0f * p1.x + p0.x) + p2.x, ay = (-2.0f * p1.y + p0.y) + p2.y;
  const float bx = (-2.0f * p2.x + p1.x) + p3.x, by = (-2.0f * p2.y + p1.y) + p3.y;
  const float a0 = area_idx(ax, ay), a1 = area_idy(ax, ay), b0 = area_idx(bx, by), b1 = area_idy(bx, by);
  const float m0 = a0 * a0 + a1 * a1, m1 = b0 * b0 + b1 * b1;
  return area_count_of(ceilf(sqrtf(sqrtf((m0 < m1 ? m1 : m0) * 9.0f))));
}

struct AreaSegPts { V2 p0, p1, p2, p3; };
SKB_HD AreaSegPts area_seg_points(const skb_dl_seg& s, const float* ctm) {
  AreaSegPts q;
  q.p0 = xform(ctm, v2(s.p[0], s.p[1]));
  q.p1 = xform(ctm, v2(s.p[2], s.p[3]));
  q.p2 = xform(ctm, v2(s.p[4], s.p[5]));
  q.p3 = xform(ctm, v2(s.p[6], s.p[7]));
  return q;
}

// Lines the segment flattens to (0 for a lone point).
SKB_HDN int area_seg_line_count(const skb_dl_seg& s, const float* ctm) {
  const uint32_t type = s.type_flags & SKB_SEG_TYPE_MASK;
  const AreaSegPts q = area_seg_points(s, ctm);
  switch (type) {
    case SKB_SEG_LINE:
    case SKB_SEG_CLOSE:
      return 1;
    case SKB_SEG_QUAD:
      return area_quad_count(q.p0, q.p1, q.p2);
    case SKB_SEG_CONIC: {   // HandleConicTo: Conic::ChopIntoQuadsPOW2(quads, 1) of the transformed conic, then two quads
      V2 c[5];
      conic_to_quads(q.p0, q.p1, q.p2, s.w, c);
      return area_quad_count(q.p0, c[1], c[2]) + area_quad_count(c[2], c[3], c[4]);
    }
    case SKB_SEG_CUBIC:
      return area_cubic_count(q.p0, q.p1, q.p2, q.p3);
    default:
      return 0;
  }
}

SKB_HD void area_quad_line(V2 p0, V2 p1, V2 p2, int k, int n, V2* from, V2* to) {
  if (n <= 1) { *from = p0; *to = p2; return; }
  const QuadCoeff c = quad_coeff(p0, p1, p2);
  *from = k == 0 ? p0 : quad_eval(c, (float)k / (float)n);
  *to = k + 1 == n ? p2 : quad_eval(c, (float)(k + 1) / (float)n);
}

// k-th of the n lines of a segment.
SKB_HDN void area_seg_line(const skb_dl_seg& s, const float* ctm, int k, int n, V2* from, V2* to) {
  const uint32_t type = s.type_flags & SKB_SEG_TYPE_MASK;
  const AreaSegPts q = area_seg_points(s, ctm);
  switch (type) {
    case SKB_SEG_QUAD:
      area_quad_line(q.p0, q.p1, q.p2, k, n, from, to);
      return;
    case SKB_SEG_CONIC: {
      V2 c[5];
      conic_to_quads(q.p0, q.p1, q.p2, s.w, c);
      const int n1 = area_quad_count(q.p0, c[1], c[2]);
      if (k < n1) area_quad_line(q.p0, c[1], c[2], k, n1, from, to);
      else area_quad_line(c[2], c[3], c[4], k - n1, n - n1, from, to);
      return;
    }
    case SKB_SEG_CUBIC: {
      if (n <= 1) { *from = q.p0; *to = q.p3; return; }
      const CubicCoeff c = cubic_coeff(q.p0, q.p1, q.p2, q.p3);
      *from = k == 0 ? q.p0 : cubic_eval(c, (float)k / (float)n);
      *to = k + 1 == n ? q.p3 : cubic_eval(c, (float)(k + 1) / (float)n);
      return;
    }
    default:   // LINE, CLOSE
      *from = q.p0;
      *to = q.p1;
      return;
  }
}

// ---- binning ------------------------------------------------------------------------------------------------------
SKB_HD uint32_t area_pack_fixed(float value) {   // PackFixed, coverage_aa_tiler.cc:45-49 (std::round: half away from zero)
  int32_t fixed = (int32_t)roundf(value * 256.0f);
  fixed = fixed < 0 ? 0 : fixed;
  fixed = fixed > SKB_AREA_FIXED_LIMIT ? SKB_AREA_FIXED_LIMIT : fixed;
  return (uint32_t)fixed;
}

// A line clipped to tile (tx, ty) in the tile's 8.8 coordinates (AddTileLine :136-181, without the tile-domain test).
// Returns 0: dropped (horizontal after quantisation); 1: a line, words = from_x | from_y << 16, to_x | to_y << 16;
// 2: a full-height line on the tile's left edge, *local = +-1 for the tile's local backdrop.
SKB_HDN int area_tile_line(V2 from, V2 to, int tx, int ty, uint32_t* w0, uint32_t* w1, int* local) {
  const float tile_left = (float)tx * (float)SKB_AREA_TILE, tile_top = (float)ty * (float)SKB_AREA_TILE;
  const uint32_t fx_ = area_pack_fixed(from.x - tile_left), fy_ = area_pack_fixed(from.y - tile_top);
  const uint32_t tx_ = area_pack_fixed(to.x - tile_left), ty_ = area_pack_fixed(to.y - tile_top);
  if (fy_ == ty_) return 0;
  const uint32_t ymin = fy_ < ty_ ? fy_ : ty_, ymax = fy_ < ty_ ? ty_ : fy_;
  if (fx_ == 0 && tx_ == 0 && ymin == 0 && ymax == SKB_AREA_FIXED_LIMIT) {
    *local = fy_ > ty_ ? 1 : -1;
    return 2;
  }
  *w0 = fx_ | (fy_ << 16);
  *w1 = tx_ | (ty_ << 16);
  return 1;
}

SKB_HD V2 area_sample(V2 from, V2 to, float t) {   // Sample :33-43
  if (t == 0.f) return from;
  if (t == 1.f) return to;
  return v2(from.x + (to.x - from.x) * t, from.y + (to.y - from.y) * t);
}

// ProcessGlobalLine (:213-311) as a visitor: for every tile the line passes through, in order,
//   sink.line(a, b, tx, ty, aux)   the clipped line (aux = 0) and, where the reference adds one, the auxiliary line on the
//                                  tile's left edge (aux = 1) — both still in global coordinates;
//   sink.backdrop(tx, ty, delta)   a crossing of a horizontal tile boundary.
// The sink applies the tile domain (AddTileLine's Contains test, AddBackdropDelta's row / column rules).
template <class Sink>
SKB_HDN void area_walk_line(V2 from, V2 to, Sink& sink) {
  if (from.x == to.x && from.y == to.y) return;
  const float tw = (float)SKB_AREA_TILE;
  int tx = (int)floorf(from.x / tw), ty = (int)floorf(from.y / tw);
  const int ttx = (int)floorf(to.x / tw), tty = (int)floorf(to.y / tw);
  const float vx = to.x - from.x, vy = to.y - from.y;
  const int step_x = vx < 0.0f ? -1 : 1, step_y = vy < 0.0f ? -1 : 1;
  const float first_x = (float)(tx + (vx >= 0.0f ? 1 : 0)) * tw;
  const float first_y = (float)(ty + (vy >= 0.0f ? 1 : 0)) * tw;
  const float inf = skb_inf();
  float t_max_x = vx == 0.0f ? inf : (first_x - from.x) / vx;
  float t_max_y = vy == 0.0f ? inf : (first_y - from.y) / vy;
  const float t_delta_x = vx == 0.0f ? inf : fabsf(tw / vx);
  const float t_delta_y = vy == 0.0f ? inf : fabsf(tw / vy);
  V2 cur = from;
  bool has_last = false, last_is_x = true;
  // a line of finite length passes through a bounded number of tiles; the cap only guards against NaN input
  for (int guard = 0; guard < (1 << 22); guard++) {
    const bool next_is_x = t_max_x < t_max_y ? true : (t_max_x > t_max_y ? false : step_x > 0);
    float next_t = next_is_x ? t_max_x : t_max_y;
    if (!(next_t < 1.0f)) next_t = 1.0f;
    const bool has_next = tx != ttx || ty != tty;
    const V2 next = area_sample(from, to, next_t);
    sink.line(cur, next, tx, ty, 0);
    if (step_x < 0 && has_next && next_is_x) {   // AddLeftBoundaryLine(tile, next.y, downward)
      const float left = (float)tx * tw, top = (float)ty * tw, bottom = top + tw;
      float y = next.y < top ? top : next.y;
      y = y > bottom ? bottom : y;
      sink.line(v2(left, y), v2(left, bottom), tx, ty, 1);
    } else if (step_x > 0 && has_last && last_is_x) {   // AddLeftBoundaryLine(tile, cur.y, upward)
      const float left = (float)tx * tw, top = (float)ty * tw, bottom = top + tw;
      float y = cur.y < top ? top : cur.y;
      y = y > bottom ? bottom : y;
      sink.line(v2(left, bottom), v2(left, y), tx, ty, 1);
    }
    if (step_y < 0 && has_next && !next_is_x) sink.backdrop(tx, ty, 1);
    else if (step_y > 0 && has_last && !last_is_x) sink.backdrop(tx, ty, -1);
    if (!has_next) break;
    if (next_is_x) {
      if (tx == ttx) break;
      t_max_x += t_delta_x;
      tx += step_x;
    } else {
      if (ty == tty) break;
      t_max_y += t_delta_y;
      ty += step_y;
    }
    cur = next;
    last_is_x = next_is_x;
    has_last = true;
  }
}

// ---- per pixel ----------------------------------------------------------------------------------------------------
// One binned line, unpacked for the pixels of a tile: the parts of coverage_aa_edge_contribution that do not depend
// on the pixel.
struct AreaLine {
  float fx_, fy_, tx_, ty_;   // tile coordinates (8.8 / 256)
  float dx, dy, sign, y_slope, x_slope, edge_top, edge_bottom, left_endpoint_y;
};
SKB_HD AreaLine area_line_unpack(uint32_t w0, uint32_t w1) {
  AreaLine l;
  l.fx_ = (float)(w0 & 0xFFFFu) / 256.0f;
  l.fy_ = (float)(w0 >> 16) / 256.0f;
  l.tx_ = (float)(w1 & 0xFFFFu) / 256.0f;
  l.ty_ = (float)(w1 >> 16) / 256.0f;
  l.edge_top = l.fy_ < l.ty_ ? l.fy_ : l.ty_;
  l.edge_bottom = l.fy_ < l.ty_ ? l.ty_ : l.fy_;
  l.dx = l.tx_ - l.fx_;
  l.dy = l.ty_ - l.fy_;
  l.sign = l.dy < 0.0f ? 1.0f : -1.0f;
  l.y_slope = l.dy / l.dx;   // unused (inf / nan) when dx == 0: that case returns before it is read
  l.x_slope = l.dx / l.dy;
  l.left_endpoint_y = l.fx_ <= l.tx_ ? l.fy_ : l.ty_;
  return l;
}
SKB_HD float area_clampf(float v, float lo, float hi) {   // WGSL clamp(e, low, high) = min(max(e, low), high)
  const float m = v < lo ? lo : v;
  return m > hi ? hi : m;
}
// coverage_aa_edge_contribution for pixel (px, py) of the tile, given the line's y overlap with the pixel row
// [y_min, y_max) (non-empty).
SKB_HD float area_edge_contribution(const AreaLine& l, float px, float y_min, float y_max) {
  const float pixel_left = px, pixel_right = px + 1.0f;
  if (l.dx == 0.0f) {
    const float h = y_max - y_min;
    const float covered_width = area_clampf(pixel_right - l.fx_, 0.0f, 1.0f);
    return l.sign * h * covered_width;
  }
  const float lpy = area_clampf(l.fy_ + (pixel_left - l.fx_) * l.y_slope, y_min, y_max);
  const float rpy = area_clampf(l.fy_ + (pixel_right - l.fx_) * l.y_slope, y_min, y_max);
  const float h = fabsf(rpy - lpy);
  const float lpx = l.fx_ + (lpy - l.fy_) * l.x_slope, rpx = l.fx_ + (rpy - l.fy_) * l.x_slope;
  const float area = h * (pixel_right - 0.5f * (lpx + rpx));
  const float cover = fabsf(lpy - area_clampf(l.left_endpoint_y, y_min, y_max));
  return l.sign * (cover + area);
}
// coverage_aa_resolve_alpha (WGSL round = half to even) and our A8 quantisation
SKB_HD float area_resolve_alpha(float winding, int even_odd) {
  float a;
  if (even_odd) {
    const float even_winding = 2.0f * rintf(0.5f * winding);
    a = fabsf(winding - even_winding);
  } else {
    a = fabsf(winding);
  }
  return a < 1.0f ? a : 1.0f;
}
SKB_HD uint32_t area_alpha_u8(float a) { return (uint32_t)(uint8_t)(a * 255.0f + 0.5f); }

// Coverage of pixel (px, py) of a tile from its lines in key order (w0, w1 pairs) and its backdrop: the reference's
// coverage_aa_resolve_pixel.  The plain per-pixel form (CPU simulation, and the check of the kernel's row-wise form).
SKB_HDN uint32_t area_pixel(const uint32_t* words, int n_lines, int stride_words, int backdrop, int even_odd, int px, int py) {
  if (n_lines == 0) return area_alpha_u8(area_resolve_alpha((float)backdrop, even_odd));
  float winding = (float)backdrop;
  const float pixel_top = (float)py, pixel_bottom = (float)py + 1.0f;
  for (int k = 0; k < n_lines; k++) {
    const AreaLine l = area_line_unpack(words[(size_t)k * stride_words], words[(size_t)k * stride_words + 1]);
    const float y_min = l.edge_top < pixel_top ? pixel_top : l.edge_top;
    const float y_max = l.edge_bottom < pixel_bottom ? l.edge_bottom : pixel_bottom;
    if (y_min >= y_max) continue;   // contribution 0.0: winding + 0.0 == winding
    winding = winding + area_edge_contribution(l, (float)px, y_min, y_max);
  }
  return area_alpha_u8(area_resolve_alpha(winding, even_odd));
}

}  // namespace skb

#endif  // SKB_AREA_CUH
