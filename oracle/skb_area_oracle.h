/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Included once by skb_oracle.c (uses its v2 / xform / curve helpers).
 *
 * Plain-C restatement of the reference's COVERAGE-AA path (its GPU backends' analytic coverage: tile-binned lines,
 * signed-area accumulation, backdrop prefix sums) — what the CUDA backend runs in SKB_COVERAGE_AREA mode
 * (north star stages 2-3).  Sequential and in the reference's own data structures, so that it shares nothing with
 * the device decomposition it checks:
 *   flattening    PathVisitor::VisitPath / HandleQuadTo / HandleConicTo / HandleCubicTo, src/graphic/path_visitor.cc:46-211,
 *                 Wang's formula src/geometry/wangs_formula.hpp:114-161 (precision 4, identity vector transform: the
 *                 path is transformed first, CoverageAAPathTiler::Tile coverage_aa_tiler.cc:73-94)
 *   tiling        CoverageAAPathTiler::Reset / ProcessGlobalLine / AddTileLine / AddLeftBoundaryLine / AddBackdropDelta /
 *                 ResolveBackdrops, src/render/hw/coverage/coverage_aa_tiler.cc:96-324; lines grouped per tile in
 *                 emission order (EncodeCoverageAALines, coverage_aa_line_encoder.cc:39-75)
 *   per pixel     coverage_aa_edge_contribution / coverage_aa_resolve_alpha / coverage_aa_resolve_pixel,
 *                 src/render/hw/coverage/wgsl_coverage_aa_common.hpp:10-103 (the variant without conflation
 *                 correction, the default), in fp32 with one rounding per operation, lines summed in range order
 * PINNING: the tiler part is checked against the reference's own compiled CoverageAAPathTiler (oracle/_ref,
 * ref_coverage_aa_tile) tile by tile and line by line on random paths and on the shapes of the reference's tiler
 * unit tests (test/ut/render/hw/coverage_aa_path_tiler_test.cc); the per-pixel part against the reference's golden
 * image coverage_aa_images/canonical_edges_exact.png (exact-match rule, test/golden/cases/shape/shape.cc:666-672).
 * The WGSL itself cannot run here (no GPU API): its fp32 evaluation order is restated as written.
 * What is ours, not the reference's: coverage is quantised to A8 as uint8(alpha * 255 + 0.5) and handed to the same
 * span brush as the software path (the reference multiplies alpha into the fragment colour in float).
 */
#ifndef SKB_AREA_ORACLE_H
#define SKB_AREA_ORACLE_H

#define AREA_TILE 16
#define AREA_SUBPX 256
#define AREA_FIXED_LIMIT (AREA_TILE * AREA_SUBPX)
#define AREA_NO_RANGE 0xFFFFFFFFu

typedef struct { v2 from, to; } area_gline;
typedef struct { uint16_t from_x, from_y, to_x, to_y; uint32_t range; } area_tline;
typedef struct { uint32_t range; int16_t backdrop_delta, local_backdrop; } area_tstate;
typedef struct { int32_t tile_x, tile_y; uint32_t range; int32_t backdrop; } area_tile;
typedef struct {
  int ox, oy, w, h; /* tile_bounds_ */
  int32_t* row_backdrops;
  area_tstate* states;
  area_tline* lines; size_t n_lines, cap_lines;
  uint32_t* range_counts; size_t n_ranges, cap_ranges;
  area_tile* tiles; size_t n_tiles, cap_tiles;
} area_tiler;

static void area_tiler_free(area_tiler* t) {
  free(t->row_backdrops); free(t->states); free(t->lines); free(t->range_counts); free(t->tiles);
  memset(t, 0, sizeof(*t));
}

/* PackFixed — coverage_aa_tiler.cc:45-49 */
static uint16_t area_pack_fixed(float value) {
  int32_t fixed = (int32_t)roundf(value * (float)AREA_SUBPX);
  if (fixed < 0) fixed = 0;
  if (fixed > AREA_FIXED_LIMIT) fixed = AREA_FIXED_LIMIT;
  return (uint16_t)fixed;
}
static int area_contains(const area_tiler* t, int tx, int ty) {
  int64_t dx = (int64_t)tx - t->ox, dy = (int64_t)ty - t->oy;
  return dx >= 0 && dy >= 0 && dx < t->w && dy < t->h;
}
static size_t area_index(const area_tiler* t, int tx, int ty) { return (size_t)(ty - t->oy) * (size_t)t->w + (size_t)(tx - t->ox); }

/* AddTileLine — coverage_aa_tiler.cc:136-181 */
static void area_add_tile_line(area_tiler* t, area_gline line, int tx, int ty) {
  if (!area_contains(t, tx, ty)) return;
  float tile_left = (float)tx * AREA_TILE, tile_top = (float)ty * AREA_TILE;
  area_tline tl;
  tl.from_x = area_pack_fixed(line.from.x - tile_left);
  tl.from_y = area_pack_fixed(line.from.y - tile_top);
  tl.to_x = area_pack_fixed(line.to.x - tile_left);
  tl.to_y = area_pack_fixed(line.to.y - tile_top);
  tl.range = AREA_NO_RANGE;
  if (tl.from_y == tl.to_y) return;
  int on_left = tl.from_x == 0 && tl.to_x == 0;
  uint16_t ymin = tl.from_y < tl.to_y ? tl.from_y : tl.to_y, ymax = tl.from_y < tl.to_y ? tl.to_y : tl.from_y;
  area_tstate* st = &t->states[area_index(t, tx, ty)];
  if (on_left && ymin == 0 && ymax == AREA_FIXED_LIMIT) {
    st->local_backdrop += tl.from_y > tl.to_y ? 1 : -1;
    return;
  }
  if (st->range == AREA_NO_RANGE) { /* GetOrCreateLineRangeId :200-211 */
    if (t->n_ranges == t->cap_ranges) {
      t->cap_ranges = t->cap_ranges ? t->cap_ranges * 2 : 64;
      t->range_counts = (uint32_t*)realloc(t->range_counts, t->cap_ranges * sizeof(uint32_t));
    }
    st->range = (uint32_t)t->n_ranges;
    t->range_counts[t->n_ranges++] = 0;
  }
  tl.range = st->range;
  t->range_counts[tl.range]++;
  if (t->n_lines == t->cap_lines) {
    t->cap_lines = t->cap_lines ? t->cap_lines * 2 : 256;
    t->lines = (area_tline*)realloc(t->lines, t->cap_lines * sizeof(area_tline));
  }
  t->lines[t->n_lines++] = tl;
}

/* AddBackdropDelta — :183-198 */
static void area_add_backdrop_delta(area_tiler* t, int tx, int ty, int32_t delta) {
  int ofx = tx - t->ox, ofy = ty - t->oy;
  if (ofy < 0 || ofy >= t->h || ofx >= t->w) return;
  if (ofx < 0) { t->row_backdrops[ofy] += delta; return; }
  t->states[area_index(t, tx, ty)].backdrop_delta += (int16_t)delta;
}

/* AddLeftBoundaryLine — :313-324 */
static void area_add_left_boundary_line(area_tiler* t, int tx, int ty, float crossing_y, int upward) {
  float left = (float)tx * AREA_TILE, top = (float)ty * AREA_TILE, bottom = top + AREA_TILE;
  float y = crossing_y < top ? top : crossing_y;
  y = y > bottom ? bottom : y; /* std::clamp(v, lo, hi) */
  area_gline b;
  if (upward) { b.from = V(left, bottom); b.to = V(left, y); }
  else { b.from = V(left, y); b.to = V(left, bottom); }
  area_add_tile_line(t, b, tx, ty);
}

static v2 area_sample(area_gline l, float tt) { /* Sample — :33-43 */
  if (tt == 0.f) return l.from;
  if (tt == 1.f) return l.to;
  return V(l.from.x + (l.to.x - l.from.x) * tt, l.from.y + (l.to.y - l.from.y) * tt);
}

/* ProcessGlobalLine — :213-311 */
static void area_process_global_line(area_tiler* t, area_gline line) {
  if (line.from.x == line.to.x && line.from.y == line.to.y) return;
  int ftx = (int)floorf(line.from.x / (float)AREA_TILE), fty = (int)floorf(line.from.y / (float)AREA_TILE);
  int ttx = (int)floorf(line.to.x / (float)AREA_TILE), tty = (int)floorf(line.to.y / (float)AREA_TILE);
  float vx = line.to.x - line.from.x, vy = line.to.y - line.from.y;
  int step_x = vx < 0.0f ? -1 : 1, step_y = vy < 0.0f ? -1 : 1;
  float first_x = (float)(ftx + (vx >= 0.0f ? 1 : 0)) * AREA_TILE;
  float first_y = (float)(fty + (vy >= 0.0f ? 1 : 0)) * AREA_TILE;
  float t_max_x = vx == 0.0f ? INFINITY : (first_x - line.from.x) / vx;
  float t_max_y = vy == 0.0f ? INFINITY : (first_y - line.from.y) / vy;
  float t_delta_x = vx == 0.0f ? INFINITY : fabsf((float)AREA_TILE / vx);
  float t_delta_y = vy == 0.0f ? INFINITY : fabsf((float)AREA_TILE / vy);
  v2 cur = line.from;
  int tx = ftx, ty = fty;
  int has_last = 0, last_is_x = 1;
  for (;;) {
    int next_is_x = t_max_x < t_max_y ? 1 : (t_max_x > t_max_y ? 0 : (step_x > 0 ? 1 : 0));
    float next_t = next_is_x ? t_max_x : t_max_y;
    if (!(next_t < 1.0f)) next_t = 1.0f; /* std::min(v, 1.0f) */
    int has_next = tx != ttx || ty != tty;
    v2 next = area_sample(line, next_t);
    area_gline clipped;
    clipped.from = cur;
    clipped.to = next;
    area_add_tile_line(t, clipped, tx, ty);
    if (step_x < 0 && has_next && next_is_x) area_add_left_boundary_line(t, tx, ty, next.y, 0);
    else if (step_x > 0 && has_last && last_is_x) area_add_left_boundary_line(t, tx, ty, cur.y, 1);
    if (step_y < 0 && has_next && !next_is_x) area_add_backdrop_delta(t, tx, ty, 1);
    else if (step_y > 0 && has_last && !last_is_x) area_add_backdrop_delta(t, tx, ty, -1);
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
    has_last = 1;
  }
}

/* Wang's formula at precision 4 with the identity vector transform — wangs_formula.hpp:114-161 (LengthTermP2<2>(4) = 1,
 * LengthTermP2<3>(4) = 9), Root4 = sqrtf(sqrtf(x)); the visitor takes the ceiling (path_visitor.cc:126-131,188-193). */
static float area_xf(float a, float b) { return 1.0f * a + 0.0f * b; } /* VectorXform()(v): fC0 * v.x + fC1 * v.y */
static float area_yf(float a, float b) { return 0.0f * a + 1.0f * b; }
static float area_wang_quad(v2 p0, v2 p1, v2 p2) {
  float vx = (-2.0f * p1.x + p0.x) + p2.x, vy = (-2.0f * p1.y + p0.y) + p2.y;
  float wx = area_xf(vx, vy), wy = area_yf(vx, vy);
  return ceilf(sqrtf(sqrtf((wx * wx + wy * wy) * 1.0f)));
}
static float area_wang_cubic(v2 p0, v2 p1, v2 p2, v2 p3) {
  float ax = (-2.0f * p1.x + p0.x) + p2.x, ay = (-2.0f * p1.y + p0.y) + p2.y;
  float bx = (-2.0f * p2.x + p1.x) + p3.x, by = (-2.0f * p2.y + p1.y) + p3.y;
  float a0 = area_xf(ax, ay), a1 = area_yf(ax, ay), b0 = area_xf(bx, by), b1 = area_yf(bx, by);
  float m0 = a0 * a0 + a1 * a1, m1 = b0 * b0 + b1 * b1;
  return ceilf(sqrtf(sqrtf((m0 < m1 ? m1 : m0) * 9.0f))); /* std::max(a, b) = a < b ? b : a */
}

static void area_emit(area_tiler* t, v2 a, v2 b) {
  area_gline l;
  l.from = a;
  l.to = b;
  area_process_global_line(t, l);
}
/* HandleQuadTo — path_visitor.cc:112-152 */
static void area_quad(area_tiler* t, v2 p1, v2 p2, v2 p3) {
  float num = area_wang_quad(p1, p2, p3);
  if (num <= 1.0f) { area_emit(t, p1, p3); return; }
  if (!(num < 1024.f)) num = 1023.f; /* the reference DEBUG_CHECKs num < 1 << 10; guard for release inputs */
  int n = (int)num;
  quad_coeff c = quad_coeff_make(p1, p2, p3);
  v2 prev = p1;
  for (int i = 1; i <= n; i++) {
    float tt = (float)i / n;
    v2 cur = i == n ? p3 : quad_eval(&c, tt);
    area_emit(t, prev, cur);
    prev = cur;
  }
}
/* HandleCubicTo — :176-209 */
static void area_cubic(area_tiler* t, v2 p1, v2 p2, v2 p3, v2 p4) {
  float num = area_wang_cubic(p1, p2, p3, p4);
  if (num <= 1.0f) { area_emit(t, p1, p4); return; }
  if (!(num < 1024.f)) num = 1023.f;
  int n = (int)num;
  cubic_coeff c = cubic_coeff_make(p1, p2, p3, p4);
  v2 prev = p1;
  for (int i = 1; i <= n; i++) {
    float tt = (float)i / n;
    v2 cur = i == n ? p4 : cubic_eval(&c, tt);
    area_emit(t, prev, cur);
    prev = cur;
  }
}

/* The tiled path of one draw: CoverageAAPathTiler::Tile (:73-94) on the display list's segments (the source path's
 * verbs with implicit closes materialised; Path::Iter with force_close walks the same lines).  `scissor4` = l t r b or
 * NULL.  Returns 0 when nothing is to be drawn. */
static int area_tile_path(area_tiler* t, const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* scissor4) {
  memset(t, 0, sizeof(*t));
  /* Path::CopyWithMatrix + GetBounds: bounds of every point of the transformed path */
  int have = 0;
  float l = 0, tp = 0, r = 0, b = 0;
  for (uint32_t i = 0; i < n_segs; i++) {
    const skb_dl_seg* s = &segs[i];
    uint32_t type = s->type_flags & SKB_SEG_TYPE_MASK;
    int first = 0, last = -1;
    if (type == SKB_SEG_POINT) { v2 q = xform(ctm, V(s->start[0], s->start[1])); if (!have) { l = r = q.x; tp = b = q.y; have = 1; } else { if (q.x < l) l = q.x; if (q.x > r) r = q.x; if (q.y < tp) tp = q.y; if (q.y > b) b = q.y; } continue; }
    if (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) last = 1;
    else if (type == SKB_SEG_QUAD || type == SKB_SEG_CONIC) last = 2;
    else if (type == SKB_SEG_CUBIC) last = 3;
    for (int k = first; k <= last; k++) {
      v2 q = xform(ctm, V(s->p[2 * k], s->p[2 * k + 1]));
      if (!have) { l = r = q.x; tp = b = q.y; have = 1; }
      else { if (q.x < l) l = q.x; if (q.x > r) r = q.x; if (q.y < tp) tp = q.y; if (q.y > b) b = q.y; }
    }
  }
  if (!have || !(l < r && tp < b)) return 0; /* Rect::IsEmpty */
  if (scissor4) { /* Rect::Intersect — rect.cc:160-172 */
    float il = l > scissor4[0] ? l : scissor4[0], ir = r < scissor4[2] ? r : scissor4[2];
    float it = tp > scissor4[1] ? tp : scissor4[1], ib = b < scissor4[3] ? b : scissor4[3];
    if (!(il < ir && it < ib)) return 0;
    l = il; tp = it; r = ir; b = ib;
  }
  /* Reset — :96-109 */
  int min_x = (int)floorf(l / AREA_TILE), min_y = (int)floorf(tp / AREA_TILE);
  int max_x = (int)ceilf(r / AREA_TILE), max_y = (int)ceilf(b / AREA_TILE);
  t->ox = min_x; t->oy = min_y; t->w = max_x - min_x; t->h = max_y - min_y;
  if (t->w <= 0 || t->h <= 0) return 0;
  t->row_backdrops = (int32_t*)calloc((size_t)t->h, sizeof(int32_t));
  t->states = (area_tstate*)calloc((size_t)t->w * t->h, sizeof(area_tstate));
  for (size_t i = 0; i < (size_t)t->w * t->h; i++) t->states[i].range = AREA_NO_RANGE;
  /* PathTilingVisitor: every line of the flattened, transformed path */
  for (uint32_t i = 0; i < n_segs; i++) {
    const skb_dl_seg* s = &segs[i];
    uint32_t type = s->type_flags & SKB_SEG_TYPE_MASK;
    v2 p0 = xform(ctm, V(s->p[0], s->p[1])), p1 = xform(ctm, V(s->p[2], s->p[3]));
    v2 p2 = xform(ctm, V(s->p[4], s->p[5])), p3 = xform(ctm, V(s->p[6], s->p[7]));
    switch (type) {
      case SKB_SEG_LINE: case SKB_SEG_CLOSE: area_emit(t, p0, p1); break;
      case SKB_SEG_QUAD: area_quad(t, p0, p1, p2); break;
      case SKB_SEG_CONIC: { /* HandleConicTo :154-174: ChopIntoQuadsPOW2(quads, 1) of the transformed conic */
        v2 q[5];
        conic_to_quads(p0, p1, p2, s->w, q);
        q[0] = p0;
        area_quad(t, q[0], q[1], q[2]);
        area_quad(t, q[2], q[3], q[4]);
      } break;
      case SKB_SEG_CUBIC: area_cubic(t, p0, p1, p2, p3); break;
      default: break;
    }
  }
  return 1;
}

/* ResolveBackdrops — :111-134 */
static void area_resolve_backdrops(area_tiler* t, int even_odd) {
  for (int y = 0; y < t->h; y++) {
    int32_t acc = t->row_backdrops[y];
    for (int x = 0; x < t->w; x++) {
      area_tstate* st = &t->states[(size_t)y * t->w + x];
      int32_t backdrop = acc + st->local_backdrop;
      acc += st->backdrop_delta;
      int has = even_odd ? (backdrop % 2 != 0) : (backdrop != 0);
      if (st->range != AREA_NO_RANGE || has) {
        if (t->n_tiles == t->cap_tiles) {
          t->cap_tiles = t->cap_tiles ? t->cap_tiles * 2 : 64;
          t->tiles = (area_tile*)realloc(t->tiles, t->cap_tiles * sizeof(area_tile));
        }
        area_tile* o = &t->tiles[t->n_tiles++];
        o->tile_x = t->ox + x; o->tile_y = t->oy + y; o->range = st->range; o->backdrop = backdrop;
      }
    }
  }
}

/* coverage_aa_edge_contribution — wgsl_coverage_aa_common.hpp:10-54 */
static float area_clampf(float v, float lo, float hi) { float m = v < lo ? lo : v; return m > hi ? hi : m; } /* WGSL clamp = min(max(e, low), high) */
static float area_edge_contribution(float fx_, float fy_, float tx_, float ty_, float px, float py) {
  float pixel_left = px, pixel_top = py, pixel_right = pixel_left + 1.0f, pixel_bottom = pixel_top + 1.0f;
  float edge_top = fy_ < ty_ ? fy_ : ty_, edge_bottom = fy_ < ty_ ? ty_ : fy_;
  float y_min = edge_top < pixel_top ? pixel_top : edge_top, y_max = edge_bottom < pixel_bottom ? edge_bottom : pixel_bottom;
  if (y_min >= y_max) return 0.0f;
  float dx = tx_ - fx_, dy = ty_ - fy_;
  float sign = dy < 0.0f ? 1.0f : -1.0f;
  if (dx == 0.0f) {
    float h = y_max - y_min;
    float covered_width = area_clampf(pixel_right - fx_, 0.0f, 1.0f);
    return sign * h * covered_width;
  }
  float y_slope = dy / dx, x_slope = dx / dy;
  float lpy = area_clampf(fy_ + (pixel_left - fx_) * y_slope, y_min, y_max);
  float rpy = area_clampf(fy_ + (pixel_right - fx_) * y_slope, y_min, y_max);
  float h = fabsf(rpy - lpy);
  float lpx = fx_ + (lpy - fy_) * x_slope, rpx = fx_ + (rpy - fy_) * x_slope;
  float area = h * (pixel_right - 0.5f * (lpx + rpx));
  float left_endpoint_y = fx_ <= tx_ ? fy_ : ty_;
  float cover = fabsf(lpy - area_clampf(left_endpoint_y, y_min, y_max));
  return sign * (cover + area);
}
/* coverage_aa_resolve_alpha — :57-65 (WGSL round: half to even) */
static float area_resolve_alpha(float winding, int even_odd) {
  if (even_odd) {
    float even_winding = 2.0f * rintf(0.5f * winding);
    float a = fabsf(winding - even_winding);
    return a < 1.0f ? a : 1.0f;
  }
  float a = fabsf(winding);
  return a < 1.0f ? a : 1.0f;
}
static uint8_t area_alpha_u8(float a) { return (uint8_t)(a * 255.0f + 0.5f); }

/* The draw's coverage as spans (runs of equal non-zero A8 per pixel row), restricted to the integer scan rectangle
 * floor/ceil(path bounds ∩ clip) ∩ surface — the pixels the software path may touch for the same draw.  Lines of a
 * tile are summed in the order the tiler emitted them (EncodeCoverageAALines keeps it). */
static void area_raster_path(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd,
                             int surf_w, int surf_h, spanvec* out) {
  area_tiler t;
  if (!area_tile_path(&t, segs, n_segs, ctm, clip)) { area_tiler_free(&t); return; }
  area_resolve_backdrops(&t, even_odd);
  /* group the lines by range, keeping their order */
  uint32_t* off = (uint32_t*)calloc(t.n_ranges + 1, sizeof(uint32_t));
  for (size_t i = 0; i < t.n_ranges; i++) off[i + 1] = off[i] + t.range_counts[i];
  uint32_t* cursor = (uint32_t*)malloc((t.n_ranges + 1) * sizeof(uint32_t));
  memcpy(cursor, off, (t.n_ranges + 1) * sizeof(uint32_t));
  area_tline* sorted = (area_tline*)malloc((t.n_lines + 1) * sizeof(area_tline));
  for (size_t i = 0; i < t.n_lines; i++) sorted[cursor[t.lines[i].range]++] = t.lines[i];
  /* the scan rectangle (same rule as SWRaster::RastePath applies to its bounds, sw_raster.cc:741-780) */
  int sx0, sy0, sx1, sy1;
  {
    /* bounds ∩ clip were taken in area_tile_path; recompute the float rectangle */
    int have = 0;
    float l = 0, tp = 0, r = 0, b = 0;
    for (uint32_t i = 0; i < n_segs; i++) {
      const skb_dl_seg* s = &segs[i];
      uint32_t type = s->type_flags & SKB_SEG_TYPE_MASK;
      int last = -1;
      if (type == SKB_SEG_POINT) { v2 q = xform(ctm, V(s->start[0], s->start[1])); if (!have) { l = r = q.x; tp = b = q.y; have = 1; } else { if (q.x < l) l = q.x; if (q.x > r) r = q.x; if (q.y < tp) tp = q.y; if (q.y > b) b = q.y; } continue; }
      if (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) last = 1;
      else if (type == SKB_SEG_QUAD || type == SKB_SEG_CONIC) last = 2;
      else if (type == SKB_SEG_CUBIC) last = 3;
      for (int k = 0; k <= last; k++) {
        v2 q = xform(ctm, V(s->p[2 * k], s->p[2 * k + 1]));
        if (!have) { l = r = q.x; tp = b = q.y; have = 1; }
        else { if (q.x < l) l = q.x; if (q.x > r) r = q.x; if (q.y < tp) tp = q.y; if (q.y > b) b = q.y; }
      }
    }
    float il = l > clip[0] ? l : clip[0], ir = r < clip[2] ? r : clip[2];
    float it = tp > clip[1] ? tp : clip[1], ib = b < clip[3] ? b : clip[3];
    sx0 = (int)floorf(il); sy0 = (int)floorf(it); sx1 = (int)ceilf(ir); sy1 = (int)ceilf(ib);
    if (sx0 < 0) sx0 = 0;
    if (sy0 < 0) sy0 = 0;
    if (sx1 > surf_w) sx1 = surf_w;
    if (sy1 > surf_h) sy1 = surf_h;
  }
  /* pixels, tile by tile; spans are emitted row-major per tile (order between pixels does not matter to the brush) */
  for (size_t ti = 0; ti < t.n_tiles; ti++) {
    const area_tile* tl = &t.tiles[ti];
    uint32_t n = tl->range == AREA_NO_RANGE ? 0 : t.range_counts[tl->range], o = tl->range == AREA_NO_RANGE ? 0 : off[tl->range];
    for (int py = 0; py < AREA_TILE; py++) {
      int y = tl->tile_y * AREA_TILE + py;
      if (y < sy0 || y >= sy1) continue;
      if (g_band_y1 > g_band_y0 && (y < g_band_y0 || y >= g_band_y1)) continue;
      int run_x = 0, run_len = 0, run_cover = 0;
      for (int px = 0; px < AREA_TILE; px++) {
        int x = tl->tile_x * AREA_TILE + px;
        int a8 = 0;
        if (x >= sx0 && x < sx1) {
          float alpha;
          if (n == 0) {
            alpha = area_resolve_alpha((float)tl->backdrop, even_odd);
          } else {
            float winding = (float)tl->backdrop;
            for (uint32_t k = 0; k < n; k++) {
              const area_tline* ln = &sorted[o + k];
              winding = winding + area_edge_contribution((float)ln->from_x / 256.0f, (float)ln->from_y / 256.0f,
                                                         (float)ln->to_x / 256.0f, (float)ln->to_y / 256.0f, (float)px, (float)py);
            }
            alpha = area_resolve_alpha(winding, even_odd);
          }
          a8 = area_alpha_u8(alpha);
        }
        if (run_len && a8 == run_cover) { run_len++; continue; }
        if (run_len && run_cover) sv_push(out, run_x, y, run_len, run_cover);
        run_x = x; run_len = 1; run_cover = a8;
      }
      if (run_len && run_cover) sv_push(out, run_x, y, run_len, run_cover);
    }
  }
  free(off); free(cursor); free(sorted);
  area_tiler_free(&t);
}

#endif /* SKB_AREA_ORACLE_H */
