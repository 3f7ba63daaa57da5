// skb_stages.cuh — the per-thread bodies of the flatten / setup / coverage stages,
// shared by the CUDA kernels (skb_kernels.cu) and the CPU simulation (tests/sim/).
#ifndef SKB_STAGES_CUH
#define SKB_STAGES_CUH

#include <cstddef>

#include "skity_b200/csrc/skb_core.cuh"
#include "skity_b200/csrc/skb_walk.cuh"

namespace skb {

// ---- order-preserving float <-> int keys for atomicMin/atomicMax on path bounds ----------------
SKB_HD int32_t float_key(float f) {
  int32_t i;
#if defined(__CUDA_ARCH__)
  i = __float_as_int(f);
#else
  union { float f; int32_t i; } u;
  u.f = f;
  i = u.i;
#endif
  return i >= 0 ? i : (i ^ 0x7FFFFFFF);
}
SKB_HD float key_float(int32_t k) {
  int32_t i = k >= 0 ? k : (k ^ 0x7FFFFFFF);
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  union { float f; int32_t i; } u;
  u.i = i;
  return u.f;
#endif
}

// Geometry of one raster op after SWRaster::RastePath's bounds logic (sw_raster.cc:741-780).
struct OpGeom {
  int32_t bmin_x, bmin_y, bmax_x, bmax_y;  // float_key()s of the transformed path bounds (atomics)
  int32_t scan_l, scan_t, scan_r, scan_b;  // integer scan rectangle (floor/ceil of bounds ∩ clip)
  float scan_top_f, scan_bottom_f;         // same, as the floats CanBeIgnored sees
  int32_t start_y, stop_y;                 // WalkEdges start_y (bounds top) / stop_y (scan bottom)
  fx left_clip, right_clip;
  int32_t empty;                           // nothing to rasterise
  // tiles of the target surface overlapped by the scan rectangle
  int32_t tx0, ty0, ntx, nty;
  uint32_t slot_base;                      // first edge slot (2 sentinels + 2 per primitive)
  uint32_t n_slots;
  uint32_t row_base;                       // first entry in the row table
  uint32_t item_base;                      // first (op, tile) work item
  uint32_t first_prim;
  uint32_t color;                          // SOLID paints: premultiplied colour, A<<24|R<<16|G<<8|B
  uint32_t fast_solid;                     // SOLID paint blended kSrcOver: the fine pass needs nothing but `color`
  uint32_t culled;                         // the draw cannot reach this device's band: it was given no primitives (k_op_init)
  uint32_t area;                           // coverage mode AREA takes this draw (an unclipped fill): no edges, no sweep; its
                                           // bounds are those of its transformed control points (skb_area.cuh)
};
static_assert(offsetof(OpGeom, color) % 8 == 0 && sizeof(OpGeom) % 8 == 0, "the fine pass loads (color, fast_solid) as one 64-bit word");

// Second half of RastePath's prologue: bounds_ = floor/ceil(path bounds); scan = bounds ∩ clip,
// floor/ceil'd; empty test; WalkEdges arguments.  surf_w/h bound the tile rectangle.
// extra_right = 1 for ops that go through the clip stage: FindSpan's `+ 1` can reach the pixel just right of
// the scan rectangle, so their tile rectangle is one pixel wider.
// unbounded = 1 for clip paths: their spans count wherever they fall (HasClip(), nested clips), so the tile
// rectangle is the whole scan rectangle, on the surface or not.
SKB_HDN void op_setup(OpGeom& g, const float clip[4], uint32_t surf_w, uint32_t surf_h, bool have_points, int extra_right = 0,
                      int unbounded = 0) {
  g.empty = 1;
  g.ntx = g.nty = 0;
  g.tx0 = g.ty0 = 0;
  if (!have_points) return;
  float l = key_float(g.bmin_x), t = key_float(g.bmin_y), r = key_float(g.bmax_x), b = key_float(g.bmax_y);
  if (!(finite_f(l) && finite_f(t) && finite_f(r) && finite_f(b))) return;  // Rect::SetBoundsCheck -> empty
  float bt = floorf(t);
  float il = l > clip[0] ? l : clip[0], ir = r < clip[2] ? r : clip[2];
  float it = t > clip[1] ? t : clip[1], ib = b < clip[3] ? b : clip[3];
  if (!(il < ir && it < ib)) { il = it = ir = ib = 0.f; }  // Rect::Intersect failure -> SetEmpty
  float sl = floorf(il), st = floorf(it), sr = ceilf(ir), sb = ceilf(ib);
  if (!(sl < sr && st < sb)) return;
  g.scan_top_f = st;
  g.scan_bottom_f = sb;
  g.scan_l = f2i_trunc(sl);
  g.scan_t = f2i_trunc(st);
  g.scan_r = f2i_trunc(sr);
  g.scan_b = f2i_trunc(sb);
  g.start_y = f2i_trunc(bt);
  g.stop_y = f2i_trunc(sb);
  g.left_clip = (fx)(f2u_wrap(sl) << 16);
  g.right_clip = (fx)(f2u_wrap(sr) << 16);
  g.empty = 0;
  // tiles: scan rectangle ∩ surface (SWSpanBrush::Brush clips spans to the bitmap, sw_span_brush.cc:80-99)
  int x0 = g.scan_l, y0 = g.scan_t, x1 = g.scan_r + extra_right, y1 = g.scan_b;
  if (!unbounded) {
    x0 = x0 < 0 ? 0 : x0;
    y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > (int)surf_w ? (int)surf_w : x1;
    y1 = y1 > (int)surf_h ? (int)surf_h : y1;
  }
  if (x0 >= x1 || y0 >= y1) return;  // rasterised but entirely off-surface: no tiles
  g.tx0 = x0 >> 4;  // floor division by SKB_TILE, also for negative coordinates
  g.ty0 = y0 >> 4;
  g.ntx = ((x1 + SKB_TILE - 1) >> 4) - g.tx0;
  g.nty = ((y1 + SKB_TILE - 1) >> 4) - g.ty0;
}

// Flatten one lowered primitive (line or quad, already transformed) into its 0..2 edges
// (SWEdgeBuilder::AddLine / ChopQuadAtYExtrema + AddQuad, sw_edge.cc:299-336).  Edges go to
// slot[0], slot[1]; the valid bit (24) tells the walker which are live, bit 25 marks quadratics
// whose cull extent (q_first_y, q_last_y) is parked in prev/next.
SKB_HDN void flatten_prim(int npts, const V2 p[3], Edge slot[2], QuadState qslot[2], int wide = 0) {
  slot[0].curve = 0;
  slot[1].curve = 0;
  if (npts == 2) {
    Edge e;
    e.curve = 0;
    e.prev = e.next = -1;
    if (set_line(e, p[0].x, p[0].y, p[1].x, p[1].y, wide)) {
      e.curve |= 1 << 24;
      slot[0] = e;
    }
  } else if (npts == 3) {
    V2 mono[5];
    int k = chop_quad_y(p, mono);
    for (int j = 0; j < k; j++) {
      Edge e;
      QuadState qs;
      e.curve = 0;
      float q[6] = {mono[2 * j].x, mono[2 * j].y, mono[2 * j + 1].x, mono[2 * j + 1].y, mono[2 * j + 2].x, mono[2 * j + 2].y};
      fx fy, ly;
      if (set_quad(e, qs, q, &fy, &ly, wide)) {
        e.curve |= (1 << 24) | (1 << 25);
        e.prev = fy;
        e.next = ly;
        slot[j] = e;
        qslot[j] = qs;
      }
    }
  }
}

// Coverage of one pixel of one path from its trapezoid rows: `direct` is the value a directly
// emitted span gives it (RealSpanBuilder, full-height band), `accum` the saturating sum of the
// rows that go through the accumulating SpanBuilder (sw_raster.cc:54-91).  In the reference these
// are two separate spans, blended one after the other (direct first).
struct PixelCover { uint8_t direct, accum; };

SKB_HDN PixelCover cover_pixel(const TrapRec* pool, uint2 row, int x) {
  PixelCover pc;
  pc.direct = 0;
  pc.accum = 0;
  uint32_t idx = row.x;
  uint32_t acc = 0;
  for (uint32_t k = 0; k < row.y; k++, idx++) {
    TrapRec r = pool[idx];
    if (r.flags & SKB_REC_LINK) {
      idx = (uint32_t)r.y;
      r = pool[idx];
    }
    uint8_t a;
    const TrapPrep pr = trap_prepare(r);
    if (!trap_prep_alpha(pr, x, &a)) continue;
    const uint32_t full = r.flags & 0xFF;
    if (full == 0xFF && !((r.flags >> 8) & 1)) {
      pc.direct = a;  // direct spans of one row never overlap (edges_too_close routes overlaps to accum)
    } else {
      acc += a;
    }
  }
  pc.accum = (uint8_t)(acc > 255 ? 255 : acc);
  return pc;
}

}  // namespace skb

#endif  // SKB_STAGES_CUH
