// skb_backend.cu — CUDA kernels of the raster pipeline and the C ABI (include/skb.h).
//
// Stages per frame (all on the surface's stream; sm_100a only):
//   1 flatten   k_op_init, k_seg_count, scan, k_flatten   curves -> lowered primitives -> edges
//   2 setup     k_op_setup, scans                          bounds, scan rectangles, work-item tables
//   3 walk      k_walk                                     active-edge sweep -> trapezoid rows (TrapRec)
//   4 coverage  k_cover                                    per (op, 16x16 tile): A8 masks + tile classes
//   5 bin       scan, k_scatter                            per-tile command lists
//   6 fine      k_fine                                     paint + SrcOver, 128-bit RGBA8 loads/stores
//   7 blur      k_blur_h, k_blur_v                         StackBlur-exact triangular filter
// HBM layout: see DESIGN.md §3.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "include/skb.h"
#include "skity_b200/csrc/skb_clip.cuh"
#include "skity_b200/csrc/skb_stages.cuh"
#include "skity_b200/csrc/skb_rowwalk.cuh"
#include "skity_b200/csrc/skb_area.cuh"

namespace skb {

static thread_local std::string g_last_error;
static void set_error(const std::string& s) { g_last_error = s; }

#define SKB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
      return SKB_ERROR_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

// ------------------------------------------------------------------ device tables
// NVTX ranges named after the reference's own trace hooks (SKITY_TRACE_EVENT, src/tracing.hpp:14-30), so that a
// timeline of this backend reads like a trace of the software backend: SWRaster_RastePath (sw_raster.cc:734) covers
// flatten / setup / walk / coverage, SWCanvas_OnClipPath (sw_canvas.cc:316) the clip stage, SWSpanBrush_Brush
// (sw_span_brush.cc:67) bin + fine, SWCanvas_HandleFilter the blur passes.  Free when no tool is attached.
struct NvtxStage {
  bool open = false;
  void next(const char* name) {
    if (open) nvtxRangePop();
    nvtxRangePushA(name);
    open = true;
  }
  ~NvtxStage() {
    if (open) nvtxRangePop();
  }
};

struct SurfDesc {
  uint8_t* px;          // premultiplied RGBA8, rows `pitch` bytes apart, padded to whole tiles
  uint32_t w, h;        // logical size
  uint32_t pitch;       // bytes
  uint32_t tiles_x, tiles_y;
  uint32_t tile_base;   // first global tile index
  uint32_t row0, row1;  // rows this device renders (band), whole surface for temporaries
  uint32_t level;       // pass in which the surface is composited: after every surface its draws sample or are blurred from
  uint32_t pad;
  uint8_t* px_out;      // where the fine pass stores the finished pixels: px, or the same canvas on ANOTHER GPU (peer
                        // memory over NVLink) when the bands of one canvas are rendered by several devices
};

#define SKB_CMD_SOLID 0x80000000u
#define SKB_CMD_ZERO 0x40000000u   // the item has a zero-coverage map (CoverArgs::zmask)
#define SKB_CMD_ITEM_MASK 0x3FFFFFFFu

__device__ __forceinline__ uint32_t find_interval(const uint32_t* off, uint32_t n, uint32_t v) {
  // largest i in [0, n) with off[i] <= v   (off is non-decreasing, off[0] <= v)
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------- scan
// Exclusive prefix sum of n uint32 in place, 2048 elements per CTA, recursive over block sums.
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block(uint32_t* data, uint32_t n, uint32_t* block_sums) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    v[i] = idx < n ? data[idx] : 0u;
    sum += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;  // exclusive
    if (lane == SCAN_THREADS / 32 - 1 && block_sums) block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + (incl - sum);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    if (idx < n) data[idx] = run;
    run += v[i];
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* data, uint32_t n, const uint32_t* block_offsets) {
  const uint32_t add = block_offsets[blockIdx.x];
  const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    uint32_t idx = base + i;
    if (idx < n) data[idx] += add;
  }
}

// ------------------------------------------------------------------ stage 1: flatten
struct FrameTables {
  const skb_dl_op* ops;
  const skb_dl_path* paths;
  const skb_dl_seg* segs;
  const skb_dl_paint* paints;
  const float* stops;
  uint32_t n_ops, n_segs;
  uint32_t wide;  // wide-coordinate mode (skb_surface_set_coord_mode): 24.8 -> 16.16 without the reference's int32 wrap
};

// Also culls, before anything is flattened, the draws that cannot reach the rows this device renders (band split of
// one canvas over several GPUs, skb_surface_set_band): the control points of a path bound its lowered quads up to half
// of their extent (a quad's control point from Cubic::ToQuads is (3(c1'+c2') - (p0'+p3'))/4 with all four inside the
// control polygon's bounds), so a path whose control points' y range, widened by half its height plus 2 px, misses the
// band gets no primitives at all (`culled`: k_seg_count counts 0 for its segments).  Clip paths are never culled.
__global__ void k_op_init(FrameTables t, OpGeom* geom, uint32_t* seg_op, const SurfDesc* surfs, int area_mode) {
  uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= t.n_ops) return;
  OpGeom g;
  memset(&g, 0, sizeof(g));
  g.bmin_x = g.bmin_y = INT_MAX;
  g.bmax_x = g.bmax_y = INT_MIN;
  g.empty = 1;
  const skb_dl_op o = t.ops[op];
  if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
    const skb_dl_path p = t.paths[o.path];
    const SurfDesc sd = surfs[o.surface];
    const bool banded = o.kind == SKB_OP_FILL && (sd.row0 > 0 || sd.row1 < sd.h);
    // coverage mode AREA (skb_surface_set_coverage_mode): unclipped fills are binned as lines and never swept; their
    // bounds are those of the transformed path's points (Path::GetBounds in CoverageAAPathTiler::Tile)
    const bool area = area_mode != 0 && o.kind == SKB_OP_FILL && o.clip_in == 0;
    g.area = area ? 1u : 0u;
    float ymin = 3.0e38f, ymax = -3.0e38f;
    bool all_finite = true;
    for (uint32_t i = 0; i < p.n_segs; i++) {
      uint32_t s = p.seg_off + i;
      seg_op[s] = op;
      const uint32_t type = t.segs[s].type_flags & SKB_SEG_TYPE_MASK;
      if (area && type != SKB_SEG_POINT) {
        const skb_dl_seg& sg = t.segs[s];
        const int last = (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) ? 1 : type == SKB_SEG_CUBIC ? 3 : 2;
        for (int k = 0; k <= last; k++) {
          const V2 q = xform(o.ctm, v2(sg.p[2 * k], sg.p[2 * k + 1]));
          const int32_t kx = float_key(q.x), ky = float_key(q.y);
          g.bmin_x = min(g.bmin_x, kx); g.bmax_x = max(g.bmax_x, kx);
          g.bmin_y = min(g.bmin_y, ky); g.bmax_y = max(g.bmax_y, ky);
        }
      }
      if (banded) {
        const skb_dl_seg& sg = t.segs[s];
        int first = 1, last = 0;  // control points p[2*first .. 2*last+1]; p[0..1] repeats the start point
        if (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) last = 1;
        else if (type == SKB_SEG_QUAD || type == SKB_SEG_CONIC) last = 2;
        else if (type == SKB_SEG_CUBIC) last = 3;
        const V2 st = xform(o.ctm, seg_start_point(t.segs, s));
        all_finite &= finite_f(st.y);
        ymin = fminf(ymin, st.y); ymax = fmaxf(ymax, st.y);
        for (int k = first; k <= last; k++) {
          const V2 q = xform(o.ctm, v2(sg.p[2 * k], sg.p[2 * k + 1]));
          all_finite &= finite_f(q.y);
          ymin = fminf(ymin, q.y); ymax = fmaxf(ymax, q.y);
        }
      }
      if (type == SKB_SEG_POINT) {
        V2 q = xform(o.ctm, seg_start_point(t.segs, s));
        int32_t kx = float_key(q.x), ky = float_key(q.y);
        g.bmin_x = min(g.bmin_x, kx); g.bmax_x = max(g.bmax_x, kx);
        g.bmin_y = min(g.bmin_y, ky); g.bmax_y = max(g.bmax_y, ky);
      }
    }
    if (banded && all_finite && p.n_segs > 0) {
      const float pad = 0.5f * (ymax - ymin) + 2.0f;
      if (ymax + pad < (float)sd.row0 || ymin - pad > (float)sd.row1) {
        g.culled = 1;
        g.bmin_x = g.bmin_y = INT_MAX;   // no bounds: k_op_setup leaves the op empty
        g.bmax_x = g.bmax_y = INT_MIN;
      }
    }
  }
  geom[op] = g;
}

__global__ void k_seg_count(FrameTables t, uint32_t* prim_cnt, const uint32_t* seg_op, const OpGeom* geom) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= t.n_segs) return;
  const OpGeom& g = geom[seg_op[s]];
  prim_cnt[s] = (g.culled || g.area) ? 0u : (uint32_t)seg_prim_count(t.segs[s]);
}

// One thread per segment, looping over the primitives it lowers to (a line / quad gives one, a conic two, a cubic a
// handful — their first index is the segment's entry in the scanned counts): evaluate and transform the control points,
// fold them into the path bounds, emit each primitive's 0..2 edges.  (One thread per primitive had to find its segment by
// a binary search over the scanned counts — 23 dependent loads at 6M segments; C4a flatten stage 2.1 -> see DESIGN.)
__global__ void k_flatten(FrameTables t, const uint32_t* prim_off, uint32_t n_prims, const uint32_t* seg_op, OpGeom* geom,
                          Edge* edges, QuadState* quads, uint32_t* chord_cap, uint32_t* slot_op) {
  const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= t.n_segs) return;
  const uint32_t p0 = prim_off[seg];
  const int n = (int)(prim_off[seg + 1] - p0);
  if (n <= 0) return;
  const uint32_t op = seg_op[seg];
  const float* ctm = t.ops[op].ctm;
  OpGeom* g = &geom[op];
  // Each path owns one contiguous region of 64 bytes per slot: its Edge array followed by its
  // QuadState array, so the sweep's working set per path stays within a few cache lines.
  const skb_dl_path pa = t.paths[t.ops[op].path];
  const uint32_t first_prim = prim_off[pa.seg_off];
  const uint32_t n_slots = 2 + 2 * (prim_off[pa.seg_off + pa.n_segs] - first_prim);
  const size_t slot_base = (size_t)2 * first_prim + (size_t)2 * op;
  uint8_t* region = reinterpret_cast<uint8_t*>(edges) + slot_base * (sizeof(Edge) + sizeof(QuadState));
  Edge* E = reinterpret_cast<Edge*>(region);
  QuadState* Q = reinterpret_cast<QuadState*>(region + (size_t)n_slots * sizeof(Edge));
  int32_t bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
  for (int k = 0; k < n; k++) {
    const uint32_t prim = p0 + (uint32_t)k;
    V2 p[3];
    int np = seg_prim(t.segs, seg, k, n, ctm, p);
    for (int j = 0; j < np; j++) {
      int32_t kx = float_key(p[j].x), ky = float_key(p[j].y);
      bx0 = min(bx0, kx); bx1 = max(bx1, kx);
      by0 = min(by0, ky); by1 = max(by1, ky);
    }
    Edge slot[2];
    QuadState qslot[2];
    flatten_prim(np, p, slot, qslot, (int)t.wide);
    const uint32_t at = 2 + 2 * (prim - first_prim);
    E[at] = slot[0];
    E[at + 1] = slot[1];
    if ((slot[0].curve >> 25) & 1) Q[at] = qslot[0];
    if ((slot[1].curve >> 25) & 1) Q[at + 1] = qslot[1];
    if (chord_cap) {
      // row-parallel walk (skb_rowwalk.cuh): an upper bound of the chords each edge can give — a line is one, a
      // quadratic its first chord plus one per remaining subdivision — and the op that owns the slot
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const bool valid = (slot[j].curve >> 24) & 1, quad = (slot[j].curve >> 25) & 1;
        chord_cap[slot_base + at + j] = valid ? (quad ? 1u + (uint32_t)edge_count(slot[j]) : 1u) : 0u;
        slot_op[slot_base + at + j] = op;
      }
    }
  }
  if (bx0 <= bx1) {
    atomicMin(&g->bmin_x, bx0); atomicMax(&g->bmax_x, bx1);
    atomicMin(&g->bmin_y, by0); atomicMax(&g->bmax_y, by1);
  }
}

// ------------------------------------------------------------------- stage 2: setup
__global__ void k_op_setup(FrameTables t, const uint32_t* prim_off, OpGeom* geom, const SurfDesc* surfs, uint32_t* row_cnt,
                           uint32_t* item_cnt, uint32_t* too_big, RwOp* rwops, uint32_t* wrow_cnt) {
  uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= t.n_ops) return;
  const skb_dl_op o = t.ops[op];
  uint32_t rows = 0, items = 0;
  if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
    OpGeom g = geom[op];
    const skb_dl_path p = t.paths[o.path];
    uint32_t first_prim = prim_off[p.seg_off];
    uint32_t n_prims = prim_off[p.seg_off + p.n_segs] - first_prim;
    g.first_prim = first_prim;
    g.slot_base = 2 * first_prim + 2 * op;
    g.n_slots = 2 + 2 * n_prims;
    const SurfDesc sd = surfs[o.surface];
    const bool is_clip = o.kind == SKB_OP_CLIP;
    op_setup(g, o.clip_bounds, sd.w, sd.h, p.n_segs > 0, (is_clip || o.clip_in != 0) ? 1 : 0, is_clip ? 1 : 0);
    if (is_clip && !g.empty && (int64_t)g.ntx * g.nty > (1 << 20)) {  // a clip path reaching absurdly far off the surface (> 2^28 pixels)
      *too_big = 1;
      g.ntx = g.nty = 0;
    }
    // a clip state is needed whole on every device: no band restriction for clip paths
    if (!is_clip && !g.empty && g.ntx > 0) {
      // keep only the tile rows this device renders (band split); the rows outside are another GPU's
      int ty0 = max(g.ty0, (int)(sd.row0 / SKB_TILE));
      int ty1 = min(g.ty0 + g.nty, (int)((sd.row1 + SKB_TILE - 1) / SKB_TILE));
      if (ty1 <= ty0) {
        g.ntx = g.nty = 0;
      } else {
        g.ty0 = ty0;
        g.nty = ty1 - ty0;
      }
    }
    if (!g.empty && g.ntx > 0) {
      rows = (uint32_t)(g.nty * SKB_TILE);
      items = is_clip ? 0u : (uint32_t)(g.ntx * g.nty);  // clip paths fill clip tables, not tile masks
    }
    if (o.kind == SKB_OP_FILL) {
      const skb_dl_paint pt = t.paints[o.paint];
      // kept in the order the fine pass blends in (swap_rb)
      g.color = pt.type == SKB_PAINT_SOLID ? swap_rb(color4f_to_pm_word(pt.color[0], pt.color[1], pt.color[2], pt.color[3])) : 0u;
      g.fast_solid = (pt.type == SKB_PAINT_SOLID && pt.blend == 0 && SKB_PAINT_CF_OFFSET(pt) == 0) ? 1u : 0u;
    }
    geom[op] = g;
  }
  row_cnt[op] = rows;
  item_cnt[op] = items;
  {
    // 64-bit totals beside the 32-bit prefix scans (too_big + 10: rows, + 12: items): a frame whose scan rows or
    // (op, tile) items do not fit 32 bits is refused by run_frame instead of wrapping into undersized buffers
    const unsigned m = __activemask();
    const uint32_t wr = __reduce_add_sync(m, rows), wi = __reduce_add_sync(m, items);
    if ((int)(threadIdx.x & 31) == __ffs((int)m) - 1) {
      if (wr) atomicAdd(reinterpret_cast<unsigned long long*>(too_big + 10), (unsigned long long)wr);
      if (wi) atomicAdd(reinterpret_cast<unsigned long long*>(too_big + 12), (unsigned long long)wi);
    }
  }
  if (rwops) {
    // row-parallel walk: the rows the sweep covers, from WalkEdges' start_y (the top of the path's bounds — above the
    // scan rectangle when the path is clipped at the top) to where the records stop being read
    RwOp r;
    r.wrow_base = 0;
    r.n_wrows = 0;
    r.origin_row = 0;
    r.y0q = INT_MAX;
    r.fail = 0;
    r.rec_base = r.n_recs = r.pad = 0;
    uint32_t wr = 0;
    if (rows) {
      const OpGeom g = geom[op];
      const int stop = min(g.stop_y, (g.ty0 + g.nty) * SKB_TILE);
      const int64_t n = (int64_t)stop - (int64_t)g.start_y;
      r.origin_row = g.start_y;
      if (n <= 0 || n * 4 > SKB_RW_MAXQ) {
        r.fail = SKB_RWFAIL_HARD;   // taller than the chord table addresses: swept sequentially
      } else {
        r.n_wrows = (int32_t)n;
        wr = ((uint32_t)n + 1u + 15u) & ~15u;   // + the table's end entry, in groups of 16 rows
      }
    }
    rwops[op] = r;
    wrow_cnt[op] = wr;
  }
}

// -------------------------------------------------------------------- stage 3: walk
// One thread sweeps one path.  The sweep is a long chain of dependent, branchy integer work, so a
// warp of 32 different paths serialises heavily.  When a frame has far fewer paths than the GPU has
// thread slots, the paths are therefore SPREAD: only every (32/lanes)-th lane of a warp carries a
// path, which cuts the divergence per warp and multiplies the number of warps the schedulers can
// interleave.  k_walk_list first compacts the non-empty ops into a list.
// Also writes the owner of every tile row (trow_op): the coverage and clip stages start from a tile row or a
// pixel row and would otherwise find its op by a binary search over row_base — 20 dependent loads at 1M ops.
// When the paths are few, the frame's sweep lasts as long as its longest path, and a path sharing its warp with
// others advances only as fast as the warp gets through every lane's branch of every iteration.  The longest paths (by
// rows x edge slots, a histogram in half octaves: k_walk_hist, k_walk_pick) are therefore put at the front of the list
// and get a warp each; the others share warps as before.
__device__ __forceinline__ uint32_t walk_work_bucket(const OpGeom& g) {
  const uint32_t w = (uint32_t)g.nty * g.n_slots;
  if (w == 0) return 0;
  const int e = 31 - __clz((int)w);
  const uint32_t frac = e >= 1 ? (w >> (e - 1)) & 1u : 0u;
  return min(63u, 2u * (uint32_t)e + frac);
}
__global__ void k_walk_hist(FrameTables t, const OpGeom* geom, uint32_t* hist) {
  uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= t.n_ops) return;
  const OpGeom g = geom[op];
  if (g.empty || g.ntx == 0 || g.nty == 0 || g.area) return;
  atomicAdd(&hist[walk_work_bucket(g)], 1u);
}
// pick[0] = the bucket above which a path is "long", pick[1] = how many those are (at most max_long)
// A path is long only when it is well above the frame's median (more than two half octaves: about three times the
// median's work) — in a frame of similar paths (C4b: 64 canvases of 1 000 random paths) nobody gains from a warp of
// its own, and the warps spent on it are missing elsewhere.
__global__ void k_walk_pick(const uint32_t* hist, uint32_t max_long, uint32_t* pick) {
  uint32_t total = 0;
  for (int b = 0; b < 64; b++) total += hist[b];
  int median = 0;
  for (uint32_t below = 0; median < 63; median++) {
    below += hist[median];
    if (2 * below >= total) break;
  }
  uint32_t cum = 0;
  int b = 63;
  for (; b > median + 2; b--) {
    if (cum + hist[b] > max_long) break;
    cum += hist[b];
  }
  pick[0] = (uint32_t)b;
  pick[1] = cum;
}

__global__ void k_walk_list(FrameTables t, const OpGeom* geom, uint32_t* count, uint32_t* list, const uint32_t* row_base,
                            uint32_t* trow_op, RwOp* rwops, const uint32_t* wrow_base, uint32_t* wgrp_op,
                            const uint32_t* pick, uint32_t* cursors) {
  uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= t.n_ops) return;
  const OpGeom g = geom[op];
  if (g.empty || g.ntx == 0 || g.nty == 0) return;
  {
    uint32_t* o = trow_op + row_base[op] / SKB_TILE;
    for (int r = 0; r < g.nty; r++) o[r] = op;
  }
  if (g.area) return;   // coverage mode AREA: binned as lines (k_area_*), not swept
  if (rwops) {   // row-parallel walk: where the op's walk rows are, and who owns each group of 16 of them
    const uint32_t wb = wrow_base[op], we = wrow_base[op + 1];
    rwops[op].wrow_base = wb;
    for (uint32_t gq = wb >> 4; gq < (we >> 4); gq++) wgrp_op[gq] = op;
    return;   // the sequential sweep's list is made later, of the paths the row-parallel form gives up on
  }
  if (pick) {   // long paths to the front of the list (pick[1] of them), the others behind
    const bool is_long = (int)walk_work_bucket(g) > (int)pick[0];
    const uint32_t pos = is_long ? atomicAdd(&cursors[0], 1u) : pick[1] + atomicAdd(&cursors[1], 1u);
    list[pos] = op;
    atomicAdd(count, 1u);
    return;
  }
  // warp-aggregated append
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  list[base + __popc(m & ((1u << lane) - 1))] = op;
}

struct WalkArgs {
  FrameTables t;
  const OpGeom* geom;
  const uint32_t* row_base;
  Edge* edges;
  uint8_t* edges2;   // same layout and size as `edges`: the sweep's compact copy in sweep order (walk_prologue), or null
  QuadState* quads;
  int32_t* ord;
  TrapRec* pool;
  uint32_t* pool_next;
  uint32_t pool_cap;
  uint32_t* overflow;
  uint2* rows;
  const uint32_t* list;   // ops of this class
  const uint32_t* count;  // how many
  const uint32_t* pick;   // [1] = how many paths at the front of the list get a warp each (null: none)
};

__device__ __forceinline__ void walk_setup_sink(const WalkArgs& a, uint32_t op, const OpGeom& g, RecSink& sink) {
  sink.pool = a.pool;
  sink.pool_next = a.pool_next;
  sink.pool_cap = a.pool_cap;
  sink.overflow = a.overflow;
  sink.rows = a.rows + a.row_base[op];
  sink.row0 = g.ty0 * SKB_TILE;
  sink.n_rows = g.nty * SKB_TILE;
  sink_init(sink);
}

#ifndef WALK_BLOCK
#define WALK_BLOCK 64
#endif
#ifdef WALK_MINB  // minimum resident blocks per SM (caps the registers); unset = the compiler's own choice
#define WALK_BOUNDS __launch_bounds__(WALK_BLOCK, WALK_MINB)
#else
#define WALK_BOUNDS __launch_bounds__(WALK_BLOCK)
#endif
__global__ void WALK_BOUNDS k_walk(WalkArgs a, int lane_stride) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_list = *a.count;
  // the first n_long paths of the list have a warp each (its lane 0); the threads behind those warps take the rest
  const uint32_t n_long = a.pick ? min(a.pick[1], n_list) : 0u;
  uint32_t first, step;
  if ((tid >> 5) < n_long) {
    if (tid & 31u) return;
    first = tid >> 5;
    step = 0xFFFFFFFFu - first;   // one path only
  } else {
    tid -= n_long * 32u;
    if (tid % (uint32_t)lane_stride) return;
    // a grid smaller than the list (SKB_WALK_RESIDENT: a cap on the paths in flight, i.e. on the sweep's cache
    // footprint) takes the paths in strides of the grid
    first = n_long + tid / (uint32_t)lane_stride;
    step = max(1u, (gridDim.x * blockDim.x - n_long * 32u) / (uint32_t)lane_stride);
  }
  for (uint32_t i = first; i < n_list; i += step) {
  const uint32_t op = a.list[i];
  const OpGeom g = a.geom[op];
  if (g.empty || g.ntx == 0 || g.nty == 0) continue;
  RecSink sink;
  walk_setup_sink(a, op, g, sink);
  // rows below the last tile row are never read: stop the sweep there (rows do not depend on later ones)
  int stop_y = min(g.stop_y, sink.row0 + sink.n_rows);
  uint8_t* region = reinterpret_cast<uint8_t*>(a.edges) + (size_t)g.slot_base * (sizeof(Edge) + sizeof(QuadState));
  Edge* E = reinterpret_cast<Edge*>(region);
  QuadState* Q = reinterpret_cast<QuadState*>(region + (size_t)g.n_slots * sizeof(Edge));
  Edge* E2 = nullptr;
  QuadState* Q2 = nullptr;
  if (a.edges2) {
    uint8_t* region2 = a.edges2 + (size_t)g.slot_base * (sizeof(Edge) + sizeof(QuadState));
    E2 = reinterpret_cast<Edge*>(region2);   // the quadratic states follow the compact edges (walk_prologue)
  }
  walk_path(E, Q, (int)g.n_slots, a.ord + g.slot_base, g.scan_top_f, g.scan_bottom_f, g.start_y, stop_y, g.left_clip,
            g.right_clip, (int)a.t.ops[op].fill_type, sink, (int)a.t.wide, E2, Q2);
  }
}

// ------------------------------------------------- stage 3 (row-parallel form): skb_rowwalk.cuh
// The sweep as short kernels whose threads own a slot, a path or a (path, pixel row):
//   k_rw_events   slot      chords' y (forward differencing of y alone), events of every walk row, first y of the path
//   k_rw_bands0   path      band tables from the events alone
//   k_rw_chain    slot      chords with their x, walked through the bands of the tables
//   k_rw_rows<0>  (path,row) the row swept under both hypotheses about its first band: forcing bits, masks, record counts
//   k_rw_resolve  path      composes the rows' f_in -> f_out maps; final band tables, record offsets
//   k_rw_chain    slot      again, with the final tables
//   k_rw_rows<1>  (path,row) the row swept with the exact x: emits the records, re-derives every table entry; a
//                            mismatch flags the path (retried once from its new tables with the sort ranks; what still
//                            fails is swept by the sequential k_walk)
struct RwArgs {
  FrameTables t;
  const OpGeom* geom;
  const uint32_t* row_base;
  uint8_t* edges;
  RwOp* ops;
  const uint32_t* wgrp_op;   // owner of every group of 16 walk rows
  uint32_t n_wgrps;
  const uint32_t* slot_op;
  uint32_t n_slots_total;
  const uint32_t* chord_base;   // scanned chord capacities, n_slots_total + 1
  SlotInfo* slots;
  Chord* chords;
  uint32_t* ev_words;
  RowBand* tab;
  uint32_t* res;
  uint32_t* rec_off;
  uint32_t* rec_cnt;            // per op, scanned into the ops' first records
  uint16_t* rank;
  int32_t* ord;
  TrapRec* pool;
  uint32_t pool_cap;
  uint32_t* pool_next;
  uint32_t* overflow;
  uint2* rows;
  uint32_t* n_failed;           // statistics: [0] retried, [1] swept sequentially
};

__device__ __forceinline__ bool rw_op_active(uint32_t fail, int phase) {
  return phase == 0 ? fail == 0 : (fail & (SKB_RWFAIL_RETRY | SKB_RWFAIL_HARD)) == SKB_RWFAIL_RETRY;
}
__device__ __forceinline__ const Edge* rw_op_edges(const RwArgs& a, const OpGeom& g) {
  return reinterpret_cast<const Edge*>(a.edges + (size_t)g.slot_base * (sizeof(Edge) + sizeof(QuadState)));
}

__global__ void __launch_bounds__(128) k_rw_events(RwArgs a) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_slots_total) return;
  const uint32_t cbase = a.chord_base[s];
  const uint32_t cap = a.chord_base[s + 1] - cbase;
  SlotInfo si;
  si.chord_base = cbase;
  si.n_chords = 0;
  si.q_first = si.q_last = 0;
  if (cap) {
    const uint32_t op = a.slot_op[s];
    const RwOp ro = a.ops[op];
    if (ro.fail == 0 && ro.n_wrows > 0) {
      const OpGeom g = a.geom[op];
      const Edge* E = rw_op_edges(a, g);
      const QuadState* Q = reinterpret_cast<const QuadState*>(E + g.n_slots);
      const uint32_t local = s - g.slot_base;
      const Edge e = E[local];
      const bool quad = (e.curve >> 25) & 1;
      const fx y0 = quad ? e.prev : e.upper_y, y1 = quad ? e.next : e.lower_y;
      if (!can_be_ignored(g.scan_top_f, g.scan_bottom_f, y0, y1, (int)a.t.wide)) {
        QuadState q;
        if (quad) q = Q[local];
        const int n = rw_chain(e, q, ro.origin_row, ro.n_wrows * 4, nullptr, a.chords + cbase, (int)cap, a.ev_words + (ro.wrow_base >> 2));
        if (n <= 0) {
          atomicOr(&a.ops[op].fail, SKB_RWFAIL_HARD);
        } else {
          si.n_chords = (uint32_t)n;
          si.q_first = chord_uq(a.chords[cbase].yy);
          si.q_last = chord_lq(a.chords[cbase + n - 1].yy);
          atomicMin(&a.ops[op].y0q, si.q_first);
        }
      }
    }
  }
  a.slots[s] = si;
}

// phase 1 (retry): also the sort ranks, and the op's fail bits move from "soft" to "being retried"
__global__ void __launch_bounds__(128) k_rw_bands0(RwArgs a, int phase) {
  const uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= a.t.n_ops) return;
  RwOp ro = a.ops[op];
  if (phase == 1) {
    if ((ro.fail & SKB_RWFAIL_HARD) || !(ro.fail & SKB_RWFAIL_SOFT)) return;
    a.ops[op].fail = SKB_RWFAIL_RETRY;
    atomicAdd(&a.n_failed[0], 1u);
    const OpGeom g = a.geom[op];
    {  // row entries the first attempt's final pass wrote before one of its rows failed
      uint2* r = a.rows + a.row_base[op];
      for (int i = 0; i < g.nty * SKB_TILE; i++) r[i] = make_uint2(0u, 0u);
    }
    rw_sort_ranks(rw_op_edges(a, g), (int)g.n_slots, a.ord + g.slot_base, g.scan_top_f, g.scan_bottom_f, (int)a.t.wide, a.rank + g.slot_base);
  } else if (ro.fail) {
    return;
  }
  if (ro.n_wrows <= 0 || ro.y0q == INT_MAX) return;
  rw_bands_from_events(reinterpret_cast<const uint8_t*>(a.ev_words) + ro.wrow_base, ro.n_wrows, ro.y0q, a.tab + ro.wrow_base);
}

__global__ void __launch_bounds__(128) k_rw_chain(RwArgs a, int phase) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_slots_total) return;
  const SlotInfo si = a.slots[s];
  if (si.n_chords == 0) return;
  const uint32_t op = a.slot_op[s];
  const RwOp ro = a.ops[op];
  if (!rw_op_active(ro.fail, phase)) return;
  const OpGeom g = a.geom[op];
  const Edge* E = rw_op_edges(a, g);
  const QuadState* Q = reinterpret_cast<const QuadState*>(E + g.n_slots);
  const uint32_t local = s - g.slot_base;
  const Edge e = E[local];
  if (!((e.curve >> 25) & 1)) {   // a line: its one chord is the edge itself
    Chord c = a.chords[si.chord_base];
    c.x = e.x; c.dx = e.dx; c.dy = e.dy;
    a.chords[si.chord_base] = c;
    return;
  }
  const QuadState q = Q[local];
  const int cap = (int)(a.chord_base[s + 1] - si.chord_base);
  const int n = rw_chain(e, q, ro.origin_row, ro.n_wrows * 4, a.tab + ro.wrow_base, a.chords + si.chord_base, cap, nullptr);
  if (n != (int)si.n_chords) atomicOr(&a.ops[op].fail, SKB_RWFAIL_HARD);
}

__device__ __forceinline__ void rw_fill_row_in(const RwArgs& a, const RwOp& ro, const OpGeom& g, uint32_t op, int row, int phase, RwRowIn& in) {
  in.rank = phase == 1 ? a.rank + g.slot_base : nullptr;
  in.slots = a.slots + g.slot_base;
  in.n_slots = (int)g.n_slots;
  in.chords = a.chords;
  in.tab = a.tab + ro.wrow_base;
  in.ev = reinterpret_cast<const uint8_t*>(a.ev_words) + ro.wrow_base;
  in.row = row;
  in.y0q = ro.y0q;
  in.stop_q = ro.n_wrows * 4;
  in.origin_fx = i_to_fx(ro.origin_row);
  in.left_clip = g.left_clip;
  in.right_clip = g.right_clip;
  in.even_odd = (int)a.t.ops[op].fill_type;
  in.exact = false;
}

#ifndef RW_ROWS_BLOCK
#define RW_ROWS_BLOCK 64
#endif
template <int FINAL>
__global__ void __launch_bounds__(RW_ROWS_BLOCK) k_rw_rows(RwArgs a, int phase) {
  const uint32_t wr = blockIdx.x * blockDim.x + threadIdx.x;   // global walk row
  // chunked records (retries, sequential sweeps) follow the linear ones, on a chunk boundary
  if (FINAL && phase == 0 && wr == 0) *a.pool_next = (a.rec_cnt[a.t.n_ops] + SKB_CHUNK - 1) & ~(SKB_CHUNK - 1);
  if ((wr >> 4) >= a.n_wgrps) return;
  const uint32_t op = a.wgrp_op[wr >> 4];
  const RwOp ro = a.ops[op];
  if (!rw_op_active(ro.fail, phase)) return;
  const int row = (int)(wr - ro.wrow_base);
  if (row >= ro.n_wrows || ro.y0q == INT_MAX) return;
  const OpGeom g = a.geom[op];
  RwRowIn in;
  rw_fill_row_in(a, ro, g, op, row, phase, in);
  // rows above the first tile row of the op's scan rectangle are swept (they shape the tables) but emit nothing
  const int rel = ro.origin_row + row - g.ty0 * SKB_TILE;
  const bool emits = rel >= 0 && rel < g.nty * SKB_TILE;
  if (!FINAL) {
    RwRowOut o0, o1;
    rw_row(in, false, nullptr, 0, o0);
    rw_row(in, true, nullptr, 0, o1);
    if (o0.n_recs > 255) o0.fail = SKB_RWF_EMIT;
    if (o1.n_recs > 255) o1.fail = SKB_RWF_EMIT;
    a.res[wr] = rw_pack(o0.f_ins, o0.f_surv, o1.f_surv, o0.mask, o1.mask, o0.fail != 0, o1.fail != 0, emits ? o0.n_recs & 255 : 0,
                        emits ? o1.n_recs & 255 : 0);
  } else {
    const RowBand tb = in.tab[row];
    in.exact = true;
    const uint32_t first = ro.rec_base + a.rec_off[wr];
    const int n_alloc = (int)((row + 1 < ro.n_wrows ? a.rec_off[wr + 1] : ro.n_recs) - a.rec_off[wr]);
    RwRowOut o;
    const bool room = (uint64_t)first + (uint64_t)n_alloc <= (uint64_t)a.pool_cap;
    if (!room) *a.overflow = 1;
    rw_row(in, (tb.fl & SKB_RB_FIN) != 0, (emits && room) ? a.pool + first : nullptr, n_alloc, o);
    const bool in_rows = row * 4 + 4 > ro.y0q;
    bool bad = o.fail != 0;
    if (in_rows && (o.mask != tb.mask || o.f_surv != ((tb.fl & SKB_RB_FSURV) != 0) || o.f_ins != ((tb.fl & SKB_RB_FINS) != 0))) bad = true;
    if (emits && o.n_recs != n_alloc) bad = true;
    if (bad) {
      atomicOr(&a.ops[op].fail, phase == 0 ? SKB_RWFAIL_SOFT : SKB_RWFAIL_HARD);
    } else if (emits && n_alloc > 0 && room) {
      a.rows[a.row_base[op] + (uint32_t)rel] = make_uint2(first, (uint32_t)n_alloc | SKB_ROW_LINEAR);
    }
  }
}

__global__ void __launch_bounds__(128) k_rw_resolve(RwArgs a, int phase) {
  const uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= a.t.n_ops) return;
  if (phase == 0) a.rec_cnt[op] = 0;
  const RwOp ro = a.ops[op];
  if (!rw_op_active(ro.fail, phase) || ro.n_wrows <= 0 || ro.y0q == INT_MAX) return;
  const uint32_t total = rw_bands_resolve(a.res + ro.wrow_base, ro.n_wrows, ro.y0q, a.tab + ro.wrow_base, a.rec_off + ro.wrow_base);
  if (total == 0xFFFFFFFFu) {
    a.ops[op].fail = phase == 0 ? SKB_RWFAIL_SOFT : SKB_RWFAIL_HARD;
    return;
  }
  a.ops[op].n_recs = total;
  if (phase == 0) {
    a.rec_cnt[op] = total;
  } else {
    const uint32_t base = atomicAdd(a.pool_next, (total + SKB_CHUNK - 1) & ~(SKB_CHUNK - 1));   // keeps the sequential sweep's chunks aligned
    if ((uint64_t)base + total > (uint64_t)a.pool_cap) *a.overflow = 1;
    a.ops[op].rec_base = base;
  }
}

// after the scan of rec_cnt: the ops' first records
__global__ void __launch_bounds__(128) k_rw_rec_base(RwArgs a) {
  const uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= a.t.n_ops) return;
  a.ops[op].rec_base = a.rec_cnt[op];
}

// The paths the row-parallel form gave up on: their row entries are cleared (the final pass may have written some
// before another row of the path failed) and they are listed for the sequential sweep.
__global__ void __launch_bounds__(128) k_rw_fallback_list(RwArgs a, uint32_t* count, uint32_t* list) {
  const uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= a.t.n_ops) return;
  const uint32_t fail = a.ops[op].fail;
  if (!(fail & (SKB_RWFAIL_HARD | SKB_RWFAIL_SOFT))) return;
  const OpGeom g = a.geom[op];
  if (g.empty || g.ntx == 0 || g.nty == 0) return;
  uint2* r = a.rows + a.row_base[op];
  for (int i = 0; i < g.nty * SKB_TILE; i++) r[i] = make_uint2(0u, 0u);
  atomicAdd(&a.n_failed[1], 1u);
  list[atomicAdd(count, 1u)] = op;
}

// ---------------------------------------------------------------- stage 4: coverage
// One warp per (op, tile).  Lane L owns the 8 horizontally adjacent pixels (L&1)*8.. of tile row
// L>>1, so a tile's A8 mask is written as 32 x 8 contiguous bytes = one coalesced 256-byte store.
struct CoverArgs {
  const OpGeom* geom;
  const uint32_t* item_base;  // n_ops + 1
  const uint32_t* row_base;
  const uint32_t* trow_op;    // owner of every tile row (n_trows)
  uint32_t n_ops, n_items, n_trows;
  const skb_dl_op* ops;
  const SurfDesc* surfs;
  const TrapRec* pool;
  const uint2* rows;
  uint8_t* mask0;  // n_items * 256: plane blended first (direct spans, or the only plane)
  uint8_t* mask1;  // n_items * 256: accumulated spans where a pixel of the tile has both
  uint8_t* mask[SKB_CLIP_PLANES];  // all planes (0,1 as above; 2.. only used by clipped draws)
  uint16_t* item_flags;
  uint32_t* tile_cnt;
  const skb_dl_paint* paints;
  uint8_t* zmask;  // n_items * 256, only with blend modes that act on zero-coverage pixels: 1 = touched by a direct span
  uint8_t* zplane[SKB_CLIP_PLANES];  // the same per coverage plane, for clipped draws ([0] = zmask)
};
// item flags (u16): bit k = coverage plane k present (k < SKB_CLIP_PLANES = 8), bit 8 = plane 0 is solid 255,
// bit 9 = zmask holds this item's zero-coverage map
#define SKB_ITEM_PLANE0 1u
#define SKB_ITEM_PLANE1 2u
#define SKB_ITEM_SOLID 256u
#define SKB_ITEM_ZERO 512u
#define SKB_ITEM_PLANE_MASK 255u

#ifndef COVER_WARPS
#define COVER_WARPS 1
#endif
#ifndef COVER_CW
#define COVER_CW 128                 // pixels per chunk of a tile row (8 tiles) held in shared memory
#endif
#ifndef COVER_ND
#define COVER_ND 64                  // trapezoid records summarised per pass
#endif
#ifndef COVER_UNIT
#define COVER_UNIT 2                 // edge pixels evaluated per work unit
#endif
#ifndef COVER_INLINE_PX
#define COVER_INLINE_PX 8            // records with at most this many edge-zone pixels are evaluated by the lane that summarised them
#endif

struct alignas(16) CoverWarpSmem {
  uint32_t D[16][COVER_CW / 4];      // directly emitted coverage, one byte per pixel
  uint32_t A[16][COVER_CW / 2];      // accumulated coverage, 16 bits per pixel (saturated when read)
  uint32_t z_rec[COVER_ND];          // record index in the pool
  int32_t z_base[COVER_ND + 1];      // first work unit of the record (exclusive prefix)
  int16_t z_L[COVER_ND], z_R[COVER_ND];    // pixels [L, R) get a value; chunk-relative, clipped
  int16_t z_jl[COVER_ND], z_jr[COVER_ND];  // pixels [jl, jr) get `full`
  uint16_t z_fa[COVER_ND];           // full | accum << 8 | row << 12
  int32_t r_pre[16], r_cnt[16];
  uint32_t r_first[16];
};

// One warp per (op, tile row).  The tile row is processed in chunks of COVER_CW pixels whose coverage
// is built in shared memory, then cut into 16x16 A8 masks:
//   1. the row's trapezoid records (all 16 pixel rows, COVER_ND per pass) are summarised in parallel
//      as ranges: outside / edge zone / interior;
//   2. the edge-zone pixels — the only ones that need the triangle/ramp formulas — are split into
//      units of COVER_UNIT pixels and evaluated by all 32 lanes (unit -> record by binary search);
//      values go straight into the chunk (byte store for direct spans, 16-bit atomic add for
//      accumulated ones: safely_add_alpha saturates, and a saturating sum is order-independent);
//   3. interiors are filled one record per lane;
//   4. lane L then owns pixel row L>>1 and the 8-pixel half L&1 of each tile: two vector loads,
//      classification (empty / solid / one plane / two planes) by ballot, coalesced 256-byte stores.
#ifndef COVER_MINB
#define COVER_MINB 32  // 64 registers.  Measured, C1 / C4a coverage (ms): unset 0.629 / 23.6, 24 blocks 0.595 / 22.4, 28: 0.609 / 23.0,
                       // 32: 0.581 / 21.5; two warps per block with 14 / 16 blocks: 0.632 / 24.1, 0.638 / 24.5; four warps: 0.694 / 27.6
#endif
#ifdef COVER_MINB  // minimum resident blocks per SM (caps the registers); unset = the compiler's own choice
#define COVER_BOUNDS __launch_bounds__(COVER_WARPS * 32, COVER_MINB)
#else
#define COVER_BOUNDS __launch_bounds__(COVER_WARPS * 32)
#endif
__global__ void COVER_BOUNDS k_cover(CoverArgs c) {
  __shared__ CoverWarpSmem sm_all[COVER_WARPS];
  const int wib = threadIdx.x >> 5;
  const uint32_t trow = blockIdx.x * COVER_WARPS + wib;
  const int lane = threadIdx.x & 31;
  if (trow >= c.n_trows) return;
  CoverWarpSmem& sm = sm_all[wib];
  const uint32_t op = c.trow_op[trow];
  const OpGeom g = c.geom[op];
  const uint32_t tr = trow - c.row_base[op] / SKB_TILE;
  const int ty = g.ty0 + (int)tr;
  const skb_dl_op o = c.ops[op];
  if (o.kind != SKB_OP_FILL || o.clip_in != 0) return;  // clip paths and clipped draws: k_clip_rows
  const SurfDesc sd = c.surfs[o.surface];
  const int xmax = min(g.scan_r, (int)sd.w);
  const int xmin = max(g.scan_l, 0);
  // blend modes that change the destination under a zero source: remember which pixels a direct span touched
  const bool zmode = c.zmask != nullptr && (blend_zero_src_matters(paint_blend_mode(c.paints[o.paint])) ||
                                            SKB_PAINT_CF_OFFSET(c.paints[o.paint]) != 0);
  // lanes 0..15 own the row table entries of the 16 pixel rows
  uint2 row = make_uint2(0u, 0u);
  {
    const int y = ty * SKB_TILE + lane;
    if (lane < 16 && y >= g.scan_t && y < g.scan_b && y < (int)sd.h) row = c.rows[c.row_base[op] + tr * SKB_TILE + (uint32_t)lane];
  }
  // rows written by the row-parallel walk hold their records one after the other (no chunk links)
  const uint32_t lin_mask = __ballot_sync(0xffffffffu, (row.y & SKB_ROW_LINEAR) != 0);
  row.y &= ~SKB_ROW_LINEAR;
  int pre = (int)row.y;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, pre, d);
    if (lane >= d) pre += t;
  }
  const int n_tot = __shfl_sync(0xffffffffu, pre, 31);
  if (n_tot == 0) return;  // nothing in this tile row (item flags were zeroed)
  if (lane < 16) {
    sm.r_pre[lane] = pre - (int)row.y;
    sm.r_cnt[lane] = (int)row.y;
    sm.r_first[lane] = row.x;
  }
  __syncwarp();
  const uint32_t item_row = c.item_base[op] + tr * (uint32_t)g.ntx;
  uint8_t* const Db = reinterpret_cast<uint8_t*>(&sm.D[0][0]);
  uint32_t* const Aw = &sm.A[0][0];

  const int x_end_all = (g.tx0 + g.ntx) * SKB_TILE;
  for (int cx = g.tx0 * SKB_TILE; cx < x_end_all; cx += COVER_CW) {
    const int cx_end = min(cx + COVER_CW, x_end_all);
    const int cl = max(xmin, cx), cr = min(xmax, cx_end);
    if (cr <= cl) continue;
    // zero the chunk
    __syncwarp();
    {
      uint4* z = reinterpret_cast<uint4*>(&sm.D[0][0]);
      constexpr int n16 = (int)((sizeof(sm.D) + sizeof(sm.A)) / 16);
      static_assert(offsetof(CoverWarpSmem, A) == sizeof(sm.D), "D and A are zeroed as one block");
#pragma unroll
      for (int i = 0; i < n16 / 32; i++) z[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    }
    int lo = INT_MAX, hi = INT_MIN;
    for (int pbase = 0; pbase < n_tot; pbase += COVER_ND) {
      const int npass = min(COVER_ND, n_tot - pbase);
      __syncwarp();
      // ---- 1. range summaries
      for (int p = lane; p < COVER_ND; p += 32) {
        int units = 0;
        if (p < npass) {
          const int P = pbase + p;
          // the row that owns record P: the last row whose prefix is <= P (it has records: the next prefix is > P)
          int rr_ = sm.r_pre[8] <= P ? 8 : 0;
          rr_ += sm.r_pre[rr_ + 4] <= P ? 4 : 0;
          rr_ += sm.r_pre[rr_ + 2] <= P ? 2 : 0;
          rr_ += sm.r_pre[rr_ + 1] <= P ? 1 : 0;
          const uint32_t first = sm.r_first[rr_];
          // k-th record of the row: consecutive slots, the last slot of every chunk links to the next chunk
          uint32_t idx;
          if ((lin_mask >> rr_) & 1u) {
            idx = first + (uint32_t)(P - sm.r_pre[rr_]);
          } else {
            uint32_t pos = (first & (SKB_CHUNK - 1)) + (uint32_t)(P - sm.r_pre[rr_]);
            uint32_t cb = first & ~(SKB_CHUNK - 1);
            while (pos >= SKB_CHUNK - 1) {
              cb = (uint32_t)c.pool[cb | (SKB_CHUNK - 1)].y;
              pos -= SKB_CHUNK - 1;
            }
            idx = cb + pos;
          }
          const TrapPrep pr = trap_prepare(c.pool[idx]);
          int L = max(pr.L, cl), R = min(pr.mode ? pr.R : pr.L, cr);
          if (R > L) {
            lo = min(lo, L);
            hi = max(hi, R);
          }
          R = max(R, L);
          const int jl = min(max(pr.jl, L), R);
          const int jr = min(max(pr.jr, jl), R);
          sm.z_rec[p] = idx;
          sm.z_L[p] = (int16_t)(L - cx);
          sm.z_R[p] = (int16_t)(R - cx);
          sm.z_jl[p] = (int16_t)(jl - cx);
          sm.z_jr[p] = (int16_t)(jr - cx);
          sm.z_fa[p] = (uint16_t)(pr.full | (pr.accum ? 0x100u : 0u) | ((uint32_t)rr_ << 12));
          if ((jl - L) + (R - jr) <= COVER_INLINE_PX) {
            // short edge zones (steep edges: a pixel or two per side) are evaluated right here, while the prepared
            // record is in registers; only the long ones go to the unit queue, where all lanes share them
            for (int xr = L - cx; xr < R - cx; xr++) {
              if (xr == jl - cx) xr = jr - cx;
              if (xr >= R - cx) break;
              uint8_t v = 0;
              if (!trap_prep_alpha(pr, cx + xr, &v)) continue;
              if (v == 0) {
                if (zmode && !pr.accum) atomicAdd(&Aw[(rr_ * COVER_CW + xr) >> 1], 0x8000u << (16 * (xr & 1)));
                continue;
              }
              if (pr.accum) atomicAdd(&Aw[(rr_ * COVER_CW + xr) >> 1], (uint32_t)v << (16 * (xr & 1)));
              else Db[rr_ * COVER_CW + xr] = v;
            }
          } else {
            units = (jl - L + COVER_UNIT - 1) / COVER_UNIT + (R - jr + COVER_UNIT - 1) / COVER_UNIT;
          }
        }
        sm.z_base[p] = units;
      }
      __syncwarp();
      // exclusive prefix of the unit counts
      int carry = 0;
#pragma unroll
      for (int b = 0; b < COVER_ND; b += 32) {
        const int v = sm.z_base[b + lane];
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += t;
        }
        sm.z_base[b + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      const int total = carry;
      if (lane == 0) sm.z_base[COVER_ND] = total;
      __syncwarp();
      // ---- 2. edge-zone pixels
      for (int u = lane; u < total; u += 32) {
        int lo_d = 0, hi_d = COVER_ND;
        while (hi_d - lo_d > 1) {
          int mid = (lo_d + hi_d) >> 1;
          if (sm.z_base[mid] <= u) lo_d = mid; else hi_d = mid;
        }
        const int e = lo_d;
        const int off = u - sm.z_base[e];
        const int L = sm.z_L[e], jl = sm.z_jl[e], jr = sm.z_jr[e], R = sm.z_R[e];
        const int nl = (jl - L + COVER_UNIT - 1) / COVER_UNIT;
        int xs, xe;
        if (off < nl) {
          xs = L + off * COVER_UNIT;
          xe = min(xs + COVER_UNIT, jl);
        } else {
          xs = jr + (off - nl) * COVER_UNIT;
          xe = min(xs + COVER_UNIT, R);
        }
        const uint32_t fa = sm.z_fa[e];
        const int rr_ = (int)(fa >> 12);
        const bool accum = (fa >> 8) & 1;
        const TrapPrep pr = trap_prepare(c.pool[sm.z_rec[e]]);
        for (int xr = xs; xr < xe; xr++) {
          uint8_t v = 0;
          if (!trap_prep_alpha(pr, cx + xr, &v)) continue;
          if (v == 0) {  // bit 15 of the pixel's 16-bit sum: touched by a direct span whose coverage is 0
            if (zmode && !accum) atomicAdd(&Aw[(rr_ * COVER_CW + xr) >> 1], 0x8000u << (16 * (xr & 1)));
            continue;
          }
          if (accum) atomicAdd(&Aw[(rr_ * COVER_CW + xr) >> 1], (uint32_t)v << (16 * (xr & 1)));
          else Db[rr_ * COVER_CW + xr] = v;
        }
      }
      // ---- 3. interiors
      for (int p = lane; p < npass; p += 32) {
        int x = sm.z_jl[p];
        const int xe = sm.z_jr[p];
        if (xe <= x) continue;
        const uint32_t fa = sm.z_fa[p];
        const int rr_ = (int)(fa >> 12);
        const uint32_t full = fa & 0xFF;
        if ((fa >> 8) & 1) {
          if (full == 0) continue;
          uint32_t* Ar = Aw + rr_ * (COVER_CW / 2);
          if (x & 1) {
            atomicAdd(&Ar[x >> 1], full << 16);
            x++;
          }
          for (; x + 2 <= xe; x += 2) atomicAdd(&Ar[x >> 1], full | (full << 16));
          if (x < xe) atomicAdd(&Ar[x >> 1], full);
        } else {
          uint8_t* Dr = Db + rr_ * COVER_CW;
          for (; x < xe && (x & 3); x++) Dr[x] = (uint8_t)full;
          for (; x + 4 <= xe; x += 4) *reinterpret_cast<uint32_t*>(Dr + x) = full * 0x01010101u;
          for (; x < xe; x++) Dr[x] = (uint8_t)full;
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    __syncwarp();
    if (hi <= lo) continue;

    // ---- 4. cut the chunk into tiles, classify and store
    const int tx_begin = lo / SKB_TILE;
    const int tx_end = (hi + SKB_TILE - 1) / SKB_TILE;
    const int prow = lane >> 1, half = lane & 1;
    for (int tx = tx_begin; tx < tx_end; tx++) {
      const int xr = tx * SKB_TILE - cx + half * 8;
      const uint2 dv = *reinterpret_cast<const uint2*>(Db + prow * COVER_CW + xr);
      uint4 av = *reinterpret_cast<const uint4*>(Aw + ((prow * COVER_CW + xr) >> 1));
      uint32_t z0 = 0, z1 = 0;
      if (zmode) {
        z0 = __byte_perm((av.x >> 15) & 0x00010001u, (av.y >> 15) & 0x00010001u, 0x6420);
        z1 = __byte_perm((av.z >> 15) & 0x00010001u, (av.w >> 15) & 0x00010001u, 0x6420);
        av.x &= 0x7FFF7FFFu;
        av.y &= 0x7FFF7FFFu;
        av.z &= 0x7FFF7FFFu;
        av.w &= 0x7FFF7FFFu;
      }
      // saturate the 16-bit sums and pack them to bytes
      av.x = __vminu2(av.x, 0x00FF00FFu);
      av.y = __vminu2(av.y, 0x00FF00FFu);
      av.z = __vminu2(av.z, 0x00FF00FFu);
      av.w = __vminu2(av.w, 0x00FF00FFu);
      const uint32_t a0 = __byte_perm(av.x, av.y, 0x6420), a1 = __byte_perm(av.z, av.w, 0x6420);
      // bytes that are non-zero, as 0x80 flags
      const uint32_t nd0 = ((dv.x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | dv.x) & 0x80808080u;
      const uint32_t nd1 = ((dv.y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | dv.y) & 0x80808080u;
      const uint32_t na0 = ((a0 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | a0) & 0x80808080u;
      const uint32_t na1 = ((a1 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | a1) & 0x80808080u;
      const bool any = __any_sync(0xffffffffu, (dv.x | dv.y | a0 | a1 | z0 | z1) != 0);
      if (!any) continue;
      const bool anyz = zmode && __any_sync(0xffffffffu, (z0 | z1) != 0);
      // a pixel carries two spans when it has an accumulated value and a direct one — including a direct span
      // whose coverage is 0 where that still blends (z)
      const bool both = __any_sync(0xffffffffu, (((nd0 | (z0 << 7)) & na0) | ((nd1 | (z1 << 7)) & na1)) != 0);
      const bool solid = __all_sync(0xffffffffu, (dv.x & dv.y) == 0xFFFFFFFFu && (a0 | a1) == 0);
      const uint32_t item = item_row + (uint32_t)(tx - g.tx0);
      uint32_t flags = SKB_ITEM_PLANE0;
      if (anyz) {
        flags |= SKB_ITEM_ZERO;
        reinterpret_cast<uint2*>(c.zmask + (size_t)item * 256)[lane] = make_uint2(z0, z1);
      }
      if (solid) {
        flags |= SKB_ITEM_SOLID;
      } else {
        reinterpret_cast<uint2*>(c.mask0 + (size_t)item * 256)[lane] = both ? dv : make_uint2(dv.x | a0, dv.y | a1);
        if (both) {
          flags |= SKB_ITEM_PLANE1;
          reinterpret_cast<uint2*>(c.mask1 + (size_t)item * 256)[lane] = make_uint2(a0, a1);
        }
      }
      if (lane == 0) {
        c.item_flags[item] = (uint16_t)flags;
        uint32_t tile = sd.tile_base + (uint32_t)ty * sd.tiles_x + (uint32_t)tx;
        atomicAdd(&c.tile_cnt[tile], (flags & SKB_ITEM_PLANE1) ? 2u : 1u);
      }
    }
  }
}

// ------------------------------------------------ stages 2-3, coverage mode AREA (skb_area.cuh)
// Unclipped fills under SKB_COVERAGE_AREA take this route instead of walk + k_cover: lines are flattened and binned to
// (op, tile) items, the backdrop deltas are prefix-summed along each tile row, and one warp per tile row accumulates
// the signed areas of each tile's lines into the same A8 masks / item flags k_cover writes.
struct AreaArgs {
  FrameTables t;
  const OpGeom* geom;
  const uint32_t* seg_op;
  const uint32_t* line_off;    // n_segs + 1: first global line of every segment (scanned k_area_seg_count)
  uint32_t n_lines;
  uint32_t* item_cnt;          // n_items + 1: lines binned to the item; scanned in place into item offsets
  uint32_t* item_cursor;       // n_items: write cursors of the second binning pass
  int32_t* item_local;         // n_items: local backdrop; k_area_backdrop turns it into the tile's backdrop
  int32_t* item_delta;         // n_items: backdrop delta applied to the tiles on the right
  int32_t* row_backdrop;       // n_trows: crossings left of the op's tile rectangle
  uint4* lines;                // binned lines: from_x | from_y << 16, to_x | to_y << 16 (8.8), key, 0
  CoverArgs c;
};

__global__ void k_area_seg_count(AreaArgs a, uint32_t* cnt) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.t.n_segs) return;
  const uint32_t op = a.seg_op[s];
  const OpGeom& g = a.geom[op];
  cnt[s] = (g.area && !g.culled) ? (uint32_t)area_seg_line_count(a.t.segs[s], a.t.ops[op].ctm) : 0u;
}

struct AreaBinSink {
  int tx0, ty0, ntx, nty;
  uint32_t item_base, trow_base, key;
  bool write;
  const AreaArgs* a;
  __device__ __forceinline__ void line(V2 p, V2 q, int tx, int ty, int aux) {
    const int dx = tx - tx0, dy = ty - ty0;
    if (dx < 0 || dy < 0 || dx >= ntx || dy >= nty) return;   // CoverageAATileRect::Contains
    uint32_t w0 = 0, w1 = 0;
    int local = 0;
    const int kind = area_tile_line(p, q, tx, ty, &w0, &w1, &local);
    if (kind == 0) return;
    const uint32_t item = item_base + (uint32_t)(dy * ntx + dx);
    if (kind == 2) {
      if (!write) atomicAdd(&a->item_local[item], local);
      return;
    }
    if (!write) {
      atomicAdd(&a->item_cnt[item], 1u);
    } else {
      const uint32_t pos = a->item_cnt[item] + atomicAdd(&a->item_cursor[item], 1u);
      a->lines[pos] = make_uint4(w0, w1, key | (uint32_t)aux, 0u);
    }
  }
  __device__ __forceinline__ void backdrop(int tx, int ty, int delta) {
    if (write) return;
    const int dx = tx - tx0, dy = ty - ty0;
    if (dy < 0 || dy >= nty || dx >= ntx) return;
    if (dx < 0) atomicAdd(&a->row_backdrop[trow_base + (uint32_t)dy], delta);
    else atomicAdd(&a->item_delta[item_base + (uint32_t)(dy * ntx + dx)], delta);
  }
};

// One thread per flattened line, run twice: pass 0 counts the lines of every item and accumulates the backdrops (integer
// atomics: order-independent), pass 1 writes the lines behind the scanned offsets.  The key (global line number, auxiliary
// bit) is the order the reference's sequential tiler would have emitted them in.
__global__ void k_area_bin(AreaArgs a, int write) {
  uint32_t ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= a.n_lines) return;
  const uint32_t seg = find_interval(a.line_off, a.t.n_segs, ln);
  const int k = (int)(ln - a.line_off[seg]);
  const int n = (int)(a.line_off[seg + 1] - a.line_off[seg]);
  const uint32_t op = a.seg_op[seg];
  const OpGeom& g = a.geom[op];
  if (g.empty || g.ntx <= 0 || g.nty <= 0) return;
  V2 from, to;
  area_seg_line(a.t.segs[seg], a.t.ops[op].ctm, k, n, &from, &to);
  if (!(finite_f(from.x) && finite_f(from.y) && finite_f(to.x) && finite_f(to.y))) return;
  AreaBinSink sink;
  sink.tx0 = g.tx0;
  sink.ty0 = g.ty0;
  sink.ntx = g.ntx;
  sink.nty = g.nty;
  sink.item_base = a.c.item_base[op];
  sink.trow_base = a.c.row_base[op] / SKB_TILE;
  sink.key = ln << 1;
  sink.write = write != 0;
  sink.a = &a;
  area_walk_line(from, to, sink);
}

// ResolveBackdrops: one thread per tile row of an AREA op.
__global__ void k_area_backdrop(AreaArgs a) {
  uint32_t trow = blockIdx.x * blockDim.x + threadIdx.x;
  if (trow >= a.c.n_trows) return;
  const uint32_t op = a.c.trow_op[trow];
  const OpGeom& g = a.geom[op];
  if (!g.area) return;
  const uint32_t tr = trow - a.c.row_base[op] / SKB_TILE;
  const uint32_t item0 = a.c.item_base[op] + tr * (uint32_t)g.ntx;
  int acc = a.row_backdrop[trow];
  for (int x = 0; x < g.ntx; x++) {
    const int b = acc + a.item_local[item0 + x];
    acc += a.item_delta[item0 + x];
    a.item_local[item0 + x] = b;
  }
}

// Block shape of k_area_cover, measured on C4a (coverage stage, ms): 4 warps / 128 lines / compiler's registers (80) 67.5;
// 4 warps / 32 lines with at least 8 / 10 / 12 / 14 blocks per SM 49.7 / 45.2 / 46.2 / 51.5; 2 warps / 32 lines / 24 blocks
// (42 registers, 48 warps per SM) 44.0 — the kernel is a chain of short dependent phases per tile and needs the warps.
#ifndef AREA_WARPS
#define AREA_WARPS 2
#endif
#ifndef AREA_CAP
#define AREA_CAP 32                // lines of a tile held in shared memory at a time (more: the plain form below)
#endif
#ifndef AREA_MINB
#define AREA_MINB 24
#endif
struct alignas(16) AreaWarpSmem {
  AreaLine line[AREA_CAP];
  uint32_t key[AREA_CAP];
  int32_t cover[16][17];   // per pixel row: 8.8 sums of the lines' "pixel lies right of the line" terms, scattered at the
                           // first column right of each line, then prefix-summed along the row
  uint32_t mixed[16];      // per pixel row: bit c = pixel (row, c) lies under some line (needs the full formula, in line order)
  uint8_t amix[256];       // alphas of those pixels
  uint8_t list[256];       // which pixels they are
};

// One warp per (op, tile row); tiles left to right.  A tile's lines are put in key order (rank sort in shared memory)
// and unpacked once.  The reference sums a pixel's contributions in fp32 in line order, so the order must be kept — but
// only where it matters: a line contributes exactly 0 to the pixels left of it and an exact multiple of 1/256 (the
// height of its overlap with the pixel row, signed) to the pixels right of it; only the one or two pixels per row that
// lie UNDER the line get an inexact area term.  So per tile:
//   1. every (line, pixel row) pair is classified by a lane: the columns under the line (with the reference's own
//      expression for the line's y at a pixel edge, so that "left" and "right" are exactly the cases in which the
//      reference's clamps collapse) are marked in a per-row bit mask, the right-of-line term goes into an integer 8.8
//      grid at the first column right of the line (integer adds: order-free);
//   2. the grid is prefix-summed along the rows: the winding of every pixel no line passes over, exactly (sums of
//      multiples of 1/256 below 2^15 are exact in fp32 in any order);
//   3. the marked pixels — typically a tenth of the tile — are dealt to the lanes, 32 at a time, and each is evaluated
//      the reference's way: all lines in key order, full formula;
//   4. every lane resolves the alphas of its 8 pixels (lane -> pixel row lane >> 1, half lane & 1; integer arithmetic for
//      the pixels of step 2, on which the reference's fp32 resolve is exact) and the tile's mask is stored.
// A tile with more than AREA_CAP lines (first sorted in place in global memory by selection, then streamed) takes the
// plain form: every lane accumulates its 8 pixels over all lines that reach its row.
__global__ void __launch_bounds__(AREA_WARPS * 32, AREA_MINB) k_area_cover(AreaArgs a) {
  __shared__ AreaWarpSmem sm_all[AREA_WARPS];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t trow = blockIdx.x * AREA_WARPS + wib;
  if (trow >= a.c.n_trows) return;
  const uint32_t op = a.c.trow_op[trow];
  const OpGeom g = a.geom[op];
  if (!g.area || g.empty || g.ntx <= 0) return;
  AreaWarpSmem& sm = sm_all[wib];
  const skb_dl_op& o = a.c.ops[op];
  const SurfDesc sd = a.c.surfs[o.surface];
  const uint32_t tr = trow - a.c.row_base[op] / SKB_TILE;
  const int ty = g.ty0 + (int)tr;
  const int even_odd = (int)o.fill_type;
  const int xmin = max(g.scan_l, 0), xmax = min(g.scan_r, (int)sd.w);
  const int ymin = max(g.scan_t, 0), ymax = min(g.scan_b, (int)sd.h);
  const int prow = lane >> 1, half = lane & 1;
  const int y = ty * SKB_TILE + prow;
  const bool row_in = y >= ymin && y < ymax;
  const float pixel_top = (float)prow, pixel_bottom = (float)prow + 1.0f;
  const uint32_t item_row = a.c.item_base[op] + tr * (uint32_t)g.ntx;
  for (int txi = 0; txi < g.ntx; txi++) {
    const uint32_t item = item_row + (uint32_t)txi;
    const uint32_t beg = a.item_cnt[item];
    const int n = (int)(a.item_cnt[item + 1] - beg);
    const int backdrop = a.item_local[item];
    if (n == 0 && (backdrop == 0 || (even_odd && !(backdrop & 1)))) continue;   // nothing inside: alpha 0 (uniform over the warp)
    const int x0 = (g.tx0 + txi) * SKB_TILE + half * 8;
    // bit i: pixel i of this lane's 8 lies in the scan rectangle — all 8, as a span of bits (columns lo .. hi - 1)
    uint32_t inmask = 0;
    if (row_in) {
      const int lo = min(max(xmin - x0, 0), 8), hi = min(max(xmax - x0, 0), 8);
      inmask = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
    }
    uint32_t d0 = 0, d1 = 0;
    if (n == 0) {
      const uint32_t v = area_alpha_u8(area_resolve_alpha((float)backdrop, even_odd));
      if (v == 0) continue;   // uniform over the warp
#pragma unroll
      for (int i = 0; i < 4; i++) {
        d0 |= ((inmask >> i) & 1u) ? v << (8 * i) : 0u;
        d1 |= ((inmask >> (4 + i)) & 1u) ? v << (8 * i) : 0u;
      }
    } else {
      if (n > AREA_CAP) {
        // in-place selection sort by key (rare: more lines in one 16x16 tile than the shared-memory window holds)
        for (int i = 0; i < n - 1; i++) {
          uint32_t best = 0xFFFFFFFFu;
          int best_j = i;
          for (int j = i + lane; j < n; j += 32) {
            const uint32_t kj = a.lines[beg + j].z;
            if (kj < best) { best = kj; best_j = j; }
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oj = __shfl_xor_sync(0xffffffffu, best_j, d);
            if (ob < best) { best = ob; best_j = oj; }
          }
          if (lane == 0 && best_j != i) {
            const uint4 t0 = a.lines[beg + i];
            a.lines[beg + i] = a.lines[beg + best_j];
            a.lines[beg + best_j] = t0;
          }
          __syncwarp();
        }
      }
      if (n <= AREA_CAP) {
        const int m = n;
        __syncwarp();
        {   // lines in key order, unpacked
          uint4 mine[AREA_CAP / 32];
#pragma unroll
          for (int q = 0; q < AREA_CAP / 32; q++) {
            const int j = lane + 32 * q;
            if (j < m) {
              mine[q] = a.lines[beg + j];
              sm.key[j] = mine[q].z;
            }
          }
          for (int j = lane; j < 16 * 17; j += 32) (&sm.cover[0][0])[j] = 0;
          if (lane < 16) sm.mixed[lane] = 0u;
          __syncwarp();
#pragma unroll
          for (int q = 0; q < AREA_CAP / 32; q++) {
            const int j = lane + 32 * q;
            if (j < m) {
              int rank = 0;
              for (int i = 0; i < m; i++) rank += sm.key[i] < mine[q].z ? 1 : 0;
              sm.line[rank] = area_line_unpack(mine[q].x, mine[q].y);
            }
          }
        }
        __syncwarp();
        // 1. classify (line, pixel row) pairs: two lines at a time, 16 rows each
        for (int k0 = 0; k0 < m; k0 += 2) {
          const int k = k0 + (lane >> 4), r = lane & 15;
          if (k < m) {
            const AreaLine l = sm.line[k];
            const float rt = (float)r, rb = (float)r + 1.0f;
            const float y_min = l.edge_top < rt ? rt : l.edge_top;
            const float y_max = l.edge_bottom < rb ? l.edge_bottom : rb;
            if (y_min < y_max) {
              int lo, hi;      // columns [lo, hi] lie under the line in this row; columns > hi right of it, < lo left of it
              float rterm;     // what a pixel right of the line receives
              if (l.dx == 0.0f) {
                // vertical: sign * h * clamp(pixel_right - x, 0, 1) — 0 for pixel_right <= x, sign * h for pixel_right >= x + 1
                lo = hi = (int)ceilf(l.fx_) - 1;
                rterm = l.sign * (y_max - y_min) * 1.0f;
              } else {
                const bool up = l.y_slope > 0.0f;   // the line's y grows with x
                const float xa = l.fx_ + (y_min - l.fy_) * l.x_slope, xb = l.fx_ + (y_max - l.fy_) * l.x_slope;
                lo = (int)floorf(xa < xb ? xa : xb);
                hi = (int)floorf(xa < xb ? xb : xa);
                lo = lo < 0 ? 0 : (lo > 15 ? 15 : lo);
                hi = hi < lo ? lo : (hi > 15 ? 15 : hi);
                // Widen until the reference's own expressions say "right" / "left": the line's y at the left edge of
                // column hi + 1 is at or beyond the row overlap's far end (then both clamped ys of every pixel from there
                // on coincide: zero area, full cover term), its y at the right edge of column lo - 1 at or before the near
                // end (zero area, zero cover).  The expression is monotone in the column, so one test each side decides.
                while (hi < 15) {
                  const float ly = l.fy_ + ((float)(hi + 1) - l.fx_) * l.y_slope;
                  if (up ? ly >= y_max : ly <= y_min) break;
                  hi++;
                }
                while (lo > 0) {
                  const float ry = l.fy_ + ((float)lo - l.fx_) * l.y_slope;   // right edge of column lo - 1
                  if (up ? ry <= y_min : ry >= y_max) break;
                  lo--;
                }
                const float ley = area_clampf(l.left_endpoint_y, y_min, y_max);
                rterm = l.sign * fabsf((up ? y_max : y_min) - ley);
              }
              if (hi >= 0 && lo <= 15) {
                const int c0 = lo < 0 ? 0 : lo, c1 = hi > 15 ? 15 : hi;
                if (c1 >= c0) atomicOr(&sm.mixed[r], ((2u << c1) - 1u) & ~((1u << c0) - 1u));
              }
              if (hi < 15) atomicAdd(&sm.cover[r][hi < -1 ? 0 : hi + 1], (int32_t)(rterm * 256.0f));
            }
          }
        }
        __syncwarp();
        // 2. windings of the pixels no line passes over; 3a. the list of the others
        int cnt = 0;
        if (lane < 16) {
          int acc = 0;
#pragma unroll
          for (int c = 0; c < 16; c++) {
            acc += sm.cover[lane][c];
            sm.cover[lane][c] = acc;
          }
          cnt = __popc(sm.mixed[lane]);
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane < 16) {
          int at = incl - cnt;
          for (uint32_t mm = sm.mixed[lane]; mm; mm &= mm - 1) sm.list[at++] = (uint8_t)(lane * 16 + (__ffs((int)mm) - 1));
        }
        __syncwarp();
        // 3b. the pixels under lines: the reference's loop, all lines in key order
        for (int b = 0; b < total; b += 32) {
          const int idx = b + lane;
          if (idx < total) {
            const int pid = sm.list[idx];
            const float pt = (float)(pid >> 4), pb = (float)(pid >> 4) + 1.0f, pxf = (float)(pid & 15);
            float wv = (float)backdrop;
            for (int k = 0; k < m; k++) {
              const AreaLine& l = sm.line[k];
              const float y_min = l.edge_top < pt ? pt : l.edge_top;
              const float y_max = l.edge_bottom < pb ? l.edge_bottom : pb;
              if (y_min >= y_max) continue;
              wv = wv + area_edge_contribution(l, pxf, y_min, y_max);
            }
            sm.amix[pid] = (uint8_t)area_alpha_u8(area_resolve_alpha(wv, even_odd));
          }
        }
        __syncwarp();
        // 4. alphas of this lane's 8 pixels.  A pixel no line passes over has a winding k / 256 with k an integer, and
        //    coverage_aa_resolve_alpha + the A8 quantisation are exact on such values: |w| (or its distance to the nearest
        //    even integer) clamped to 1, times 255, plus 0.5, truncated = (a * 255 + 128) >> 8 with a in 1/256ths.
        const uint32_t mrow = sm.mixed[prow];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int c = half * 8 + i;
          const int kw = backdrop * 256 + sm.cover[prow][c];
          uint32_t ak = (uint32_t)(kw < 0 ? -kw : kw);
          if (even_odd) {
            const uint32_t kk = ak & 511u;
            ak = kk > 256u ? 512u - kk : kk;
          } else {
            ak = ak > 256u ? 256u : ak;
          }
          uint32_t v = (ak * 255u + 128u) >> 8;
          if ((mrow >> c) & 1u) v = sm.amix[prow * 16 + c];
          if (!((inmask >> i) & 1u)) v = 0u;
          if (i < 4) d0 |= v << (8 * i);
          else d1 |= v << (8 * (i - 4));
        }
      } else {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] = (float)backdrop;
      for (int c0 = 0; c0 < n; c0 += AREA_CAP) {
        const int m = min(AREA_CAP, n - c0);
        __syncwarp();
        if (n <= AREA_CAP) {
          uint4 mine[AREA_CAP / 32];
#pragma unroll
          for (int q = 0; q < AREA_CAP / 32; q++) {
            const int j = lane + 32 * q;
            if (j < m) {
              mine[q] = a.lines[beg + j];
              sm.key[j] = mine[q].z;
            }
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < AREA_CAP / 32; q++) {
            const int j = lane + 32 * q;
            if (j < m) {
              int rank = 0;
              for (int i = 0; i < m; i++) rank += sm.key[i] < mine[q].z ? 1 : 0;
              sm.line[rank] = area_line_unpack(mine[q].x, mine[q].y);
            }
          }
        } else {
          for (int j = lane; j < m; j += 32) {
            const uint4 l = a.lines[beg + c0 + j];
            sm.line[j] = area_line_unpack(l.x, l.y);
          }
        }
        __syncwarp();
        for (int k = 0; k < m; k++) {
          const AreaLine& l = sm.line[k];
          const float y_min = l.edge_top < pixel_top ? pixel_top : l.edge_top;
          const float y_max = l.edge_bottom < pixel_bottom ? l.edge_bottom : pixel_bottom;
          if (y_min >= y_max) continue;
          const AreaLine lv = l;
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] = w[i] + area_edge_contribution(lv, (float)(half * 8 + i), y_min, y_max);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        d0 |= ((inmask >> i) & 1u) ? area_alpha_u8(area_resolve_alpha(w[i], even_odd)) << (8 * i) : 0u;
        d1 |= ((inmask >> (4 + i)) & 1u) ? area_alpha_u8(area_resolve_alpha(w[4 + i], even_odd)) << (8 * i) : 0u;
      }
      }
    }
    const bool any = __any_sync(0xffffffffu, (d0 | d1) != 0);
    if (!any) continue;
    const bool solid = __all_sync(0xffffffffu, (d0 & d1) == 0xFFFFFFFFu);
    uint32_t flags = SKB_ITEM_PLANE0;
    if (solid) flags |= SKB_ITEM_SOLID;
    else reinterpret_cast<uint2*>(a.c.mask0 + (size_t)item * 256)[lane] = make_uint2(d0, d1);
    if (lane == 0) {
      a.c.item_flags[item] = (uint16_t)flags;
      const uint32_t tile = sd.tile_base + (uint32_t)ty * sd.tiles_x + (uint32_t)(g.tx0 + txi);
      atomicAdd(&a.c.tile_cnt[tile], 1u);
    }
  }
}

// ------------------------------------------------------------- stage 4b: path clips
// Clip states and clipped draws are produced row by row (skb_clip.cuh): one thread sweeps one pixel
// row of one op.  A clip state is a per-pixel table of up to SKB_CLIP_MAXE (span start, coverage)
// entries over the clip path's scan rectangle; a clipped draw becomes up to SKB_CLIP_PLANES
// coverage planes (the k-th coverage blended into a pixel lives in plane k).
struct ClipStateDesc {
  int32_t rx0, ry0, rw, rh;  // region covered by the table (scan rectangle + 1 column, on the surface)
  uint32_t table_off;        // first pixel of the table (in pixels)
  uint32_t nonempty;         // SWCanvas::State::HasClip(): some span exists
  uint32_t op;
  uint32_t kind;             // 0: intersecting state (per-pixel entry table); 1: ClipOp::kDifference state (per-row sorted span lists)
};
// A difference state lives in the same arena as an intersecting one (8 words per pixel of its region, rows of rw pixels):
// row r holds [n_spans, 0, (x, len) * n_spans]; a row of rw pixels has at most rw direct and rw accumulated spans, and
// 2 + 4 * rw <= 8 * rw.
#define SKB_CLIP_KIND_DIFF 1u

// A ClipOp::kDifference clip applied while a path clip is in force ("T2": RecursiveClip's spans_subtraction(clip spans,
// fresh spans), sw_canvas.cc:188-189): the result is an intersecting state over the PARENT's region — the region of the
// first clip of the chain (`region_op`) —, made by k_clip_t2 from the parent's table and the fresh path's sorted row
// spans.  The records are worked out on the host (validate_dl) in op order.
struct ClipT2Rec {
  uint32_t op, region_op, depth, pad;
};

// the scan rectangle of a clip path as a table region (+ the column FindSpan's `+ 1` can reach), on the surface or not:
// HasClip() and nested clips see every span of a clip path (sw_canvas.cc:315-336)
__device__ __forceinline__ void clip_region_of(const OpGeom& g, int32_t* rx0, int32_t* ry0, int32_t* rw, int32_t* rh) {
  *rx0 = *ry0 = *rw = *rh = 0;
  if (!g.empty && g.ntx > 0) {
    *rx0 = g.scan_l;
    *ry0 = g.scan_t;
    *rw = g.scan_r + 1 - g.scan_l;
    *rh = g.scan_b - g.scan_t;
    if (*rw <= 0 || *rh <= 0) *rw = *rh = 0;
  }
}

__global__ void k_clip_sizes(FrameTables t, const OpGeom* geom, const SurfDesc* surfs, ClipStateDesc* states, uint32_t* px_cnt,
                             const ClipT2Rec* t2, uint32_t n_t2) {
  uint32_t op = blockIdx.x * blockDim.x + threadIdx.x;
  if (op >= t.n_ops) return;
  const skb_dl_op o = t.ops[op];
  if (o.kind != SKB_OP_CLIP) return;
  ClipStateDesc d;
  d.table_off = 0;
  d.nonempty = 0;
  d.op = op;
  d.kind = o.aux == 0 ? SKB_CLIP_KIND_DIFF : 0u;
  clip_region_of(geom[op], &d.rx0, &d.ry0, &d.rw, &d.rh);
  uint32_t px = (uint32_t)(d.rw * d.rh);
  if (o.aux == 0 && o.clip_in != 0) {
    // T2: the state's arena holds the fresh path's sorted row spans (difference layout, the fresh path's own region)
    // followed by the new intersecting table over the parent's region
    uint32_t lo = 0, hi = n_t2;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (t2[mid].op < op) lo = mid + 1; else hi = mid;
    }
    const uint32_t fresh_px = px;
    d.kind = 0u;
    clip_region_of(geom[lo < n_t2 && t2[lo].op == op ? t2[lo].region_op : op], &d.rx0, &d.ry0, &d.rw, &d.rh);
    d.table_off = fresh_px;
    px = fresh_px + (uint32_t)(d.rw * d.rh);
  }
  states[o.clip_out] = d;
  px_cnt[o.clip_out] = px;
}

struct ClipArgs {
  CoverArgs c;
  ClipStateDesc* states;
  const uint32_t* state_px_off;  // scanned px_cnt
  uint32_t* table;               // SKB_CLIP_MAXE entries per pixel
  const uint8_t* op_depth;       // nesting depth of the state a CLIP op defines
  uint32_t* overflow;            // set when a pixel needs more entries / planes than provided
  uint32_t n_rows;
  const ClipT2Rec* t2;           // difference clips on top of a path clip, in op order
  uint32_t n_t2;
};

// Row y of a state's table (intersecting: SKB_CLIP_MAXE entries per pixel; difference: [n, 0, spans ...]).
__device__ __forceinline__ uint32_t* clip_state_row(const ClipArgs& a, uint32_t state, const ClipStateDesc& d, int y) {
  return a.table + ((size_t)a.state_px_off[state] + (size_t)d.table_off + (size_t)(y - d.ry0) * (size_t)d.rw) * SKB_CLIP_MAXE;
}

// mode 0: build the clip states of nesting depth `level`; mode 1: rasterise the clipped draws.
// One WARP per row: every lane takes a run of consecutive pixels (at least SKB_CLIP_SEG_MIN) and, except the
// first, seeks the sweep state to its first pixel (clip_row_seek).  Rows that are not this launch's business
// cost one warp-uniform early exit.
#ifndef SKB_CLIP_SEG_MIN
#define SKB_CLIP_SEG_MIN 2   // measured on C2 with clips: 1: 4.47, 2: 4.51, 4: 4.68, 8: 5.12, 16: 6.13 ms
#endif
__global__ void __launch_bounds__(128) k_clip_rows(ClipArgs a, int mode, int level) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= a.n_rows) return;
  const int seg = (int)(threadIdx.x & 31);
  const CoverArgs& c = a.c;
  const uint32_t op = c.trow_op[r / SKB_TILE];
  const skb_dl_op o = c.ops[op];
  if (mode == 0) {
    if (o.kind != SKB_OP_CLIP || a.op_depth[op] != (uint8_t)level || o.aux == 0) return;   // difference clips: k_clip_diff
    if (o.clip_in != 0 && a.states[o.clip_in].kind == SKB_CLIP_KIND_DIFF) return;            // on top of one: k_clip_diff mode 2
  } else {
    if (o.kind != SKB_OP_FILL || o.clip_in == 0 || a.states[o.clip_in].kind == SKB_CLIP_KIND_DIFF) return;
  }
  const OpGeom g = c.geom[op];
  if (g.empty || g.ntx == 0) return;
  const SurfDesc sd = c.surfs[o.surface];
  const int y = g.ty0 * SKB_TILE + (int)(r - c.row_base[op]);
  if (y < g.scan_t || y >= g.scan_b || (mode == 1 && y >= (int)sd.h)) return;
  uint2 row = c.rows[r];
  row.y &= ~SKB_ROW_LINEAR;   // consecutive records simply contain no chunk link
  if (row.y == 0) return;

  // parent clip state (C side)
  ClipStateDesc par;
  par.rw = par.rh = 0;
  bool clipped = false;
  if (o.clip_in != 0) {
    par = a.states[o.clip_in];
    clipped = par.nonempty != 0;
  }
  const bool c_row_in = clipped && y >= par.ry0 && y < par.ry0 + par.rh;
  const uint32_t* c_row = c_row_in ? clip_state_row(a, o.clip_in, par, y) : nullptr;
  // own state (mode 0)
  ClipStateDesc own;
  uint32_t* own_row = nullptr;
  if (mode == 0) {
    own = a.states[o.clip_out];
    if (own.rw == 0 || y < own.ry0 || y >= own.ry0 + own.rh) return;
    own_row = clip_state_row(a, o.clip_out, own, y);
  }

  // the row's prepared records: one array per warp in shared memory, every lane prepares its share
  __shared__ TrapPrep s_prep[4][SKB_CLIP_RMAX];
  ClipRowState st;
  clip_row_begin(st, c.pool, row, s_prep[threadIdx.x >> 5], seg, 32);
  __syncwarp();
  int x_first = g.scan_l, x_last = g.scan_r;  // inclusive: the pixel after the last span can receive the `+ 1`
  if (st.n_prep >= 0) {
    int lo = INT_MAX, hi = INT_MIN;
    for (int k = 0; k < st.n_prep; k++) {
      if (st.prep[k].mode == 0) continue;
      lo = min(lo, st.prep[k].L);
      hi = max(hi, st.prep[k].R);
    }
    if (hi <= lo) return;
    x_first = max(x_first, lo);
    x_last = min(x_last, hi);
  }
  if (mode == 1) x_last = min(x_last, (int)sd.w - 1);
  // clipped draws store the pixels of the surface only; those left of it merely feed the sweep state
  const int x_out = mode == 1 ? max(x_first, 0) : x_first;
  if (st.n_prep >= 0) {
    const int width = x_last - x_out + 1;
    const int seg_px = max(SKB_CLIP_SEG_MIN, (width + 31) / 32);
    const int x0 = x_out + seg * seg_px;
    if (x0 > x_last) return;
    clip_row_seek(st, c.pool, row, x_first, x0);
    x_first = x0;
    x_last = min(x_last, x0 + seg_px - 1);
    clip_row_focus(st, x_first, x_last);
  } else if (seg != 0) {
    return;  // a row with more records than the state holds is swept by one thread
  }
  const int cap = mode == 0 ? SKB_CLIP_MAXE : SKB_CLIP_PLANES;
  // a clipped draw whose blend mode / colour filter acts on zero-coverage pixels keeps the spans of coverage 0
  bool zm = false;
  if (mode == 1 && c.zplane[1] != nullptr) {
    const skb_dl_paint pt = c.paints[o.paint];
    zm = blend_zero_src_matters(paint_blend_mode(pt)) || SKB_PAINT_CF_OFFSET(pt) != 0;
  }
  const uint32_t item_row = c.item_base[op] + (uint32_t)((y / SKB_TILE) - g.ty0) * (uint32_t)g.ntx;
  bool wrote = false, over = false;
  for (int x = x_first; x <= x_last; x++) {
    SpanSide ld, od, la, oa;
    clip_row_step(st, c.pool, row, x, ld, od, la, oa);
    if (x < x_out) continue;
    if ((mode == 0 || zm) ? !(ld.present | od.present | la.present | oa.present) : (ld.cover | od.cover | la.cover | oa.cover) == 0) continue;
    const uint32_t* clist = nullptr;
    const uint32_t* cprev = nullptr;
    int n_c = 0, n_p = 0;
    if (c_row && x >= par.rx0 && x < par.rx0 + par.rw) {
      clist = c_row + (size_t)(x - par.rx0) * SKB_CLIP_MAXE;
      while (n_c < SKB_CLIP_MAXE && clist[n_c]) n_c++;
    }
    if (mode == 0 && c_row && x - 1 >= par.rx0 && x - 1 < par.rx0 + par.rw) {  // parent spans ending here (zero-length sub-spans)
      cprev = c_row + (size_t)(x - 1 - par.rx0) * SKB_CLIP_MAXE;
      while (n_p < SKB_CLIP_MAXE && cprev[n_p]) n_p++;
    }
    ClipOut out;
    clip_combine(x, ld, od, la, oa, clist, n_c, cprev, n_p, clipped, cap, mode == 0 ? 2 : (zm ? 1 : 0), out);
    over |= out.overflow;
    if (out.n == 0) continue;
    if (mode == 0) {
      if (x >= own.rx0 && x < own.rx0 + own.rw) {
        uint32_t* e = own_row + (size_t)(x - own.rx0) * SKB_CLIP_MAXE;
        for (int k = 0; k < out.n; k++) e[k] = out.e[k];
        wrote = true;
      }
    } else {
      const int tx = x / SKB_TILE;
      if (tx < g.tx0 || tx >= g.tx0 + g.ntx) continue;
      const size_t at = (size_t)(item_row + (uint32_t)(tx - g.tx0)) * 256 + (size_t)(y % SKB_TILE) * SKB_TILE + (x % SKB_TILE);
      for (int k = 0; k < out.n; k++) {
        c.mask[k][at] = (uint8_t)clip_entry_cover(out.e[k]);
        if (zm) c.zplane[k][at] = 1;  // a span reaches the pixel, whatever its coverage
      }
    }
  }
  if (wrote) a.states[o.clip_out].nonempty = 1;
  if (over) *a.overflow = 1;
}

// ClipOp::kDifference (skb_clip.cuh).  One warp per pixel row of an op: the lanes prepare the row's records into shared
// memory, lane 0 walks the row's spans in the reference's list order.
//   mode 0  a difference CLIP op (its state has no parent: validate_dl): the row's spans are stored and sorted by x the
//           way std::sort leaves them;
//   mode 1  a FILL under a difference state: every span of the draw is cut by the state's spans of that row; what is
//           left goes to coverage plane 0 (directly emitted spans) or 1 (accumulated spans) — a pixel has at most one
//           of each, blended in that order, as for an unclipped draw.
// Rare and small next to the draws themselves; no attempt is made to share a row among lanes.
struct DiffStoreD {
  uint2* spans;
  int n;
  __device__ void operator()(int x, int len, uint32_t) { spans[n++] = make_uint2((uint32_t)x, (uint32_t)len); }
};
struct DiffStoreA {
  uint2* spans;   // filled from the END of the row's array backwards (the accumulated spans follow the direct ones in the list)
  int cap, n;
  __device__ void operator()(int x, int len, uint32_t) { spans[cap - 1 - n++] = make_uint2((uint32_t)x, (uint32_t)len); }
};
struct DiffPiece {
  uint8_t* plane;
  uint8_t* zplane;
  const OpGeom* g;
  uint32_t item_row;
  int y, w;
  uint32_t cover;
  __device__ void operator()(int x, int len) {
    for (int px = max(x, 0); px < min(x + len, w); px++) {
      const int tx = px / SKB_TILE;
      if (tx < g->tx0 || tx >= g->tx0 + g->ntx) continue;
      const size_t at = (size_t)(item_row + (uint32_t)(tx - g->tx0)) * 256 + (size_t)(y % SKB_TILE) * SKB_TILE + (px % SKB_TILE);
      plane[at] = (uint8_t)cover;
      if (zplane) zplane[at] = 1;
    }
  }
};
// A piece of a new INTERSECTING state's span (mode 2): appended to the entry lists of the pixels it covers, a
// zero-length piece as a marker entry (skb_clip.cuh) — the same table k_clip_rows builds for nested intersecting clips.
struct DiffTablePiece {
  uint32_t* own_row;
  int rx0, rw;
  uint32_t cover;
  bool wrote, over;
  __device__ void append(int px, uint32_t e) {
    if (px < rx0 || px >= rx0 + rw) return;
    uint32_t* slot = own_row + (size_t)(px - rx0) * SKB_CLIP_MAXE;
    int k = 0;
    while (k < SKB_CLIP_MAXE && slot[k]) k++;
    if (k == SKB_CLIP_MAXE) {
      over = true;
      return;
    }
    slot[k] = e;
    wrote = true;
  }
  __device__ void operator()(int x, int len) {
    if (len == 0) {
      append(x, clip_entry(x, cover) | SKB_CLIP_MARKER);
      return;
    }
    for (int px = x; px < x + len; px++) append(px, clip_entry(x, cover));
  }
};
template <class Piece>
struct DiffCutT {
  Piece piece;
  const uint2* ms;
  int n_ms;
  bool keep_zero;   // spans of coverage 0 matter: the blend mode / colour filter acts on them, or a clip state is being built
  __device__ void operator()(int x, int len, uint32_t cover) {
    if (cover == 0 && !keep_zero) return;
    piece.cover = cover;
    span_subtract(x, len, ms, n_ms, piece);
  }
};
typedef DiffCutT<DiffPiece> DiffCut;

__global__ void __launch_bounds__(128) k_clip_diff(ClipArgs a, int mode, int level) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= a.n_rows) return;
  const int lane = (int)(threadIdx.x & 31);
  const CoverArgs& c = a.c;
  const uint32_t op = c.trow_op[r / SKB_TILE];
  const skb_dl_op o = c.ops[op];
  if (mode == 0) {
    // level 0: the difference states proper (no parent); level > 0: the fresh row spans of the difference clips applied
    // on top of a path clip at that nesting depth (k_clip_t2 consumes them)
    if (o.kind != SKB_OP_CLIP || o.aux != 0) return;
    if (level == 0 ? o.clip_in != 0 : (o.clip_in == 0 || a.op_depth[op] != (uint8_t)level)) return;
  } else if (mode == 1) {
    if (o.kind != SKB_OP_FILL || o.clip_in == 0 || a.states[o.clip_in].kind != SKB_CLIP_KIND_DIFF) return;
  } else {
    if (o.kind != SKB_OP_CLIP || o.aux != 1 || a.op_depth[op] != (uint8_t)level || o.clip_in == 0 ||
        a.states[o.clip_in].kind != SKB_CLIP_KIND_DIFF)
      return;
  }
  const OpGeom g = c.geom[op];
  if (g.empty || g.ntx == 0) return;
  const SurfDesc sd = c.surfs[o.surface];
  const int y = g.ty0 * SKB_TILE + (int)(r - c.row_base[op]);
  if (y < g.scan_t || y >= g.scan_b || (mode == 1 && y >= (int)sd.h)) return;
  uint2 row = c.rows[r];
  row.y &= ~SKB_ROW_LINEAR;
  if (row.y == 0) return;
  __shared__ TrapPrep s_prep[4][SKB_CLIP_RMAX];
  ClipRowState st;
  clip_row_begin(st, c.pool, row, s_prep[threadIdx.x >> 5], lane, 32);
  __syncwarp();
  if (lane != 0) return;
  int x_first = g.scan_l, x_last = g.scan_r;
  if (st.n_prep >= 0) {
    int lo = INT_MAX, hi = INT_MIN;
    for (int k = 0; k < st.n_prep; k++) {
      if (st.prep[k].mode == 0) continue;
      lo = min(lo, st.prep[k].L);
      hi = max(hi, st.prep[k].R);
    }
    if (hi <= lo) return;
    x_first = max(x_first, lo);
    x_last = min(x_last, hi);
  }
  if (mode == 0) {
    // where the row's list goes: the state's own table (a difference state), or the head of a T2 state's arena
    ClipStateDesc own = a.states[o.clip_out];
    if (o.clip_in != 0) {
      clip_region_of(g, &own.rx0, &own.ry0, &own.rw, &own.rh);
      own.table_off = 0;
    }
    if (own.rw == 0 || y < own.ry0 || y >= own.ry0 + own.rh) return;
    uint32_t* base = clip_state_row(a, o.clip_out, own, y);
    uint2* spans = reinterpret_cast<uint2*>(base + 2);
    DiffStoreD sd_{spans, 0};
    DiffStoreA sa_{spans, 2 * own.rw, 0};
    x_first = max(x_first, own.rx0);
    x_last = min(x_last, own.rx0 + own.rw - 1);
    clip_row_spans(st, c.pool, row, x_first, x_last, sd_, sa_);
    // the accumulated spans behind the direct ones, in their own order
    for (int k = 0; k < sa_.n; k++) {
      const uint2 v = spans[2 * own.rw - 1 - k];
      spans[sd_.n + k] = v;
    }
    const int n = sd_.n + sa_.n;
    std_sort_replica(spans, n, SpanXLess());
    base[0] = (uint32_t)n;
    if (n && o.clip_in == 0) a.states[o.clip_out].nonempty = 1;
  } else {
    const ClipStateDesc par = a.states[o.clip_in];
    const uint2* ms = nullptr;
    int n_ms = 0;
    if (par.nonempty && par.rw > 0 && y >= par.ry0 && y < par.ry0 + par.rh) {
      const uint32_t* base = clip_state_row(a, o.clip_in, par, y);
      n_ms = (int)base[0];
      ms = reinterpret_cast<const uint2*>(base + 2);
    }
    if (mode == 2) {
      // an intersecting clip on top of a difference state: RecursiveClip's spans_subtraction(fresh spans, clip spans)
      // (sw_canvas.cc:186-187); the result is an intersecting state (:328-330), every span of the fresh path kept
      // (coverage 0 included), every piece an entry
      const ClipStateDesc own = a.states[o.clip_out];
      if (own.rw == 0 || y < own.ry0 || y >= own.ry0 + own.rh) return;
      DiffCutT<DiffTablePiece> cd, ca_;
      cd.piece.own_row = clip_state_row(a, o.clip_out, own, y);
      cd.piece.rx0 = own.rx0;
      cd.piece.rw = own.rw;
      cd.piece.cover = 0;
      cd.piece.wrote = cd.piece.over = false;
      cd.ms = ms;
      cd.n_ms = n_ms;
      cd.keep_zero = true;
      // the direct spans precede the accumulated ones in the list: two walks of the row, so that every pixel's entries
      // come out in list order (all pieces of direct spans, then all pieces of accumulated spans)
      struct Skip { __device__ void operator()(int, int, uint32_t) {} } skip;
      ClipRowState st2 = st;
      clip_row_spans(st, c.pool, row, x_first, x_last, cd, skip);
      ca_ = cd;
      clip_row_spans(st2, c.pool, row, x_first, x_last, skip, ca_);
      if (cd.piece.wrote || ca_.piece.wrote) a.states[o.clip_out].nonempty = 1;
      if (cd.piece.over || ca_.piece.over) *a.overflow = 1;
      return;
    }
    bool zm = false;
    if (c.zplane[1] != nullptr) {
      const skb_dl_paint pt = c.paints[o.paint];
      zm = blend_zero_src_matters(paint_blend_mode(pt)) || SKB_PAINT_CF_OFFSET(pt) != 0;
    }
    const uint32_t item_row = c.item_base[op] + (uint32_t)((y / SKB_TILE) - g.ty0) * (uint32_t)g.ntx;
    DiffCut cd, ca_;
    cd.piece = DiffPiece{c.mask[0], zm ? c.zplane[0] : nullptr, &g, item_row, y, (int)sd.w, 0u};
    cd.ms = ms;
    cd.n_ms = n_ms;
    cd.keep_zero = zm;
    ca_ = cd;
    ca_.piece.plane = c.mask[1];
    ca_.piece.zplane = zm ? c.zplane[1] : nullptr;
    clip_row_spans(st, c.pool, row, x_first, x_last, cd, ca_);
  }
}

// T2 (see ClipT2Rec): one thread per row of the parent's region.  The parent's spans are read back from its per-pixel
// table by a left-to-right sweep that keeps the spans open at the current pixel (an entry = span start + coverage; the
// same word in consecutive pixels is one span); when a span ends it is cut by the fresh path's row spans the way
// spans_subtraction does, and every piece is written to the SAME slot its span occupied in the parent's entry list of
// each pixel — the reference appends the pieces span by span in list order, so the order of a pixel's entries is the
// parent's.  Zero-length spans (marker entries) are cut on their own.  Rows the fresh path does not reach are copied.
// Two spans of the parent with the same start and coverage open at one pixel cannot be told apart in the table: flagged
// (the frame is refused).  A parent that turned out EMPTY makes the op an ordinary difference clip (sw_canvas.cc:331-334):
// the state becomes a difference state over the fresh path's region, whose sorted rows are already in place.
struct T2Piece {
  const uint32_t* par_row;
  uint32_t* own_row;
  int rx0, rw;
  uint32_t value;    // the parent's entry word of the span being cut
  bool wrote, lost;
  __device__ void put(int px, uint32_t e) {
    if (px < rx0 || px >= rx0 + rw) return;
    const uint32_t* ps = par_row + (size_t)(px - rx0) * SKB_CLIP_MAXE;
    for (int j = 0; j < SKB_CLIP_MAXE; j++) {
      if (ps[j] == value) {
        own_row[(size_t)(px - rx0) * SKB_CLIP_MAXE + j] = e;
        wrote = true;
        return;
      }
    }
    lost = true;
  }
  __device__ void operator()(int x, int len) {
    const uint32_t cover = clip_entry_cover(value);
    if (len == 0) {
      put(x, clip_entry(x, cover) | SKB_CLIP_MARKER);
      return;
    }
    for (int px = x; px < x + len; px++) put(px, clip_entry(x, cover));
  }
};

__global__ void __launch_bounds__(128) k_clip_t2(ClipArgs a, int level) {
  const ClipT2Rec rec = a.t2[blockIdx.y];
  if ((int)rec.depth != level) return;
  const CoverArgs& c = a.c;
  const skb_dl_op o = c.ops[rec.op];
  ClipStateDesc own = a.states[o.clip_out];
  const ClipStateDesc par = a.states[o.clip_in];
  const OpGeom g = c.geom[rec.op];
  int fx0, fy0, fw, fh;   // the fresh path's region: where its sorted row spans are
  clip_region_of(g, &fx0, &fy0, &fw, &fh);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, n_thr = gridDim.x * blockDim.x;
  if (par.kind == SKB_CLIP_KIND_DIFF) {   // difference on difference (PerformMerge): only reachable through an empty parent
    if (tid == 0) *a.overflow = 1;
    return;
  }
  if (!par.nonempty) {
    // HasClip() is false: the clip simply becomes the state, with its own op
    if (tid == 0) {
      ClipStateDesc d = own;
      d.kind = SKB_CLIP_KIND_DIFF;
      d.rx0 = fx0;
      d.ry0 = fy0;
      d.rw = fw;
      d.rh = fh;
      d.table_off = 0;
      bool any = false;
      for (int r = 0; r < fh && !any; r++) any = clip_state_row(a, o.clip_out, d, fy0 + r)[0] != 0;
      d.nonempty = any ? 1u : 0u;
      a.states[o.clip_out] = d;
    }
    return;
  }
  if (own.rw == 0) return;
  bool wrote = false, bad = false;
  for (uint32_t r = tid; r < (uint32_t)own.rh; r += n_thr) {
    const int y = own.ry0 + (int)r;
    const uint32_t* prow = clip_state_row(a, o.clip_in, par, y);   // same region as ours
    uint32_t* orow = clip_state_row(a, o.clip_out, own, y);
    const uint2* ms = nullptr;
    int n_ms = 0;
    if (fw > 0 && y >= fy0 && y < fy0 + fh) {
      ClipStateDesc f = own;
      f.rx0 = fx0;
      f.ry0 = fy0;
      f.rw = fw;
      f.rh = fh;
      f.table_off = 0;
      const uint32_t* base = clip_state_row(a, o.clip_out, f, y);
      n_ms = (int)base[0];
      ms = reinterpret_cast<const uint2*>(base + 2);
    }
    if (n_ms == 0) {   // "no spans in this line means minus zero": the row's spans stay as they are
      for (int k = 0; k < own.rw * SKB_CLIP_MAXE; k++) {
        const uint32_t v = prow[k];
        orow[k] = v;
        wrote |= v != 0;
      }
      continue;
    }
    T2Piece piece;
    piece.par_row = prow;
    piece.own_row = orow;
    piece.rx0 = own.rx0;
    piece.rw = own.rw;
    piece.value = 0;
    piece.wrote = piece.lost = false;
    uint32_t open_v[2 * SKB_CLIP_MAXE];   // spans open at the previous pixel
    int n_open = 0;
    for (int x = 0; x <= own.rw; x++) {
      uint32_t here[SKB_CLIP_MAXE];
      int n_here = 0;
      if (x < own.rw) {
        const uint32_t* ps = prow + (size_t)x * SKB_CLIP_MAXE;
        for (int j = 0; j < SKB_CLIP_MAXE && ps[j]; j++) {
          const uint32_t v = ps[j];
          if (clip_entry_is_marker(v)) {   // a zero-length span of the parent at this pixel
            piece.value = v;
            struct ZeroLen {
              T2Piece* p;
              int px;
              __device__ void operator()(int x_, int len) {
                if (len == 0) p->put(px, p->value);   // handed down as it is (start and coverage kept)
                // a zero-length span has no pixels: a piece with length cannot come out of it
              }
            } zl{&piece, own.rx0 + x};
            span_subtract(clip_entry_start(v), 0, ms, n_ms, zl);
            continue;
          }
          for (int k = 0; k < n_here; k++) bad |= here[k] == v;   // twins: indistinguishable in the table
          here[n_here++] = v;
        }
      }
      // spans that were open and are not here any more end at this pixel
      for (int k = 0; k < n_open; k++) {
        bool goes_on = false;
        for (int j = 0; j < n_here; j++) goes_on |= here[j] == open_v[k];
        if (goes_on) continue;
        const int start = clip_entry_start(open_v[k]);
        piece.value = open_v[k];
        span_subtract(start, own.rx0 + x - start, ms, n_ms, piece);
      }
      // a span whose start lies left of the pixel where it first shows (it began off the region) still counts from
      // its start: the open list is simply what is here now
      n_open = n_here;
      for (int j = 0; j < n_here; j++) open_v[j] = here[j];
    }
    // close the gaps the removed pixels left in the entry lists (order kept)
    for (int x = 0; x < own.rw; x++) {
      uint32_t* e = orow + (size_t)x * SKB_CLIP_MAXE;
      int k = 0;
      for (int j = 0; j < SKB_CLIP_MAXE; j++) {
        const uint32_t v = e[j];
        if (v) {
          e[j] = 0;
          e[k++] = v;
        }
      }
    }
    wrote |= piece.wrote;
    bad |= piece.lost;
  }
  if (wrote) a.states[o.clip_out].nonempty = 1;
  if (bad) *a.overflow = 1;
}

// One warp per (op, tile) item of a clipped draw: which planes are present, is plane 0 solid.
__global__ void __launch_bounds__(128) k_clip_classify(CoverArgs c) {
  const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (item >= c.n_items) return;
  const uint32_t op = find_interval(c.item_base, c.n_ops, item);
  const skb_dl_op o = c.ops[op];
  if (o.kind != SKB_OP_FILL || o.clip_in == 0) return;
  bool zm = false;
  if (c.zplane[1] != nullptr) {
    const skb_dl_paint pt = c.paints[o.paint];
    zm = blend_zero_src_matters(paint_blend_mode(pt)) || SKB_PAINT_CF_OFFSET(pt) != 0;
  }
  uint32_t flags = 0;
  bool solid = true;
#pragma unroll
  for (int k = 0; k < SKB_CLIP_PLANES; k++) {
    const uint2 v = reinterpret_cast<const uint2*>(c.mask[k] + (size_t)item * 256)[lane];
    bool nz = (v.x | v.y) != 0;
    if (zm) {
      const uint2 z = reinterpret_cast<const uint2*>(c.zplane[k] + (size_t)item * 256)[lane];
      nz |= (z.x | z.y) != 0;
    }
    if (__any_sync(0xffffffffu, nz)) flags |= 1u << k;
    if (k == 0) solid = (v.x == 0xFFFFFFFFu && v.y == 0xFFFFFFFFu);
  }
  solid = __all_sync(0xffffffffu, solid) && flags == 1u;
  if (solid) flags |= SKB_ITEM_SOLID;
  if (zm && flags) flags |= SKB_ITEM_ZERO;
  if (lane == 0) {
    c.item_flags[item] = (uint16_t)flags;
    if (flags) {
      const OpGeom g = c.geom[op];
      const uint32_t local = item - c.item_base[op];
      const int tx = g.tx0 + (int)(local % (uint32_t)g.ntx);
      const int ty = g.ty0 + (int)(local / (uint32_t)g.ntx);
      const SurfDesc sd = c.surfs[o.surface];
      atomicAdd(&c.tile_cnt[sd.tile_base + (uint32_t)ty * sd.tiles_x + (uint32_t)tx], (uint32_t)__popc(flags & SKB_ITEM_PLANE_MASK));
    }
  }
}

// --------------------------------------------------------------------- stage 5: bin
// One thread per (op, tile row) — the owner of a tile row comes from the table the walk list wrote, where a thread per
// item had to find its op by a binary search over a million item offsets — appending the row's non-empty items to their
// tiles' command lists.
__global__ void k_scatter(CoverArgs c, const uint32_t* tile_off, uint32_t* tile_fill, uint2* cmds) {
  const uint32_t trow = blockIdx.x * blockDim.x + threadIdx.x;
  if (trow >= c.n_trows) return;
  const uint32_t op = c.trow_op[trow];
  const skb_dl_op o = c.ops[op];
  if (o.kind != SKB_OP_FILL) return;
  const OpGeom g = c.geom[op];
  if (g.empty || g.ntx == 0 || g.nty == 0) return;
  const uint32_t tr = trow - c.row_base[op] / SKB_TILE;
  if (tr >= (uint32_t)g.nty) return;
  const SurfDesc sd = c.surfs[o.surface];
  const uint32_t item0 = c.item_base[op] + tr * (uint32_t)g.ntx;
  const uint32_t tile0 = sd.tile_base + (uint32_t)(g.ty0 + (int)tr) * sd.tiles_x + (uint32_t)g.tx0;
  for (int i = 0; i < g.ntx; i++) {
    const uint32_t item = item0 + (uint32_t)i;
    const uint32_t flags = c.item_flags[item];
    if (!flags) continue;
    const uint32_t tile = tile0 + (uint32_t)i;
    const uint32_t n = (uint32_t)__popc(flags & SKB_ITEM_PLANE_MASK);
    uint32_t pos = tile_off[tile] + atomicAdd(&tile_fill[tile], n);
    for (uint32_t k = 0; k < SKB_CLIP_PLANES; k++) {
      if (!((flags >> k) & 1u)) continue;
      cmds[pos++] = make_uint2((op << 3) | k, item | ((k == 0 && (flags & SKB_ITEM_SOLID)) ? SKB_CMD_SOLID : 0u) |
                                                  (((flags & SKB_ITEM_ZERO) && (k == 0 || o.clip_in != 0)) ? SKB_CMD_ZERO : 0u));
    }
  }
}

// -------------------------------------------------------------------- stage 6: fine
struct FineArgs {
  const uint32_t* tile_off;
  const uint2* cmds;
  uint2* cmds_sorted;  // scratch for tiles whose list does not fit the shared-memory sorter
  uint32_t tile_begin, tile_end;
  uint32_t level;  // which surfaces this launch composites
  uint32_t remote_store;  // the canvas band is stored into another GPU's canvas: empty tiles are stored as well
  const SurfDesc* surfs;
  const uint32_t* surf_tile_base;  // n_surfaces + 1
  uint32_t n_surfaces;
  const OpGeom* geom;
  const skb_dl_op* ops;
  const skb_dl_paint* paints;
  const float* stops;
  const uint8_t* mask[SKB_CLIP_PLANES];
  const uint8_t* zmask;
  const uint8_t* zplane[SKB_CLIP_PLANES];  // [0] = zmask
};
#ifndef FINE_WARPS
#define FINE_WARPS 2
#endif
#define FINE_SORT_CAP 256

#ifndef FINE_MINB
#define FINE_MINB 16  // 64 registers: measured C1 0.27 -> 0.235 ms, C3 7.6 -> 6.6 ms against the compiler's own choice
#endif
#ifdef FINE_MINB  // minimum resident blocks per SM (caps the registers); unset = the compiler's own choice
#define FINE_BOUNDS __launch_bounds__(FINE_WARPS * 32, FINE_MINB)
#else
#define FINE_BOUNDS __launch_bounds__(FINE_WARPS * 32)
#endif
__global__ void FINE_BOUNDS k_fine(FineArgs a) {
  __shared__ uint2 s_cmd[FINE_WARPS][FINE_SORT_CAP];
  __shared__ uint32_t s_src[FINE_WARPS][256];   // paint colours of a tile's covered pixels (gradient commands)
  __shared__ uint8_t s_pix[FINE_WARPS][256];    // which pixels those are
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* const s_requant = nullptr;  // the sampler's byte round trip is the identity (skb_core.cuh)
  const uint32_t tile = a.tile_begin + blockIdx.x * FINE_WARPS + warp;
  if (tile >= a.tile_end) return;
  const uint32_t c0 = a.tile_off[tile];
  const uint32_t n = a.tile_off[tile + 1] - c0;
  if (n == 0 && !a.remote_store) return;
  const uint32_t s = find_interval(a.surf_tile_base, a.n_surfaces, tile);
  const SurfDesc sd = a.surfs[s];
  if (sd.level != a.level) return;
  const uint32_t local = tile - sd.tile_base;
  const int tx = (int)(local % sd.tiles_x), ty = (int)(local / sd.tiles_x);
  const int y = ty * SKB_TILE + (lane >> 1);
  const int x0 = tx * SKB_TILE + (lane & 1) * 8;
  if (n == 0) {
    // The band's pixels are stored into another GPU's canvas (sd.px_out != sd.px): tiles nothing was drawn into are
    // stored too, so that the gathering canvas needs no clear of its own for the rows other devices own.
    if (sd.px_out != sd.px) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(sd.px + (size_t)y * sd.pitch) + x0);
      uint4* dstp = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(sd.px_out + (size_t)y * sd.pitch) + x0);
      dstp[0] = src[0];
      dstp[1] = src[1];
    }
    return;
  }

  // order the tile's commands by (op, plane): draws must be composited in draw order
  const uint2* list;
  if (n <= 32) {
    // one command per lane; its position is the number of smaller keys (keys are unique)
    const uint2 mine = lane < n ? a.cmds[c0 + lane] : make_uint2(0xFFFFFFFFu, 0u);
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) rank += __shfl_sync(0xffffffffu, mine.x, (int)j) < mine.x ? 1u : 0u;
    if (lane < n) s_cmd[warp][rank] = mine;
    __syncwarp();
    list = s_cmd[warp];
  } else if (n <= FINE_SORT_CAP) {
    uint32_t N = 64;
    while (N < n) N <<= 1;
    for (uint32_t i = lane; i < N; i += 32) s_cmd[warp][i] = i < n ? a.cmds[c0 + i] : make_uint2(0xFFFFFFFFu, 0u);
    __syncwarp();
    for (uint32_t k = 2; k <= N; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = lane; i < N; i += 32) {
          uint32_t ixj = i ^ j;
          if (ixj > i) {
            uint2 u = s_cmd[warp][i], v = s_cmd[warp][ixj];
            bool up = (i & k) == 0;
            if ((u.x > v.x) == up) {
              s_cmd[warp][i] = v;
              s_cmd[warp][ixj] = u;
            }
          }
        }
        __syncwarp();
      }
    }
    list = s_cmd[warp];
  } else {
    // rank sort through global scratch (keys are unique): rare, only for very deep tiles
    for (uint32_t i = lane; i < n; i += 32) {
      uint2 u = a.cmds[c0 + i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n; j++) rank += a.cmds[c0 + j].x < u.x;
      a.cmds_sorted[c0 + rank] = u;
    }
    __threadfence_block();
    __syncwarp();
    list = a.cmds_sorted + c0;
  }

  uint32_t* prow = reinterpret_cast<uint32_t*>(sd.px + (size_t)y * sd.pitch) + x0;
  uint4 q0 = reinterpret_cast<uint4*>(prow)[0];
  uint4 q1 = reinterpret_cast<uint4*>(prow)[1];
  // blended in the reference's register order (see swap_rb), swapped back when stored
  uint32_t dst[8] = {swap_rb(q0.x), swap_rb(q0.y), swap_rb(q0.z), swap_rb(q0.w),
                     swap_rb(q1.x), swap_rb(q1.y), swap_rb(q1.z), swap_rb(q1.w)};

  // The loops over a lane's eight pixels that stay rolled (the paints with a lot of code per pixel) work on dst[0] and
  // rotate the array once per pixel: every index is a constant, so dst[] lives in registers — indexed by the loop
  // counter it sat in local memory, loaded and stored around every command, the branch-free solid path included.
#define SKB_ROTATE_DST(d)                                                                                       \
  do {                                                                                                          \
    dst[0] = dst[1]; dst[1] = dst[2]; dst[2] = dst[3]; dst[3] = dst[4];                                         \
    dst[4] = dst[5]; dst[5] = dst[6]; dst[6] = dst[7]; dst[7] = (d);                                            \
  } while (0)
  // A command that covers the whole tile with an opaque solid colour (SrcOver) leaves nothing of what was blended
  // before it: start at the last such command.
  uint32_t first = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + (uint32_t)lane;
    bool opaque = false;
    if (i < n) {
      const uint2 cmd = list[i];
      if (cmd.y & SKB_CMD_SOLID) {
        const uint2 gc = *reinterpret_cast<const uint2*>(&a.geom[cmd.x >> 3].color);  // color, fast_solid
        opaque = gc.y != 0 && (gc.x >> 24) == 255u;
      }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, opaque);
    if (b) first = base + 31u - (uint32_t)__clz((int)b);
  }
  for (uint32_t i = first; i < n; i++) {
    const uint2 cmd = list[i];
    const uint32_t op = cmd.x >> 3;
    uint32_t lo, hi;
    if (cmd.y & SKB_CMD_SOLID) {
      lo = hi = 0xFFFFFFFFu;
    } else {
      const uint8_t* m = a.mask[cmd.x & 7u] + (size_t)(cmd.y & SKB_CMD_ITEM_MASK) * 256;
      uint2 mv = reinterpret_cast<const uint2*>(m)[lane];
      lo = mv.x;
      hi = mv.y;
    }
    const uint2 gc = *reinterpret_cast<const uint2*>(&a.geom[op].color);  // color, fast_solid
    if (gc.y) {
      // Solid colour, SrcOver: blend_cover() without its branches.  Its three shortcuts are what the general
      // expression gives anyway: coverage 255 leaves the colour as it is (= AlphaMulQ by 256), a source alpha of 0
      // adds 0 to AlphaMulQ(dst, 256) = dst (a premultiplied colour scaled to alpha 0 is 0), and a source alpha of
      // 255 adds the colour to AlphaMulQ(dst, 1) = 0.  All lanes of the warp run the same instructions.
      const uint32_t color = gc.x;
      if (cmd.y & SKB_CMD_SOLID) {
        const uint32_t inv = 256u - (color >> 24);
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = color + alpha_mul_q(dst[j], inv);
      } else {
        const uint32_t c_rb = color & 0x00FF00FFu, c_ag = (color >> 8) & 0x00FF00FFu;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint32_t cv = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF;
          const uint32_t scale = cv + ((cv + 1u) >> 8);  // 255 -> 256
          const uint32_t src = (((c_rb * scale) >> 8) & 0x00FF00FFu) | ((c_ag * scale) & 0xFF00FF00u);
          dst[j] = src + alpha_mul_q(dst[j], 256u - (src >> 24));
        }
      }
    } else {
      uint32_t zlo = 0, zhi = 0;
      if (cmd.y & SKB_CMD_ZERO) {
        uint2 zv = reinterpret_cast<const uint2*>(a.zplane[cmd.x & 7u] + (size_t)(cmd.y & SKB_CMD_ITEM_MASK) * 256)[lane];
        zlo = zv.x;
        zhi = zv.y;
      }
      const uint32_t pidx = a.ops[op].paint;
      const uint32_t ptype = a.paints[pidx].type;
      // gradients (and bilinear images) are evaluated for the tile's covered pixels dealt evenly to the lanes; the
      // per-lane shortcut below would leave that warp-wide step
      const bool dealt = (ptype >= SKB_PAINT_LINEAR && ptype <= SKB_PAINT_SWEEP) || ptype == SKB_PAINT_CONICAL;
      if (!dealt && (lo | hi | zlo | zhi) == 0) continue;
      const skb_dl_paint pt = a.paints[pidx];
      const uint32_t galpha = ptype == SKB_PAINT_IMAGE ? (pt.global_alpha & 0xFF) : 0xFFu;
      SurfaceView img;
      img.px = nullptr;
      img.w = img.h = img.pitch = 0;
      if (ptype == SKB_PAINT_IMAGE) {
        const SurfDesc is = a.surfs[pt.image_surface];
        img.px = is.px;
        img.w = is.w;
        img.h = is.h;
        img.pitch = is.pitch;
      }
      const uint32_t mode = paint_blend_mode(pt);
      // a colour filter can turn a zero source into something: then zero-coverage pixels matter as well
      const uint32_t* cf = SKB_PAINT_CF_OFFSET(pt) ? reinterpret_cast<const uint32_t*>(a.stops) + (SKB_PAINT_CF_OFFSET(pt) - 1) : nullptr;
      const bool zmode = blend_zero_src_matters(mode) || cf != nullptr;
      if (ptype == SKB_PAINT_IMAGE && !(pt.tile_mode & SKB_PAINT_IMAGE_LINEAR)) {
        // Nearest-sampled image (blurred temporaries, layers, DrawImage): the same per-pixel work as the general
        // loop below with what is constant over the tile kept in registers — paint_color()'s matrix row for this
        // pixel row, the tile modes, the blend mode (SrcOver inline).
        const float fyc = y + 0.5f;
        const float uy = fyc * pt.m[1], vy = fyc * pt.m[4];
        const float m0 = pt.m[0], m2 = pt.m[2], m3 = pt.m[3], m5 = pt.m[5];
        const uint32_t tmode = pt.tile_mode;
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
          uint32_t cv = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF;
          const bool touched = cv != 0 || (((j < 4 ? zlo : zhi) >> (8 * (j & 3))) & 0xFF) != 0;
          cv &= galpha;
          uint32_t d = dst[0];
          if (cv || (touched && zmode)) {
            const float fxc = (x0 + j) + 0.5f;
            const float u = fxc * m0 + uy + m2;  // paint_color: fxc * m[0] + fyc * m[1] + m[2], same order
            const float v = fxc * m3 + vy + m5;
            uint32_t src = swap_rb(sample_image_nearest(tmode, img, u, v, s_requant));
            if (cv != 255) src = alpha_mul_q(src, cv);
            if (cf) src = apply_color_filter(cf, src);
            if (mode == SKB_BLEND_SRC_OVER) d = (src >> 24) == 0 ? d : src + alpha_mul_q(d, 256 - (src >> 24));
            else d = porter_duff(src, d, mode);
          }
          SKB_ROTATE_DST(d);
        }
        continue;
      }
      if (dealt) {
        // Gradient paint: paint_color() is by far the most expensive step (fp32 with IEEE division / square root, the
        // fdlibm atan2f of the sweep gradient) and a mask covers only part of a tile — evaluated by the owner of each
        // pixel, half of the lanes wait for the other half.  Instead the covered pixels of the tile are listed (ballot-
        // free prefix over the lanes' counts) and dealt out 32 at a time: every lane computes the colour of ONE covered
        // pixel per round, into shared memory; then each lane scales / filters / blends its own eight from there.
        uint32_t need = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint32_t cvj = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF;
          const bool touched = cvj != 0 || (((j < 4 ? zlo : zhi) >> (8 * (j & 3))) & 0xFF) != 0;
          if ((cvj & galpha) || (touched && zmode)) need |= 1u << j;
        }
        const int cnt = __popc(need);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;   // uniform over the warp
        {
          int k = incl - cnt;
          for (uint32_t m = need; m; m &= m - 1) s_pix[warp][k++] = (uint8_t)(lane * 8 + (__ffs((int)m) - 1));
        }
        __syncwarp();
        for (int b = 0; b < total; b += 32) {
          const int idx = b + lane;
          if (idx < total) {
            const int pid = s_pix[warp][idx];
            const int ln = pid >> 3;
            const int px_ = tx * SKB_TILE + (ln & 1) * 8 + (pid & 7), py_ = ty * SKB_TILE + (ln >> 1);
            s_src[warp][pid] = paint_color(pt, a.stops, img, px_, py_, s_requant);
          }
        }
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
          uint32_t d = dst[0];
          if ((need >> j) & 1u) {
            const uint32_t cv = (((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF) & galpha;
            uint32_t src = swap_rb(s_src[warp][lane * 8 + j]);
            if (cv != 255) src = alpha_mul_q(src, cv);
            if (cf) src = apply_color_filter(cf, src);
            d = porter_duff(src, d, mode);
          }
          SKB_ROTATE_DST(d);
        }
        __syncwarp();   // the next command reuses the lists
        continue;
      }
#pragma unroll 1
      for (int j = 0; j < 8; j++) {
        uint32_t cv = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF;
        // a span reaches the pixel when its coverage is non-zero, or zero on a direct span (zmask)
        const bool touched = cv != 0 || (((j < 4 ? zlo : zhi) >> (8 * (j & 3))) & 0xFF) != 0;
        cv &= galpha;  // `cover & global_alpha_` (sw_span_brush.cc:101)
        uint32_t d = dst[0];
        if (cv || (touched && zmode)) {
          // BrushH: colour, scaled by the coverage, colour filter, blend (sw_span_brush.cc:108-133)
          uint32_t src = swap_rb(paint_color(pt, a.stops, img, x0 + j, y, s_requant));
          if (cv != 255) src = alpha_mul_q(src, cv);
          if (cf) src = apply_color_filter(cf, src);
          d = porter_duff(src, d, mode);
        }
        SKB_ROTATE_DST(d);
      }
    }
  }
  uint32_t* orow = reinterpret_cast<uint32_t*>(sd.px_out + (size_t)y * sd.pitch) + x0;
  reinterpret_cast<uint4*>(orow)[0] = make_uint4(swap_rb(dst[0]), swap_rb(dst[1]), swap_rb(dst[2]), swap_rb(dst[3]));
  reinterpret_cast<uint4*>(orow)[1] = make_uint4(swap_rb(dst[4]), swap_rb(dst[5]), swap_rb(dst[6]), swap_rb(dst[7]));
}

// -------------------------------------------------------------------- stage 7: blur
// SWStackBlur (sw_stack_blur.cc:18-284) as the triangular filter it computes:
//   sum(x) = SUM_{i=-r..r} (r+1-|i|) * p[clamp(x+i, 0, n-1)],   out = uint8((sum * mul[r]) >> shr[r])
// With m = r+1, T = prefix sum of the clamp-extended row and U = prefix sum of T:
//   sum(x) = U[x+m+1] - 2*U[x+1] + U[x-m+1]      (box_m * box_m = triangle)
// All sums are kept mod 2^32 — the result (< 2^24) is exact.
struct BlurJob {
  uint32_t src, dst;  // surface ids
  int32_t radius;
  uint32_t style;     // 0 plain blur; BlurStyle kSolid 2 / kOuter 3 / kInner 4 (mask_filter.cc:64-100); 5 drop shadow;
                      // 6 dilate, 7 erode
  uint32_t color;     // style 5: the shadow colour (unpremultiplied A<<24|R<<16|G<<8|B)
  float morph_rx, morph_ry;  // styles 6 / 7 (dilate / erode): the filter's radii, no blur
};

// What MaskFilterOnFilter / DropShadowImageFilter::OnFilter do to the blurred bitmap with the unblurred
// one at hand (src/effect/mask_filter.cc:64-100, src/effect/image_filter.cc:222-233), one thread per pixel.
__global__ void k_blur_style(SurfDesc raw, SurfDesc dst, uint32_t style, uint32_t color) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)dst.w * dst.h) return;
  const uint32_t x = (uint32_t)(i % dst.w), y = (uint32_t)(i / dst.w);
  uint32_t* out = reinterpret_cast<uint32_t*>(dst.px + (size_t)y * dst.pitch) + x;
  const uint32_t raw_c = reinterpret_cast<const uint32_t*>(raw.px + (size_t)y * raw.pitch)[x];
  const uint32_t blur_c = *out;
  const uint32_t raw_a = raw_c >> 24, blur_a = blur_c >> 24;
  if (style == 2) {  // kSolid: the shape itself over its blur
    if (raw_a > 0) *out = raw_c;
  } else if (style == 3) {  // kOuter: only the halo
    if (raw_a > 0 && raw_a >= blur_a) *out = 0;
  } else if (style == 4) {  // kInner: the blur inside the shape
    const float a_factor = 1.f / 255.f;
    *out = raw_a > 0 ? alpha_mul_q(blur_c, (uint32_t)(a_factor * raw_a * blur_a)) : 0u;
  } else if (style == 5) {  // drop shadow: the shadow colour (kept unpremultiplied) with the blurred alpha
    *out = blur_a > 0 ? (((color >> 16) & 0xFF) | (((color >> 8) & 0xFF) << 8) | ((color & 0xFF) << 16) | (blur_a << 24)) : 0u;
  }
}


// morph<type, direction> (src/effect/image_filter.cc:294-340): per channel max (dilate) or min (erode) over the window
// [i - radius, i + radius] clamped to the line, along x (dir 0) or y (dir 1).  One thread per pixel.
__global__ void k_morph(const uint8_t* src, uint32_t src_pitch, uint8_t* dst, uint32_t dst_pitch, int w, int h, int radius,
                        int dir, int erode) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)w * h) return;
  const int x = (int)(i % (uint32_t)w), y = (int)(i / (uint32_t)w);
  const int n = dir == 0 ? w : h, at = dir == 0 ? x : y;
  radius = min(radius, n - 1);
  const int lo = max(at - radius, 0), hi = min(at + radius, n - 1);
  uint32_t acc = erode ? 0xFFFFFFFFu : 0u;
  for (int k = lo; k <= hi; k++) {
    const uint32_t p = *reinterpret_cast<const uint32_t*>(src + (size_t)(dir == 0 ? y : k) * src_pitch + (size_t)(dir == 0 ? k : x) * 4);
    acc = erode ? __vminu4(acc, p) : __vmaxu4(acc, p);
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)y * dst_pitch + (size_t)x * 4) = acc;
}

// Horizontal pass: one BLOCK of BLUR_H_WARPS warps per (job, row).  The source row comes into shared memory with one
// bulk copy (TMA, cp.async.bulk + mbarrier).  Each thread owns a run of consecutive samples of the extended row
// (n + 2m + 1 samples): it sums them (T), the thread totals are scanned (shuffles, then across the warps), it then
// builds its part of U (prefix of T) in shared memory, the totals are scanned again; the output loop reads U at the
// three positions.  Measured on C3 (2k paths, 8192^2; the H pass alone): runs read from global memory, even run
// length 5.9 ms; odd run length (no bank conflicts on U) 5.0 ms; row staged by TMA 4.7 ms; two warps per row instead
// of one (the shared memory a row needs limits the rows in flight, so more threads per row) see DESIGN.md.
#ifndef BLUR_H_WARPS
#define BLUR_H_WARPS 2
#endif
#ifndef BLUR_H_TMA
#define BLUR_H_TMA 1
#endif
// exclusive scan of one value per thread over the block's threads (in thread order), four channels at once
__device__ __forceinline__ void blur_row_scan(const uint32_t v[4], uint32_t off[4], int lane, int wib, uint32_t (*tot)[4]) {
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint32_t inc = v[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    off[c] = inc - v[c];
    if (BLUR_H_WARPS > 1 && lane == 31) tot[wib][c] = inc;
  }
  if (BLUR_H_WARPS > 1) {
    __syncthreads();
#pragma unroll
    for (int w = 0; w < BLUR_H_WARPS - 1; w++) {
      if (w < wib) {
#pragma unroll
        for (int c = 0; c < 4; c++) off[c] += tot[w][c];
      }
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(BLUR_H_WARPS * 32) k_blur_h(const BlurJob* jobs, const uint32_t* job_row_base, uint32_t n_jobs,
                                                              const SurfDesc* surfs, uint32_t first_row, uint32_t end_row,
                                                              uint32_t max_len) {
  // U[max_len] (uint4), the source row (max_len words), the warps' scan totals, one mbarrier
  extern __shared__ __align__(16) uint32_t sm[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NT = BLUR_H_WARPS * 32;
  const uint32_t grow = first_row + blockIdx.x;
  if (grow >= end_row) return;
  const uint32_t job = find_interval(job_row_base, n_jobs, grow);
  const BlurJob jb = jobs[job];
  const SurfDesc S = surfs[jb.src], D = surfs[jb.dst];
  const int y = (int)(grow - job_row_base[job]);
  const int n = (int)S.w;
  const uint32_t* srow = reinterpret_cast<const uint32_t*>(S.px + (size_t)y * S.pitch);
  uint32_t* drow = reinterpret_cast<uint32_t*>(D.px + (size_t)y * D.pitch);
  if (jb.style >= 6) return;  // dilate / erode: k_morph
  int r = jb.radius > 254 ? 254 : jb.radius;
  if (r <= 1) {
    for (int x = tid; x < n; x += NT) drow[x] = srow[x];
    return;
  }
  const int m = r + 1;
  const int len = n + 2 * m + 1;  // extended index j = k - m for k in [0, len): j in [-m, n+m]
  uint4* U = reinterpret_cast<uint4*>(sm);  // U[k] = the four channels' exclusive double prefix
  uint32_t* rowbuf = reinterpret_cast<uint32_t*>(U + max_len);
  uint32_t(*tot)[4] = reinterpret_cast<uint32_t(*)[4]>(rowbuf + max_len);
  uint64_t* bar = reinterpret_cast<uint64_t*>(tot + BLUR_H_WARPS);
#if BLUR_H_TMA
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  if (tid == 0) {
    const uint32_t bytes = ((uint32_t)n * 4u + 15u) & ~15u;  // the pitch is padded to whole tiles: reading on is safe
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(rowbuf)),
                 "l"(srow), "r"(bytes), "r"(bar_a)
                 : "memory");
  }
  if (BLUR_H_WARPS > 1) __syncthreads();  // the barrier word is initialised before anyone waits on it
  else __syncwarp();
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "BLUR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra BLUR_DONE;\n"
      "bra BLUR_WAIT;\n"
      "BLUR_DONE:\n"
      "}" ::"r"(bar_a)
      : "memory");
  const uint32_t* rsrc = rowbuf;
#else
  const uint32_t* rsrc = srow;
#endif
  // an odd run length keeps the threads' 128-bit accesses to U (and their reads of the staged row) on different banks
  const int chunk = ((len + NT - 1) / NT) | 1;
  const int k0 = min(tid * chunk, len), k1 = min(k0 + chunk, len);
  auto sample = [&](int k) -> uint32_t {
    int j = k - m;
    j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    return rsrc[j];
  };
  // One pass over my run builds T and U relative to the run's start (T_rel = sum of the run's samples so far, U_rel =
  // sum of T_rel so far) and leaves U_rel in shared memory; the run totals are scanned over the row; with t_off = T at
  // the run's start the true values are T = t_off + T_rel and U[k] = u_off + (k - k0) t_off + U_rel[k], where u_off
  // is the scan of the runs' sums of T = n_run t_off + U_rel(end).
  uint32_t tv[4] = {0, 0, 0, 0}, ua[4] = {0, 0, 0, 0};
  for (int k = k0; k < k1; k++) {
    const uint32_t px = sample(k);
    U[k] = make_uint4(ua[0], ua[1], ua[2], ua[3]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      ua[c] += tv[c];                       // U_rel[k+1] = U_rel[k] + T_rel[k]
      tv[c] += (px >> (8 * c)) & 0xFF;      // T_rel[k+1] = T_rel[k] + e[k]
    }
  }
  uint32_t t_off[4];
  blur_row_scan(tv, t_off, lane, wib, tot);
  const uint32_t n_run = (uint32_t)(k1 - k0);
#pragma unroll
  for (int c = 0; c < 4; c++) ua[c] += n_run * t_off[c];
  uint32_t u_off[4];
  blur_row_scan(ua, u_off, lane, wib, tot);
  for (int k = k0; k < k1; k++) {
    uint4 v = U[k];
    U[k] = make_uint4(v.x + u_off[0], v.y + u_off[1], v.z + u_off[2], v.w + u_off[3]);
#pragma unroll
    for (int c = 0; c < 4; c++) u_off[c] += t_off[c];   // (k - k0) t_off, one step at a time
  }
  if (BLUR_H_WARPS > 1) __syncthreads();
  else __syncwarp();
  uint32_t mul;
  int shr;
  blur_mul_shr(r, &mul, &shr);
  for (int x = tid; x < n; x += NT) {
    // extended index j maps to k = j + m;  U[j] here means prefix over samples < j
    const uint4 ul = U[x + m + 1 + m], um = U[x + 1 + m], ug = U[x - m + 1 + m];
    const uint32_t s0 = ul.x - 2u * um.x + ug.x, s1 = ul.y - 2u * um.y + ug.y;
    const uint32_t s2 = ul.z - 2u * um.z + ug.z, s3 = ul.w - 2u * um.w + ug.w;
    drow[x] = (uint32_t)(uint8_t)(((uint64_t)s0 * mul) >> shr) | ((uint32_t)(uint8_t)(((uint64_t)s1 * mul) >> shr) << 8) |
              ((uint32_t)(uint8_t)(((uint64_t)s2 * mul) >> shr) << 16) | ((uint32_t)(uint8_t)(((uint64_t)s3 * mul) >> shr) << 24);
  }
}

// Vertical pass, in place on the destination: one thread per column, three running (T, U) pairs
// at rows y+m+1, y+1, y-m+1.  Reproduces the reference's seeding of out_sum.b/.a from the G channel
// (sw_stack_blur.cc:178-179): bytes 0 and 3 drift by -y*(r+1)*(g0 - b0|a0) in 64-bit arithmetic.
// Reads of the column lead/lag the writes by >= 1 row only through values already consumed, so the
// pass needs a separate source: `tmp` holds the H-blurred rows (the job's dst is written last).
__global__ void __launch_bounds__(128) k_blur_v(const BlurJob* jobs, const uint32_t* job_col_base, uint32_t n_jobs,
                                                const SurfDesc* surfs, const uint8_t* const* tmp_px, uint32_t first_col,
                                                uint32_t end_col) {
  const uint32_t gcol = first_col + blockIdx.x * blockDim.x + threadIdx.x;
  if (gcol >= end_col) return;
  const uint32_t job = find_interval(job_col_base, n_jobs, gcol);
  const BlurJob jb = jobs[job];
  const SurfDesc D = surfs[jb.dst];
  if (jb.style >= 6) return;
  int r = jb.radius > 254 ? 254 : jb.radius;
  if (r <= 1) return;  // the horizontal kernel already copied
  const int x = (int)(gcol - job_col_base[job]);
  const int h = (int)D.h;
  const int m = r + 1;
  const uint8_t* src = tmp_px[job];
  const size_t pitch = D.pitch;
  auto sample = [&](int j) -> uint32_t {
    j = j < 0 ? 0 : (j > h - 1 ? h - 1 : j);
    return *reinterpret_cast<const uint32_t*>(src + (size_t)j * pitch + (size_t)x * 4);
  };
  // running prefix pairs: P(k) = (T[k], U[k]) with T[k] = sum_{j<k} e[j], U[k] = sum_{j<k} T[j], extended index from -m
  uint32_t Tl[4] = {0, 0, 0, 0}, Ul[4] = {0, 0, 0, 0};  // at index y+m+1
  uint32_t Tm[4] = {0, 0, 0, 0}, Um[4] = {0, 0, 0, 0};  // at index y+1
  uint32_t Tg[4] = {0, 0, 0, 0}, Ug[4] = {0, 0, 0, 0};  // at index y-m+1
  auto advance = [&](uint32_t* Tt, uint32_t* Uu, int j) {
    uint32_t px = sample(j);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      Uu[c] += Tt[c];
      Tt[c] += (px >> (8 * c)) & 0xFF;
    }
  };
  // all three start at extended index -m (T = U = 0) and are brought to their y = 0 positions: m+1, 1 and -m+1.
  // The samples above the top row are the top row (clamp), so after k of them T = k p0 and U = p0 k (k-1) / 2: only
  // the leading pair has real rows (0..m) to walk through.
  const uint32_t p0 = sample(0);
  const uint32_t um = (uint32_t)m, tri_m = um * (um - 1u) / 2u, tri_m1 = (um + 1u) * um / 2u;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const uint32_t e0 = (p0 >> (8 * c)) & 0xFF;
    Tl[c] = um * e0;
    Ul[c] = e0 * tri_m;
    Tm[c] = (um + 1u) * e0;
    Um[c] = e0 * tri_m1;
    Tg[c] = e0;
    Ug[c] = 0u;
  }
  for (int j = 0; j < m + 1; j++) advance(Tl, Ul, j);       // -> index m+1
  const uint64_t g0 = (p0 >> 8) & 0xFF, b0 = p0 & 0xFF, a0 = (p0 >> 24) & 0xFF;
  const uint64_t drift0 = (uint64_t)m * (g0 - b0), drift3 = (uint64_t)m * (g0 - a0);
  uint32_t mul;
  int shr;
  blur_mul_shr(r, &mul, &shr);
  uint8_t* dcol = D.px + (size_t)x * 4;
  // From here on only S = Ul - 2 Um + Ug and its first difference Dd = Tl - 2 Tm + Tg are carried (U' = U + T and
  // T' = T + e give S' = S + Dd and Dd' = Dd + e[y+m+1] - 2 e[y+1] + e[y-m+1], all mod 2^32): 8 running values and
  // 7 operations per channel and row instead of 24 and 12.
  uint32_t S[4], Dd[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    S[c] = Ul[c] - 2u * Um[c] + Ug[c];
    Dd[c] = Tl[c] - 2u * Tm[c] + Tg[c];
  }
  for (int y = 0; y < h; y++) {
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint64_t sum = (uint64_t)S[c];
      if (c == 0) sum -= (uint64_t)y * drift0;
      if (c == 3) sum -= (uint64_t)y * drift3;
      out |= (uint32_t)(uint8_t)((sum * (uint64_t)mul) >> shr) << (8 * c);
    }
    *reinterpret_cast<uint32_t*>(dcol + (size_t)y * pitch) = out;
    const uint32_t pl = sample(y + m + 1), pm = sample(y + 1), pg = sample(y - m + 1);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      S[c] += Dd[c];
      Dd[c] += ((pl >> (8 * c)) & 0xFF) - 2u * ((pm >> (8 * c)) & 0xFF) + ((pg >> (8 * c)) & 0xFF);
    }
  }
}

// ----------------------------------------------------------------------- debug tap
__global__ void k_read_coverage(CoverArgs c, uint32_t op, int x, int y, uint32_t w, uint32_t h, uint8_t* direct, uint8_t* accum) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  int px = x + (int)(i % w), py = y + (int)(i / w);
  uint8_t dv = 0, av = 0;
  const OpGeom g = c.geom[op];
  if (!g.empty && g.ntx > 0 && px >= 0 && py >= 0) {
    int tx = px / SKB_TILE, ty = py / SKB_TILE;
    if (tx >= g.tx0 && tx < g.tx0 + g.ntx && ty >= g.ty0 && ty < g.ty0 + g.nty) {
      uint32_t item = c.item_base[op] + (uint32_t)((ty - g.ty0) * g.ntx + (tx - g.tx0));
      uint32_t f = c.item_flags[item];
      int lx = px % SKB_TILE, ly = py % SKB_TILE;
      if (f & SKB_ITEM_SOLID) {
        dv = 255;
      } else if (f & SKB_ITEM_PLANE0) {
        // masks do not say which span kind produced a lone value; report it as `direct`
        dv = c.mask0[(size_t)item * 256 + ly * 16 + lx];
        if (f & SKB_ITEM_PLANE1) av = c.mask1[(size_t)item * 256 + ly * 16 + lx];
        if (dv == 0 && av != 0) {  // report the blend SEQUENCE: zero coverage is never blended
          dv = av;
          av = 0;
        }
      }
    }
  }
  direct[i] = dv;
  accum[i] = av;
}

// =========================================================================== host side
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace skb

using namespace skb;

struct skb_device_s {
  int ordinal = 0;
  int sm_count = 0;
};

// Structure of an encoded frame, worked out by the ONE pass validate_dl makes over the op table, so that neither
// skb_frame_encode nor run_frame loops over the ops again (at 1M ops every such loop is ~10 ms of host time per frame, and
// inside run_frame it is idle GPU time).
struct FramePlan {
  bool clip_ops = false, clipped_fills = false, diff_clips = false;
  bool zero_blend = false;               // some draw's paint blends with a mode (or colour filter) that acts on zero-coverage pixels
  int max_depth = 0;
  std::vector<uint8_t> op_depth;         // nesting depth of the clip state a CLIP op defines (empty without clip ops)
  std::vector<skb_dl_op> blur_ops;       // the BLUR ops, in op order
  std::vector<ClipT2Rec> t2;             // difference clips applied on top of a path clip, in op order
  std::vector<uint32_t> surf_level;      // per surface: dependency depth (see SurfDesc::level)
  uint32_t max_level = 0;
  std::vector<uint8_t> surf_drawn;       // per surface: some FILL op targets it
};

struct skb_surface_s {
  skb_device dev = nullptr;
  uint32_t w = 0, h = 0;
  uint32_t band_y0 = 0, band_y1 = 0;
  int coord_mode = SKB_COORD_AUTO;
  FramePlan plan;   // structure of the encoded frame (validate_dl)
  uint8_t* stage[2] = {nullptr, nullptr};   // page-locked staging for reads into pageable memory (skb_surface_read_pixels)
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  int walk_mode = 0;
  int coverage_mode = SKB_COVERAGE_EXACT;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[12] = {};
  // host-mapped words the device writes counts into: reading them does not queue behind another
  // surface's read-back on the copy engine
  uint32_t* mapped_host = nullptr;
  uint32_t* mapped_dev = nullptr;
  // canvas pixels (persistent)
  uint8_t* canvas = nullptr;
  uint8_t* remote_canvas = nullptr;  // peer mapping of the same canvas on the gathering device (cudaIpcOpenMemHandle)
  uint32_t pitch = 0, tiles_x = 0, tiles_y = 0;
  // frame
  std::vector<uint8_t> host_dl;
  bool have_frame = false;
  cudaEvent_t ev_blur[2 * 16] = {};  // around the blur section of every level
  uint32_t n_levels_timed = 0;
  bool flushed = false;
  skb_frame_stats stats = {};
  // device buffers (grow-only)
  Buf area_line_cnt, area_item_cnt, area_item_cursor, area_item_local, area_item_delta, area_row_backdrop, area_lines;
  Buf rw_chord_cnt, rw_slot_op, rw_slots, rw_rank, rw_ops, rw_wrow_cnt, rw_rec_cnt, rw_chords, rw_wgrp_op, rw_ev, rw_tab, rw_res, rw_rec_off;
  Buf dl, geom, seg_op, prim_cnt, edges, edges2, quads, walk_lists, mask_extra[SKB_CLIP_PLANES - 2], clip_states, clip_px, clip_table, clip_t2, op_depth, ord, row_cnt, item_cnt, rows, trow_op, pool, counters, mask0, mask1, zmask, zplane_extra[SKB_CLIP_PLANES - 1], item_flags, tile_cnt,
      tile_fill, cmds, cmds_sorted, surfs, surf_tile_base, temp_px, scan_tmp, blur_jobs, blur_rows, blur_cols, blur_tmp_ptrs,
      blur_tmp;
  // host mirrors kept for the debug tap
  uint32_t n_ops = 0, n_items = 0;
  std::vector<SurfDesc> h_surfs;
};

namespace skb {

__global__ void k_fetch_words(uint32_t* dst, const uint32_t* src, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

__global__ void k_gather5(uint32_t* dst, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, const uint32_t* e) {
  if (threadIdx.x == 0) {
    dst[0] = *a;
    dst[1] = *b;
    dst[2] = *c;
    dst[3] = d ? *d : 0u;
    dst[4] = e ? *e : 0u;
  }
}

// Brings n (<= 16) device words to the host and waits for them.
static skb_result fetch_words(skb_surface s, uint32_t* out, const uint32_t* dev_src, int n) {
  k_fetch_words<<<1, 32, 0, s->stream>>>(s->mapped_dev, dev_src, n);
  SKB_CUDA(cudaGetLastError());
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < n; i++) out[i] = ((volatile uint32_t*)s->mapped_host)[i];
  return SKB_SUCCESS;
}

static skb_result buf_reserve(Buf& b, size_t bytes) {
  if (bytes <= b.cap) return SKB_SUCCESS;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    return SKB_ERROR_OUT_OF_MEMORY;
  }
  b.cap = want;
  return SKB_SUCCESS;
}
static void buf_free(Buf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

#define SKB_TRY(expr)                     \
  do {                                    \
    skb_result _r = (expr);               \
    if (_r != SKB_SUCCESS) return _r;     \
  } while (0)

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// exclusive scan of data[0..n] (n+1 entries, data[n] must be 0 on entry) -> data[n] = total
static skb_result scan_exclusive(skb_surface s, uint32_t* data, uint32_t n_plus_1, uint32_t* launches) {
  // levels of block sums live in scan_tmp
  std::vector<uint32_t> sizes;
  uint32_t n = n_plus_1;
  size_t total = 0;
  while (n > SCAN_BLOCK) {
    n = cdiv(n, SCAN_BLOCK);
    sizes.push_back(n);
    total += n;
  }
  SKB_TRY(buf_reserve(s->scan_tmp, (total + 1) * 4));
  std::vector<uint32_t*> lv;
  uint32_t* base = (uint32_t*)s->scan_tmp.p;
  for (uint32_t sz : sizes) {
    lv.push_back(base);
    base += sz;
  }
  // up-sweep
  uint32_t* cur = data;
  n = n_plus_1;
  for (size_t l = 0; l <= sizes.size(); l++) {
    uint32_t blocks = cdiv(n, SCAN_BLOCK);
    k_scan_block<<<blocks, SCAN_THREADS, 0, s->stream>>>(cur, n, l < sizes.size() ? lv[l] : nullptr);
    (*launches)++;
    if (l < sizes.size()) {
      cur = lv[l];
      n = sizes[l];
    }
  }
  // down-sweep
  for (size_t l = sizes.size(); l-- > 0;) {
    uint32_t* child = l == 0 ? data : lv[l - 1];
    uint32_t child_n = l == 0 ? n_plus_1 : sizes[l - 1];
    k_scan_add<<<cdiv(child_n, SCAN_BLOCK), SCAN_THREADS, 0, s->stream>>>(child, child_n, lv[l]);
    (*launches)++;
  }
  SKB_CUDA(cudaGetLastError());
  return SKB_SUCCESS;
}

// Validates the list and, when `plan` is given, works out the frame's structure in the same pass: the op table (72 B per
// op), the path table and each draw's paint are read exactly once.
static skb_result validate_dl(const uint8_t* dl, size_t bytes, FramePlan* plan = nullptr) {
  if (bytes < sizeof(skb_dl_header)) return SKB_ERROR_BAD_DISPLAY_LIST;
  skb_dl_header h;
  memcpy(&h, dl, sizeof(h));
  if (h.magic != SKB_DL_MAGIC || h.version != SKB_DL_VERSION || h.total_bytes > bytes) {
    set_error("display list: bad magic/version/size");
    return SKB_ERROR_BAD_DISPLAY_LIST;
  }
  auto in_range = [&](uint32_t off, uint64_t n, size_t sz) { return (uint64_t)off + n * sz <= h.total_bytes && off % 16 == 0; };
  if (!in_range(h.off_surfaces, h.n_surfaces, sizeof(skb_dl_surface)) || !in_range(h.off_ops, h.n_ops, sizeof(skb_dl_op)) ||
      !in_range(h.off_paths, h.n_paths, sizeof(skb_dl_path)) || !in_range(h.off_segs, h.n_segs, sizeof(skb_dl_seg)) ||
      !in_range(h.off_paints, h.n_paints, sizeof(skb_dl_paint)) || !in_range(h.off_stops, h.n_stop_floats, 4) ||
      h.n_surfaces == 0) {
    set_error("display list: section out of range");
    return SKB_ERROR_BAD_DISPLAY_LIST;
  }
  // the surface and op tables stay on the host as dl[0 .. off_paths) (skb_frame_encode): they must lie before the paths
  if ((uint64_t)h.off_surfaces + (uint64_t)h.n_surfaces * sizeof(skb_dl_surface) > h.off_ops ||
      (uint64_t)h.off_ops + (uint64_t)h.n_ops * sizeof(skb_dl_op) > h.off_paths || h.off_paths > h.off_segs) {
    set_error("display list: sections out of order (surfaces, ops, paths, segments)");
    return SKB_ERROR_BAD_DISPLAY_LIST;
  }
  if (h.n_clip_states > h.n_ops) {
    set_error("display list: more clip states than ops");
    return SKB_ERROR_BAD_DISPLAY_LIST;
  }
  const skb_dl_op* ops = (const skb_dl_op*)(dl + h.off_ops);
  const skb_dl_path* paths = (const skb_dl_path*)(dl + h.off_paths);
  const skb_dl_paint* paints = (const skb_dl_paint*)(dl + h.off_paints);
  const skb_dl_surface* vsurfs = (const skb_dl_surface*)(dl + h.off_surfaces);
  for (uint32_t i = 1; i < h.n_surfaces; i++) {  // surface 0 takes the size of the skb_surface it is rendered into
    if (vsurfs[i].width == 0 || vsurfs[i].height == 0 || vsurfs[i].width > 65535u * 16 || vsurfs[i].height > 65535u * 16) {
      set_error("display list: surface size out of range");
      return SKB_ERROR_BAD_DISPLAY_LIST;
    }
  }
  for (uint32_t i = 0; i < h.n_surfaces; i++) {
    if (!(vsurfs[i].flags & SKB_SURFACE_IMAGE)) continue;
    if (i == 0 || (uint64_t)vsurfs[i].reserved + (uint64_t)vsurfs[i].width * vsurfs[i].height * 4 > h.total_bytes) {
      set_error("display list: image pixels out of range");
      return SKB_ERROR_BAD_DISPLAY_LIST;
    }
  }
  std::vector<uint8_t> diff_state((size_t)h.n_clip_states + 1, 0);   // states defined by a ClipOp::kDifference clip
  std::vector<uint32_t> region_of_state((size_t)h.n_clip_states + 1, 0);   // the op whose scan rectangle is the state's table region
  std::vector<int> state_depth(plan ? (size_t)h.n_clip_states + 1 : 0, 0);
  std::vector<uint32_t> levels(h.n_surfaces, 0);     // dependency depth of every surface: a surface is composited after the
  std::vector<uint2> uses;                           // surfaces its draws sample and after the source of the blur that makes it
  if (plan) {
    *plan = FramePlan();
    plan->surf_drawn.assign(h.n_surfaces, 0);
  }
  // The flatten stage lays edge regions out from "op i owns path i's segments, paths in op order, every segment owned":
  // path k (k-th FILL/CLIP op) must be path index k and the paths must tile the segment table without gaps or overlap.
  uint64_t next_seg = 0;
  uint32_t next_path = 0;
  for (uint32_t i = 0; i < h.n_ops; i++) {
    const skb_dl_op& o = ops[i];
    uint32_t src = 0xFFFFFFFFu;   // surface this op samples
    if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
      if (o.path != next_path || next_path >= h.n_paths || paths[next_path].seg_off != next_seg) {
        set_error("display list: every fill / clip op owns the next path, and paths follow each other in the segment table");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      next_seg += paths[next_path].n_segs;
      next_path++;
    }
    if (o.surface < h.n_surfaces && (vsurfs[o.surface].flags & SKB_SURFACE_IMAGE)) {
      set_error("display list: an image surface is read-only");
      return SKB_ERROR_BAD_DISPLAY_LIST;
    }
    if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
      if (o.path >= h.n_paths || o.surface >= h.n_surfaces || (o.kind == SKB_OP_FILL && o.paint >= h.n_paints) ||
          o.clip_in > h.n_clip_states) {
        set_error("display list: op index out of range");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      const skb_dl_path& p = paths[o.path];
      if ((uint64_t)p.seg_off + p.n_segs > h.n_segs) {
        set_error("display list: path segments out of range");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      if (o.kind == SKB_OP_FILL) {
        const skb_dl_paint& pt = paints[o.paint];
        const bool gradient = (pt.type >= SKB_PAINT_LINEAR && pt.type <= SKB_PAINT_SWEEP) || pt.type == SKB_PAINT_CONICAL;
        if (pt.type > SKB_PAINT_CONICAL || (pt.type == SKB_PAINT_IMAGE && pt.image_surface >= h.n_surfaces) ||
            (gradient && ((uint64_t)pt.stop_off + 5ull * pt.n_colors + (pt.type == SKB_PAINT_CONICAL ? 16u : 0u) >
                              h.n_stop_floats ||
                          pt.n_colors < 1))) {
          set_error("display list: bad paint");
          return SKB_ERROR_BAD_DISPLAY_LIST;
        }
        if (pt.blend > 22 || (pt.blend > 15 && pt.blend != 22)) {
          set_error("display list: blend mode outside what SWRenderTarget implements (kClear..kScreen, kSoftLight)");
          return SKB_ERROR_BAD_DISPLAY_LIST;
        }
        if (SKB_PAINT_CF_OFFSET(pt)) {
          const uint64_t off = SKB_PAINT_CF_OFFSET(pt) - 1;
          const uint32_t* blk = (const uint32_t*)(dl + h.off_stops) + off;
          if (off + 4 > h.n_stop_floats || blk[0] < SKB_CF_BLEND || blk[0] > SKB_CF_TABLE ||
              off + (blk[0] == SKB_CF_TABLE ? 68u : 16u) > h.n_stop_floats || (blk[0] == SKB_CF_BLEND && blk[1] > 21)) {
            set_error("display list: bad colour filter block");
            return SKB_ERROR_BAD_DISPLAY_LIST;
          }
        }
      }
      if (o.kind == SKB_OP_FILL) {
        const skb_dl_paint& pt = paints[o.paint];
        if (pt.type == SKB_PAINT_IMAGE) src = pt.image_surface;
        if (plan) {
          plan->surf_drawn[o.surface] = 1;
          if (o.clip_in != 0) plan->clipped_fills = true;
          if ((pt.blend && blend_zero_src_matters(pt.blend - 1)) || SKB_PAINT_CF_OFFSET(pt)) plan->zero_blend = true;
        }
      }
      if (o.kind == SKB_OP_CLIP && (o.clip_out == 0 || o.clip_out > h.n_clip_states)) {
        set_error("display list: clip state id out of range");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      if (o.kind == SKB_OP_CLIP) {
        // ClipOp::kDifference: a state of its own (k_clip_diff mode 0), refined by intersecting clips afterwards
        // (mode 2), or applied on top of intersecting path clips (k_clip_t2).  Difference on difference goes through
        // PerformMerge (sw_canvas.cc:194-217), a std::sort of the two whole span lists with ties: not on the device.
        if (o.aux > 1) {
          set_error("display list: unknown clip op");
          return SKB_ERROR_BAD_DISPLAY_LIST;
        }
        if (o.aux == 0 && o.clip_in != 0 && diff_state[o.clip_in]) {
          set_error("ClipOp::kDifference on top of a ClipOp::kDifference clip (PerformMerge) is not implemented on the device");
          return SKB_ERROR_UNSUPPORTED;
        }
        // a difference clip on top of an intersecting state leaves an intersecting state over the parent's region,
        // which is the region of the first clip of the chain
        const bool t2 = o.aux == 0 && o.clip_in != 0;
        diff_state[o.clip_out] = o.aux == 0 && !t2;
        region_of_state[o.clip_out] = t2 ? region_of_state[o.clip_in] : i;
        if (t2 && plan) {
          ClipT2Rec rec;
          rec.op = i;
          rec.region_op = region_of_state[o.clip_in];
          rec.depth = 0;   // set below, with the op's depth
          rec.pad = 0;
          plan->t2.push_back(rec);
        }
        if (plan) {
          if (!plan->clip_ops) plan->op_depth.assign(h.n_ops, 0);
          plan->clip_ops = true;
          if (o.aux == 0) plan->diff_clips = true;
          const int d = state_depth[o.clip_in] + 1;
          if (d > 250) {
            set_error("clip stack deeper than 250");
            return SKB_ERROR_UNSUPPORTED;
          }
          state_depth[o.clip_out] = d;
          plan->op_depth[i] = (uint8_t)d;
          plan->max_depth = std::max(plan->max_depth, d);
          if (!plan->t2.empty() && plan->t2.back().op == i) plan->t2.back().depth = (uint32_t)d;
        }
      }
    } else if (o.kind == SKB_OP_BLUR) {
      if (o.surface >= h.n_surfaces || o.aux >= h.n_surfaces || o.surface == 0 || o.aux == 0 || o.surface == o.aux) {
        set_error("display list: blur surface out of range");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      if (o.fill_type == 1 || o.fill_type > 7) {
        set_error("display list: unknown blur style");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      src = o.aux;
      if (plan) plan->blur_ops.push_back(o);
    } else {
      set_error("display list: unknown op kind");
      return SKB_ERROR_BAD_DISPLAY_LIST;
    }
    if (src != 0xFFFFFFFFu) {
      levels[o.surface] = std::max(levels[o.surface], levels[src] + 1);
      uses.push_back(make_uint2(src, o.surface));
    }
  }
  if (next_path != h.n_paths || next_seg != h.n_segs) {
    set_error("display list: paths or segments that no op owns");
    return SKB_ERROR_BAD_DISPLAY_LIST;
  }
  for (const uint2& u : uses) {   // a source must have been complete when it was used
    if (levels[u.x] >= levels[u.y]) {
      set_error("display list: a surface is sampled before the draws and blurs that feed it");
      return SKB_ERROR_BAD_DISPLAY_LIST;
    }
  }
  if (plan) {
    for (uint32_t l : levels) plan->max_level = std::max(plan->max_level, l);
    plan->surf_level = std::move(levels);
  }
  return SKB_SUCCESS;
}


static skb_result run_frame(skb_surface s) {
  const uint8_t* dl = s->host_dl.data();
  skb_dl_header h;
  memcpy(&h, dl, sizeof(h));
  const skb_dl_surface* hs = (const skb_dl_surface*)(dl + h.off_surfaces);
  cudaStream_t st = s->stream;
  skb_frame_stats& S = s->stats;
  memset(&S, 0, sizeof(S));
  S.n_ops = h.n_ops;
  S.n_segs = h.n_segs;
  uint32_t launches = 0;

  if (hs[0].width != s->w || hs[0].height != s->h) {
    set_error("display list canvas size differs from the surface");
    return SKB_ERROR_INVALID_ARGUMENT;
  }

  // ---- surfaces: canvas + temporaries (zeroed like Bitmap's calloc, src/io/pixmap.cc:77)
  std::vector<SurfDesc>& surfs = s->h_surfs;
  surfs.assign(h.n_surfaces, SurfDesc());
  std::vector<uint32_t> tile_base(h.n_surfaces + 1, 0);
  size_t temp_bytes = 0;
  std::vector<size_t> temp_off(h.n_surfaces, 0);
  for (uint32_t i = 0; i < h.n_surfaces; i++) {
    SurfDesc& d = surfs[i];
    d.w = hs[i].width;
    d.h = hs[i].height;
    d.tiles_x = cdiv(d.w, SKB_TILE);
    d.tiles_y = cdiv(d.h, SKB_TILE);
    d.pitch = d.tiles_x * SKB_TILE * 4;
    d.tile_base = tile_base[i];
    d.row0 = 0;
    d.row1 = d.h;
    d.level = s->plan.surf_level[i];
    d.pad = 0;
    tile_base[i + 1] = tile_base[i] + d.tiles_x * d.tiles_y;
    if (i > 0) {
      temp_off[i] = temp_bytes;
      temp_bytes += (size_t)d.pitch * d.tiles_y * SKB_TILE;
    }
  }
  if (s->band_y1 > 0) {
    surfs[0].row0 = s->band_y0;
    surfs[0].row1 = s->band_y1 < s->h ? s->band_y1 : s->h;
  }
  SKB_TRY(buf_reserve(s->temp_px, temp_bytes + 256));
  surfs[0].px = s->canvas;
  for (uint32_t i = 1; i < h.n_surfaces; i++) surfs[i].px = (uint8_t*)s->temp_px.p + temp_off[i];
  for (uint32_t i = 0; i < h.n_surfaces; i++) surfs[i].px_out = surfs[i].px;
  if (s->remote_canvas) surfs[0].px_out = s->remote_canvas;
  if (temp_bytes) SKB_CUDA(cudaMemsetAsync(s->temp_px.p, 0, temp_bytes, st));
  for (uint32_t i = 1; i < h.n_surfaces; i++) {  // application images: their pixels came with the display list
    if (hs[i].flags & SKB_SURFACE_IMAGE)
      SKB_CUDA(cudaMemcpy2DAsync(surfs[i].px, surfs[i].pitch, (const uint8_t*)s->dl.p + hs[i].reserved, (size_t)surfs[i].w * 4,
                                 (size_t)surfs[i].w * 4, surfs[i].h, cudaMemcpyDeviceToDevice, st));
  }
  const uint32_t n_tiles = tile_base[h.n_surfaces];
  S.n_tiles = n_tiles;
  SKB_TRY(buf_reserve(s->surfs, surfs.size() * sizeof(SurfDesc)));
  SKB_TRY(buf_reserve(s->surf_tile_base, tile_base.size() * 4));
  SKB_CUDA(cudaMemcpyAsync(s->surfs.p, surfs.data(), surfs.size() * sizeof(SurfDesc), cudaMemcpyHostToDevice, st));
  SKB_CUDA(cudaMemcpyAsync(s->surf_tile_base.p, tile_base.data(), tile_base.size() * 4, cudaMemcpyHostToDevice, st));

  FrameTables t;
  const uint8_t* ddl = (const uint8_t*)s->dl.p;
  t.ops = (const skb_dl_op*)(ddl + h.off_ops);
  t.paths = (const skb_dl_path*)(ddl + h.off_paths);
  t.segs = (const skb_dl_seg*)(ddl + h.off_segs);
  t.paints = (const skb_dl_paint*)(ddl + h.off_paints);
  t.stops = (const float*)(ddl + h.off_stops);
  t.n_ops = h.n_ops;
  t.n_segs = h.n_segs;
  t.wide = s->coord_mode == SKB_COORD_WIDE || (s->coord_mode == SKB_COORD_AUTO && (s->w > 8192 || s->h > 8192)) ? 1u : 0u;
  const uint32_t n_ops = h.n_ops, n_segs = h.n_segs;
  s->n_ops = n_ops;
  if (n_ops == 0) {
    // nothing to launch; the (tiny) upload must still have left the caller's buffer when this returns, as it has
    // on every other path (the first count fetched from the device waits for it)
    SKB_CUDA(cudaStreamSynchronize(st));
    s->flushed = true;
    return SKB_SUCCESS;
  }

  SKB_TRY(buf_reserve(s->geom, (size_t)n_ops * sizeof(OpGeom)));
  SKB_TRY(buf_reserve(s->seg_op, (size_t)(n_segs + 1) * 4));
  SKB_TRY(buf_reserve(s->prim_cnt, (size_t)(n_segs + 1) * 4));
  SKB_TRY(buf_reserve(s->row_cnt, (size_t)(n_ops + 1) * 4));
  SKB_TRY(buf_reserve(s->item_cnt, (size_t)(n_ops + 1) * 4));
  SKB_TRY(buf_reserve(s->counters, 1024));   // words 0..31 counters, 64..127 the sweep's work histogram, 128..129 its pick
  OpGeom* geom = (OpGeom*)s->geom.p;
  uint32_t* seg_op = (uint32_t*)s->seg_op.p;
  uint32_t* prim_off = (uint32_t*)s->prim_cnt.p;
  uint32_t* row_base = (uint32_t*)s->row_cnt.p;
  uint32_t* item_base = (uint32_t*)s->item_cnt.p;
  uint32_t* counters = (uint32_t*)s->counters.p;  // [0] pool_next, [1] overflow

  NvtxStage nvtx;
  nvtx.next("SWRaster_RastePath/flatten");
  cudaEventRecord(s->ev[0], st);
  // ---- stage 1: flatten
  SKB_CUDA(cudaMemsetAsync(seg_op, 0, (size_t)(n_segs + 1) * 4, st));  // segments no op refers to count as op 0's (validate_dl rejects such lists)
  // coverage mode (skb_surface_set_coverage_mode / SKB_COVERAGE_MODE): AREA sends the unclipped fills through
  // k_area_* (tile-binned lines, signed-area accumulation) instead of flatten-to-edges + walk + k_cover
  const bool area_mode = getenv("SKB_COVERAGE_MODE") ? atoi(getenv("SKB_COVERAGE_MODE")) == SKB_COVERAGE_AREA
                                                      : s->coverage_mode == SKB_COVERAGE_AREA;
  k_op_init<<<cdiv(n_ops, 128), 128, 0, st>>>(t, geom, seg_op, (const SurfDesc*)s->surfs.p, area_mode ? 1 : 0);
  launches++;
  uint32_t n_prims = 0, n_area_lines = 0;
  AreaArgs aa;
  memset(&aa, 0, sizeof(aa));
  if (n_segs) {
    SKB_CUDA(cudaMemsetAsync(prim_off + n_segs, 0, 4, st));
    k_seg_count<<<cdiv(n_segs, 256), 256, 0, st>>>(t, prim_off, seg_op, geom);
    launches++;
    SKB_TRY(scan_exclusive(s, prim_off, n_segs + 1, &launches));
    uint32_t two[2] = {0, 0};
    if (area_mode) {
      SKB_TRY(buf_reserve(s->area_line_cnt, (size_t)(n_segs + 1) * 4));
      uint32_t* line_off = (uint32_t*)s->area_line_cnt.p;
      aa.t = t;
      aa.geom = geom;
      aa.seg_op = seg_op;
      aa.line_off = line_off;
      SKB_CUDA(cudaMemsetAsync(line_off + n_segs, 0, 4, st));
      k_area_seg_count<<<cdiv(n_segs, 256), 256, 0, st>>>(aa, line_off);
      launches++;
      SKB_TRY(scan_exclusive(s, line_off, n_segs + 1, &launches));
      k_gather5<<<1, 32, 0, st>>>(counters + 8, prim_off + n_segs, line_off + n_segs, counters + 6, nullptr, nullptr);
      SKB_TRY(fetch_words(s, two, counters + 8, 2));
      launches += 2;
    } else {
      SKB_TRY(fetch_words(s, two, prim_off + n_segs, 1));
      launches++;
    }
    n_prims = two[0];
    n_area_lines = two[1];
  }
  S.n_prims = n_prims;
  S.n_area_lines = n_area_lines;
  S.n_area_tile_lines = 0;
  const size_t n_slots = (size_t)2 * n_prims + (size_t)2 * n_ops + 2;
  S.n_edges_slots = (uint32_t)n_slots;
  SKB_TRY(buf_reserve(s->edges, n_slots * (sizeof(Edge) + sizeof(QuadState))));
  SKB_TRY(buf_reserve(s->walk_lists, (size_t)n_ops * 4 + 16));
  SKB_TRY(buf_reserve(s->ord, n_slots * 4));
  Edge* edges = (Edge*)s->edges.p;
  // stage 3: the sequential sweep (one thread per path), or — skb_surface_set_walk_mode / SKB_WALK_MODE=1 — its
  // row-parallel form (skb_rowwalk.cuh), which emits the same records
  const bool rowwalk = getenv("SKB_WALK_MODE") ? atoi(getenv("SKB_WALK_MODE")) == 1 : s->walk_mode == 1;
  uint32_t* chord_base = nullptr;
  uint32_t* wrow_base = nullptr;
  if (rowwalk) {
    SKB_TRY(buf_reserve(s->rw_chord_cnt, (n_slots + 1) * 4));
    SKB_TRY(buf_reserve(s->rw_slot_op, n_slots * 4));
    SKB_TRY(buf_reserve(s->rw_slots, n_slots * sizeof(SlotInfo)));
    SKB_TRY(buf_reserve(s->rw_rank, n_slots * 2));
    SKB_TRY(buf_reserve(s->rw_ops, (size_t)n_ops * sizeof(RwOp)));
    SKB_TRY(buf_reserve(s->rw_wrow_cnt, (size_t)(n_ops + 1) * 4));
    SKB_TRY(buf_reserve(s->rw_rec_cnt, (size_t)(n_ops + 1) * 4));
    chord_base = (uint32_t*)s->rw_chord_cnt.p;
    wrow_base = (uint32_t*)s->rw_wrow_cnt.p;
  }

  uint64_t n_rows = 0, n_items = 0, n_wrows = 0, n_chords = 0;
  uint32_t pool_cap = 0;
  for (int attempt = 0;; attempt++) {
    if (n_prims) {
      if (rowwalk && attempt == 0) SKB_CUDA(cudaMemsetAsync(chord_base, 0, (n_slots + 1) * 4, st));
      k_flatten<<<cdiv(n_segs, 128), 128, 0, st>>>(t, prim_off, n_prims, seg_op, geom, edges, nullptr, (rowwalk && attempt == 0) ? chord_base : nullptr,
                                                    (uint32_t*)s->rw_slot_op.p);
      launches++;
    }
    if (attempt == 0) {
      cudaEventRecord(s->ev[1], st);
      nvtx.next("SWRaster_RastePath/setup");
      // ---- stage 2: setup
      SKB_CUDA(cudaMemsetAsync(row_base + n_ops, 0, 4, st));
      SKB_CUDA(cudaMemsetAsync(item_base + n_ops, 0, 4, st));
      SKB_CUDA(cudaMemsetAsync(counters + 6, 0, 4, st));
      SKB_CUDA(cudaMemsetAsync(counters + 16, 0, 16, st));   // 64-bit row / item totals (k_op_setup)
      if (rowwalk) SKB_CUDA(cudaMemsetAsync(wrow_base + n_ops, 0, 4, st));
      k_op_setup<<<cdiv(n_ops, 128), 128, 0, st>>>(t, prim_off, geom, (const SurfDesc*)s->surfs.p, row_base, item_base, counters + 6,
                                                   rowwalk ? (RwOp*)s->rw_ops.p : nullptr, wrow_base);
      launches++;
      SKB_TRY(scan_exclusive(s, row_base, n_ops + 1, &launches));
      SKB_TRY(scan_exclusive(s, item_base, n_ops + 1, &launches));
      if (rowwalk) {
        SKB_TRY(scan_exclusive(s, wrow_base, n_ops + 1, &launches));
        SKB_TRY(scan_exclusive(s, chord_base, (uint32_t)n_slots + 1, &launches));
      }
      uint32_t tot[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // one round trip for all the totals: [0..4] gathered, [8..11] the 64-bit totals
      k_gather5<<<1, 32, 0, st>>>(counters + 8, row_base + n_ops, item_base + n_ops, counters + 6, rowwalk ? wrow_base + n_ops : nullptr,
                                  rowwalk ? chord_base + n_slots : nullptr);
      SKB_TRY(fetch_words(s, tot, counters + 8, 12));
      const uint32_t too_big = tot[2];
      launches += 2;
      {
        const uint64_t rows64 = (uint64_t)tot[8] | ((uint64_t)tot[9] << 32), items64 = (uint64_t)tot[10] | ((uint64_t)tot[11] << 32);
        if (rows64 > 0xFFFFFFF0ull || items64 > 0xFFFFFFF0ull) {
          set_error("the frame's scan rows or (draw, tile) items exceed 2^32: too many large draws for one display list");
          return SKB_ERROR_OUT_OF_MEMORY;
        }
      }
      if (too_big) {
        set_error("the scan rectangle of a clip path exceeds 2^28 pixels (it reaches far beyond the surface)");
        return SKB_ERROR_UNSUPPORTED;
      }
      n_rows = tot[0];
      n_items = tot[1];
      n_wrows = tot[3];
      n_chords = tot[4];
      S.n_rows = n_rows;
      S.n_items = n_items;
      s->n_items = (uint32_t)n_items;
      uint64_t want = 2 * n_rows + (uint64_t)SKB_CHUNK * 2 * n_ops + 4096;
      if (rowwalk) want = 2 * n_rows + n_rows / 4 + 65536;   // linear records; chunks only for the few paths swept sequentially
      if (area_mode && n_prims == 0) want = 4096;   // every draw took the AREA route: nothing is swept
      if (want > 0x7FFFFFF0ull) want = 0x7FFFFFF0ull;
      pool_cap = (uint32_t)want;
      SKB_TRY(buf_reserve(s->rows, (n_rows + 1) * sizeof(uint2)));
      SKB_TRY(buf_reserve(s->trow_op, (n_rows / SKB_TILE + 1) * 4));
      SKB_TRY(buf_reserve(s->mask0, (n_items + 1) * 256));
      SKB_TRY(buf_reserve(s->mask1, (area_mode && n_prims == 0) ? 256 : (n_items + 1) * 256));   // AREA coverage is one plane
      SKB_TRY(buf_reserve(s->item_flags, 2 * n_items + 16));
      SKB_TRY(buf_reserve(s->tile_cnt, (size_t)(n_tiles + 1) * 4));
      SKB_TRY(buf_reserve(s->tile_fill, (size_t)(n_tiles + 1) * 4));
      if (rowwalk) {
        SKB_TRY(buf_reserve(s->rw_chords, (n_chords + 1) * sizeof(Chord)));
        SKB_TRY(buf_reserve(s->rw_wgrp_op, (n_wrows / 16 + 1) * 4));
        SKB_TRY(buf_reserve(s->rw_ev, n_wrows + 64));
        SKB_TRY(buf_reserve(s->rw_tab, (n_wrows + 1) * sizeof(RowBand)));
        SKB_TRY(buf_reserve(s->rw_res, (n_wrows + 1) * 4));
        SKB_TRY(buf_reserve(s->rw_rec_off, (n_wrows + 2) * 4));
      }
      cudaEventRecord(s->ev[2], st);
    }
    SKB_TRY(buf_reserve(s->pool, (size_t)pool_cap * sizeof(TrapRec)));
    pool_cap = (uint32_t)(s->pool.cap / sizeof(TrapRec));
    S.pool_capacity = pool_cap;
    // ---- stage 3: walk
    nvtx.next("SWRaster_RastePath/WalkEdges");
    SKB_CUDA(cudaMemsetAsync(counters, 0, 64, st));
    if (n_rows) SKB_CUDA(cudaMemsetAsync(s->rows.p, 0, n_rows * sizeof(uint2), st));
    WalkArgs wa;
    wa.t = t;
    wa.geom = geom;
    wa.row_base = row_base;
    wa.edges = edges;
    wa.edges2 = nullptr;
#ifndef SKB_WALK_NO_COMPACT
    SKB_TRY(buf_reserve(s->edges2, n_slots * (sizeof(Edge) + sizeof(QuadState))));
    wa.edges2 = (uint8_t*)s->edges2.p;
#endif
    wa.quads = nullptr;
    wa.ord = (int32_t*)s->ord.p;
    wa.pool = (TrapRec*)s->pool.p;
    wa.pool_next = counters;
    wa.pool_cap = pool_cap;
    wa.overflow = counters + 1;
    wa.rows = (uint2*)s->rows.p;
    wa.list = (const uint32_t*)s->walk_lists.p;
    wa.count = counters + 4;
    // counters: [0] pool_next, [1] overflow, [4] number of ops to sweep sequentially, [10] retried, [11] swept sequentially
    // few paths (the GPU has thread slots to spare): the longest ones get a warp each (k_walk_hist / k_walk_pick)
    // ... when the paths share their warps eight or sixteen to a warp (lane stride 4 or 2 below); with four paths to a
    // warp the longest path is no faster alone (measured: C1 1.10 vs 1.17 ms; C2 with clips 3.94 -> 3.02 ms)
    static const bool walk_no_long = getenv("SKB_WALK_NO_LONG") != nullptr;
    const uint64_t walk_slots = (uint64_t)s->dev->sm_count * 32 * 32;
    const bool long_first = !rowwalk && !walk_no_long && (uint64_t)n_ops * 2 <= walk_slots && (uint64_t)n_ops * 4 * 2 > walk_slots;
    const uint32_t max_long = long_first ? std::min<uint32_t>(n_ops, (uint32_t)s->dev->sm_count * 8u) : 0u;
    wa.pick = nullptr;
    if (long_first) {
      SKB_CUDA(cudaMemsetAsync(counters + 64, 0, 66 * 4, st));
      k_walk_hist<<<cdiv(n_ops, 128), 128, 0, st>>>(t, geom, counters + 64);
      k_walk_pick<<<1, 1, 0, st>>>(counters + 64, max_long, counters + 128);
      launches += 2;
      wa.pick = counters + 128;
    }
    k_walk_list<<<cdiv(n_ops, 128), 128, 0, st>>>(t, geom, counters + 4, (uint32_t*)s->walk_lists.p, row_base, (uint32_t*)s->trow_op.p,
                                                  (rowwalk && attempt == 0) ? (RwOp*)s->rw_ops.p : nullptr, wrow_base, (uint32_t*)s->rw_wgrp_op.p,
                                                  wa.pick, counters + 12);
    launches++;
    uint32_t n_seq = n_ops;   // upper bound of the paths the sequential sweep gets
    if (rowwalk && n_wrows && attempt == 0) {
      RwArgs ra;
      ra.t = t;
      ra.geom = geom;
      ra.row_base = row_base;
      ra.edges = (uint8_t*)edges;
      ra.ops = (RwOp*)s->rw_ops.p;
      ra.wgrp_op = (const uint32_t*)s->rw_wgrp_op.p;
      ra.n_wgrps = (uint32_t)(n_wrows / 16);
      ra.slot_op = (const uint32_t*)s->rw_slot_op.p;
      ra.n_slots_total = (uint32_t)n_slots;
      ra.chord_base = chord_base;
      ra.slots = (SlotInfo*)s->rw_slots.p;
      ra.chords = (Chord*)s->rw_chords.p;
      ra.ev_words = (uint32_t*)s->rw_ev.p;
      ra.tab = (RowBand*)s->rw_tab.p;
      ra.res = (uint32_t*)s->rw_res.p;
      ra.rec_off = (uint32_t*)s->rw_rec_off.p;
      ra.rec_cnt = (uint32_t*)s->rw_rec_cnt.p;
      ra.rank = (uint16_t*)s->rw_rank.p;
      ra.ord = (int32_t*)s->ord.p;
      ra.pool = (TrapRec*)s->pool.p;
      ra.pool_cap = pool_cap;
      ra.pool_next = counters;
      ra.overflow = counters + 1;
      ra.rows = (uint2*)s->rows.p;
      ra.n_failed = counters + 10;
      const uint32_t g_slots = cdiv(n_slots, 128), g_ops = cdiv(n_ops, 128), g_rows = cdiv(n_wrows, RW_ROWS_BLOCK);
      SKB_CUDA(cudaMemsetAsync(s->rw_ev.p, 0, n_wrows + 64, st));
      k_rw_events<<<g_slots, 128, 0, st>>>(ra);
      k_rw_bands0<<<g_ops, 128, 0, st>>>(ra, 0);
      launches += 2;
      for (int phase = 0; phase < 2; phase++) {
        if (phase == 1) {
          k_rw_bands0<<<g_ops, 128, 0, st>>>(ra, 1);
          launches++;
        }
        k_rw_chain<<<g_slots, 128, 0, st>>>(ra, phase);
        k_rw_rows<0><<<g_rows, RW_ROWS_BLOCK, 0, st>>>(ra, phase);
        k_rw_resolve<<<g_ops, 128, 0, st>>>(ra, phase);
        launches += 3;
        if (phase == 0) {
          SKB_CUDA(cudaMemsetAsync(ra.rec_cnt + n_ops, 0, 4, st));
          SKB_TRY(scan_exclusive(s, ra.rec_cnt, n_ops + 1, &launches));
          k_rw_rec_base<<<g_ops, 128, 0, st>>>(ra);
          launches++;
        }
        k_rw_chain<<<g_slots, 128, 0, st>>>(ra, phase);
        k_rw_rows<1><<<g_rows, RW_ROWS_BLOCK, 0, st>>>(ra, phase);
        launches += 2;
      }
      k_rw_fallback_list<<<g_ops, 128, 0, st>>>(ra, counters + 4, (uint32_t*)s->walk_lists.p);
      launches++;
    }
    {
      // the sequential sweep: every path (SKB_WALK_MODE=0, or a re-run after a pool overflow), else only those listed.
      // Paths are spread over warps while the GPU has spare thread slots (about 32 warps per SM wanted; measured on C1:
      // stride 1 5.2 ms, 4 2.6 ms, 8 2.0 ms, 32 2.6 ms)
      int lane_stride = 1;
      const uint64_t want_threads = (uint64_t)s->dev->sm_count * 32 * 32;
      while (lane_stride < 8 && (uint64_t)n_seq * lane_stride * 2 <= want_threads) lane_stride *= 2;
      if (getenv("SKB_WALK_LANE_STRIDE")) lane_stride = atoi(getenv("SKB_WALK_LANE_STRIDE"));
      // SKB_WALK_SMEM: dynamic shared memory per block, only to cap the blocks (= paths) in flight per SM — the sweep's
      // working set is ~3 KB of edges per path, and with every thread slot taken it spills from L2 to DRAM
      static const int walk_smem = getenv("SKB_WALK_SMEM") ? atoi(getenv("SKB_WALK_SMEM")) : 0;
      if (walk_smem > 48 * 1024) cudaFuncSetAttribute(k_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, walk_smem);
      static const int walk_carveout = getenv("SKB_WALK_CARVEOUT") ? atoi(getenv("SKB_WALK_CARVEOUT")) : -2;
      if (walk_carveout >= -1) cudaFuncSetAttribute(k_walk, cudaFuncAttributePreferredSharedMemoryCarveout, walk_carveout);
      uint32_t walk_grid = cdiv((uint64_t)n_seq * lane_stride + (uint64_t)max_long * 32, WALK_BLOCK);
      static const int walk_resident = getenv("SKB_WALK_RESIDENT") ? atoi(getenv("SKB_WALK_RESIDENT")) : 0;   // blocks per SM
      if (walk_resident > 0) walk_grid = std::min(walk_grid, (uint32_t)(walk_resident * s->dev->sm_count));
      k_walk<<<walk_grid, WALK_BLOCK, walk_smem, st>>>(wa, lane_stride);
      launches++;
    }
    uint32_t hc[12];
    SKB_TRY(fetch_words(s, hc, counters, 12));
    launches++;
    S.n_records = hc[0];
    S.n_rw_retried = hc[10];
    S.n_rw_sequential = hc[11];
    if (!hc[1]) break;
    if (attempt >= 6 || pool_cap >= 0x7FFFFFF0u) {
      set_error("trapezoid record pool exhausted");
      return SKB_ERROR_OUT_OF_MEMORY;
    }
    S.n_retries++;
    uint64_t bigger = (uint64_t)pool_cap * 2;
    pool_cap = bigger > 0x7FFFFFF0ull ? 0x7FFFFFF0u : (uint32_t)bigger;
    // the sweep mutates the edges: rebuild them and walk again (the re-run is the sequential sweep of every path)
  }
  cudaEventRecord(s->ev[3], st);

  // ---- stage 4: coverage
  nvtx.next("SWRaster_RastePath/SpanBuilder");
  CoverArgs ca;
  ca.geom = geom;
  ca.item_base = item_base;
  ca.row_base = row_base;
  ca.trow_op = (const uint32_t*)s->trow_op.p;
  ca.n_ops = n_ops;
  ca.n_items = (uint32_t)n_items;
  ca.n_trows = (uint32_t)(n_rows / SKB_TILE);
  ca.ops = t.ops;
  ca.surfs = (const SurfDesc*)s->surfs.p;
  ca.pool = (const TrapRec*)s->pool.p;
  ca.rows = (const uint2*)s->rows.p;
  ca.mask0 = (uint8_t*)s->mask0.p;
  ca.mask1 = (uint8_t*)s->mask1.p;
  for (int k = 0; k < SKB_CLIP_PLANES; k++) ca.mask[k] = nullptr;
  ca.mask[0] = ca.mask0;
  ca.mask[1] = ca.mask1;
  ca.item_flags = (uint16_t*)s->item_flags.p;
  ca.tile_cnt = (uint32_t*)s->tile_cnt.p;
  ca.paints = t.paints;
  ca.zmask = nullptr;
  for (int k = 0; k < SKB_CLIP_PLANES; k++) ca.zplane[k] = nullptr;
  if (s->plan.zero_blend) {
    SKB_TRY(buf_reserve(s->zmask, (n_items + 1) * 256));
    ca.zmask = (uint8_t*)s->zmask.p;
    ca.zplane[0] = ca.zmask;
  }
  // clip structure of the frame (worked out by skb_frame_encode): nesting depth of every clip state, clipped draws present?
  const bool has_clip_ops = s->plan.clip_ops, has_clipped_fills = s->plan.clipped_fills;
  const int max_depth = s->plan.max_depth;
  const std::vector<uint8_t>& op_depth = s->plan.op_depth;
  if (has_clipped_fills) {
    for (int k = 2; k < SKB_CLIP_PLANES; k++) {
      SKB_TRY(buf_reserve(s->mask_extra[k - 2], (n_items + 1) * 256));
      ca.mask[k] = (uint8_t*)s->mask_extra[k - 2].p;
    }
    // clipped draws write single bytes of their planes: start from zero
    for (int k = 0; k < SKB_CLIP_PLANES; k++) SKB_CUDA(cudaMemsetAsync(ca.mask[k], 0, n_items * 256, st));
    if (s->plan.zero_blend) {  // ... and of the per-plane "a span reaches this pixel" maps, when some paint needs them
      for (int k = 1; k < SKB_CLIP_PLANES; k++) {
        SKB_TRY(buf_reserve(s->zplane_extra[k - 1], (n_items + 1) * 256));
        ca.zplane[k] = (uint8_t*)s->zplane_extra[k - 1].p;
      }
      for (int k = 0; k < SKB_CLIP_PLANES; k++) SKB_CUDA(cudaMemsetAsync(ca.zplane[k], 0, n_items * 256, st));
    }
  }
  SKB_CUDA(cudaMemsetAsync(s->tile_cnt.p, 0, (size_t)(n_tiles + 1) * 4, st));
  SKB_CUDA(cudaMemsetAsync(s->tile_fill.p, 0, (size_t)(n_tiles + 1) * 4, st));
  if (n_items) {
    SKB_CUDA(cudaMemsetAsync(s->item_flags.p, 0, 2 * n_items, st));
    if (!area_mode || n_prims) {   // nothing was swept when every draw took the AREA route
      k_cover<<<cdiv(ca.n_trows, COVER_WARPS), COVER_WARPS * 32, 0, st>>>(ca);
      launches++;
    }
  }
  if (area_mode && n_area_lines && n_items) {
    // ---- stages 2-3 of the AREA mode: bin the lines (count, scan, write), resolve the backdrops, accumulate
    const uint32_t n_trows = ca.n_trows;
    SKB_TRY(buf_reserve(s->area_item_cnt, (n_items + 1) * 4));
    SKB_TRY(buf_reserve(s->area_item_cursor, (n_items + 1) * 4));
    SKB_TRY(buf_reserve(s->area_item_local, (n_items + 1) * 4));
    SKB_TRY(buf_reserve(s->area_item_delta, (n_items + 1) * 4));
    SKB_TRY(buf_reserve(s->area_row_backdrop, ((size_t)n_trows + 1) * 4));
    aa.n_lines = n_area_lines;
    aa.item_cnt = (uint32_t*)s->area_item_cnt.p;
    aa.item_cursor = (uint32_t*)s->area_item_cursor.p;
    aa.item_local = (int32_t*)s->area_item_local.p;
    aa.item_delta = (int32_t*)s->area_item_delta.p;
    aa.row_backdrop = (int32_t*)s->area_row_backdrop.p;
    aa.lines = nullptr;
    aa.c = ca;
    SKB_CUDA(cudaMemsetAsync(aa.item_cnt, 0, (n_items + 1) * 4, st));
    SKB_CUDA(cudaMemsetAsync(aa.item_cursor, 0, (n_items + 1) * 4, st));
    SKB_CUDA(cudaMemsetAsync(aa.item_local, 0, (n_items + 1) * 4, st));
    SKB_CUDA(cudaMemsetAsync(aa.item_delta, 0, (n_items + 1) * 4, st));
    SKB_CUDA(cudaMemsetAsync(aa.row_backdrop, 0, ((size_t)n_trows + 1) * 4, st));
    k_area_bin<<<cdiv(n_area_lines, 128), 128, 0, st>>>(aa, 0);
    launches++;
    SKB_TRY(scan_exclusive(s, aa.item_cnt, (uint32_t)n_items + 1, &launches));
    uint32_t n_tile_lines = 0;
    SKB_TRY(fetch_words(s, &n_tile_lines, aa.item_cnt + n_items, 1));
    launches++;
    S.n_area_tile_lines = n_tile_lines;
    SKB_TRY(buf_reserve(s->area_lines, ((size_t)n_tile_lines + 1) * sizeof(uint4)));
    aa.lines = (uint4*)s->area_lines.p;
    k_area_bin<<<cdiv(n_area_lines, 128), 128, 0, st>>>(aa, 1);
    k_area_backdrop<<<cdiv(n_trows, 128), 128, 0, st>>>(aa);
    k_area_cover<<<cdiv(n_trows, AREA_WARPS), AREA_WARPS * 32, 0, st>>>(aa);
    launches += 3;
  }
  cudaEventRecord(s->ev[9], st);
  nvtx.next("SWCanvas_OnClipPath");
  if (has_clip_ops && n_rows) {
    // ---- stage 4b: clip states (by nesting depth), then the clipped draws
    const uint32_t n_states = h.n_clip_states;
    SKB_TRY(buf_reserve(s->clip_states, (size_t)(n_states + 2) * sizeof(ClipStateDesc)));
    SKB_TRY(buf_reserve(s->clip_px, (size_t)(n_states + 2) * 4));
    SKB_TRY(buf_reserve(s->op_depth, n_ops));
    SKB_CUDA(cudaMemsetAsync(s->clip_states.p, 0, (size_t)(n_states + 2) * sizeof(ClipStateDesc), st));
    SKB_CUDA(cudaMemsetAsync(s->clip_px.p, 0, (size_t)(n_states + 2) * 4, st));
    SKB_CUDA(cudaMemcpyAsync(s->op_depth.p, op_depth.data(), n_ops, cudaMemcpyHostToDevice, st));
    const uint32_t n_t2 = (uint32_t)s->plan.t2.size();
    SKB_TRY(buf_reserve(s->clip_t2, (size_t)(n_t2 + 1) * sizeof(ClipT2Rec)));
    if (n_t2) SKB_CUDA(cudaMemcpyAsync(s->clip_t2.p, s->plan.t2.data(), (size_t)n_t2 * sizeof(ClipT2Rec), cudaMemcpyHostToDevice, st));
    k_clip_sizes<<<cdiv(n_ops, 128), 128, 0, st>>>(t, geom, (const SurfDesc*)s->surfs.p, (ClipStateDesc*)s->clip_states.p,
                                                 (uint32_t*)s->clip_px.p, (const ClipT2Rec*)s->clip_t2.p, n_t2);
    launches++;
    SKB_TRY(scan_exclusive(s, (uint32_t*)s->clip_px.p, n_states + 2, &launches));
    uint32_t total_px = 0;
    SKB_TRY(fetch_words(s, &total_px, (uint32_t*)s->clip_px.p + n_states + 1, 1));
    launches++;
    SKB_TRY(buf_reserve(s->clip_table, ((size_t)total_px + 1) * SKB_CLIP_MAXE * 4));
    SKB_CUDA(cudaMemsetAsync(s->clip_table.p, 0, ((size_t)total_px + 1) * SKB_CLIP_MAXE * 4, st));
    ClipArgs cl;
    cl.c = ca;
    cl.states = (ClipStateDesc*)s->clip_states.p;
    cl.state_px_off = (const uint32_t*)s->clip_px.p;
    cl.table = (uint32_t*)s->clip_table.p;
    cl.op_depth = (const uint8_t*)s->op_depth.p;
    cl.overflow = counters + 2;
    cl.n_rows = (uint32_t)n_rows;
    cl.t2 = (const ClipT2Rec*)s->clip_t2.p;
    cl.n_t2 = n_t2;
    const uint32_t clip_grid = cdiv(n_rows * 32, 128);
    if (s->plan.diff_clips && clip_grid) {   // difference states have no parent: all of them first
      k_clip_diff<<<clip_grid, 128, 0, st>>>(cl, 0, 0);
      launches++;
    }
    for (int level = 1; level <= max_depth && clip_grid; level++) {
      k_clip_rows<<<clip_grid, 128, 0, st>>>(cl, 0, level);
      launches++;
      if (s->plan.diff_clips && level >= 2) {   // intersecting clips on top of a difference state
        k_clip_diff<<<clip_grid, 128, 0, st>>>(cl, 2, level);
        launches++;
      }
      if (n_t2 && level >= 2) {   // difference clips on top of an intersecting state: their row spans, then the cut
        k_clip_diff<<<clip_grid, 128, 0, st>>>(cl, 0, level);
        k_clip_t2<<<dim3(32, n_t2), 128, 0, st>>>(cl, level);
        launches += 2;
      }
    }
    if (has_clipped_fills && clip_grid && n_items) {
      k_clip_rows<<<clip_grid, 128, 0, st>>>(cl, 1, 0);
      launches++;
      if (s->plan.diff_clips) {
        k_clip_diff<<<clip_grid, 128, 0, st>>>(cl, 1, 0);
        launches++;
      }
      k_clip_classify<<<cdiv(n_items * 32, 128), 128, 0, st>>>(ca);
      launches++;
    }
    uint32_t over = 0;
    SKB_TRY(fetch_words(s, &over, counters + 2, 1));
    launches++;
    if (over) {
      set_error("clip stack: a pixel is covered by more clip spans / coverage planes than the device tables hold, two clip spans of one row are indistinguishable (same start and coverage), or a difference clip met an empty parent that was itself a difference clip");
      return SKB_ERROR_UNSUPPORTED;
    }
  }
  cudaEventRecord(s->ev[4], st);
  nvtx.next("SWSpanBrush_Brush/bin");
  // ---- stage 5: bin
  SKB_TRY(scan_exclusive(s, (uint32_t*)s->tile_cnt.p, n_tiles + 1, &launches));
  uint32_t n_cmds = 0;
  SKB_TRY(fetch_words(s, &n_cmds, (uint32_t*)s->tile_cnt.p + n_tiles, 1));
  launches++;
  S.n_cmds = n_cmds;
  SKB_TRY(buf_reserve(s->cmds, ((size_t)n_cmds + 1) * sizeof(uint2)));
  SKB_TRY(buf_reserve(s->cmds_sorted, ((size_t)n_cmds + 1) * sizeof(uint2)));
  if (n_items) {
    k_scatter<<<cdiv(ca.n_trows, 128), 128, 0, st>>>(ca, (const uint32_t*)s->tile_cnt.p, (uint32_t*)s->tile_fill.p, (uint2*)s->cmds.p);
    launches++;
  }
  cudaEventRecord(s->ev[5], st);

  // ---- stages 6 + 7: fine pass over temporaries, blur, fine pass over the canvas
  nvtx.next("SWSpanBrush_Brush/fine+SWCanvas_HandleFilter");
  FineArgs fa;
  fa.tile_off = (const uint32_t*)s->tile_cnt.p;
  fa.cmds = (const uint2*)s->cmds.p;
  fa.cmds_sorted = (uint2*)s->cmds_sorted.p;
  fa.surfs = (const SurfDesc*)s->surfs.p;
  fa.surf_tile_base = (const uint32_t*)s->surf_tile_base.p;
  fa.n_surfaces = h.n_surfaces;
  fa.remote_store = s->remote_canvas ? 1u : 0u;
  fa.geom = geom;
  fa.ops = t.ops;
  fa.paints = t.paints;
  fa.stops = t.stops;
  for (int k = 0; k < SKB_CLIP_PLANES; k++) fa.mask[k] = ca.mask[k];
  fa.zmask = ca.zmask;
  for (int k = 0; k < SKB_CLIP_PLANES; k++) fa.zplane[k] = ca.zplane[k];
  // blur jobs of the whole frame, sorted by the level of their destination
  std::vector<BlurJob> jobs;
  for (const skb_dl_op& bo : s->plan.blur_ops) {
    {
      BlurJob j;
      j.src = bo.aux;
      j.dst = bo.surface;
      j.radius = (int32_t)bo.clip_bounds[0];
      j.style = bo.fill_type;
      j.color = bo.paint;
      j.morph_rx = bo.clip_bounds[0];
      j.morph_ry = bo.clip_bounds[1];
      if (surfs[j.src].w != surfs[j.dst].w || surfs[j.src].h != surfs[j.dst].h) {
        set_error("blur: source and destination surfaces differ in size");
        return SKB_ERROR_BAD_DISPLAY_LIST;
      }
      jobs.push_back(j);
    }
  }
  std::stable_sort(jobs.begin(), jobs.end(), [&](const BlurJob& x, const BlurJob& y) { return surfs[x.dst].level < surfs[y.dst].level; });
  const uint32_t nj = (uint32_t)jobs.size();
  std::vector<uint32_t> rowb(nj + 1, 0), colb(nj + 1, 0);
  std::vector<const uint8_t*> tmp_ptrs(nj);
  size_t blur_smem = 0;
  uint32_t blur_max_len = 0;
  if (nj) {
    size_t max_len = 0, tmp_bytes = 0;
    std::vector<size_t> tmp_off(nj);
    for (uint32_t i = 0; i < nj; i++) {
      const SurfDesc& d = surfs[jobs[i].dst];
      rowb[i + 1] = rowb[i] + d.h;
      colb[i + 1] = colb[i] + d.w;
      int r = jobs[i].radius > 254 ? 254 : jobs[i].radius;
      if (r > 1) max_len = std::max(max_len, (size_t)d.w + 2 * (size_t)(r + 1) + 1);
      tmp_off[i] = tmp_bytes;
      tmp_bytes += (size_t)d.pitch * d.tiles_y * SKB_TILE;
    }
    blur_max_len = (uint32_t)max_len;
    max_len = (max_len + 3) & ~(size_t)3;
    blur_max_len = (uint32_t)max_len;
    blur_smem = max_len * 16 + max_len * 4 + 16 * BLUR_H_WARPS + 16;
    if (blur_smem > 200 * 1024) {
      set_error("blur: surface too wide for the shared-memory row scan");
      return SKB_ERROR_UNSUPPORTED;
    }
    // H pass writes into a scratch copy of each destination, V pass reads it and writes the destination
    SKB_TRY(buf_reserve(s->blur_tmp, tmp_bytes + 256));
    SKB_TRY(buf_reserve(s->blur_jobs, nj * sizeof(BlurJob)));
    SKB_TRY(buf_reserve(s->blur_rows, (nj + 1) * 4));
    SKB_TRY(buf_reserve(s->blur_cols, (nj + 1) * 4));
    SKB_TRY(buf_reserve(s->blur_tmp_ptrs, nj * sizeof(void*)));
    for (uint32_t i = 0; i < nj; i++) tmp_ptrs[i] = (const uint8_t*)s->blur_tmp.p + tmp_off[i];
    // the H kernel addresses its destination through SurfDesc: give it a table where dst.px = scratch
    std::vector<SurfDesc> htab = surfs;
    for (uint32_t i = 0; i < nj; i++) htab[jobs[i].dst].px = (uint8_t*)tmp_ptrs[i];
    Buf& htab_buf = s->scan_tmp;  // reuse: scans are done
    SKB_TRY(buf_reserve(htab_buf, htab.size() * sizeof(SurfDesc)));
    SKB_CUDA(cudaMemcpyAsync(htab_buf.p, htab.data(), htab.size() * sizeof(SurfDesc), cudaMemcpyHostToDevice, st));
    SKB_CUDA(cudaMemcpyAsync(s->blur_jobs.p, jobs.data(), nj * sizeof(BlurJob), cudaMemcpyHostToDevice, st));
    SKB_CUDA(cudaMemcpyAsync(s->blur_rows.p, rowb.data(), (nj + 1) * 4, cudaMemcpyHostToDevice, st));
    SKB_CUDA(cudaMemcpyAsync(s->blur_cols.p, colb.data(), (nj + 1) * 4, cudaMemcpyHostToDevice, st));
    SKB_CUDA(cudaMemcpyAsync(s->blur_tmp_ptrs.p, tmp_ptrs.data(), nj * sizeof(void*), cudaMemcpyHostToDevice, st));
  }
  if (s->plan.max_level >= 16) {
    set_error("more than 16 dependent passes (nested layers / filters)");
    return SKB_ERROR_UNSUPPORTED;
  }
  // Level by level: first the blurs that produce surfaces of this level, then the fine pass of the
  // surfaces drawn at this level (their image sources and blur inputs are complete by construction).
  s->n_levels_timed = s->plan.max_level + 1;
  uint32_t j0 = 0;
  for (uint32_t level = 0; level <= s->plan.max_level; level++) {
    if (!s->ev_blur[2 * level]) {
      SKB_CUDA(cudaEventCreate(&s->ev_blur[2 * level]));
      SKB_CUDA(cudaEventCreate(&s->ev_blur[2 * level + 1]));
    }
    cudaEventRecord(s->ev_blur[2 * level], st);
    uint32_t j1 = j0;
    while (j1 < nj && surfs[jobs[j1].dst].level == level) j1++;
    if (j1 > j0) {
      k_blur_h<<<rowb[j1] - rowb[j0], BLUR_H_WARPS * 32, blur_smem, st>>>(
          (const BlurJob*)s->blur_jobs.p, (const uint32_t*)s->blur_rows.p, nj, (const SurfDesc*)s->scan_tmp.p, rowb[j0], rowb[j1],
          blur_max_len);
      launches++;
      // radius <= 1: the H kernel copied src into scratch; finish with a plain copy into dst
      for (uint32_t i = j0; i < j1; i++) {
        int r = jobs[i].radius > 254 ? 254 : jobs[i].radius;
        if (r <= 1 && jobs[i].style < 6) {
          const SurfDesc& d = surfs[jobs[i].dst];
          SKB_CUDA(cudaMemcpyAsync(d.px, tmp_ptrs[i], (size_t)d.pitch * d.h, cudaMemcpyDeviceToDevice, st));
        }
      }
      k_blur_v<<<cdiv(colb[j1] - colb[j0], 128), 128, 0, st>>>((const BlurJob*)s->blur_jobs.p, (const uint32_t*)s->blur_cols.p, nj,
                                                                 (const SurfDesc*)s->surfs.p,
                                                                 (const uint8_t* const*)s->blur_tmp_ptrs.p, colb[j0], colb[j1]);
      launches++;
      for (uint32_t i = j0; i < j1; i++) {
        if (jobs[i].style < 6) continue;
        // MorphologyImageFilter::OnFilter (image_filter.cc:342-385): x then y when both radii are positive, else the one
        const SurfDesc& sdesc = surfs[jobs[i].src];
        const SurfDesc& d = surfs[jobs[i].dst];
        const int erode = jobs[i].style == 7, mw = (int)d.w, mh = (int)d.h;
        const uint32_t mgrid = cdiv((uint64_t)mw * mh, 256);
        const float rxf = jobs[i].morph_rx, ryf = jobs[i].morph_ry;
        if (rxf > 0 && ryf > 0) {
          k_morph<<<mgrid, 256, 0, st>>>(sdesc.px, sdesc.pitch, (uint8_t*)tmp_ptrs[i], d.pitch, mw, mh, (int)rxf, 0, erode);
          k_morph<<<mgrid, 256, 0, st>>>(tmp_ptrs[i], d.pitch, d.px, d.pitch, mw, mh, (int)ryf, 1, erode);
          launches += 2;
        } else if (rxf > 0) {
          k_morph<<<mgrid, 256, 0, st>>>(sdesc.px, sdesc.pitch, d.px, d.pitch, mw, mh, (int)rxf, 0, erode);
          launches++;
        } else if (ryf > 0) {
          k_morph<<<mgrid, 256, 0, st>>>(sdesc.px, sdesc.pitch, d.px, d.pitch, mw, mh, (int)ryf, 1, erode);
          launches++;
        }
      }
      for (uint32_t i = j0; i < j1; i++) {
        if (jobs[i].style == 0 || jobs[i].style >= 6) continue;
        const SurfDesc& d = surfs[jobs[i].dst];
        k_blur_style<<<cdiv((uint64_t)d.w * d.h, 256), 256, 0, st>>>(surfs[jobs[i].src], d, jobs[i].style, jobs[i].color);
        launches++;
      }
    }
    j0 = j1;
    cudaEventRecord(s->ev_blur[2 * level + 1], st);
    // fine pass of this level
    fa.level = level;
    if (surfs[0].level == level) {  // the canvas: only the tiles of this device's band
      uint32_t ty0 = surfs[0].row0 / SKB_TILE, ty1 = cdiv(surfs[0].row1, SKB_TILE);
      fa.tile_begin = ty0 * surfs[0].tiles_x;
      fa.tile_end = ty1 * surfs[0].tiles_x;
      if (fa.tile_end > fa.tile_begin) {
        k_fine<<<cdiv(fa.tile_end - fa.tile_begin, FINE_WARPS), FINE_WARPS * 32, 0, st>>>(fa);
        launches++;
      }
    }
    bool others = false;
    for (uint32_t i = 1; i < h.n_surfaces && !others; i++) others = surfs[i].level == level && s->plan.surf_drawn[i];
    if (others) {
      fa.tile_begin = tile_base[1];
      fa.tile_end = n_tiles;
      k_fine<<<cdiv(fa.tile_end - fa.tile_begin, FINE_WARPS), FINE_WARPS * 32, 0, st>>>(fa);
      launches++;
    }
  }
  cudaEventRecord(s->ev[8], st);
  SKB_CUDA(cudaGetLastError());
  S.n_launches = launches;
  // algorithmic bytes: fine = canvas band load+store + 8 B/command + 256 B/mask command;
  // coverage = 32 B/record in + 256 B/non-empty mask out
  uint64_t band_px = (uint64_t)(surfs[0].row1 - surfs[0].row0) * s->w;
  S.bytes_fine = band_px * 8 + (uint64_t)n_cmds * (8 + 256);
  S.bytes_cover = S.n_records * 32 + (uint64_t)n_cmds * 256;
  // AREA coverage pass (k_area_cover): 16 B per binned line in, 4 B backdrop per item in, 256 B per non-empty mask out
  S.bytes_area = (uint64_t)S.n_area_tile_lines * 16 + (area_mode ? n_items * 4 + (uint64_t)n_cmds * 256 : 0);
  S.bytes_walk = (uint64_t)S.n_edges_slots * (sizeof(Edge) + sizeof(QuadState)) + S.n_records * 32 + S.n_rows * 8;
  if (rowwalk) S.bytes_walk += n_chords * sizeof(Chord) * 2 + n_wrows * (sizeof(RowBand) + 8);  // chords written once and read by the rows; band table, row results
  S.bytes_blur = 0;
  for (uint32_t i = 0; i < nj; i++) {
    const int r = jobs[i].radius > 254 ? 254 : jobs[i].radius;
    if (jobs[i].style < 6 && r > 1) S.bytes_blur += (uint64_t)surfs[jobs[i].dst].w * surfs[jobs[i].dst].h * 16;
  }
  s->flushed = true;
  return SKB_SUCCESS;
}

}  // namespace skb

// ================================================================================ C ABI
extern "C" {

const char* skb_get_last_error_string(void) { return g_last_error.c_str(); }
const char* skb_version_string(void) { return "skity-b200 0.1 (sm_100a)"; }

skb_result skb_device_create(int ordinal, skb_device* out) {
  if (!out) return SKB_ERROR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)");
    return SKB_ERROR_NO_DEVICE;
  }
  if (ordinal < 0 || ordinal >= count) {
    set_error("device ordinal out of range");
    return SKB_ERROR_INVALID_ARGUMENT;
  }
  SKB_CUDA(cudaSetDevice(ordinal));
  cudaDeviceProp prop;
  SKB_CUDA(cudaGetDeviceProperties(&prop, ordinal));
  if (prop.major < 10) {
    set_error(std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only");
    return SKB_ERROR_NO_DEVICE;
  }
  skb_device d = new skb_device_s();
  d->ordinal = ordinal;
  // process-wide kernel attribute, set once here (to the most run_frame ever asks for) rather than per frame: surfaces
  // of one device may be driven from different threads
  SKB_CUDA(cudaFuncSetAttribute(k_blur_h, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  d->sm_count = prop.multiProcessorCount;
  *out = d;
  return SKB_SUCCESS;
}

void skb_device_destroy(skb_device d) { delete d; }

skb_result skb_device_sm_count(skb_device d, int* out) {
  if (!d || !out) return SKB_ERROR_INVALID_ARGUMENT;
  *out = d->sm_count;
  return SKB_SUCCESS;
}

skb_result skb_surface_create(skb_device d, uint32_t w, uint32_t h, skb_surface* out) {
  if (!d || !out || w == 0 || h == 0 || w > 65535 * 16 || h > 65535 * 16) return SKB_ERROR_INVALID_ARGUMENT;
  *out = nullptr;
  SKB_CUDA(cudaSetDevice(d->ordinal));
  skb_surface s = new skb_surface_s();
  s->dev = d;
  s->w = w;
  s->h = h;
  s->tiles_x = cdiv(w, SKB_TILE);
  s->tiles_y = cdiv(h, SKB_TILE);
  s->pitch = s->tiles_x * SKB_TILE * 4;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->canvas, (size_t)s->pitch * s->tiles_y * SKB_TILE);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->canvas, 0, (size_t)s->pitch * s->tiles_y * SKB_TILE, s->stream);
  for (int i = 0; i < 12 && e == cudaSuccess; i++) e = cudaEventCreate(&s->ev[i]);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->mapped_host, 64, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&s->mapped_dev, s->mapped_host, 0);
  if (e != cudaSuccess) {
    set_error(std::string("surface create: ") + cudaGetErrorString(e));
    skb_surface_destroy(s);
    return e == cudaErrorMemoryAllocation ? SKB_ERROR_OUT_OF_MEMORY : SKB_ERROR_CUDA;
  }
  *out = s;
  return SKB_SUCCESS;
}

void skb_surface_destroy(skb_surface s) {
  if (!s) return;
  cudaSetDevice(s->dev->ordinal);
  if (s->stream) cudaStreamSynchronize(s->stream);
  Buf* bufs[] = {&s->area_line_cnt, &s->area_item_cnt, &s->area_item_cursor, &s->area_item_local, &s->area_item_delta, &s->area_row_backdrop, &s->area_lines,
                 &s->rw_chord_cnt, &s->rw_slot_op, &s->rw_slots, &s->rw_rank, &s->rw_ops, &s->rw_wrow_cnt, &s->rw_rec_cnt, &s->rw_chords, &s->rw_wgrp_op, &s->rw_ev, &s->rw_tab, &s->rw_res, &s->rw_rec_off, &s->dl, &s->geom, &s->seg_op, &s->prim_cnt, &s->edges, &s->edges2, &s->quads, &s->walk_lists, &s->mask_extra[0], &s->mask_extra[1], &s->mask_extra[2], &s->mask_extra[3], &s->mask_extra[4], &s->mask_extra[5], &s->clip_states, &s->clip_px, &s->clip_table, &s->clip_t2, &s->op_depth, &s->ord, &s->row_cnt, &s->item_cnt, &s->rows, &s->trow_op, &s->pool,
                 &s->counters, &s->mask0, &s->mask1, &s->zmask, &s->zplane_extra[0], &s->zplane_extra[1], &s->zplane_extra[2], &s->zplane_extra[3], &s->zplane_extra[4], &s->zplane_extra[5], &s->zplane_extra[6], &s->item_flags, &s->tile_cnt, &s->tile_fill, &s->cmds, &s->cmds_sorted,
                 &s->surfs, &s->surf_tile_base, &s->temp_px, &s->scan_tmp, &s->blur_jobs, &s->blur_rows, &s->blur_cols,
                 &s->blur_tmp_ptrs, &s->blur_tmp};
  for (Buf* b : bufs) buf_free(*b);
  if (s->remote_canvas) cudaIpcCloseMemHandle(s->remote_canvas);
  if (s->canvas) cudaFree(s->canvas);
  if (s->mapped_host) cudaFreeHost(s->mapped_host);
  for (int i = 0; i < 2; i++) {
    if (s->stage[i]) cudaFreeHost(s->stage[i]);
    if (s->stage_ev[i]) cudaEventDestroy(s->stage_ev[i]);
  }
  for (int i = 0; i < 12; i++)
    if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  for (int i = 0; i < 32; i++)
    if (s->ev_blur[i]) cudaEventDestroy(s->ev_blur[i]);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// x + w <= surface width and y + h <= surface height, without the uint32 wrap of the sums
static bool rect_inside(skb_surface s, uint32_t x, uint32_t y, uint32_t w, uint32_t h) {
  return x <= s->w && w <= s->w - x && y <= s->h && h <= s->h - y;
}

skb_result skb_surface_set_band(skb_surface s, uint32_t y0, uint32_t y1) {
  if (!s) return SKB_ERROR_INVALID_ARGUMENT;
  if (y1 == 0) {
    s->band_y0 = s->band_y1 = 0;
    return SKB_SUCCESS;
  }
  if (y0 >= y1 || y0 >= s->h || y0 % SKB_TILE != 0 || (y1 % SKB_TILE != 0 && y1 < s->h)) {
    set_error("band rows must be multiples of 16");
    return SKB_ERROR_INVALID_ARGUMENT;
  }
  s->band_y0 = y0;
  s->band_y1 = y1;
  return SKB_SUCCESS;
}

skb_result skb_surface_set_coord_mode(skb_surface s, int mode) {
  if (!s || mode < SKB_COORD_AUTO || mode > SKB_COORD_WIDE) return SKB_ERROR_INVALID_ARGUMENT;
  s->coord_mode = mode;
  return SKB_SUCCESS;
}

skb_result skb_surface_set_coverage_mode(skb_surface s, int mode) {
  if (!s || (mode != SKB_COVERAGE_EXACT && mode != SKB_COVERAGE_AREA)) return SKB_ERROR_INVALID_ARGUMENT;
  s->coverage_mode = mode;
  return SKB_SUCCESS;
}

skb_result skb_surface_set_walk_mode(skb_surface s, int mode) {
  if (!s || mode < 0 || mode > 1) return SKB_ERROR_INVALID_ARGUMENT;
  s->walk_mode = mode;
  return SKB_SUCCESS;
}

skb_result skb_frame_begin(skb_surface s, int clear) {
  if (!s) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  if (!clear && s->remote_canvas) {
    set_error("a band whose pixels are stored into another GPU's canvas starts from a cleared band: skb_frame_begin(clear = 0) "
              "would blend onto this device's copy, not onto the gathering canvas");
    return SKB_ERROR_INVALID_ARGUMENT;
  }
  if (clear) {
    // with a band set only the band's rows are this device's business: the rows of a gathering canvas that other
    // devices store into (skb_surface_set_remote_canvas on their side) must not be cleared under their feet
    size_t r0 = 0, r1 = (size_t)s->tiles_y * SKB_TILE;
    if (s->band_y1 > 0) {
      r0 = s->band_y0;
      r1 = ((size_t)(s->band_y1 < s->h ? s->band_y1 : s->h) + SKB_TILE - 1) / SKB_TILE * SKB_TILE;
    }
    SKB_CUDA(cudaMemsetAsync(s->canvas + r0 * s->pitch, 0, (r1 - r0) * s->pitch, s->stream));
  }
  // the last encoded display list stays resident: a frame may be flushed again without re-uploading it
  s->flushed = false;
  return SKB_SUCCESS;
}

skb_result skb_display_list_validate(const void* dl, size_t bytes) {
  if (!dl) return SKB_ERROR_INVALID_ARGUMENT;
  return validate_dl((const uint8_t*)dl, bytes);
}

// The part of a display list that can reach rows [row0, row1) of the canvas (surface 0): the list of one band of a
// canvas split over several GPUs.  Dropped are fills of the canvas whose path cannot reach those rows — by the test
// k_op_init applies on the device (control points' y range widened by half its height + 2 px), with one more pixel of
// margin, so every op dropped here is one the device would have culled anyway and the band's pixels are the same as with
// the whole list.  Everything else stays: clip paths, blurs, draws into other surfaces.  Paths, segments and paints of the
// dropped ops go too; what is kept keeps its order.
skb_result skb_display_list_cull_rows(const void* dl_v, size_t bytes, int32_t row0, int32_t row1, void* out_v, size_t out_capacity,
                                      size_t* out_bytes) {
  if (!dl_v || !out_bytes || row1 < row0) return SKB_ERROR_INVALID_ARGUMENT;
  const uint8_t* dl = (const uint8_t*)dl_v;
  skb_result vr = validate_dl(dl, bytes);
  if (vr != SKB_SUCCESS) return vr;
  skb_dl_header h;
  memcpy(&h, dl, sizeof(h));
  const skb_dl_surface* surfs = (const skb_dl_surface*)(dl + h.off_surfaces);
  const skb_dl_op* ops = (const skb_dl_op*)(dl + h.off_ops);
  const skb_dl_path* paths = (const skb_dl_path*)(dl + h.off_paths);
  const skb_dl_seg* segs = (const skb_dl_seg*)(dl + h.off_segs);
  const skb_dl_paint* paints = (const skb_dl_paint*)(dl + h.off_paints);
  auto align16 = [](uint64_t v) { return (v + 15) & ~(uint64_t)15; };
  const uint64_t tail_old = align16((uint64_t)h.off_stops + (uint64_t)h.n_stop_floats * 4);   // image pixels, if any
  for (uint32_t i = 0; i < h.n_surfaces; i++)
    if ((surfs[i].flags & SKB_SURFACE_IMAGE) && surfs[i].reserved < tail_old) {
      set_error("display list: image pixels before the end of the stop pool (not a layout this function rewrites)");
      return SKB_ERROR_UNSUPPORTED;
    }
  const bool canvas0 = !(surfs[0].flags & SKB_SURFACE_IMAGE);
  std::vector<uint8_t> keep(h.n_ops, 1);
  std::vector<uint32_t> paint_map(h.n_paints, 0xFFFFFFFFu);
  uint64_t n_ops = 0, n_paths = 0, n_segs = 0, n_paints = 0;
  const int n_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  {
    auto classify = [&](uint32_t a, uint32_t b) {
      for (uint32_t i = a; i < b; i++) {
        const skb_dl_op& o = ops[i];
        if (o.kind != SKB_OP_FILL || o.surface != 0 || !canvas0) continue;
        const skb_dl_path& p = paths[o.path];
        if (p.n_segs == 0) continue;
        float ymin = 3.0e38f, ymax = -3.0e38f;
        bool all_finite = true;
        for (uint32_t k = 0; k < p.n_segs; k++) {
          const uint32_t si = p.seg_off + k;
          const skb_dl_seg& sg = segs[si];
          const uint32_t type = sg.type_flags & SKB_SEG_TYPE_MASK;
          int last = 0;
          if (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) last = 1;
          else if (type == SKB_SEG_QUAD || type == SKB_SEG_CONIC) last = 2;
          else if (type == SKB_SEG_CUBIC) last = 3;
          const V2 st = xform(o.ctm, seg_start_point(segs, si));
          all_finite &= std::isfinite(st.y);
          ymin = std::min(ymin, st.y); ymax = std::max(ymax, st.y);
          for (int c = 1; c <= last; c++) {
            const V2 q = xform(o.ctm, v2(sg.p[2 * c], sg.p[2 * c + 1]));
            all_finite &= std::isfinite(q.y);
            ymin = std::min(ymin, q.y); ymax = std::max(ymax, q.y);
          }
        }
        if (!all_finite) continue;
        const float pad = 0.5f * (ymax - ymin) + 3.0f;
        if (ymax + pad < (float)row0 || ymin - pad > (float)row1) keep[i] = 0;
      }
    };
    std::vector<std::thread> th;
    const uint32_t per = (h.n_ops + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
      const uint32_t a = std::min(h.n_ops, (uint32_t)t * per), b = std::min(h.n_ops, a + per);
      if (a < b) th.emplace_back(classify, a, b);
    }
    for (auto& t : th) t.join();
  }
  for (uint32_t i = 0; i < h.n_ops; i++) {
    if (!keep[i]) continue;
    const skb_dl_op& o = ops[i];
    n_ops++;
    if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
      n_paths++;
      n_segs += paths[o.path].n_segs;
    }
    if (o.kind == SKB_OP_FILL && paint_map[o.paint] == 0xFFFFFFFFu) paint_map[o.paint] = (uint32_t)n_paints++;
  }
  skb_dl_header nh = h;
  nh.n_ops = (uint32_t)n_ops;
  nh.n_paths = (uint32_t)n_paths;
  nh.n_segs = (uint32_t)n_segs;
  nh.n_paints = (uint32_t)n_paints;
  uint64_t off = align16(sizeof(skb_dl_header));
  nh.off_surfaces = (uint32_t)off; off = align16(off + (uint64_t)h.n_surfaces * sizeof(skb_dl_surface));
  nh.off_ops = (uint32_t)off;      off = align16(off + n_ops * sizeof(skb_dl_op));
  nh.off_paths = (uint32_t)off;    off = align16(off + n_paths * sizeof(skb_dl_path));
  nh.off_segs = (uint32_t)off;     off = align16(off + n_segs * sizeof(skb_dl_seg));
  nh.off_paints = (uint32_t)off;   off = align16(off + n_paints * sizeof(skb_dl_paint));
  nh.off_stops = (uint32_t)off;    off = align16(off + (uint64_t)h.n_stop_floats * 4);
  const uint64_t tail_new = off;
  const uint64_t tail_bytes = h.total_bytes > tail_old ? h.total_bytes - tail_old : 0;
  off += tail_bytes;
  nh.total_bytes = (uint32_t)off;
  *out_bytes = (size_t)off;
  if (!out_v) return SKB_SUCCESS;   // size query
  if (out_capacity < off) {
    set_error("skb_display_list_cull_rows: output buffer too small");
    return SKB_ERROR_INVALID_ARGUMENT;
  }
  uint8_t* out = (uint8_t*)out_v;
  memset(out, 0, nh.off_ops);
  memcpy(out, &nh, sizeof(nh));
  skb_dl_surface* nsurfs = (skb_dl_surface*)(out + nh.off_surfaces);
  memcpy(nsurfs, surfs, (size_t)h.n_surfaces * sizeof(skb_dl_surface));
  for (uint32_t i = 0; i < h.n_surfaces; i++)
    if (nsurfs[i].flags & SKB_SURFACE_IMAGE) nsurfs[i].reserved = (uint32_t)(nsurfs[i].reserved - tail_old + tail_new);
  skb_dl_op* nops = (skb_dl_op*)(out + nh.off_ops);
  skb_dl_path* npaths = (skb_dl_path*)(out + nh.off_paths);
  skb_dl_seg* nsegs = (skb_dl_seg*)(out + nh.off_segs);
  skb_dl_paint* npaints = (skb_dl_paint*)(out + nh.off_paints);
  uint32_t wo = 0, wp = 0, ws = 0;
  for (uint32_t i = 0; i < h.n_ops; i++) {
    if (!keep[i]) continue;
    skb_dl_op o = ops[i];
    if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
      const skb_dl_path& p = paths[o.path];
      skb_dl_path np = p;
      np.seg_off = ws;
      memcpy(nsegs + ws, segs + p.seg_off, (size_t)p.n_segs * sizeof(skb_dl_seg));
      ws += p.n_segs;
      npaths[wp] = np;
      o.path = wp++;
    }
    if (o.kind == SKB_OP_FILL) {
      npaints[paint_map[o.paint]] = paints[o.paint];
      o.paint = paint_map[o.paint];
    }
    nops[wo++] = o;
  }
  memcpy(out + nh.off_stops, dl + h.off_stops, (size_t)h.n_stop_floats * 4);
  if (tail_bytes) memcpy(out + tail_new, dl + tail_old, (size_t)tail_bytes);
  return SKB_SUCCESS;
}

skb_result skb_frame_encode(skb_surface s, const void* dl, size_t bytes) {
  if (!s || !dl) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  skb_dl_header h;
  if (bytes >= sizeof(h)) {
    // the upload does not wait for the validation: the copy engine moves the list (456 MB for 1M paths) while this
    // thread reads the op table; nothing is launched on it before validate_dl has accepted it (skb_frame_flush)
    memcpy(&h, dl, sizeof(h));
    if (h.magic == SKB_DL_MAGIC && h.version == SKB_DL_VERSION && h.total_bytes <= bytes && h.total_bytes >= sizeof(h)) {
      SKB_TRY(buf_reserve(s->dl, h.total_bytes));
      SKB_CUDA(cudaMemcpyAsync(s->dl.p, dl, h.total_bytes, cudaMemcpyHostToDevice, s->stream));
    }
  }
  s->have_frame = false;
  SKB_TRY(validate_dl((const uint8_t*)dl, bytes, &s->plan));
  memcpy(&h, dl, sizeof(h));
  // the header and the surface table are also read on the host while launching
  s->host_dl.assign((const uint8_t*)dl, (const uint8_t*)dl + h.off_ops);
  s->have_frame = true;
  return SKB_SUCCESS;
}

skb_result skb_frame_flush(skb_surface s) {
  if (!s) return SKB_ERROR_INVALID_ARGUMENT;
  if (!s->have_frame) {
    set_error("skb_frame_flush without skb_frame_encode");
    return SKB_ERROR_INVALID_ARGUMENT;
  }
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  return run_frame(s);
}

skb_result skb_surface_sync(skb_surface s) {
  if (!s) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  return SKB_SUCCESS;
}

// Copies rows [r0, r1) of a staged chunk into the caller's buffer with a few host threads: the destination of a large
// read-back is usually memory nobody has touched yet (a fresh Pixmap), and one thread page-faulting its way through it
// is several times slower than the PCIe copy that feeds it.
static void copy_rows_parallel(uint8_t* dst, size_t dst_stride, const uint8_t* src, size_t src_stride, size_t row_bytes, uint32_t rows) {
  const size_t bytes = (size_t)rows * row_bytes;
  unsigned nt = bytes >= ((size_t)8 << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1u;
  auto work = [&](uint32_t a, uint32_t b) {
    if (dst_stride == row_bytes && src_stride == row_bytes) {
      memcpy(dst + (size_t)a * row_bytes, src + (size_t)a * row_bytes, (size_t)(b - a) * row_bytes);
    } else {
      for (uint32_t r = a; r < b; r++) memcpy(dst + (size_t)r * dst_stride, src + (size_t)r * src_stride, row_bytes);
    }
  };
  if (nt <= 1) {
    work(0, rows);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back(work, (uint32_t)((uint64_t)rows * t / nt), (uint32_t)((uint64_t)rows * (t + 1) / nt));
  for (auto& t : th) t.join();
}

#define SKB_STAGE_BYTES ((size_t)32 << 20)

skb_result skb_surface_read_pixels(skb_surface s, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void* dst, size_t stride) {
  if (!s || !dst || !rect_inside(s, x, y, w, h) || stride < (size_t)w * 4) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  const size_t row_bytes = (size_t)w * 4;
  // Pageable destination and a transfer worth the trouble: the driver would stage it through its own small bounce
  // buffer at a few GB/s.  Instead the rectangle travels in chunks through two page-locked buffers of the surface, the
  // copy of chunk k + 1 over PCIe overlapping the host copy of chunk k into the destination.
  cudaPointerAttributes pa;
  const bool pageable = cudaPointerGetAttributes(&pa, dst) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (pageable && row_bytes * h >= ((size_t)4 << 20) && row_bytes <= SKB_STAGE_BYTES) {
    for (int i = 0; i < 2; i++) {
      if (!s->stage[i]) SKB_CUDA(cudaHostAlloc((void**)&s->stage[i], SKB_STAGE_BYTES, cudaHostAllocDefault));
      if (!s->stage_ev[i]) SKB_CUDA(cudaEventCreateWithFlags(&s->stage_ev[i], cudaEventDisableTiming));
    }
    const uint32_t rows_per = (uint32_t)(SKB_STAGE_BYTES / row_bytes);
    uint32_t issued = 0, done = 0;
    int k_issue = 0, k_done = 0;
    uint32_t rows_of[2] = {0, 0}, row0_of[2] = {0, 0};
    while (done < h) {
      while (issued < h && k_issue - k_done < 2) {
        const int b = k_issue & 1;
        const uint32_t n = std::min(rows_per, h - issued);
        SKB_CUDA(cudaMemcpy2DAsync(s->stage[b], row_bytes, s->canvas + (size_t)(y + issued) * s->pitch + (size_t)x * 4, s->pitch,
                                   row_bytes, n, cudaMemcpyDeviceToHost, s->stream));
        SKB_CUDA(cudaEventRecord(s->stage_ev[b], s->stream));
        rows_of[b] = n;
        row0_of[b] = issued;
        issued += n;
        k_issue++;
      }
      const int b = k_done & 1;
      SKB_CUDA(cudaEventSynchronize(s->stage_ev[b]));
      copy_rows_parallel((uint8_t*)dst + (size_t)row0_of[b] * stride, stride, s->stage[b], row_bytes, row_bytes, rows_of[b]);
      done += rows_of[b];
      k_done++;
    }
    return SKB_SUCCESS;
  }
  SKB_CUDA(cudaMemcpy2DAsync(dst, stride, s->canvas + (size_t)y * s->pitch + (size_t)x * 4, s->pitch, row_bytes, h,
                             cudaMemcpyDeviceToHost, s->stream));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  return SKB_SUCCESS;
}

skb_result skb_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return SKB_ERROR_INVALID_ARGUMENT;
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(std::string("page-locked host allocation failed: ") + cudaGetErrorString(e));
    return SKB_ERROR_OUT_OF_MEMORY;
  }
  return SKB_SUCCESS;
}

void skb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

skb_result skb_surface_read_pixels_async(skb_surface s, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void* dst,
                                         size_t stride) {
  if (!s || !dst || !rect_inside(s, x, y, w, h) || stride < (size_t)w * 4) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaMemcpy2DAsync(dst, stride, s->canvas + (size_t)y * s->pitch + (size_t)x * 4, s->pitch, (size_t)w * 4, h,
                             cudaMemcpyDeviceToHost, s->stream));
  return SKB_SUCCESS;
}

skb_result skb_surface_export_canvas(skb_surface s, void* handle64) {
  if (!s || !handle64) return SKB_ERROR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI hands the IPC handle over as 64 bytes");
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  cudaIpcMemHandle_t hdl;
  SKB_CUDA(cudaIpcGetMemHandle(&hdl, s->canvas));
  memcpy(handle64, &hdl, 64);
  return SKB_SUCCESS;
}

skb_result skb_surface_set_remote_canvas(skb_surface s, const void* handle64) {
  if (!s) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  if (s->remote_canvas) {
    SKB_CUDA(cudaIpcCloseMemHandle(s->remote_canvas));
    s->remote_canvas = nullptr;
  }
  if (!handle64) return SKB_SUCCESS;
  cudaIpcMemHandle_t hdl;
  memcpy(&hdl, handle64, 64);
  void* p = nullptr;
  SKB_CUDA(cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess));
  s->remote_canvas = (uint8_t*)p;
  return SKB_SUCCESS;
}

skb_result skb_surface_write_pixels(skb_surface s, uint32_t x, uint32_t y, uint32_t w, uint32_t h, const void* src,
                                    size_t stride) {
  if (!s || !src || !rect_inside(s, x, y, w, h) || stride < (size_t)w * 4) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaMemcpy2DAsync(s->canvas + (size_t)y * s->pitch + (size_t)x * 4, s->pitch, src, stride, (size_t)w * 4, h,
                             cudaMemcpyHostToDevice, s->stream));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  return SKB_SUCCESS;
}

skb_result skb_frame_read_surface(skb_surface s, uint32_t index, void* dst, size_t stride) {
  if (!s || !dst || !s->flushed || index == 0 || index >= s->h_surfs.size()) return SKB_ERROR_INVALID_ARGUMENT;
  const SurfDesc& d = s->h_surfs[index];
  if (stride < (size_t)d.w * 4) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaMemcpy2DAsync(dst, stride, d.px, d.pitch, (size_t)d.w * 4, d.h, cudaMemcpyDeviceToHost, s->stream));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  return SKB_SUCCESS;
}

skb_result skb_surface_device_ptr(skb_surface s, void** out_ptr, size_t* out_pitch) {
  if (!s || !out_ptr) return SKB_ERROR_INVALID_ARGUMENT;
  *out_ptr = s->canvas;
  if (out_pitch) *out_pitch = s->pitch;
  return SKB_SUCCESS;
}

skb_result skb_surface_stream(skb_surface s, void** out_stream) {
  if (!s || !out_stream) return SKB_ERROR_INVALID_ARGUMENT;
  *out_stream = (void*)s->stream;
  return SKB_SUCCESS;
}

skb_result skb_frame_get_stats(skb_surface s, skb_frame_stats* out) {
  if (!s || !out) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  SKB_CUDA(cudaStreamSynchronize(s->stream));
  if (s->flushed && s->n_ops) {
    float ms = 0;
    // ev: 0 start, 1 after flatten, 2 after setup, 3 after walk, 4 after cover, 5 after bin, 8 after the last fine pass
    //     9 after k_cover (ev 3..9 = coverage, 9..4 = clip stage)
    static const int from_ev[7] = {0, 1, 2, 3, 9, 4, 5};
    static const int to_ev[7] = {1, 2, 3, 9, 4, 5, 8};
    static const int stage_of[7] = {0, 1, 2, 3, 7, 4, 5};
    for (int i = 0; i < 8; i++) s->stats.ms_stage[i] = 0;
    for (int i = 0; i < 7; i++) {
      if (cudaEventElapsedTime(&ms, s->ev[from_ev[i]], s->ev[to_ev[i]]) == cudaSuccess) s->stats.ms_stage[stage_of[i]] += ms;
    }
    // the blur sections sit inside ev 5..8: move their time from "fine" to "blur"
    for (uint32_t l = 0; l < s->n_levels_timed; l++) {
      if (cudaEventElapsedTime(&ms, s->ev_blur[2 * l], s->ev_blur[2 * l + 1]) == cudaSuccess) {
        s->stats.ms_stage[6] += ms;
        s->stats.ms_stage[5] -= ms;
      }
    }
    if (cudaEventElapsedTime(&ms, s->ev[0], s->ev[8]) == cudaSuccess) s->stats.ms_total = ms;
    cudaGetLastError();
  }
  *out = s->stats;
  return SKB_SUCCESS;
}

skb_result skb_debug_read_coverage(skb_surface s, uint32_t op, int32_t x, int32_t y, uint32_t w, uint32_t h, uint8_t* direct,
                                   uint8_t* accum) {
  if (!s || !direct || !accum || !s->flushed || op >= s->n_ops || w == 0 || h == 0) return SKB_ERROR_INVALID_ARGUMENT;
  SKB_CUDA(cudaSetDevice(s->dev->ordinal));
  uint8_t *dd = nullptr, *da = nullptr;
  SKB_CUDA(cudaMalloc((void**)&dd, (size_t)w * h));
  SKB_CUDA(cudaMalloc((void**)&da, (size_t)w * h));
  CoverArgs ca;
  memset(&ca, 0, sizeof(ca));
  ca.geom = (const OpGeom*)s->geom.p;
  ca.item_base = (const uint32_t*)s->item_cnt.p;
  ca.n_ops = s->n_ops;
  ca.n_items = s->n_items;
  ca.mask0 = (uint8_t*)s->mask0.p;
  ca.mask1 = (uint8_t*)s->mask1.p;
  ca.item_flags = (uint16_t*)s->item_flags.p;
  k_read_coverage<<<cdiv((uint64_t)w * h, 256), 256, 0, s->stream>>>(ca, op, x, y, w, h, dd, da);
  cudaError_t e = cudaMemcpyAsync(direct, dd, (size_t)w * h, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(accum, da, (size_t)w * h, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(dd);
  cudaFree(da);
  if (e != cudaSuccess) {
    set_error(std::string("read coverage: ") + cudaGetErrorString(e));
    return SKB_ERROR_CUDA;
  }
  return SKB_SUCCESS;
}

}  // extern "C"
