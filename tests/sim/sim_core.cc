// CPU simulation of the flatten -> setup -> walk -> coverage stages: runs the very same
// per-thread functions the CUDA kernels run (skity_b200/csrc/skb_*.cuh), one "thread" after
// the other, so kernel logic can be checked against the oracle where there is no GPU.
// Test infrastructure only; built by tests/simlib.py with g++.
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>

#include "skity_b200/csrc/skb_stages.cuh"
#include "tests/sim/walk_nested.hpp"

using namespace skb;

// SKB_SIM_WALK_MODE selects the sweep variant (skb_walk.cuh walk_path `mode`); default = what the GPU runs.
static int sim_walk_mode() {
  const char* m = getenv("SKB_SIM_WALK_MODE");
  return m ? atoi(m) : 2;
}


static int g_sim_wide = 0;  // wide-coordinate mode (include/skb.h SKB_COORD_WIDE)
extern "C" void sim_set_wide(int w) { g_sim_wide = w; }

extern "C" {

// Returns number of trapezoid records, or <0.  direct/accum are surf_w*surf_h planes (pre-zeroed
// by the caller).  stats[0]=n_prims stats[1]=n_edges stats[2]=pixels with both planes nonzero
// stats[3]=pixels where two direct rows overlapped.
long sim_path_cover(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd,
                    int surf_w, int surf_h, uint8_t* direct, uint8_t* accum, int64_t* stats) {
  std::vector<uint32_t> prim_off(n_segs + 1, 0);
  for (uint32_t i = 0; i < n_segs; i++) prim_off[i + 1] = prim_off[i] + (uint32_t)seg_prim_count(segs[i]);
  uint32_t n_prims = prim_off[n_segs];
  OpGeom g;
  std::memset(&g, 0, sizeof(g));
  g.bmin_x = g.bmin_y = INT_MAX;
  g.bmax_x = g.bmax_y = INT_MIN;
  bool have = false;
  auto bound = [&](V2 p) {
    have = true;
    int32_t kx = float_key(p.x), ky = float_key(p.y);
    if (kx < g.bmin_x) g.bmin_x = kx;
    if (kx > g.bmax_x) g.bmax_x = kx;
    if (ky < g.bmin_y) g.bmin_y = ky;
    if (ky > g.bmax_y) g.bmax_y = ky;
  };
  std::vector<Edge> E(2 + 2 * (size_t)n_prims);
  std::vector<QuadState> Q(E.size());
  std::memset(E.data(), 0, E.size() * sizeof(Edge));
  std::memset(Q.data(), 0, Q.size() * sizeof(QuadState));
  for (uint32_t i = 0; i < n_segs; i++) {
    if ((segs[i].type_flags & SKB_SEG_TYPE_MASK) == SKB_SEG_POINT) bound(xform(ctm, seg_start_point(segs, i)));
    int n = (int)(prim_off[i + 1] - prim_off[i]);
    for (int k = 0; k < n; k++) {
      V2 p[3];
      int np = seg_prim(segs, i, k, n, ctm, p);
      for (int j = 0; j < np; j++) bound(p[j]);
      flatten_prim(np, p, &E[2 + 2 * (size_t)(prim_off[i] + k)], &Q[2 + 2 * (size_t)(prim_off[i] + k)], g_sim_wide);
    }
  }
  op_setup(g, clip, (uint32_t)surf_w, (uint32_t)surf_h, have);
  stats[0] = n_prims;
  stats[1] = stats[2] = stats[3] = 0;
  for (size_t i = 2; i < E.size(); i++) stats[1] += (E[i].curve >> 24) & 1;
  if (g.empty) return 0;
  int n_rows = g.scan_b - g.scan_t;
  std::vector<uint2> rows((size_t)n_rows);
  std::memset(rows.data(), 0, rows.size() * sizeof(uint2));
  std::vector<TrapRec> pool((size_t)1 << 16);
  uint32_t pool_next = 0, overflow = 0;
  std::vector<int32_t> ord(E.size());
  for (;;) {
    std::vector<Edge> Ew = E;
    std::vector<QuadState> Qw = Q;
    RecSink sink;
    sink.pool = pool.data();
    sink.pool_next = &pool_next;
    sink.pool_cap = (uint32_t)pool.size();
    sink.overflow = &overflow;
    sink.rows = rows.data();
    sink.row0 = g.scan_t;
    sink.n_rows = n_rows;
    sink_init(sink);
    sim_walk_path(Ew.data(), Qw.data(), (int)Ew.size(), ord.data(), g.scan_top_f, g.scan_bottom_f, g.start_y, g.stop_y, g.left_clip,
              g.right_clip, even_odd, sink, sim_walk_mode(), g_sim_wide);
    if (!overflow) break;
    pool.resize(pool.size() * 4);
    pool_next = 0;
    overflow = 0;
    std::memset(rows.data(), 0, rows.size() * sizeof(uint2));
  }
  long n_recs = 0;
  for (int r = 0; r < n_rows; r++) n_recs += rows[r].y;
  for (int y = g.scan_t; y < g.scan_b; y++) {
    if (y < 0 || y >= surf_h) continue;
    uint2 row = rows[y - g.scan_t];
    if (row.y == 0) continue;
    for (int x = g.scan_l < 0 ? 0 : g.scan_l; x < g.scan_r && x < surf_w; x++) {
      PixelCover pc = cover_pixel(pool.data(), row, x);
      direct[(size_t)y * surf_w + x] = pc.direct;
      accum[(size_t)y * surf_w + x] = pc.accum;
      if (pc.direct && pc.accum) stats[2]++;
    }
  }
  return n_recs;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Whole-frame CPU simulation (fills, gradients, path clips; no blur): every op goes through the same
// per-thread stage functions the kernels use, pixels are composited in op order.
#include "skity_b200/csrc/skb_clip.cuh"

namespace {

struct SimOp {
  OpGeom g;
  std::vector<uint2> rows;
  std::vector<TrapRec> pool;
  bool ok = false;
};

void sim_raster_op(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd, int surf_w,
                   int surf_h, SimOp& out) {
  std::vector<uint32_t> prim_off(n_segs + 1, 0);
  for (uint32_t i = 0; i < n_segs; i++) prim_off[i + 1] = prim_off[i] + (uint32_t)seg_prim_count(segs[i]);
  uint32_t n_prims = prim_off[n_segs];
  OpGeom& g = out.g;
  std::memset(&g, 0, sizeof(g));
  g.bmin_x = g.bmin_y = INT_MAX;
  g.bmax_x = g.bmax_y = INT_MIN;
  bool have = false;
  auto bound = [&](V2 p) {
    have = true;
    int32_t kx = float_key(p.x), ky = float_key(p.y);
    if (kx < g.bmin_x) g.bmin_x = kx;
    if (kx > g.bmax_x) g.bmax_x = kx;
    if (ky < g.bmin_y) g.bmin_y = ky;
    if (ky > g.bmax_y) g.bmax_y = ky;
  };
  std::vector<Edge> E(2 + 2 * (size_t)n_prims);
  std::vector<QuadState> Q(E.size());
  std::memset(E.data(), 0, E.size() * sizeof(Edge));
  std::memset(Q.data(), 0, Q.size() * sizeof(QuadState));
  for (uint32_t i = 0; i < n_segs; i++) {
    if ((segs[i].type_flags & SKB_SEG_TYPE_MASK) == SKB_SEG_POINT) bound(xform(ctm, seg_start_point(segs, i)));
    int n = (int)(prim_off[i + 1] - prim_off[i]);
    for (int k = 0; k < n; k++) {
      V2 p[3];
      int np = seg_prim(segs, i, k, n, ctm, p);
      for (int j = 0; j < np; j++) bound(p[j]);
      flatten_prim(np, p, &E[2 + 2 * (size_t)(prim_off[i] + k)], &Q[2 + 2 * (size_t)(prim_off[i] + k)], g_sim_wide);
    }
  }
  op_setup(g, clip, (uint32_t)surf_w, (uint32_t)surf_h, have);
  out.ok = !g.empty;
  if (g.empty) return;
  int n_rows = g.scan_b - g.scan_t;
  out.rows.assign((size_t)n_rows, uint2{0, 0});
  out.pool.assign((size_t)1 << 14, TrapRec());
  uint32_t pool_next = 0, overflow = 0;
  std::vector<int32_t> ord(E.size());
  for (;;) {
    std::vector<Edge> Ew = E;
    std::vector<QuadState> Qw = Q;
    RecSink sink;
    sink.pool = out.pool.data();
    sink.pool_next = &pool_next;
    sink.pool_cap = (uint32_t)out.pool.size();
    sink.overflow = &overflow;
    sink.rows = out.rows.data();
    sink.row0 = g.scan_t;
    sink.n_rows = n_rows;
    sink_init(sink);
    sim_walk_path(Ew.data(), Qw.data(), (int)Ew.size(), ord.data(), g.scan_top_f, g.scan_bottom_f, g.start_y, g.stop_y,
              g.left_clip, g.right_clip, even_odd, sink, sim_walk_mode(), g_sim_wide);
    if (!overflow) break;
    out.pool.resize(out.pool.size() * 4);
    pool_next = 0;
    overflow = 0;
    std::fill(out.rows.begin(), out.rows.end(), uint2{0, 0});
  }
}

struct SimClipState {
  int rx0 = 0, ry0 = 0, rw = 0, rh = 0;
  std::vector<uint32_t> entries;  // rw*rh*MAXE
  bool nonempty = false;
};

}  // namespace

// How many bytes the sampler's u8 -> float -> u8 round trip changes (requant in skb_core.cuh): must be 0, the device
// code relies on it.
extern "C" int sim_requant_changes() {
  int bad = 0;
  for (uint32_t c = 0; c < 256; c++) bad += requant(c) != c;
  return bad;
}

extern "C" int sim_render_dl(const uint8_t* dl, size_t bytes, uint8_t* out_rgba, int64_t* stats) {
  (void)bytes;
  const skb_dl_header* h = (const skb_dl_header*)dl;
  const skb_dl_surface* sd = (const skb_dl_surface*)(dl + h->off_surfaces);
  const skb_dl_op* ops = (const skb_dl_op*)(dl + h->off_ops);
  const skb_dl_path* paths = (const skb_dl_path*)(dl + h->off_paths);
  const skb_dl_seg* segs = (const skb_dl_seg*)(dl + h->off_segs);
  const skb_dl_paint* paints = (const skb_dl_paint*)(dl + h->off_paints);
  const float* pool = (const float*)(dl + h->off_stops);
  if (h->n_surfaces != 1) return -10;  // blur temporaries are not simulated
  const int W = (int)sd[0].width, H = (int)sd[0].height;
  std::vector<uint32_t> canvas((size_t)W * H, 0u);
  std::vector<SimClipState> states(h->n_clip_states + 1);
  stats[0] = stats[1] = stats[2] = stats[3] = 0;  // [0] plane overflows, [1] most planes on a pixel, [3] clip_row_seek mismatches
  SurfaceView none;
  none.px = nullptr;
  none.w = none.h = none.pitch = 0;
  for (uint32_t i = 0; i < h->n_ops; i++) {
    const skb_dl_op& o = ops[i];
    if (o.kind != SKB_OP_FILL && o.kind != SKB_OP_CLIP) return -11;
    SimOp so;
    const skb_dl_path& p = paths[o.path];
    sim_raster_op(segs + p.seg_off, p.n_segs, o.ctm, o.clip_bounds, (int)o.fill_type, W, H, so);
    const SimClipState* parent = o.clip_in ? &states[o.clip_in] : nullptr;
    const bool clipped = parent && parent->nonempty;
    SimClipState* target = nullptr;
    if (o.kind == SKB_OP_CLIP) {
      if (o.aux != 1) return -12;  // only intersecting clips
      target = &states[o.clip_out];
      if (so.ok) {
        const OpGeom& g = so.g;
        // the whole scan rectangle, on the surface or not: HasClip() and nested clips see every span
        target->rx0 = g.scan_l;
        target->ry0 = g.scan_t;
        target->rw = g.scan_r + 1 - g.scan_l;
        target->rh = g.scan_b - g.scan_t;
        target->entries.assign((size_t)target->rw * target->rh * SKB_CLIP_MAXE, 0u);
      }
    }
    if (!so.ok) continue;
    const OpGeom& g = so.g;
    const skb_dl_paint* paint = o.kind == SKB_OP_FILL ? &paints[o.paint] : nullptr;
    const bool is_clip = o.kind == SKB_OP_CLIP;
    for (int y = g.scan_t; y < g.scan_b; y++) {
      if (!is_clip && (y < 0 || y >= H)) continue;
      uint2 row = so.rows[(size_t)(y - g.scan_t)];
      if (row.y == 0) continue;
      ClipRowState st;
      TrapPrep prep_storage[SKB_CLIP_RMAX];
      clip_row_begin(st, so.pool.data(), row, prep_storage);
      for (int x = g.scan_l; x <= g.scan_r && (is_clip || x < W); x++) {
        SpanSide ld, od, la, oa;
        // the GPU lets several threads share a row: a thread entering at x must reconstruct this very state
        // ... and looks only at the records that reach its run of pixels (clip_row_focus): here runs of 7
        if (st.n_prep >= 0 && (x - g.scan_l) % 7 == 0) clip_row_focus(st, x, x + 6);
        if (st.n_prep >= 0 && x > g.scan_l && (x - g.scan_l) % 7 == 0) {
          ClipRowState t = st;
          clip_row_unfocus(t);
          t.prev_d = t.prev_a = 0xDEAD;
          t.prev_d_start = t.prev_a_start = -12345;
          t.prev_d_ends = true;
          clip_row_seek(t, so.pool.data(), row, g.scan_l, x);
          const bool same = t.prev_d == st.prev_d && t.prev_a == st.prev_a && t.prev_d_ends == st.prev_d_ends &&
                            (st.prev_d == 0 || t.prev_d_start == st.prev_d_start) &&
                            (st.prev_a == 0 || t.prev_a_start == st.prev_a_start);
          if (!same) stats[3]++;
        }
        clip_row_step(st, so.pool.data(), row, x, ld, od, la, oa);
        if (!is_clip && x < 0) continue;
        const uint32_t* clist = nullptr;
        int n_c = 0;
        if (clipped && x >= parent->rx0 && x < parent->rx0 + parent->rw && y >= parent->ry0 && y < parent->ry0 + parent->rh) {
          clist = &parent->entries[((size_t)(y - parent->ry0) * parent->rw + (x - parent->rx0)) * SKB_CLIP_MAXE];
          while (n_c < SKB_CLIP_MAXE && clist[n_c]) n_c++;
        }
        const uint32_t* cprev = nullptr;
        int n_p = 0;
        if (is_clip && clipped && x - 1 >= parent->rx0 && x - 1 < parent->rx0 + parent->rw && y >= parent->ry0 && y < parent->ry0 + parent->rh) {
          cprev = &parent->entries[((size_t)(y - parent->ry0) * parent->rw + (x - 1 - parent->rx0)) * SKB_CLIP_MAXE];
          while (n_p < SKB_CLIP_MAXE && cprev[n_p]) n_p++;
        }
        ClipOut out;
        clip_combine(x, ld, od, la, oa, clist, n_c, cprev, n_p, clipped, is_clip ? SKB_CLIP_MAXE : SKB_CLIP_PLANES, is_clip ? 2 : 0, out);
        if (getenv("SKB_SIM_DEBUG_XY")) {
          int dx_ = 0, dy_ = 0;
          sscanf(getenv("SKB_SIM_DEBUG_XY"), "%d,%d", &dx_, &dy_);
          if (y == dy_ && x >= dx_ - 2 && x <= dx_ + 2)
            fprintf(stderr, "op %u kind %u x %d y %d: ld(%d,%u,%d) od(%d,%u,%d) la(%d,%u,%d) oa(%d,%u,%d) clipped %d n_c %d n_p %d out.n %d\n", i, o.kind, x, y,
                    (int)ld.present, ld.cover, ld.start, (int)od.present, od.cover, od.start, (int)la.present, la.cover, la.start,
                    (int)oa.present, oa.cover, oa.start, (int)clipped, n_c, n_p, out.n);
        }
        if (out.overflow) stats[0]++;
        if (out.n > stats[1]) stats[1] = out.n;
        if (o.kind == SKB_OP_CLIP) {
          if (x >= target->rx0 && x < target->rx0 + target->rw && y >= target->ry0 && y < target->ry0 + target->rh) {
            uint32_t* e = &target->entries[((size_t)(y - target->ry0) * target->rw + (x - target->rx0)) * SKB_CLIP_MAXE];
            for (int k = 0; k < out.n; k++) e[k] = out.e[k];
            if (out.n) target->nonempty = true;
          }
        } else {
          for (int k = 0; k < out.n; k++) {
            uint32_t cv = clip_entry_cover(out.e[k]);
            uint32_t galpha = paint->type == SKB_PAINT_IMAGE ? (paint->global_alpha & 0xFF) : 0xFFu;
            cv &= galpha;
            if (cv) {
              uint32_t src = swap_rb(paint_color(*paint, pool, none, x, y));
              if (cv != 255) src = alpha_mul_q(src, cv);
              if (SKB_PAINT_CF_OFFSET(*paint)) src = apply_color_filter((const uint32_t*)pool + (SKB_PAINT_CF_OFFSET(*paint) - 1), src);
              canvas[(size_t)y * W + x] = swap_rb(porter_duff(src, swap_rb(canvas[(size_t)y * W + x]), paint_blend_mode(*paint)));
            }
          }
        }
      }
    }
  }
  std::memcpy(out_rgba, canvas.data(), (size_t)W * H * 4);
  return 0;
}


// Number of inputs (out of n pseudo-random (y, x) pairs) on which skb_atan2f differs from the C library's atan2f.
extern "C" long sim_atan2f_mismatches(long n, unsigned long long seed) {
  unsigned long long s = seed ? seed : 88172645463325252ull;
  long bad = 0;
  for (long i = 0; i < n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    float a = (float)((double)(s & 0xFFFFFF) / 0xFFFFFF * 8000.0 - 4000.0);
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    float b = (float)((double)(s & 0xFFFFFF) / 0xFFFFFF * 8000.0 - 4000.0);
    if (i % 3 == 0) {  // a third of the samples with widely different magnitudes
      a = ldexpf(a, (int)((s >> 40) % 80) - 40);
      b = ldexpf(b, (int)((s >> 48) % 80) - 40);
    }
    float r = atan2f(a, b), m = skb::skb_atan2f(a, b);
    if (memcmp(&r, &m, 4) != 0) bad++;
  }
  return bad;
}

// Number of inputs on which fx_div (FP64 with an exact correction step) differs from the integer form of SWFixedDiv:
// n pseudo-random operand pairs of every magnitude, plus the corner values against each other.
extern "C" long sim_fx_div_mismatches(long n, unsigned long long seed) {
  unsigned long long s = seed ? seed : 88172645463325252ull;
  long bad = 0;
  const int32_t corner[] = {1, -1, 2, -2, 3, 255, 256, 65535, 65536, 65537, -65536, 0x7FFF, 0x8000, 0x10000, 0xFFFFF, 0x100000,
                            0x1FFFFF, 0x200000, 0x7FFFFFFF, -0x7FFFFFFF, (int32_t)0x80000000, 0x40000000, -0x40000000, 0x7FFFFFFE,
                            0x55555555, 0x33333333, 46341, 46340, 92681, 1000003};
  const int nc = (int)(sizeof(corner) / sizeof(corner[0]));
  for (int i = 0; i < nc; i++)
    for (int j = 0; j < nc; j++)
      for (int z = 0; z < 2; z++) {
        const int32_t a = z ? 0 : corner[i], b = corner[j];
        if (skb::fx_div(a, b) != skb::fx_div_int(a, b)) bad++;
      }
  for (long i = 0; i < n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    int32_t a = (int32_t)(s >> 32) >> (int)(s & 31);
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    int32_t b = (int32_t)(s >> 32) >> (int)(s & 31);
    if (b == 0) b = 1;
    if (i % 5 == 0) a = (int32_t)((int64_t)b * (int32_t)((s >> 8) & 0xFFFF) >> 16);   // near-exact multiples
    if (skb::fx_div(a, b) != skb::fx_div_int(a, b)) bad++;
  }
  return bad;
}

// ---------------------------------------------------------------------------------------------
// Row-parallel walk (skb_rowwalk.cuh) against the sequential sweep (skb_walk.cuh), record by record.
#include "skity_b200/csrc/skb_rowwalk.cuh"

namespace {
struct RwSimPath {
  OpGeom g;
  std::vector<Edge> E;
  std::vector<QuadState> Q;
  bool ok = false;
};

void rw_sim_flatten(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int surf_w, int surf_h, RwSimPath& out) {
  std::vector<uint32_t> prim_off(n_segs + 1, 0);
  for (uint32_t i = 0; i < n_segs; i++) prim_off[i + 1] = prim_off[i] + (uint32_t)seg_prim_count(segs[i]);
  uint32_t n_prims = prim_off[n_segs];
  OpGeom& g = out.g;
  std::memset(&g, 0, sizeof(g));
  g.bmin_x = g.bmin_y = INT_MAX;
  g.bmax_x = g.bmax_y = INT_MIN;
  bool have = false;
  auto bound = [&](V2 p) {
    have = true;
    int32_t kx = float_key(p.x), ky = float_key(p.y);
    if (kx < g.bmin_x) g.bmin_x = kx;
    if (kx > g.bmax_x) g.bmax_x = kx;
    if (ky < g.bmin_y) g.bmin_y = ky;
    if (ky > g.bmax_y) g.bmax_y = ky;
  };
  out.E.assign(2 + 2 * (size_t)n_prims, Edge());
  out.Q.assign(out.E.size(), QuadState());
  std::memset(out.E.data(), 0, out.E.size() * sizeof(Edge));
  std::memset(out.Q.data(), 0, out.Q.size() * sizeof(QuadState));
  for (uint32_t i = 0; i < n_segs; i++) {
    if ((segs[i].type_flags & SKB_SEG_TYPE_MASK) == SKB_SEG_POINT) bound(xform(ctm, seg_start_point(segs, i)));
    int n = (int)(prim_off[i + 1] - prim_off[i]);
    for (int k = 0; k < n; k++) {
      V2 p[3];
      int np = seg_prim(segs, i, k, n, ctm, p);
      for (int j = 0; j < np; j++) bound(p[j]);
      flatten_prim(np, p, &out.E[2 + 2 * (size_t)(prim_off[i] + k)], &out.Q[2 + 2 * (size_t)(prim_off[i] + k)], g_sim_wide);
    }
  }
  op_setup(g, clip, (uint32_t)surf_w, (uint32_t)surf_h, have);
  out.ok = !g.empty;
}
}  // namespace

// Returns 0: the row-parallel form produced exactly the sequential sweep's records; 1: it flagged the path for the
// sequential fallback; 2: MISMATCH (a bug); 3: path empty.  stats[0] = records, [1] = walk rows, [2] = chords, [3] = fail stage.
extern "C" int sim_rowwalk_check(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd,
                                 int surf_w, int surf_h, int64_t* stats) {
  RwSimPath P;
  rw_sim_flatten(segs, n_segs, ctm, clip, surf_w, surf_h, P);
  stats[0] = stats[1] = stats[2] = stats[3] = 0;
  if (!P.ok) return 3;
  const OpGeom& g = P.g;
  // ---- sequential sweep (reference for this check)
  const int n_rows = g.scan_b - g.scan_t;
  std::vector<uint2> rows((size_t)n_rows, uint2{0, 0});
  std::vector<TrapRec> pool((size_t)1 << 16);
  {
    uint32_t pool_next = 0, overflow = 0;
    std::vector<int32_t> ord(P.E.size());
    for (;;) {
      std::vector<Edge> Ew = P.E;
      std::vector<QuadState> Qw = P.Q;
      RecSink sink;
      sink.pool = pool.data();
      sink.pool_next = &pool_next;
      sink.pool_cap = (uint32_t)pool.size();
      sink.overflow = &overflow;
      sink.rows = rows.data();
      sink.row0 = g.scan_t;
      sink.n_rows = n_rows;
      sink_init(sink);
      sim_walk_path(Ew.data(), Qw.data(), (int)Ew.size(), ord.data(), g.scan_top_f, g.scan_bottom_f, g.start_y, g.stop_y, g.left_clip,
                g.right_clip, even_odd, sink, 1, g_sim_wide);
      if (!overflow) break;
      pool.resize(pool.size() * 4);
      pool_next = 0;
      overflow = 0;
      std::fill(rows.begin(), rows.end(), uint2{0, 0});
    }
  }
  // ---- row-parallel form, kernel by kernel
  const int n_slots = (int)P.E.size();
  const int origin = g.start_y;
  const int n_wrows = g.stop_y - g.start_y;
  if (n_wrows <= 0 || n_wrows * 4 > SKB_RW_MAXQ) { stats[3] = 1; return 1; }
  const int stop_q = n_wrows * 4;
  stats[1] = n_wrows;
  std::vector<SlotInfo> slots((size_t)n_slots);
  std::vector<uint32_t> cap((size_t)n_slots, 0), base((size_t)n_slots + 1, 0);
  for (int s = 2; s < n_slots; s++) {
    const Edge& e = P.E[s];
    if (!((e.curve >> 24) & 1)) continue;
    const bool quad = (e.curve >> 25) & 1;
    cap[s] = quad ? 1u + (uint32_t)edge_count(e) : 1u;
  }
  for (int s = 0; s < n_slots; s++) base[s + 1] = base[s] + cap[s];
  std::vector<Chord> chords(base[n_slots] + 1);
  std::vector<uint32_t> ev_words((size_t)(n_wrows + 4) / 4 + 1, 0u);
  int y0q = INT_MAX;
  // K2: chord y's and events
  for (int s = 0; s < n_slots; s++) {
    slots[s] = SlotInfo{base[s], 0u, 0, 0};
    if (!cap[s]) continue;
    const Edge& e = P.E[s];
    const bool quad = (e.curve >> 25) & 1;
    const fx y0 = quad ? e.prev : e.upper_y, y1 = quad ? e.next : e.lower_y;
    if (can_be_ignored(g.scan_top_f, g.scan_bottom_f, y0, y1, g_sim_wide)) continue;
    const int n = rw_chain(e, P.Q[s], origin, stop_q, nullptr, &chords[base[s]], (int)cap[s], ev_words.data());
    if (n <= 0) { stats[3] = 2; return 1; }
    slots[s].n_chords = (uint32_t)n;
    slots[s].q_first = chord_uq(chords[base[s]].yy);
    slots[s].q_last = chord_lq(chords[base[s] + n - 1].yy);
    if (slots[s].q_first < y0q) y0q = slots[s].q_first;
    stats[2] += n;
  }
  if (y0q == INT_MAX) {  // every edge culled: the sweep has nothing to do (walk_prologue returns false)
    for (int r = 0; r < n_rows; r++)
      if (rows[(size_t)r].y) return 2;
    return 0;
  }
  const uint8_t* ev = reinterpret_cast<const uint8_t*>(ev_words.data());
  std::vector<RowBand> tab((size_t)n_wrows + 1);
  std::vector<uint32_t> res((size_t)n_wrows, 0u), rec_off((size_t)n_wrows, 0u);
  rw_bands_from_events(ev, n_wrows, y0q, tab.data());
  RwRowIn in;
  in.slots = slots.data();
  in.n_slots = n_slots;
  in.chords = chords.data();
  in.tab = tab.data();
  in.ev = ev;
  in.y0q = y0q;
  in.stop_q = stop_q;
  in.origin_fx = i_to_fx(origin);
  in.left_clip = g.left_clip;
  in.right_clip = g.right_clip;
  in.even_odd = even_odd;
  in.rank = nullptr;
  uint32_t total = 0;
  std::vector<TrapRec> recs;
  std::vector<uint16_t> rank((size_t)n_slots, 0);
  std::vector<int32_t> ord2((size_t)n_slots, 0);
  int max_rounds = getenv("SKB_SIM_RW_ROUNDS") ? atoi(getenv("SKB_SIM_RW_ROUNDS")) : 3;
  int round = 0;
  int why = 0;
  for (;; round++) {
    if (round >= max_rounds) { stats[3] = why; return 1; }
    if (round == 1) {  // retry: with the sort ranks, and from the tables of the first attempt
      rw_sort_ranks(P.E.data(), n_slots, ord2.data(), g.scan_top_f, g.scan_bottom_f, g_sim_wide, rank.data());
      in.rank = rank.data();
    }
    why = 0;
    // K4: chain with the current tables
    for (int s = 0; s < n_slots && !why; s++) {
      if (!slots[s].n_chords) continue;
      const int n = rw_chain(P.E[s], P.Q[s], origin, stop_q, tab.data(), &chords[base[s]], (int)cap[s], nullptr);
      if (n != (int)slots[s].n_chords) why = 4;
    }
    if (why) { stats[3] = why; return 1; }   // structural: retrying does not help
    // K5: both hypotheses per row
    in.exact = false;
    for (int r = 0; r < n_wrows; r++) {
      in.row = r;
      RwRowOut o0, o1;
      rw_row(in, false, nullptr, 0, o0);
      rw_row(in, true, nullptr, 0, o1);
      if (o0.n_recs > 255) o0.fail = SKB_RWF_EMIT;
      if (o1.n_recs > 255) o1.fail = SKB_RWF_EMIT;
      if (getenv("SKB_SIM_VERBOSE") && (o0.fail || o1.fail)) fprintf(stderr, "  round %d row %d fail codes %d %d\n", round, r, o0.fail, o1.fail);
      res[r] = rw_pack(o0.f_ins, o0.f_surv, o1.f_surv, o0.mask, o1.mask, o0.fail != 0, o1.fail != 0, o0.n_recs & 255, o1.n_recs & 255);
    }
    // K6
    total = rw_bands_resolve(res.data(), n_wrows, y0q, tab.data(), rec_off.data());
    if (total == 0xFFFFFFFFu) {  // a row gave up: the tables it left are no basis for the retry
      why = 6;
      rw_bands_from_events(ev, n_wrows, y0q, tab.data());
      continue;
    }
    // K4 again with the final tables
    for (int s = 0; s < n_slots && !why; s++) {
      if (!slots[s].n_chords) continue;
      const int n = rw_chain(P.E[s], P.Q[s], origin, stop_q, tab.data(), &chords[base[s]], (int)cap[s], nullptr);
      if (n != (int)slots[s].n_chords) why = 5;
    }
    if (why) { stats[3] = why; return 1; }
    // K7: final rows
    recs.assign((size_t)total + 1, TrapRec());
    in.exact = true;
    for (int r = 0; r < n_wrows && !why; r++) {
      in.row = r;
      const int n_alloc = (int)((r + 1 < n_wrows ? rec_off[r + 1] : total) - rec_off[r]);
      RwRowOut o;
      rw_row(in, (tab[r].fl & SKB_RB_FIN) != 0, recs.data() + rec_off[r], n_alloc, o);
      const bool in_rows = r * 4 + 4 > y0q;
      if (o.fail) {
        why = 70 + o.fail;
        if (getenv("SKB_SIM_VERBOSE")) fprintf(stderr, "  round %d final row %d fail code %d (fin %d mask tab %u)\n", round, r, o.fail, (int)(tab[r].fl & 1), tab[r].mask);
        break;
      }
      if (in_rows && (o.mask != tab[r].mask || o.f_surv != ((tab[r].fl & SKB_RB_FSURV) != 0) || o.f_ins != ((tab[r].fl & SKB_RB_FINS) != 0) ||
                      o.n_recs != n_alloc)) {
        why = 8;
        if (getenv("SKB_SIM_VERBOSE"))
          fprintf(stderr, "  round %d verify row %d: mask %u/%u fsurv %d/%d fins %d/%d recs %d/%d fin %d res %08x\n", round, r, o.mask, tab[r].mask, (int)o.f_surv,
                  (int)((tab[r].fl & SKB_RB_FSURV) != 0), (int)o.f_ins, (int)((tab[r].fl & SKB_RB_FINS) != 0), o.n_recs, n_alloc,
                  (int)(tab[r].fl & SKB_RB_FIN), res[r]);
      }
    }
    if (!why) break;
  }
  stats[3] = round;
  // ---- compare with the sequential sweep's records, row by row, in order
  stats[0] = total;
  for (int r = 0; r < n_wrows; r++) {
    const int y = origin + r;
    const int rel = y - g.scan_t;
    const int n_alloc = (int)((r + 1 < n_wrows ? rec_off[r + 1] : total) - rec_off[r]);
    if (rel < 0 || rel >= n_rows) continue;   // rows above the scan rectangle are swept but emit nothing
    const uint2 row = rows[(size_t)rel];
    if ((int)row.y != n_alloc) {
      if (getenv("SKB_SIM_VERBOSE")) fprintf(stderr, "row %d (y %d): %d records, sequential %u\n", r, y, n_alloc, row.y);
      return 2;
    }
    uint32_t idx = row.x;
    for (uint32_t k = 0; k < row.y; k++, idx++) {
      TrapRec a = pool[idx];
      if (a.flags & SKB_REC_LINK) {
        idx = (uint32_t)a.y;
        a = pool[idx];
      }
      const TrapRec& b = recs[rec_off[r] + k];
      if (std::memcmp(&a, &b, sizeof(TrapRec)) != 0) {
        if (getenv("SKB_SIM_VERBOSE"))
          fprintf(stderr, "row %d (y %d) rec %u: seq y %d ul %d ur %d ll %d lr %d ldy %d rdy %d fl %x | row y %d ul %d ur %d ll %d lr %d ldy %d rdy %d fl %x\n",
                  r, y, k, a.y, a.ul, a.ur, a.ll, a.lr, a.ldy, a.rdy, a.flags, b.y, b.ul, b.ur, b.ll, b.lr, b.ldy, b.rdy, b.flags);
        return 2;
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Coverage mode AREA (skb_area.cuh): the device stages k_area_seg_count / k_area_bin / k_area_backdrop / k_area_cover
// on one path, thread by thread in an arbitrary (here: reversed) order — binning must not depend on it.
#include <algorithm>
#include <map>

#include "skity_b200/csrc/skb_area.cuh"

namespace {
struct SimAreaLine { uint32_t w0, w1, key; };
struct SimAreaSink {
  int tx0, ty0, ntx, nty;
  uint32_t key;
  std::vector<std::vector<SimAreaLine>>* items;
  std::vector<int>* local;
  std::vector<int>* delta;
  std::vector<int>* row_backdrop;
  void line(V2 p, V2 q, int tx, int ty, int aux) {
    const int dx = tx - tx0, dy = ty - ty0;
    if (dx < 0 || dy < 0 || dx >= ntx || dy >= nty) return;
    uint32_t w0 = 0, w1 = 0;
    int lc = 0;
    const int kind = area_tile_line(p, q, tx, ty, &w0, &w1, &lc);
    if (kind == 0) return;
    const size_t item = (size_t)dy * ntx + dx;
    if (kind == 2) { (*local)[item] += lc; return; }
    (*items)[item].push_back(SimAreaLine{w0, w1, key | (uint32_t)aux});
  }
  void backdrop(int tx, int ty, int d) {
    const int dx = tx - tx0, dy = ty - ty0;
    if (dy < 0 || dy >= nty || dx >= ntx) return;
    if (dx < 0) (*row_backdrop)[dy] += d;
    else (*delta)[(size_t)dy * ntx + dx] += d;
  }
};
}  // namespace

extern "C" {
// cover: surf_w * surf_h bytes, pre-zeroed.  stats[0] = lines, stats[1] = binned (line, tile) pairs.  Returns 0.
int sim_area_cover(const skb_dl_seg* segs, uint32_t n_segs, const float* ctm, const float* clip, int even_odd, int surf_w,
                   int surf_h, uint8_t* cover, int64_t* stats) {
  OpGeom g;
  std::memset(&g, 0, sizeof(g));
  g.bmin_x = g.bmin_y = INT_MAX;
  g.bmax_x = g.bmax_y = INT_MIN;
  bool have = false;
  for (uint32_t i = 0; i < n_segs; i++) {   // k_op_init, area branch (+ the POINT rule)
    const uint32_t type = segs[i].type_flags & SKB_SEG_TYPE_MASK;
    int last = (type == SKB_SEG_LINE || type == SKB_SEG_CLOSE) ? 1 : type == SKB_SEG_CUBIC ? 3 : 2;
    V2 pts[4];
    int np = 0;
    if (type == SKB_SEG_POINT) pts[np++] = xform(ctm, seg_start_point(segs, i));
    else for (int k = 0; k <= last; k++) pts[np++] = xform(ctm, v2(segs[i].p[2 * k], segs[i].p[2 * k + 1]));
    for (int k = 0; k < np; k++) {
      have = true;
      const int32_t kx = float_key(pts[k].x), ky = float_key(pts[k].y);
      g.bmin_x = std::min(g.bmin_x, kx); g.bmax_x = std::max(g.bmax_x, kx);
      g.bmin_y = std::min(g.bmin_y, ky); g.bmax_y = std::max(g.bmax_y, ky);
    }
  }
  op_setup(g, clip, (uint32_t)surf_w, (uint32_t)surf_h, have);
  stats[0] = stats[1] = 0;
  if (g.empty || g.ntx <= 0 || g.nty <= 0) return 0;
  std::vector<uint32_t> line_off(n_segs + 1, 0);
  for (uint32_t i = 0; i < n_segs; i++) line_off[i + 1] = line_off[i] + (uint32_t)area_seg_line_count(segs[i], ctm);
  const uint32_t n_lines = line_off[n_segs];
  stats[0] = n_lines;
  std::vector<std::vector<SimAreaLine>> items((size_t)g.ntx * g.nty);
  std::vector<int> local(items.size(), 0), delta(items.size(), 0), row_backdrop((size_t)g.nty, 0);
  for (uint32_t r = 0; r < n_lines; r++) {
    const uint32_t ln = n_lines - 1 - r;   // any order
    const uint32_t seg = (uint32_t)(std::upper_bound(line_off.begin(), line_off.end(), ln) - line_off.begin()) - 1;
    V2 from, to;
    area_seg_line(segs[seg], ctm, (int)(ln - line_off[seg]), (int)(line_off[seg + 1] - line_off[seg]), &from, &to);
    if (!(finite_f(from.x) && finite_f(from.y) && finite_f(to.x) && finite_f(to.y))) continue;
    SimAreaSink sink{g.tx0, g.ty0, g.ntx, g.nty, ln << 1, &items, &local, &delta, &row_backdrop};
    area_walk_line(from, to, sink);
  }
  const int xmin = std::max(g.scan_l, 0), xmax = std::min(g.scan_r, surf_w);
  const int ymin = std::max(g.scan_t, 0), ymax = std::min(g.scan_b, surf_h);
  for (int tr = 0; tr < g.nty; tr++) {
    int acc = row_backdrop[tr];
    for (int txi = 0; txi < g.ntx; txi++) {
      const size_t item = (size_t)tr * g.ntx + txi;
      const int backdrop = acc + local[item];
      acc += delta[item];
      auto& v = items[item];
      stats[1] += (int64_t)v.size();
      std::sort(v.begin(), v.end(), [](const SimAreaLine& a, const SimAreaLine& b) { return a.key < b.key; });
      std::vector<uint32_t> words(v.size() * 2 + 2);
      for (size_t k = 0; k < v.size(); k++) { words[2 * k] = v[k].w0; words[2 * k + 1] = v[k].w1; }
      for (int py = 0; py < 16; py++) {
        const int y = (g.ty0 + tr) * 16 + py;
        if (y < ymin || y >= ymax) continue;
        for (int px = 0; px < 16; px++) {
          const int x = (g.tx0 + txi) * 16 + px;
          if (x < xmin || x >= xmax) continue;
          cover[(size_t)y * surf_w + x] = (uint8_t)area_pixel(words.data(), (int)v.size(), 2, backdrop, even_odd, px, py);
        }
      }
    }
  }
  return 0;
}
}
