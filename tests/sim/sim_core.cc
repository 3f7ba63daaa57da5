// CPU simulation of the flatten -> setup -> walk -> coverage stages: runs the very same
// per-thread functions the CUDA kernels run (skity_b200/csrc/skb_*.cuh), one "thread" after
// the other, so kernel logic can be checked against the oracle where there is no GPU.
// Test infrastructure only; built by tests/simlib.py with g++.
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "skity_b200/csrc/skb_stages.cuh"

using namespace skb;

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
      flatten_prim(np, p, &E[2 + 2 * (size_t)(prim_off[i] + k)], &Q[2 + 2 * (size_t)(prim_off[i] + k)]);
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
    walk_path(Ew.data(), Qw.data(), nullptr, (int)Ew.size(), ord.data(), g.scan_top_f, g.scan_bottom_f, g.start_y, g.stop_y, g.left_clip,
              g.right_clip, even_odd, sink);
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
