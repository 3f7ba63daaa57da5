// skb_walk.cuh — stage "walk": one thread sweeps the active-edge list of one path
// and emits one TrapRec per (band, inside interval), i.e. per blit_trapezoid_row
// call the reference's WalkEdges would make (src/render/sw/sw_raster.cc:546-677).
//
// The sweep is inherently sequential in y (band heights and every edge's x are
// functions of the whole history), so the parallelism of this stage is across
// paths; all per-pixel work is deferred to the coverage stage, which consumes the
// records in parallel.  Pure per-thread code: also compiled by g++ for the CPU
// simulation used by the tests.
#ifndef SKB_WALK_CUH
#define SKB_WALK_CUH

#include "skity_b200/csrc/skb_core.cuh"

namespace skb {

#define SKB_CHUNK 32u  // records per pool chunk; the last slot of a chunk links to the next chunk

#if defined(__CUDA_ARCH__)
#define SKB_ATOMIC_ADD_U32(p, v) atomicAdd((p), (v))
#else
#define SKB_ATOMIC_ADD_U32(p, v) skb_host_fetch_add((p), (v))
SKB_HD uint32_t skb_host_fetch_add(uint32_t* p, uint32_t v) {
  uint32_t o = *p;
  *p = o + v;
  return o;
}
#endif

// Records and row entries are written once by the sweep and read once, a whole stage later, by the coverage kernels:
// they are stored with the streaming (evict-first) hint so that they do not push the sweep's own working set — the
// edges of the paths in flight, re-read and re-written every band — out of the L2 (C4a: 27.9 -> 26.4 ms).
#if defined(__CUDA_ARCH__) && !defined(SKB_WALK_NO_STREAM)
#define SKB_STORE_REC(p, r)                                                                                      \
  do {                                                                                                           \
    __stcs(reinterpret_cast<uint4*>(p), make_uint4((uint32_t)(r).y, (uint32_t)(r).ul, (uint32_t)(r).ur, (uint32_t)(r).ll)); \
    __stcs(reinterpret_cast<uint4*>(p) + 1, make_uint4((uint32_t)(r).lr, (uint32_t)(r).ldy, (uint32_t)(r).rdy, (r).flags)); \
  } while (0)
#define SKB_STORE_ROW(p, e) __stcs((p), (e))
#else
#define SKB_STORE_REC(p, r) (*(p) = (r))
#define SKB_STORE_ROW(p, e) (*(p) = (e))
#endif

// Where a walking thread puts its records.  rows[] has one (first index, count) pair per scan
// row of the path, rows processed top-down.
struct RecSink {
  TrapRec* pool;
  uint32_t* pool_next;  // bump allocator shared by all paths (in records)
  uint32_t pool_cap;
  uint32_t* overflow;   // set to 1 when the pool is exhausted
  uint2* rows;          // this path's row table
  int row0;             // first scan row (scan_top)
  int n_rows;
  // thread-private cursor
  uint32_t cur;
  uint32_t left;
  int cur_row;          // relative row of the records being written, -1 = none
  uint32_t row_first;   // first record and record count of that row (flushed to rows[] when it ends)
  uint32_t row_count;
};

SKB_HD void sink_init(RecSink& s) {
  s.cur = 0xFFFFFFFFu;
  s.left = 0;
  s.cur_row = -1;
  s.row_first = 0;
  s.row_count = 0;
}

// Writes the finished row's (first record, count) entry; called when the sweep moves to another row
// and once more after the sweep.
SKB_HD void sink_flush_row(RecSink& s) {
  if (s.cur_row >= 0 && s.row_count > 0) {
    uint2 e;
    e.x = s.row_first;
    e.y = s.row_count;
    SKB_STORE_ROW(&s.rows[s.cur_row], e);
  }
}

SKB_HDN void sink_emit(RecSink& s, const TrapRec& r) {
  int rel = r.y - s.row0;
  if (rel < 0 || rel >= s.n_rows) return;  // RealSpanBuilder/SpanBuilder drop rows above scan_bounds.Top()
  if (s.left == 0) {
    uint32_t c = SKB_ATOMIC_ADD_U32(s.pool_next, SKB_CHUNK);
    if (c + SKB_CHUNK > s.pool_cap) {
      *s.overflow = 1;
      return;
    }
    if (s.cur != 0xFFFFFFFFu) {
      TrapRec link;
      link.y = (int32_t)c;
      link.ul = link.ur = link.ll = link.lr = link.ldy = link.rdy = 0;
      link.flags = SKB_REC_LINK;
      s.pool[s.cur] = link;
    }
    s.cur = c;
    s.left = SKB_CHUNK - 1;
  }
  if (rel != s.cur_row) {
    sink_flush_row(s);
    s.cur_row = rel;
    s.row_first = s.cur;
    s.row_count = 0;
  }
  SKB_STORE_REC(&s.pool[s.cur], r);
  s.cur++;
  s.left--;
  s.row_count++;
}

// ---- SortEdges: libstdc++ std::sort (GCC 13 bits/stl_algo.h) on an index array --------------
// The reference sorts with std::sort by (upper_y, x, dx) (sw_raster.cc:679-697); the sort is not
// stable and stroke outlines contain many edges with identical keys, so the exact algorithm
// (introsort, threshold 16, median-of-three, unguarded partition, final insertion sort, heapsort
// at the depth limit) is reproduced to obtain the same permutation.
SKB_HD bool edge_less(const Edge* E, int a, int b) {
  int va = E[a].upper_y, vb = E[b].upper_y;
  if (va == vb) { va = E[a].x; vb = E[b].x; }
  if (va == vb) { va = E[a].dx; vb = E[b].dx; }
  return va < vb;
}
SKB_HDN void ss_unguarded_linear_insert(const Edge* E, int32_t* v, int last) {
  int32_t val = v[last];
  int next = last - 1;
  while (edge_less(E, val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}
SKB_HDN void ss_insertion_sort(const Edge* E, int32_t* v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (edge_less(E, v[i], v[first])) {
      int32_t val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1];
      v[first] = val;
    } else {
      ss_unguarded_linear_insert(E, v, i);
    }
  }
}
SKB_HDN void ss_adjust_heap(const Edge* E, int32_t* v, int first, int hole, int len, int32_t value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (edge_less(E, v[first + child], v[first + child - 1])) child--;
    v[first + hole] = v[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[first + hole] = v[first + child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && edge_less(E, v[first + parent], value)) {
    v[first + hole] = v[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[first + hole] = value;
}
SKB_HDN void ss_heap_sort(const Edge* E, int32_t* v, int first, int last) {
  int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      ss_adjust_heap(E, v, first, parent, len, v[first + parent]);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    int32_t val = v[last];
    v[last] = v[first];
    ss_adjust_heap(E, v, first, 0, last - first, val);
  }
}
SKB_HDN void sort_edge_indices(const Edge* E, int32_t* v, int n) {
  if (n <= 1) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) lg++;
  // iterative __introsort_loop: the recursion on [cut,last) becomes an explicit stack
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 0;
  stack_first[0] = 0;
  stack_last[0] = n;
  stack_depth[0] = lg * 2;
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
        ss_heap_sort(E, v, first, last);
        break;
      }
      --depth;
      int mid = first + (last - first) / 2;
      int a = first + 1, b = mid, c = last - 1;
      int pick;
      if (edge_less(E, v[a], v[b])) {
        if (edge_less(E, v[b], v[c])) pick = b;
        else if (edge_less(E, v[a], v[c])) pick = c;
        else pick = a;
      } else if (edge_less(E, v[a], v[c])) pick = a;
      else if (edge_less(E, v[b], v[c])) pick = c;
      else pick = b;
      { int32_t t = v[first]; v[first] = v[pick]; v[pick] = t; }
      int lo = first + 1, hi = last;
      for (;;) {
        while (edge_less(E, v[lo], v[first])) ++lo;
        --hi;
        while (edge_less(E, v[first], v[hi])) --hi;
        if (!(lo < hi)) break;
        int32_t t = v[lo]; v[lo] = v[hi]; v[hi] = t;
        ++lo;
      }
      // std: recurse on [lo,last) first, then continue with [first,lo).  The two ranges are
      // disjoint, so doing the left part first yields the same arrangement.
      if (sp < 64) {
        stack_first[sp] = lo;
        stack_last[sp] = last;
        stack_depth[sp] = depth;
        sp++;
      }
      last = lo;
    }
  }
  if (n > 16) {
    ss_insertion_sort(E, v, 0, 16);
    for (int i = 16; i != n; ++i) ss_unguarded_linear_insert(E, v, i);
  } else {
    ss_insertion_sort(E, v, 0, n);
  }
}

// ---- active edge list helpers (indices into the path's slot array; 0 = head, 1 = tail) ---------
#define SKB_HEAD 0
#define SKB_TAIL 1
SKB_HD void upd_nny(fx y, fx next_y, fx* nny) { if (y > next_y && y < *nny) *nny = y; }
SKB_HD void check_intersection(Edge* E, int e, fx next_y, fx* nny) {
  int p = E[e].prev;
  if (E[p].prev >= 0 && fx_add(E[p].x, E[p].dx) > fx_add(E[e].x, E[e].dx)) *nny = fx_add(next_y, SKB_FX1 >> 2);
}
SKB_HD void remove_edge(Edge* E, int e) {
  E[E[e].prev].next = E[e].next;
  E[E[e].next].prev = E[e].prev;
}
SKB_HD void insert_after(Edge* E, int e, int after) {
  E[e].prev = after;
  E[e].next = E[after].next;
  E[E[after].next].prev = e;
  E[after].next = e;
}
SKB_HDN void backward_insert_on_x(Edge* E, int e) {
  fx x = E[e].x;
  int prev = E[e].prev;
  while (E[prev].prev >= 0 && E[prev].x > x) prev = E[prev].prev;
  if (E[prev].next != e) {
    remove_edge(E, e);
    insert_after(E, e, prev);
  }
}
SKB_HDN void insert_new_edges(Edge* E, int ne, fx y, fx* nny) {  // sw_raster.cc:207-247
  if (E[ne].upper_y > y) {
    upd_nny(E[ne].upper_y, y, nny);
    return;
  }
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
  int start = prev;
  while (E[start].prev >= 0 && E[start].x > E[ne].x) start = E[start].prev;
  do {
    int next = E[ne].next;
    bool placed = false;
    for (;;) {
      if (E[start].next == ne) { placed = true; break; }
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
SKB_HD bool too_close_edges(const Edge* E, int prev, int next, fx lowerY) {  // sw_raster.cc:139-143
  return next >= 0 && prev >= 0 && E[next].upper_y < lowerY &&
         fx_add(E[prev].x, SKB_FX1) >= fx_sub(E[next].x, fx_abs(E[next].dx));
}

// Sweep one path.  E[0..n_slots): slot 0/1 are the sentinels, the others hold candidate edges
// (valid bit = curve bit 24, set by the flatten stage).  `ord` is scratch for n_slots ints.
// Bounds are the integers SWRaster::RastePath derives (sw_raster.cc:741-780).
// Q holds the quadratic state of slot i at Q[i].
struct WalkState {
  fx y, nny;
};

// SWEdgeBuilder culling (sw_edge.cc:322-336) + SortEdges + ProcessEdges (sw_raster.cc:679-729) and
// the prologue of WalkEdges (:549-563).  Returns false when the path has no edge to sweep.
// With E2 (room for 2 * n_slots entries; Q2 null = the quadratic states go right behind the compact edges) the edges that take part are copied there in sweep order, one after the
// other (slots 2..n+1, the links simply i-1 / i+1), and the sweep runs on the copy: the edges that are active together
// — neighbours in sweep order far more often than along the contour — then share cache lines, and the invalid slots
// between them are gone.  Slot numbers are only names to the sweep (ties of the sort are decided before the copy), so
// the records are the same.  (Measured on C4a: 27.9 -> 24.7 ms; the copy kept in shared memory instead — 10 to 30 edges
// per thread — is slower at every size, 35 to 62 ms: the sweep lives on the number of paths in flight.)
SKB_HDN bool walk_prologue(Edge*& E, QuadState*& Q, int n_slots, int32_t* ord, float scan_top_f, float scan_bottom_f, int start_y,
                           fx left_clip, fx right_clip, WalkState& ws, int wide = 0, Edge* E2 = nullptr, QuadState* Q2 = nullptr) {
  int n = 0;
  for (int i = 2; i < n_slots; i++) {
    if (!((E[i].curve >> 24) & 1)) continue;
    // quadratic edges are culled on their whole y extent (q_first_y, q_last_y), which the flatten
    // stage parks in the not-yet-used link fields
    const bool quad = (E[i].curve >> 25) & 1;
    fx y0 = quad ? E[i].prev : E[i].upper_y;
    fx y1 = quad ? E[i].next : E[i].lower_y;
    if (can_be_ignored(scan_top_f, scan_bottom_f, y0, y1, wide)) continue;
    ord[n++] = i;
  }
  if (n == 0) return false;
  sort_edge_indices(E, ord, n);
  int first = ord[0], last = ord[n - 1];
  if (E2) {
    if (!Q2) Q2 = reinterpret_cast<QuadState*>(E2 + (n + 2));   // right behind the compact edges (same element size)
    for (int i = 0; i < n; i++) {
      const int src = ord[i];
      Edge e = E[src];
      const bool quad = (e.curve >> 25) & 1;
      e.prev = i == 0 ? SKB_HEAD : i + 1;
      e.next = i == n - 1 ? SKB_TAIL : i + 3;
      E2[i + 2] = e;
      if (quad) Q2[i + 2] = Q[src];
    }
    E = E2;
    Q = Q2;
    first = 2;
    last = n + 1;
  } else {
    for (int i = 0; i < n; i++) {
      E[ord[i]].prev = i == 0 ? SKB_HEAD : ord[i - 1];
      E[ord[i]].next = i == n - 1 ? SKB_TAIL : ord[i + 1];
    }
  }
  Edge& H = E[SKB_HEAD];
  Edge& T = E[SKB_TAIL];
  H.prev = -1; H.next = first;
  H.upper_y = H.lower_y = SKB_FX_MIN; H.dx = 0; H.dy = SKB_FX_MAX;
  H.curve = 0;
  T.prev = last; T.next = -1;
  T.upper_y = T.lower_y = SKB_FX_MAX; T.dx = 0; T.dy = SKB_FX_MAX;
  T.curve = 0;
  // WalkEdges (sw_raster.cc:546-677)
  H.x = left_clip;
  T.x = right_clip;
  fx y = fx_max(E[H.next].upper_y, i_to_fx(start_y));
  fx nny = SKB_FX_MAX;
  int e;
  for (e = H.next; E[e].upper_y <= y; e = E[e].next) {
    Edge& q = E[e];
    // SWEdge::GoY(dst) (sw_edge.hpp:43-51) with y == upper_y and upper_x == x, as they are before the sweep: one row
    // below the top the edge takes a dx step; anywhere else GoY computes upper_x + dx * (y - upper_y) = x
    if (y == fx_add(q.upper_y, SKB_FX1)) q.x = fx_add(q.x, q.dx);
    upd_nny(q.lower_y, y, &nny);
  }
  upd_nny(E[e].upper_y, y, &nny);
  ws.y = y;
  ws.nny = nny;
  return true;
}

// The same sweep as ONE loop: every iteration handles one active edge and, when that was the last
// edge of the band, the band's epilogue and the next band's prologue.  A warp of 32 paths then runs
// the same instruction stream however many edges and bands each path has — the nested form makes
// every lane wait for the longest inner loop of the warp at every band.
SKB_HDN void walk_bands_flat(Edge* E, QuadState* Q, WalkState ws, int stop_y, fx left_clip, fx right_clip,
                             int even_odd, RecSink& sink) {
  const int mask = even_odd ? 1 : -1;
  const fx stop_fx = i_to_fx(stop_y);
  fx y = ws.y, nny = ws.nny;
  int w = 0, cur = 0, left_edge = SKB_HEAD, prev_right = 0, y_shift = 0;
  bool in_interval = false;
  fx prev_x = 0, next_y = 0, left = 0, left_dy = 0;
  uint32_t full = 0;
#define SKB_BEGIN_BAND()                                                        \
  do {                                                                          \
    w = 0;                                                                      \
    in_interval = false;                                                        \
    prev_x = left_clip;                                                         \
    next_y = fx_min(nny, fx_ceil_fx(fx_add(y, 1)));                             \
    cur = E[SKB_HEAD].next;                                                     \
    left_edge = SKB_HEAD;                                                       \
    left = left_clip;                                                           \
    left_dy = 0;                                                                \
    prev_right = fx_floor_i(left_clip);                                         \
    nny = SKB_FX_MAX;                                                           \
    y_shift = 0;                                                                \
    if (fx_sub(next_y, y) & (SKB_FX1 >> 2)) {                                   \
      y_shift = 2;                                                              \
      next_y = fx_add(y, SKB_FX1 >> 2);                                         \
    } else if (fx_sub(next_y, y) & (SKB_FX1 >> 1)) {                            \
      y_shift = 1;                                                              \
    }                                                                           \
    full = (uint32_t)(uint8_t)fx_round_i((fx)(0xFF * fx_sub(next_y, y)));       \
    cur_upper = E[cur].upper_y;                                                 \
    lk_real = false;                                                            \
  } while (0)
  // cur_upper = E[cur].upper_y, loaded one iteration ahead (nothing an iteration does changes the upper_y of the
  // edge after it).  lk_*: the edge that sits right before the not-yet-visited part of the list — the last visited
  // edge that stayed where it was — as registers: its x + dx is all check_intersection needs of it.
  fx cur_upper = 0, lk_sum = 0;
  bool lk_real = false;
  SKB_BEGIN_BAND();
  bool done = false;
  while (!done) {
    if (cur_upper <= y) {
      Edge c = E[cur];
      const fx next_upper = E[c.next].upper_y;
      w += edge_winding(c);
      const bool prev_in = in_interval;
      in_interval = (w & mask) != 0;
      const bool is_left = in_interval && !prev_in, is_right = !in_interval && prev_in;
      const fx old_x = c.x;
      c.x = fx_add(c.x, c.dx >> y_shift);
      if (is_left) {
        left = fx_max(old_x, left_clip);
        left_dy = c.dy;
        left_edge = cur;
      } else if (is_right) {
        const fx right = fx_min(right_clip, old_x);
        const fx le_x = E[left_edge].x;
        TrapRec r;
        r.y = y >> 16;
        r.ul = left;
        r.ur = right;
        r.ll = fx_max(left_clip, le_x);
        r.lr = fx_min(right_clip, c.x);
        r.ldy = left_dy;
        r.rdy = c.dy;
        bool no_real = false;
        if (full == 0xFF) {
          const int nx = c.next;  // too_close_edges(cur, cur->next, next_y) with cur's advanced x
          no_real = (prev_right > fx_floor_i(left) || prev_right > fx_floor_i(le_x)) ||
                    (next_upper < next_y && fx_add(c.x, SKB_FX1) >= fx_sub(E[nx].x, fx_abs(E[nx].dx)));
        }
        r.flags = full | (no_real ? 0x100u : 0u);
        sink_emit(sink, r);
        prev_right = fx_ceil_i(fx_max(right, c.x));
      }
      const int next = c.next;
      bool chord = false;
      while (c.lower_y <= next_y) {
        if (edge_count(c) > 0) {
          QuadState& q = Q[cur];
          chord = true;
          // SWQuadEdge::KeepContinuous (sw_edge.cc:294-297): the next chord starts where the sweep has brought the edge
          if (!update_quad(c, q, c.x, next_y)) break;
        } else {
          break;
        }
      }
      // write back what changed (the links are edited in place below): x always, the rest only when the edge took
      // its next chord
      E[cur].x = c.x;
      if (chord) {
        E[cur].dx = c.dx; E[cur].dy = c.dy;
        E[cur].upper_y = c.upper_y; E[cur].lower_y = c.lower_y; E[cur].curve = c.curve;
      }
      if (c.lower_y <= next_y) {
        remove_edge(E, cur);
      } else {
        upd_nny(c.lower_y, next_y, &nny);
        const fx sum = fx_add(c.x, c.dx);
        if (c.x < prev_x) {
          backward_insert_on_x(E, cur);
          check_intersection(E, cur, next_y, &nny);
          if (E[cur].next == next) {  // not moved (only the head was before it)
            lk_real = true;
            lk_sum = sum;
          }
        } else {
          prev_x = c.x;
          // check_intersection with the previous edge's x + dx at hand
          if (lk_real && lk_sum > sum) nny = fx_add(next_y, SKB_FX1 >> 2);
          lk_real = true;
          lk_sum = sum;
        }
      }
      cur = next;
      cur_upper = next_upper;
    }
    if (cur_upper > y) {
      if (in_interval) {
        TrapRec r;
        r.y = y >> 16;
        r.ul = left;
        r.ur = right_clip;
        r.ll = fx_max(left_clip, E[left_edge].x);
        r.lr = right_clip;
        r.ldy = left_dy;
        r.rdy = 0;
        bool no_real = full == 0xFF && too_close_edges(E, E[left_edge].prev, left_edge, next_y);
        r.flags = full | (no_real ? 0x100u : 0u);
        sink_emit(sink, r);
      }
      y = next_y;
      if (y >= stop_fx) {
        done = true;
      } else {
        insert_new_edges(E, cur, y, &nny);
        SKB_BEGIN_BAND();
      }
    }
  }
#undef SKB_BEGIN_BAND
  sink_flush_row(sink);
}

// The band loop in the reference's own shape (a loop over bands around a loop over the active edges) lives in
// tests/sim/walk_nested.hpp: the CPU simulation uses it to cross-check the flat loop; the product never runs it.
SKB_HDN void walk_path(Edge* E, QuadState* Q, int n_slots, int32_t* ord, float scan_top_f, float scan_bottom_f, int start_y,
                       int stop_y, fx left_clip, fx right_clip, int even_odd, RecSink& sink, int wide = 0,
                       Edge* E2 = nullptr, QuadState* Q2 = nullptr) {
  WalkState ws;
  if (!walk_prologue(E, Q, n_slots, ord, scan_top_f, scan_bottom_f, start_y, left_clip, right_clip, ws, wide, E2, Q2)) return;
  walk_bands_flat(E, Q, ws, stop_y, left_clip, right_clip, even_odd, sink);
}


}  // namespace skb

#endif  // SKB_WALK_CUH
