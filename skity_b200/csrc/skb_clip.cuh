// skb_clip.cuh — the path-clip stack, per pixel row.
//
// The reference keeps a clip as a LIST OF SPANS and clips a draw by intersecting span lists
// (SWCanvas::State::PerformClip / FindSpan, src/render/sw/sw_canvas.cc:158-265).  Observable
// consequences that a per-pixel min(coverage) would miss, all reproduced here:
//   * a pixel is blended once per (draw span, clip span) PAIR that covers it, in list order;
//   * FindSpan's `+ 1` (:257): a clip span that starts inside a draw span S and reaches past its
//     end also yields the pixel just after S — so the span ending at a pixel matters too;
//   * the nested clip is the list of those sub-spans (RecursiveClip :178-193), starts included;
//   * HasClip() is "list non-empty": a clip that rasterised to nothing clips nothing (sw_canvas.hpp:36).
// The span structure of a rasterised path follows from its trapezoid rows: a directly emitted row
// gives one-pixel spans under its slanted edges and ONE span for its fully covered interior
// (blit_full_alpha, sw_raster.cc:303-310); accumulated rows are run-length encoded by equal value
// (SpanBuilder::Flush, :108-136) and come after the direct spans of the row.
//
// One thread sweeps one pixel row left to right (pure per-thread code, shared with the CPU
// simulation).  Per pixel it forms the "S side" (own direct / own accumulated span, and the direct /
// accumulated span ENDING at this pixel), reads the "C side" (the clip state's spans covering the
// pixel: start + coverage, in list order) and emits the ordered list of resulting coverages.
#ifndef SKB_CLIP_CUH
#define SKB_CLIP_CUH

#include "skity_b200/csrc/skb_core.cuh"
#include "skity_b200/csrc/skb_sort.cuh"

namespace skb {

#define SKB_CLIP_MAXE 8          // spans of a clip state that may cover one pixel
#define SKB_CLIP_PLANES 8        // coverage planes of a clipped draw
#define SKB_CLIP_RMAX 48         // prepared records per row (shared memory on the GPU; 96 costs 15 % of the clip stage in occupancy); rows with more are swept
                                 // by one thread that re-reads the records for every pixel
#define SKB_CLIP_START_BIAS (1 << 22)

// A clip-state entry: coverage (bits 0-7) | (span start x + bias) << 8 | marker flag (bit 31).  0 = no entry.
// Besides the spans that clip something, a state keeps the spans the reference keeps although they cover nothing —
// they make HasClip() true (a non-empty span list) and they breed more of their kind in nested clips:
//   * spans of coverage 0 (a directly emitted span whose coverage came out as 0; FindSpan keeps min(.., 0)),
//     stored like any span;
//   * zero-LENGTH spans, which FindSpan makes when a parent span ends exactly where an own span starts
//     (`clip.x + clip.len == span.x`, sw_canvas.cc:228-241) and which it hands down whenever an own span covers or
//     ends at their position: stored as a MARKER entry at that pixel.
#define SKB_CLIP_MARKER 0x80000000u
SKB_HD uint32_t clip_entry(int start, uint32_t cover) { return cover | ((uint32_t)(start + SKB_CLIP_START_BIAS) << 8); }
SKB_HD int clip_entry_start(uint32_t e) { return (int)((e & ~SKB_CLIP_MARKER) >> 8) - SKB_CLIP_START_BIAS; }
SKB_HD uint32_t clip_entry_cover(uint32_t e) { return e & 0xFF; }
SKB_HD bool clip_entry_is_marker(uint32_t e) { return (e & SKB_CLIP_MARKER) != 0; }

struct SpanSide {   // one span on the S side at the current pixel
  uint32_t cover;   // may be 0 for a directly emitted span (see above)
  int start;
  bool present;
};

// Ordered output of one pixel.
struct ClipOut {
  uint32_t e[SKB_CLIP_MAXE > SKB_CLIP_PLANES ? SKB_CLIP_MAXE : SKB_CLIP_PLANES];
  int n;
  bool overflow;
};

// keep: 2 = building a clip state (spans that cover nothing are kept, zero-length markers included); 1 = a clipped draw
// whose blend mode or colour filter acts on zero-coverage pixels (spans of coverage 0 kept, markers dropped: they
// have no pixel); 0 = any other clipped draw (both dropped).
SKB_HD void clip_out_push(ClipOut& o, int cap, int keep, int start, uint32_t cover, bool marker) {
  if (marker ? keep < 2 : (cover == 0 && keep < 1)) return;
  if (o.n >= cap) { o.overflow = true; return; }
  o.e[o.n++] = clip_entry(start, cover) | (marker ? SKB_CLIP_MARKER : 0u);
}

// Combine the S side of pixel x with the clip spans covering it (FindSpan, all three cases).
//   own   : S contains the pixel            -> every C gives min(cover), sub-span starts at max(S.x, C.x)
//   left  : S ends exactly at the pixel     -> only C with C.x > S.x (the `+ 1`), sub-span starts at C.x
//   C a zero-length marker here             -> a marker again, for an own and for a left S alike
//   C ends exactly here, S starts here      -> a new marker (c_prev = entries of pixel x - 1)
// Order: spans in emission order (direct before accumulated, left before own), C in list order.
SKB_HDN void clip_combine(int x, const SpanSide& left_d, const SpanSide& own_d, const SpanSide& left_a, const SpanSide& own_a,
                          const uint32_t* clist, int n_c, const uint32_t* c_prev, int n_prev, bool clipped, int cap,
                          int keep_ghosts, ClipOut& out) {
  out.n = 0;
  out.overflow = false;
  if (!clipped) {  // HasClip() false: the raster spans themselves
    if (own_d.present) clip_out_push(out, cap, keep_ghosts, own_d.start, own_d.cover, false);
    if (own_a.present) clip_out_push(out, cap, keep_ghosts, own_a.start, own_a.cover, false);
    return;
  }
  const SpanSide* seq[4] = {&left_d, &own_d, &left_a, &own_a};
  for (int k = 0; k < 4; k++) {
    const SpanSide& s = *seq[k];
    if (!s.present) continue;
    const bool is_left = (k & 1) == 0;
    if (!is_left && keep_ghosts == 2 && s.start == x) {
      // parent spans that end exactly where this span starts
      for (int i = 0; i < n_prev; i++) {
        if (clip_entry_is_marker(c_prev[i])) continue;
        bool continues = false;
        for (int j = 0; j < n_c; j++) continues |= clist[j] == c_prev[i];
        if (continues) continue;
        const uint32_t pc = clip_entry_cover(c_prev[i]);
        clip_out_push(out, cap, 2, x, pc < s.cover ? pc : s.cover, true);
      }
    }
    for (int i = 0; i < n_c; i++) {
      const int cstart = clip_entry_start(clist[i]);
      const uint32_t ccover = clip_entry_cover(clist[i]);
      const uint32_t m = ccover < s.cover ? ccover : s.cover;
      if (clip_entry_is_marker(clist[i])) {
        clip_out_push(out, cap, keep_ghosts, cstart, m, true);
      } else if (is_left) {
        if (cstart > s.start) clip_out_push(out, cap, keep_ghosts, cstart, m, false);
      } else {
        clip_out_push(out, cap, keep_ghosts, cstart > s.start ? cstart : s.start, m, false);
      }
    }
  }
}

// Sweep state of one row.
struct ClipRowState {
  TrapPrep* prep;  // SKB_CLIP_RMAX prepared records, storage given by the caller: the threads that share a row on the
                   // GPU share ONE array in shared memory (each prepares a part of it)
  int n_prep;      // records prepared (row has at most SKB_CLIP_RMAX) or -1: evaluate generically
  // the records clip_row_step looks at, in record order: all of them, or (clip_row_focus) those that reach the
  // pixels the caller is going to step through
  uint8_t act[SKB_CLIP_RMAX];
  int n_act;
  // previous pixel
  uint32_t prev_d, prev_a;
  int prev_d_start, prev_a_start;
  bool prev_d_ends;  // the direct span covering the previous pixel ends at the current pixel
};

// S side of pixel x of a row given its trapezoid records; advances the sweep state.
// Must be called for consecutive x starting at the row's first possibly covered pixel.
SKB_HDN void clip_row_step(ClipRowState& st, const TrapRec* pool, uint2 row, int x, SpanSide& left_d, SpanSide& own_d,
                           SpanSide& left_a, SpanSide& own_a) {
  uint32_t d = 0, acc = 0;
  int d_start = x;
  bool d_ends_next = true;
  bool d_touched = false;
  if (st.n_prep >= 0) {
    for (int i = 0; i < st.n_act; i++) {
      const TrapPrep& p = st.prep[st.act[i]];
      uint8_t v;
      if (!trap_prep_alpha(p, x, &v)) continue;
      if (!p.accum) {
        d = v;
        d_touched = true;
        if (p.mode == 1 && x >= p.jl && x < p.jr) {  // interior of a direct row: one long span
          d_start = p.jl;
          d_ends_next = (x + 1 == p.jr);
        } else {
          d_start = x;
          d_ends_next = true;
        }
      } else {
        acc += v;
      }
    }
  } else {
    uint32_t idx = row.x;
    for (uint32_t k = 0; k < row.y; k++, idx++) {
      TrapRec r = pool[idx];
      if (r.flags & SKB_REC_LINK) {
        idx = (uint32_t)r.y;
        r = pool[idx];
      }
      const TrapPrep p = trap_prepare(r);
      uint8_t v;
      if (!trap_prep_alpha(p, x, &v)) continue;
      if (!p.accum) {
        d = v;
        d_touched = true;
        if (p.mode == 1 && x >= p.jl && x < p.jr) {
          d_start = p.jl;
          d_ends_next = (x + 1 == p.jr);
        } else {
          d_start = x;
          d_ends_next = true;
        }
      } else {
        acc += v;
      }
    }
  }
  const uint32_t a = acc > 255u ? 255u : acc;
  // spans ending at this pixel
  left_d.present = st.prev_d_ends;
  left_d.cover = st.prev_d_ends ? st.prev_d : 0u;
  left_d.start = st.prev_d_start;
  const bool a_run_continues = a != 0 && a == st.prev_a;
  left_a.cover = (st.prev_a != 0 && !a_run_continues) ? st.prev_a : 0u;
  left_a.present = left_a.cover != 0;
  left_a.start = st.prev_a_start;
  // spans covering this pixel
  own_d.cover = d;
  own_d.present = d_touched;   // a directly emitted span exists here even when its coverage came out as 0
  own_d.start = d_start;
  own_a.cover = a;
  own_a.present = a != 0;
  own_a.start = a_run_continues ? st.prev_a_start : x;
  // advance
  st.prev_d = d;
  st.prev_d_start = d_start;
  st.prev_d_ends = d_touched ? d_ends_next : false;
  st.prev_a = a;
  st.prev_a_start = own_a.start;
}

// Accumulated coverage of pixel x from the prepared records (saturated), without touching the sweep state.
SKB_HDN uint32_t clip_row_accum_at(const ClipRowState& st, int x) {
  uint32_t acc = 0;
  for (int k = 0; k < st.n_prep; k++) {
    const TrapPrep& p = st.prep[k];
    uint8_t v;
    if (p.accum && trap_prep_alpha(p, x, &v)) acc += v;
  }
  return acc > 255u ? 255u : acc;
}

// Puts the sweep state where a sweep from `x_first` would have it after pixel x0 - 1, so that several threads
// can share one row (each takes a run of pixels).  Only for rows whose records are all prepared (n_prep >= 0).
// Everything about pixel x0 - 1 follows from the records alone except the START of the accumulated run it
// belongs to — the run of equal, non-zero accumulated coverage — which is found by walking left: whole
// stretches where every record contributes a constant are skipped at once, only edge zones are walked pixel by
// pixel.
SKB_HDN void clip_row_seek(ClipRowState& st, const TrapRec* pool, uint2 row, int x_first, int x0) {
  if (x0 <= x_first) return;
  SpanSide ld, od, la, oa;
  st.prev_d = st.prev_a = 0;
  st.prev_d_ends = false;
  clip_row_step(st, pool, row, x0 - 1, ld, od, la, oa);  // from a blank state: everything but prev_a_start is right
  const uint32_t a0 = st.prev_a;
  if (a0 == 0) return;
  int cur = x0 - 1;
  for (;;) {
    // leftmost pixel zl <= cur such that every accumulating record is constant on [zl, cur]
    int zl = x_first;
    for (int k = 0; k < st.n_prep; k++) {
      const TrapPrep& p = st.prep[k];
      if (!p.accum || p.mode == 0) continue;
      int z;
      if (cur >= p.R) z = p.R;                          // right of the record: 0 back to its end
      else if (cur < p.L) continue;                     // left of it: 0 all the way
      else if (cur >= p.jl && cur < p.jr) z = p.jl;     // interior: `full` back to jl
      else z = cur;                                     // edge zone: nothing assumed
      zl = z > zl ? z : zl;
    }
    cur = zl;
    if (cur <= x_first) break;
    if (clip_row_accum_at(st, cur - 1) != a0) break;
    cur--;
  }
  st.prev_a_start = cur;
}

// `storage`: SKB_CLIP_RMAX entries.  `part` / `n_parts`: this caller prepares the records k with k % n_parts == part
// (1 part = all of them); callers that share the storage synchronise before they step.
SKB_HDN void clip_row_begin(ClipRowState& st, const TrapRec* pool, uint2 row, TrapPrep* storage, int part = 0, int n_parts = 1) {
  st.prep = storage;
  st.prev_d = st.prev_a = 0;
  st.prev_d_start = st.prev_a_start = 0;
  st.prev_d_ends = false;
  st.n_act = 0;
  if (row.y > (uint32_t)SKB_CLIP_RMAX) {
    st.n_prep = -1;
    return;
  }
  st.n_prep = (int)row.y;
  st.n_act = (int)row.y;
  uint32_t idx = row.x;
  for (uint32_t k = 0; k < row.y; k++, idx++) {
    if (pool[idx].flags & SKB_REC_LINK) idx = (uint32_t)pool[idx].y;
    if ((int)(k % (uint32_t)n_parts) == part) st.prep[k] = trap_prepare(pool[idx]);
    st.act[k] = (uint8_t)k;
  }
}

SKB_HD void clip_row_unfocus(ClipRowState& st) {
  for (int k = 0; k < st.n_prep; k++) st.act[k] = (uint8_t)k;
  st.n_act = st.n_prep > 0 ? st.n_prep : 0;
}

// From here on clip_row_step is only called for pixels xa..xb: keep the records that give one of them a value (a
// record gives nothing outside [L, R)).  A thread that shares a row with others steps through a run of a few pixels
// and most of the row's records do not reach it.
SKB_HDN void clip_row_focus(ClipRowState& st, int xa, int xb) {
  if (st.n_prep < 0) return;
  int n = 0;
  for (int k = 0; k < st.n_prep; k++) {
    const TrapPrep& p = st.prep[k];
    if (p.mode == 0 || p.R <= xa || p.L > xb) continue;
    st.act[n++] = (uint8_t)k;
  }
  st.n_act = n;
}

// ---- ClipOp::kDifference ------------------------------------------------------------------------------------------
// A state whose op is kDifference keeps the clip path's spans as they were rasterised and a draw is cut span by span
// (SWCanvas::State::PerformClip -> spans_subtraction, src/render/sw/sw_canvas.cc:56-133,158-161).  The reference's
// subtraction is sequential over the row's clip spans SORTED BY x (std::sort, ties as libstdc++ leaves them), ignores
// their coverage, and — by its own comment "not correct" — keeps what a clip span overlaps on the LEFT of the running
// span; all of it is observable and reproduced.  A difference state is therefore stored as per-row sorted span lists
// (x, len), not as the per-pixel entry table of intersecting states.
//
// The spans of one rasterised row in the order the reference's list holds them: the directly emitted spans as the
// sweep emits them (ascending x: one-pixel spans under slanted edges, one span per fully covered interior), then the
// row's accumulated coverage run-length encoded (SpanBuilder::Flush, sw_raster.cc:108-136) — an accumulated row is
// flushed after the sweep has left it.  Calls onD(x, len, cover) / onA(x, len, cover).
template <class OnD, class OnA>
SKB_HDN void clip_row_spans(ClipRowState& st, const TrapRec* pool, uint2 row, int x_first, int x_last, OnD& onD, OnA& onA) {
  for (int x = x_first; x <= x_last; x++) {
    SpanSide ld, od, la, oa;
    clip_row_step(st, pool, row, x, ld, od, la, oa);
    if (la.present) onA(la.start, x - la.start, la.cover);
    if (od.present && st.prev_d_ends) onD(od.start, x + 1 - od.start, od.cover);
  }
  if (st.prev_a != 0) onA(st.prev_a_start, x_last + 1 - st.prev_a_start, st.prev_a);
}

struct SpanXLess {
  SKB_HD bool operator()(const uint2& a, const uint2& b) const { return (int)a.x < (int)b.x; }
};

// spans_subtraction for ONE span [sx, sx + slen) against the row's clip spans ms[0..n) (x, len; sorted by x).
// Calls emit(x, len) for every piece the reference appends, zero-length pieces included.
template <class Emit>
SKB_HDN void span_subtract(int sx, int slen, const uint2* ms, int n, Emit& emit) {
  if (n == 0) {   // "no spans in this line means minus zero"
    emit(sx, slen);
    return;
  }
  int cx = sx, cl = slen;
  for (int j = 0; j < n; j++) {
    const int mx = (int)ms[j].x, ml = (int)ms[j].y;
    if (mx + ml < cx || mx > cx + cl) continue;
    if (mx < cx) {
      if (mx + ml > cx + cl) {
        cl = 0;
        break;
      }
      const int last = cx + cl, len = mx + ml - cx;
      if (len == 0) continue;
      emit(cx, len);   // the overlapped part is KEPT (sw_canvas.cc:84-98)
      cx += len;
      cl = last - cx;
    } else {
      if (mx + ml < cx + cl) {
        const int last = cx + cl;
        emit(cx, mx - cx);
        cx = mx + ml;
        cl = last - cx;
      } else {
        emit(cx, mx - cx);
        cl = 0;
      }
    }
    if (cl <= 0) break;
  }
  if (cl > 0) emit(cx, cl);
}

}  // namespace skb

#endif  // SKB_CLIP_CUH
