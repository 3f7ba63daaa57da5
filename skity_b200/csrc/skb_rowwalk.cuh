// skb_rowwalk.cuh — stage "walk", row-parallel form: the reference's active-edge sweep
// (WalkEdges, src/render/sw/sw_raster.cc:546-677) evaluated by one thread per (path, pixel row)
// instead of one thread per path, with bit-identical trapezoid records.
//
// Why this is possible.  The sweep is sequential in y only through three things:
//   (1) band structure — every edge start/end inside a pixel row splits the row into ¼ / ½ bands
//       for ALL edges (sw_raster.cc:571-591); these events are known up front: line ends are given,
//       and the y of every chord of a quadratic comes from forward differencing of y alone
//       (SWQuadEdge::UpdateQuad, sw_edge.cc:233-292 — newy, the snapped y and the success of a chord
//       do not depend on x);
//   (2) an edge's x at a band top = upper_x + Σ (dx >> y_shift) over the bands it crossed
//       (SWEdge::GoY, sw_edge.hpp:55-58): whole-pixel bands add dx exactly, ½ and ¼ bands truncate —
//       so x is a closed form of the NUMBER of rows of each band pattern between the edge's start
//       and the row (four patterns: 1 | ½+½ | ¼+¼+½ in any order | 4×¼), i.e. of three per-path
//       prefix counts over rows; the next chord of a quadratic starts at the walked x
//       (KeepContinuous, sw_edge.cc:294-297), so a quadratic is a short sequential chain (≤ 64 chords)
//       that only needs those counts;
//   (3) check_intersection (sw_raster.cc:164-169): two list neighbours about to cross force the NEXT
//       band to ¼.  Inside a row the row's thread simulates this itself; across a row boundary it is
//       one bit f(Y), and a row maps f_in to f_out — a function on {0,1}.  Every row is evaluated
//       under both hypotheses in parallel and the per-path composition over rows (a trivial serial
//       pass per path) yields the true bit of every row.
// The forced ¼ bands change the truncation losses of (2) by a few 2^-16 px, which feeds back into
// (3) only through comparisons that are within those few units of a tie.  So the bits are first
// computed from tables without forced bands (T0), the tables rebuilt with them (T1), chords chained
// again, and the final pass — which emits the records — RE-DERIVES every bit and band pattern with
// the exact x and compares: a path whose tables are self-consistent is, by induction over bands,
// exactly the sequential sweep; any path that is not (or that meets a case this form does not
// model: unresolved ties in x, too many edges in a row, a right edge culled away, ...) is flagged and
// swept by the sequential walker (skb_walk.cuh) afterwards.  Nothing is approximated.
//
// Pure per-thread code (SKB_HD): compiled into the kernels of skb_backend.cu and into the CPU
// simulation (tests/sim/), where the records are compared one by one with the sequential walker's.
#ifndef SKB_ROWWALK_CUH
#define SKB_ROWWALK_CUH

#include "skity_b200/csrc/skb_walk.cuh"

namespace skb {

#define SKB_RW_MAXA 24        // active edges of one row a thread tracks (more: sequential fallback)
#define SKB_RW_MAXN 12         // edges starting within one row
#define SKB_RW_MAXQ 32767     // quarter rows a path may span (15 bits)
#define SKB_ROW_LINEAR 0x80000000u  // rows[].y flag: the row's records are consecutive in the pool (no chunk links)

// One chord of an edge (a line edge has one): what SWEdge holds while the chord is active.
struct alignas(16) Chord {
  fx x, dx, dy;   // x at the chord's upper end, slope, |1/slope|
  uint32_t yy;    // upper quarter row | lower quarter row << 15 | (winding < 0) << 30   (relative to the op's origin row)
};
SKB_HD int chord_uq(uint32_t yy) { return (int)(yy & 0x7FFF); }
SKB_HD int chord_lq(uint32_t yy) { return (int)((yy >> 15) & 0x7FFF); }
SKB_HD int chord_w(uint32_t yy) { return (yy >> 30) & 1 ? -1 : 1; }
SKB_HD uint32_t chord_yy(int uq, int lq, int w) { return (uint32_t)uq | ((uint32_t)lq << 15) | (w < 0 ? 1u << 30 : 0u); }

// Per edge slot: where its chords are.
struct alignas(16) SlotInfo {
  uint32_t chord_base;
  uint32_t n_chords;   // 0: no edge (empty slot, degenerate, or culled by CanBeIgnored)
  int32_t q_first, q_last;  // quarter rows [first chord's upper, last chord's lower)
};

// Band table entry of one walk row (rows counted from the op's origin row = WalkEdges' start_y).
struct RowBand {
  uint16_t nB, nC, nD;  // rows ABOVE this one with band pattern ½+½ / ¼+¼+½ (any order) / 4×¼
  uint8_t mask;         // band boundaries inside this row: bit 0 at ¼, bit 1 at ½, bit 2 at ¾
  uint8_t fl;           // bit 0: the row's first band is forced to ¼ (f_in); bit 1: its last band forces the next row's (f_out of the
                        // surviving edges); bit 2: the edges inserted at the row's top force it (f_ins)
};
#define SKB_RB_FIN 1u
#define SKB_RB_FSURV 2u
#define SKB_RB_FINS 4u

// RwOp::fail bits
#define SKB_RWFAIL_HARD 1u    // sweep this path sequentially (skb_walk.cuh)
#define SKB_RWFAIL_SOFT 2u    // the first attempt's tables were not self-consistent (or need the sort ranks): retry once
#define SKB_RWFAIL_RETRY 16u  // being retried

// Per op.
struct RwOp {
  uint32_t wrow_base;   // first entry in the band table / event table (n_wrows + 1 entries)
  int32_t n_wrows;      // rows the sweep covers: [origin_row, origin_row + n_wrows)
  int32_t origin_row;   // start_y
  int32_t y0q;          // quarter row of the sweep's first y (min upper_y of the edges), INT_MAX if no edge
  uint32_t fail;        // != 0: sweep this path sequentially
  uint32_t rec_base;    // first record of the op's linear record region
  uint32_t n_recs;      // records the final pass is to emit (sum over the rows)
  uint32_t pad;
};

// ---- band arithmetic --------------------------------------------------------------------------
// Actual band boundaries of a row given the quarter positions of its events (bits 0..2), the quarter the sweep enters
// the row at and whether the first band is forced to ¼ (sw_raster.cc:571-591).  Forcing inside the row is the row
// thread's business (rw_row); this is the table-side form without it.
SKB_HD uint32_t rw_band_mask(uint32_t ev, int p0, bool forced_first) {
  uint32_t mask = 0;
  int p = p0;
  bool forced = forced_first;
  while (p < 4) {
    const uint32_t above = (ev | 8u) >> p;          // bit i: an event (or the row end) at quarter p + 1 + i
    int h = 1;
    while (!((above >> (h - 1)) & 1u)) h++;
    if (forced || (h & 1)) h = 1;
    forced = false;
    p += h;
    if (p < 4) mask |= 1u << (p - 1);
  }
  return mask;
}
// pattern of a whole row: 0 = one band, 1 = ½+½, 2 = ¼+¼+½ in some order, 3 = 4×¼
SKB_HD int rw_row_class(uint32_t mask) { return mask == 0 ? 0 : mask == 2 ? 1 : mask == 7 ? 3 : 2; }

// x after the bands of one row between quarters q0 and q1 (both band boundaries of `mask`, or 0 / 4)
SKB_HD fx rw_adv_partial(fx x, fx dx, uint32_t mask, int q0, int q1) {
  int p = q0;
  while (p < q1) {
    const uint32_t above = (mask | 8u) >> p;
    int h = 1;
    while (!((above >> (h - 1)) & 1u)) h++;
    x = fx_add(x, dx >> (h == 1 ? 2 : h == 2 ? 1 : 0));
    p += h;
  }
  return x;
}

// x of an edge (x0 at quarter row q0, slope dx) at quarter row q1 >= q0, both band boundaries.  rows = the op's band table.
SKB_HDN fx rw_walk_x(fx x, fx dx, int q0, int q1, const RowBand* rows) {
  if (q1 <= q0) return x;
  int r0 = q0 >> 2;
  const int r1 = q1 >> 2;
  if (r0 == r1) return rw_adv_partial(x, dx, rows[r0].mask, q0 & 3, q1 & 3);
  if (q0 & 3) {
    x = rw_adv_partial(x, dx, rows[r0].mask, q0 & 3, 4);
    r0++;
  }
  if (r1 > r0) {
    const uint32_t nB = (uint16_t)(rows[r1].nB - rows[r0].nB), nC = (uint16_t)(rows[r1].nC - rows[r0].nC),
                   nD = (uint16_t)(rows[r1].nD - rows[r0].nD);
    const uint32_t nA = (uint32_t)(r1 - r0) - nB - nC - nD;
    const uint32_t h = (uint32_t)(dx >> 1), q = (uint32_t)(dx >> 2);
    x = (fx)((uint32_t)x + nA * (uint32_t)dx + nB * (2u * h) + nC * (2u * q + h) + nD * (4u * q));
  }
  if (q1 & 3) x = rw_adv_partial(x, dx, rows[r1].mask, 0, q1 & 3);
  return x;
}

// ---- chords of one edge slot ----------------------------------------------------------------------
// Walks the chain of chords the sweep would give this edge: chord k+1 is UpdateQuad from where the sweep brought
// chord k (KeepContinuous).  tab == nullptr: first pass — x is advanced without the band tables (only the y of the
// chords is meant: it does not depend on x), the chords' quarter rows are stored, `ev` receives the events;
// tab != nullptr: second pass — exact x; the stored quarter rows must come out the same (returns -1 otherwise).
// `origin_fx` = origin row << 16, `stop_q` = quarter row the sweep stops at.  Returns the number of chords, -1 on a
// case the row-parallel form does not model.
SKB_HDN int rw_chain(const Edge& e0, const QuadState& q0, int origin_row, int stop_q, const RowBand* tab, Chord* out, int cap,
                     uint32_t* ev_words) {
  Edge c = e0;
  QuadState q = q0;
  const bool quad = (c.curve >> 25) & 1;
  const fx origin_fx = i_to_fx(origin_row);
  int n = 0;
  for (;;) {
    const fx rel_u = fx_sub(c.upper_y, origin_fx), rel_l = fx_sub(c.lower_y, origin_fx);
    if (rel_u < 0 || (rel_u & 0x3FFF) || (rel_l & 0x3FFF) || rel_l <= rel_u) return -1;
    const int uq = rel_u >> 14;
    int lq = rel_l >> 14;
    if (lq > SKB_RW_MAXQ) {
      if (uq >= stop_q) lq = SKB_RW_MAXQ;   // entirely below the sweep's last row: never looked at
      else return -1;
    }
    if (uq > SKB_RW_MAXQ || n >= cap) return -1;
    const uint32_t yy = chord_yy(uq, lq, edge_winding(c));
    if (tab) {
      if (out[n].yy != yy) return -1;
      out[n].x = c.x;
      out[n].dx = c.dx;
      out[n].dy = c.dy;
    } else {
      out[n].yy = yy;
      if (ev_words) {
        // events strictly inside a pixel row (quarter positions 1..3) — row-boundary events change nothing
        if ((uq & 3) && uq < stop_q) {
#if defined(__CUDA_ARCH__)
          atomicOr(&ev_words[uq >> 4], 1u << (8 * ((uq >> 2) & 3) + (uq & 3) - 1));
#else
          ev_words[uq >> 4] |= 1u << (8 * ((uq >> 2) & 3) + (uq & 3) - 1);
#endif
        }
        if ((lq & 3) && lq < stop_q) {
#if defined(__CUDA_ARCH__)
          atomicOr(&ev_words[lq >> 4], 1u << (8 * ((lq >> 2) & 3) + (lq & 3) - 1));
#else
          ev_words[lq >> 4] |= 1u << (8 * ((lq >> 2) & 3) + (lq & 3) - 1);
#endif
        }
      }
    }
    n++;
    if (!quad || edge_count(c) == 0 || lq >= stop_q) break;
    // the sweep arrives at the chord's end with x walked through the bands
    fx xw;
    if (tab) xw = rw_walk_x(c.x, c.dx, uq, lq, tab);
    else xw = fx_add(c.x, (fx)(((int64_t)c.dx * (int64_t)(lq - uq)) >> 2));
    const fx next_y = c.lower_y;
    c.x = xw;
    const int w_before = edge_winding(c);
    // WalkEdges: while (lower_y <= next_y) { KeepContinuous; if (!UpdateQuad()) break; }   (sw_raster.cc:630-640)
    bool alive = false;
    while (edge_count(c) > 0) {
      if (!update_quad(c, q, c.x, next_y)) break;
      if (c.lower_y > next_y) {
        alive = true;
        break;
      }
      return -1;  // a chord that runs backwards in y (it flips the winding and restarts from its far end): sequential sweep
    }
    if (!alive) break;
    if (c.upper_y != next_y) return -1;
    (void)w_before;
  }
  return n;
}

// ---- band tables of one op ------------------------------------------------------------------------
// What a row pass reports per row: bit 0 f_ins | bit 1, 2 f_surv under f_in = 0, 1 | bits 3-5, 6-8 the band masks |
// bit 9, 10 failed under f_in = 0, 1 | bits 11-18, 19-26 records emitted.
SKB_HD uint32_t rw_pack(bool f_ins, bool fs0, bool fs1, uint32_t m0, uint32_t m1, bool fail0, bool fail1, int n0, int n1) {
  return (f_ins ? 1u : 0u) | (fs0 ? 2u : 0u) | (fs1 ? 4u : 0u) | (m0 << 3) | (m1 << 6) | (fail0 ? 1u << 9 : 0u) | (fail1 ? 1u << 10 : 0u) |
         ((uint32_t)n0 << 11) | ((uint32_t)n1 << 19);
}

// First tables: bands from the events alone (no forced bands).  tab has n_wrows + 1 entries.
SKB_HDN void rw_bands_from_events(const uint8_t* ev, int n_wrows, int y0q, RowBand* tab) {
  uint32_t nB = 0, nC = 0, nD = 0;
  for (int r = 0; r <= n_wrows; r++) {
    RowBand b;
    b.nB = (uint16_t)nB; b.nC = (uint16_t)nC; b.nD = (uint16_t)nD;
    b.mask = 0;
    b.fl = 0;
    if (r < n_wrows && r * 4 + 4 > y0q) {
      const int p0 = y0q > r * 4 ? y0q - r * 4 : 0;
      b.mask = (uint8_t)rw_band_mask(ev[r] & 7u, p0, false);
      const int cls = rw_row_class(b.mask);
      nB += cls == 1; nC += cls == 2; nD += cls == 3;
    }
    tab[r] = b;
  }
}

// Final tables: composes the rows' f_in -> f_out maps from the top, takes every row's mask and record count under its
// true f_in.  rec_off[r] receives the row's first record relative to the op's.  Returns the op's record count, or
// 0xFFFFFFFF when a row failed under the hypothesis that holds.
SKB_HDN uint32_t rw_bands_resolve(const uint32_t* res, int n_wrows, int y0q, RowBand* tab, uint32_t* rec_off) {
  uint32_t nB = 0, nC = 0, nD = 0, total = 0;
  bool f = false, failed = false;
  for (int r = 0; r <= n_wrows; r++) {
    RowBand b;
    b.nB = (uint16_t)nB; b.nC = (uint16_t)nC; b.nD = (uint16_t)nD;
    b.mask = 0;
    b.fl = 0;
    if (r < n_wrows) rec_off[r] = total;
    if (r < n_wrows && r * 4 + 4 > y0q) {
      const uint32_t v = res[r];
      const int h = f ? 1 : 0;
      failed |= ((v >> (9 + h)) & 1u) != 0;
      b.mask = (uint8_t)((v >> (3 + 3 * h)) & 7u);
      const bool fs = ((v >> (1 + h)) & 1u) != 0;
      b.fl = (uint8_t)((f ? SKB_RB_FIN : 0u) | (fs ? SKB_RB_FSURV : 0u) | ((v & 1u) ? SKB_RB_FINS : 0u));
      total += (v >> (11 + 8 * h)) & 0xFFu;
      f = fs;
      const int cls = rw_row_class(b.mask);
      nB += cls == 1; nC += cls == 2; nD += cls == 3;
    }
    tab[r] = b;
  }
  return failed ? 0xFFFFFFFFu : total;
}

// Position of every live edge slot in the order SortEdges gives the edges (std::sort by (upper_y, x, dx), whose
// arrangement of equal keys is reproduced by sort_edge_indices): decides between edges that start together with
// identical keys.  E = the op's slots as the flatten stage left them; ord = scratch of n_slots ints.
SKB_HDN void rw_sort_ranks(const Edge* E, int n_slots, int32_t* ord, float scan_top_f, float scan_bottom_f, int wide, uint16_t* rank) {
  int n = 0;
  for (int i = 2; i < n_slots; i++) {
    rank[i] = 0;
    if (!((E[i].curve >> 24) & 1)) continue;
    const bool quad = (E[i].curve >> 25) & 1;
    const fx y0 = quad ? E[i].prev : E[i].upper_y, y1 = quad ? E[i].next : E[i].lower_y;
    if (can_be_ignored(scan_top_f, scan_bottom_f, y0, y1, wide)) continue;
    ord[n++] = i;
  }
  sort_edge_indices(E, ord, n);
  for (int i = 0; i < n; i++) rank[ord[i]] = (uint16_t)i;
}

// ---- one row -----------------------------------------------------------------------------------
struct RwEdge {   // an active edge of the row sweep
  fx x, dx, dy;
  int32_t lq;       // quarter row its current chord ends at
  int32_t w;        // winding
  uint32_t cidx;    // current chord (index into the chord table)
  uint32_t cend;    // one past the last chord of its slot
  uint32_t rank;    // position of its slot in SortEdges' order (only when RwRowIn::rank is given)
  uint32_t cbase;   // first chord of its slot (identifies the edge)
  int32_t uq;       // quarter row its current chord starts at
};

struct RwRowIn {
  const uint16_t* rank;    // per slot: position in the reference's std::sort order (rw_sort_ranks); nullptr: not computed — edges with
                           // identical sort keys then fail the path (SKB_RWF_SAME_KEYS), to be retried with ranks
  const SlotInfo* slots;   // the op's slots [0, n_slots)
  int n_slots;
  const Chord* chords;     // global chord table
  const RowBand* tab;      // the op's band table
  const uint8_t* ev;       // the op's event masks, one byte per walk row
  int row;                 // walk row (relative to the origin row)
  int y0q;                 // the sweep's first quarter row
  int stop_q;              // quarter row the sweep stops at (a multiple of 4)
  fx origin_fx;            // origin row << 16
  fx left_clip, right_clip;
  int even_odd;
  bool exact;              // final pass: the chord table is consistent with the band table — check it
};

struct RwRowOut {
  uint32_t mask;     // actual band boundaries of the row
  bool f_ins;        // forcing by the edges inserted at the row's top
  bool f_surv;       // forcing of the next row's first band by the row's last band
  int n_recs;
  int fail;          // 0 = fine; otherwise why the row-parallel form gives up on this path (SKB_RWF_*)
};
#define SKB_RWF_MANY_NEW 1      // more than SKB_RW_MAXN edges start in one row
#define SKB_RWF_SAME_KEYS 2     // edges with identical SortEdges keys
#define SKB_RWF_CHORD_GAP 3     // the slot's chords do not tile its y range
#define SKB_RWF_MANY_ACTIVE 4   // more than SKB_RW_MAXA active edges
#define SKB_RWF_TIE 5           // equal x at the row's top that the slopes do not order
#define SKB_RWF_EVENTS 6        // events of the table differ from the events met
#define SKB_RWF_EMIT 7          // more records than allocated
#define SKB_RWF_CHAIN 8         // chord chain and row sweep disagree
#define SKB_RWF_OPEN 9          // interval open at the end of a band (right edge culled away)
#define SKB_RWF_ORDER 10

// Are the chords cbase_a .. cidx_a and cbase_b .. cidx_b the same sequence (start, end, x, slope)?
SKB_HDN bool rw_twins(const Chord* chords, uint32_t cbase_a, uint32_t cidx_a, uint32_t cbase_b, uint32_t cidx_b) {
  if (cidx_a - cbase_a != cidx_b - cbase_b) return false;
  for (uint32_t k = 0; k <= cidx_a - cbase_a; k++) {
    const Chord a = chords[cbase_a + k], b = chords[cbase_b + k];
    // the last chord may end differently (the twins part ways below this row); everything above must coincide
    const bool last = k == cidx_a - cbase_a;
    if (a.x != b.x || a.dx != b.dx || chord_uq(a.yy) != chord_uq(b.yy) || (!last && chord_lq(a.yy) != chord_lq(b.yy))) return false;
  }
  return true;
}

struct RwRowIn;
SKB_HDN bool rw_x_above(const RwRowIn& in, uint32_t cbase, uint32_t cidx, int q_top, fx* x_out);

// Edge order of SortEdges for edges that start together: (x, dx); a tie beyond that is decided by the introsort's
// permutation of equal keys, which this form does not know.
SKB_HD int rw_new_edge_cmp(const RwEdge& a, const RwEdge& b, bool have_rank) {
  if (a.x != b.x) return a.x < b.x ? -1 : 1;
  if (a.dx != b.dx) return a.dx < b.dx ? -1 : 1;
  if (have_rank && a.rank != b.rank) return a.rank < b.rank ? -1 : 1;
  return 0;
}

// insert_new_edges (sw_raster.cc:207-247) for the edges M[0..nM) that start at the current y, M in SortEdges order.
// Returns true when one of them makes check_intersection fire.
SKB_HDN bool rw_insert_new(RwEdge* A, int& nA, const RwEdge* M, int nM, int& fail) {
  bool forced = false;
  if (nM == 0) return false;
  if (nA + nM > SKB_RW_MAXA) {
    fail = SKB_RWF_MANY_ACTIVE;
    return false;
  }
  // backward_insert_start: from the last active edge back to the first whose x is not greater
  int s = nA - 1;
  while (s >= 0 && A[s].x > M[0].x) s--;
  for (int k = 0; k < nM; k++) {
    const RwEdge m = M[k];
    while (s + 1 < nA && A[s + 1].x < m.x) s++;
    for (int j = nA; j > s + 1; j--) A[j] = A[j - 1];
    A[s + 1] = m;
    nA++;
    if (s >= 0 && fx_add(A[s].x, A[s].dx) > fx_add(m.x, m.dx)) forced = true;
    s = s + 1;
  }
  return forced;
}

// x of an edge at the top of the band that ends at q_top (a row's top): the band above is the last band of the row above,
// whose top follows from that row's band mask.  false when the edge was not active there.
SKB_HDN bool rw_x_above(const RwRowIn& in, uint32_t cbase, uint32_t cidx, int q_top, fx* x_out) {
  const uint32_t m = in.tab[in.row - 1].mask & 7u;
  int q_prev = q_top - 4 + (m & 4u ? 3 : m & 2u ? 2 : m & 1u ? 1 : 0);
  if (q_prev < in.y0q) q_prev = in.y0q;
  Chord c = in.chords[cidx];
  if (chord_uq(c.yy) >= q_top) {   // the chord starts at q_top: the one before it was active above
    if (cidx == cbase) return false;
    c = in.chords[cidx - 1];
  }
  if (chord_uq(c.yy) > q_prev || chord_lq(c.yy) < q_top) return false;
  *x_out = rw_walk_x(c.x, c.dx, chord_uq(c.yy), q_prev, in.tab);
  return true;
}

// Sweeps one walk row.  f_in: the row's first band is forced to ¼ by the LAST band of the row above (the forcing by the
// edges inserted at this row's top is found here).  emit != nullptr: records are written to emit[0..) in the sweep's order.
SKB_HDN void rw_row(const RwRowIn& in, bool f_in, TrapRec* emit, int emit_cap, RwRowOut& out) {
  out.mask = 0;
  out.f_ins = out.f_surv = false;
  out.n_recs = 0;
  out.fail = 0;
  const int q_row = in.row * 4;
  const int q_end = q_row + 4;
  if (q_end <= in.y0q || q_row >= in.stop_q) return;  // above the first edge / below the last row: nothing happens
  const bool first_row = in.y0q >= q_row;
  const int q_top = first_row ? in.y0q : q_row;
  RwEdge A[SKB_RW_MAXA];
  RwEdge P[SKB_RW_MAXN];   // edges that start in this row (at its top included), by (start, SortEdges order)
  int P_uq[SKB_RW_MAXN];
  int nA = 0, nP = 0;
  // ---- gather
  for (int sidx = 0; sidx < in.n_slots; sidx++) {
    const SlotInfo si = in.slots[sidx];
    if (si.n_chords == 0 || si.q_first >= q_end || si.q_last <= q_top) continue;
    RwEdge e;
    e.cend = si.chord_base + si.n_chords;
    e.rank = in.rank ? in.rank[sidx] : 0u;
    e.cbase = si.chord_base;
    if (si.q_first >= q_top) {
      const Chord c = in.chords[si.chord_base];
      e.x = c.x; e.dx = c.dx; e.dy = c.dy;
      e.lq = chord_lq(c.yy);
      e.w = chord_w(c.yy);
      e.cidx = si.chord_base;
      e.uq = si.q_first;
      if (nP >= SKB_RW_MAXN) { out.fail = SKB_RWF_MANY_NEW; return; }
      // insertion sort by (start, x, dx)
      int pos = nP;
      while (pos > 0) {
        const int cmp = P_uq[pos - 1] != si.q_first ? (P_uq[pos - 1] < si.q_first ? -1 : 1) : rw_new_edge_cmp(P[pos - 1], e, in.rank != nullptr);
        if (cmp == 0) { out.fail = SKB_RWF_SAME_KEYS; return; }   // identical sort keys: the reference's order is std::sort's whim
        if (cmp < 0) break;
        pos--;
      }
      for (int j = nP; j > pos; j--) { P[j] = P[j - 1]; P_uq[j] = P_uq[j - 1]; }
      P[pos] = e;
      P_uq[pos] = si.q_first;
      nP++;
    } else {
      // the chord active at q_top: first chord whose lower end is below it
      uint32_t lo = 0, hi = si.n_chords - 1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (chord_lq(in.chords[si.chord_base + mid].yy) > q_top) hi = mid; else lo = mid + 1;
      }
      const Chord c = in.chords[si.chord_base + lo];
      const int uq = chord_uq(c.yy);
      if (uq > q_top || chord_lq(c.yy) <= q_top) { out.fail = SKB_RWF_CHORD_GAP; return; }
      e.dx = c.dx; e.dy = c.dy;
      e.lq = chord_lq(c.yy);
      e.w = chord_w(c.yy);
      e.cidx = si.chord_base + lo;
      e.uq = uq;
      e.x = rw_walk_x(c.x, c.dx, uq, q_top, in.tab);
      if (nA >= SKB_RW_MAXA) { out.fail = SKB_RWF_MANY_ACTIVE; return; }
      // the list at the row's top is x-sorted; equal x: the edge that came from the left (larger slope) is first —
      // decidable when the slopes differ by at least one unit per quarter band, otherwise the order is history
      int pos = nA;
      while (pos > 0) {
        if (A[pos - 1].x < e.x) break;
        if (A[pos - 1].x == e.x) {
          // Twins — two edges whose chords so far are identical one by one (same start, same x, same slope): they have
          // been one on top of the other in every band, the stable re-sorting has kept them in the order they were
          // inserted in, which is SortEdges' (the ranks).
          if (in.rank != nullptr && rw_twins(in.chords, A[pos - 1].cbase, A[pos - 1].cidx, e.cbase, e.cidx)) {
            if (A[pos - 1].rank < e.rank) break;
            pos--;
            continue;
          }
          // Otherwise the re-sorting at the end of the band above left them in the order they had at that band's top
          // (it moves an edge only past edges with a strictly greater x): compare their x there.
          if (first_row || in.row == 0) { out.fail = SKB_RWF_TIE; return; }
          fx xa, xb;
          if (!rw_x_above(in, A[pos - 1].cbase, A[pos - 1].cidx, q_top, &xa) || !rw_x_above(in, e.cbase, e.cidx, q_top, &xb)) {
            out.fail = SKB_RWF_TIE;
            return;
          }
          if (xa == xb) {
            // coincident above as well: twins up to the chords that end here?
            const uint32_t ia = chord_uq(in.chords[A[pos - 1].cidx].yy) >= q_top ? A[pos - 1].cidx - 1 : A[pos - 1].cidx;
            const uint32_t ib = chord_uq(in.chords[e.cidx].yy) >= q_top ? e.cidx - 1 : e.cidx;
            if (in.rank == nullptr || !rw_twins(in.chords, A[pos - 1].cbase, ia, e.cbase, ib)) { out.fail = SKB_RWF_TIE; return; }
            if (A[pos - 1].rank < e.rank) break;
            pos--;
            continue;
          }
          if (xa < xb) break;
        }
        pos--;
      }
      for (int j = nA; j > pos; j--) A[j] = A[j - 1];
      A[pos] = e;
      nA++;
    }
  }
  // ---- the row's bands
  const uint32_t ev = in.ev[in.row] & 7u;
  const int mask_w = in.even_odd ? 1 : -1;
  int q = q_top;
  int p_next = 0;            // next pending edge
  bool forced = f_in;
  // edges that start at the top: the sweep's initial list (first row: sorted order, no checks) or insert_new_edges
  {
    int n_top = 0;
    while (p_next + n_top < nP && P_uq[p_next + n_top] == q_top) n_top++;
    if (n_top) {
      const bool fi = rw_insert_new(A, nA, P + p_next, n_top, out.fail);
      if (out.fail) return;
      if (!first_row) {
        out.f_ins = fi;
        forced = forced || fi;
      }
      p_next += n_top;
    }
  }
  while (q < q_end) {
    // next event
    int nxt = q_end;
    for (int k = 0; k < nA; k++) nxt = A[k].lq < nxt ? A[k].lq : nxt;
    if (p_next < nP && P_uq[p_next] < nxt) nxt = P_uq[p_next];
    {
      // the events of the table are the events met here (the chords' y do not depend on anything the passes vary)
      const uint32_t above = (ev | 8u) >> (q & 3);
      int h_ev = 1;
      while (!((above >> (h_ev - 1)) & 1u)) h_ev++;
      if (q + h_ev != nxt) { out.fail = SKB_RWF_EVENTS; return; }
    }
    int h = nxt - q;
    if (forced || (h & 1)) h = 1;
    const int y_shift = h == 1 ? 2 : h == 2 ? 1 : 0;
    const uint32_t full = h == 1 ? 64u : h == 2 ? 128u : 255u;   // fixed_to_alpha(h / 4) (sw_raster.cc:151-153,249)
    const int next_q = q + h;
    if (next_q < q_end) out.mask |= 1u << ((next_q & 3) - 1);
    forced = false;
    // ---- the band: active edges in list order
    int w = 0, nN = 0, prev_right = fx_floor_i(in.left_clip);
    bool in_interval = false;
    fx left = in.left_clip, left_dy = 0;
    fx left_edge_x = 0;   // advanced x of the current left edge
    // the left edge itself, for the band that ends with its interval still open: which edge, whether it ended in this band
    // and then which edge preceded it in the list at that moment (0xFFFFFFFF: the head), its slope when it ended
    uint32_t left_id = 0xFFFFFFFFu, left_prev_id = 0xFFFFFFFFu;
    bool left_removed = false;
    fx left_old_dx = 0;
    const int y_px = fx_floor_i(in.origin_fx) + (q >> 2);
    for (int k = 0; k < nA; k++) {
      RwEdge e = A[k];
      w += e.w;
      const bool prev_in = in_interval;
      in_interval = (w & mask_w) != 0;
      const bool is_left = in_interval && !prev_in, is_right = !in_interval && prev_in;
      const fx old_x = e.x;
      e.x = fx_add(e.x, e.dx >> y_shift);
      if (is_left) {
        left = fx_max(old_x, in.left_clip);
        left_dy = e.dy;
        left_edge_x = e.x;
        left_id = e.cbase;
        left_removed = false;
      } else if (is_right) {
        const fx right = fx_min(in.right_clip, old_x);
        TrapRec r;
        r.y = y_px;
        r.ul = left;
        r.ur = right;
        r.ll = fx_max(in.left_clip, left_edge_x);
        r.lr = fx_min(in.right_clip, e.x);
        r.ldy = left_dy;
        r.rdy = e.dy;
        bool no_real = false;
        if (full == 0xFF) {
          no_real = prev_right > fx_floor_i(left) || prev_right > fx_floor_i(left_edge_x);
          if (!no_real && k + 1 < nA) {   // too_close_edges(cur, cur->next, next_y): the next active edge, not yet advanced
            const RwEdge& nx = A[k + 1];
            no_real = fx_add(e.x, SKB_FX1) >= fx_sub(nx.x, fx_abs(nx.dx));
          }
        }
        r.flags = full | (no_real ? 0x100u : 0u);
        if (emit) {
          if (out.n_recs < emit_cap) emit[out.n_recs] = r;
          else { out.fail = SKB_RWF_EMIT; return; }
        }
        out.n_recs++;
        prev_right = fx_ceil_i(fx_max(right, e.x));
      }
      // chord end
      bool removed = false;
      if (e.lq <= next_q) {
        if (e.cidx + 1 < e.cend) {
          const Chord c = in.chords[e.cidx + 1];
          if (chord_uq(c.yy) != next_q || chord_lq(c.yy) <= next_q) { out.fail = SKB_RWF_CHORD_GAP; return; }
          if (in.exact && c.x != e.x) { out.fail = SKB_RWF_CHAIN; return; }   // the chain and the row sweep walked the same bands
          e.dx = c.dx; e.dy = c.dy;
          e.lq = chord_lq(c.yy);
          e.w = chord_w(c.yy);
          e.uq = next_q;
          e.cidx++;
        } else {
          removed = true;
        }
      }
      if (removed) {
        if (e.cbase == left_id) {   // remove_edge leaves the edge's own links as they were: its predecessor at this moment
          left_removed = true;
          left_prev_id = nN > 0 ? A[nN - 1].cbase : 0xFFFFFFFFu;
          left_old_dx = e.dx;
        }
        continue;
      }
      // re-sort (backward_insert_edge_based_on_x) and check_intersection against the new list predecessor
      int pos = nN;
      while (pos > 0 && A[pos - 1].x > e.x) pos--;
      if (pos > 0 && fx_add(A[pos - 1].x, A[pos - 1].dx) > fx_add(e.x, e.dx)) forced = true;
      for (int j = nN; j > pos; j--) A[j] = A[j - 1];
      A[pos] = e;
      nN++;
    }
    if (in_interval) {
      // the right edge of the interval was culled away (an edge wholly outside the scan rows, SWEdge::CanBeIgnored): the
      // interval runs to the right clip (sw_raster.cc:661-668)
      TrapRec r;
      r.y = y_px;
      r.ul = left;
      r.ur = in.right_clip;
      r.ll = fx_max(in.left_clip, left_edge_x);
      r.lr = in.right_clip;
      r.ldy = left_dy;
      r.rdy = 0;
      bool no_real = false;
      if (full == 0xFF) {
        // edges_too_close(left_edge->prev, left_edge, next_y)
        fx prev_x = in.left_clip, le_dx = left_old_dx;
        bool started_before = true;
        const uint32_t want = left_removed ? left_prev_id : left_id;
        bool found = left_removed && left_prev_id == 0xFFFFFFFFu;
        for (int j = 0; j < nN; j++) {
          if (A[j].cbase != want) continue;
          found = true;
          if (left_removed) {
            prev_x = A[j].x;
          } else {
            prev_x = j > 0 ? A[j - 1].x : in.left_clip;
            le_dx = A[j].dx;
            started_before = A[j].uq < next_q;
          }
        }
        if (!found) { out.fail = SKB_RWF_OPEN; return; }
        no_real = started_before && fx_add(prev_x, SKB_FX1) >= fx_sub(left_edge_x, fx_abs(le_dx));
      }
      r.flags = full | (no_real ? 0x100u : 0u);
      if (emit) {
        if (out.n_recs < emit_cap) emit[out.n_recs] = r;
        else { out.fail = SKB_RWF_EMIT; return; }
      }
      out.n_recs++;
    }
    nA = nN;
    q = next_q;
    if (q >= q_end) break;
    // insert_new_edges at the band boundary
    {
      int n_new = 0;
      while (p_next + n_new < nP && P_uq[p_next + n_new] == q) n_new++;
      if (n_new) {
        const bool fi = rw_insert_new(A, nA, P + p_next, n_new, out.fail);
        if (out.fail) return;
        forced = forced || fi;
        p_next += n_new;
      } else if (p_next < nP && P_uq[p_next] < q) {
        out.fail = SKB_RWF_ORDER;
        return;
      }
    }
  }
  out.f_surv = forced;
}

}  // namespace skb

#endif  // SKB_ROWWALK_CUH
