// tests/sim/walk_nested.hpp — test infrastructure, not product code.
//
// The reference's band sweep in its own shape (WalkEdges, src/render/sw/sw_raster.cc:546-677: a loop over bands
// around a loop over the active edges), on the product's edge slots.  The CPU simulation runs it beside the flat
// single-loop form the GPU runs (skb_walk.cuh walk_bands_flat) and compares the trapezoid records one by one
// (SKB_SIM_WALK_MODE=0).
#ifndef SKB_TESTS_SIM_WALK_NESTED_HPP
#define SKB_TESTS_SIM_WALK_NESTED_HPP

#include <vector>

#include "skity_b200/csrc/skb_walk.cuh"

namespace skb {

// The band loop as the reference writes it: a loop over bands around a loop over the active edges.
SKB_HDN void walk_bands_nested(Edge* E, QuadState* Q, const uint16_t* qmap, WalkState ws, int stop_y, fx left_clip,
                               fx right_clip, int even_odd, RecSink& sink) {
  Edge& H = E[SKB_HEAD];
  fx y = ws.y, nny = ws.nny;
  const int mask = even_odd ? 1 : -1;
  for (;;) {
    int w = 0;
    bool in_interval = false;
    fx prev_x = H.x;
    fx next_y = fx_min(nny, fx_ceil_fx(fx_add(y, 1)));
    int cur = H.next, left_edge = SKB_HEAD;
    fx left = left_clip, left_dy = 0;
    int prev_right = fx_floor_i(left_clip);
    nny = SKB_FX_MAX;
    int y_shift = 0;
    if (fx_sub(next_y, y) & (SKB_FX1 >> 2)) {
      y_shift = 2;
      next_y = fx_add(y, SKB_FX1 >> 2);
    } else if (fx_sub(next_y, y) & (SKB_FX1 >> 1)) {
      y_shift = 1;
    }
    // fixed_to_alpha(next_y - y) = SWFixedRoundToInt(0xFF * h) (sw_raster.cc:151-153,249)
    const uint32_t full = (uint32_t)(uint8_t)fx_round_i((fx)(0xFF * fx_sub(next_y, y)));
    while (E[cur].upper_y <= y) {
      Edge& c = E[cur];
      w += edge_winding(c);
      bool prev_in = in_interval;
      in_interval = (w & mask) != 0;
      bool is_left = in_interval && !prev_in, is_right = !in_interval && prev_in;
      if (is_left) {
        left = fx_max(c.x, left_clip);
        left_dy = c.dy;
        left_edge = cur;
        c.x = fx_add(c.x, c.dx >> y_shift);
      } else if (is_right) {
        fx right = fx_min(right_clip, c.x);
        c.x = fx_add(c.x, c.dx >> y_shift);
        TrapRec r;
        r.y = y >> 16;
        r.ul = left;
        r.ur = right;
        r.ll = fx_max(left_clip, E[left_edge].x);
        r.lr = fx_min(right_clip, c.x);
        r.ldy = left_dy;
        r.rdy = c.dy;
        bool no_real = full == 0xFF && ((prev_right > fx_floor_i(left) || prev_right > fx_floor_i(E[left_edge].x)) ||
                                        too_close_edges(E, cur, c.next, next_y));
        r.flags = full | (no_real ? 0x100u : 0u);
        sink_emit(sink, r);
        prev_right = fx_ceil_i(fx_max(right, c.x));
      } else {
        c.x = fx_add(c.x, c.dx >> y_shift);
      }
      int next = c.next;
      while (c.lower_y <= next_y) {
        if (edge_count(c) > 0) {
          QuadState& q = Q[qmap ? (int)qmap[cur] : cur];
          // SWQuadEdge::KeepContinuous (sw_edge.cc:294-297): the next chord starts where the sweep has brought the edge
          if (!update_quad(c, q, c.x, next_y)) break;
        } else {
          break;
        }
      }
      if (c.lower_y <= next_y) {
        remove_edge(E, cur);
      } else {
        upd_nny(c.lower_y, next_y, &nny);
        fx new_x = c.x;
        if (new_x < prev_x) backward_insert_on_x(E, cur);
        else prev_x = new_x;
        check_intersection(E, cur, next_y, &nny);
      }
      cur = next;
    }
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
    if (y >= i_to_fx(stop_y)) break;
    insert_new_edges(E, cur, y, &nny);
  }
  sink_flush_row(sink);
}


// mode 0: nested loops (cross-check); otherwise the flat loop
inline void sim_walk_path(Edge* E, QuadState* Q, int n_slots, int32_t* ord, float scan_top_f, float scan_bottom_f, int start_y,
                          int stop_y, fx left_clip, fx right_clip, int even_odd, RecSink& sink, int mode, int wide) {
  if (mode == 2) {  // the flat loop on the compact copy in sweep order (what k_walk runs)
    std::vector<Edge> E2((size_t)n_slots * 2);
    walk_path(E, Q, n_slots, ord, scan_top_f, scan_bottom_f, start_y, stop_y, left_clip, right_clip, even_odd, sink, wide,
              E2.data(), nullptr);
    return;
  }
  if (mode != 0) {
    walk_path(E, Q, n_slots, ord, scan_top_f, scan_bottom_f, start_y, stop_y, left_clip, right_clip, even_odd, sink, wide);
    return;
  }
  WalkState ws;
  if (!walk_prologue(E, Q, n_slots, ord, scan_top_f, scan_bottom_f, start_y, left_clip, right_clip, ws, wide)) return;
  walk_bands_nested(E, Q, nullptr, ws, stop_y, left_clip, right_clip, even_odd, sink);
}

}  // namespace skb

#endif  // SKB_TESTS_SIM_WALK_NESTED_HPP
