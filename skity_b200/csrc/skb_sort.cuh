// skb_sort.cuh — std::sort as libstdc++ performs it (GCC 13 bits/stl_algo.h: introsort with threshold 16,
// median-of-three moved to the front, unguarded partition, heapsort at the depth limit, final insertion sort), on an
// array of values with a strict-weak `less`.  The reference sorts span lists with comparators that leave ties
// (spans_subtraction sorts by x alone, src/render/sw/sw_canvas.cc:71-72) and what it does next depends on how the
// ties come out, so the permutation itself has to be reproduced.  Same algorithm as sort_edge_indices (skb_walk.cuh),
// which sorts an index array; pure per-thread code.
#ifndef SKB_SORT_CUH
#define SKB_SORT_CUH

#include "skity_b200/csrc/skb_core.cuh"

namespace skb {

template <class T, class Less>
SKB_HDN void sr_unguarded_linear_insert(T* v, int last, Less less) {
  const T val = v[last];
  int next = last - 1;
  while (less(val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}
template <class T, class Less>
SKB_HDN void sr_insertion_sort(T* v, int first, int last, Less less) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(v[i], v[first])) {
      const T val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1];
      v[first] = val;
    } else {
      sr_unguarded_linear_insert(v, i, less);
    }
  }
}
template <class T, class Less>
SKB_HDN void sr_adjust_heap(T* v, int first, int hole, int len, T value, Less less) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(v[first + child], v[first + child - 1])) child--;
    v[first + hole] = v[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[first + hole] = v[first + child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && less(v[first + parent], value)) {
    v[first + hole] = v[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[first + hole] = value;
}
template <class T, class Less>
SKB_HDN void sr_heap_sort(T* v, int first, int last, Less less) {
  const int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      sr_adjust_heap(v, first, parent, len, v[first + parent], less);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    const T val = v[last];
    v[last] = v[first];
    sr_adjust_heap(v, first, 0, last - first, val, less);
  }
}
template <class T, class Less>
SKB_HDN void std_sort_replica(T* v, int n, Less less) {
  if (n <= 1) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) lg++;
  // __introsort_loop recurses on [cut, last) and loops on [first, cut): the ranges are disjoint, so an explicit stack
  // that does the left part first gives the same arrangement
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 1;
  stack_first[0] = 0;
  stack_last[0] = n;
  stack_depth[0] = lg * 2;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
        sr_heap_sort(v, first, last, less);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      const int a = first + 1, b = mid, c = last - 1;
      int pick;
      if (less(v[a], v[b])) {
        if (less(v[b], v[c])) pick = b;
        else if (less(v[a], v[c])) pick = c;
        else pick = a;
      } else if (less(v[a], v[c])) pick = a;
      else if (less(v[b], v[c])) pick = c;
      else pick = b;
      { const T t = v[first]; v[first] = v[pick]; v[pick] = t; }
      int lo = first + 1, hi = last;
      for (;;) {
        while (less(v[lo], v[first])) ++lo;
        --hi;
        while (less(v[first], v[hi])) --hi;
        if (!(lo < hi)) break;
        const T t = v[lo]; v[lo] = v[hi]; v[hi] = t;
        ++lo;
      }
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
    sr_insertion_sort(v, 0, 16, less);
    for (int i = 16; i != n; ++i) sr_unguarded_linear_insert(v, i, less);
  } else {
    sr_insertion_sort(v, 0, n, less);
  }
}

}  // namespace skb

#endif  // SKB_SORT_CUH
