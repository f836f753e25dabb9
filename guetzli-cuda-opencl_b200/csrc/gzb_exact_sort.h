// gzb_exact_sort.h -- libstdc++'s std::sort, restated, for host and device code.
//
// The reference orders candidates with std::sort on float keys in two places: the <= 189 non-zero AC
// coefficients of a block (guetzli/processor.cc:410-412) and the back end's global_order
// (processor.cc:825-828). std::sort is not stable, so wherever two keys are equal the arrangement the
// search sees is the one the LIBRARY produces, and the golden files come from a libstdc++ build:
//   std::sort(first, last)  =  __introsort_loop(first, last, 2 * floor(log2(n)))   [bits/stl_algo.h]
//                              + __final_insertion_sort(first, last)
//   __introsort_loop: while the range is longer than 16: depth budget used up -> heap sort the range
//       (__partial_sort(first, last, last) = __make_heap + __sort_heap); else median-of-three of
//       (first + 1, middle, last - 1) to `first`, unguarded Hoare partition around it, recurse into the
//       right part, loop on the left part;
//   __final_insertion_sort: insertion sort (each element moves left past the strictly greater ones).
// This file is that algorithm written out (no calls into the library's private functions), generic over
// the element type, compiled for the host (gzb_encoder.cc: the short ranges of the back end's order) and
// for the device (gzb_zeroing.cuh: a block whose keys tie). tests/test_host_cpu.py checks it against the
// std::sort / std::partial_sort of the build.
#pragma once
#include <cstddef>

#ifdef __CUDACC__
#define GZB_HD __host__ __device__ __forceinline__
#else
#define GZB_HD inline
#endif

namespace gzb {
namespace xsort {

template <typename E>
GZB_HD void swap_e(E& a, E& b) { const E t = a; a = b; b = t; }

// std::__move_median_to_first(result, a, b, c)
template <typename E, typename Less>
GZB_HD void median_to_first(E* result, E* a, E* b, E* c, Less less) {
  if (less(*a, *b)) {
    if (less(*b, *c)) swap_e(*result, *b);
    else if (less(*a, *c)) swap_e(*result, *c);
    else swap_e(*result, *a);
  } else if (less(*a, *c)) swap_e(*result, *a);
  else if (less(*b, *c)) swap_e(*result, *c);
  else swap_e(*result, *b);
}

// std::__unguarded_partition_pivot(first, last)
template <typename E, typename Less>
GZB_HD E* partition_pivot(E* first, E* last, Less less) {
  E* mid = first + (last - first) / 2;
  median_to_first(first, first + 1, mid, last - 1, less);
  const E* pivot = first;
  E* lo = first + 1;
  E* hi = last;
  for (;;) {
    while (less(*lo, *pivot)) ++lo;
    --hi;
    while (less(*pivot, *hi)) --hi;
    if (!(lo < hi)) return lo;
    swap_e(*lo, *hi);
    ++lo;
  }
}

// std::__insertion_sort
template <typename E, typename Less>
GZB_HD void insertion_sort(E* first, E* last, Less less) {
  for (E* i = first; i < last; ++i) {
    const E v = *i;
    E* j = i;
    while (j > first && less(v, *(j - 1))) { *j = *(j - 1); --j; }
    *j = v;
  }
}

// std::__push_heap / std::__adjust_heap (bits/stl_heap.h)
template <typename E, typename Less>
GZB_HD void push_heap_at(E* first, ptrdiff_t hole, ptrdiff_t top, const E value, Less less) {
  ptrdiff_t parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
template <typename E, typename Less>
GZB_HD void adjust_heap(E* first, ptrdiff_t hole, ptrdiff_t len, const E value, Less less) {
  const ptrdiff_t top = hole;
  ptrdiff_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(first[child], first[child - 1])) --child;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  push_heap_at(first, hole, top, value, less);
}
// std::__partial_sort(first, last, last): __make_heap, then __sort_heap
template <typename E, typename Less>
GZB_HD void heap_sort(E* first, E* last, Less less) {
  const ptrdiff_t len = last - first;
  if (len < 2) return;
  for (ptrdiff_t parent = (len - 2) / 2;; --parent) {
    const E v = first[parent];
    adjust_heap(first, parent, len, v, less);
    if (parent == 0) break;
  }
  while (last - first > 1) {
    --last;
    const E v = *last;          // __pop_heap(first, last, last)
    *last = *first;
    adjust_heap(first, static_cast<ptrdiff_t>(0), last - first, v, less);
  }
}

// std::__introsort_loop(first, last, depth) without recursion: the right-hand parts wait on a stack
// (at most one per level, and the budget is 2 * log2(n) levels).
template <int kStack, typename E, typename Less>
GZB_HD void introsort_loop(E* first, E* last, int depth, Less less) {
  struct Pending { E* first; E* last; int depth; };
  Pending stack[kStack];   // kStack > depth
  int top = 0;
  for (;;) {
    while (last - first > 16) {
      if (depth == 0) { heap_sort(first, last, less); break; }
      --depth;
      E* cut = partition_pivot(first, last, less);
      // the library recurses into [cut, last) first; the two parts are disjoint, so the order in which they
      // are finished does not matter
      stack[top].first = cut; stack[top].last = last; stack[top].depth = depth;
      ++top;
      last = cut;
    }
    if (top == 0) return;
    --top;
    first = stack[top].first; last = stack[top].last; depth = stack[top].depth;
  }
}

// Final arrangement of a range std::sort's partitioning has isolated with `depth` budget left. The closing
// insertion sort never moves an entry across a partition boundary (everything on the left is <= the pivot
// <= everything on the right, and entries only pass strictly greater ones), so it can be run per range.
template <int kStack, typename E, typename Less>
GZB_HD void finish_range(E* first, E* last, int depth, Less less) {
  if (last - first > 1) {
    introsort_loop<kStack>(first, last, depth, less);
    insertion_sort(first, last, less);
  }
}

GZB_HD int depth_budget(size_t n) {   // 2 * std::__lg(n)
  int lg = 0;
  while ((n >> (lg + 1)) != 0) ++lg;
  return 2 * lg;
}

// std::sort(first, last, less); kStack > 2 * log2(last - first)
template <int kStack, typename E, typename Less>
GZB_HD void sort(E* first, E* last, Less less) {
  if (first != last) finish_range<kStack>(first, last, depth_budget(static_cast<size_t>(last - first)), less);
}

}  // namespace xsort
}  // namespace gzb
