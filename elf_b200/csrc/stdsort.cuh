// stdsort.cuh -- the permutation libstdc++'s std::sort produces, for one thread.
//
// MCTSActor::pi2response (src_cpp/elfgames/go/mcts/mcts.h:289-295) sorts the 362 (move, probability)
// pairs of a network reply with std::sort and a comparator on the probability alone; the order of moves
// with EQUAL probabilities -- and through it the insertion order of the edges into the node's
// unordered_map, i.e. every container-order tie-break later on -- is whatever that library's introsort does
// with them.  Half-precision networks make equal probabilities the norm (equal fp16 logits), so a search
// that is to reproduce the reference's games bit for bit has to reproduce this permutation as well
// (option std_sort_ties, k_expand<N, true>).  libstdc++ is not part of the reference tree; its published
// algorithm (GCC <bits/stl_algo.h>, <bits/stl_heap.h>, unchanged here since GCC 4.x) is restated:
// std::__sort = __introsort_loop (median of first+1 / middle / last-1 moved to first, unguarded Hoare
// partition around *first, recursion on the right part, loop on the left, depth limit 2*floor(log2 n), heap
// sort of a range when it reaches 0) + __final_insertion_sort (threshold 16).  Elements are 16-bit indices
// into a key array, comp(a, b) = key[a] > key[b] (descending probability: the bit patterns of
// non-negative floats order like the values).  Sequential by nature: one lane runs it on shared memory,
// only for leaves that actually hold equal probabilities.  Pinned against the real std::sort (this file
// compiled for the host by tests/cxx/stdsort_check.cc: random arrays full of duplicates, adversarial inputs
// that reach the heap sort) and through the search (tests/test_emu_kernels.py::test_equal_priors_follow_std_sort).
#pragma once
#include <cstdint>

namespace elfb200 {

struct StdSortCtx {
  const uint32_t* key;  // key[element]
};

#define SS_COMP(c, a, b) ((c)->key[(a)] > (c)->key[(b)])

__device__ inline void ss_swap(uint16_t* x, uint16_t* y) {
  const uint16_t t = *x;
  *x = *y;
  *y = t;
}

// ---- <bits/stl_heap.h> ----
__device__ inline void ss_push_heap(StdSortCtx* c, uint16_t* first, int hole, int top, uint16_t value) {
  int parent = (hole - 1) / 2;
  while (hole > top && SS_COMP(c, first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

__device__ inline void ss_adjust_heap(StdSortCtx* c, uint16_t* first, int hole, int len, uint16_t value) {
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (SS_COMP(c, first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  ss_push_heap(c, first, hole, top, value);
}

/* std::__partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap */
__device__ inline void ss_heap_sort(StdSortCtx* c, uint16_t* first, int len) {
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const uint16_t value = first[parent];
      ss_adjust_heap(c, first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) { /* __sort_heap: __pop_heap(first, last, last) */
    --last;
    const uint16_t value = first[last];
    first[last] = first[0];
    ss_adjust_heap(c, first, 0, last, value);
  }
}

// ---- <bits/stl_algo.h> ----
__device__ inline void ss_move_median_to_first(StdSortCtx* c, uint16_t* result, uint16_t* a, uint16_t* b, uint16_t* cc) {
  if (SS_COMP(c, *a, *b)) {
    if (SS_COMP(c, *b, *cc))
      ss_swap(result, b);
    else if (SS_COMP(c, *a, *cc))
      ss_swap(result, cc);
    else
      ss_swap(result, a);
  } else if (SS_COMP(c, *a, *cc)) {
    ss_swap(result, a);
  } else if (SS_COMP(c, *b, *cc)) {
    ss_swap(result, cc);
  } else {
    ss_swap(result, b);
  }
}

__device__ inline int ss_unguarded_partition(StdSortCtx* c, uint16_t* v, int first, int last, int pivot) {
  for (;;) {
    while (SS_COMP(c, v[first], v[pivot])) ++first;
    --last;
    while (SS_COMP(c, v[pivot], v[last])) --last;
    if (!(first < last)) return first;
    ss_swap(&v[first], &v[last]);
    ++first;
  }
}

__device__ inline void ss_unguarded_linear_insert(StdSortCtx* c, uint16_t* v, int last) {
  const uint16_t val = v[last];
  int next = last - 1;
  while (SS_COMP(c, val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}

__device__ inline void ss_insertion_sort(StdSortCtx* c, uint16_t* v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (SS_COMP(c, v[i], v[first])) {
      const uint16_t val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1]; /* move_backward(first, i, i + 1) */
      v[first] = val;
    } else {
      ss_unguarded_linear_insert(c, v, i);
    }
  }
}

/* v[0..n): the elements (indices into key) in input order; sorted in place as std::sort would */
__device__ inline void ss_sort(StdSortCtx* c, uint16_t* v, int n) {
  if (n <= 0) return;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) lg++; /* std::__lg */
  /* __introsort_loop: explicit stack of the right-hand parts (they are disjoint: any order) */
  int stack_first[24], stack_last[24], stack_depth[24], sp = 0;  // depth <= 2*floor(log2 n) + 1 = 17 for n <= 511
  stack_first[0] = 0, stack_last[0] = n, stack_depth[0] = 2 * lg, sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
        ss_heap_sort(c, v + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      ss_move_median_to_first(c, &v[first], &v[first + 1], &v[mid], &v[last - 1]);
      const int cut = ss_unguarded_partition(c, v, first + 1, last, first);
      stack_first[sp] = cut, stack_last[sp] = last, stack_depth[sp] = depth, sp++;
      last = cut;
    }
  }
  /* __final_insertion_sort */
  if (n > 16) {
    ss_insertion_sort(c, v, 0, 16);
    for (int i = 16; i != n; ++i) ss_unguarded_linear_insert(c, v, i);
  } else {
    ss_insertion_sort(c, v, 0, n);
  }
}

}  // namespace elfb200
