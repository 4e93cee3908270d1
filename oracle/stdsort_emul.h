/* oracle/stdsort_emul.h -- TEST INFRASTRUCTURE ONLY (part of the C restatement, never linked into the
 * product).
 *
 * The permutation libstdc++'s std::sort produces, restated: MCTSActor::pi2response
 * (src_cpp/elfgames/go/mcts/mcts.h:289-295) sorts the 362 (move, probability) pairs of a network reply
 * with std::sort and a comparator on the probability alone, so the order of moves with EQUAL
 * probabilities -- and through it the insertion order of the edges into the node's unordered_map -- is
 * whatever that library's introsort does with them.  libstdc++ (GCC's <bits/stl_algo.h>, <bits/stl_heap.h>;
 * unchanged in this part since GCC 4.x; the reference is built against the system's copy) is not part of
 * /root/reference, so its published algorithm is restated here: std::__sort = __introsort_loop (median of
 * three of first+1 / middle / last-1 moved to first, unguarded Hoare partition around *first, recursion on
 * the right part, loop on the left, depth limit 2*floor(log2 n), heap sort of a range when it hits 0)
 * followed by __final_insertion_sort (threshold 16).  Elements are indices into a key array and
 * comp(a, b) = key[a] > key[b] (descending probability; probabilities are non-negative floats, whose bit
 * patterns order like the values).  Pinned against the real std::sort by tests/test_stdsort_emul.py
 * (random arrays full of duplicates, sizes 0..400, and median-of-three killer inputs that reach the heap
 * sort). */
#ifndef ORACLE_STDSORT_EMUL_H
#define ORACLE_STDSORT_EMUL_H
#include <stdint.h>

#ifndef SSE_FN
#define SSE_FN static inline
#endif

typedef struct {
  const uint32_t* key; /* key[element] */
  long heap_sorts;     /* how often the depth limit was hit (statistics for the tests) */
} SseCtx;

#define SSE_COMP(c, a, b) ((c)->key[(a)] > (c)->key[(b)])

SSE_FN void sse_swap(uint16_t* x, uint16_t* y) {
  const uint16_t t = *x;
  *x = *y;
  *y = t;
}

/* ---- <bits/stl_heap.h> ---- */
SSE_FN void sse_push_heap(SseCtx* c, uint16_t* first, int hole, int top, uint16_t value) {
  int parent = (hole - 1) / 2;
  while (hole > top && SSE_COMP(c, first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

SSE_FN void sse_adjust_heap(SseCtx* c, uint16_t* first, int hole, int len, uint16_t value) {
  const int top = hole;
  int second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (SSE_COMP(c, first[second], first[second - 1])) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  sse_push_heap(c, first, hole, top, value);
}

/* std::__partial_sort(first, last, last): __heap_select degenerates to __make_heap, then __sort_heap */
SSE_FN void sse_heap_sort(SseCtx* c, uint16_t* first, int len) {
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const uint16_t value = first[parent];
      sse_adjust_heap(c, first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) { /* __sort_heap: __pop_heap(first, last, last) */
    --last;
    const uint16_t value = first[last];
    first[last] = first[0];
    sse_adjust_heap(c, first, 0, last, value);
  }
}

/* ---- <bits/stl_algo.h> ---- */
SSE_FN void sse_move_median_to_first(SseCtx* c, uint16_t* result, uint16_t* a, uint16_t* b, uint16_t* cc) {
  if (SSE_COMP(c, *a, *b)) {
    if (SSE_COMP(c, *b, *cc))
      sse_swap(result, b);
    else if (SSE_COMP(c, *a, *cc))
      sse_swap(result, cc);
    else
      sse_swap(result, a);
  } else if (SSE_COMP(c, *a, *cc)) {
    sse_swap(result, a);
  } else if (SSE_COMP(c, *b, *cc)) {
    sse_swap(result, cc);
  } else {
    sse_swap(result, b);
  }
}

SSE_FN int sse_unguarded_partition(SseCtx* c, uint16_t* v, int first, int last, int pivot) {
  for (;;) {
    while (SSE_COMP(c, v[first], v[pivot])) ++first;
    --last;
    while (SSE_COMP(c, v[pivot], v[last])) --last;
    if (!(first < last)) return first;
    sse_swap(&v[first], &v[last]);
    ++first;
  }
}

SSE_FN void sse_unguarded_linear_insert(SseCtx* c, uint16_t* v, int last) {
  const uint16_t val = v[last];
  int next = last - 1;
  while (SSE_COMP(c, val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}

SSE_FN void sse_insertion_sort(SseCtx* c, uint16_t* v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (SSE_COMP(c, v[i], v[first])) {
      const uint16_t val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1]; /* move_backward(first, i, i + 1) */
      v[first] = val;
    } else {
      sse_unguarded_linear_insert(c, v, i);
    }
  }
}

/* v[0..n): the elements (indices into key) in input order; sorted in place as std::sort would */
SSE_FN void sse_sort(SseCtx* c, uint16_t* v, int n) {
  if (n <= 0) return;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) lg++; /* std::__lg */
  /* __introsort_loop: explicit stack of the right-hand parts (they are disjoint: any order) */
  int stack_first[64], stack_last[64], stack_depth[64], sp = 0;
  stack_first[0] = 0, stack_last[0] = n, stack_depth[0] = 2 * lg, sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
        c->heap_sorts++;
        sse_heap_sort(c, v + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      sse_move_median_to_first(c, &v[first], &v[first + 1], &v[mid], &v[last - 1]);
      const int cut = sse_unguarded_partition(c, v, first + 1, last, first);
      stack_first[sp] = cut, stack_last[sp] = last, stack_depth[sp] = depth, sp++;
      last = cut;
    }
  }
  /* __final_insertion_sort */
  if (n > 16) {
    sse_insertion_sort(c, v, 0, 16);
    for (int i = 16; i != n; ++i) sse_unguarded_linear_insert(c, v, i);
  } else {
    sse_insertion_sort(c, v, 0, n);
  }
}
#endif
