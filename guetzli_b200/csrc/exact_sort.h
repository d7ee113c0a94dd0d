// Prefix-exact replay of libstdc++'s std::sort (GCC 13, bits/stl_algo.h) for the
// selection-walk order (g/processor.cc:675): the reference sorts (block, key) pairs
// with a key-only comparator, so the arrangement of equal keys is whatever introsort
// leaves behind, and the walk's result can depend on it.  The walk consumes only a
// prefix of the sorted array; introsort's recursion never lets a sub-range
// influence anything outside itself, so sub-ranges that start beyond the wanted
// prefix can be skipped: O(n) partition work instead of O(n log n).
//
// partial_std_sort(a, n, want) returns k_end >= min(want, n) such that a[0..k_end)
// is element-for-element what std::sort(a, a+n, less) would have produced.
#pragma once
#include <stddef.h>

#include <utility>

namespace gb200 {
namespace exact_sort {

typedef std::pair<int, float> Item;
inline bool less(const Item& a, const Item& b) { return a.second < b.second; }

// std::__adjust_heap + std::__push_heap
inline void adjust_heap(Item* first, ptrdiff_t hole, ptrdiff_t len, Item value) {
  const ptrdiff_t top = hole;
  ptrdiff_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  ptrdiff_t parent = (hole - 1) / 2;
  while (hole > top && less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
inline void heap_sort(Item* first, ptrdiff_t n) {
  if (n >= 2) {
    ptrdiff_t parent = (n - 2) / 2;
    for (;;) {
      Item v = first[parent];
      adjust_heap(first, parent, n, v);
      if (parent == 0) break;
      parent--;
    }
  }
  for (ptrdiff_t last = n; last > 1;) {
    --last;
    Item v = first[last];
    first[last] = first[0];
    adjust_heap(first, 0, last, v);
  }
}

inline void unguarded_linear_insert(Item* a, ptrdiff_t last) {
  Item val = a[last];
  ptrdiff_t next = last - 1;
  while (less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

inline void insertion_sort(Item* a, ptrdiff_t first, ptrdiff_t last) {
  if (first == last) return;
  for (ptrdiff_t i = first + 1; i != last; ++i) {
    if (less(a[i], a[first])) {
      Item val = a[i];
      for (ptrdiff_t j = i; j > first; --j) a[j] = a[j - 1];
      a[first] = val;
    } else {
      unguarded_linear_insert(a, i);
    }
  }
}

// std::__introsort_loop on a[first, last) with `depth` levels left, restricted to the
// sub-ranges that start before `want`; *k_end grows to the end of the last leaf range
// (<= 16 elements, not yet insertion-sorted) visited.  partition(first, last) may be
// replaced by an equivalent implementation (order_exact.h runs large ranges on the device).
inline void introsort_prefix(Item* a, ptrdiff_t first0, ptrdiff_t last0, ptrdiff_t depth0, ptrdiff_t want,
                             ptrdiff_t* k_end) {
  ptrdiff_t st_first[160], st_last[160], st_depth[160];
  int sp = 0;
  st_first[sp] = first0;
  st_last[sp] = last0;
  st_depth[sp] = depth0;
  ++sp;
  while (sp > 0) {
    --sp;
    ptrdiff_t first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
    if (first >= want) continue;
    while (last - first > 16) {
      if (depth == 0) {
        heap_sort(a + first, last - first);
        break;
      }
      --depth;
      // __move_median_to_first(first, first+1, mid, last-1)
      const ptrdiff_t mid = first + (last - first) / 2;
      const ptrdiff_t ia = first + 1, ib = mid, ic = last - 1;
      if (less(a[ia], a[ib])) {
        if (less(a[ib], a[ic])) std::swap(a[first], a[ib]);
        else if (less(a[ia], a[ic])) std::swap(a[first], a[ic]);
        else std::swap(a[first], a[ia]);
      } else if (less(a[ia], a[ic])) {
        std::swap(a[first], a[ia]);
      } else if (less(a[ib], a[ic])) {
        std::swap(a[first], a[ic]);
      } else {
        std::swap(a[first], a[ib]);
      }
      // __unguarded_partition(first+1, last, pivot = *first)
      ptrdiff_t lo = first + 1, hi = last;
      for (;;) {
        while (less(a[lo], a[first])) ++lo;
        --hi;
        while (less(a[first], a[hi])) --hi;
        if (!(lo < hi)) break;
        std::swap(a[lo], a[hi]);
        ++lo;
      }
      const ptrdiff_t cut = lo;
      if (cut < want) {  // the right part is needed too
        st_first[sp] = cut;
        st_last[sp] = last;
        st_depth[sp] = depth;
        ++sp;
      }
      last = cut;
    }
    if (last > *k_end) *k_end = last;  // leaf range [first, last), first < want
  }
}

inline ptrdiff_t introsort_depth_limit(ptrdiff_t n) {
  ptrdiff_t depth_limit = 0;
  for (ptrdiff_t m = n; m > 1; m >>= 1) ++depth_limit;
  return 2 * depth_limit;
}

// std::__final_insertion_sort of an n-element array restricted to the prefix [0, k_end).
inline void final_insertion_prefix(Item* a, ptrdiff_t n, ptrdiff_t* k_end) {
  if (n > 16) {
    const ptrdiff_t head = *k_end < 16 ? *k_end : 16;
    insertion_sort(a, 0, head);
    for (ptrdiff_t i = 16; i < *k_end; ++i) unguarded_linear_insert(a, i);
  } else {
    insertion_sort(a, 0, n);
    *k_end = n;
  }
}

inline size_t partial_std_sort(Item* a, size_t n_, size_t want_) {
  const ptrdiff_t n = static_cast<ptrdiff_t>(n_);
  const ptrdiff_t want = static_cast<ptrdiff_t>(want_ < n_ ? want_ : n_);
  if (n < 2) return n_;
  // ranges are visited right-to-left within a parent, so k_end is the maximum end over all
  // leaf ranges that start before `want`; they tile [0, k_end).
  ptrdiff_t k_end = 0;
  introsort_prefix(a, 0, n, introsort_depth_limit(n), want, &k_end);
  final_insertion_prefix(a, n, &k_end);
  return static_cast<size_t>(k_end);
}

}  // namespace exact_sort
}  // namespace gb200
