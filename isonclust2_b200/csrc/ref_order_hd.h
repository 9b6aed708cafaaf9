// libstdc++'s std::sort on (Size, payload) keys, restated so that it runs on the DEVICE as well as on the host.
//
// SortMinimizerHits (reference cluster.cpp:622-636) std::sort()s the candidates by Size, descending; std::sort is not
// stable, so the order of equally scored candidates is defined by libstdc++'s introsort (ref_order.h explains how the
// input sequence — the unordered_map's iteration order — is obtained).  Only ONE group of equally scored candidates is
// ever consulted, so, as in ref_order.h, ranges that hold no member of that group are not partitioned (the final
// insertion sort never moves an element across a range boundary and keeps equal elements in place).
//
// Everything std::sort does is restated here without the library: __introsort_loop's median-of-three + unguarded Hoare
// partition, and its depth-exhausted branch std::partial_sort(first, last, last) = __heap_select + __sort_heap
// (stl_algo.h / stl_heap.h of libstdc++).  tests/cpp/test_ref_order.cpp checks it against the real std::sort, including
// inputs that exhaust the depth budget; icb_ref_order_selfcheck() (cluster.cu) repeats a small check at start-up.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ICB_HD __host__ __device__ __forceinline__
#else
#define ICB_HD inline
#endif

namespace icb_sort_hd {

typedef uint64_t Key;  // Size << 32 | payload; only Size takes part in comparisons
ICB_HD bool gt(Key a, Key b) { return (a >> 32) > (b >> 32); }
ICB_HD void swp(Key& a, Key& b)
{
    const Key t = a;
    a = b;
    b = t;
}

// ---- stl_heap.h with comp = gt (a "max-heap" under comp: the root is the element no other is gt-before, i.e. the smallest Size)
ICB_HD void push_heap(Key* first, long hole, long top, Key value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && gt(first[parent], value)) {
	first[hole] = first[parent];
	hole = parent;
	parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
ICB_HD void adjust_heap(Key* first, long hole, long len, Key value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
	child = 2 * (child + 1);
	if (gt(first[child], first[child - 1])) child--;
	first[hole] = first[child];
	hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
	child = 2 * (child + 1);
	first[hole] = first[child - 1];
	hole = child - 1;
    }
    push_heap(first, hole, top, value);
}
// std::partial_sort(first, last, last, gt): __heap_select with an empty tail = __make_heap, then __sort_heap
ICB_HD void heap_sort(Key* first, Key* last)
{
    const long len = long(last - first);
    if (len >= 2) {
	long parent = (len - 2) / 2;
	for (;;) {
	    const Key v = first[parent];
	    adjust_heap(first, parent, len, v);
	    if (parent == 0) break;
	    parent--;
	}
    }
    while (last - first > 1) {
	--last;
	const Key v = *last;
	*last = *first;
	adjust_heap(first, 0, long(last - first), v);
    }
}

// ---- stl_algo.h
ICB_HD void median_to_first(Key* result, Key* a, Key* b, Key* c)
{
    if (gt(*a, *b)) {
	if (gt(*b, *c)) swp(*result, *b);
	else if (gt(*a, *c)) swp(*result, *c);
	else swp(*result, *a);
    }
    else if (gt(*a, *c)) swp(*result, *a);
    else if (gt(*b, *c)) swp(*result, *c);
    else swp(*result, *b);
}
ICB_HD Key* partition_pivot(Key* first, Key* last)
{
    Key* mid = first + (last - first) / 2;
    median_to_first(first, first + 1, mid, last - 1);
    Key* lo = first + 1;
    Key* hi = last;
    const Key pivot = *first;
    for (;;) {
	while (gt(*lo, pivot)) ++lo;
	--hi;
	while (gt(pivot, *hi)) --hi;
	if (!(lo < hi)) return lo;
	swp(*lo, *hi);
	++lo;
    }
}
ICB_HD bool holds(const Key* first, const Key* last, uint32_t target)
{
    for (; first != last; ++first)
	if (uint32_t(*first >> 32) == target) return true;
    return false;
}

// __introsort_loop over [k0, k0 + n), restricted to the ranges that hold an element with Size == target.  The recursion
// of the library (right part recursed into, left part iterated) becomes an explicit stack: at most one frame per level
// of the depth budget 2 * floor(log2 n) <= 62.
ICB_HD void introsort_group(Key* k0, long n, uint32_t target)
{
    struct Frame {
	Key *first, *last;
	long depth;
    } stack[64];
    int sp = 0;
    long lg = 0;
    for (unsigned long v = (unsigned long)n; v > 1; v >>= 1) lg++;
    Key* first = k0;
    Key* last = k0 + n;
    long depth = 2 * lg;
    for (;;) {
	while (last - first > 16) {
	    if (!holds(first, last, target)) break;
	    if (depth == 0) {
		heap_sort(first, last);
		break;
	    }
	    --depth;
	    Key* cut = partition_pivot(first, last);
	    // the library recurses into [cut, last) first, then continues with [first, cut): the two ranges are disjoint, so
	    // the order of processing does not matter; the left part is stacked, the right part processed next
	    stack[sp].first = first;
	    stack[sp].last = cut;
	    stack[sp].depth = depth;
	    sp++;
	    first = cut;
	}
	if (sp == 0) break;
	sp--;
	first = stack[sp].first;
	last = stack[sp].last;
	depth = stack[sp].depth;
    }
}

}  // namespace icb_sort_hd
