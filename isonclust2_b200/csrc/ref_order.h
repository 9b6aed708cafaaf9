// The reference's candidate ORDER among equally scored candidates, reproduced on the host (plain C++, no CUDA).
//
// SortMinimizerHits (reference cluster.cpp:622-636) walks the std::unordered_map GetMinimizerHits filled
// (minimizer.cpp:44-76: keys (cluster, strand), hash StrandedClsHash minimizer.h:53-58, 20*(|mins|+|revMins|) buckets,
// minimizer.cpp:48) and std::sort()s the result by Size, descending.  std::sort is not stable, so which of two equally
// scored candidates comes first depends on the container's iteration order and on libstdc++'s introsort.  The final sort
// is done with the real std::sort on the same sequence; the iteration order is computed directly:
//
//   libstdc++'s _Hashtable keeps ONE singly linked list.  A node whose bucket is empty is pushed at the head of the list;
//   otherwise it is linked in at the beginning of its bucket (_M_insert_bucket_begin).  Hence: buckets appear in reverse
//   order of their first insertion, and inside a bucket nodes appear newest first.  No rehash happens while the number of
//   keys stays within the bucket count the constructor chose (_Prime_rehash_policy::_M_next_bkt(hint), max load 1.0);
//   above that the real container is used.
//
// Insertion order of GetMinimizerHits: forward minimizers by Index, each posting list by ascending cluster id, then the
// reverse minimizers likewise -> keys ordered by (strand: +1 first, Index of the first hit, cluster id).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <unordered_map>
#include <utility>
#include <vector>

struct IcbPairHash {  // minimizer.h:53-58 StrandedClsHash
    std::size_t operator()(const std::pair<int, int>& u) const { return std::size_t(int(u.first * u.second)); }
};

struct IcbSortedCand {	// SortedHit (cluster.h): Size = hits, Cls, Strand
    unsigned Size;
    unsigned Cls;
    int Strand;
};

struct IcbRefOrderScratch {
    std::vector<uint32_t> count;		   // counting sort by (strand, first hit)
    std::vector<int32_t> ins;			   // candidate index by insertion rank
    std::vector<uint32_t> stamp, first_seen, last;  // per bucket
    std::vector<int32_t> prev;			   // per insertion rank: older node of the same bucket
    std::vector<uint32_t> bucket_of;		   // per insertion rank
    std::vector<uint64_t> keys;			   // icb_reference_order_group: Size << 32 | candidate, in iteration order
    struct Slot {
	uint32_t bucket, first_seen, last;
    };
    std::vector<Slot> tab;			   // icb_reference_order_group: occupied buckets, open addressing
    uint32_t cur = 0;
};

// cands[0..n): distinct (id, strand) keys with id = the cluster id the reference sees, ASCENDING by id; `first` = Index of
// the key's first hit among the read's minimizers of that strand.  out: the reference's `order` vector after the sort.
template <class Cand>
void icb_reference_order(const Cand* cands, int n, size_t bucket_hint, IcbRefOrderScratch& S, std::vector<IcbSortedCand>& out)
{
    out.clear();
    if (n <= 0) return;
    // ---- insertion order: stable counting sort by (strand, first hit) of the id-sorted input
    uint32_t maxFirst = 0;
    for (int i = 0; i < n; i++) maxFirst = std::max(maxFirst, uint32_t(cands[i].tmapped));
    const uint32_t M1 = maxFirst + 1;
    S.count.assign(size_t(2) * M1 + 1, 0u);
    for (int i = 0; i < n; i++) S.count[(cands[i].strand < 0 ? M1 : 0u) + uint32_t(cands[i].tmapped) + 1]++;
    for (size_t b = 1; b < S.count.size(); b++) S.count[b] += S.count[b - 1];
    S.ins.resize(n);
    for (int i = 0; i < n; i++) S.ins[S.count[(cands[i].strand < 0 ? M1 : 0u) + uint32_t(cands[i].tmapped)]++] = i;
    // ---- iteration order of the unordered_map
    const size_t bkt = std::__detail::_Prime_rehash_policy()._M_next_bkt(bucket_hint);
    out.reserve(n);
    if (size_t(n) > bkt) {
	// the container would rehash on the way: use it
	std::unordered_map<std::pair<int, int>, unsigned, IcbPairHash> m(bucket_hint);
	for (int t = 0; t < n; t++) {
	    const Cand& c = cands[S.ins[t]];
	    m[std::make_pair(int(c.id), int(c.strand))] = unsigned(c.count);
	}
	for (auto& kv : m) out.push_back(IcbSortedCand{kv.second, unsigned(kv.first.first), kv.first.second});
    }
    else {
	if (S.stamp.size() < bkt) {
	    S.stamp.assign(bkt, 0u);
	    S.first_seen.resize(bkt);
	    S.last.resize(bkt);
	    S.cur = 0;
	}
	if (++S.cur == 0) {
	    std::fill(S.stamp.begin(), S.stamp.end(), 0u);
	    S.cur = 1;
	}
	S.prev.resize(n);
	S.bucket_of.resize(n);
	const IcbPairHash H;
	for (int t = 0; t < n; t++) {
	    const Cand& c = cands[S.ins[t]];
	    const uint32_t b = uint32_t(H(std::make_pair(int(c.id), int(c.strand))) % bkt);
	    S.bucket_of[t] = b;
	    if (S.stamp[b] != S.cur) {
		S.stamp[b] = S.cur;
		S.first_seen[b] = uint32_t(t);
		S.prev[t] = -1;
	    }
	    else
		S.prev[t] = int32_t(S.last[b]);
	    S.last[b] = uint32_t(t);
	}
	for (int t = n - 1; t >= 0; t--) {
	    const uint32_t b = S.bucket_of[t];
	    if (S.first_seen[b] != uint32_t(t)) continue;  // listed with its bucket, at the place of the bucket's oldest node
	    for (int32_t u = int32_t(S.last[b]); u >= 0; u = S.prev[u]) {
		const Cand& c = cands[S.ins[u]];
		out.push_back(IcbSortedCand{unsigned(c.count), unsigned(c.id), int(c.strand)});
	    }
	}
    }
    std::sort(out.begin(), out.end(), [](const IcbSortedCand& a, const IcbSortedCand& b) { return a.Size > b.Size; });
}

// ------------------------------------------------------------------------------------------------------------------
// Only one group of equally scored candidates is ever consulted (the ones tied at the top count for the alignment test,
// or one count group for the mapping test), so the full sort is not needed: what matters is the order in which the
// elements of that group leave std::sort.  libstdc++'s std::sort is introsort: __introsort_loop (median-of-three pivot
// moved to the front, unguarded Hoare partition, recursion on the right part, iteration on the left, heap sort when the
// depth budget 2*floor(log2 n) is used up) until ranges are at most 16 long, then one insertion sort over everything.
// Ranges are disjoint and ordered against each other, the insertion sort never moves an element past an element that is
// not smaller, and it keeps equal elements in place — so the group's final order is the order in which its members
// stand after the introsort loop, and a range that holds no member of the group need not be partitioned at all.
// The loop below performs exactly the partition steps std::sort performs on the ranges that hold group members
// (quick-select-like O(n) instead of O(n log n)); the depth-exhausted case runs the real std::partial_sort.
// tests/cpp/test_ref_order.cpp checks it against the real std::sort.
namespace icb_introsort {
typedef uint64_t Key;  // Size << 32 | position in iteration order
inline bool gt(Key a, Key b) { return (a >> 32) > (b >> 32); }
inline bool holds(const Key* first, const Key* last, uint32_t target)
{
    for (; first != last; ++first)
	if (uint32_t(*first >> 32) == target) return true;
    return false;
}
inline void median_to_first(Key* result, Key* a, Key* b, Key* c)
{
    if (gt(*a, *b)) {
	if (gt(*b, *c)) std::swap(*result, *b);
	else if (gt(*a, *c)) std::swap(*result, *c);
	else std::swap(*result, *a);
    }
    else if (gt(*a, *c)) std::swap(*result, *a);
    else if (gt(*b, *c)) std::swap(*result, *c);
    else std::swap(*result, *b);
}
inline Key* partition_pivot(Key* first, Key* last)
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
	std::swap(*lo, *hi);
	++lo;
    }
}
inline void loop(Key* first, Key* last, long depth, uint32_t target)
{
    while (last - first > 16) {
	if (!holds(first, last, target)) return;
	if (depth == 0) {
	    std::partial_sort(first, last, last, [](Key a, Key b) { return gt(a, b); });
	    return;
	}
	--depth;
	Key* cut = partition_pivot(first, last);
	loop(cut, last, depth, target);
	last = cut;
    }
}
}  // namespace icb_introsort

// Insertion order of GetMinimizerHits for an id-sorted candidate list: stable counting sort by (strand: +1 first, first
// hit).  perm[rank] = index into cands.
template <class Cand>
void icb_insertion_order(const Cand* cands, int n, IcbRefOrderScratch& S, std::vector<int32_t>& perm)
{
    perm.resize(size_t(n > 0 ? n : 0));
    if (n <= 0) return;
    uint32_t maxFirst = 0;
    for (int i = 0; i < n; i++) maxFirst = std::max(maxFirst, uint32_t(cands[i].tmapped));
    const uint32_t M1 = maxFirst + 1;
    S.count.assign(size_t(2) * M1 + 1, 0u);
    for (int i = 0; i < n; i++) S.count[(cands[i].strand < 0 ? M1 : 0u) + uint32_t(cands[i].tmapped) + 1]++;
    for (size_t b = 1; b < S.count.size(); b++) S.count[b] += S.count[b - 1];
    for (int i = 0; i < n; i++) perm[S.count[(cands[i].strand < 0 ? M1 : 0u) + uint32_t(cands[i].tmapped)]++] = i;
}

// The members of the group Size == target, in the order the reference's sorted `order` vector lists them.
// PRESORTED: cands is in insertion order already (icb_insertion_order applied by the caller; dropping elements from such a
// list keeps it in insertion order) instead of ascending by id.
template <bool PRESORTED = false, class Cand>
void icb_reference_order_group(const Cand* cands, int n, size_t bucket_hint, uint32_t target, IcbRefOrderScratch& S,
			       std::vector<IcbSortedCand>& out)
{
    out.clear();
    if (n <= 0) return;
    if (PRESORTED) {
	S.ins.resize(n);
	for (int i = 0; i < n; i++) S.ins[i] = i;
    }
    else
	icb_insertion_order(cands, n, S, S.ins);
    // ---- iteration order of the unordered_map -> S.keys[position] = Size << 32 | candidate index
    const size_t bkt = std::__detail::_Prime_rehash_policy()._M_next_bkt(bucket_hint);
    S.keys.clear();
    S.keys.reserve(n);
    if (size_t(n) > bkt || bkt > 0xffffffffull) {
	// the container would rehash on the way: use it
	std::unordered_map<std::pair<int, int>, unsigned, IcbPairHash> m(bucket_hint);
	for (int t = 0; t < n; t++) {
	    const Cand& c = cands[S.ins[t]];
	    m[std::make_pair(int(c.id), int(c.strand))] = unsigned(S.ins[t]);
	}
	for (auto& kv : m) S.keys.push_back(uint64_t(uint32_t(cands[kv.second].count)) << 32 | kv.second);
    }
    else {
	// One singly linked list, buckets in reverse order of their first insertion, newest node first inside a bucket.
	// Only the <= n occupied buckets matter: they are kept in a small open-addressing table (L1-resident) instead of
	// arrays over all `bkt` buckets.
	uint32_t tsize = 64;
	while (tsize < 2u * uint32_t(n)) tsize <<= 1;
	const uint32_t tmask = tsize - 1;
	S.tab.assign(size_t(tsize), IcbRefOrderScratch::Slot{0xffffffffu, 0u, 0u});
	S.prev.resize(n);
	S.bucket_of.resize(n);	// slot of the node's bucket
	// bucket = size_t(int(id * strand)) % bkt (StrandedClsHash, sign-extended): id % bkt on the forward strand,
	// (2^64 - id) % bkt on the reverse strand; 32-bit remainders by multiplication (Lemire's fastmod)
	const uint32_t b32 = uint32_t(bkt);
	const uint64_t fm = ~uint64_t(0) / b32 + 1;
	const uint32_t two64 = uint32_t((((~uint64_t(0)) % bkt) + 1) % bkt);
	for (int t = 0; t < n; t++) {
	    const Cand& c = cands[S.ins[t]];
	    const uint32_t r = uint32_t((__uint128_t(fm * uint32_t(c.id)) * b32) >> 64);  // id % bkt
	    uint32_t b = r;
	    if (c.strand < 0 && c.id != 0) b = two64 >= r ? two64 - r : two64 + b32 - r;
	    uint32_t slot = (b * 2654435761u) & tmask;
	    while (S.tab[slot].bucket != b && S.tab[slot].bucket != 0xffffffffu) slot = (slot + 1) & tmask;
	    IcbRefOrderScratch::Slot& e = S.tab[slot];
	    S.bucket_of[t] = slot;
	    if (e.bucket == 0xffffffffu) {
		e.bucket = b;
		e.first_seen = uint32_t(t);
		S.prev[t] = -1;
	    }
	    else
		S.prev[t] = int32_t(e.last);
	    e.last = uint32_t(t);
	}
	for (int t = n - 1; t >= 0; t--) {
	    const IcbRefOrderScratch::Slot& e = S.tab[S.bucket_of[t]];
	    if (e.first_seen != uint32_t(t)) continue;	// listed with its bucket, at the place of the bucket's oldest node
	    for (int32_t u = int32_t(e.last); u >= 0; u = S.prev[u]) {
		const int ci = S.ins[u];
		S.keys.push_back(uint64_t(uint32_t(cands[ci].count)) << 32 | uint32_t(ci));
	    }
	}
    }
    // ---- the group's order after std::sort(order, Size descending)
    icb_introsort::Key* k0 = S.keys.data();
    long lg = 0;
    for (size_t v = size_t(n); v > 1; v >>= 1) lg++;
    icb_introsort::loop(k0, k0 + n, 2 * lg, target);
    for (int t = 0; t < n; t++)
	if (uint32_t(k0[t] >> 32) == target) {
	    const Cand& c = cands[uint32_t(k0[t])];
	    out.push_back(IcbSortedCand{unsigned(c.count), unsigned(c.id), int(c.strand)});
	}
}
