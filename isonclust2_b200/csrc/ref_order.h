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
