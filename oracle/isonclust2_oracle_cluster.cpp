// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement of isONclust2's greedy, order-dependent assignment loop and
// the index/candidate functions it calls (a11, a13, a14, a17, a22, a23 and
// cluster.cpp:67-322 with consensus off).  C++ rather than C because the order
// in which equally-scored candidates are tried is *defined* by libstdc++:
// the iteration order of an unordered_map with hash size_t(int(cls*strand))
// and 20*(|mins|+|revMins|) buckets (minimizer.cpp:48, minimizer.h:53-58)
// followed by an unstable std::sort (cluster.cpp:622-636).  The same two
// library facilities are used here on flat data.
//
// Pinned against oracle/_ref (the reference compiled unmodified) by
// tests/test_oracle_cluster.py: identical cluster membership, strands and
// alignment-invocation counts on seeded synthetic batches in all three modes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <iterator>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "isonclust2_oracle.h"

namespace {

struct PairHash {
    // minimizer.h:53-58 StrandedClsHash
    std::size_t operator()(const std::pair<int, int>& u) const { return std::size_t(int(u.first * u.second)); }
};
struct Hit {
    uint32_t Pos, Index;
};
typedef std::unordered_map<std::pair<int, int>, std::vector<Hit>, PairHash> HitMap;
struct Cand {
    unsigned Size;
    unsigned Cls;
    int Strand;
};
struct IdHash {
    std::size_t operator()(const unsigned& u) const { return std::size_t(u); }
};
typedef std::unordered_map<unsigned, std::vector<unsigned>, IdHash> MinDB;

std::string revcomp(const std::string& s)
{
    std::string r(s.size(), 'N');
    orc_revcomp(s.data(), int64_t(s.size()), &r[0]);
    return r;
}

}  // namespace

namespace {

// state of the greedy loop: the MinDB and, per cluster, the read that represents it
struct LoopState {
    MinDB db{1000000};	// cluster.cpp:92-94 MIN_DB_RESERVE
    std::vector<int32_t> founder;
    int64_t aligned_calls = 0;
};

// the skips of ClusterSortedReads, cluster.cpp:116-160
bool skipped(const orc_cluster_params* P, const orc_reads* R, int64_t i)
{
    if (R->status[i] != 1) return true;
    if (R->score[i] < 0) return true;
    int64_t rawLen = R->raw_offs[i + 1] - R->raw_offs[i];
    if (uint64_t(rawLen) < uint64_t(unsigned(2 * P->k))) return true;
    if (uint64_t(R->hpc_len[i]) < uint64_t(unsigned(2 * P->k))) return true;
    if (orc_quality_filtered(R->raw_err[i], P->min_qual)) return true;
    return false;
}

// AddMinimizers, minimizer.cpp:31-42
void add_minimizers(LoopState& S, const orc_reads* R, int64_t i, unsigned newId)
{
    const orc_minimizer* mins = R->mins + R->min_offs[i];
    const int64_t nm = R->min_offs[i + 1] - R->min_offs[i];
    for (int64_t t = 0; t < nm; t++) {
	auto v = S.db.find(mins[t].Min);
	if (v == S.db.end()) S.db[mins[t].Min] = std::vector<unsigned>{newId};
	else if (v->second.empty() || newId > v->second.back()) v->second.push_back(newId);
    }
}

// UpdateMinDB, minimizer.cpp:124-160, on minimizer VALUES (only .Min matters to it)
void update_min_db(MinDB& db, unsigned best, const std::vector<uint32_t>& oldv, const std::vector<uint32_t>& newv)
{
    std::vector<uint32_t> o(oldv), n(newv);  // the two std::set<unsigned> of minimizer.cpp:127-137
    std::sort(o.begin(), o.end());
    o.erase(std::unique(o.begin(), o.end()), o.end());
    std::sort(n.begin(), n.end());
    n.erase(std::unique(n.begin(), n.end()), n.end());
    std::vector<uint32_t> to_del, to_ins;
    std::set_difference(o.begin(), o.end(), n.begin(), n.end(), std::back_inserter(to_del));  // :139-140
    std::set_difference(n.begin(), n.end(), o.begin(), o.end(), std::back_inserter(to_ins));  // :141-142
    for (uint32_t m : to_del) {	 // :144-153: the list becomes the sorted set of its ids without `best`
	std::vector<unsigned>& pl = db[m];
	std::sort(pl.begin(), pl.end());
	pl.erase(std::unique(pl.begin(), pl.end()), pl.end());
	pl.erase(std::remove(pl.begin(), pl.end(), best), pl.end());
    }
    for (uint32_t m : to_ins) {	 // :155-159
	std::vector<unsigned>& pl = db[m];
	pl.push_back(best);
	std::sort(pl.begin(), pl.end());
    }
}

// getBestCluster (cluster.cpp:530-568) for read i against the clusters of S: best = -1 -> the read founds a cluster
void best_cluster(const orc_cluster_params* P, const orc_reads* R, LoopState& S, int64_t i, int& best, int& bestStrand)
{
    const std::vector<int32_t>& founder = S.founder;
    MinDB& db = S.db;
    const int64_t rawLen = R->raw_offs[i + 1] - R->raw_offs[i];
    const orc_minimizer* mins = R->mins + R->min_offs[i];
    int64_t nm = R->min_offs[i + 1] - R->min_offs[i];
    const orc_minimizer* rmins = R->rev_mins + R->rev_offs[i];
    int64_t nr = R->rev_offs[i + 1] - R->rev_offs[i];

    // GetMinimizerHits, minimizer.cpp:44-76 (+ ConsolidateMinimizerHits cluster.cpp:609-615)
    HitMap hits(20 * (nm + nr));
    for (int64_t t = 0; t < nm; t++) {
	auto it = db.find(mins[t].Min);
	if (it == db.end()) continue;
	for (auto c : it->second) hits[std::make_pair(int(c), 1)].push_back(Hit{mins[t].Pos, mins[t].Index});
    }
    for (int64_t t = 0; t < nr; t++) {
	auto it = db.find(rmins[t].Min);
	if (it == db.end()) continue;
	for (auto c : it->second) hits[std::make_pair(int(c), -1)].push_back(Hit{rmins[t].Pos, rmins[t].Index});
    }
    // SortMinimizerHits, cluster.cpp:622-636
    std::vector<Cand> order;
    order.reserve(hits.size());
    for (auto& h : hits) order.push_back(Cand{unsigned(h.second.size()), unsigned(h.first.first), h.first.second});
    std::sort(order.begin(), order.end(), [](const Cand& a, const Cand& b) { return a.Size > b.Size; });

    best = -1;
    bestStrand = 0;
    if (order.empty()) return;
    bool mapped = false;
    if (P->mode == 0 || P->mode == 1) {
	// getBestClusterMapping, cluster.cpp:355-406
	unsigned top = order[0].Size;
	if (!(top < unsigned(P->min_shared))) {
	    for (auto& c : order) {
		if (int(c.Size) < int(double(top) * P->min_frac)) break;
		auto& hv = hits.at(std::make_pair(int(c.Cls), c.Strand));
		int cl = orc_err_class(R->hpc_err[founder[c.Cls]]) * 15 + orc_err_class(R->hpc_err[i]);
		std::vector<uint32_t> pos(hv.size()), idx(hv.size());
		for (size_t t = 0; t < hv.size(); t++) {
		    pos[t] = hv[t].Pos;
		    idx[t] = hv[t].Index;
		}
		int64_t M = c.Strand == 1 ? nm : nr;
		float mr = float(orc_mapped_ratio(R->hpc_len[i], M, pos.data(), idx.data(), int64_t(hv.size()),
						  P->p_shared[cl], P->min_prob_no_hits));
		if (mr >= P->mapped_th) {
		    best = int(c.Cls);
		    bestStrand = c.Strand;
		    mapped = true;
		    break;
		}
	    }
	}
    }
    if (!mapped && !(order[0].Size < unsigned(P->min_shared)) && P->mode != 1) {
	// getBestClusterAln, cluster.cpp:461-515
	S.aligned_calls++;
	unsigned top = order[0].Size;
	std::string read(R->raw_seq + R->raw_offs[i], size_t(rawLen));
	for (auto& c : order) {
	    if (c.Size < top) break;
	    int f = founder[c.Cls];
	    std::string rep(R->raw_seq + R->raw_offs[f], size_t(R->raw_offs[f + 1] - R->raw_offs[f]));
	    if (c.Strand == -1) rep = revcomp(rep);
	    double e1 = R->raw_err[i], e2 = R->raw_err[f];
	    int go = orc_set_gap_open(e1 + e2);
	    int64_t al = orc_align(read.data(), int64_t(read.size()), rep.data(), int64_t(rep.size()), go, e1 + e2, P->k, nullptr,
				   nullptr, nullptr, nullptr, nullptr);
	    double ratio = double(al) / unsigned(read.size());
	    if (ratio >= P->aligned_th) {
		best = int(c.Cls);
		bestStrand = c.Strand;
		break;
	    }
	}
    }
}

// one step of ClusterSortedReads (cluster.cpp:108-261) for read i; returns false if the read is skipped
bool loop_step(const orc_cluster_params* P, const orc_reads* R, LoopState& S, int64_t i, int32_t& cls, int32_t& strand)
{
    cls = -1;
    strand = 0;
    if (skipped(P, R, i)) return false;
    int best, bestStrand;
    best_cluster(P, R, S, i, best, bestStrand);
    if (best == -1) {
	// new cluster, cluster.cpp:177-222
	unsigned newId = unsigned(S.founder.size());
	add_minimizers(S, R, i, newId);
	S.founder.push_back(int32_t(i));
	cls = int32_t(newId);
	strand = 1;
    }
    else {
	// join, cluster.cpp:223-261 (strand flip :235-246)
	cls = best;
	strand = bestStrand == -1 ? -1 : 1;
    }
    return true;
}

}  // namespace

extern "C" int64_t orc_cluster_reads(const orc_cluster_params* P, const orc_reads* R, int32_t* out_cluster,
				     int32_t* out_strand, int32_t* out_founder, int64_t* n_aligned)
{
    LoopState S;
    for (int64_t i = 0; i < R->n_reads; i++) loop_step(P, R, S, i, out_cluster[i], out_strand[i]);
    for (size_t c = 0; c < S.founder.size(); c++) out_founder[c] = S.founder[c];
    if (n_aligned) *n_aligned = S.aligned_calls;
    return int64_t(S.founder.size());
}

// The loop on clusters that exist already (the left batch of a merge, cluster.cpp:67-106) whose representatives are then
// replaced (the representative swap + UpdateMinDB the reference does after a consensus update, cluster.cpp:279-291,
// minimizer.cpp:124-160): cluster c is seeded from read seeds[c] (AddMinimizers of its forward sketch), update u makes read
// upd_read[u] the representative of cluster upd_cls[u], then the reads queries[0..n_q) are clustered in that order.
// out_founder needs room for n_seeds + n_q entries.
extern "C" int64_t orc_cluster_seeded(const orc_cluster_params* P, const orc_reads* R, const int32_t* seeds, int64_t n_seeds,
				      const int32_t* upd_cls, const int32_t* upd_read, int64_t n_upd, const int32_t* queries,
				      int64_t n_q, int32_t* out_cluster, int32_t* out_strand, int32_t* out_founder, int64_t* n_aligned)
{
    LoopState S;
    auto values = [&](int64_t i) {
	std::vector<uint32_t> v;
	for (int64_t j = R->min_offs[i]; j < R->min_offs[i + 1]; j++) v.push_back(R->mins[j].Min);
	return v;
    };
    for (int64_t c = 0; c < n_seeds; c++) {
	add_minimizers(S, R, seeds[c], unsigned(c));
	S.founder.push_back(seeds[c]);
    }
    for (int64_t u = 0; u < n_upd; u++) {
	const int32_t c = upd_cls[u];
	if (c < 0 || size_t(c) >= S.founder.size()) return -1;
	update_min_db(S.db, unsigned(c), values(S.founder[c]), values(upd_read[u]));
	S.founder[c] = upd_read[u];
    }
    for (int64_t q = 0; q < n_q; q++) loop_step(P, R, S, queries[q], out_cluster[q], out_strand[q]);
    for (size_t c = 0; c < S.founder.size(); c++) out_founder[c] = S.founder[c];
    if (n_aligned) *n_aligned = S.aligned_calls;
    return int64_t(S.founder.size());
}

// a12 — UpdateMinDB (minimizer.cpp:124-160) restated on the flat MinDB above, driven like
// ref_mindb_after_updates of oracle/ref_shim.cpp: AddMinimizers (minimizer.cpp:31-42) for nCls founders, then update u
// replaces the minimizers of cluster upd_cls[u] (its current ones are the "old" set).  Export sorted by key, empty
// posting lists included (the reference does not erase emptied keys, minimizer.cpp:150-152).
extern "C" int64_t orc_mindb_after_updates(const orc_minimizer* cls_mins, const int64_t* cls_offs, int n_cls, const int32_t* upd_cls,
					   const orc_minimizer* upd_mins, const int64_t* upd_offs, int n_upd, uint32_t* keys,
					   int64_t* offs, uint32_t* cls, int64_t key_cap, int64_t cls_cap)
{
    MinDB db(1000000);
    std::vector<std::vector<uint32_t>> cur(n_cls);  // current minimizer VALUES per cluster (only .Min matters to UpdateMinDB)
    for (int c = 0; c < n_cls; c++) {
	for (int64_t j = cls_offs[c]; j < cls_offs[c + 1]; j++) {
	    const uint32_t m = cls_mins[j].Min;
	    cur[c].push_back(m);
	    auto v = db.find(m);
	    if (v == db.end()) db[m] = std::vector<unsigned>{unsigned(c)};
	    else if (v->second.empty() || unsigned(c) > v->second.back()) v->second.push_back(unsigned(c));
	}
    }
    for (int u = 0; u < n_upd; u++) {
	const unsigned best = unsigned(upd_cls[u]);
	std::vector<uint32_t> nw;
	for (int64_t j = upd_offs[u]; j < upd_offs[u + 1]; j++) nw.push_back(upd_mins[j].Min);
	update_min_db(db, best, cur[best], nw);
	cur[best] = nw;
    }
    std::vector<uint32_t> ks;
    int64_t tot = 0;
    for (auto& kv : db) {
	ks.push_back(kv.first);
	tot += int64_t(kv.second.size());
    }
    if (int64_t(ks.size()) > key_cap || tot > cls_cap) return -1 - std::max<int64_t>(int64_t(ks.size()), tot);
    std::sort(ks.begin(), ks.end());
    int64_t p = 0;
    for (size_t i = 0; i < ks.size(); i++) {
	keys[i] = ks[i];
	offs[i] = p;
	for (unsigned c : db[ks[i]]) cls[p++] = c;
    }
    offs[ks.size()] = p;
    return int64_t(ks.size());
}
