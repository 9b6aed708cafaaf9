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

extern "C" int64_t orc_cluster_reads(const orc_cluster_params* P, const orc_reads* R, int32_t* out_cluster,
				     int32_t* out_strand, int32_t* out_founder, int64_t* n_aligned)
{
    MinDB db(1000000);	// cluster.cpp:92-94 MIN_DB_RESERVE
    int32_t nmax[225];
    orc_nmax_table(P->p_shared, P->min_prob_no_hits, nmax);
    std::vector<int32_t> founder;  // cluster id -> read index of its representative
    int64_t aligned_calls = 0;

    for (int64_t i = 0; i < R->n_reads; i++) {
	out_cluster[i] = -1;
	out_strand[i] = 0;
	// skips, cluster.cpp:116-160
	if (R->status[i] != 1) continue;
	if (R->score[i] < 0) continue;
	int64_t rawLen = R->raw_offs[i + 1] - R->raw_offs[i];
	if (uint64_t(rawLen) < uint64_t(unsigned(2 * P->k))) continue;
	if (uint64_t(R->hpc_len[i]) < uint64_t(unsigned(2 * P->k))) continue;
	if (orc_quality_filtered(R->raw_err[i], P->min_qual)) continue;

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

	int best = -1, bestStrand = 0;
	// getBestCluster, cluster.cpp:530-568
	if (!order.empty()) {
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
		aligned_calls++;
		unsigned top = order[0].Size;
		std::string read(R->raw_seq + R->raw_offs[i], size_t(rawLen));
		for (auto& c : order) {
		    if (c.Size < top) break;
		    int f = founder[c.Cls];
		    std::string rep(R->raw_seq + R->raw_offs[f], size_t(R->raw_offs[f + 1] - R->raw_offs[f]));
		    if (c.Strand == -1) rep = revcomp(rep);
		    double e1 = R->raw_err[i], e2 = R->raw_err[f];
		    int go = orc_set_gap_open(e1 + e2);
		    int64_t al = orc_align(read.data(), int64_t(read.size()), rep.data(), int64_t(rep.size()), go, e1 + e2,
					   P->k, nullptr, nullptr, nullptr, nullptr, nullptr);
		    double ratio = double(al) / unsigned(read.size());
		    if (ratio >= P->aligned_th) {
			best = int(c.Cls);
			bestStrand = c.Strand;
			break;
		    }
		}
	    }
	}
	if (best == -1) {
	    // new cluster, cluster.cpp:177-222 ; AddMinimizers minimizer.cpp:31-42
	    unsigned newId = unsigned(founder.size());
	    for (int64_t t = 0; t < nm; t++) {
		auto v = db.find(mins[t].Min);
		if (v == db.end()) db[mins[t].Min] = std::vector<unsigned>{newId};
		else if (v->second.empty() || newId > v->second.back()) v->second.push_back(newId);
	    }
	    founder.push_back(int32_t(i));
	    out_cluster[i] = int32_t(newId);
	    out_strand[i] = 1;
	}
	else {
	    // join, cluster.cpp:223-261 (strand flip :235-246)
	    out_cluster[i] = best;
	    out_strand[i] = bestStrand == -1 ? -1 : 1;
	}
    }
    for (size_t c = 0; c < founder.size(); c++) out_founder[c] = founder[c];
    if (n_aligned) *n_aligned = aligned_calls;
    return int64_t(founder.size());
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
	std::vector<uint32_t> o(cur[best]), n(nw);  // the two std::set<unsigned> of minimizer.cpp:127-137
	std::sort(o.begin(), o.end());
	o.erase(std::unique(o.begin(), o.end()), o.end());
	std::sort(n.begin(), n.end());
	n.erase(std::unique(n.begin(), n.end()), n.end());
	std::vector<uint32_t> to_del, to_ins;
	std::set_difference(o.begin(), o.end(), n.begin(), n.end(), std::back_inserter(to_del));  // :139-140
	std::set_difference(n.begin(), n.end(), o.begin(), o.end(), std::back_inserter(to_ins));  // :141-142
	for (uint32_t m : to_del) {  // :144-153: the list becomes the sorted set of its ids without `best`
	    std::vector<unsigned>& pl = db[m];
	    std::sort(pl.begin(), pl.end());
	    pl.erase(std::unique(pl.begin(), pl.end()), pl.end());
	    pl.erase(std::remove(pl.begin(), pl.end(), best), pl.end());
	}
	for (uint32_t m : to_ins) {  // :155-159
	    std::vector<unsigned>& pl = db[m];
	    pl.push_back(best);
	    std::sort(pl.begin(), pl.end());
	}
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
