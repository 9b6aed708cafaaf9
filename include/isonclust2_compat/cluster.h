// cluster.h of the reference's header face (reference src/cluster.h:15-38, cluster.cpp:324-353, 408-459, 609-636).
// ClusterSortedReads / getBestCluster work on the reference's Batch objects (serialisation layer); their B200 form is
// icb_cluster_reads / icb_cluster_batches of the C ABI.
#ifndef CLUSTER_H_INCLUDED
#define CLUSTER_H_INCLUDED

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "icb_compat.h"
#include "minimizer.h"
#include "p_emp_prob.h"
#include "parasail.h"
#include "seq.h"

#define REP 0

// what SortMinimizerHits' second parameter needs to be here: the cluster lists themselves belong to the batch layer
struct ProcSeqStub {};
typedef std::vector<ProcSeqStub> Cluster;
typedef std::vector<Cluster> Clusters;

// cluster.cpp:324-353.  pError comes from the host table and pow() from the host libm, as in the reference: a run of n missed
// minimizers counts as mapped iff pow(pError, n) >= minProbNoHits, which is n <= n_max (monotone); the walk over the hits runs
// on the device in integers and the division is the reference's.
inline double getMappedRatio(const Seq& hpcSeq, const Seq& clHpcSeq, const Minimizers& mins, const MinimizerHitVector& hits,
			     const MinSharedMap& sharedMinTab, double minProbNoHits)
{
    const double pError = 1.0 - GetPMinShared(clHpcSeq.ErrorRate(), hpcSeq.ErrorRate(), sharedMinTab);
    int32_t nmax = -1;
    for (int32_t t = 0; t <= int32_t(mins.size()); t++) {
	if (pow(pError, double(t)) >= minProbNoHits) nmax = t;
	else break;
    }
    std::vector<uint32_t> pos(hits.size()), idx(hits.size());
    for (size_t i = 0; i < hits.size(); i++) {
	pos[i] = hits[i].Pos;
	idx[i] = hits[i].Index;
    }
    int64_t total = 0;
    icb_compat::check(icb_mapped_length(icb_compat::ctx(), pos.data(), idx.data(), int64_t(hits.size()), int64_t(mins.size()),
					int64_t(hpcSeq.Str().length()), nmax, &total),
		      "getMappedRatio");
    return double(total) / double(hpcSeq.Str().length());
}

inline int setGapOpen(double e) { return e <= 0.01 ? 5 : (e <= 0.04 ? 4 : (e <= 0.1 ? 3 : 2)); }

// semi-global alignment + traceback of one pair on the device (parasail_sg_trace_scan_16 / _32 semantics)
inline parasail_result_t* ParasailAlign(const std::string& read, const std::string& ref, int gapOpen, int gapExtend,
					const parasail_matrix_t* user_matrix)
{
    if (gapExtend != 1 || !user_matrix || user_matrix->match != 2 || user_matrix->mismatch != -2)
	throw std::invalid_argument("isonclust2_compat: ParasailAlign supports the path's scoring only (match 2, mismatch -2, extend 1)");
    const std::string seq = read + ref, qual(seq.size(), 'I');
    const int64_t offs[3] = {0, int64_t(read.size()), int64_t(seq.size())};
    icb_readset* rs = nullptr;
    icb_compat::check(icb_readset_create(icb_compat::ctx(), seq.data(), qual.data(), offs, 2, &rs), "ParasailAlign");
    icb_pair p{0, 1, 1, gapOpen, 0};
    icb_align_result res;
    std::string comp(seq.size() + 1, ' ');
    const int rc = icb_align_traceback(rs, &p, 1, &res, &comp[0], int64_t(comp.size()));
    icb_readset_destroy(rs);
    icb_compat::check(rc, "ParasailAlign");
    comp.resize(size_t(res.comp_len));
    return new parasail_result_t{res.score, res.end_query, res.end_ref, comp};
}

// cluster.cpp:442-459: windows counted on the device, limit and ratio are the reference's doubles
inline double getAlnRatio(const std::string& comp, double e, unsigned slen, unsigned kmerSize)
{
    const double limit = floor((1.0 - e) * kmerSize);
    const int lim = limit < -1.0 ? -1 : (limit > double(kmerSize + 1) ? int(kmerSize + 1) : int(limit));
    int64_t aligned = 0;
    icb_compat::check(icb_count_windows(icb_compat::ctx(), comp.data(), int64_t(comp.size()), int(kmerSize), lim, &aligned), "getAlnRatio");
    return double(aligned) / slen;
}

// cluster.cpp:622-636: the hit map is iterated and sorted with the host's std::sort, exactly as the reference does
inline SortedHits SortMinimizerHits(const MinimizerHits& hits, const Clusters&)
{
    SortedHits sorted;
    sorted.reserve(hits.size());
    for (auto& h : hits) sorted.emplace_back(new SortedHit{unsigned(h.second.size()), unsigned(h.first.first), h.first.second});
    std::sort(sorted.begin(), sorted.end(), [](const SortedHitP& a, const SortedHitP& b) { return a->Size > b->Size; });
    return sorted;
}

#endif
