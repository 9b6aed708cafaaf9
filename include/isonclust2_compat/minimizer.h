// minimizer.h of the reference's header face (reference src/minimizer.h:20-93, minimizer.cpp:31-160).
#ifndef MINIMIZER_H_INCLUDED
#define MINIMIZER_H_INCLUDED

#include <algorithm>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "icb_compat.h"
#include "kmer_index.h"

class Seq;

typedef icb_minimizer Minimizer;  // {Min, Pos, Index}: the same three unsigned fields (minimizer.h:20-29)
typedef std::vector<Minimizer> Minimizers;
inline bool operator==(const Minimizer& a, const Minimizer& b) { return a.Min == b.Min && a.Pos == b.Pos && a.Index == b.Index; }

// window minimizers of a k-mer sequence, on the device
inline Minimizers GetKmerMinimizers(const KmerSeq& kmerSeq, int kmerSize, int windowSize)
{
    Minimizers out(kmerSeq.size());
    int64_t n = 0;
    icb_compat::check(icb_kmer_minimizers(icb_compat::ctx(), kmerSeq.data(), int64_t(kmerSeq.size()), kmerSize, windowSize, out.data(), &n),
		      "GetKmerMinimizers");
    out.resize(size_t(n));
    return out;
}

typedef std::pair<int, int> StrandedCluster;
struct StrandedClsHash {
    std::size_t operator()(const StrandedCluster& u) const { return std::size_t(int(u.first * u.second)); }
};

// The minimizer -> clusters index lives on the device; the object is a handle (copies share the index).
class MinimizerDB {
public:
    icb_index* handle() const
    {
	if (!h_) {
	    icb_index* ix = nullptr;
	    icb_compat::check(icb_index_create(icb_compat::ctx(), &ix), "MinimizerDB");
	    h_.reset(ix, icb_index_destroy);
	}
	return h_.get();
    }

private:
    mutable std::shared_ptr<icb_index> h_;
};

inline void AddMinimizers(const Minimizers& mins, unsigned cls, MinimizerDB& db)
{
    icb_compat::check(icb_index_add(db.handle(), cls, mins.data(), int64_t(mins.size())), "AddMinimizers");
}

inline void UpdateMinDB(int best, const Minimizers& oldMins, const Minimizers& newMins, MinimizerDB& db)
{
    icb_compat::check(icb_index_update_cluster(db.handle(), unsigned(best), oldMins.data(), int64_t(oldMins.size()), newMins.data(),
					       int64_t(newMins.size())),
		      "UpdateMinDB");
}

typedef struct {
    unsigned Pos;
    unsigned Index;
} MinimizerHit;
typedef std::vector<MinimizerHit> MinimizerHitVector;
typedef std::unordered_map<StrandedCluster, MinimizerHitVector, StrandedClsHash> MinimizerHits;

// The device returns the hits in the order the reference appends them; they go into the same container, built with the same
// bucket count (minimizer.cpp:48) in the same order, so iterating it — what SortMinimizerHits does — behaves as in the reference.
inline MinimizerHits GetMinimizerHits(const Minimizers& mins, const Minimizers& revMins, const MinimizerDB& db)
{
    MinimizerHits res(20 * (mins.size() + revMins.size()));
    std::vector<icb_hit> hits(4 * (mins.size() + revMins.size()) + 64);
    int64_t n = 0;
    for (;;) {
	icb_compat::check(icb_index_query_hits(db.handle(), mins.data(), int64_t(mins.size()), revMins.data(), int64_t(revMins.size()),
					       hits.data(), int64_t(hits.size()), &n),
			  "GetMinimizerHits");
	if (n <= int64_t(hits.size())) break;
	hits.resize(size_t(n));
    }
    for (int64_t i = 0; i < n; i++) res[std::make_pair(int(hits[i].cls), int(hits[i].strand))].push_back(MinimizerHit{hits[i].pos, hits[i].index});
    return res;
}

typedef struct {
    unsigned Size;
    unsigned Cls;
    int Strand;
} SortedHit;
typedef std::unique_ptr<SortedHit> SortedHitP;
typedef std::vector<SortedHitP> SortedHits;

#endif
