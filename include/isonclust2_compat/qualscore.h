// qualscore.h of the reference's header face (reference src/qualscore.h:14-24, qualscore.cpp:14-37, 107-180).
// PrepareSortedBatch is the batched entry point icb_sketch_reads of the C ABI (it returns the reference's Batch object,
// which belongs to the serialisation layer, not to this face).
#ifndef QUALSCORE_H_INCLUDED
#define QUALSCORE_H_INCLUDED

#include <stdlib.h>

#include <string>
#include <vector>

#include "icb_compat.h"
#include "seq.h"

typedef std::vector<double> QualTab;

inline QualTab InitQualTab()
{
    QualTab t(129);
    icb_init_qual_tab(t.data(), 0);
    return t;
}

inline QualTab InitQualTabNomin()
{
    QualTab t(129);
    icb_init_qual_tab(t.data(), 1);
    return t;
}

namespace icb_compat {
struct FlatReads {
    std::string seq, qual;
    std::vector<int64_t> offs{0};
    void add(const Seq& s)
    {
	seq += s.Str();
	qual += s.Qual();
	offs.push_back(int64_t(seq.size()));
    }
};
}  // namespace icb_compat

// Score and error rate of every read, on the device (one launch for all of them)
inline void FillQualScores(SequencesP& sequences, int kmerSize, int windowSize, const QualTab& qualTab, const QualTab& qualTabNomin)
{
    (void)windowSize;
    if (sequences.empty()) return;
    icb_compat::FlatReads f;
    for (auto& s : sequences) f.add(*s);
    icb_readset* rs = nullptr;
    icb_compat::check(icb_readset_create(icb_compat::ctx(), f.seq.data(), f.qual.data(), f.offs.data(), int64_t(sequences.size()), &rs),
		      "FillQualScores");
    std::vector<double> score(sequences.size()), err(sequences.size());
    int rc = icb_score_reads(rs, kmerSize, qualTab.data(), qualTabNomin.data());
    if (rc == ICB_OK) rc = icb_scores_download(rs, score.data(), err.data());
    icb_readset_destroy(rs);
    icb_compat::check(rc, "FillQualScores");
    for (size_t i = 0; i < sequences.size(); i++) {
	sequences[i]->SetScore(score[i]);
	sequences[i]->SetErrorRate(err[i]);
    }
}

inline double CalcQualScore(const Seq& s, int kmerSize, const QualTab& qualTab)
{
    SequencesP one;
    one.emplace_back(new Seq(s));
    // FillQualScores' guard (reads of at most 2k bases score -1, qualscore.cpp:20-33) sits above CalcQualScore in the reference;
    // a read of at most k bases scores -1 inside it as well (qualscore.cpp:110-112)
    FillQualScores(one, kmerSize, 0, qualTab, qualTab);
    return one[0]->Score();
}

// stable, descending by score (the permutation comes from the library's stable merge sort)
inline void SortByQualScores(SequencesP& sequences)
{
    std::vector<double> score(sequences.size());
    for (size_t i = 0; i < sequences.size(); i++) score[i] = sequences[i]->Score();
    std::vector<int32_t> order(sequences.size());
    icb_compat::check(icb_sort_by_score(score.data(), int64_t(score.size()), order.data()), "SortByQualScores");
    SequencesP out;
    out.reserve(sequences.size());
    for (int32_t i : order) out.push_back(std::move(sequences[size_t(i)]));
    sequences.swap(out);
}

inline double CalcErrorRate(const std::string& quals, const QualTab& qualTab)
{
    double e = 0.0;
    const int64_t offs[2] = {0, int64_t(quals.size())};
    icb_compat::check(icb_calc_error_rate(icb_compat::ctx(), quals.data(), offs, 1, qualTab.data(), &e), "CalcErrorRate");
    return e;
}

#endif
