// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI wrapper around the UNMODIFIED reference sources under
// /root/reference (compiled in place by oracle/Makefile into
// oracle/_ref/libisonclust2_ref.so).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it, and only as the
// checker / the CPU baseline.  Nothing here is an algorithm: every function
// marshals flat buffers into the reference's own types and calls the
// reference's own function, cited beside each wrapper.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "args.h"
#include "cluster.h"
#include "cluster_data.h"
#include "consensus.h"
#include "hpc.h"
#include "kmer_index.h"
#include "minimizer.h"
#include "p_emp_prob.h"
#include "parasail.h"
#include "qualscore.h"
#include "seq.h"
#include "serialize.h"
#include "util.h"

extern bool VERBOSE;					    // args.cpp:8
extern UnsignedHash uh;					    // cluster.cpp:23
extern std::unique_ptr<spoa::AlignmentEngine> SpoaEngine;  // consensus.cpp:13

namespace {

const QualTab& qtab()
{
    static QualTab t = InitQualTab();  // qualscore.cpp:156
    return t;
}
const QualTab& qtabNomin()
{
    static QualTab t = InitQualTabNomin();  // qualscore.cpp:171
    return t;
}

struct RefBatch {
    BatchP batch;		// the reference's own Batch
    std::vector<double> score;	// per ORIGINAL read, after FillQualScores
    std::vector<double> err;
    std::vector<int32_t> order;	// sorted slot -> original read index
};

int idToIndex(const std::string& id)
{
    // shim names every read "read_<original index>"
    auto p = id.rfind('_');
    return atoi(id.c_str() + p + 1);
}

void ensureSpoa(int algo)
{
    // main.cpp:285-324
    auto algorithm = spoa::AlignmentType::kSW;
    if (algo == 1) algorithm = spoa::AlignmentType::kNW;
    if (algo == 2) algorithm = spoa::AlignmentType::kOV;
    SpoaEngine = std::unique_ptr<spoa::AlignmentEngine>(
	spoa::AlignmentEngine::Create(algorithm, 4, -8, -8, -4, -20, -1));
}

}  // namespace

extern "C" {

// hpc.cpp:4-32
int ref_hpc(const char* seq, const char* qual, int n, char* outSeq,
	    char* outQual)
{
    std::unique_ptr<Seq> s(
	new Seq("r", std::string(seq, n), std::string(qual, n), 0.0));
    std::unique_ptr<Seq> h(HomopolymerCompress(s));
    memcpy(outSeq, h->Str().data(), h->Str().size());
    memcpy(outQual, h->Qual().data(), h->Qual().size());
    return int(h->Str().size());
}

// util.cpp:13-38 ; returns -1 where the reference throws
int ref_revcomp(const char* seq, int n, char* out)
{
    try {
	auto r = RevComp(std::string(seq, n));
	memcpy(out, r.data(), r.size());
    }
    catch (...) {
	return -1;
    }
    return n;
}

// kmer_index.cpp:5-17
int ref_kmer_encode(const char* seq, int n, int k, uint32_t* out)
{
    auto v = KmerEncodeSeq(std::string(seq, n), unsigned(k));
    memcpy(out, v.data(), v.size() * sizeof(uint32_t));
    return int(v.size());
}

// minimizer.cpp:78-123 ; out = {Min,Pos,Index} triples
int ref_minimizers(const uint32_t* kmers, int n, int k, int w, uint32_t* out)
{
    KmerSeq ks(kmers, kmers + n);
    auto m = GetKmerMinimizers(ks, k, w);
    for (size_t i = 0; i < m.size(); i++) {
	out[3 * i] = m[i].Min;
	out[3 * i + 1] = m[i].Pos;
	out[3 * i + 2] = m[i].Index;
    }
    return int(m.size());
}

// qualscore.cpp:156-180
void ref_qualtabs(double* tab, double* tabNomin)
{
    memcpy(tab, qtab().data(), 129 * sizeof(double));
    memcpy(tabNomin, qtabNomin().data(), 129 * sizeof(double));
}

// qualscore.cpp:107-136
double ref_calc_qual_score(const char* qual, int n, int k)
{
    Seq s("r", std::string(n, 'A'), std::string(qual, n), 0.0);
    return CalcQualScore(s, k, qtab());
}

// qualscore.cpp:147-154
double ref_calc_error_rate(const char* qual, int n, int nomin)
{
    return CalcErrorRate(std::string(qual, n), nomin ? qtabNomin() : qtab());
}

// p_emp_prob.cpp:22-94 ; returns NaN where the reference throws
double ref_pmin_shared(int k, int w, double e1, double e2)
{
    static int ck = -1, cw = -1;
    static MinSharedMap msm;
    if (ck != k || cw != w) {
	msm = InitMinSharedMap(k, w);
	ck = k;
	cw = w;
    }
    try {
	return GetPMinShared(e1, e2, msm);
    }
    catch (...) {
	return std::nan("");
    }
}

// cluster.cpp:324-353 ; hits = {Pos,Index} pairs ; nMins = mins.size()
double ref_mapped_ratio(int k, int w, int hpcLen, double readHpcErr,
			double clsHpcErr, int nMins, const uint32_t* hits,
			int nHits, double minProbNoHits)
{
    static int ck = -1, cw = -1;
    static MinSharedMap msm;
    if (ck != k || cw != w) {
	msm = InitMinSharedMap(k, w);
	ck = k;
	cw = w;
    }
    Seq a("a", std::string(hpcLen, 'A'), std::string(hpcLen, 'I'), 0.0);
    a.SetErrorRate(readHpcErr);
    Seq b("b", "A", "I", 0.0);
    b.SetErrorRate(clsHpcErr);
    Minimizers mins(nMins);
    MinimizerHitVector hv(nHits);
    for (int i = 0; i < nHits; i++) {
	hv[i].Pos = hits[2 * i];
	hv[i].Index = hits[2 * i + 1];
    }
    return getMappedRatio(a, b, mins, hv, msm, minProbNoHits);
}

// cluster.cpp:425-440
int ref_set_gap_open(double e) { return setGapOpen(e); }

// cluster.cpp:408-423 + traceback.c:171-182 + cluster.cpp:442-459.
// comp (may be NULL) must hold n+m+1 bytes. Returns the aligned-window COUNT
// (getAlnRatio's numerator) so the caller compares an integer; ratio is
// written to *ratio as the reference computes it.
int ref_align(const char* read, int n, const char* ref, int m, int gapOpen,
	      double e, int k, int* score, int* endQuery, int* endRef,
	      char* comp, int* compLen, double* ratio, int* saturated16)
{
    auto mat = parasail_matrix_create("ACGT", 2, -2);  // cluster.cpp:479
    std::string a(read, n), b(ref, m);
    // saturation flag of the 16-bit pass, for the tests' information
    {
	auto r16 = parasail_sg_trace_scan_16(a.c_str(), n, b.c_str(), m,
					     gapOpen, 1, mat);
	if (saturated16) *saturated16 = parasail_result_is_saturated(r16);
	parasail_result_free(r16);
    }
    auto res = ParasailAlign(a, b, gapOpen, 1, mat);
    if (score) *score = res->score;
    if (endQuery) *endQuery = res->end_query;
    if (endRef) *endRef = res->end_ref;
    auto tr = parasail_result_get_traceback(res, a.c_str(), n, b.c_str(), m,
					    mat, '|', ' ', ' ');
    parasail_result_free(res);
    std::string c(tr->comp);
    parasail_traceback_free(tr);
    parasail_matrix_free(mat);
    if (comp) memcpy(comp, c.c_str(), c.size() + 1);
    if (compLen) *compLen = int(c.size());
    double r = getAlnRatio(c, e, unsigned(n), unsigned(k));
    if (ratio) *ratio = r;
    // r = aligned / n with aligned an exact small integer
    return int(r * double(n) + 0.5);
}

// scalar spec of the trace table: parasail sg_trace.c ("orig") through
// parasail_sg_trace ; fills the n*m int8 trace table for inspection.
int ref_align_scalar_trace(const char* read, int n, const char* ref, int m,
			   int gapOpen, int8_t* trace, int* endQuery,
			   int* endRef)
{
    auto mat = parasail_matrix_create("ACGT", 2, -2);
    auto res = parasail_sg_trace(read, n, ref, m, gapOpen, 1, mat);
    int score = res->score;
    *endQuery = res->end_query;
    *endRef = res->end_ref;
    if (trace)
	memcpy(trace, parasail_result_get_trace_table(res),
	       size_t(n) * size_t(m));
    parasail_result_free(res);
    parasail_matrix_free(mat);
    return score;
}

// ---------------------------------------------------------------- batches

// FASTQ-less front end of mainSort (main.cpp:118-122): FillQualScores +
// SortByQualScores, then PrepareSortedBatch over the whole sorted range
// (main.cpp:177-186).  Reads are named "read_<i>".  nThreads<=0: TBB default.
void* ref_sort_and_sketch(const char* seqs, const char* quals,
			  const int64_t* offs, int nReads, int k, int w,
			  double minQual, int doSketch)
{
    auto rb = new RefBatch;
    SequencesP sequences;
    sequences.reserve(nReads);
    for (int i = 0; i < nReads; i++) {
	auto len = size_t(offs[i + 1] - offs[i]);
	sequences.emplace_back(
	    new Seq("read_" + std::to_string(i), std::string(seqs + offs[i], len),
		    std::string(quals + offs[i], len), 0.0));
    }
    FillQualScores(sequences, k, w, qtab(), qtabNomin());
    rb->score.resize(nReads);
    rb->err.resize(nReads);
    for (int i = 0; i < nReads; i++) {
	rb->score[i] = sequences[i]->Score();
	rb->err[i] = sequences[i]->ErrorRate();
    }
    SortByQualScores(sequences);
    rb->order.resize(nReads);
    unsigned long bases = 0;
    for (int i = 0; i < nReads; i++) {
	rb->order[i] = idToIndex(sequences[i]->Name());
	bases += sequences[i]->Str().length();
    }
    if (doSketch && nReads > 0) {
	rb->batch = BatchP(PrepareSortedBatch(sequences, 0, nReads - 1,
					       int(bases), k, w, minQual, qtab(),
					       qtabNomin()));
	rb->batch->BatchNr = 0;
	rb->batch->BatchBases = bases;
	rb->batch->TotalReads = 0;  // uninitialised in the reference
	CmdArgs a;
	a.KmerSize = k;
	a.WindowSize = w;
	a.MinQual = minQual;
	rb->batch->SortArgs = a;
    }
    return rb;
}

void ref_batch_free(void* h) { delete static_cast<RefBatch*>(h); }

void ref_batch_scores(void* h, double* score, double* err, int32_t* order)
{
    auto rb = static_cast<RefBatch*>(h);
    memcpy(score, rb->score.data(), rb->score.size() * sizeof(double));
    memcpy(err, rb->err.data(), rb->err.size() * sizeof(double));
    memcpy(order, rb->order.data(), rb->order.size() * sizeof(int32_t));
}

// totals of the sketch, to size the export buffers
void ref_batch_sketch_sizes(void* h, int64_t* hpcBytes, int64_t* nMins,
			    int64_t* nRevMins)
{
    auto rb = static_cast<RefBatch*>(h);
    int64_t a = 0, b = 0, c = 0;
    for (auto& cl : rb->batch->Cls) {
	auto& p = cl->at(0);
	if (p->HpcSeq) a += int64_t(p->HpcSeq->Str().size());
	b += int64_t(p->Mins.size());
	c += int64_t(p->RevMins.size());
    }
    *hpcBytes = a;
    *nMins = b;
    *nRevMins = c;
}

// Flat export of a freshly sketched batch, slot order = sorted order.
// status: 0 = filtered (Raw and Hpc null), 1 = sketched, 2 = raw kept / no hpc
void ref_batch_sketch_export(void* h, int32_t* status, char* hpcSeq,
			     char* hpcQual, int64_t* hpcOffs, double* hpcErr,
			     double* rawScore, uint32_t* mins, int64_t* minOffs,
			     uint32_t* revMins, int64_t* revOffs)
{
    auto rb = static_cast<RefBatch*>(h);
    int64_t ho = 0, mo = 0, ro = 0;
    size_t i = 0;
    for (auto& cl : rb->batch->Cls) {
	auto& p = cl->at(0);
	hpcOffs[i] = ho;
	minOffs[i] = mo;
	revOffs[i] = ro;
	status[i] = p->RawSeq ? (p->HpcSeq ? 1 : 2) : 0;
	hpcErr[i] = p->HpcSeq ? p->HpcSeq->ErrorRate() : -1.0;
	rawScore[i] = p->RawSeq ? p->RawSeq->Score() : -2.0;
	if (p->HpcSeq) {
	    auto n = p->HpcSeq->Str().size();
	    memcpy(hpcSeq + ho, p->HpcSeq->Str().data(), n);
	    memcpy(hpcQual + ho, p->HpcSeq->Qual().data(), n);
	    ho += int64_t(n);
	}
	for (auto& m : p->Mins) {
	    mins[3 * mo] = m.Min;
	    mins[3 * mo + 1] = m.Pos;
	    mins[3 * mo + 2] = m.Index;
	    mo++;
	}
	for (auto& m : p->RevMins) {
	    revMins[3 * ro] = m.Min;
	    revMins[3 * ro + 1] = m.Pos;
	    revMins[3 * ro + 2] = m.Index;
	    ro++;
	}
	i++;
    }
    hpcOffs[i] = ho;
    minOffs[i] = mo;
    revOffs[i] = ro;
}

// Split [from, to] (sorted slots, inclusive) out of a sketched batch as a new
// batch handle numbered batchNr, the way mainSort cuts batches
// (main.cpp:149-199).  The clusters are MOVED out of the parent.
void* ref_batch_slice(void* h, int from, int to, int batchNr)
{
    auto rb = static_cast<RefBatch*>(h);
    auto nb = new RefBatch;
    nb->batch = BatchP(new Batch);
    auto& b = nb->batch;
    unsigned long bases = 0;
    for (int i = from; i <= to; i++) {
	auto& cl = rb->batch->Cls[i];
	if (cl->at(0)->RawSeq) bases += cl->at(0)->RawSeq->Str().length();
	b->Cls.push_back(cl);
    }
    b->NrCls = int(b->Cls.size());
    b->BatchStart = from;
    b->BatchEnd = to;
    b->Depth = -1;
    b->BatchNr = batchNr;
    b->BatchBases = bases;
    b->TotalReads = 0;
    b->SortArgs = rb->batch->SortArgs;
    return nb;
}

// mainCluster (main.cpp:238-382) without file I/O.  right == NULL: initial
// clustering through CreatePseudoBatch; else merge left <- right.
// mode: 0 sahlin, 1 fast, 2 furious.  Returns number of clusters.
int ref_cluster(void* left, void* right, int mode, int minShared,
		double mappedTh, double alignedTh, double minFrac,
		double minProbNoHits, int minClsSize, int spoaAlgo)
{
    auto lb = static_cast<RefBatch*>(left);
    auto& leftBatch = lb->batch;
    BatchP pseudo;
    BatchP* rightP = nullptr;
    auto setArgs = [&](CmdArgs& a) {
	a.Mode = ClsMode(mode);
	a.MinShared = minShared;
	a.MappedThreshold = mappedTh;
	a.AlignedThreshold = alignedTh;
	a.MinFraction = minFrac;
	a.MinProbNoHits = minProbNoHits;
	if (minClsSize > 0) a.MinClsSize = minClsSize;
    };
    setArgs(leftBatch->SortArgs);
    if (right == nullptr) {
	pseudo = CreatePseudoBatch(leftBatch);	// main.cpp:262
	leftBatch->Cls.clear();
	if (leftBatch->Depth > 0) leftBatch->Depth = -leftBatch->Depth;
	leftBatch->NrCls = 0;
	leftBatch->MinDB = MinimizerDB(MIN_DB_RESERVE, uh);
	rightP = &pseudo;
    }
    else {
	auto rbp = static_cast<RefBatch*>(right);
	rbp->batch->MinDB = MinimizerDB(0, uh);
	rightP = &rbp->batch;
    }
    setArgs((*rightP)->SortArgs);
    ensureSpoa(spoaAlgo);
    VERBOSE = false;
    ClusterSortedReads(leftBatch, *rightP, true, false);
    return int(leftBatch->Cls.size());
}

unsigned ref_aln_invoked() { return AlnInvoked(); }

// cluster membership: for cluster c, members (original read indices parsed
// from ProcSeq::Id) and their MatchStrand, INCLUDING the synthetic rep at
// slot 0 (its Id is the founder's).  sizes[c] = members in cluster c.
// number of clusters of a (loaded or clustered) batch
int ref_n_clusters(void* h) { return int(static_cast<RefBatch*>(h)->batch->Cls.size()); }

void ref_cluster_sizes(void* h, int32_t* sizes)
{
    auto rb = static_cast<RefBatch*>(h);
    size_t i = 0;
    for (auto& cl : rb->batch->Cls) sizes[i++] = int32_t(cl->size());
}

void ref_cluster_members(void* h, int32_t* ids, int32_t* strands)
{
    auto rb = static_cast<RefBatch*>(h);
    size_t i = 0;
    for (auto& cl : rb->batch->Cls) {
	for (auto& p : *cl) {
	    ids[i] = idToIndex(p->Id);
	    strands[i] = p->MatchStrand;
	    i++;
	}
    }
}

// MinDB in container iteration order (what SaveBatch would serialise)
int64_t ref_mindb_size(void* h, int64_t* totalPostings)
{
    auto rb = static_cast<RefBatch*>(h);
    int64_t t = 0;
    for (auto& kv : rb->batch->MinDB) t += int64_t(kv.second.size());
    *totalPostings = t;
    return int64_t(rb->batch->MinDB.size());
}

void ref_mindb_export(void* h, uint32_t* keys, int64_t* offs, uint32_t* cls)
{
    auto rb = static_cast<RefBatch*>(h);
    int64_t o = 0;
    size_t i = 0;
    for (auto& kv : rb->batch->MinDB) {
	keys[i] = kv.first;
	offs[i] = o;
	for (auto c : kv.second) cls[o++] = c;
	i++;
    }
    offs[i] = o;
}

// serialize.cpp:5-10
void ref_save_batch(void* h, const char* path)
{
    auto rb = static_cast<RefBatch*>(h);
    SaveBatch(rb->batch, std::string(path));
}

// serialize.cpp:12-27
void* ref_load_batch(const char* path)
{
    auto rb = new RefBatch;
    rb->batch = LoadBatch(std::string(path));
    return rb;
}

// getBestCluster's candidate order for ONE read against a MinDB built from
// `nCls` founders' forward minimizers: GetMinimizerHits (minimizer.cpp:44-76)
// + SortMinimizerHits (cluster.cpp:622-636).  out = {Size,Cls,Strand} triples
// in the reference's (libstdc++-dependent) order.
int ref_sorted_hits(const uint32_t* clsMins, const int64_t* clsOffs, int nCls,
		    const uint32_t* mins, int nMins, const uint32_t* revMins,
		    int nRev, int32_t* out)
{
    MinimizerDB db(MIN_DB_RESERVE, uh);
    for (int c = 0; c < nCls; c++) {
	Minimizers m;
	for (int64_t j = clsOffs[c]; j < clsOffs[c + 1]; j++)
	    m.push_back(Minimizer{clsMins[3 * j], clsMins[3 * j + 1],
				  clsMins[3 * j + 2]});
	AddMinimizers(m, unsigned(c), db);
    }
    Minimizers a, b;
    for (int j = 0; j < nMins; j++)
	a.push_back(Minimizer{mins[3 * j], mins[3 * j + 1], mins[3 * j + 2]});
    for (int j = 0; j < nRev; j++)
	b.push_back(
	    Minimizer{revMins[3 * j], revMins[3 * j + 1], revMins[3 * j + 2]});
    auto hits = GetMinimizerHits(a, b, db);
    Clusters cls;
    auto order = SortMinimizerHits(hits, cls);
    int i = 0;
    for (auto& s : order) {
	out[3 * i] = int32_t(s->Size);
	out[3 * i + 1] = int32_t(s->Cls);
	out[3 * i + 2] = s->Strand;
	i++;
    }
    return i;
}

// UpdateMinDB (minimizer.cpp:124-160) on a MinDB built by AddMinimizers (minimizer.cpp:31-42) from `nCls` founders'
// forward minimizers: update u replaces the minimizers of cluster updCls[u] (its current ones are the "old" set) by
// updMins[updOffs[u] .. updOffs[u+1]).  Exports the resulting map sorted by key (empty posting lists included):
// keys[], offs[nKeys+1], cls[].  Returns nKeys, or -1 - needed if a capacity is too small.
int64_t ref_mindb_after_updates(const uint32_t* clsMins, const int64_t* clsOffs,
				int nCls, const int32_t* updCls,
				const uint32_t* updMins, const int64_t* updOffs,
				int nUpd, uint32_t* keys, int64_t* offs,
				uint32_t* cls, int64_t keyCap, int64_t clsCap)
{
    MinimizerDB db(MIN_DB_RESERVE, uh);
    std::vector<Minimizers> cur(nCls);
    for (int c = 0; c < nCls; c++) {
	for (int64_t j = clsOffs[c]; j < clsOffs[c + 1]; j++)
	    cur[c].push_back(Minimizer{clsMins[3 * j], clsMins[3 * j + 1],
				       clsMins[3 * j + 2]});
	AddMinimizers(cur[c], unsigned(c), db);
    }
    for (int u = 0; u < nUpd; u++) {
	Minimizers nm;
	for (int64_t j = updOffs[u]; j < updOffs[u + 1]; j++)
	    nm.push_back(Minimizer{updMins[3 * j], updMins[3 * j + 1],
				   updMins[3 * j + 2]});
	UpdateMinDB(updCls[u], cur[updCls[u]], nm, db);
	cur[updCls[u]] = nm;
    }
    std::vector<uint32_t> ks;
    int64_t tot = 0;
    for (auto& kv : db) {
	ks.push_back(kv.first);
	tot += int64_t(kv.second.size());
    }
    if (int64_t(ks.size()) > keyCap || tot > clsCap)
	return -1 - std::max<int64_t>(int64_t(ks.size()), tot);
    std::sort(ks.begin(), ks.end());
    int64_t o = 0;
    for (size_t i = 0; i < ks.size(); i++) {
	keys[i] = ks[i];
	offs[i] = o;
	for (auto c : db[ks[i]]) cls[o++] = c;
    }
    offs[ks.size()] = o;
    return int64_t(ks.size());
}

// ---------------------------------------------------------------- spoa (vendor/spoa 4.0.9), as consensus.cpp drives it
// A graph handle is a spoa::Graph; ref_poa_add = AddSeqToGraph (consensus.cpp:15-22), ref_poa_consensus =
// Graph::GenerateConsensus (graph.cpp:368-375), ref_poa_save / ref_poa_load = the cereal binary archive of the graph alone
// (graph.hpp:188-276), ref_poa_align = AlignmentEngine::Align without adding.
void* ref_poa_new() { return new spoa::Graph(); }
void ref_poa_free(void* g) { delete static_cast<spoa::Graph*>(g); }
void ref_poa_clear(void* g) { static_cast<spoa::Graph*>(g)->Clear(); }
int64_t ref_poa_num_sequences(void* g) { return int64_t(static_cast<spoa::Graph*>(g)->sequences().size()); }
int64_t ref_poa_num_nodes(void* g) { return int64_t(static_cast<spoa::Graph*>(g)->nodes().size()); }
void ref_poa_add(void* g, const char* seq, int n, unsigned weight, int algo)
{
    ensureSpoa(algo);
    AddSeqToGraph(std::string(seq, size_t(n)), static_cast<spoa::Graph*>(g), SpoaEngine.get(), weight);
}
int64_t ref_poa_align(void* g, const char* seq, int n, int algo, int32_t* nodes, int32_t* pos, int64_t cap, int32_t* score)
{
    ensureSpoa(algo);
    std::int32_t sc = 0;
    auto a = SpoaEngine->Align(std::string(seq, size_t(n)), *static_cast<spoa::Graph*>(g), &sc);
    if (score) *score = sc;
    if (int64_t(a.size()) <= cap)
	for (size_t i = 0; i < a.size(); i++) {
	    nodes[i] = a[i].first;
	    pos[i] = a[i].second;
	}
    return int64_t(a.size());
}
int64_t ref_poa_consensus(void* g, char* out, int64_t cap)
{
    std::string c = static_cast<spoa::Graph*>(g)->GenerateConsensus();
    if (int64_t(c.size()) <= cap) memcpy(out, c.data(), c.size());
    return int64_t(c.size());
}
int64_t ref_poa_save(void* g, char* out, int64_t cap)
{
    std::ostringstream os(std::ios::binary);
    {
	cereal::BinaryOutputArchive ar(os);
	ar(*static_cast<spoa::Graph*>(g));
    }
    const std::string s = os.str();
    if (int64_t(s.size()) <= cap) memcpy(out, s.data(), s.size());
    return int64_t(s.size());
}
void* ref_poa_load(const char* bytes, int64_t n)
{
    std::istringstream is(std::string(bytes, size_t(n)), std::ios::binary);
    auto g = new spoa::Graph();
    cereal::BinaryInputArchive ar(is);
    ar(*g);
    return g;
}

}  // extern "C"

