// The greedy, order-dependent assignment loop (host) on top of the stage-2 and stage-3 kernels.
//
// Mirrors ClusterSortedReads (reference cluster.cpp:67-322, consensus off — the default, args.h:19) and
// getBestCluster / getBestClusterMapping / getBestClusterAln (cluster.cpp:355-406, 461-515, 530-568).  The
// host keeps exactly what is sequential in the reference: walking the sorted reads in order, deciding
// join-or-found, and — for the rare reads where two equally scored candidates both pass — the candidate
// ORDER, which the reference defines through libstdc++ (iteration order of an unordered_map with
// StrandedClsHash and 20*(|mins|+|revMins|) buckets, minimizer.cpp:48, minimizer.h:53-58, followed by an
// unstable std::sort, cluster.cpp:622-636); the same two library facilities are used here on the device's
// (cluster, strand, count, first-hit) records.  Everything data-parallel — hit counting, mapped lengths,
// alignments — is computed on the device per CHUNK of reads, speculatively where the sequential dependency
// demands it (see index.cu), and consumed here in order.
#include <algorithm>
#include <chrono>
#include <ctime>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <unordered_map>

#include <atomic>
#include <functional>
#include <thread>

#include "align.cuh"
#include "dloop.cuh"
#include "index.cuh"
#include "helpers.h"
#include "ref_order.h"
#include "ref_order_hd.h"

// reads per chunk of the greedy loop when the caller does not say: fewer, longer rounds of the device loop than 256 (7.4 k
// instead of 12.4 k rounds per 1 M reads) at a few more candidate-order evaluations per read
#define ICB_DEFAULT_CHUNK 512

namespace {

typedef IcbSortedCand SortedCand;

struct AlnKey {
    int pos, id, strand;
    bool operator<(const AlnKey& o) const
    {
	if (pos != o.pos) return pos < o.pos;
	if (id != o.id) return id < o.id;
	return strand < o.strand;
    }
};

enum Outcome { OUT_JOIN, OUT_NEW, OUT_NEED, OUT_NONE };

// per-thread scratch of the decision function (no allocation in steady state) + what it reports besides the outcome
struct DecideScratch {
    std::vector<IcbCand> ro_real, v, cs;
    std::vector<SortedCand> ro_out;
    std::vector<const IcbCand*> pass, tied;
    IcbRefOrderScratch ro;
    std::vector<AlnKey> needPairs;
    std::vector<int32_t> needFull;
    bool usedAln = false, usedTie = false;
    int alnCalls = 0;  // DP calls the reference makes for the read being decided (getBestClusterAln stops at the first pass)
    double ms_order = 0;
    int64_t order_calls = 0, decisions = 0, hyp_decisions = 0;
};

double thread_cpu_ms()
{
    timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

struct Stopwatch {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap()
    {
	auto t1 = std::chrono::steady_clock::now();
	double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	t0 = t1;
	return ms;
    }
};

}  // namespace

struct icb_clusterer {
    icb_ctx* ctx = nullptr;  // the context (stream) this clusterer launches on: the read set's, or a worker's child context
    icb_readset* rs = nullptr;
    icb_cluster_params P{};
    int32_t nmax[225];
    // host mirror of per-read scalars lives in the read set (shared by all clusterers on it)
    // clusters
    std::vector<int32_t> founder;  // cluster id -> founder read
    // device state
    IcbIndex main_idx, chunk_idx;
    int32_t* d_nmax = nullptr;
    int32_t* d_cls_founder = nullptr;
    int64_t cls_founder_cap = 0;
    int32_t* d_chunk_reads = nullptr;
    int32_t* d_list = nullptr;
    int32_t* d_ids = nullptr;
    IcbQueryHeader* d_hdr = nullptr;
    IcbCand* d_cands = nullptr;
    int64_t cands_cap = 0;
    int64_t* d_full_off = nullptr;
    uint32_t* d_scratch = nullptr;
    int64_t scratch_words = 0;
    int chunk_cap = 0;
    icb_cluster_stats stats{};
    std::vector<std::unique_ptr<DecideScratch>> scratch;  // [0] the batch worker itself, [1..] tasks handed to the helper pool
    // pinned host staging (asynchronous copies, no hidden stream synchronisation on pageable memory)
    int32_t* h_i32 = nullptr;	 // 3 * chunk_cap: chunk reads / ids / list
    IcbQueryHeader* h_hdr = nullptr;
    IcbCand* h_cands = nullptr;
    int64_t h_cands_cap = 0;
    int64_t* h_off = nullptr;	 // chunk_cap + 1

    // device-resident loop (dloop.cu): the chunk's records, pools and work queues
    bool use_dloop = true;
    int dl_reason = 0;
    bool dl_retry = false;   // the device loop enlarged a pool: the chunk goes through it again
    int force_fallback = 0;  // test hook (ICB_DLOOP_FORCE_FALLBACK=m): chunks with (C0 + N) % m == 0 go back to the host loop
    IcbDLoopView V{};
    IcbQueryHeader* d_hdr_full = nullptr;
    int32_t* d_n_ins = nullptr;
    uint32_t cand_cap = 0;
    unsigned long long full_cap = 0;
    // one round of the device loop as a captured graph: one driver call instead of ~17 launches per round (35 worker threads
    // launching kernel by kernel are bound by the driver's launch rate, not by the device)
    struct RoundKey {
	IcbDLoopView view;  // every pointer and parameter the kernels of the round take by value
	const void* p[12];
	size_t smem_f;
	int nb_f, smem_cnt_f, iters, mode, max_m, k, spec, aln_priority;
	bool operator==(const RoundKey& o) const { return memcmp(this, &o, sizeof(RoundKey)) == 0; }
    };
    cudaGraphExec_t round_exec = nullptr;
    RoundKey round_key{};
    int round_nodes = 0;
    int qf_bound = 0;		       // snapshot size the full-list query's launch configuration is planned for
    IcbQueryArgs* d_qf_args = nullptr;
    IcbQueryArgs* h_qf_args = nullptr;  // pinned
    IcbDLoopState* h_state = nullptr;  // pinned
    int32_t* h_out = nullptr;	       // pinned: 3 * chunk_cap (cluster, strand, founder positions)
    double rounds_avg = 3.0;

    int init_buffers();
    void free_buffers();
    int init_dloop();
    void free_dloop();
    int grow_cand_pool(uint32_t want);
    int grow_full_pool(unsigned long long want);
    int ensure_founder_cap(int64_t n);
    int plan_query(int Ctot, int cap, int mode, IcbQueryArgs& a, int n_items_max, int& n_blocks, size_t& smem_bytes);
    int process_chunk_device(const std::vector<int32_t>& chunk, const std::vector<int64_t>& slots, int32_t* cluster_of, int32_t* strand_of,
			     bool& handled);
    int reset();
    int seed(const int32_t* rep_reads, int64_t n);
    int update_representative(int32_t cluster, int32_t rep_read);

    int ensure_scratch(int64_t words);
    int ensure_cands(int64_t n);
    int ensure_cands_host(int64_t n);
    int query(const std::vector<int32_t>& chunk, int C0, const std::vector<int32_t>* list, int mode, int cap,
	      const std::vector<int32_t>* live);
    int upload_founders(int64_t from);
    int process_chunk(const std::vector<int32_t>& chunk, const std::vector<int64_t>& slots, int32_t* cluster_of, int32_t* strand_of);
    // the device-resident loop; the host loop takes the chunks it hands back
    int run_chunk(const std::vector<int32_t>& chunk, const std::vector<int64_t>& slots, int32_t* cluster_of, int32_t* strand_of)
    {
	bool handled = false;
	if (use_dloop) {
	    for (int attempt = 0; attempt < 6 && !handled; attempt++) {
		dl_retry = false;
		int rc = process_chunk_device(chunk, slots, cluster_of, strand_of, handled);
		if (rc) return rc;
		if (!dl_retry) break;
	    }
	    if (!handled) {
		stats.dl_fallback_chunks++;
		stats.dl_fallback_reasons |= int64_t(1) << (dl_reason & 31);
	    }
	}
	return handled ? ICB_OK : process_chunk(chunk, slots, cluster_of, strand_of);
    }
};

int icb_clusterer::ensure_scratch(int64_t words)
{
    if (words <= scratch_words) return ICB_OK;
    ICB_CUDA(ctx, ctx->wait());
    icb_dfree(ctx, d_scratch);
    d_scratch = nullptr;
    scratch_words = 0;
    const int64_t want = words + words / 2;  // geometric: the index grows with every chunk
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_scratch, size_t(want) * 4));
    scratch_words = want;
    return ICB_OK;
}

// pinned mirror only (pool-mode results are copied from the candidate pool)
int icb_clusterer::ensure_cands_host(int64_t n)
{
    if (n > h_cands_cap) {
	ICB_CUDA(ctx, ctx->wait());
	cudaFreeHost(h_cands);
	h_cands = nullptr;
	h_cands_cap = 0;
	const int64_t want = std::max<int64_t>(n + n / 2, 1 << 16);
	ICB_CUDA(ctx, cudaMallocHost((void**)&h_cands, size_t(want) * sizeof(IcbCand)));
	h_cands_cap = want;
    }
    return ICB_OK;
}

int icb_clusterer::ensure_cands(int64_t n)
{
    if (n > cands_cap) {
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, d_cands);
	d_cands = nullptr;
	cands_cap = 0;
	const int64_t want = std::max<int64_t>(n + n / 2, 1 << 16);
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_cands, size_t(want) * sizeof(IcbCand)));
	cands_cap = want;
    }
    if (n > h_cands_cap) {
	ICB_CUDA(ctx, ctx->wait());
	cudaFreeHost(h_cands);
	h_cands = nullptr;
	h_cands_cap = 0;
	const int64_t want = std::max<int64_t>(n + n / 2, 1 << 16);
	ICB_CUDA(ctx, cudaMallocHost((void**)&h_cands, size_t(want) * sizeof(IcbCand)));
	h_cands_cap = want;
    }
    return ICB_OK;
}

int icb_clusterer::init_buffers()
{
    const int cs = chunk_cap;
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_nmax, sizeof(int32_t) * 225));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_chunk_reads, sizeof(int32_t) * cs));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_list, sizeof(int32_t) * cs));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_ids, sizeof(int32_t) * cs));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_hdr, sizeof(IcbQueryHeader) * cs));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_full_off, sizeof(int64_t) * (cs + 1)));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_i32, sizeof(int32_t) * (3 * cs + 256)));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_hdr, sizeof(IcbQueryHeader) * cs));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_off, sizeof(int64_t) * (cs + 1)));
    memcpy(h_i32 + 3 * cs, nmax, sizeof(int32_t) * 225);
    ICB_CUDA(ctx, cudaMemcpyAsync(d_nmax, h_i32 + 3 * cs, sizeof(int32_t) * 225, cudaMemcpyHostToDevice, ctx->stream));
    int rc = main_idx.create(ctx, 1u << 16, 1 << 16);
    if (rc) return rc;
    rc = chunk_idx.create(ctx, 1u << 18, 1 << 17);
    if (rc) return rc;
    rc = init_dloop();
    if (rc) return rc;
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

// buffers of the device-resident loop; also used by the host loop for reads with more relevant candidates than its
// fixed per-read capacity (pool-mode query)
int icb_clusterer::init_dloop()
{
    const int cs = chunk_cap;
    const int mw = (cs + 31) / 32;
    // read per clusterer (tests switch engines inside one process)
    const bool off = getenv("ICB_HOST_LOOP") != nullptr;
    use_dloop = !off && mw <= DL_MAXW;
    force_fallback = getenv("ICB_DLOOP_FORCE_FALLBACK") ? atoi(getenv("ICB_DLOOP_FORCE_FALLBACK")) : 0;
    V = IcbDLoopView{};
    V.mw = mw;
    auto dm = [&](void** p, size_t bytes) { return icb_dmalloc(ctx, p, bytes); };
    ICB_CUDA(ctx, dm((void**)&V.st, sizeof(IcbDLoopState)));
    ICB_CUDA(ctx, dm((void**)&V.hdr, sizeof(IcbQueryHeader) * cs));
    ICB_CUDA(ctx, dm((void**)&d_hdr_full, sizeof(IcbQueryHeader) * cs));
    ICB_CUDA(ctx, dm((void**)&V.founder_cls, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.view_cls, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.outcome, cs));
    ICB_CUDA(ctx, dm((void**)&V.flags, cs));
    ICB_CUDA(ctx, dm((void**)&V.dec_cls, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.dec_strand, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.aln_calls, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.plan_n, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.plan_i, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.need_slot, 4 * cs * DL_SPEC));
    ICB_CUDA(ctx, dm((void**)&V.need_n, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.req_count, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.ord_req, cs));
    ICB_CUDA(ctx, dm((void**)&V.ord_target, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.order_list_s, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.order_list_l, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.full_off, 8 * cs));
    ICB_CUDA(ctx, dm((void**)&V.full_n, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.depmask, size_t(4) * cs * mw));
    ICB_CUDA(ctx, dm((void**)&V.seenmask, size_t(4) * cs * mw));
    ICB_CUDA(ctx, dm((void**)&V.out_cls, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.out_strand, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.founder_pos, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.descs16, sizeof(IcbPairDesc) * cs * DL_SPEC));
    ICB_CUDA(ctx, dm((void**)&V.descs32, sizeof(IcbPairDesc) * cs * DL_SPEC));
    ICB_CUDA(ctx, dm((void**)&V.full_list, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.full_list_off, 8 * cs));
    ICB_CUDA(ctx, dm((void**)&V.ins_reads, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&V.ins_ids, 4 * cs));
    ICB_CUDA(ctx, dm((void**)&d_n_ins, 16));
    ICB_CUDA(ctx, dm((void**)&d_qf_args, sizeof(IcbQueryArgs)));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_qf_args, sizeof(IcbQueryArgs)));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_state, sizeof(IcbDLoopState)));
    ICB_CUDA(ctx, cudaMallocHost((void**)&h_out, sizeof(int32_t) * 3 * cs));
    cand_cap = 0;
    full_cap = 0;
    int rc = grow_cand_pool(uint32_t(cs) * 256u);
    if (rc) return rc;
    if (use_dloop) rc = grow_full_pool((unsigned long long)cs * 4096ull);
    return rc;
}

void icb_clusterer::free_dloop()
{
    void* ps[] = {V.st, V.hdr, d_hdr_full, V.founder_cls, V.view_cls, V.outcome, V.flags, V.dec_cls, V.dec_strand, V.aln_calls, V.plan_n, V.plan_i,
		  V.need_slot, V.need_n, V.req_count, V.ord_req, V.ord_target, V.order_list_s, V.order_list_l, V.full_off, V.full_n, V.depmask, V.seenmask, V.out_cls, V.out_strand, V.founder_pos, V.descs16, V.descs32,
		  V.full_list, V.full_list_off, V.ins_reads, V.ins_ids, d_n_ins, V.cands, V.aln, V.plan, V.full};
    for (void* p : ps) icb_dfree(ctx, p);
    icb_dfree(ctx, d_qf_args);
    d_qf_args = nullptr;
    if (round_exec) cudaGraphExecDestroy(round_exec);
    round_exec = nullptr;
    cudaFreeHost(h_qf_args);
    h_qf_args = nullptr;
    cudaFreeHost(h_state);
    cudaFreeHost(h_out);
    V = IcbDLoopView{};
    h_state = nullptr;
    h_out = nullptr;
}

int icb_clusterer::grow_cand_pool(uint32_t want)
{
    if (want <= cand_cap) return ICB_OK;
    ICB_CUDA(ctx, ctx->wait());
    icb_dfree(ctx, V.cands);
    icb_dfree(ctx, V.aln);
    icb_dfree(ctx, V.plan);
    V.cands = nullptr;
    V.aln = V.plan = nullptr;
    cand_cap = 0;
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&V.cands, sizeof(IcbCand) * size_t(want)));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&V.aln, 4 * size_t(want)));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&V.plan, 4 * size_t(want)));
    cand_cap = want;
    return ICB_OK;
}

int icb_clusterer::grow_full_pool(unsigned long long want)
{
    if (want <= full_cap) return ICB_OK;
    ICB_CUDA(ctx, ctx->wait());
    icb_dfree(ctx, V.full);
    V.full = nullptr;
    full_cap = 0;
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&V.full, sizeof(IcbCand) * size_t(want)));
    full_cap = want;
    return ICB_OK;
}

void icb_clusterer::free_buffers()
{
    free_dloop();
    main_idx.destroy();
    chunk_idx.destroy();
    icb_dfree(ctx, d_nmax);
    icb_dfree(ctx, d_cls_founder);
    icb_dfree(ctx, d_chunk_reads);
    icb_dfree(ctx, d_list);
    icb_dfree(ctx, d_ids);
    icb_dfree(ctx, d_hdr);
    icb_dfree(ctx, d_cands);
    icb_dfree(ctx, d_full_off);
    icb_dfree(ctx, d_scratch);
    cudaFreeHost(h_i32);
    cudaFreeHost(h_hdr);
    cudaFreeHost(h_cands);
    cudaFreeHost(h_off);
}

// forget every cluster (the buffers and their capacities stay): the next batch starts from an empty index
int icb_clusterer::reset()
{
    founder.clear();
    stats = icb_cluster_stats{};
    return main_idx.clear();
}

// common part of every query launch: read set, indices, parameters
static void query_base(icb_clusterer* cl, IcbQueryArgs& a, int n_chunk, int C0)
{
    icb_readset* rs = cl->rs;
    a = IcbQueryArgs{};
    a.arena = rs->d_mins;
    a.min_off = rs->d_min_off;
    a.nf = rs->d_nf;
    a.nr = rs->d_nr;
    a.hpc_len = rs->d_hpc_len;
    a.read_class = rs->d_class;
    a.chunk_reads = cl->d_chunk_reads;
    a.n_chunk = n_chunk;
    a.main_idx = cl->main_idx.v;
    a.chunk_idx = cl->chunk_idx.v;
    a.C0 = C0;
    a.cls_founder = cl->d_cls_founder;
    a.nmax = cl->d_nmax;
    a.min_shared = cl->P.min_shared;
    a.min_frac = cl->P.min_fraction;
}

// shared-memory / global-scratch plan of one query launch: dense counters (Ctot words) + candidate list (2*cap, mode 0
// only) + bitmaps in shared memory when they fit, the counters (and mode 1's first-hit table) in per-CTA global scratch
// otherwise
int icb_clusterer::plan_query(int Ctot, int cap, int mode, IcbQueryArgs& a, int n_items_max, int& n_blocks, size_t& smem_bytes)
{
    const int bw = (rs->max_m + 31) / 32;
    const int bmWords = std::max(bw, 1024);
    int64_t words_smem = int64_t(Ctot) + 2 * int64_t(cap) + bmWords;
    const int64_t smem_limit_words = (200 * 1024) / 4;
    n_blocks = std::min(n_items_max, ctx->sm_count * 8);
    a.smem_cnt = words_smem <= smem_limit_words;
    int64_t stride = 0;
    if (!a.smem_cnt) {
	words_smem = 2 * int64_t(cap) + bmWords;
	stride += Ctot;
    }
    if (mode == 1 || mode == 2) stride += 2 * int64_t(Ctot);  // first-hit table
    if (words_smem > smem_limit_words) return ctx->fail(ICB_ERR_ARG, "a read has too many minimizers for the candidate query");
    a.cap = cap;
    a.mode = mode;
    a.smem_words = int32_t(words_smem);
    a.g_stride = stride;
    a.g_scratch = nullptr;
    if (stride > 0) {
	int rc = ensure_scratch(stride * n_blocks);
	if (rc) return rc;
	a.g_scratch = d_scratch;
    }
    smem_bytes = size_t(words_smem) * 4;
    return ICB_OK;
}

// one query launch of the HOST loop over (a subset of) the chunk; results land in the pinned mirrors h_hdr / h_cands.
// mode 0: `cap` records per item at item*cap.  mode 1: item t at h_off[t], sized by `live` (the n_total a mode-0
// query reported for that read).  mode 2: the relevant candidates of the listed reads without a per-read capacity
// (pool-mode query): item t at h_hdr[t].cand_off, h_hdr[t].n_cand records; `live` = the n_rel_all a mode-0 query reported.
int icb_clusterer::query(const std::vector<int32_t>& chunk, int C0, const std::vector<int32_t>* list, int mode, int cap,
			 const std::vector<int32_t>* live)
{
    cudaStream_t st = ctx->stream;
    const int n_chunk = int(chunk.size());
    const int n_items = list ? int(list->size()) : n_chunk;
    if (n_items == 0) return ICB_OK;
    const int Ctot = C0 + n_chunk;
    IcbQueryArgs a;
    query_base(this, a, n_chunk, C0);
    if (list) {
	int32_t* h_list = h_i32 + 2 * chunk_cap;
	memcpy(h_list, list->data(), sizeof(int32_t) * n_items);
	ICB_CUDA(ctx, cudaMemcpyAsync(d_list, h_list, sizeof(int32_t) * n_items, cudaMemcpyHostToDevice, st));
	a.list = d_list;
	a.n_list = n_items;
    }
    a.hdr = d_hdr;
    int64_t n_records;
    if (mode == 1) {
	h_off[0] = 0;
	for (int i = 0; i < n_items; i++) h_off[i + 1] = h_off[i] + (*live)[i];
	n_records = h_off[n_items];
	ICB_CUDA(ctx, cudaMemcpyAsync(d_full_off, h_off, sizeof(int64_t) * n_items, cudaMemcpyHostToDevice, st));
	a.full_off = d_full_off;
    }
    else if (mode == 2) {
	n_records = 0;
	for (int i = 0; i < n_items; i++) {
	    if ((*live)[i] > 32767) return ctx->fail(ICB_ERR_ARG, "a read has more than 32767 candidate clusters within minFraction of its best");
	    n_records += (*live)[i];
	}
	int rc = grow_cand_pool(uint32_t(std::min<int64_t>(n_records + n_records / 2 + 1024, 0xfffffff0ll)));
	if (rc) return rc;
	if (n_records > int64_t(cand_cap)) return ctx->fail(ICB_ERR_NOMEM, "candidate pool too large");
	rc = icb_dl_init(ctx, V, n_chunk, C0, cand_cap, full_cap, 0, 0);
	if (rc) return rc;
	a.dl_state = V.st;
	a.dl_pool = V.cands;
	a.dl_pool_aln = V.aln;
	a.dl_depmask = V.depmask;
	a.dl_mw = V.mw;
	cap = 0;
    }
    else
	n_records = int64_t(n_items) * cap;
    int rc = mode == 2 ? ensure_cands_host(std::max<int64_t>(n_records, 1)) : ensure_cands(std::max<int64_t>(n_records, 1));
    if (rc) return rc;
    a.cands = d_cands;
    int n_blocks;
    size_t smem_bytes;
    rc = plan_query(Ctot, cap, mode, a, n_items, n_blocks, smem_bytes);
    if (rc) return rc;
    rc = icb_query_launch(ctx, a, n_blocks, smem_bytes);
    if (rc) return rc;
    ICB_CUDA(ctx, cudaMemcpyAsync(h_hdr, d_hdr, sizeof(IcbQueryHeader) * n_items, cudaMemcpyDeviceToHost, st));
    if (n_records > 0)
	ICB_CUDA(ctx, cudaMemcpyAsync(h_cands, mode == 2 ? V.cands : d_cands, sizeof(IcbCand) * n_records, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

// AddMinimizers (minimizer.cpp:31-42) for clusters that exist already: the left batch of a merge
// (cluster.cpp:67-106 works on leftBatch->MinDB as loaded).  rep_reads[c] = the read whose forward sketch
// represents cluster `n_clusters + c`.
int icb_clusterer::seed(const int32_t* rep_reads, int64_t n)
{
    if (n <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    const int64_t C0 = int64_t(founder.size());
    int64_t tot = 0;
    for (int64_t i = 0; i < n; i++) {
	const int32_t r = rep_reads[i];
	if (r < 0 || r >= rs->n_reads || rs->h_status[r] != 1) return ctx->fail(ICB_ERR_ARG, "seed: representative is not a sketched read");
	tot += rs->h_nf[r];
    }
    int rc = main_idx.reserve(tot);
    if (rc) return rc;
    founder.insert(founder.end(), rep_reads, rep_reads + n);
    rc = upload_founders(C0);
    if (rc) return rc;
    // ids and reads of the new clusters, staged through the pinned block in pieces of chunk_cap
    for (int64_t done = 0; done < n; done += chunk_cap) {
	const int cnt = int(std::min<int64_t>(chunk_cap, n - done));
	for (int i = 0; i < cnt; i++) {
	    h_i32[i] = rep_reads[done + i];
	    h_i32[chunk_cap + i] = int32_t(C0 + done + i);
	}
	ICB_CUDA(ctx, cudaMemcpyAsync(d_chunk_reads, h_i32, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(d_ids, h_i32 + chunk_cap, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
	rc = icb_index_insert(ctx, main_idx, rs, d_chunk_reads, d_ids, cnt);
	if (rc) return rc;
	ICB_CUDA(ctx, ctx->wait());
    }
    main_idx.nodes_upper += tot;
    return ICB_OK;
}

// UpdateMinDB (minimizer.cpp:124-160) + the representative swap around it (cluster.cpp:279-291): from now on cluster
// `cluster` is represented by read `rep_read` — its forward sketch in the index, its compressed length / error class in
// the mapping test, its sequence in the alignment test.
int icb_clusterer::update_representative(int32_t cluster, int32_t rep_read)
{
    if (cluster < 0 || size_t(cluster) >= founder.size()) return ctx->fail(ICB_ERR_ARG, "update_representative: no such cluster");
    if (rep_read < 0 || rep_read >= rs->n_reads || rs->h_status[rep_read] != 1)
	return ctx->fail(ICB_ERR_ARG, "update_representative: representative is not a sketched read");
    const int32_t old = founder[cluster];
    if (old == rep_read) return ICB_OK;
    int rc = main_idx.reserve(rs->h_nf[rep_read]);
    if (rc) return rc;
    rc = icb_index_update(ctx, main_idx, rs, old, rep_read, cluster, rs->h_nf[old], rs->h_nf[rep_read]);
    if (rc) return rc;
    main_idx.nodes_upper += rs->h_nf[rep_read];
    founder[cluster] = rep_read;
    int32_t* stage = h_i32 + 2 * chunk_cap;
    stage[0] = rep_read;
    ICB_CUDA(ctx, cudaMemcpyAsync(d_cls_founder + cluster, stage, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

// room for n entries in the device table cluster -> representative read (the first founder.size() entries are kept)
int icb_clusterer::ensure_founder_cap(int64_t n)
{
    if (n <= cls_founder_cap) return ICB_OK;
    cudaStream_t st = ctx->stream;
    const int64_t keep = std::min<int64_t>(int64_t(founder.size()), cls_founder_cap);
    const int64_t ncap = std::max<int64_t>(cls_founder_cap * 2, n + 4096);
    int32_t* nd = nullptr;
    ICB_CUDA(ctx, ctx->wait());
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&nd, sizeof(int32_t) * ncap));
    if (keep > 0) ICB_CUDA(ctx, cudaMemcpyAsync(nd, d_cls_founder, sizeof(int32_t) * keep, cudaMemcpyDeviceToDevice, st));
    ICB_CUDA(ctx, ctx->wait());
    icb_dfree(ctx, d_cls_founder);
    d_cls_founder = nd;
    cls_founder_cap = ncap;
    return ICB_OK;
}

// device copy of founder[from..]
int icb_clusterer::upload_founders(int64_t from)
{
    cudaStream_t st = ctx->stream;
    const int64_t n = int64_t(founder.size());
    int rc = ensure_founder_cap(n);
    if (rc) return rc;
    int32_t* stage = h_i32 + 2 * chunk_cap;
    for (int64_t done = from; done < n; done += chunk_cap) {
	const int64_t cnt = std::min<int64_t>(chunk_cap, n - done);
	memcpy(stage, founder.data() + done, sizeof(int32_t) * cnt);
	ICB_CUDA(ctx, cudaMemcpyAsync(d_cls_founder + done, stage, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
	if (done + cnt < n) ICB_CUDA(ctx, ctx->wait());  // the staging block is reused by the next piece
    }
    return ICB_OK;
}

int icb_clusterer::process_chunk(const std::vector<int32_t>& chunk, const std::vector<int64_t>& slots, int32_t* cluster_of,
				 int32_t* strand_of)
{
    IcbRange nvtx_range("chunk: host loop");
    cudaStream_t st = ctx->stream;
    const std::vector<int32_t>&nf = rs->h_nf, &nr = rs->h_nr, &hpc_len = rs->h_hpc_len;
    const std::vector<double>& raw_err = rs->h_err;
    const int N = int(chunk.size());
    const int C0 = int(founder.size());
    const int k = P.kmer_size;
    const int CAP = 64;
    stats.chunks++;
    Stopwatch sw;
    // ---- chunk index: every read of the chunk as a provisional founder (id = C0 + position)
    {
	int64_t tot = 0;
	for (int i = 0; i < N; i++) {
	    h_i32[i] = chunk[i];
	    h_i32[chunk_cap + i] = C0 + i;
	    tot += nf[chunk[i]];
	}
	ICB_CUDA(ctx, cudaMemcpyAsync(d_chunk_reads, h_i32, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(d_ids, h_i32 + chunk_cap, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
	int rc = chunk_idx.clear();
	if (rc) return rc;
	rc = chunk_idx.reserve(tot);
	if (rc) return rc;
	rc = icb_index_insert(ctx, chunk_idx, rs, d_chunk_reads, d_ids, N);
	if (rc) return rc;
	chunk_idx.nodes_upper = tot;
    }
    stats.ms_insert += sw.lap();
    // ---- candidates of every read against snapshot + provisional founders
    int rc = query(chunk, C0, nullptr, 0, CAP, nullptr);
    if (rc) return rc;
    stats.reads_queried += N;
    std::vector<IcbQueryHeader> hdr(h_hdr, h_hdr + N);
    // per-read candidate lists (reads whose list overflowed are re-queried with room for every key)
    std::vector<std::vector<IcbCand>> cl(N);
    std::vector<int32_t> over;
    for (int i = 0; i < N; i++) {
	if (hdr[i].overflow) over.push_back(i);
	else cl[i].assign(h_cands + int64_t(i) * CAP, h_cands + int64_t(i) * CAP + hdr[i].n_cand);
    }
    if (!over.empty()) {
	// more relevant candidates than the fixed per-read capacity: asked again through the pool-mode query, which sizes
	// every read's list by what the first pass counted
	std::vector<int32_t> want(over.size());
	for (size_t t = 0; t < over.size(); t++) want[t] = hdr[over[t]].n_rel_all;
	rc = query(chunk, C0, &over, 2, 0, &want);
	if (rc) return rc;
	for (size_t t = 0; t < over.size(); t++) {
	    if (h_hdr[t].overflow) return ctx->fail(ICB_ERR_STATE, "candidate pool overflow");
	    cl[over[t]].assign(h_cands + h_hdr[t].cand_off, h_cands + h_hdr[t].cand_off + h_hdr[t].n_cand);
	}
    }

    stats.ms_query += sw.lap();
    std::vector<int> founder_cls(N, -1);  // provisional founder -> real cluster id once committed
    // A read's decision is a function of (its candidate lists, which of the provisional founders IN those lists are
    // committed, the alignment results it asked for).  Decisions are cached and only re-evaluated when one of the read's own
    // provisional candidates is committed (`dirty`) or device work it waited for has arrived (outcome OUT_NEED).
    struct Decision {
	int cls = -1, strand = 0, alnCalls = 0;
	bool usedAln = false, usedTie = false;
    };
    std::vector<Decision> dec(N);
    std::vector<char> eval_outcome(N, OUT_NONE);
    std::vector<char> dirty(N, 1);
    std::vector<std::vector<int32_t>> dependents(N);  // provisional founder -> later reads that list it as a candidate
    std::vector<char> committed(N, 0);
    // alignment results per read: (provisional id, strand) -> aligned window count; a read asks for one or two pairs
    struct AlnRes {
	int id, strand, aligned;
    };
    std::vector<std::vector<AlnRes>> aln(N);
    auto aln_find = [&](int pos, int id, int strand) -> const AlnRes* {
	for (auto& a : aln[pos])
	    if (a.id == id && a.strand == strand) return &a;
	return nullptr;
    };
    std::vector<std::vector<IcbCand>> full(N);	// full candidate lists of tie reads
    std::vector<char> have_full(N, 0);
    int nextCls = C0;
    int commit_pos = 0;
    for (int i = 0; i < N; i++)
	for (auto& c : cl[i])
	    if (c.id >= C0) dependents[c.id - C0].push_back(i);

    // `fmap` below is the map provisional founder -> cluster id a decision is taken under: founder_cls itself (the committed
    // founders: the decision counts), or founder_hyp (committed + predicted founders: only the device work the decision
    // would ask for is kept, see the speculative pass of the main loop)
    std::vector<int> founder_hyp;
    auto real_id_in = [&](const std::vector<int>& fmap, int id) { return id < C0 ? id : fmap[id - C0]; };
    auto rep_read = [&](int id) { return id < C0 ? founder[id] : chunk[id - C0]; };

    // exact candidate order of the reference for one read, restricted to real clusters (ref_order.h) and to the group of
    // candidates with `target` hits; `all` is the device's full list, ascending by id — the map provisional id -> real id
    // is monotone, so the order survives it
    struct OrderSlot {
	uint64_t set_hash = 0;
	int n_real = -1, target = -1;
	std::vector<SortedCand> order;
    };
    std::vector<OrderSlot> order_cache(2 * size_t(N));	// per read: [0] under the committed founders, [1] under the predicted ones
    auto reference_order = [&](DecideScratch& T, const std::vector<int>& fmap, int bank, int pos, const std::vector<IcbCand>& all,
			       int target) -> const std::vector<SortedCand>& {
	T.ro_real.clear();
	uint64_t h = 0x9E3779B97F4A7C15ull;
	for (auto& c : all) {
	    int rid = real_id_in(fmap, c.id);
	    if (rid < 0) continue;
	    T.ro_real.push_back(c);
	    T.ro_real.back().id = rid;
	    h = (h ^ (uint64_t(uint32_t(c.id)) << 32 | uint32_t(rid))) * 0xD1B54A32D192ED03ull;
	    h ^= h >> 29;
	}
	// the order is a function of the real candidates (which provisional ones count, and under which id)
	OrderSlot& slot = order_cache[2 * size_t(pos) + bank];
	if (slot.n_real == int(T.ro_real.size()) && slot.target == target && slot.set_hash == h) return slot.order;
	// the other bank may hold it: a prediction that came true leaves the committed founders with exactly the set (and
	// ids) the read was decided under one round earlier.  The passes over the two banks never overlap in time.
	const OrderSlot& other = order_cache[2 * size_t(pos) + (bank ^ 1)];
	if (other.n_real == int(T.ro_real.size()) && other.target == target && other.set_hash == h) return other.order;
	const int r = chunk[pos];
	Stopwatch t;
	// full[pos] is kept in the reference's insertion order (see where it is filled), and so is every subset of it
	icb_reference_order_group<true>(T.ro_real.data(), int(T.ro_real.size()), 20 * (size_t(nf[r]) + size_t(nr[r])), uint32_t(target),
					T.ro, T.ro_out);
	T.ms_order += t.lap();
	T.order_calls++;
	slot.n_real = int(T.ro_real.size());
	slot.set_hash = h;
	slot.target = target;
	slot.order = T.ro_out;
	return slot.order;
    };

    // decision for the read at `pos` given the founders committed so far.  Returns OUT_NEED when device work is
    // missing (alignment pairs appended to T.needPairs, tie reads to T.needFull).  Reads shared state only; writes
    // T and the read's own order_cache slot, so different reads may be decided concurrently.
    auto decide = [&](DecideScratch& T, const std::vector<int>& fmap, int bank, int pos, int& outCls, int& outStrand) -> Outcome {
	auto real_id = [&](int id) { return real_id_in(fmap, id); };
	bool& usedAln = T.usedAln;
	bool& usedTie = T.usedTie;
	int& alnCalls = T.alnCalls;
	std::vector<IcbCand>&v = T.v, &cs = T.cs;
	std::vector<const IcbCand*>&pass = T.pass, &tied = T.tied;
	std::vector<AlnKey>& needPairs = T.needPairs;
	std::vector<int32_t>& needFull = T.needFull;
	T.decisions++;
	usedAln = usedTie = false;
	alnCalls = 0;
	const int r = chunk[pos];
	v.clear();
	int top = hdr[pos].top_snap;
	for (auto& c : cl[pos]) {
	    if (c.id >= C0 && fmap[c.id - C0] < 0) continue;  // provisional founder that is no founder (yet)
	    v.push_back(c);
	    top = std::max(top, c.count);
	}
	outCls = -1;
	outStrand = 0;
	if (top == 0) return OUT_NEW;  // no hits at all (cluster.cpp:541-543)
	const bool mappingMode = P.mode == 0 || P.mode == 1;
	if (mappingMode && !(unsigned(top) < unsigned(P.min_shared))) {
	    const int thr = int(double(unsigned(top)) * P.min_fraction);  // cluster.cpp:387
	    // groups of equal count, best first
	    cs.clear();
	    for (auto& c : v)
		if (!(c.count < thr)) cs.push_back(c);
	    std::sort(cs.begin(), cs.end(), [](const IcbCand& a, const IcbCand& b) { return a.count > b.count; });
	    size_t g = 0;
	    while (g < cs.size()) {
		size_t e = g;
		pass.clear();
		while (e < cs.size() && cs[e].count == cs[g].count) {
		    float mr = float(double(cs[e].tmapped) / double(size_t(hpc_len[r])));  // cluster.cpp:351,390
		    if (mr >= P.mapped_threshold) pass.push_back(&cs[e]);
		    e++;
		}
		if (pass.size() == 1) {
		    outCls = real_id(pass[0]->id);
		    outStrand = pass[0]->strand;
		    return OUT_JOIN;
		}
		if (pass.size() > 1) {
		    // several equally scored candidates pass: the reference takes the first in ITS order
		    if (!have_full[pos]) {
			needFull.push_back(pos);
			return OUT_NEED;
		    }
		    usedTie = true;
		    const auto& order = reference_order(T, fmap, bank, pos, full[pos], cs[g].count);
		    for (auto& o : order) {
			for (auto* p : pass)
			    if (unsigned(real_id(p->id)) == o.Cls && p->strand == o.Strand) {
				outCls = int(o.Cls);
				outStrand = o.Strand;
				return OUT_JOIN;
			    }
		    }
		    return OUT_NEW;  // unreachable
		}
		g = e;
	    }
	}
	if (unsigned(top) < unsigned(P.min_shared)) return OUT_NEW;  // cluster.cpp:553-556
	if (P.mode == 1) return OUT_NEW;			      // fast: minimizers only
	// getBestClusterAln: candidates tied with the top count, in the reference's order, first passing wins
	// (cluster.cpp:480-512).  Like the reference, a candidate is only aligned once every candidate before it failed:
	// with several tied candidates the order is needed BEFORE the first alignment, so the full list is fetched first.
	tied.clear();
	for (auto& c : v)
	    if (c.count == top) tied.push_back(&c);
	const unsigned readLen = unsigned(rs->h_offs[r + 1] - rs->h_offs[r]);
	if (tied.size() == 1) {
	    const IcbCand* c = tied[0];
	    const AlnKey key{pos, c->id, c->strand};
	    const AlnRes* ia = aln_find(pos, c->id, c->strand);
	    if (!ia) {
		needPairs.push_back(key);
		return OUT_NEED;
	    }
	    usedAln = true;
	    alnCalls = 1;
	    if (double(ia->aligned) / readLen >= P.aligned_threshold) {	// cluster.cpp:458, 503
		outCls = real_id(c->id);
		outStrand = c->strand;
		return OUT_JOIN;
	    }
	    return OUT_NEW;
	}
	if (!have_full[pos]) {
	    needFull.push_back(pos);
	    return OUT_NEED;
	}
	usedTie = true;
	const auto& order = reference_order(T, fmap, bank, pos, full[pos], top);
	alnCalls = 0;
	for (auto& o : order) {
	    if (int(o.Size) != top) break;  // sorted by Size, descending: the tied candidates come first
	    const IcbCand* c = nullptr;
	    for (auto* p : tied)
		if (unsigned(real_id(p->id)) == o.Cls && p->strand == o.Strand) {
		    c = p;
		    break;
		}
	    if (!c) continue;
	    const AlnKey key{pos, c->id, c->strand};
	    const AlnRes* ia = aln_find(pos, c->id, c->strand);
	    if (!ia) {
		needPairs.push_back(key);
		return OUT_NEED;
	    }
	    alnCalls++;
	    if (double(ia->aligned) / readLen >= P.aligned_threshold) {
		usedAln = true;
		outCls = int(o.Cls);
		outStrand = o.Strand;
		return OUT_JOIN;
	    }
	}
	usedAln = true;
	return OUT_NEW;
    };

    if (scratch.empty()) scratch.emplace_back(new DecideScratch);
    IcbHelperPool& pool = IcbHelperPool::get();
    std::vector<int32_t> todo;
    std::vector<char> hyp_mark(N, 0);
    static const bool speculate_founders = !(getenv("ICB_NO_FOUNDER_SPECULATION"));
    // (re)decides the read if one of its provisional candidates was committed since, or if it waited for device work
    auto evaluate = [&](DecideScratch& T, int pos) -> Outcome {
	if (!dirty[pos] && eval_outcome[pos] != OUT_NEED) return Outcome(eval_outcome[pos]);
	int c, s;
	const Outcome o = decide(T, founder_cls, 0, pos, c, s);
	dirty[pos] = 0;
	eval_outcome[pos] = char(o);
	Decision& d = dec[pos];
	d.cls = c;
	d.strand = s;
	d.alnCalls = T.alnCalls;
	d.usedAln = T.usedAln;
	d.usedTie = T.usedTie;
	return o;
    };
    while (commit_pos < N) {
	stats.rounds++;
	std::vector<AlnKey> needPairs;
	std::vector<int32_t> needFull;
	DecideScratch& T0 = *scratch[0];
	T0.needPairs.clear();
	T0.needFull.clear();
	// ---- in order: commit until the first read that waits for device work
	int pos = commit_pos;
	for (; pos < N; pos++) {
	    const Outcome o = evaluate(T0, pos);
	    if (o == OUT_NEED) break;
	    const Decision& d = dec[pos];
	    if (d.usedAln) {
		stats.aln_invoked++;
		stats.aln_calls_ref += d.alnCalls;
	    }
	    if (d.usedTie) stats.tie_fallbacks++;
	    if (o == OUT_NEW) {
		founder_cls[pos] = nextCls;
		for (int32_t dep : dependents[pos]) dirty[dep] = 1;
		founder.push_back(chunk[pos]);
		cluster_of[slots[pos]] = nextCls;
		strand_of[slots[pos]] = 1;
		nextCls++;
	    }
	    else {
		cluster_of[slots[pos]] = d.cls;
		strand_of[slots[pos]] = d.strand;
	    }
	    committed[pos] = 1;
	    commit_pos = pos + 1;
	}
	// ---- behind the frontier: the set of founders is frozen now; every later read that has to be (re)decided is
	// evaluated only for the device work it will need — independent of each other, shared with the helper pool
	if (pos < N) {
	    todo.clear();
	    for (int q = pos + 1; q < N; q++)
		if (dirty[q] || eval_outcome[q] == OUT_NEED) todo.push_back(q);
	    const int per_task = 6;
	    int n_tasks = std::min((int(todo.size()) + per_task - 1) / per_task, pool.helpers() + 1);
	    if (n_tasks <= 1) {
		for (int32_t q : todo) evaluate(T0, q);
	    }
	    else {
		while (int(scratch.size()) < n_tasks + 1) scratch.emplace_back(new DecideScratch);
		for (int t = 1; t <= n_tasks; t++) {
		    scratch[t]->needPairs.clear();
		    scratch[t]->needFull.clear();
		}
		std::atomic<int> cursor(0);
		const std::function<void(int)> fn = [&](int t) {
		    DecideScratch& T = *scratch[t + 1];
		    for (;;) {
			const int i0 = cursor.fetch_add(per_task);
			if (i0 >= int(todo.size())) break;
			const int i1 = std::min(i0 + per_task, int(todo.size()));
			for (int i = i0; i < i1; i++) evaluate(T, todo[i]);
		    }
		};
		pool.parallel_for(n_tasks, fn);
	    }
	}
	// ---- one step further: reads behind the frontier that will found a cluster if nothing before them changes are
	// taken as founders already (with the ids they would get), and every later read that lists one of them is decided
	// once more under that assumption — only to collect the alignments / full lists it would ask for, so that they are
	// on the device's to-do list of THIS round trip instead of the one after the founder is committed.  Nothing of these
	// decisions is kept; results of the requested work are valid whatever happens (they are keyed by provisional id).
	if (pos < N && speculate_founders) {
	    founder_hyp = founder_cls;
	    int hyp_next = nextCls, n_pred = 0;
	    std::fill(hyp_mark.begin(), hyp_mark.end(), 0);
	    for (int q = pos + 1; q < N; q++)
		if (eval_outcome[q] == OUT_NEW) {
		    founder_hyp[q] = hyp_next++;
		    n_pred++;
		    for (int32_t dep : dependents[q]) hyp_mark[dep] = 1;
		}
	    todo.clear();
	    if (n_pred)
		for (int q = pos + 1; q < N; q++)
		    if (hyp_mark[q]) todo.push_back(q);
	    if (!todo.empty()) {
		const int per_task = 6;
		const int n_tasks = std::max(1, std::min((int(todo.size()) + per_task - 1) / per_task, pool.helpers() + 1));
		while (int(scratch.size()) < n_tasks + 1) scratch.emplace_back(new DecideScratch);
		std::atomic<int> cursor(0);
		const std::function<void(int)> fn = [&](int t) {
		    DecideScratch& T = *scratch[t + 1];
		    for (;;) {
			const int i0 = cursor.fetch_add(per_task);
			if (i0 >= int(todo.size())) break;
			const int i1 = std::min(i0 + per_task, int(todo.size()));
			for (int i = i0; i < i1; i++) {
			    int c, s;
			    decide(T, founder_hyp, 1, todo[i], c, s);
			    T.decisions--;  // not a decision of the greedy loop
			    T.hyp_decisions++;
			}
		    }
		};
		pool.parallel_for(n_tasks, fn);
	    }
	}
	for (auto& T : scratch) {
	    stats.hyp_decisions += T->hyp_decisions;
	    T->hyp_decisions = 0;
	    needPairs.insert(needPairs.end(), T->needPairs.begin(), T->needPairs.end());
	    needFull.insert(needFull.end(), T->needFull.begin(), T->needFull.end());
	    T->needPairs.clear();
	    T->needFull.clear();
	    stats.ms_order += T->ms_order;
	    stats.order_calls += T->order_calls;
	    stats.decisions += T->decisions;
	    T->ms_order = 0;
	    T->order_calls = T->decisions = 0;
	}
	stats.ms_host += sw.lap();
	if (commit_pos >= N) break;
	if (needPairs.empty() && needFull.empty()) return ctx->fail(ICB_ERR_STATE, "greedy loop made no progress");
	if (!needFull.empty()) {
	    std::sort(needFull.begin(), needFull.end());
	    needFull.erase(std::unique(needFull.begin(), needFull.end()), needFull.end());
	    std::vector<int32_t> live(needFull.size());
	    for (size_t t = 0; t < needFull.size(); t++) live[t] = hdr[needFull[t]].n_total;
	    rc = query(chunk, C0, &needFull, 1, 0, &live);
	    if (rc) return rc;
	    for (size_t t = 0; t < needFull.size(); t++) {
		const int pos = needFull[t];
		// the device writes the list ascending by id; it is kept in the reference's insertion order (strand, first hit,
		// id: GetMinimizerHits, minimizer.cpp:44-76), which every later evaluation of the candidate order starts from
		{
		    const IcbCand* src = h_cands + h_off[t];
		    const int nc = h_hdr[t].n_cand;
		    DecideScratch& Ts = *scratch[0];
		    icb_insertion_order(src, nc, Ts.ro, Ts.ro.ins);
		    full[pos].resize(size_t(nc));
		    for (int q = 0; q < nc; q++) full[pos][q] = src[Ts.ro.ins[q]];
		}
		have_full[pos] = 1;
		// the candidate ORDER depends on every real candidate of the full list, however few hits it has
		for (auto& c : full[pos])
		    if (c.id >= C0 && founder_cls[c.id - C0] < 0) dependents[c.id - C0].push_back(pos);
	    }
	    stats.ms_full += sw.lap();
	}
	if (!needPairs.empty()) {
	    std::sort(needPairs.begin(), needPairs.end());
	    needPairs.erase(std::unique(needPairs.begin(), needPairs.end(), [](const AlnKey& a, const AlnKey& b) {
				return a.pos == b.pos && a.id == b.id && a.strand == b.strand;
			    }),
			    needPairs.end());
	    std::vector<icb_pair> pairs(needPairs.size());
	    for (size_t t = 0; t < needPairs.size(); t++) {
		const AlnKey& key = needPairs[t];
		const int r = chunk[key.pos], rep = rep_read(key.id);
		const double e = raw_err[r] + raw_err[rep];  // cluster.cpp:494-496: e1 (read) + e2 (representative)
		icb_pair& p = pairs[t];
		p.read = r;
		p.rep = rep;
		p.strand = key.strand;
		p.gap_open = e <= 0.01 ? 5 : (e <= 0.04 ? 4 : (e <= 0.1 ? 3 : 2));  // setGapOpen, cluster.cpp:425-440
		double lim = std::floor((1.0 - e) * unsigned(k));		     // getAlnRatio, cluster.cpp:446
		p.limit = lim < -1.0 ? -1 : (lim > double(k + 1) ? k + 1 : int(lim));
	    }
	    std::vector<icb_align_result> res(pairs.size());
	    rc = icb_align_pairs_on(ctx, rs, pairs.data(), int64_t(pairs.size()), k, res.data());
	    if (rc) return rc;
	    stats.pairs_aligned += int64_t(pairs.size());
	    for (size_t t = 0; t < needPairs.size(); t++) {
		aln[needPairs[t].pos].push_back(AlnRes{needPairs[t].id, needPairs[t].strand, res[t].aligned});
		stats.cells_aligned += int64_t(rs->h_offs[pairs[t].read + 1] - rs->h_offs[pairs[t].read]) *
				       int64_t(rs->h_offs[pairs[t].rep + 1] - rs->h_offs[pairs[t].rep]);
	    }
	    stats.ms_align += sw.lap();
	}
    }
    // ---- AddMinimizers for the clusters founded in this chunk (cluster.cpp:178-180)
    std::vector<int32_t> fr, fi;
    int64_t tot = 0;
    for (int i = 0; i < N; i++)
	if (founder_cls[i] >= 0) {
	    fr.push_back(chunk[i]);
	    fi.push_back(founder_cls[i]);
	    tot += nf[chunk[i]];
	}
    if (!fr.empty()) {
	rc = main_idx.reserve(tot);
	if (rc) return rc;
	rc = upload_founders(C0);
	if (rc) return rc;
	memcpy(h_i32, fr.data(), sizeof(int32_t) * fr.size());
	memcpy(h_i32 + chunk_cap, fi.data(), sizeof(int32_t) * fi.size());
	ICB_CUDA(ctx, cudaMemcpyAsync(d_chunk_reads, h_i32, sizeof(int32_t) * fr.size(), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(d_ids, h_i32 + chunk_cap, sizeof(int32_t) * fi.size(), cudaMemcpyHostToDevice, st));
	rc = icb_index_insert(ctx, main_idx, rs, d_chunk_reads, d_ids, int(fr.size()));
	if (rc) return rc;
	main_idx.nodes_upper += tot;
	ICB_CUDA(ctx, ctx->wait());  // the pinned staging block is rewritten by the next chunk
    }
    stats.ms_insert += sw.lap();
    return ICB_OK;
}

// One chunk through the device-resident loop (dloop.cu).  The host enqueues the chunk's kernels — chunk index, pool-mode
// query, then rounds of commit / decide / gather / full lists / alignments — and reads one state block back per batch of
// rounds; nothing of the chunk is decided on the host.  handled = false: the device loop met a case it leaves to the host
// loop (a read with more candidates than its shared-memory plan, a hit map that would rehash, a pool that was too small
// and has been enlarged for the next chunks) and nothing of the chunk has been applied.
int icb_clusterer::process_chunk_device(const std::vector<int32_t>& chunk, const std::vector<int64_t>& slots, int32_t* cluster_of,
					int32_t* strand_of, bool& handled)
{
    IcbRange nvtx_range("chunk: device loop");
    handled = false;
    cudaStream_t st = ctx->stream;
    const int N = int(chunk.size());
    const int C0 = int(founder.size());
    const int Ctot = C0 + N;
    const int k = P.kmer_size;
    Stopwatch sw;
    // ---- capacities that may reallocate (and synchronise) before anything of the chunk is queued
    int rc = ensure_founder_cap(Ctot);
    if (rc) return rc;
    int64_t tot = 0;
    for (int i = 0; i < N; i++) tot += rs->h_nf[chunk[i]];
    rc = main_idx.reserve(tot);
    if (rc) return rc;
    if (!ctx->align_ws) ctx->align_ws = new IcbAlignWorkspace;
    IcbAlignWorkspace& ws = *ctx->align_ws;
    {
	// trace arena of this stream: room for the largest pair of the read set at least, a few hundred average pairs by default
	static const size_t trace_mb = getenv("ICB_DLOOP_TRACE_MB") ? size_t(atol(getenv("ICB_DLOOP_TRACE_MB"))) : 256;
	const size_t L = size_t(std::max<int64_t>(rs->max_len, 64));
	const size_t one = ((L + 7) / 8) * ((L + 71) / 8 + 1) * 32 + 4096;
	const size_t lines_one = (4 * L + 64) * 4;
	rc = ws.ensure(ctx, std::max(trace_mb << 20, one), std::max(lines_one * size_t(N), size_t(4) << 20), int64_t(DL_SPEC) * chunk_cap + 16);
	if (rc) return rc;
    }
    const unsigned long long tiles_cap = ws.trace_bytes / 32 - 4, words_cap = ws.lines_bytes / 4 - 32;
    IcbQueryArgs qa, qf;
    int nb_a, nb_f;
    size_t smem_a, smem_f;
    query_base(this, qa, N, C0);
    rc = plan_query(Ctot, 0, 2, qa, N, nb_a, smem_a);
    if (rc) return rc;
    qa.hdr = V.hdr;
    qa.dl_state = V.st;
    qa.dl_pool = V.cands;
    qa.dl_pool_aln = V.aln;
    qa.dl_depmask = V.depmask;
    qa.dl_mw = V.mw;
    // measured on configs[1] (1 M reads): fewer rounds (6.7 k instead of 7.4 k) but more candidate orders computed ahead of
    // need and more speculative alignments: 395 k instead of 420 k reads/s -> off unless asked for
    static const bool early_full = getenv("ICB_DLOOP_EARLY_FULL") != nullptr;
    if (early_full && P.mode != 1) {
	qa.dl_full_pool = V.full;
	qa.dl_full_off = V.full_off;
	qa.dl_full_n = V.full_n;
    }
    query_base(this, qf, N, C0);
    // the full-list query is launched from the round graph: its launch configuration is planned for a snapshot size with
    // headroom, so that it (and the graph) only changes when the clusterer has outgrown it
    if (Ctot > qf_bound) {
	qf_bound = 4096;
	while (qf_bound < Ctot + Ctot / 4) qf_bound *= 2;
    }
    rc = plan_query(qf_bound, 0, 1, qf, chunk_cap, nb_f, smem_f);  // same scratch block as the pool-mode plan: never used at the same time
    if (rc) return rc;
    if (qa.g_stride > 0) qa.g_scratch = d_scratch;  // the second plan may have enlarged (moved) the scratch block
    qf.g_scratch = d_scratch;
    qf.hdr = d_hdr_full;
    qf.list = V.full_list;
    qf.dl_n_list = &V.st->n_full;
    qf.full_off = V.full_list_off;
    qf.cands = V.full;
    qf.dl_depmask = V.depmask;
    qf.dl_mw = V.mw;
    qf.dl_full_off = V.full_off;
    qf.dl_full_n = V.full_n;
    const int max_m = int(std::min<int64_t>(rs->max_len, 8000));
    static const int commit_iters = getenv("ICB_DLOOP_COMMIT_ITERS") ? std::max(1, atoi(getenv("ICB_DLOOP_COMMIT_ITERS"))) : 2;
    // the view's per-run pointers
    V.chunk_reads = d_chunk_reads;
    V.offs = rs->d_offs;
    V.raw_err = rs->d_err;
    V.hpc_len = rs->d_hpc_len;
    V.nf = rs->d_nf;
    V.nr = rs->d_nr;
    V.bkt = rs->d_bkt;
    V.cls_founder = d_cls_founder;
    V.mode = P.mode;
    V.min_shared = P.min_shared;
    V.k = k;
    V.min_frac = P.min_fraction;
    V.mapped_th = P.mapped_threshold;
    V.aligned_th = P.aligned_threshold;
    V.force_fallback = force_fallback;
    static const int spec_pairs = getenv("ICB_DLOOP_SPEC_PAIRS") ? std::max(1, std::min(DL_SPEC, atoi(getenv("ICB_DLOOP_SPEC_PAIRS")))) : 1;
    V.spec_pairs = spec_pairs;
    static const int spec_window = getenv("ICB_DLOOP_SPEC_WINDOW") ? atoi(getenv("ICB_DLOOP_SPEC_WINDOW")) : 1 << 20;
    V.spec_window = spec_window;
    // ---- chunk index: every read of the chunk as a provisional founder (id = C0 + position)
    for (int i = 0; i < N; i++) {
	h_i32[i] = chunk[i];
	h_i32[chunk_cap + i] = C0 + i;
    }
    ICB_CUDA(ctx, cudaMemcpyAsync(d_chunk_reads, h_i32, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_ids, h_i32 + chunk_cap, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    rc = chunk_idx.clear();
    if (rc) return rc;
    rc = chunk_idx.reserve(tot);
    if (rc) return rc;
    qa.res_tag = qf.res_tag = uint32_t((unsigned long long)(V.descs16) >> 8) & 0xfffffu;
    qa.chunk_idx = qf.chunk_idx = chunk_idx.v;
    qa.main_idx = qf.main_idx = main_idx.v;
    rc = icb_index_insert(ctx, chunk_idx, rs, d_chunk_reads, d_ids, N);
    if (rc) return rc;
    chunk_idx.nodes_upper = tot;
    // ---- candidates of every read, first decisions
    rc = icb_dl_init(ctx, V, N, C0, cand_cap, full_cap, tiles_cap, words_cap);
    if (rc) return rc;
    rc = icb_query_launch(ctx, qa, nb_a, smem_a);
    if (rc) return rc;
    rc = icb_dl_decide(ctx, V, N);
    if (rc) return rc;
    // the full-list query reads its arguments from device memory
    *h_qf_args = qf;
    ICB_CUDA(ctx, cudaMemcpyAsync(d_qf_args, h_qf_args, sizeof(IcbQueryArgs), cudaMemcpyHostToDevice, st));
    const int G = chunk_cap;  // grids of the round's kernels: the chunk capacity (positions past the chunk's end return at once)
    const int GP = chunk_cap * V.spec_pairs;  // capacity of the pair queues = grid of the alignment kernels
    auto enqueue_round = [&]() -> int {
	// commit the decided prefix, re-decide what the new founders invalidated (in parallel), commit on: the walk only
	// stops for good at a read that waits for an alignment or for its full candidate list
	for (int it = 0; it < commit_iters; it++) {
	    int r2 = icb_dl_commit(ctx, V);
	    if (r2) return r2;
	    r2 = icb_dl_decide(ctx, V, G);
	    if (r2) return r2;
	}
	int r2 = icb_dl_gather(ctx, V, GP);
	if (r2) return r2;
	r2 = icb_query_launch_ptr(ctx, d_qf_args, qf.smem_cnt != 0, nb_f, smem_f);
	if (r2) return r2;
	if (P.mode != 1) {
	    r2 = icb_align16_run_dev(ctx, ws, rs->d_seq, V.descs16, &V.st->n_pairs16, GP, max_m, k, V.aln);
	    if (r2) return r2;
	    // pairs the packed kernel cannot take (longer than 8000, or refused for a character outside ACGT)
	    r2 = icb_align32_run_dev(ctx, ws, rs->d_seq, V.descs32, &V.st->n_pairs32, GP, k, V.aln);
	    if (r2) return r2;
	}
	return ICB_OK;
    };
    static const bool no_graph = getenv("ICB_NO_GRAPH") != nullptr;
    const bool use_graph = !no_graph && !ctx->timing;
    if (use_graph) {
	RoundKey key;
	memset(&key, 0, sizeof key);
	const void* ptrs[12] = {V.st, V.cands, V.full, V.cls_founder, ws.d_trace, ws.d_lines, ws.d_ends, ws.d_ticket, rs->d_seq, rs->d_bkt, d_qf_args,
				V.hdr};
	memcpy(key.p, ptrs, sizeof ptrs);
	memcpy(&key.view, &V, sizeof V);
	key.smem_f = smem_f;
	key.nb_f = nb_f;
	key.smem_cnt_f = qf.smem_cnt;
	key.iters = commit_iters;
	key.mode = P.mode;
	key.max_m = max_m;
	key.k = k;
	key.spec = V.spec_window * 8 + V.spec_pairs;
	key.aln_priority = ctx->aln_priority * 2 + (ctx->aln_latency ? 1 : 0);
	if (!round_exec || !(key == round_key)) {
	    if (round_exec) cudaGraphExecDestroy(round_exec);
	    round_exec = nullptr;
	    const int64_t lc0 = ctx->launch_count;
	    ICB_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
	    const int rcap = enqueue_round();
	    cudaGraph_t g = nullptr;
	    const cudaError_t e = cudaStreamEndCapture(st, &g);
	    round_nodes = int(ctx->launch_count - lc0);
	    ctx->launch_count = lc0;
	    if (rcap) {
		if (g) cudaGraphDestroy(g);
		return rcap;
	    }
	    ICB_CUDA(ctx, e);
	    const cudaError_t e2 = cudaGraphInstantiate(&round_exec, g, 0);
	    cudaGraphDestroy(g);
	    ICB_CUDA(ctx, e2);
	    round_key = key;
	}
    }
    int enqueued = 0;
    for (;;) {
	const int R = std::max(2, std::min(16, int(rounds_avg + 1.5) - enqueued));
	for (int q = 0; q < R; q++) {
	    if (use_graph) {
		ICB_CUDA(ctx, cudaGraphLaunch(round_exec, st));
		ctx->launch_count += round_nodes;
	    }
	    else {
		rc = enqueue_round();
		if (rc) return rc;
	    }
	}
	enqueued += R;
	rc = icb_dl_commit(ctx, V);  // a last walk, so that a chunk whose work is complete ends in this batch of rounds
	if (rc) return rc;
	rc = icb_dl_finalize(ctx, V, d_n_ins);
	if (rc) return rc;
	rc = icb_index_insert_dev(ctx, main_idx, rs, V.ins_reads, V.ins_ids, d_n_ins, N);
	if (rc) return rc;
	ICB_CUDA(ctx, cudaMemcpyAsync(h_state, V.st, sizeof(IcbDLoopState), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(h_out, V.out_cls, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(h_out + chunk_cap, V.out_strand, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(h_out + 2 * chunk_cap, V.founder_pos, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, ctx->wait());
	if (h_state->fallback || h_state->done == 2) break;
	if (enqueued > 64 * N + 1024) return ctx->fail(ICB_ERR_STATE, "device loop made no progress");
    }
    stats.ms_align += sw.lap();
    if (h_state->fallback) {
	// handled stays false: nothing was applied (the founders' insert is skipped unless the chunk finished).  A pool that
	// was too small is enlarged and the chunk goes through the device loop once more; anything else is the host loop's.
	dl_retry = false;
	dl_reason = h_state->fallback;
	if (h_state->fallback == DL_FB_POOL && cand_cap < 0x40000000u) {
	    rc = grow_cand_pool(cand_cap * 4);
	    if (rc) return rc;
	    dl_retry = true;
	}
	if (h_state->fallback == DL_FB_FULLPOOL && full_cap < (1ull << 31)) {
	    rc = grow_full_pool(full_cap * 4);
	    if (rc) return rc;
	    dl_retry = true;
	}
	return ICB_OK;
    }
    // ---- apply
    const IcbDLoopState& S = *h_state;
    rounds_avg = 0.75 * rounds_avg + 0.25 * double(S.rounds);
    for (int i = 0; i < N; i++) {
	cluster_of[slots[i]] = h_out[i];
	strand_of[slots[i]] = h_out[chunk_cap + i];
    }
    int64_t ftot = 0;
    for (int i = 0; i < S.n_founders; i++) {
	const int32_t r = chunk[h_out[2 * chunk_cap + i]];
	founder.push_back(r);
	ftot += rs->h_nf[r];
    }
    main_idx.nodes_upper += ftot;
    stats.chunks++;
    stats.dl_chunks++;
    stats.rounds += S.rounds;
    stats.reads_queried += N;
    stats.aln_invoked += S.aln_invoked;
    stats.aln_calls_ref += S.aln_calls_ref;
    stats.pairs_aligned += S.pairs_aligned;
    stats.cells_aligned += S.cells_aligned;
    stats.tie_fallbacks += S.tie_fallbacks;
    stats.decisions += S.decisions;
    stats.order_calls += S.order_calls;
    stats.hits += S.hits;
    stats.cands += S.cands;
    stats.dl_stop_full += S.stop_full;
    stats.dl_stop_pair_first += S.stop_pair_first;
    stats.dl_stop_pair_again += S.stop_pair_again;
    handled = true;
    return ICB_OK;
}

// per-read scalars the greedy loop reads on the host + the device error classes: once per read set
int icb_readset_prepare_host(icb_readset* rs)
{
    icb_ctx* ctx = rs->ctx;
    if (rs->host_ready) return ICB_OK;
    const int64_t n = rs->n_reads;
    const int64_t cap = std::max(rs->n_reads, rs->cap_reads);  // room for appended reads (icb_readset_append_from)
    cudaStream_t st = ctx->stream;
    rs->h_status.assign(cap, 0);
    rs->h_hpc_len.assign(cap, 0);
    rs->h_nf.assign(cap, 0);
    rs->h_nr.assign(cap, 0);
    rs->h_score.assign(cap, 0.0);
    rs->h_err.assign(cap, 0.0);
    if (n) {
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_status.data(), rs->d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_hpc_len.data(), rs->d_hpc_len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_nf.data(), rs->d_nf, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_nr.data(), rs->d_nr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_score.data(), rs->d_score, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->h_err.data(), rs->d_err, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    if (!rs->d_class) ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&rs->d_class, cap + 16));
    if (!rs->d_bkt) ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&rs->d_bkt, sizeof(uint32_t) * (cap + 16)));
    int rc0 = icb_classify_errors(ctx, rs->d_hpc_err, rs->d_status, n, rs->d_class);
    if (rc0) return rc0;
    ICB_CUDA(ctx, ctx->wait());
    rs->max_m = 0;
    for (int64_t i = 0; i < n; i++) rs->max_m = std::max(rs->max_m, std::max(rs->h_nf[i], rs->h_nr[i]));
    if (cap > n && rs->sk_w >= rs->sk_k) {
	// appended reads are as long as max_len at most: leave their sketches the expected density and a third on top (the
	// clusterers size their bitmaps and tables by max_m; icb_readset_append_from refuses a sketch above it)
	const int64_t est = rs->max_len * 2 / (rs->sk_w - rs->sk_k + 2) * 4 / 3 + 64;
	rs->max_m = int(std::min<int64_t>(32767, std::max<int64_t>(rs->max_m, est)));
    }
    // bucket count of the reference's hit map per read: unordered_map(20 * (|mins| + |revMins|)) (minimizer.cpp:48) asks the
    // HOST libstdc++ for the next prime; the device loop's candidate-order emulation reads it from this table
    {
	rs->h_bkt.assign(size_t(cap), 0);
	std::__detail::_Prime_rehash_policy pol;
	size_t last_hint = ~size_t(0), last_bkt = 0;
	for (int64_t i = 0; i < n; i++) {
	    const size_t hint = 20 * (size_t(rs->h_nf[i]) + size_t(rs->h_nr[i]));
	    if (hint != last_hint) {
		last_hint = hint;
		last_bkt = pol._M_next_bkt(hint);
	    }
	    rs->h_bkt[size_t(i)] = uint32_t(std::min<size_t>(last_bkt, 0xffffffffu));
	}
	if (n) ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_bkt, rs->h_bkt.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, ctx->wait());
    }
    rs->host_ready = true;
    return ICB_OK;
}

// The stored error rates of a loaded batch instead of the ones the quality strings give (see the header)
extern "C" int icb_readset_set_errors(icb_readset* rs, const double* raw_err, const double* raw_score, const double* hpc_err)
{
    if (!rs) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!rs->scored || !rs->sketched) return ctx->fail(ICB_ERR_STATE, "set_errors needs scored and sketched reads");
    const size_t bytes = sizeof(double) * size_t(rs->n_reads);
    if (bytes) {
	if (raw_err) ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_err, raw_err, bytes, cudaMemcpyHostToDevice, ctx->stream));
	if (raw_score) ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_score, raw_score, bytes, cudaMemcpyHostToDevice, ctx->stream));
	if (hpc_err) ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_hpc_err, hpc_err, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    ICB_CUDA(ctx, ctx->wait());
    rs->host_ready = false;  // host mirrors and error classes are rebuilt by the next clusterer
    return ICB_OK;
}

// A sketched read of another read set becomes read n_reads of this one (see the header).  Everything the clusterers read
// about a read is filled in: raw and compressed bases, both sketches (appended to the arena), the per-read scalars with the
// caller's error rates, the error class, the hit map's bucket count, the host mirrors.
extern "C" int icb_readset_append_from(icb_readset* rs, icb_readset* src, int64_t read, double raw_err, double raw_score, double hpc_err,
				       int64_t* index)
{
    if (!rs || !src || !index || rs == src) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    try {
	ICB_CUDA(ctx, cudaSetDevice(ctx->device));
	if (src->ctx != ctx) return ctx->fail(ICB_ERR_ARG, "append: both read sets must belong to one context");
	if (!rs->sketched || !src->sketched || rs->sk_k != src->sk_k || rs->sk_w != src->sk_w)
	    return ctx->fail(ICB_ERR_STATE, "append: both read sets must be sketched with the same k and w");
	if (read < 0 || read >= src->n_reads) return ctx->fail(ICB_ERR_ARG, "append: no such read in the source read set");
	int rc = icb_readset_prepare_host(rs);
	if (rc) return rc;
	cudaStream_t st = ctx->stream;
	// the source read's scalars
	int32_t sc[4] = {0, 0, 0, 0};  // status, hpc_len, nf, nr
	int64_t smo = 0;
	ICB_CUDA(ctx, cudaMemcpyAsync(&sc[0], src->d_status + read, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(&sc[1], src->d_hpc_len + read, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(&sc[2], src->d_nf + read, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(&sc[3], src->d_nr + read, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(&smo, src->d_min_off + read, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, ctx->wait());
	if (sc[0] != 1) return ctx->fail(ICB_ERR_STATE, "append: the source read has no sketch (too short after compression?)");
	const int64_t so = src->h_offs[size_t(read)], L = src->h_offs[size_t(read) + 1] - so;
	const int64_t idx = rs->n_reads, base = rs->h_offs[size_t(idx)];
	const int64_t nm = int64_t(sc[2]) + int64_t(sc[3]);
	if (idx >= rs->cap_reads || base + L > rs->cap_bases)
	    return ctx->fail(ICB_ERR_NOMEM, "append: the read set's reserve is used up (icb_readset_create_reserve)");
	if (rs->arena_used + nm > rs->mins_capacity) return ctx->fail(ICB_ERR_NOMEM, "append: the minimizer arena's reserve is used up");
	if (L > rs->max_len) return ctx->fail(ICB_ERR_ARG, "append: the read is longer than the reserve's extra_max_len");
	if (sc[2] > rs->max_m || sc[3] > rs->max_m) return ctx->fail(ICB_ERR_ARG, "append: the read has more minimizers than the read set was prepared for");
	// bases and sketches, device to device
	if (L > 0) {
	    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_seq + base, src->d_seq + so, size_t(L), cudaMemcpyDeviceToDevice, st));
	    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_qual + base, src->d_qual + so, size_t(L), cudaMemcpyDeviceToDevice, st));
	    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_hpc_seq + base, src->d_hpc_seq + so, size_t(sc[1]), cudaMemcpyDeviceToDevice, st));
	    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_hpc_qual + base, src->d_hpc_qual + so, size_t(sc[1]), cudaMemcpyDeviceToDevice, st));
	}
	if (nm > 0)
	    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_mins + rs->arena_used, src->d_mins + smo, sizeof(icb_minimizer) * size_t(nm),
					  cudaMemcpyDeviceToDevice, st));
	// per-read scalars (pageable sources: each copy is staged before the call returns)
	const int64_t end = base + L, mo = rs->arena_used;
	const int32_t one = 1;
	std::__detail::_Prime_rehash_policy pol;
	const uint32_t bkt = uint32_t(std::min<size_t>(pol._M_next_bkt(20 * (size_t(sc[2]) + size_t(sc[3]))), 0xffffffffu));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_offs + idx + 1, &end, sizeof(int64_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_status + idx, &one, sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_hpc_len + idx, &sc[1], sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_nf + idx, &sc[2], sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_nr + idx, &sc[3], sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_min_off + idx, &mo, sizeof(int64_t), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_hpc_err + idx, &hpc_err, sizeof(double), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_err + idx, &raw_err, sizeof(double), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_score + idx, &raw_score, sizeof(double), cudaMemcpyHostToDevice, st));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_bkt + idx, &bkt, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
	rc = icb_classify_errors(ctx, rs->d_hpc_err + idx, rs->d_status + idx, 1, rs->d_class + idx);
	if (rc) return rc;
	ICB_CUDA(ctx, ctx->wait());
	// host side
	rs->h_offs.push_back(end);
	rs->h_status[size_t(idx)] = 1;
	rs->h_hpc_len[size_t(idx)] = sc[1];
	rs->h_nf[size_t(idx)] = sc[2];
	rs->h_nr[size_t(idx)] = sc[3];
	rs->h_score[size_t(idx)] = raw_score;
	rs->h_err[size_t(idx)] = raw_err;
	rs->h_bkt[size_t(idx)] = bkt;
	rs->n_reads = idx + 1;
	rs->n_bases = end;
	rs->arena_used += nm;
	rs->total_f = rs->total_r = -1;
	*index = idx;
	return ICB_OK;
    }
    catch (const std::bad_alloc&) {
	return ctx->fail(ICB_ERR_NOMEM, "append: out of host memory");
    }
    catch (...) {
	return ctx->fail(ICB_ERR_STATE, "append: unexpected exception");
    }
}

// The candidate-order emulation (ref_order.h, ref_order_hd.h and their device form in dloop.cu) restates what THIS
// libstdc++ does (bucket counts, list order of _Hashtable, introsort).  Once per process it is checked against the real
// std::unordered_map + std::sort on random candidate lists; a library that behaves differently makes every clusterer
// creation fail loudly instead of clustering differently from the reference built with it.
static bool icb_ref_order_selfcheck()
{
    static const bool ok = [] {
	struct C4 {
	    int32_t id, strand, count, tmapped;
	};
	uint64_t s = 0x243F6A8885A308D3ull;
	auto rnd = [&]() {
	    s ^= s << 13;
	    s ^= s >> 7;
	    s ^= s << 17;
	    return uint32_t(s >> 11);
	};
	IcbRefOrderScratch S;
	std::vector<IcbSortedCand> grp;
	for (int it = 0; it < 24; it++) {
	    const int ncls = 40 + int(rnd() % 1500), nf = 50 + int(rnd() % 500), nr = 50 + int(rnd() % 500);
	    const size_t hint = size_t(20) * size_t(nf + nr);
	    std::vector<C4> v;
	    for (int id = 0; id < ncls; id++)
		for (int sd = 0; sd < 2; sd++)
		    if (rnd() % 3) v.push_back(C4{id, sd ? -1 : 1, 1 + int(rnd() % 5), int(rnd() % (sd ? nr : nf))});
	    std::vector<C4> ins(v);
	    std::stable_sort(ins.begin(), ins.end(), [](const C4& a, const C4& b) {
		if (a.strand != b.strand) return a.strand > b.strand;
		return a.tmapped < b.tmapped;
	    });
	    std::unordered_map<std::pair<int, int>, unsigned, IcbPairHash> m(hint);
	    for (auto& c : ins) m[std::make_pair(c.id, c.strand)] = unsigned(c.count);
	    std::vector<IcbSortedCand> want;
	    std::vector<icb_sort_hd::Key> keys;
	    for (auto& kv : m) {
		keys.push_back(icb_sort_hd::Key(kv.second) << 32 | icb_sort_hd::Key(want.size()));
		want.push_back(IcbSortedCand{kv.second, unsigned(kv.first.first), kv.first.second});
	    }
	    const std::vector<IcbSortedCand> iter_order(want);
	    std::sort(want.begin(), want.end(), [](const IcbSortedCand& a, const IcbSortedCand& b) { return a.Size > b.Size; });
	    const unsigned target = want.empty() ? 0 : want[0].Size;
	    icb_reference_order_group(v.data(), int(v.size()), hint, target, S, grp);
	    icb_sort_hd::introsort_group(keys.data(), long(keys.size()), target);
	    size_t q = 0, q2 = 0;
	    for (auto& w : want)
		if (w.Size == target) {
		    if (q >= grp.size() || grp[q].Cls != w.Cls || grp[q].Strand != w.Strand) return false;
		    q++;
		    while (q2 < keys.size() && uint32_t(keys[q2] >> 32) != target) q2++;
		    if (q2 >= keys.size()) return false;
		    const IcbSortedCand& e = iter_order[uint32_t(keys[q2])];
		    if (e.Cls != w.Cls || e.Strand != w.Strand) return false;
		    q2++;
		}
	    if (q != grp.size()) return false;
	}
	return true;
    }();
    return ok;
}

// parameters of a clustering run -> clusterer (validated; the n_max table is evaluated with the HOST libm)
static int clusterer_configure(icb_ctx* ctx, icb_clusterer* cl, icb_readset* rs, const icb_cluster_params* p)
{
    if (!rs->scored || !rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_clusterer_create needs scored and sketched reads");
    if (p->kmer_size != rs->sk_k || p->window_size != rs->sk_w)
	return ctx->fail(ICB_ERR_ARG, "k / w differ from the ones the reads were sketched with");
    if (p->mode < 0 || p->mode > 2) return ctx->fail(ICB_ERR_ARG, "mode must be 0 (sahlin), 1 (fast) or 2 (furious)");
    if (!rs->host_ready) return ctx->fail(ICB_ERR_STATE, "read set host mirror missing");
    if (rs->max_m >= 32768) return ctx->fail(ICB_ERR_ARG, "a read has 32768 or more minimizers per strand");
    if (!icb_ref_order_selfcheck())
	return ctx->fail(ICB_ERR_STATE, "the C++ library's unordered_map / std::sort do not behave like the ones the candidate-order "
					"emulation restates (libstdc++): results would differ from a reference built with this library");
    cl->ctx = ctx;
    cl->rs = rs;
    cl->P = *p;
    if (cl->P.chunk_size <= 0) cl->P.chunk_size = ICB_DEFAULT_CHUNK;
    // n_max table (cluster.cpp:335-349: pow(pError, n) >= minProbNoHits with the HOST libm; monotone in n)
    if (p->mode != 2) {
	for (int c = 0; c < 225; c++) {
	    double ps = p->p_shared[c];
	    if (!(ps == ps)) return ctx->fail(ICB_ERR_NO_PEMP, "Empirical probability lookup failure!");
	    double pe = 1.0 - ps;
	    int32_t n = -1;
	    // a run of missed minimizers is never longer than a sketch: the scan stops there (with pError = 1, or
	    // minProbNoHits <= 0, every run is bridged and the scan would otherwise walk to its cap for all 225 classes)
	    const int32_t t_max = rs->max_m + 1;
	    for (int32_t t = 0; t <= t_max; t++) {
		if (std::pow(pe, double(t)) >= p->min_prob_no_hits) n = t;
		else break;
	    }
	    cl->nmax[c] = n;
	}
    }
    else
	for (int c = 0; c < 225; c++) cl->nmax[c] = 0;
    return ICB_OK;
}

// a clusterer on an explicit context (the read set's own, or a worker's child context with its own stream);
// errors are reported on `ctx`
static int clusterer_create_on(icb_ctx* ctx, icb_readset* rs, const icb_cluster_params* p, icb_clusterer** out)
{
    *out = nullptr;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<icb_clusterer> cl(new icb_clusterer);
    int rc = clusterer_configure(ctx, cl.get(), rs, p);
    if (rc) return rc;
    cl->chunk_cap = cl->P.chunk_size;
    rc = cl->init_buffers();
    if (rc) {
	cl->free_buffers();
	return rc;
    }
    *out = cl.release();
    return ICB_OK;
}

// a pooled worker's clusterer for another run: new parameters / read set, same buffers
static int clusterer_rebind(icb_clusterer* cl, icb_readset* rs, const icb_cluster_params* p)
{
    icb_ctx* ctx = cl->ctx;
    int rc = clusterer_configure(ctx, cl, rs, p);
    if (rc) return rc;
    memcpy(cl->h_i32 + 3 * cl->chunk_cap, cl->nmax, sizeof(int32_t) * 225);
    ICB_CUDA(ctx, cudaMemcpyAsync(cl->d_nmax, cl->h_i32 + 3 * cl->chunk_cap, sizeof(int32_t) * 225, cudaMemcpyHostToDevice, ctx->stream));
    return cl->reset();
}

// sums of the additive counters and timers (ms_batch_max / ms_wall are set by the caller)
static void add_stats(icb_cluster_stats& t, const icb_cluster_stats& s)
{
    t.chunks += s.chunks; t.rounds += s.rounds; t.reads_queried += s.reads_queried; t.aln_invoked += s.aln_invoked;
    t.pairs_aligned += s.pairs_aligned; t.cells_aligned += s.cells_aligned; t.tie_fallbacks += s.tie_fallbacks;
    t.decisions += s.decisions; t.ms_query += s.ms_query; t.ms_full += s.ms_full; t.ms_align += s.ms_align;
    t.ms_insert += s.ms_insert; t.ms_host += s.ms_host; t.ms_cpu += s.ms_cpu; t.ms_order += s.ms_order; t.order_calls += s.order_calls;
    t.aln_calls_ref += s.aln_calls_ref; t.hyp_decisions += s.hyp_decisions; t.hits += s.hits; t.cands += s.cands;
    t.dl_chunks += s.dl_chunks; t.dl_fallback_chunks += s.dl_fallback_chunks; t.dl_fallback_reasons |= s.dl_fallback_reasons;
    t.dl_stop_full += s.dl_stop_full; t.dl_stop_pair_first += s.dl_stop_pair_first; t.dl_stop_pair_again += s.dl_stop_pair_again;
}

// what icb_cluster_tree_traced keeps for a caller that rebuilds the batches of every node (the CLI's `tree`)
struct icb_tree_trace {
    struct Job {
	int32_t left, right, level;
	std::vector<int32_t> cluster_of, strand_of;  // per cluster of the right node: -1 = dropped (MinClsSize), else its cluster in the merged node
    };
    std::vector<int32_t> leaf_cluster, leaf_strand;  // per sorted slot: the leaf's own decision (batch-local ids)
    std::vector<Job> jobs;
};

// batches (and merges of one tree level) in flight per GPU
static int default_workers()
{
    static const int v = getenv("ICB_WORKERS") ? std::max(1, atoi(getenv("ICB_WORKERS"))) : 48;
    return v;
}

// launch priority of the alignment kernels of sorted batch b of n (see cluster_batches_impl)
static int batch_aln_priority(const icb_ctx* c, int b, int n_batches)
{
    static const bool on = getenv("ICB_NO_BATCH_PRIORITY") == nullptr;
    if (!on) return c->lo_priority;
    return std::min(c->lo_priority, b == 0 ? c->hi_priority + 1 : (b <= n_batches / 8 ? c->hi_priority + 2 : c->lo_priority));
}

// how many of the first sorted batches take the latency-first alignment kernel (their chains of rounds are the longest)
static int batch_latency_first()
{
    static const int v = getenv("ICB_LATENCY_BATCHES") ? atoi(getenv("ICB_LATENCY_BATCHES")) : 1;
    return v;
}

void icb_worker_pool_release(icb_ctx* ctx)
{
    for (size_t w = 0; w < ctx->pool_ctx.size(); w++) {
	if (ctx->pool_cl[w]) icb_clusterer_destroy(ctx->pool_cl[w]);
	icb_ctx_destroy_child(ctx, ctx->pool_ctx[w]);
    }
    ctx->pool_ctx.clear();
    ctx->pool_cl.clear();
}

extern "C" {

int icb_clusterer_create(icb_readset* rs, const icb_cluster_params* p, icb_clusterer** out)
{
    if (!rs || !p || !out) return ICB_ERR_ARG;
    *out = nullptr;
    icb_ctx* ctx = rs->ctx;
    if (!rs->scored || !rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_clusterer_create needs scored and sketched reads");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = icb_readset_prepare_host(rs);
    if (rc) return rc;
    return clusterer_create_on(ctx, rs, p, out);
}

void icb_clusterer_destroy(icb_clusterer* cl)
{
    if (!cl) return;
    cudaSetDevice(cl->ctx->device);
    cudaStreamSynchronize(cl->ctx->stream);
    cl->free_buffers();
    delete cl;
}

int icb_clusterer_seed(icb_clusterer* cl, const int32_t* rep_reads, int64_t n)
{
    if (!cl || n < 0 || (n > 0 && !rep_reads)) return ICB_ERR_ARG;
    ICB_CUDA(cl->ctx, cudaSetDevice(cl->ctx->device));
    return cl->seed(rep_reads, n);
}

int icb_clusterer_update_representative(icb_clusterer* cl, int32_t cluster, int32_t rep_read)
{
    if (!cl) return ICB_ERR_ARG;
    ICB_CUDA(cl->ctx, cudaSetDevice(cl->ctx->device));
    return cl->update_representative(cluster, rep_read);
}

static int cluster_batches_impl(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds,
			int32_t n_batches, int32_t n_workers, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
			icb_cluster_stats* stats)
{
    IcbRange nvtx_range("cluster of every batch (icb_cluster_batches)");
    if (!rs || !params || n_batches < 0 || (n_batches > 0 && (!order || !bounds || !cluster_of || !strand_of))) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) *stats = icb_cluster_stats{};
    if (n_batches == 0) return ICB_OK;
    Stopwatch wall;
    if (!rs->scored || !rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_cluster_batches needs scored and sketched reads");
    for (int b = 0; b < n_batches; b++)
	if (bounds[b] < 0 || bounds[b + 1] < bounds[b]) return ctx->fail(ICB_ERR_ARG, "batch bounds must be non-decreasing");
    int rc = icb_readset_prepare_host(rs);
    if (rc) return rc;
    ICB_CUDA(ctx, ctx->wait());  // scores / sketches are complete before other streams read them
    // A worker sleeps on its stream (icb_ctx::wait) but for 0.6 core-seconds per million reads: how many batches are kept
    // in flight is a matter of filling the GPU, not of host cores (round 1 sized it by the cores, 3 per core, and a rank with
    // 4 of a node's 32 cores then ran 12 of its 35 batches at a time).  ICB_WORKERS overrides.
    int nw = n_workers > 0 ? n_workers : default_workers();
    nw = std::min(nw, int(n_batches));
    std::atomic<int> next(0), first_err(0);
    std::vector<icb_cluster_stats> wstats(nw);
    std::vector<std::string> werr(nw);
    // workers (child context = stream + alignment workspace, clusterer = index and staging buffers) are kept by the parent
    // context between calls: nothing is allocated or freed while other workers' kernels are in flight
    while (int(ctx->pool_ctx.size()) < nw) {
	icb_ctx* kid = icb_ctx_create_child(ctx);
	if (!kid) return ctx->fail(ICB_ERR_CUDA, "cannot create a worker stream");
	ctx->pool_ctx.push_back(kid);
	ctx->pool_cl.push_back(nullptr);
    }
    const int want_chunk = params->chunk_size > 0 ? params->chunk_size : ICB_DEFAULT_CHUNK;
    for (int w = 0; w < nw; w++) {
	icb_ctx* c = ctx->pool_ctx[w];
	c->timing = ctx->timing;
	c->trace_base = ctx->trace_base;
	if (ctx->pool_cl[w] && ctx->pool_cl[w]->chunk_cap != want_chunk) {
	    icb_clusterer_destroy(ctx->pool_cl[w]);
	    ctx->pool_cl[w] = nullptr;
	}
	int rc = ctx->pool_cl[w] ? clusterer_rebind(ctx->pool_cl[w], rs, params) : clusterer_create_on(c, rs, params, &ctx->pool_cl[w]);
	if (rc) return ctx->fail(rc, c->last_error);
    }
    static const bool batch_log = getenv("ICB_BATCH_LOG") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto work = [&](int w) {
	icb_ctx* c = ctx->pool_ctx[w];
	cudaSetDevice(c->device);
	icb_clusterer* cl = ctx->pool_cl[w];
	int rc = ICB_OK;
	while (!rc) {
	    const int b = next.fetch_add(1);
	    if (b >= n_batches || first_err.load()) break;
	    Stopwatch bw;
	    rc = cl->reset();
	    if (rc) break;
	    // The batches of a sort hold its reads best first, and the better reads found more clusters per read: the first
	    // batches need the most rounds (batch 0 of configs[1]: 405, the median batch 220).  With every batch in flight
	    // at once the step ends when the longest chain does, so its alignment kernels go ahead of the others'.
	    c->aln_priority = batch_aln_priority(c, b, n_batches);
	    c->aln_latency = b < batch_latency_first();
	    rc = icb_cluster_reads(cl, order + bounds[b], bounds[b + 1] - bounds[b], cluster_of + bounds[b], strand_of + bounds[b]);
	    if (rc) break;
	    const double bms = bw.lap();
	    wstats[w].ms_batch_max = std::max(wstats[w].ms_batch_max, bms);
	    if (batch_log) {
		const icb_cluster_stats& q = cl->stats;
		fprintf(stderr, "[icb] batch %d worker %d reads %lld: ends at %.0f, %.0f ms (query %.0f full %.0f align %.0f insert %.0f host %.0f order %.0f cpu %.0f) "
				"chunks %lld rounds %lld pairs %lld ties %lld decisions %lld order_calls %lld clusters %zu\n",
			b, w, (long long)(bounds[b + 1] - bounds[b]),
			std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), bms, q.ms_query, q.ms_full, q.ms_align, q.ms_insert, q.ms_host, q.ms_order,
			q.ms_cpu, (long long)q.chunks, (long long)q.rounds, (long long)q.pairs_aligned, (long long)q.tie_fallbacks,
			(long long)q.decisions, (long long)q.order_calls, cl->founder.size());
	    }
	    if (n_clusters) n_clusters[b] = int64_t(cl->founder.size());
	    add_stats(wstats[w], cl->stats);
	}
	if (rc) {
	    int z = 0;
	    first_err.compare_exchange_strong(z, rc);
	    werr[w] = c->last_error;
	}
    };
    if (nw == 1) work(0);
    else {
	std::vector<std::thread> th;
	for (int w = 0; w < nw; w++) th.emplace_back(work, w);
	for (auto& t : th) t.join();
    }
    for (int w = 0; w < nw; w++) {
	if (stats) {
	    const icb_cluster_stats& s = wstats[w];
	    add_stats(*stats, s);
	    stats->ms_batch_max = std::max(stats->ms_batch_max, s.ms_batch_max);
	}
	icb_ctx_fold_child(ctx, ctx->pool_ctx[w]);  // the worker's kernel timers and launch count -> parent
    }
    if (stats) stats->ms_wall = wall.lap();
    if (first_err.load()) icb_worker_pool_release(ctx);  // a failed worker's buffers are in an unknown state
    if (first_err.load()) {
	for (auto& e : werr)
	    if (!e.empty()) return ctx->fail(first_err.load(), e);
	return ctx->fail(first_err.load(), "worker failed");
    }
    return ICB_OK;
}

// configs[4] in process: leaf clustering of every sorted batch, then the merge tree the reference's pipeline runs as
// `cluster -l A -r B` jobs over .cer files (README.md:105-117, main.cpp:238-382, cluster.cpp:67-106), on RESIDENT
// clusterers: a node's index, representatives and cluster ids stay on the device between levels, and a merge is
// icb_cluster_reads on the left node's clusterer over the right node's representatives in the right node's cluster order
// (what cluster.cpp:108-261 does with rightBatch->Cls[i][REP]).  Nothing is re-uploaded, re-sketched or re-indexed.
static int cluster_tree_impl(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds, int32_t n_batches,
		     int32_t n_workers, int32_t min_cls_size, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
		     icb_cluster_stats* stats, icb_tree_trace* trace)
{
    IcbRange nvtx_range("batches + merge tree (icb_cluster_tree)");
    if (!rs || !params || n_batches < 0 || (n_batches > 0 && (!order || !bounds || !cluster_of || !strand_of))) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (stats) *stats = icb_cluster_stats{};
    if (n_clusters) *n_clusters = 0;
    if (n_batches == 0) return ICB_OK;
    Stopwatch wall;
    if (!rs->scored || !rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_cluster_tree needs scored and sketched reads");
    for (int b = 0; b < n_batches; b++)
	if (bounds[b] < 0 || bounds[b + 1] < bounds[b]) return ctx->fail(ICB_ERR_ARG, "batch bounds must be non-decreasing");
    int rc = icb_readset_prepare_host(rs);
    if (rc) return rc;
    ICB_CUDA(ctx, ctx->wait());
    int nw = n_workers > 0 ? n_workers : default_workers();
    nw = std::min(nw, int(n_batches));
    while (int(ctx->pool_ctx.size()) < nw) {
	icb_ctx* kid = icb_ctx_create_child(ctx);
	if (!kid) return ctx->fail(ICB_ERR_CUDA, "cannot create a worker stream");
	ctx->pool_ctx.push_back(kid);
	ctx->pool_cl.push_back(nullptr);
    }
    for (int w = 0; w < nw; w++) {
	ctx->pool_ctx[w]->timing = ctx->timing;
	ctx->pool_ctx[w]->trace_base = ctx->trace_base;
    }
    // one node per batch: its clusterer (left operands keep it), its representatives, the worker whose stream it lives on
    struct Node {
	icb_clusterer* cl = nullptr;
	std::vector<int32_t> reps;
	std::vector<int32_t> csize;  // reads per cluster (the reference's Cls[i]->size() - 1: the list holds the representative's copy too)
	int worker = -1, lo = 0, hi = 0, depth = 0;
    };
    std::vector<Node> nodes(static_cast<size_t>(n_batches));
    // local cluster id of a leaf -> (cluster of the node that holds the leaf now, strand sign)
    std::vector<std::vector<int32_t>> lmap(static_cast<size_t>(n_batches)), lsign(static_cast<size_t>(n_batches));
    std::atomic<int> first_err(0);
    std::vector<std::string> werr(static_cast<size_t>(nw));
    std::vector<icb_cluster_stats> wstats(static_cast<size_t>(nw));
    auto fail_worker = [&](int w, int rcw, icb_ctx* c) {
	int z = 0;
	first_err.compare_exchange_strong(z, rcw);
	werr[size_t(w)] = c->last_error;
    };
    // a node that is only ever a RIGHT operand of the plan below keeps nothing but its representatives
    std::vector<char> is_left(static_cast<size_t>(n_batches), 0);
    {
	std::vector<int> cur(static_cast<size_t>(n_batches));
	for (int b = 0; b < n_batches; b++) cur[size_t(b)] = b;
	while (cur.size() > 1) {
	    std::vector<int> nxt;
	    for (size_t i = 0; i + 1 < cur.size(); i += 2) {
		is_left[size_t(cur[i])] = 1;
		nxt.push_back(cur[i]);
	    }
	    if (cur.size() % 2) nxt.push_back(cur.back());
	    cur.swap(nxt);
	}
	is_left[0] = 1;	 // the root keeps its clusterer until the counts are read
    }
    if (trace) {
	trace->leaf_cluster.assign(size_t(bounds[n_batches]), -1);
	trace->leaf_strand.assign(size_t(bounds[n_batches]), 0);
	trace->jobs.clear();
    }
    std::mutex trace_mutex;
    // ---- leaves
    {
	std::atomic<int> next(0);
	auto work = [&](int w) {
	    icb_ctx* c = ctx->pool_ctx[size_t(w)];
	    cudaSetDevice(c->device);
	    for (;;) {
		const int b = next.fetch_add(1);
		if (b >= n_batches || first_err.load()) break;
		Node& nd = nodes[size_t(b)];
		nd.worker = w;
		nd.lo = nd.hi = b;
		c->aln_priority = batch_aln_priority(c, b, n_batches);
		c->aln_latency = b < batch_latency_first();
		int rcw = clusterer_create_on(c, rs, params, &nd.cl);
		if (!rcw)
		    rcw = icb_cluster_reads(nd.cl, order + bounds[b], bounds[b + 1] - bounds[b], cluster_of + bounds[b], strand_of + bounds[b]);
		if (rcw) {
		    fail_worker(w, rcw, c);
		    break;
		}
		nd.reps = nd.cl->founder;
		add_stats(wstats[size_t(w)], nd.cl->stats);
		nd.cl->stats = icb_cluster_stats{};
		lmap[size_t(b)].resize(nd.reps.size());
		lsign[size_t(b)].assign(nd.reps.size(), 1);
		for (size_t q = 0; q < nd.reps.size(); q++) lmap[size_t(b)][q] = int32_t(q);
		nd.csize.assign(nd.reps.size(), 0);
		for (int64_t t = bounds[b]; t < bounds[b + 1]; t++)
		    if (cluster_of[t] >= 0) nd.csize[size_t(cluster_of[t])]++;
		if (trace) {  // the leaf's own decisions (batch-local cluster ids), before the merges rename them
		    std::copy(cluster_of + bounds[b], cluster_of + bounds[b + 1], trace->leaf_cluster.begin() + bounds[b]);
		    std::copy(strand_of + bounds[b], strand_of + bounds[b + 1], trace->leaf_strand.begin() + bounds[b]);
		}
		if (!is_left[size_t(b)]) {
		    icb_clusterer_destroy(nd.cl);
		    nd.cl = nullptr;
		}
	    }
	};
	std::vector<std::thread> th;
	for (int w = 0; w < nw; w++) th.emplace_back(work, w);
	for (auto& t : th) t.join();
    }
    // ---- merge levels: adjacent nodes are paired left to right, an odd last node is carried up (it only ever becomes a
    // right operand: the reference requires the left batch not to be shallower than the right one, cluster.cpp:80-90)
    std::vector<int> cur(static_cast<size_t>(n_batches));
    for (int b = 0; b < n_batches; b++) cur[size_t(b)] = b;
    int64_t merges = 0;
    int level_no = 0;
    while (cur.size() > 1 && !first_err.load()) {
	level_no++;
	struct Job {
	    int left, right;
	};
	std::vector<Job> jobs;
	std::vector<int> nxt;
	for (size_t i = 0; i + 1 < cur.size(); i += 2) {
	    jobs.push_back(Job{cur[i], cur[i + 1]});
	    nxt.push_back(cur[i]);
	}
	if (cur.size() % 2) nxt.push_back(cur.back());
	// a job runs on the stream its left clusterer lives on
	std::vector<std::vector<int>> per_worker(static_cast<size_t>(nw));
	for (size_t j = 0; j < jobs.size(); j++) per_worker[size_t(nodes[size_t(jobs[j].left)].worker)].push_back(int(j));
	auto work = [&](int w) {
	    icb_ctx* c = ctx->pool_ctx[size_t(w)];
	    cudaSetDevice(c->device);
	    for (int j : per_worker[size_t(w)]) {
		if (first_err.load()) break;
		Node& L = nodes[size_t(jobs[size_t(j)].left)];
		Node& R = nodes[size_t(jobs[size_t(j)].right)];
		const int64_t nr = int64_t(R.reps.size());
		// clusters of a MERGED right batch with fewer than MinClsSize reads are left out of the merge — and of the output
		// (cluster.cpp:120-124)
		std::vector<int32_t> el_reps, el_idx;
		for (int64_t q = 0; q < nr; q++)
		    if (!(R.depth > 0 && min_cls_size > 1 && R.csize[size_t(q)] < min_cls_size)) {
			el_reps.push_back(R.reps[size_t(q)]);
			el_idx.push_back(int32_t(q));
		    }
		const int64_t ne = int64_t(el_reps.size());
		std::vector<int32_t> ec(static_cast<size_t>(ne)), es(static_cast<size_t>(ne));
		c->aln_priority = c->lo_priority;
		c->aln_latency = false;
		const int rcw = icb_cluster_reads(L.cl, el_reps.data(), ne, ec.data(), es.data());
		if (rcw) {
		    fail_worker(w, rcw, c);
		    break;
		}
		add_stats(wstats[size_t(w)], L.cl->stats);
		L.cl->stats = icb_cluster_stats{};
		std::vector<int32_t> tc(static_cast<size_t>(nr), -1), ts(static_cast<size_t>(nr), 0);
		L.csize.resize(L.cl->founder.size(), 0);
		for (int64_t q = 0; q < ne; q++) {
		    tc[size_t(el_idx[size_t(q)])] = ec[size_t(q)];
		    ts[size_t(el_idx[size_t(q)])] = es[size_t(q)];
		    if (ec[size_t(q)] >= 0) L.csize[size_t(ec[size_t(q)])] += R.csize[size_t(el_idx[size_t(q)])];
		}
		// every leaf under the right node: its clusters are where the right node's clusters went
		for (int b = R.lo; b <= R.hi; b++) {
		    auto& m = lmap[size_t(b)];
		    auto& sg = lsign[size_t(b)];
		    for (size_t q = 0; q < m.size(); q++) {
			const int32_t g = m[q];
			if (g < 0) continue;
			m[q] = tc[size_t(g)];
			sg[q] = ts[size_t(g)] == -1 ? -sg[q] : sg[q];
		    }
		}
		if (trace) {
		    std::lock_guard<std::mutex> g(trace_mutex);
		    icb_tree_trace::Job tj;
		    tj.left = jobs[size_t(j)].left;
		    tj.right = jobs[size_t(j)].right;
		    tj.level = level_no;
		    tj.cluster_of = tc;
		    tj.strand_of = ts;
		    trace->jobs.push_back(std::move(tj));
		}
		L.reps = L.cl->founder;
		L.hi = R.hi;
		L.depth++;
		std::vector<int32_t>().swap(R.reps);  // the right node's clusterer is released after the level (its stream may be busy)
	    }
	};
	std::vector<std::thread> th;
	for (int w = 0; w < nw; w++)
	    if (!per_worker[size_t(w)].empty()) th.emplace_back(work, w);
	for (auto& t : th) t.join();
	for (auto& j : jobs) {
	    Node& R = nodes[size_t(j.right)];
	    if (R.cl) {
		icb_clusterer_destroy(R.cl);
		R.cl = nullptr;
	    }
	}
	merges += int64_t(jobs.size());
	cur.swap(nxt);
    }
    // ---- final ids: a read's leaf cluster, mapped through every merge its leaf went through as part of a right operand
    if (!first_err.load()) {
	for (int b = 0; b < n_batches; b++) {
	    const auto& m = lmap[size_t(b)];
	    const auto& sg = lsign[size_t(b)];
	    for (int64_t s = bounds[b]; s < bounds[b + 1]; s++) {
		const int32_t c = cluster_of[s];
		if (c < 0) continue;
		cluster_of[s] = m[size_t(c)];
		strand_of[s] = m[size_t(c)] < 0 ? 0 : strand_of[s] * sg[size_t(c)];
	    }
	}
	if (n_clusters) *n_clusters = int64_t(nodes[0].reps.size());
    }
    for (auto& nd : nodes)
	if (nd.cl) {
	    icb_clusterer_destroy(nd.cl);
	    nd.cl = nullptr;
	}
    for (int w = 0; w < nw; w++) {
	if (stats) {
	    add_stats(*stats, wstats[size_t(w)]);
	    stats->ms_batch_max = std::max(stats->ms_batch_max, wstats[size_t(w)].ms_batch_max);
	}
	icb_ctx_fold_child(ctx, ctx->pool_ctx[size_t(w)]);
    }
    if (stats) {
	stats->ms_wall = wall.lap();
	stats->tree_merges = merges;
    }
    if (trace)  // the order a sequential driver would run them in: level by level, left to right
	std::sort(trace->jobs.begin(), trace->jobs.end(), [](const icb_tree_trace::Job& a, const icb_tree_trace::Job& b) {
	    return a.level != b.level ? a.level < b.level : a.left < b.left;
	});
    if (first_err.load()) {
	for (auto& e : werr)
	    if (!e.empty()) return ctx->fail(first_err.load(), e);
	return ctx->fail(first_err.load(), "worker failed");
    }
    return ICB_OK;
}

int icb_cluster_reads(icb_clusterer* cl, const int32_t* order, int64_t n, int32_t* cluster_of, int32_t* strand_of)
{
    IcbRange nvtx_range("ClusterSortedReads (icb_cluster_reads)");
    if (!cl || n < 0 || (n > 0 && (!order || !cluster_of || !strand_of))) return ICB_ERR_ARG;
    icb_ctx* ctx = cl->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    const icb_cluster_params& P = cl->P;
    std::vector<int32_t> chunk;
    std::vector<int64_t> slots;
    struct CpuClock {
	icb_cluster_stats& st;
	double t0 = thread_cpu_ms();
	~CpuClock() { st.ms_cpu += thread_cpu_ms() - t0; }
    } cpu_clock{cl->stats};
    for (int64_t s = 0; s < n; s++) {
	cluster_of[s] = -1;
	strand_of[s] = 0;
	const int32_t r = order[s];
	if (r < 0 || r >= cl->rs->n_reads) return ctx->fail(ICB_ERR_ARG, "read index outside the read set");
	// the skips of ClusterSortedReads (cluster.cpp:116-160)
	if (cl->rs->h_status[r] != 1) continue;
	if (cl->rs->h_score[r] < 0) continue;
	const int64_t L = cl->rs->h_offs[r + 1] - cl->rs->h_offs[r];
	if (uint64_t(L) < uint64_t(unsigned(2 * P.kmer_size))) continue;
	if (uint64_t(cl->rs->h_hpc_len[r]) < uint64_t(unsigned(2 * P.kmer_size))) continue;
	if ((-10 * std::log10(cl->rs->h_err[r])) <= P.min_qual) continue;
	chunk.push_back(r);
	slots.push_back(s);
	if (int(chunk.size()) == cl->chunk_cap) {
	    int rc = cl->run_chunk(chunk, slots, cluster_of, strand_of);
	    if (rc) return rc;
	    chunk.clear();
	    slots.clear();
	}
    }
    if (!chunk.empty()) {
	int rc = cl->run_chunk(chunk, slots, cluster_of, strand_of);
	if (rc) return rc;
    }
    return ICB_OK;
}

int64_t icb_clusterer_n_clusters(const icb_clusterer* cl) { return cl ? int64_t(cl->founder.size()) : 0; }

int icb_clusterer_founders(const icb_clusterer* cl, int32_t* founder_read)
{
    if (!cl || !founder_read) return ICB_ERR_ARG;
    memcpy(founder_read, cl->founder.data(), sizeof(int32_t) * cl->founder.size());
    return ICB_OK;
}

int icb_clusterer_stats(const icb_clusterer* cl, icb_cluster_stats* out)
{
    if (!cl || !out) return ICB_ERR_ARG;
    *out = cl->stats;
    return ICB_OK;
}

// the C ABI never throws: a failing thread start (std::system_error) or allocation inside becomes an error code
int icb_cluster_batches(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds,
			int32_t n_batches, int32_t n_workers, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
			icb_cluster_stats* stats)
{
    try {
	return cluster_batches_impl(rs, params, order, bounds, n_batches, n_workers, cluster_of, strand_of, n_clusters, stats);
    }
    catch (const std::bad_alloc&) {
	return rs ? rs->ctx->fail(ICB_ERR_NOMEM, "out of host memory") : ICB_ERR_NOMEM;
    }
    catch (const std::exception& e) {
	return rs ? rs->ctx->fail(ICB_ERR_STATE, e.what()) : ICB_ERR_STATE;
    }
}

int icb_cluster_tree(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds, int32_t n_batches,
		     int32_t n_workers, int32_t min_cls_size, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
		     icb_cluster_stats* stats)
{
    try {
	return cluster_tree_impl(rs, params, order, bounds, n_batches, n_workers, min_cls_size, cluster_of, strand_of, n_clusters, stats, nullptr);
    }
    catch (const std::bad_alloc&) {
	return rs ? rs->ctx->fail(ICB_ERR_NOMEM, "out of host memory") : ICB_ERR_NOMEM;
    }
    catch (const std::exception& e) {
	return rs ? rs->ctx->fail(ICB_ERR_STATE, e.what()) : ICB_ERR_STATE;
    }
}

int icb_cluster_tree_traced(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds, int32_t n_batches,
			    int32_t n_workers, int32_t min_cls_size, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
			    icb_cluster_stats* stats, icb_tree_trace** trace)
{
    if (!trace) return ICB_ERR_ARG;
    *trace = nullptr;
    try {
	std::unique_ptr<icb_tree_trace> t(new icb_tree_trace);
	const int rc = cluster_tree_impl(rs, params, order, bounds, n_batches, n_workers, min_cls_size, cluster_of, strand_of, n_clusters, stats,
					 t.get());
	if (rc) return rc;
	*trace = t.release();
	return ICB_OK;
    }
    catch (const std::bad_alloc&) {
	return rs ? rs->ctx->fail(ICB_ERR_NOMEM, "out of host memory") : ICB_ERR_NOMEM;
    }
    catch (const std::exception& e) {
	return rs ? rs->ctx->fail(ICB_ERR_STATE, e.what()) : ICB_ERR_STATE;
    }
}

void icb_tree_trace_destroy(icb_tree_trace* t) { delete t; }

int32_t icb_tree_trace_jobs(const icb_tree_trace* t) { return t ? int32_t(t->jobs.size()) : 0; }

int icb_tree_trace_leaves(const icb_tree_trace* t, const int32_t** cluster_of, const int32_t** strand_of)
{
    if (!t || !cluster_of || !strand_of) return ICB_ERR_ARG;
    *cluster_of = t->leaf_cluster.data();
    *strand_of = t->leaf_strand.data();
    return ICB_OK;
}

int icb_tree_trace_job(const icb_tree_trace* t, int32_t j, int32_t* left_batch, int32_t* right_batch, int64_t* n_right,
		       const int32_t** cluster_of, const int32_t** strand_of)
{
    if (!t || j < 0 || size_t(j) >= t->jobs.size() || !left_batch || !right_batch || !n_right || !cluster_of || !strand_of) return ICB_ERR_ARG;
    const icb_tree_trace::Job& job = t->jobs[size_t(j)];
    *left_batch = job.left;
    *right_batch = job.right;
    *n_right = int64_t(job.cluster_of.size());
    *cluster_of = job.cluster_of.data();
    *strand_of = job.strand_of.data();
    return ICB_OK;
}

}  // extern "C"
