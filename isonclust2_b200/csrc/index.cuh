// Stage 2 — device-resident minimizer -> cluster index and the per-chunk candidate query.
#pragma once
#include "icb_common.cuh"

// Posting lists as chains hanging off an open-addressing hash of the minimizer value:
//   keys[slot]  minimizer value (EMPTY = 0xFFFFFFFF; the real value 0xFFFFFFFF lives in slot `capacity`)
//   heads[slot] newest node of the chain (NIL = 0xFFFFFFFF)
//   nodes[i]    {id, next}; id = cluster id (main index) or C0 + position in the chunk (chunk index)
// Append-only, like the reference's MinimizerDB without consensus (AddMinimizers, minimizer.cpp:31-42):
// one (value, cluster) pair per distinct minimizer value of the founder's FORWARD sketch.  UpdateMinDB
// (minimizer.cpp:124-160) marks the nodes of a cluster that left a list dead (id = 0xFFFFFFFF) and appends new ones.
struct IcbDLoopState;

struct IcbIndexView {
    uint32_t* keys;
    uint32_t* heads;
    uint2* nodes;
    uint32_t mask;	 // capacity - 1 (capacity is a power of two)
    uint32_t* node_count;
};

struct IcbIndex {
    icb_ctx* ctx = nullptr;
    IcbIndexView v{};
    uint32_t capacity = 0;
    int64_t node_capacity = 0;
    int64_t nodes_upper = 0;  // upper bound of nodes inserted so far (host-side bookkeeping)
    int create(icb_ctx* c, uint32_t cap_pow2, int64_t node_cap);
    int clear();
    int reserve(int64_t extra_nodes);  // grows / rehashes as needed before an insert of `extra_nodes`
    void destroy();
};

// per candidate record returned by the query
struct IcbCand {
    int32_t id;	     // cluster id (< C0) or C0 + chunk position of a provisional founder
    int32_t strand;  // +1 / -1
    int32_t count;   // shared minimizers (SortedHit::Size)
    int32_t tmapped; // totalMapped of getMappedRatio as an integer (MODE relevant) / first hit Index (MODE full)
};

struct IcbQueryHeader {
    int32_t top_snap;	 // largest count among snapshot clusters
    int32_t n_cand;	 // records written
    int32_t overflow;	 // 1: more relevant candidates than the per-read capacity (the host asks again with room for n_rel_all)
    int32_t n_total;	 // number of (cluster,strand) keys with at least one hit (snapshot + provisional)
    int32_t n_rel_all;	 // relevant candidates of this read (>= n_cand)
    int32_t hits;	 // posting entries matched by the read's minimizers (H of SURVEY 8d)
    uint32_t cand_off;	 // pool mode: first record of this read in the candidate pool
    int32_t pad;
};

struct IcbQueryArgs {
    // reads
    const icb_minimizer* arena;
    const int64_t* min_off;
    const int32_t* nf;
    const int32_t* nr;
    const int32_t* hpc_len;
    const uint8_t* read_class;	// error class (0..14) of each read's compressed error rate
    // the chunk
    const int32_t* chunk_reads;	// read ids, in greedy order
    const int32_t* list;	// positions inside the chunk to process (NULL = all)
    int32_t n_list;
    int32_t n_chunk;
    // indices
    IcbIndexView main_idx, chunk_idx;
    int32_t C0;			// clusters in the snapshot
    const int32_t* cls_founder; // founder read of each snapshot cluster
    // parameters
    const int32_t* nmax;	// [15*15] largest run of missed minimizers still counted as mapped
    int32_t min_shared;
    double min_frac;
    int32_t cap;		// records per read
    int32_t mode;		// 0 relevant candidates with totalMapped, 1 full list with first-hit index
    // outputs
    IcbQueryHeader* hdr;	// per processed read
    IcbCand* cands;		// cap records per processed read (mode 0) / at full_off (mode 1)
    const int64_t* full_off;	// mode 1: record offset per processed read
    // scratch
    uint32_t* g_scratch;	// per-CTA global scratch when the dense counters do not fit shared memory
    int64_t g_stride;		// in uint32 words
    int32_t smem_words;		// words of dynamic shared memory available for counters + bitmaps
    int32_t smem_cnt;		// 1: dense counters live in shared memory, 0: in g_scratch
    // device-resident loop (dloop.cuh): work lists and pools live on the device
    struct IcbDLoopState* dl_state;	// pool cursors, fallback flag; NULL on the host-loop path
    const int32_t* dl_n_list;		// number of entries of `list`, written by the gather kernel
    uint32_t res_tag;			// diagnostics: the batch's tag in the residency log (residency.cuh)
    IcbCand* dl_pool;			// mode 2: candidate pool (records) ...
    int32_t* dl_pool_aln;		// ... and its alignment results, reset to "unknown"
    IcbCand* dl_full_pool;		// mode 2: full-list pool; reads whose best candidates are tied get their full list right away
    uint32_t* dl_depmask;		// [pos][dl_mw] provisional founders among the read's candidates
    int32_t dl_mw;
    int64_t* dl_full_off;		// mode 1: per chunk position, where the read's full list starts / its length
    int32_t* dl_full_n;
};

int icb_index_insert(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, const int32_t* d_reads, const int32_t* d_ids, int n);
// the same with the number of founders in device memory (*d_n <= n_max, the grid)
int icb_index_insert_dev(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, const int32_t* d_reads, const int32_t* d_ids, const int32_t* d_n,
			 int n_max);
int icb_index_update(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, int r_old, int r_new, int cls, int nf_old, int nf_new);
int icb_query_launch(icb_ctx* ctx, const IcbQueryArgs& a, int n_blocks, size_t smem_bytes);
// arguments in device memory (the launch configuration is what a captured graph bakes in)
int icb_query_launch_ptr(icb_ctx* ctx, const IcbQueryArgs* d_args, bool smem_cnt, int n_blocks, size_t smem_bytes);
int icb_classify_errors(icb_ctx* ctx, const double* d_err, const int32_t* d_status, int64_t n, uint8_t* d_class);
