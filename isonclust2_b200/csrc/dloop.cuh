// Device-resident greedy loop — shared definitions of dloop.cu (the decision / commit / work-queue kernels), index.cu
// (the pool-mode candidate query) and the alignment back ends (device-built pair queues).
//
// The reference's ClusterSortedReads (cluster.cpp:67-322) is sequential: read i sees the clusters founded by reads < i.
// A chunk of sorted reads is scored against the index snapshot + a chunk index of provisional founders (index.cu); what
// remains sequential — which provisional founders are real, the walk of getBestClusterMapping / getBestClusterAln over a
// read's candidates, and the reference's candidate ORDER among equally scored candidates — is evaluated by kernels on the
// chunk's device-resident records, in rounds: commit the decided prefix of the chunk, re-decide what the new founders
// touched, queue the alignments / full candidate lists the undecided reads wait for, run them, repeat.  The host only
// enqueues rounds and reads one small state block per chunk.
#pragma once
#include "align.cuh"
#include "index.cuh"

constexpr int DL_MAXW = 32;	   // words of a chunk-position bitmask: chunks of up to 1024 reads
constexpr int DL_SPEC = 3;	   // alignments a read with tied candidates may have in flight at once (the next ones of its plan)
constexpr int DL_THREADS = 256;	   // CTA size of the decision kernels
constexpr int DL_NMAX_S = 4096;   // candidates (cluster, strand) of one read the order emulation sorts in shared memory: the usual
constexpr int DL_NMAX_L = 8192;   // kernel, and the one for reads that hit a large (merged) index (powers of two: bitonic padding)

enum : uint8_t { DL_NONE = 0, DL_JOIN = 1, DL_NEW = 2, DL_NEED_PAIR = 3, DL_NEED_FULL = 4, DL_NEED_ORDER = 5 };
enum : int32_t { DL_ALN_UNKNOWN = -1, DL_ALN_REDO32 = -3 };
// reasons the device loop hands a chunk back to the host loop (icb_clusterer::process_chunk)
enum : int32_t { DL_FB_POOL = 1, DL_FB_TOO_MANY = 2, DL_FB_REHASH = 3, DL_FB_FULLPOOL = 4, DL_FB_STUCK = 5, DL_FB_FORCED = 6 };

struct IcbDLoopState {
    int32_t n_chunk, C0;
    int32_t commit_pos, next_cls, n_founders, done, fallback, rounds;
    uint32_t cand_cursor, cand_cap;
    unsigned long long full_cursor, full_cap;
    int32_t n_pairs16, n_pairs32, n_full, progress;
    int32_t n_order_s, n_order_l;  // reads whose candidate order is computed next (small / large shared-memory plan)
    unsigned long long tiles_cap, words_cap;  // alignment workspace of the stream, in 32-byte tiles / 32-bit words
    long long aln_invoked, aln_calls_ref, pairs_aligned, cells_aligned, tie_fallbacks, decisions, order_calls, hits, cands;
    long long stop_full, stop_pair_first, stop_pair_again;  // what the read at the commit frontier waited for at the end of a round
    uint32_t commask[DL_MAXW];   // chunk positions that founded a cluster so far
    uint32_t viewmask[DL_MAXW];  // ... or are predicted to: undecided positions whose decision currently is "new cluster"
};

struct IcbDLoopView {
    IcbDLoopState* st;
    const int32_t* chunk_reads;
    // per chunk position
    IcbQueryHeader* hdr;
    int32_t* founder_cls;  // -1, or the cluster the read at this position founded
    int32_t* view_cls;	   // the same under the PREDICTED view: committed founders, and the reads behind the commit frontier that
			   // will found a cluster if nothing before them changes, with the ids they would get
    uint8_t* outcome;
    uint8_t* flags;	   // bit 0 usedAln, bit 1 usedTie
    int32_t* dec_cls;
    int32_t* dec_strand;
    int32_t* aln_calls;	   // DP calls the reference makes for this read (getBestClusterAln stops at the first pass)
    int32_t* plan_n;	   // alignment plan: the tied candidates (slots of the read's candidate list) in the reference's order
    int32_t* plan_i;
    int32_t* need_slot;	   // [pos][DL_SPEC] candidates (slots) whose alignment the read waits for: the next unresolved ones of its plan
    int32_t* need_n;
    int32_t* req_count;	   // alignments queued for this read so far
    uint8_t* ord_req;	   // 1: order among the passing candidates of the mapping test, 2: among the candidates tied at the top
    int32_t* ord_target;   // the count of that group
    int64_t* full_off;	   // -1: no full list yet
    int32_t* full_n;
    uint32_t* depmask;	   // [pos][mw] provisional founders among the read's candidates (relevant list and full list)
    uint32_t* seenmask;	   // [pos][mw] commask when the read was last decided
    int32_t mw;
    int32_t* out_cls;
    int32_t* out_strand;
    int32_t* founder_pos;  // chunk positions that founded clusters, in order
    // pools
    IcbCand* cands;
    int32_t* aln;  // per candidate record: aligned window count, DL_ALN_UNKNOWN, DL_ALN_REDO32
    int32_t* plan;
    IcbCand* full;
    // read set
    const int64_t* offs;
    const double* raw_err;
    const int32_t* hpc_len;
    const int32_t* nf;
    const int32_t* nr;
    const uint32_t* bkt;   // per read: bucket count of the reference's hit map, _M_next_bkt(20 * (nf + nr)) (host libstdc++)
    int32_t* cls_founder;  // cluster id -> representative read (appended to by the finalize kernel)
    // parameters
    int32_t mode, min_shared, k;
    double min_frac, mapped_th, aligned_th;
    // work queues of a round
    IcbPairDesc* descs16;
    IcbPairDesc* descs32;
    int32_t* order_list_s;   // chunk positions whose candidate order is computed next
    int32_t* order_list_l;
    int32_t* full_list;	     // chunk positions whose full list is computed this round
    int64_t* full_list_off;
    int32_t* ins_reads;	     // finalize: founders to add to the main index
    int32_t* ins_ids;
    int32_t spec_pairs;	     // 1 .. DL_SPEC: how many of a read's next tied candidates are aligned ahead of the first one's result
    int32_t spec_window;     // reads this far behind the commit frontier at most get their candidate order computed ahead of time
    int32_t force_fallback;  // test hook: chunks whose C0 + n_chunk is divisible by this hand over to the host loop
};

// launchers (dloop.cu)
int icb_dl_init(icb_ctx* ctx, const IcbDLoopView& V, int n_chunk, int C0, uint32_t cand_cap, unsigned long long full_cap,
		unsigned long long tiles_cap, unsigned long long words_cap);
int icb_dl_decide(icb_ctx* ctx, const IcbDLoopView& V, int n_chunk);  // light decisions + the candidate-order kernels
int icb_dl_commit(icb_ctx* ctx, const IcbDLoopView& V);
int icb_dl_gather(icb_ctx* ctx, const IcbDLoopView& V, int max_pairs);
int icb_dl_finalize(icb_ctx* ctx, const IcbDLoopView& V, int* d_n_ins);
