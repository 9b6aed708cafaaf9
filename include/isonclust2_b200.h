/* isonclust2_b200 — C ABI of the B200 (sm_100a) implementation of isONclust2's
 * data-parallel clustering hot path.
 *
 * The reference (nanoporetech/isONclust2) has no plugin / FFI surface: every
 * function of the path is an internal C++ free function linked into one
 * binary.  This header is the boundary a maintainer would bind instead of
 * those functions; each entry point names the reference function(s) it
 * replaces (paths relative to the reference tree).  INTEGRATION.md shows the
 * call sites on the reference side.
 *
 * Conventions: plain pointers and sizes only; every function returns ICB_OK
 * (0) or a negative ICB_ERR_* code and never throws; a context owns one CUDA
 * device and one stream and is not shared between host threads concurrently
 * (create one context per thread); host pointers may be pageable or pinned
 * (pinned makes the copies asynchronous).  Reads are passed "flat": one byte
 * string of concatenated sequences, one of concatenated Phred+33 qualities
 * and offs[n_reads+1] (int64) with read i = [offs[i], offs[i+1]).
 */
#ifndef ISONCLUST2_B200_H
#define ISONCLUST2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICB_OK 0
#define ICB_ERR_CUDA -1		/* a CUDA runtime call failed; see icb_last_error */
#define ICB_ERR_ARG -2		/* invalid argument */
#define ICB_ERR_INVALID_BASE -3 /* a sketched read holds a base outside ACGT (reference: RevComp throws, util.cpp:31-33) */
#define ICB_ERR_NO_PEMP -4	/* no empirical p_shared row for this (k, w, e1, e2) (reference: GetPMinShared throws, p_emp_prob.cpp:88) */
#define ICB_ERR_STATE -5	/* call order violated (e.g. download before compute) */
#define ICB_ERR_NOMEM -6

typedef struct icb_ctx icb_ctx;
typedef struct icb_readset icb_readset;

/* {Min, Pos, Index}: byte-identical to the reference's Minimizer (minimizer.h:20-29) */
typedef struct {
    uint32_t Min, Pos, Index;
} icb_minimizer;

/* ------------------------------------------------------------------ context */
int icb_device_count(void);
int icb_ctx_create(int device, icb_ctx** ctx);
void icb_ctx_destroy(icb_ctx* ctx);
int icb_ctx_sync(icb_ctx* ctx);
/* the context's cudaStream_t, so a caller can bracket work with its own CUDA events */
void* icb_ctx_stream(icb_ctx* ctx);
const char* icb_last_error(icb_ctx* ctx);
const char* icb_strerror(int code);

/* Per-kernel device timing (CUDA events on the context's stream).  When enabled every kernel launch is
 * bracketed by an event pair; icb_ctx_kernel_time returns the number of launches and the summed
 * milliseconds recorded for the kernel called `name` since the last reset (synchronises the stream). */
int icb_ctx_timing(icb_ctx* ctx, int enable);
int icb_ctx_timing_reset(icb_ctx* ctx);
int icb_ctx_kernel_time(icb_ctx* ctx, const char* name, int64_t* launches, double* total_ms);
/* total kernel launches issued by this context since creation (always counted) */
int64_t icb_ctx_launch_count(icb_ctx* ctx);
/* Diagnostics: while the log runs, every CTA of the packed alignment kernel (kind 1) and every warp of its traceback (kind 2)
 * leaves one record of four 32-bit words {kind | SM id << 8, start, end, n | m << 16}, times in units of 1.024 us of the
 * device's global timer (tools/residency.py rebuilds the occupancy of the SMs over a step from them).  dump stops the log. */
int icb_debug_residency_start(icb_ctx* ctx, int64_t capacity);
int icb_debug_residency_dump(icb_ctx* ctx, uint32_t* records, int64_t capacity, int64_t* n_records);

/* ------------------------------------------------------------------ host-side helpers (no device work)
 * The three expressions of the path the reference evaluates with the HOST libm / libstdc++; they stay on the
 * host so results are bit-identical:
 *   icb_init_qual_tab    InitQualTab (nomin=0) / InitQualTabNomin (nomin=1), qualscore.cpp:156-180
 *   icb_quality_filter   filtered[i] = (-10*log10(err[i])) <= min_qual, qualscore.cpp:56 and cluster.cpp:157
 *   icb_sort_by_score    SortByQualScores, qualscore.cpp:138-145: stable, descending; order[slot] = read index */
void icb_init_qual_tab(double tab[129], int nomin);
int icb_quality_filter(const double* err, int64_t n, double min_qual, uint8_t* filtered);
int icb_sort_by_score(const double* score, int64_t n, int32_t* order);

/* ------------------------------------------------------------------ read set
 * Uploads the flat reads to the device (the only copy of the inputs the device stages need). */
int icb_readset_create(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n_reads,
		       icb_readset** rs);
/* The same with room behind the uploaded reads for extra_reads reads / extra_bases bases appended later, none longer than
 * extra_max_len: the consensus representatives of `sort -c <positive>` clustering (cluster.cpp:263-303), which replace a
 * cluster's representative sequence while the batch is being clustered. */
int icb_readset_create_reserve(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n_reads, int64_t extra_reads,
			       int64_t extra_bases, int64_t extra_max_len, icb_readset** rs);
/* Appends read `read` of the scored and sketched read set `src` (same context, same k / w) to `rs` — raw sequence,
 * compressed sequence and both sketches are copied on the device — with the error rates the reference gives a consensus
 * representative instead of the ones its quality string would give: UpdateClusterConsensus (consensus.cpp:57-63, 96-124)
 * sets RawSeq's error rate and score and HpcSeq's error rate to weighted means of the cluster's and the joining read's.
 * *index = the read's index in rs (usable with icb_clusterer_update_representative, icb_align_pairs).  rs must be sketched. */
int icb_readset_append_from(icb_readset* rs, icb_readset* src, int64_t read, double raw_err, double raw_score, double hpc_err, int64_t* index);
/* Replaces the per-read error rates / scores computed by icb_score_reads and icb_sketch_reads (n_reads doubles each, NULL =
 * keep): a batch loaded from a `.cer` carries the values the reference clusters with, and for consensus representatives
 * those are not functions of the quality string.  Call before the first clusterer is created on rs. */
int icb_readset_set_errors(icb_readset* rs, const double* raw_err, const double* raw_score, const double* hpc_err);
void icb_readset_destroy(icb_readset* rs);
int64_t icb_readset_n_reads(const icb_readset* rs);
int64_t icb_readset_n_bases(const icb_readset* rs);

/* ------------------------------------------------------------------ stage 1a: quality scores
 * Replaces FillQualScores (qualscore.cpp:14-37) = CalcQualScore (:107-136) + CalcErrorRate (:147-154) for
 * every read.  qualtab / qualtab_nomin are InitQualTab() / InitQualTabNomin() (:156-180) computed by the
 * HOST libm (129 doubles each) — the device never evaluates pow().  Results stay on the device
 * (the sketch and cluster stages read them) until downloaded. */
int icb_score_reads(icb_readset* rs, int kmer_size, const double qualtab[129], const double qualtab_nomin[129]);
int icb_scores_download(icb_readset* rs, double* score, double* err);

/* ------------------------------------------------------------------ stage 1b: sketches
 * Replaces the per-read body of PrepareSortedBatch (qualscore.cpp:62-96): HomopolymerCompress (hpc.cpp:4-32),
 * CalcErrorRate on the compressed qualities, KmerEncodeSeq (kmer_index.cpp:5-17) of the compressed read and of
 * its RevComp (util.cpp:13-38), GetKmerMinimizers (minimizer.cpp:78-123) on both.
 *
 * skip (may be NULL): n_reads bytes, non-zero = the host's quality filter (qualscore.cpp:56, needs host
 * log10) rejected the read -> status 0, nothing computed.
 * status per read: 0 skipped by `skip`; 1 sketched; 2 raw read too short to compress (kept without sketch,
 * qualscore.cpp:91-96); 3 compressed read shorter than 2k or w (dropped, qualscore.cpp:65-73).
 * Returns ICB_ERR_INVALID_BASE if a read that reaches the RevComp step holds a non-ACGT base. */
int icb_sketch_reads(icb_readset* rs, int kmer_size, int window_size, const double qualtab_nomin[129],
		     const uint8_t* skip);
/* totals, to size the download buffers */
int icb_sketch_sizes(icb_readset* rs, int64_t* n_minimizers, int64_t* n_rev_minimizers);
/* Any pointer may be NULL to skip that array.
 *   status[n], hpc_len[n], hpc_err[n]
 *   hpc_seq / hpc_qual: n_bases bytes each; the compressed read i occupies [offs[i], offs[i]+hpc_len[i])
 *       — the SAME offsets as the raw read (compression never grows a read), the rest of the slot is unspecified
 *   mins: n_minimizers entries, read i = [min_offs[i], min_offs[i+1]) ; rev_mins likewise with rev_offs */
int icb_sketch_download(icb_readset* rs, int32_t* status, int32_t* hpc_len, double* hpc_err, char* hpc_seq,
			char* hpc_qual, icb_minimizer* mins, int64_t* min_offs, icb_minimizer* rev_mins,
			int64_t* rev_offs);

/* ------------------------------------------------------------------ stage 3: alignment verification
 * Replaces, per candidate pair of getBestClusterAln (cluster.cpp:461-515): ParasailAlign (cluster.cpp:408-423,
 * parasail_sg_trace_scan_16 with the _32 rerun on saturation), parasail_result_get_traceback
 * (traceback_template.c:13-190) and the window count of getAlnRatio (cluster.cpp:442-459).
 * read / rep index the read set (raw sequences); strand -1 aligns against RevComp(rep) (cluster.cpp:488-490).
 * gap_open = setGapOpen(e1+e2) (cluster.cpp:425-440) and limit = floor((1-(e1+e2))*k) (cluster.cpp:446) are
 * evaluated by the host (doubles); gap extend is 1, match +2, mismatch -2 (cluster.cpp:473-479).
 * aligned = number of k-column windows of the traceback's comparison string (all but the last one) holding at
 * least `limit` matches; the caller forms aligned / read_length >= AlignedThreshold (cluster.cpp:458, 503). */
typedef struct {
    int32_t read, rep, strand, gap_open, limit;
} icb_pair;
typedef struct {
    int32_t aligned, score, end_query, end_ref, comp_len;
} icb_align_result;
int icb_align_pairs(icb_readset* rs, const icb_pair* pairs, int64_t n_pairs, int kmer_size, icb_align_result* results);

/* ------------------------------------------------------------------ stage 2 + the greedy loop: clustering
 * Replaces ClusterSortedReads for an initial clustering (cluster.cpp:67-322 with consensus off, the default):
 * per read getBestCluster (cluster.cpp:530-568) = GetMinimizerHits (minimizer.cpp:44-76) + SortMinimizerHits
 * (cluster.cpp:622-636) + getBestClusterMapping / getMappedRatio (cluster.cpp:324-406) + getBestClusterAln
 * (cluster.cpp:461-515), then AddMinimizers (minimizer.cpp:31-42) for a new cluster.  The device keeps the
 * minimizer -> cluster index and scores chunks of reads against it; the host walks the reads in order.
 *
 * p_shared: the 15x15 grid InitMinSharedMap(k, w) builds (p_emp_prob.cpp:22-47), index [c1*15+c2] with error class
 * c = round(e*100)-1 (0.01..0.15); NaN = no row (-> ICB_ERR_NO_PEMP, where the reference throws).  It ships as
 * isonclust2_b200/data/p_emp_prob.bin.  The other fields are CmdArgs' (args.h:9-37). */
typedef struct {
    int32_t kmer_size, window_size;
    int32_t mode; /* 0 sahlin, 1 fast, 2 furious (args.h:7) */
    int32_t min_shared;
    double mapped_threshold, aligned_threshold, min_fraction, min_prob_no_hits, min_qual;
    double p_shared[225];
    int32_t chunk_size; /* reads scored per device round trip; 0 = default */
} icb_cluster_params;

typedef struct {
    int64_t chunks, rounds, reads_queried;
    int64_t aln_invoked;   /* the reference's ALN_INVOKED counter (cluster.cpp:21, 563) */
    int64_t pairs_aligned; /* DP problems solved (>= what the reference solves: speculative pairs included) */
    int64_t cells_aligned;
    int64_t tie_fallbacks; /* reads whose outcome depended on the reference's candidate order */
    int64_t decisions;	   /* evaluations of the per-read decision (>= reads: speculative re-evaluations) */
    double ms_query, ms_full, ms_align, ms_insert, ms_host; /* host wall-clock split of icb_cluster_reads */
    double ms_cpu; /* CPU time (CLOCK_THREAD_CPUTIME_ID) the calling threads spent inside icb_cluster_reads */
    double ms_order;	 /* part of ms_host spent reproducing the reference's candidate order for tie reads */
    int64_t order_calls; /* evaluations of that order */
    int64_t aln_calls_ref; /* DP problems the reference solves for the same reads (getBestClusterAln stops at the first pass) */
    double ms_batch_max;   /* icb_cluster_batches: wall clock of the slowest batch */
    double ms_wall;	   /* icb_cluster_batches: wall clock of the whole call */
    int64_t hyp_decisions; /* decisions re-taken under predicted founders, only to request device work one round trip earlier */
    int64_t hits;	   /* posting entries matched by the queried reads' minimizers (H of the index-probe byte count) */
    int64_t cands;	   /* candidate records written by the chunk queries (C of the index-probe byte count) */
    int64_t dl_chunks;	   /* chunks decided entirely on the device (device-resident loop) */
    int64_t dl_fallback_chunks; /* chunks the device loop handed back to the host loop */
    int64_t tree_merges;   /* icb_cluster_tree: merge jobs run */
    int64_t dl_fallback_reasons; /* bit r set: a chunk was handed back for reason r (1 candidate pool, 2 more candidates than the
			      order kernel's shared-memory plan, 3 the reference's hit map would rehash, 4 full-list pool, 5 workspace, 6 test hook) */
    /* device loop: what the read at the commit frontier waited for when a round ended — its full candidate list, its first
     * alignment, or another alignment (the next tied candidate, or a new first candidate after a founder changed its order) */
    int64_t dl_stop_full, dl_stop_pair_first, dl_stop_pair_again;
} icb_cluster_stats;

typedef struct icb_clusterer icb_clusterer;
/* rs must be scored and sketched with the same k / w */
int icb_clusterer_create(icb_readset* rs, const icb_cluster_params* params, icb_clusterer** cl);
void icb_clusterer_destroy(icb_clusterer* cl);
/* order[n]: read ids in sorted (greedy) order.  cluster_of[slot] = cluster the read of that slot ends in (a founder
 * gets the id of the cluster it founds), -1 if the read is skipped (cluster.cpp:116-160); strand_of[slot] = +1 / -1
 * (MatchStrand after cluster.cpp:235-246), 0 if skipped.  May be called repeatedly; later calls continue on the
 * clusters founded so far. */
int icb_cluster_reads(icb_clusterer* cl, const int32_t* order, int64_t n, int32_t* cluster_of, int32_t* strand_of);
/* Registers clusters that exist already — the left batch of a merge (`cluster -l A.cer -r B.cer`, cluster.cpp:67-106
 * works on leftBatch->MinDB as loaded): rep_reads[c] is the read whose forward sketch represents the next cluster id
 * (AddMinimizers, minimizer.cpp:31-42).  A merge is then icb_cluster_reads over the right batch's representatives in
 * the right batch's cluster order. */
int icb_clusterer_seed(icb_clusterer* cl, const int32_t* rep_reads, int64_t n);
/* UpdateMinDB (minimizer.cpp:124-160) together with the representative swap the reference does around it after a
 * consensus update (cluster.cpp:279-291): cluster `cluster` is represented by read `rep_read` of the read set from now on.
 * The cluster leaves the posting list of every forward minimizer value only the old representative had and joins the list of
 * every value only the new one has; mapping and alignment tests of later icb_cluster_reads calls use the new representative's
 * compressed length, error rate and sequence.  The host that keeps the spoa consensus appends the consensus sequence to the
 * read set it uploads (it is sketched like any read) and names it here. */
int icb_clusterer_update_representative(icb_clusterer* cl, int32_t cluster, int32_t rep_read);
int64_t icb_clusterer_n_clusters(const icb_clusterer* cl);
int icb_clusterer_founders(const icb_clusterer* cl, int32_t* founder_read);
int icb_clusterer_stats(const icb_clusterer* cl, icb_cluster_stats* stats);

/* Initial clustering of n_batches independent sorted batches of one read set — what the reference's pipeline runs as
 * one `isONclust2 cluster` process per batch file (README.md:105-117).  Batch b is order[bounds[b] .. bounds[b+1]);
 * cluster ids are batch-local.  n_workers host threads (0 = three per host core of the process's affinity mask or of ICB_HOST_CORES, at most 48; a worker sleeps while its stream is busy) each drive one CUDA stream, so
 * the device always has several batches' kernels in flight; results do not depend on n_workers.
 * n_clusters[n_batches] and stats (summed over the batches) may be NULL. */
int icb_cluster_batches(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds,
			int32_t n_batches, int32_t n_workers, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
			icb_cluster_stats* stats);

/* configs[4] in process: icb_cluster_batches for the leaves, then the merge tree the reference's pipeline runs as
 * `cluster -l A -r B` jobs over .cer files (README.md:105-117; main.cpp:238-382, cluster.cpp:67-106): adjacent nodes are
 * paired left to right level by level (an odd last node is carried up), and a merge clusters the right node's clusters, in
 * id order and through their representatives, against the left node's — on clusterers that stay RESIDENT between levels
 * (index, representatives, cluster ids), so nothing is re-parsed, re-uploaded, re-sketched or re-indexed.  cluster_of[slot]
 * / strand_of[slot] are the read's cluster and strand in the ROOT; *n_clusters the root's cluster count.  min_cls_size is
 * CmdArgs::MinClsSize of `sort` (args.h:21, default 3): clusters of a MERGED right node with fewer reads are dropped from the
 * merge and from the output (cluster.cpp:120-124), their reads end with cluster -1.  Equal to running
 * the reference's tree on the same batches (tests/test_tree_gpu.py). */
int icb_cluster_tree(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds, int32_t n_batches,
		     int32_t n_workers, int32_t min_cls_size, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
		     icb_cluster_stats* stats);

/* The same with every decision kept, for a caller that rebuilds the batch of every node — cluster lists in the order the
 * reference appends them, its MinDB, its `.cer` (the CLI's `tree`): the leaves' own decisions per sorted slot (batch-local
 * cluster ids: a read founds a cluster iff its id equals the number of clusters before it), and per merge job, in the order
 * a sequential driver runs them (level by level, left to right; nodes are named by their first batch), where each cluster
 * of the right node went: -1 dropped by MinClsSize, an id >= the left node's cluster count at that moment = founded there. */
typedef struct icb_tree_trace icb_tree_trace;
int icb_cluster_tree_traced(icb_readset* rs, const icb_cluster_params* params, const int32_t* order, const int64_t* bounds, int32_t n_batches,
			    int32_t n_workers, int32_t min_cls_size, int32_t* cluster_of, int32_t* strand_of, int64_t* n_clusters,
			    icb_cluster_stats* stats, icb_tree_trace** trace);
void icb_tree_trace_destroy(icb_tree_trace* trace);
int32_t icb_tree_trace_jobs(const icb_tree_trace* trace);
int icb_tree_trace_leaves(const icb_tree_trace* trace, const int32_t** cluster_of, const int32_t** strand_of);
int icb_tree_trace_job(const icb_tree_trace* trace, int32_t j, int32_t* left_batch, int32_t* right_batch, int64_t* n_right,
		       const int32_t** cluster_of, const int32_t** strand_of);

/* ------------------------------------------------------------------ per-primitive face
 * One entry point per reference function of the path, on HOST arrays, for a host that keeps its own ClusterSortedReads
 * (cluster.cpp:67-322) and containers, and for the C++ header shim include/isonclust2_compat/ (hpc.h, kmer_index.h,
 * minimizer.h, qualscore.h, cluster.h, p_emp_prob.h with the reference's signatures) that the reference's unit test
 * test/isONclust2_test.cpp compiles against.  Each call is one small launch; the batched entry points above are the
 * throughput path. */
/* HomopolymerCompress (hpc.cpp:4-32) of n flat reads (offs[0] = 0): compressed read i is written at offs[i] of hpc_seq /
 * hpc_qual (n_bases bytes each), hpc_len[i] long; qualities are the maximum over each run. */
int icb_hpc_compress(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n, char* hpc_seq, char* hpc_qual,
		     int32_t* hpc_len);
/* KmerEncodeSeq (kmer_index.cpp:5-17): k-mers i = 0 .. len-k-1 (the reference drops the last one) as 2-bit values,
 * BaseToNumber semantics for characters outside ACGT.  out needs len-k entries. */
int icb_kmer_encode(icb_ctx* ctx, const char* seq, int64_t len, int kmer_size, uint32_t* out, int64_t* n_out);
/* GetKmerMinimizers (minimizer.cpp:78-123): leftmost minimum of every window of w-k+1 k-mers, emitted when its position
 * changes.  out needs n entries. */
int icb_kmer_minimizers(icb_ctx* ctx, const uint32_t* kmers, int64_t n, int kmer_size, int window_size, icb_minimizer* out,
			int64_t* n_out);
/* CalcErrorRate (qualscore.cpp:147-154) of n quality strings with the given 129-entry table */
int icb_calc_error_rate(icb_ctx* ctx, const char* qual, const int64_t* offs, int64_t n, const double tab[129], double* out);
/* MinimizerDB (minimizer.h:60-61) on the device.  icb_index_add = AddMinimizers (minimizer.cpp:31-42; cluster ids must be
 * added in increasing order, as ClusterSortedReads does), icb_index_update_cluster = UpdateMinDB (minimizer.cpp:124-160),
 * icb_index_query_hits = GetMinimizerHits (minimizer.cpp:44-76): every (read minimizer, cluster) hit in the order the
 * reference appends them; *n_out > cap means nothing was written, ask again with room for *n_out hits. */
typedef struct icb_index icb_index;
typedef struct {
    uint32_t cls;
    int32_t strand; /* +1 / -1 */
    uint32_t pos, index; /* MinimizerHit {Pos, Index} (minimizer.h:63-66) */
} icb_hit;
int icb_index_create(icb_ctx* ctx, icb_index** index);
void icb_index_destroy(icb_index* index);
int icb_index_add(icb_index* index, uint32_t cls, const icb_minimizer* mins, int64_t n);
int icb_index_update_cluster(icb_index* index, uint32_t cls, const icb_minimizer* old_mins, int64_t n_old, const icb_minimizer* new_mins,
			     int64_t n_new);
int icb_index_query_hits(icb_index* index, const icb_minimizer* mins, int64_t n, const icb_minimizer* rev_mins, int64_t n_rev, icb_hit* out,
			 int64_t cap, int64_t* n_out);
/* totalMapped of getMappedRatio (cluster.cpp:324-353) for one candidate's hits (ascending Index): a run of g missed
 * minimizers is bridged iff g <= n_max, the largest g with pow(pError, g) >= minProbNoHits (host libm);
 * the caller divides by hpc_len. */
int icb_mapped_length(icb_ctx* ctx, const uint32_t* hit_pos, const uint32_t* hit_index, int64_t n_hits, int64_t n_minimizers, int64_t hpc_len,
		      int32_t n_max, int64_t* total_mapped);
/* getAlnRatio's window count (cluster.cpp:442-459) of a comparison string: windows of k columns, all but the last one,
 * with at least `limit` = floor((1-e)*k) match marks (0x7C) */
int icb_count_windows(icb_ctx* ctx, const char* comp, int64_t len, int kmer_size, int limit, int64_t* aligned);
/* ParasailAlign (cluster.cpp:408-423) + parasail_result_get_traceback (traceback_template.c:13-190) of ONE pair with the
 * comparison string materialised: comp (cap bytes, may be NULL) receives the traceback's `comp` line (0x7C match, blank
 * otherwise, free end gaps included), result->comp_len its length. */
int icb_align_traceback(icb_readset* rs, const icb_pair* pair, int kmer_size, icb_align_result* result, char* comp, int64_t cap);
/* InitMinSharedMap (p_emp_prob.cpp:22-47) as the 15 x 15 grid icb_cluster_params wants: path NULL = $ICB_PEMP_TABLE, else
 * data/p_emp_prob.bin next to this library.  Cells without a row are NaN. */
int icb_load_p_shared(const char* path, int kmer_size, int window_size, double p_shared[225]);

/* ------------------------------------------------------------------ consensus (host side of the path)
 * The partial-order alignment the reference gets from its vendored spoa 4.0.9 (AVX2 build), restated bit for bit on the
 * host (csrc/poa.cpp; north_star keeps spoa on the host): engine parameters of main.cpp:284-324 (match 4, mismatch -8,
 * convex gaps -8/-4, -20/-1), algo = the reference's `-A`: 0 local (kSW), 1 global (kNW), 2 semi-global (kOV, args.h:53).
 * A graph handle is a spoa::Graph.  None of these calls touches the device. */
typedef struct icb_poa_graph icb_poa_graph;
int icb_poa_graph_create(icb_poa_graph** graph);
void icb_poa_graph_destroy(icb_poa_graph* graph);
/* spoa::Graph::load / save through cereal's binary archive (graph.hpp:188-276): the ConsGs blobs of a `.cer`.  save writes
 * *n bytes if cap allows and always returns the size in *n. */
int icb_poa_graph_load(const void* bytes, int64_t n, icb_poa_graph** graph);
int icb_poa_graph_save(icb_poa_graph* graph, void* out, int64_t cap, int64_t* n);
/* sequences().size(), nodes().size(), edges().size() */
int icb_poa_graph_sizes(const icb_poa_graph* graph, int64_t* n_sequences, int64_t* n_nodes, int64_t* n_edges);
/* Graph::Clear (graph.cpp:648-657) */
int icb_poa_graph_clear(icb_poa_graph* graph);
/* AlignmentEngine::Align (simd_alignment_engine_implementation.hpp:703-780, 1586-2131): pairs (node id or -1, sequence
 * position or -1); *n_out pairs exist, written if cap allows. */
int icb_poa_align(icb_poa_graph* graph, const char* seq, int64_t len, int algo, int32_t* node_ids, int32_t* positions, int64_t cap,
		  int64_t* n_out, int32_t* score);
/* AddSeqToGraph (consensus.cpp:15-22): Align, then Graph::AddAlignment with one weight for every base */
int icb_poa_graph_add(icb_poa_graph* graph, const char* seq, int64_t len, uint32_t weight, int algo);
/* Graph::GenerateConsensus (graph.cpp:368-375): heaviest bundle with branch completion */
int icb_poa_graph_consensus(icb_poa_graph* graph, char* out, int64_t cap, int64_t* len);

#ifdef __cplusplus
}
#endif
#endif
