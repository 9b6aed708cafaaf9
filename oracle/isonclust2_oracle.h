/* TEST INFRASTRUCTURE — CPU restatement ("port") of isONclust2's clustering
 * hot path in plain C.  See isonclust2_oracle.c for the per-function
 * reference citations and the pinning status. */
#ifndef ISONCLUST2_ORACLE_H
#define ISONCLUST2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t Min, Pos, Index;
} orc_minimizer;

/* a1 */
void orc_init_qualtab(double tab[129], int nomin);
/* a2, a3 */
double orc_calc_qual_score(const char* qual, int64_t n, int k, const double tab[129]);
double orc_calc_error_rate(const char* qual, int64_t n, const double tab[129]);
/* a4: score/err of one read as FillQualScores stores them */
void orc_fill_qual_score(const char* qual, int64_t n, int k, const double tab[129], const double tab_nomin[129],
			 double* score, double* err);
/* a5: stable order by score desc; order[slot] = read index */
void orc_sort_by_score(const double* score, int64_t n, int32_t* order);
/* a6 */
int64_t orc_hpc(const char* seq, const char* qual, int64_t n, char* out_seq, char* out_qual);
/* a7: returns -1 on a non-ACGT base (the reference throws) */
int64_t orc_revcomp(const char* seq, int64_t n, char* out);
/* a8: returns number of values written (n-k, last k-mer dropped; 0 if n<k) */
int64_t orc_kmer_encode(const char* seq, int64_t n, int k, uint32_t* out);
/* a9: window-argmin formulation; returns number of minimizers */
int64_t orc_minimizers(const uint32_t* kmers, int64_t n, int k, int w, orc_minimizer* out);
/* a10: one read of PrepareSortedBatch, quality filter decided by the caller
 * (it needs host log10).  status: 1 sketched, 2 raw kept without hpc/mins
 * (never reached for n>=1 ... see .c), 3 too short after hpc (score:=-1, raw dropped).
 * Buffers: hpc_* hold n bytes; mins/rev_mins hold n entries each. */
int orc_sketch_read(const char* seq, const char* qual, int64_t n, int k, int w, const double tab_nomin[129],
		    char* hpc_seq, char* hpc_qual, int64_t* hpc_len, double* hpc_err, orc_minimizer* mins,
		    int64_t* n_mins, orc_minimizer* rev_mins, int64_t* n_rev);
/* quality filter of qualscore.cpp:56 / cluster.cpp:157 */
int orc_quality_filtered(double err, double min_qual);

/* a15: p_shared lookup given the 15x15 table selected for (k,w); e in real units */
int orc_err_class(double e); /* 0..14 for 0.01..0.15 */
/* a16: n_max[c1*15+c2] = largest n with pow(1-p,n) >= min_prob_no_hits */
void orc_nmax_table(const double p_shared[225], double min_prob_no_hits, int32_t nmax[225]);
/* a16 in the reference's own floating-point form */
double orc_mapped_ratio(int64_t hpc_len, int64_t n_mins, const uint32_t* hit_pos, const uint32_t* hit_idx,
			int64_t n_hits, double p_shared, double min_prob_no_hits);
/* a16 integer form: totalMapped given n_max */
int64_t orc_total_mapped(int64_t hpc_len, int64_t n_mins, const uint32_t* hit_pos, const uint32_t* hit_idx,
			 int64_t n_hits, int32_t nmax);
/* a18 */
int orc_set_gap_open(double e);
/* a19-a21: semi-global affine DP + traceback + k-window count.
 * comp may be NULL; else must hold n+m+1 bytes.  Returns aligned-window count. */
int64_t orc_align(const char* read, int64_t n, const char* ref, int64_t m, int gap_open, double e, int k,
		  int32_t* score, int64_t* end_query, int64_t* end_ref, char* comp, int64_t* comp_len);
/* a21 on a given comp string */
int64_t orc_aligned_windows(const char* comp, int64_t len, double e, int k);

/* a11, a13, a14, a17, a22, a23 + the greedy loop: cluster reads given their
 * sketches.  See .c.  Returns number of clusters. */
typedef struct {
    int k, w, mode; /* 0 sahlin 1 fast 2 furious */
    int min_shared;
    double mapped_th, aligned_th, min_frac, min_prob_no_hits, min_qual;
    double p_shared[225];
} orc_cluster_params;

typedef struct {
    int64_t n_reads;
    /* per read (sorted order) */
    const int32_t* status; /* 1 = sketched (RawSeq+HpcSeq present) else skipped */
    const double* score;
    const double* raw_err;
    const double* hpc_err;
    const char* raw_seq;
    const int64_t* raw_offs;
    const int64_t* hpc_len;
    const orc_minimizer* mins;
    const int64_t* min_offs;
    const orc_minimizer* rev_mins;
    const int64_t* rev_offs;
} orc_reads;

/* out_cluster[i] = cluster id read i ends in (-1 skipped), out_strand[i] =
 * match strand (+1 founder / joined fwd, -1 joined rev), out_founder[c] = read
 * index that founded cluster c.  n_aligned (may be NULL) = ALN_INVOKED. */
int64_t orc_cluster_reads(const orc_cluster_params* P, const orc_reads* R, int32_t* out_cluster, int32_t* out_strand,
			  int32_t* out_founder, int64_t* n_aligned);

/* The same loop on clusters that exist already and whose representatives were replaced (a12 through its caller,
 * cluster.cpp:279-291): cluster c is seeded from read seeds[c], update u makes read upd_read[u] the representative of
 * cluster upd_cls[u] (UpdateMinDB, minimizer.cpp:124-160), then queries[0..n_q) are clustered in that order.
 * out_cluster / out_strand are per query; out_founder needs n_seeds + n_q entries.  Returns the number of clusters. */
int64_t orc_cluster_seeded(const orc_cluster_params* P, const orc_reads* R, const int32_t* seeds, int64_t n_seeds,
			   const int32_t* upd_cls, const int32_t* upd_read, int64_t n_upd, const int32_t* queries, int64_t n_q,
			   int32_t* out_cluster, int32_t* out_strand, int32_t* out_founder, int64_t* n_aligned);

/* a12: UpdateMinDB (minimizer.cpp:124-160) applied to a MinDB built by AddMinimizers; export sorted by key.
 * Returns the number of keys, or -1 - needed if a capacity is too small. */
int64_t orc_mindb_after_updates(const orc_minimizer* cls_mins, const int64_t* cls_offs, int n_cls, const int32_t* upd_cls,
				const orc_minimizer* upd_mins, const int64_t* upd_offs, int n_upd, uint32_t* keys, int64_t* offs,
				uint32_t* cls, int64_t key_cap, int64_t cls_cap);

#ifdef __cplusplus
}
#endif
#endif
