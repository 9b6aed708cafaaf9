/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the data-parallel hot path of isONclust2 in
 * plain C: per-read quality scoring, homopolymer compression, 2-bit k-mer
 * encoding, window minimizers, mapped-ratio test, semi-global affine alignment
 * with traceback and the k-window count.  The order-dependent greedy loop that
 * calls them is restated in isonclust2_oracle_cluster.cpp (C++, because its
 * candidate order is defined by libstdc++'s unordered_map and std::sort).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library, and only as the checker / the CPU baseline.  The product
 * (isonclust2_b200/) never links or calls it.
 *
 * PINNING: parity pinned.  tests/test_oracle_*.py check every function here
 * against (i) the 8 golden values of the reference's own unit tests
 * (test/isONclust2_test.cpp:17-203) and (ii) the reference itself, compiled
 * unmodified into oracle/_ref/libisonclust2_ref.so, on seeded random inputs;
 * committed fixtures under tests/golden/ were produced by that library
 * (tests/golden/make_golden.py).
 *
 * Floating point: the reference is built -O3 -msse3 without FMA
 * (CMakeLists.txt:19-21,42-53), so this file must be compiled with
 * -ffp-contract=off and performs every FP64 operation in the reference's order.
 */
#include "isonclust2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ a1 */
/* qualscore.cpp:156-180  InitQualTab / InitQualTabNomin */
void orc_init_qualtab(double tab[129], int nomin)
{
    for (int i = 0; i < 129; i++) tab[i] = 0.0;
    for (int i = 33; i <= 128; i++) {
	double t = pow(10, -((i - 33) / 10.0));
	if (!nomin && t > 0.79433) t = 0.79433;
	tab[i] = t;
    }
}

/* ------------------------------------------------------------------ a2 */
/* qualscore.cpp:107-136  CalcQualScore: product of the first k no-error
 * probabilities, then a rolling update cur *= (enter/leave). */
double orc_calc_qual_score(const char* qual, int64_t n, int k, const double tab[129])
{
    if (n <= k) return -1.0;
    double cur = 1.0;
    for (int i = 0; i < k; i++) cur *= 1.0 - tab[(int)qual[i]];
    double sum = cur;
    for (int64_t i = k; i < n; i++) {
	double enter = 1.0 - tab[(int)qual[i]];
	double leave = 1.0 - tab[(int)qual[i - k]];
	cur *= (enter / leave);
	sum += cur;
    }
    return sum;
}

/* ------------------------------------------------------------------ a3 */
/* qualscore.cpp:147-154  CalcErrorRate */
double orc_calc_error_rate(const char* qual, int64_t n, const double tab[129])
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += tab[(int)qual[i]];
    return s / (double)(uint64_t)n;
}

/* ------------------------------------------------------------------ a4 */
/* qualscore.cpp:14-37  FillQualScores, one read */
void orc_fill_qual_score(const char* qual, int64_t n, int k, const double tab[129], const double tab_nomin[129],
			 double* score, double* err)
{
    if ((uint64_t)n > (uint64_t)(unsigned)(2 * k)) {
	double qs = orc_calc_qual_score(qual, n, k, tab);
	if (qs <= 0) qs = -1.0;
	*score = qs;
	*err = orc_calc_error_rate(qual, n, tab_nomin);
    }
    else {
	*score = -1.0;
	*err = 1.0;
    }
}

/* ------------------------------------------------------------------ a5 */
/* qualscore.cpp:138-145  std::stable_sort, a before b iff score(a) > score(b).
 * Restated as a bottom-up merge sort (any stable sort yields the same order). */
void orc_sort_by_score(const double* score, int64_t n, int32_t* order)
{
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
    for (int64_t width = 1; width < n; width *= 2) {
	for (int64_t lo = 0; lo < n; lo += 2 * width) {
	    int64_t mid = lo + width < n ? lo + width : n;
	    int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
	    int64_t a = lo, b = mid, o = lo;
	    while (a < mid && b < hi) {
		/* take from the right run only if it is strictly better */
		if (score[order[b]] > score[order[a]]) tmp[o++] = order[b++];
		else tmp[o++] = order[a++];
	    }
	    while (a < mid) tmp[o++] = order[a++];
	    while (b < hi) tmp[o++] = order[b++];
	}
	memcpy(order, tmp, sizeof(int32_t) * (size_t)n);
    }
    free(tmp);
}

/* ------------------------------------------------------------------ a6 */
/* hpc.cpp:4-32  HomopolymerCompress: one base per run, quality = max over the
 * run (signed char comparison, hpc.cpp:24). */
int64_t orc_hpc(const char* seq, const char* qual, int64_t n, char* out_seq, char* out_qual)
{
    if (n <= 0) return 0;
    int64_t o = 0;
    int64_t i = 0;
    while (i < n) {
	int64_t j = i;
	char q = qual[i];
	while (j + 1 < n && seq[j + 1] == seq[i]) {
	    j++;
	    if (q < qual[j]) q = qual[j];
	}
	out_seq[o] = seq[i];
	out_qual[o] = q;
	o++;
	i = j + 1;
    }
    return o;
}

/* ------------------------------------------------------------------ a7 */
/* util.cpp:13-38  RevComp */
int64_t orc_revcomp(const char* seq, int64_t n, char* out)
{
    for (int64_t i = 0; i < n; i++) {
	char c = seq[n - 1 - i];
	char r;
	switch (c) {
	    case 'A': r = 'T'; break;
	    case 'C': r = 'G'; break;
	    case 'G': r = 'C'; break;
	    case 'T': r = 'A'; break;
	    default: return -1;
	}
	out[i] = r;
    }
    return n;
}

/* ------------------------------------------------------------------ a8 */
/* kmer_index.h:29-45,59-66 + kmer_index.cpp:5-17.  value = sum code*4^(k-1-j)
 * in uint32 arithmetic (wraps for k>16); non-ACGT codes as (unsigned)-1;
 * only n-k values (the last k-mer is dropped, kmer_index.cpp:12). */
static inline uint32_t base_code(char c)
{
    switch (c) {
	case 'A': return 0u;
	case 'C': return 1u;
	case 'G': return 2u;
	case 'T': return 3u;
    }
    return 0xFFFFFFFFu;
}

int64_t orc_kmer_encode(const char* seq, int64_t n, int k, uint32_t* out)
{
    if (n < k) return 0;
    for (int64_t i = 0; i < n - k; i++) {
	uint32_t v = 0;
	for (int j = 0; j < k; j++) v = 4u * v + base_code(seq[i + j]);
	out[i] = v;
    }
    return n - k;
}

/* ------------------------------------------------------------------ a9 */
/* minimizer.cpp:78-123  GetKmerMinimizers, restated as SURVEY.md App. A1:
 * W = w-k+1 consecutive k-mers per window, minimizer = leftmost minimum, a
 * minimizer is emitted for window 0 and whenever the argmin position changes. */
int64_t orc_minimizers(const uint32_t* kmers, int64_t n, int k, int w, orc_minimizer* out)
{
    int64_t W = (int64_t)w - k + 1;
    int64_t cnt = 0;
    int64_t prev = -1;
    for (int64_t j = 0; j + W <= n; j++) {
	int64_t am = j;
	for (int64_t t = j + 1; t < j + W; t++)
	    if (kmers[t] < kmers[am]) am = t;
	if (j == 0 || am != prev) {
	    out[cnt].Min = kmers[am];
	    out[cnt].Pos = (uint32_t)am;
	    out[cnt].Index = (uint32_t)cnt;
	    cnt++;
	}
	prev = am;
    }
    return cnt;
}

/* ------------------------------------------------------------------ a10 */
/* qualscore.cpp:56 and cluster.cpp:157 */
int orc_quality_filtered(double err, double min_qual) { return (-10 * log10(err)) <= min_qual; }

/* qualscore.cpp:62-96, one read that passed the quality filter.
 * The else-branch at :91-96 (raw length <= 2k and < w) moves the raw read in
 * without a sketch -> status 2. */
int orc_sketch_read(const char* seq, const char* qual, int64_t n, int k, int w, const double tab_nomin[129],
		    char* hpc_seq, char* hpc_qual, int64_t* hpc_len, double* hpc_err, orc_minimizer* mins,
		    int64_t* n_mins, orc_minimizer* rev_mins, int64_t* n_rev)
{
    *hpc_len = 0;
    *n_mins = 0;
    *n_rev = 0;
    *hpc_err = -1.0;
    if (!((uint64_t)n > (uint64_t)(unsigned)(2 * k) || (uint64_t)n >= (uint64_t)(unsigned)w)) return 2;
    int64_t lh = orc_hpc(seq, qual, n, hpc_seq, hpc_qual);
    *hpc_len = lh;
    if ((uint64_t)lh < (uint64_t)(unsigned)(2 * k) || (uint64_t)lh < (uint64_t)(unsigned)w) return 3;
    uint32_t* kv = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)lh);
    char* rc = (char*)malloc((size_t)lh);
    int64_t nk = orc_kmer_encode(hpc_seq, lh, k, kv);
    *hpc_err = orc_calc_error_rate(hpc_qual, lh, tab_nomin);
    *n_mins = orc_minimizers(kv, nk, k, w, mins);
    if (orc_revcomp(hpc_seq, lh, rc) < 0) {
	free(kv);
	free(rc);
	return -1; /* the reference terminates here */
    }
    nk = orc_kmer_encode(rc, lh, k, kv);
    *n_rev = orc_minimizers(kv, nk, k, w, rev_mins);
    free(kv);
    free(rc);
    return 1;
}

/* ------------------------------------------------------------------ a15 */
/* p_emp_prob.cpp:66-94 with util.cpp:6-10: round to 2 decimals, clamp to
 * [0.01, 0.15]; class index = round(e*100) - 1. */
int orc_err_class(double e)
{
    int decimals = (int)pow(10, 2);
    double r = round(e * decimals) / decimals;
    if (r > 0.15) r = 0.15;
    if (r < 0.01) r = 0.01;
    int c = (int)lround(r * 100.0) - 1;
    if (c < 0) c = 0;
    if (c > 14) c = 14;
    return c;
}

/* ------------------------------------------------------------------ a16 */
/* cluster.cpp:324-353: a run of n missed minimizers counts as mapped iff
 * pow(pError, n) >= minProbNoHits.  pError in (0,1) makes that monotone in n,
 * so the test is n <= n_max. */
void orc_nmax_table(const double p_shared[225], double min_prob_no_hits, int32_t nmax[225])
{
    for (int c = 0; c < 225; c++) {
	double pe = 1.0 - p_shared[c];
	int32_t n = -1;
	/* largest n with pow(pe, n) >= th; scan, the bound is a few hundred */
	for (int32_t t = 0; t < (1 << 20); t++) {
	    if (pow(pe, (double)t) >= min_prob_no_hits) n = t;
	    else break;
	}
	nmax[c] = n;
    }
}

double orc_mapped_ratio(int64_t hpc_len, int64_t n_mins, const uint32_t* hit_pos, const uint32_t* hit_idx,
			int64_t n_hits, double p_shared, double min_prob_no_hits)
{
    double pe = 1.0 - p_shared;
    double total = 0;
    if (pow(pe, hit_idx[0]) >= min_prob_no_hits) total += (double)hit_pos[0];
    for (int64_t i = 0; i + 1 < n_hits; i++) {
	double p = pow(pe, (double)(uint32_t)(hit_idx[i + 1] - (hit_idx[i] + 1)));
	if (p >= min_prob_no_hits) total += (double)(uint32_t)(hit_pos[i + 1] - hit_pos[i]);
    }
    {
	uint32_t hi = hit_idx[n_hits - 1];
	if (pow(pe, (double)(uint64_t)((uint64_t)n_mins - (hi + 1))) >= min_prob_no_hits)
	    total += (double)(uint64_t)((uint64_t)hpc_len - hit_pos[n_hits - 1]);
    }
    return total / (double)hpc_len;
}

int64_t orc_total_mapped(int64_t hpc_len, int64_t n_mins, const uint32_t* hit_pos, const uint32_t* hit_idx,
			 int64_t n_hits, int32_t nmax)
{
    int64_t t = 0;
    if ((int64_t)hit_idx[0] <= nmax) t += hit_pos[0];
    for (int64_t i = 0; i + 1 < n_hits; i++)
	if ((int64_t)hit_idx[i + 1] - (int64_t)hit_idx[i] - 1 <= nmax) t += (int64_t)hit_pos[i + 1] - hit_pos[i];
    if (n_mins - (int64_t)hit_idx[n_hits - 1] - 1 <= nmax) t += hpc_len - hit_pos[n_hits - 1];
    return t;
}

/* ------------------------------------------------------------------ a18 */
/* cluster.cpp:425-440 */
int orc_set_gap_open(double e)
{
    if (e <= 0.01) return 5;
    if (e <= 0.04) return 4;
    if (e <= 0.1) return 3;
    return 2;
}

/* ------------------------------------------------------------- a19, a20 */
/* Semi-global (all four ends free) affine-gap DP with parasail's trace
 * semantics: scalar spec parasail sg_trace.c:97-232 (recurrence, trace bits),
 * end-cell rule as executed by sg_trace_scan_avx2_256_16.c:355-395, traceback
 * walk traceback_template.c:72-175.  Matrix = parasail_matrix_create("ACGT",
 * 2,-2) (memory.c:336-413): case-insensitive, anything outside ACGT scores 0. */
enum { T_INS = 1, T_DEL = 2, T_DIAG = 4, T_DIAG_E = 8, T_INS_E = 16, T_DIAG_F = 32, T_DEL_F = 64 };

static inline int map_base(char c)
{
    switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
    }
    return 4;
}
static inline int up(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

int64_t orc_aligned_windows(const char* comp, int64_t len, double e, int k)
{
    /* cluster.cpp:442-459: windows start at 0 .. len-k-1 (the last one is
     * never tested); nm >= limit compares a long against a double. */
    double limit = floor((1.0 - e) * (unsigned)k);
    int64_t aligned = 0;
    for (int64_t t = 0; t + k < len; t++) {
	long nm = 0;
	for (int j = 0; j < k; j++) nm += comp[t + j] == '|';
	if ((double)nm >= limit) aligned++;
    }
    return aligned;
}

int64_t orc_align(const char* read, int64_t n, const char* ref, int64_t m, int gap_open, double e, int k,
		  int32_t* score_out, int64_t* end_query, int64_t* end_ref, char* comp_out, int64_t* comp_len)
{
    const int32_t NEG = INT32_MIN / 2;
    const int o = gap_open, g = 1;
    uint8_t* tr = (uint8_t*)malloc((size_t)n * (size_t)m);
    int32_t* H = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    int32_t* F = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m + 1));
    int32_t* lastcol = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    for (int64_t j = 0; j <= m; j++) {
	H[j] = 0;
	F[j] = NEG;
    }
    for (int64_t i = 1; i <= n; i++) {
	int a = map_base(read[i - 1]);
	int32_t nh = H[0]; /* H[i-1][0] */
	int32_t wh = 0;	   /* H[i][0]   */
	int32_t E = NEG;
	H[0] = wh;
	for (int64_t j = 1; j <= m; j++) {
	    int32_t nwh = nh;
	    nh = H[j];
	    int32_t f_opn = nh - o, f_ext = F[j] - g;
	    int32_t f = f_opn > f_ext ? f_opn : f_ext;
	    F[j] = f;
	    int32_t e_opn = wh - o, e_ext = E - g;
	    E = e_opn > e_ext ? e_opn : e_ext;
	    int b = map_base(ref[j - 1]);
	    int s = (a == 4 || b == 4) ? 0 : (a == b ? 2 : -2);
	    int32_t d = nwh + s;
	    wh = d > E ? d : E;
	    if (f > wh) wh = f;
	    H[j] = wh;
	    uint8_t t = (f_opn > f_ext) ? T_DIAG_F : T_DEL_F;
	    t |= (e_opn > e_ext) ? T_DIAG_E : T_INS_E;
	    t |= (wh == d) ? T_DIAG : (wh == f) ? T_DEL : T_INS;
	    tr[(i - 1) * m + (j - 1)] = t;
	}
	lastcol[i - 1] = wh;
    }
    /* end cell: best of the last row (first maximum), then the last column
     * (strictly better wins; among equals the smallest row; a tie with the
     * row maximum only moves the end if that maximum sits in the corner). */
    int32_t best = NEG;
    int64_t ei = n - 1, ej = m - 1;
    for (int64_t j = 1; j <= m; j++)
	if (H[j] > best) {
	    best = H[j];
	    ej = j - 1;
	}
    ei = n - 1;
    for (int64_t i = 0; i < n; i++) {
	if (lastcol[i] > best) {
	    best = lastcol[i];
	    ei = i;
	    ej = m - 1;
	}
	else if (lastcol[i] == best && ej == m - 1 && i < ei) {
	    ei = i;
	}
    }
    if (score_out) *score_out = best;
    if (end_query) *end_query = ei;
    if (end_ref) *end_ref = ej;

    /* traceback: comp built backwards */
    char* comp = (char*)malloc((size_t)(n + m + 1));
    int64_t c = 0;
    if (ei + 1 == n) {
	for (int64_t t = m - 1; t > ej; t--) comp[c++] = ' ';
    }
    else {
	for (int64_t t = n - 1; t > ei; t--) comp[c++] = ' ';
    }
    int64_t i = ei, j = ej;
    int where = T_DIAG;
    while (i >= 0 || j >= 0) {
	if (i < 0) {
	    while (j >= 0) {
		comp[c++] = ' ';
		j--;
	    }
	    break;
	}
	if (j < 0) {
	    while (i >= 0) {
		comp[c++] = ' ';
		i--;
	    }
	    break;
	}
	uint8_t t = tr[i * m + j];
	if (where == T_DIAG) {
	    if (t & T_DIAG) {
		comp[c++] = (up(read[i]) == up(ref[j])) ? '|' : ' ';
		i--;
		j--;
	    }
	    else if (t & T_INS) where = T_INS;
	    else where = T_DEL;
	}
	else if (where == T_INS) {
	    comp[c++] = ' ';
	    j--;
	    where = (t & T_DIAG_E) ? T_DIAG : T_INS;
	}
	else {
	    comp[c++] = ' ';
	    i--;
	    where = (t & T_DIAG_F) ? T_DIAG : T_DEL;
	}
    }
    for (int64_t a = 0, b = c - 1; a < b; a++, b--) {
	char x = comp[a];
	comp[a] = comp[b];
	comp[b] = x;
    }
    comp[c] = 0;
    int64_t aligned = orc_aligned_windows(comp, c, e, k);
    if (comp_out) memcpy(comp_out, comp, (size_t)c + 1);
    if (comp_len) *comp_len = c;
    free(comp);
    free(tr);
    free(H);
    free(F);
    free(lastcol);
    return aligned;
}
