/* Deterministic synthetic ONT-like cDNA read generator (data tooling for tests
 * and bench.py; not part of the hot path).  Spec: SURVEY.md App. D2 / §8(d):
 * G = max(1, n/reads_per_gene) genes of length U[600,2000]; a read is a copy of
 * a random gene with per-base error probability e (chosen per read from
 * err_set, or fixed) split evenly into deletion / insertion-before /
 * substitution, reverse-complemented with probability 1/2, Phred chars
 * chr(33 + clamp(int(N(int(-10 log10 e), 3)), 2, 40)).  ACGT only (the
 * reference aborts on anything else, util.cpp:31-33).
 *
 * Two passes so the caller can allocate: synth_sizes() then synth_fill().
 * The generator state is a pure function of (seed, read index) so reads can
 * be produced in parallel or in shards and stay identical. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline double gauss(uint64_t *s)
{
    double u1 = u01(s), u2 = u01(s);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};

typedef struct {
    int n_genes;
    int64_t *goffs; /* n_genes+1 */
    char *gseq;
} genes_t;

static genes_t make_genes(uint64_t seed, int n_genes, int min_len, int max_len)
{
    genes_t g;
    g.n_genes = n_genes;
    g.goffs = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_genes + 1));
    uint64_t s = seed ^ 0xA5A5A5A55A5A5A5Aull;
    int64_t tot = 0;
    for (int i = 0; i < n_genes; i++) {
	g.goffs[i] = tot;
	tot += min_len + (int64_t)(splitmix(&s) % (uint64_t)(max_len - min_len + 1));
    }
    g.goffs[n_genes] = tot;
    g.gseq = (char *)malloc((size_t)tot);
    for (int64_t i = 0; i < tot; i++) g.gseq[i] = BASES[splitmix(&s) & 3];
    return g;
}

/* generate read i into out_seq/out_qual (may be NULL to only measure) */
static int64_t gen_read(const genes_t *g, uint64_t seed, int64_t i, const double *err_set, int n_err,
			char *out_seq, char *out_qual)
{
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + (uint64_t)i * 0xD1B54A32D192ED03ull + 12345;
    splitmix(&s);
    int gi = (int)(splitmix(&s) % (uint64_t)g->n_genes);
    double e = err_set[splitmix(&s) % (uint64_t)n_err];
    int rc = (int)(splitmix(&s) & 1);
    int qmean = (int)(-10.0 * log10(e));
    const char *src = g->gseq + g->goffs[gi];
    int64_t glen = g->goffs[gi + 1] - g->goffs[gi];
    int64_t n = 0;
    for (int64_t p = 0; p < glen; p++) {
	/* walk the gene in read orientation */
	char b = rc ? src[glen - 1 - p] : src[p];
	if (rc) b = (b == 'A') ? 'T' : (b == 'C') ? 'G' : (b == 'G') ? 'C' : 'A';
	double r = u01(&s);
	int emit = 1;
	if (r < e) {
	    double t = r / e;
	    if (t < 1.0 / 3.0) {
		emit = 0; /* deletion */
	    }
	    else if (t < 2.0 / 3.0) { /* insertion before */
		if (out_seq) out_seq[n] = BASES[splitmix(&s) & 3];
		else splitmix(&s);
		n++;
	    }
	    else { /* substitution */
		int k = (int)(splitmix(&s) % 3);
		int bi = (b == 'A') ? 0 : (b == 'C') ? 1 : (b == 'G') ? 2 : 3;
		b = BASES[(bi + 1 + k) & 3];
	    }
	}
	if (emit) {
	    if (out_seq) out_seq[n] = b;
	    n++;
	}
    }
    if (n == 0) { /* never emit an empty read */
	if (out_seq) out_seq[0] = 'A';
	n = 1;
    }
    if (out_qual) {
	uint64_t q = s ^ 0x1234567ull;
	for (int64_t p = 0; p < n; p++) {
	    int v = (int)(qmean + 3.0 * gauss(&q));
	    if (v < 2) v = 2;
	    if (v > 40) v = 40;
	    out_qual[p] = (char)(33 + v);
	}
    }
    return n;
}

/* offs[n_reads+1] filled; returns total bases.  n_genes is explicit so that a prefix of a large read set (the first
 * reads of configs[1]'s 1 M reads over 40 000 genes) can be produced without generating the rest. */
int64_t synth_sizes_g(uint64_t seed, int64_t n_reads, int n_genes, int min_len, int max_len, const double *err_set, int n_err,
		      int64_t *offs)
{
    int ng = n_genes < 1 ? 1 : n_genes;
    genes_t g = make_genes(seed, ng, min_len, max_len);
#pragma omp parallel for schedule(static, 1024)
    for (int64_t i = 0; i < n_reads; i++) offs[i + 1] = gen_read(&g, seed, i, err_set, n_err, NULL, NULL);
    int64_t tot = 0;
    offs[0] = 0;
    for (int64_t i = 0; i < n_reads; i++) {
	tot += offs[i + 1];
	offs[i + 1] = tot;
    }
    offs[n_reads] = tot;
    free(g.goffs);
    free(g.gseq);
    return tot;
}

void synth_fill_g(uint64_t seed, int64_t n_reads, int n_genes, int min_len, int max_len, const double *err_set, int n_err,
		  const int64_t *offs, char *seq, char *qual)
{
    int ng = n_genes < 1 ? 1 : n_genes;
    genes_t g = make_genes(seed, ng, min_len, max_len);
#pragma omp parallel for schedule(static, 1024)
    for (int64_t i = 0; i < n_reads; i++) gen_read(&g, seed, i, err_set, n_err, seq + offs[i], qual + offs[i]);
    free(g.goffs);
    free(g.gseq);
}

int64_t synth_sizes(uint64_t seed, int64_t n_reads, int reads_per_gene, int min_len, int max_len, const double *err_set, int n_err,
		    int64_t *offs)
{
    return synth_sizes_g(seed, n_reads, (int)(n_reads / reads_per_gene), min_len, max_len, err_set, n_err, offs);
}

void synth_fill(uint64_t seed, int64_t n_reads, int reads_per_gene, int min_len, int max_len, const double *err_set, int n_err,
		const int64_t *offs, char *seq, char *qual)
{
    synth_fill_g(seed, n_reads, (int)(n_reads / reads_per_gene), min_len, max_len, err_set, n_err, offs, seq, qual);
}
