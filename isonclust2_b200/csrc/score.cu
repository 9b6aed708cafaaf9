// Stage 1a — per-read quality scoring (FP64, bit-exact with the host reference).
//
// Replaces FillQualScores / CalcQualScore / CalcErrorRate (reference qualscore.cpp:14-37, 107-136,
// 147-154).  Both quantities are strictly sequential FP64 recurrences per read (a rolling product and
// two running sums, each rounded at every step), so the parallelism is across reads: one thread per
// read, 16 quality bytes per 128-bit load, the 16 steps unrolled with compile-time byte selects.
//
// Exactness: every operation is an explicitly rounded IEEE double op (__dmul_rn/__dadd_rn/__ddiv_rn —
// never contracted into FMA) applied in the reference's order; the 10^(-Q/10) tables come from the
// host libm.  The ratio (1-p_enter)/(1-p_leave) is looked up in a 64x64 table of correctly rounded
// quotients built per CTA in shared memory (same value the division would give), with a division
// fallback for Phred > 63.
//
// Roofline class: latency/issue bound FP64 chains (3 dependent FP64 ops per base), not HBM: the
// quality bytes are read once (1 B/base).
#include "icb_common.cuh"

namespace {

constexpr int SCORE_THREADS = 256;
constexpr int RT = 64;	// ratio table covers Phred+33 chars 33..96

struct ScoreTabs {
    double one_minus[256];  // 1 - qualtab[c]   (0 -> 1.0 for c outside 33..128)
    double pn[256];	    // qualtab_nomin[c]
};

__device__ __forceinline__ uint32_t byte_of(const uint4& v, int j)
{
    uint32_t w = (j >> 2) == 0 ? v.x : (j >> 2) == 1 ? v.y : (j >> 2) == 2 ? v.z : v.w;
    return (w >> (8 * (j & 3))) & 0xffu;
}

// bytes [off, off+16) of the 32-byte concatenation lo|hi, off in [0,15] (warp-uniform)
__device__ __forceinline__ uint4 window16(const uint4& lo, const uint4& hi, int off)
{
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int sh = 8 * (off & 3);
    uint4 r;
    switch (off >> 2) {
	case 0:
	    r.x = __funnelshift_r(w[0], w[1], sh); r.y = __funnelshift_r(w[1], w[2], sh);
	    r.z = __funnelshift_r(w[2], w[3], sh); r.w = __funnelshift_r(w[3], w[4], sh);
	    break;
	case 1:
	    r.x = __funnelshift_r(w[1], w[2], sh); r.y = __funnelshift_r(w[2], w[3], sh);
	    r.z = __funnelshift_r(w[3], w[4], sh); r.w = __funnelshift_r(w[4], w[5], sh);
	    break;
	case 2:
	    r.x = __funnelshift_r(w[2], w[3], sh); r.y = __funnelshift_r(w[3], w[4], sh);
	    r.z = __funnelshift_r(w[4], w[5], sh); r.w = __funnelshift_r(w[5], w[6], sh);
	    break;
	default:
	    r.x = __funnelshift_r(w[3], w[4], sh); r.y = __funnelshift_r(w[4], w[5], sh);
	    r.z = __funnelshift_r(w[5], w[6], sh); r.w = __funnelshift_r(w[6], w[7], sh);
	    break;
    }
    return r;
}

// MODE 0: score + error rate of the raw reads (FillQualScores).  MODE 1: error rate only, lengths from
// len[] and reads selected by status[]==1 (CalcErrorRate on the compressed qualities, qualscore.cpp:78).
template <int MODE>
__global__ void __launch_bounds__(SCORE_THREADS)
icb_k_score(const uint8_t* __restrict__ qual, const int64_t* __restrict__ offs, const int32_t* __restrict__ len,
	    const int32_t* __restrict__ status, int64_t n_reads, int k, const ScoreTabs* __restrict__ tabs,
	    double* __restrict__ score, double* __restrict__ err, unsigned long long* __restrict__ counters)
{
    __shared__ double s_1m[256];
    __shared__ double s_pn[256];
    extern __shared__ double s_ratio[];	 // RT*RT, MODE 0 only
    for (int i = threadIdx.x; i < 256; i += SCORE_THREADS) {
	s_1m[i] = tabs->one_minus[i];
	s_pn[i] = tabs->pn[i];
    }
    if (MODE == 0) {
	for (int i = threadIdx.x; i < RT * RT; i += SCORE_THREADS)
	    s_ratio[i] = __ddiv_rn(tabs->one_minus[33 + i / RT], tabs->one_minus[33 + i % RT]);
    }
    __syncthreads();

    int64_t r = int64_t(blockIdx.x) * SCORE_THREADS + threadIdx.x;
    if (r >= n_reads) return;
    const int64_t start = offs[r];
    int64_t L;
    if (MODE == 0) {
	L = offs[r + 1] - start;
	if (!(uint64_t(L) > uint64_t(unsigned(2 * k)))) {  // qualscore.cpp:21,31-34
	    score[r] = -1.0;
	    err[r] = 1.0;
	    return;
	}
    }
    else {
	if (status[r] != 1) {
	    err[r] = -1.0;
	    return;
	}
	L = len[r];
    }
    const uint8_t* q = qual + start;
    double cur = 1.0, sum = 0.0, e = 0.0;
    bool bad = false;

    auto step = [&](uint32_t qc, uint32_t ql, bool main) {
	bad |= qc > 128u;
	e = __dadd_rn(e, s_pn[qc]);
	if (MODE == 0) {
	    if (main) {
		uint32_t a = qc - 33u, b = ql - 33u;
		double ratio = (a < RT && b < RT) ? s_ratio[a * RT + b] : __ddiv_rn(s_1m[qc], s_1m[ql]);
		cur = __dmul_rn(cur, ratio);
		sum = __dadd_rn(sum, cur);
	    }
	    else {
		cur = __dmul_rn(cur, s_1m[qc]);
	    }
	}
    };

    int64_t i = 0;
    if (MODE == 0) {
	// product of the first k no-error probabilities (qualscore.cpp:115-122)
	for (; i < k; i++) step(q[i], 0, false);
	sum = cur;
    }
    // scalar steps until the enter stream is 16-byte aligned
    for (; i < L && ((start + i) & 15); i++) step(q[i], MODE == 0 ? q[i - k] : 0, true);
    if (i + 16 <= L) {
	const int cb = (k + 15) >> 4;  // how many 16-byte chunks back the leave window starts (1 or 2)
	const int off = 16 * cb - k;   // byte offset of the leave window inside that chunk pair
	uint4 prev1 = make_uint4(0, 0, 0, 0), prev2 = prev1;
	if (MODE == 0) {
	    prev1 = *reinterpret_cast<const uint4*>(q + i - 16);
	    if (cb == 2) prev2 = *reinterpret_cast<const uint4*>(q + i - 32);
	}
	for (; i + 16 <= L; i += 16) {
	    const uint4 c = *reinterpret_cast<const uint4*>(q + i);
	    uint4 lv = c;
	    if (MODE == 0) lv = (cb == 1) ? window16(prev1, c, off) : window16(prev2, prev1, off);
#pragma unroll
	    for (int j = 0; j < 16; j++) step(byte_of(c, j), byte_of(lv, j), true);
	    prev2 = prev1;
	    prev1 = c;
	}
    }
    for (; i < L; i++) step(q[i], MODE == 0 ? q[i - k] : 0, true);

    if (bad) atomicOr(&counters[1], 2ull);
    err[r] = __ddiv_rn(e, double(uint64_t(L)));
    if (MODE == 0) score[r] = (sum <= 0) ? -1.0 : sum;
}

// host-side table staging shared by both launchers
int stage_tabs(icb_ctx* ctx, const double* tab, const double* tab_nomin, ScoreTabs** d_out)
{
    ScoreTabs h;
    for (int c = 0; c < 256; c++) {
	double t = (tab && c <= 128) ? tab[c] : 0.0;
	h.one_minus[c] = 1.0 - t;  // the reference's own subtraction (qualscore.cpp:116,126), host IEEE
	h.pn[c] = (tab_nomin && c <= 128) ? tab_nomin[c] : 0.0;
    }
    ScoreTabs* d = nullptr;
    ICB_CUDA(ctx, cudaMallocAsync((void**)&d, sizeof(ScoreTabs), ctx->stream));
    // pageable source: the runtime stages it before returning, so the stack copy is safe
    ICB_CUDA(ctx, cudaMemcpyAsync(d, &h, sizeof(ScoreTabs), cudaMemcpyHostToDevice, ctx->stream));
    *d_out = d;
    return ICB_OK;
}

}  // namespace

int icb_launch_score(icb_readset* rs, int k, const double* tab, const double* tab_nomin)
{
    icb_ctx* ctx = rs->ctx;
    ScoreTabs* d_tabs = nullptr;
    int rc = stage_tabs(ctx, tab, tab_nomin, &d_tabs);
    if (rc) return rc;
    if (rs->n_reads > 0) {
	size_t smem = sizeof(double) * RT * RT;
	ICB_CUDA(ctx, cudaFuncSetAttribute(icb_k_score<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	unsigned grid = unsigned((rs->n_reads + SCORE_THREADS - 1) / SCORE_THREADS);
	ICB_LAUNCH(ctx, "score_reads",
		   icb_k_score<0><<<grid, SCORE_THREADS, smem, ctx->stream>>>(rs->d_qual, rs->d_offs, nullptr, nullptr,
									       rs->n_reads, k, d_tabs, rs->d_score,
									       rs->d_err, rs->d_counters));
    }
    ICB_CUDA(ctx, cudaFreeAsync(d_tabs, ctx->stream));
    return ICB_OK;
}

int icb_launch_error_rate(icb_readset* rs, const uint8_t* d_qual, const int64_t* d_offs, const int32_t* d_len,
			  const int32_t* d_status, const double* tab_nomin, double* d_out)
{
    icb_ctx* ctx = rs->ctx;
    ScoreTabs* d_tabs = nullptr;
    int rc = stage_tabs(ctx, nullptr, tab_nomin, &d_tabs);
    if (rc) return rc;
    if (rs->n_reads > 0) {
	unsigned grid = unsigned((rs->n_reads + SCORE_THREADS - 1) / SCORE_THREADS);
	ICB_LAUNCH(ctx, "hpc_error_rate",
		   icb_k_score<1><<<grid, SCORE_THREADS, 0, ctx->stream>>>(d_qual, d_offs, d_len, d_status, rs->n_reads, 0,
									    d_tabs, nullptr, d_out, rs->d_counters));
    }
    ICB_CUDA(ctx, cudaFreeAsync(d_tabs, ctx->stream));
    return ICB_OK;
}

extern "C" {

int icb_score_reads(icb_readset* rs, int kmer_size, const double qualtab[129], const double qualtab_nomin[129])
{
    IcbRange nvtx_range("FillQualScores (icb_score_reads)");
    if (!rs || !qualtab || !qualtab_nomin) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (kmer_size < 1 || kmer_size > 31) return ctx->fail(ICB_ERR_ARG, "kmer_size must be in [1,31]");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    ICB_CUDA(ctx, cudaMemsetAsync(rs->d_counters, 0, sizeof(unsigned long long) * 8, ctx->stream));
    int rc = icb_launch_score(rs, kmer_size, qualtab, qualtab_nomin);
    if (rc) return rc;
    rs->scored = true;
    return ICB_OK;
}

int icb_scores_download(icb_readset* rs, double* score, double* err)
{
    if (!rs) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (!rs->scored) return ctx->fail(ICB_ERR_STATE, "icb_scores_download before icb_score_reads");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long flags[8];
    ICB_CUDA(ctx, cudaMemcpyAsync(flags, rs->d_counters, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    if (score) ICB_CUDA(ctx, cudaMemcpyAsync(score, rs->d_score, sizeof(double) * rs->n_reads, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) ICB_CUDA(ctx, cudaMemcpyAsync(err, rs->d_err, sizeof(double) * rs->n_reads, cudaMemcpyDeviceToHost, ctx->stream));
    ICB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[1] & 2ull) return ctx->fail(ICB_ERR_ARG, "quality character above 128 (the reference's QualTab::at throws)");
    return ICB_OK;
}

}  // extern "C"
