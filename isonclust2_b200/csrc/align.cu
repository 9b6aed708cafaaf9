// Stage 3 — batched semi-global alignment verification of (read, representative) pairs.
//
// Replaces, per pair, ParasailAlign (reference cluster.cpp:408-423 -> parasail_sg_trace_scan_16/_32),
// parasail_result_get_traceback (vendor/parasail/src/traceback_template.c:13-190) and getAlnRatio's
// window count (cluster.cpp:442-459).  The arithmetic is parasail's semi-global (all four ends free)
// affine-gap DP with match +2 / mismatch -2 (anything outside ACGT scores 0, case-insensitive,
// memory.c:336-413), gap open o in {2..5}, extend 1, its trace-bit tie-breaks (sg_trace.c:137-152), its
// end-cell rule (sg_trace_scan_avx2_256_16.c:355-395) and its traceback walk.  All DP values are int32,
// which equals parasail's 16-bit result whenever that does not saturate and its 32-bit rerun otherwise.
//
// Design (B200): nothing here is a contraction — integer max/add on the INT pipe (DPX viaddmax / vimax3).
//   icb_k_align_dp      one warp per pair.  The matrix is cut into strips of 256 query rows; lane t owns 8
//                       consecutive rows in registers and sweeps the columns one step behind lane t-1
//                       (systolic wavefront, 3 shuffles per step).  Strip boundary rows live in an
//                       L2-resident global line.  4 trace bits per cell are accumulated with funnel
//                       shifts of sign bits and stored as full 32-byte sectors, one per 8x8 cell tile.
//   traceback_warp      the same warp then walks the 8x8 tiles back from the end cell while they are still in
//                       L2 (a 4x4-tile window and its 32+32 characters are held across the lanes), feeding a
//                       k-bit shift register whose popcount gives the per-window match count.
// Roofline class: integer ALU (GCUPS); trace traffic is 0.5 B/cell.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "residency.cuh"
#include "align.cuh"

namespace {

constexpr int AL_WARPS = 4;
constexpr int AL_THREADS = AL_WARPS * 32;
constexpr int ROWS = 8;		     // rows per lane
constexpr int STRIP = 32 * ROWS;     // rows per strip
constexpr int NEG = -(1 << 29);

// A,C,G,T (either case) -> 0..3, anything else -> 4
__device__ __forceinline__ int base_code(uint32_t c)
{
    c &= 0xdfu;	 // upper-case
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}
__device__ __forceinline__ uint32_t shf_in(uint32_t acc, int signsrc)
{
    // acc = (acc << 1) | (signsrc < 0)
    return __funnelshift_l(uint32_t(signsrc), acc, 1);
}

// Traceback of one pair by a whole warp (lane-uniform control flow).  Mirrors parasail_result_get_traceback
// (traceback_template.c:72-175) and feeds getAlnRatio's window count (cluster.cpp:442-459).
__device__ icb_align_result traceback_warp(const uint8_t* __restrict__ seq, const IcbPairDesc& d, const uint4* tr, int ncg,
					   int k, int score, int ei, int ej, uint8_t* rec)
{
    const int lane = threadIdx.x & 31;
    const int n = d.n, m = d.m;
    IcbWindowCounter st;
    st.init(k, d.limit, rec);
    int i = ei, j = ej;
    // free tail (traceback_template.c:72-94)
    if (i + 1 == n) st.push_zeros(m - 1 - j);
    else st.push_zeros(n - 1 - i);
    int where = 0;	 // 0 diag, 1 ins (consumes reference), 2 del (consumes query)
    int ti0 = 1 << 28, tj0 = 1 << 28;  // window anchor (top-left tile), invalid at start
    uint4 wt = make_uint4(0, 0, 0, 0);
    uint32_t qc = 0, rc = 0;
    const int u = lane >> 1, du = u >> 2, dv = u & 3, half = lane & 1;
    while (i >= 0 || j >= 0) {
	if (i < 0) {
	    st.push_zeros(j + 1);
	    break;
	}
	if (j < 0) {
	    st.push_zeros(i + 1);
	    break;
	}
	const int ti = i >> 3, tj = j >> 3;
	if (ti < ti0 || ti > ti0 + 3 || tj < tj0 || tj > tj0 + 3) {
	    // new window with the current tile in its bottom-right corner
	    ti0 = ti - 3;
	    tj0 = tj - 3;
	    const int a = ti0 + du, b = tj0 + dv;
	    if (a >= 0 && b >= 0) wt = __ldcg(tr + (int64_t(a) * ncg + b) * 2 + half);  // written by this kernel: L2, not the read-only path
	    const int qi = ti0 * 8 + lane, rj = tj0 * 8 + lane;
	    qc = (qi >= 0 && qi < n) ? icb_up_char(seq[d.q_off + qi]) : 0u;
	    rc = (rj >= 0 && rj < m) ? icb_up_char(icb_ref_char(seq, d, rj)) : 1u;
	}
	const int uu = (ti - ti0) * 4 + (tj - tj0), r = i & 7;
	uint32_t comp = (r & 3) == 0 ? wt.x : (r & 3) == 1 ? wt.y : (r & 3) == 2 ? wt.z : wt.w;
	const uint32_t word = __shfl_sync(0xffffffffu, comp, 2 * uu + (r >> 2));
	const uint32_t nib = (word >> (4 * (7 - (j & 7)))) & 15u;
	// nib bits: 8 = H != diagonal, 4 = H != F, 2 = DIAG_E, 1 = DIAG_F
	if (where == 0) {
	    if (!(nib & 8u)) {
		uint32_t a = __shfl_sync(0xffffffffu, qc, i - ti0 * 8), b = __shfl_sync(0xffffffffu, rc, j - tj0 * 8);
		st.push(a == b ? 1u : 0u);  // traceback.c match_char, case-insensitive
		i--;
		j--;
	    }
	    else if (nib & 4u) where = 1;  // neither diagonal nor F: INS (traceback_template.c:118-120)
	    else where = 2;
	}
	else if (where == 1) {
	    st.push(0);
	    j--;
	    where = (nib & 2u) ? 0 : 1;
	}
	else {
	    st.push(0);
	    i--;
	    where = (nib & 1u) ? 0 : 2;
	}
    }
    icb_align_result r;
    r.aligned = st.pushed >= k ? st.aligned : 0;
    r.score = score;
    r.end_query = ei;
    r.end_ref = ej;
    r.comp_len = int(st.pushed);
    return r;
}

__global__ void __launch_bounds__(AL_THREADS)
icb_k_align32_dp(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, int n_pairs, uint4* __restrict__ trace,
	       int* __restrict__ lines, int k, icb_align_result* __restrict__ out, unsigned long long* __restrict__ ticket,
	       const int32_t* __restrict__ n_pairs_dev, int32_t* __restrict__ aln_out, uint8_t* __restrict__ comp_rec)
{
    const int lane = threadIdx.x & 31;
    if (n_pairs_dev) n_pairs = *n_pairs_dev;  // device-built pair queue (dloop.cu)
    for (;;) {
	int p = 0;
	if (lane == 0) p = int(atomicAdd(ticket, 1ull));
	p = __shfl_sync(0xffffffffu, p, 0);
	if (p >= n_pairs) break;
	const IcbPairDesc d = descs[p];
	const int n = d.n, m = d.m, o = d.open;
	int* bound = lines + d.line_off;	   // (H,F) of the strip's bottom row per column
	int* lastrow = bound + 2 * m;		   // H[n-1][j]
	int* lastcol = lastrow + m;		   // H[i][m-1]
	const int ncg = (m + 7) >> 3;
	uint4* tr = trace + d.trace_off * 2;
	const int nstrips = (n + STRIP - 1) / STRIP;
	const int tstar = ((n - 1) % STRIP) / ROWS, rstar = (n - 1) % ROWS;

	for (int strip = 0; strip < nstrips; strip++) {
	    const int i0 = strip * STRIP + lane * ROWS;	 // first row of this lane
	    const bool lastStrip = strip == nstrips - 1;
	    int a[ROWS];
	    uint32_t qother = 0;
#pragma unroll
	    for (int r = 0; r < ROWS; r++) {
		int c = (i0 + r < n) ? base_code(seq[d.q_off + i0 + r]) : 4;
		a[r] = c;
		qother |= uint32_t(c > 3) << r;
	    }
	    int Hl[ROWS], El[ROWS];
	    uint32_t T[ROWS];
#pragma unroll
	    for (int r = 0; r < ROWS; r++) {
		Hl[r] = 0;
		El[r] = NEG;
		T[r] = 0;
	    }
	    int hup_prev = 0;		      // H[i0-1][j-1]
	    int hb = 0, fb = NEG;	      // this lane's bottom-row (H,F) of its previous column
	    int bcode = 0;		      // this lane's previous column's reference code
	    int pre_h = 0, pre_f = NEG, pre_b = 0;  // lane l holds column s0+l of the top boundary / reference
	    const int rowgroup = strip * 32 + lane;
	    const int nsteps = m + 31;
	    for (int s = 0; s < nsteps; s++) {
		if ((s & 31) == 0) {
		    int jj = s + lane;
		    if (jj < m) {
			pre_b = base_code(icb_ref_char(seq, d, jj));
			if (strip > 0) {
			    pre_h = bound[2 * jj];
			    pre_f = bound[2 * jj + 1];
			}
		    }
		}
		// inputs of this step: from the lane above (its previous step), lane 0 from the boundary line
		int hup = __shfl_up_sync(0xffffffffu, hb, 1);
		int fup = __shfl_up_sync(0xffffffffu, fb, 1);
		int b = __shfl_up_sync(0xffffffffu, bcode, 1);
		int h0 = __shfl_sync(0xffffffffu, pre_h, s & 31);
		int f0 = __shfl_sync(0xffffffffu, pre_f, s & 31);
		int b0 = __shfl_sync(0xffffffffu, pre_b, s & 31);
		if (lane == 0) {
		    hup = h0;
		    fup = f0;
		    b = b0;
		}
		const int j = s - lane;
		if (j >= 0 && j < m) {
		    int hdiag = hup_prev;
		    hup_prev = hup;
		    const int mat = b > 3 ? 0 : 2, mis = b > 3 ? 0 : -2;
		    if (qother == 0) {
#pragma unroll
			for (int r = 0; r < ROWS; r++) {
			    int f_ext = fup - 1;
			    int F = __viaddmax_s32(hup, -o, f_ext);
			    int e_ext = El[r] - 1;
			    int E = __viaddmax_s32(Hl[r], -o, e_ext);
			    int dd = hdiag + (a[r] == b ? mat : mis);
			    int H = __vimax3_s32(dd, E, F);
			    uint32_t t = T[r];
			    t = shf_in(t, dd - H);	  // 1: H != diagonal
			    t = shf_in(t, F - H);	  // 1: H != F
			    t = shf_in(t, e_ext - E);	  // 1: E opened here (DIAG_E)
			    t = shf_in(t, f_ext - F);	  // 1: F opened here (DIAG_F)
			    T[r] = t;
			    hdiag = Hl[r];
			    Hl[r] = H;
			    El[r] = E;
			    hup = H;
			    fup = F;
			}
		    }
		    else {
#pragma unroll
			for (int r = 0; r < ROWS; r++) {
			    int f_ext = fup - 1;
			    int F = __viaddmax_s32(hup, -o, f_ext);
			    int e_ext = El[r] - 1;
			    int E = __viaddmax_s32(Hl[r], -o, e_ext);
			    int sc = (a[r] > 3) ? 0 : (a[r] == b ? mat : mis);
			    int dd = hdiag + sc;
			    int H = __vimax3_s32(dd, E, F);
			    uint32_t t = T[r];
			    t = shf_in(t, dd - H);
			    t = shf_in(t, F - H);
			    t = shf_in(t, e_ext - E);
			    t = shf_in(t, f_ext - F);
			    T[r] = t;
			    hdiag = Hl[r];
			    Hl[r] = H;
			    El[r] = E;
			    hup = H;
			    fup = F;
			}
		    }
		    hb = hup;
		    fb = fup;
		    bcode = b;
		    if (lane == 31 && !lastStrip) {
			bound[2 * j] = hb;
			bound[2 * j + 1] = fb;
		    }
		    if (lastStrip && lane == tstar) {
			int hsel = Hl[0];
#pragma unroll
			for (int r = 1; r < ROWS; r++) hsel = (r == rstar) ? Hl[r] : hsel;
			lastrow[j] = hsel;
		    }
		    if (j == m - 1) {
#pragma unroll
			for (int r = 0; r < ROWS; r++)
			    if (i0 + r < n) lastcol[i0 + r] = Hl[r];
		    }
		    if (((j & 7) == 7 || j == m - 1) && i0 < n) {
			const int sh = 4 * (7 - (j & 7));
			uint4* dst = tr + (int64_t(rowgroup) * ncg + (j >> 3)) * 2;
			dst[0] = make_uint4(T[0] << sh, T[1] << sh, T[2] << sh, T[3] << sh);
			dst[1] = make_uint4(T[4] << sh, T[5] << sh, T[6] << sh, T[7] << sh);
		    }
		}
	    }
	    __syncwarp();
	}
	// ---- end cell (sg_trace_scan_avx2_256_16.c:355-395): first maximum of the last row, then the last
	// column: strictly larger wins, the smallest row among equals; a tie with the row maximum only moves
	// the end if that maximum sits in the corner.
	__syncwarp();
	int best = NEG * 2, bj = m - 1;
	for (int j = lane; j < m; j += 32) {
	    int v = lastrow[j];
	    if (v > best) {
		best = v;
		bj = j;
	    }
	}
#pragma unroll
	for (int dlt = 16; dlt; dlt >>= 1) {
	    int ov = __shfl_xor_sync(0xffffffffu, best, dlt), oj = __shfl_xor_sync(0xffffffffu, bj, dlt);
	    if (ov > best || (ov == best && oj < bj)) {
		best = ov;
		bj = oj;
	    }
	}
	int cbest = NEG * 2, ci = n - 1;
	for (int i = lane; i < n; i += 32) {
	    int v = lastcol[i];
	    if (v > cbest) {
		cbest = v;
		ci = i;
	    }
	}
#pragma unroll
	for (int dlt = 16; dlt; dlt >>= 1) {
	    int ov = __shfl_xor_sync(0xffffffffu, cbest, dlt), oi = __shfl_xor_sync(0xffffffffu, ci, dlt);
	    if (ov > cbest || (ov == cbest && oi < ci)) {
		cbest = ov;
		ci = oi;
	    }
	}
	int score = best, ei = n - 1, ej = bj;
	if (cbest > best || (cbest == best && bj == m - 1)) {
	    score = cbest;
	    ei = ci;
	    ej = m - 1;
	}
	// ---- traceback by the same warp while the trace tiles are still in L2 (all lanes walk in lockstep; the
	// tiles of a 4x4-tile window and the 32 query / reference characters it spans are held across the lanes)
	icb_align_result res = traceback_warp(seq, d, tr, ncg, k, score, ei, ej, p == 0 ? comp_rec : nullptr);
	if (lane == 0) {
	    if (aln_out) aln_out[d.out] = res.aligned;
	    else out[d.out] = res;
	}
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side

int IcbAlignWorkspace::ensure(icb_ctx* ctx, size_t trace_need, size_t lines_need, int64_t pairs_need)
{
    // growth happens a handful of times per context; steady state allocates nothing
    if (trace_need > trace_bytes) {
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, d_trace);
	d_trace = nullptr;
	trace_bytes = 0;
	size_t want = std::max(trace_need + trace_need / 4, size_t(64) << 20);
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_trace, want));
	trace_bytes = want;
    }
    if (lines_need > lines_bytes) {
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, d_lines);
	d_lines = nullptr;
	lines_bytes = 0;
	size_t want = std::max(lines_need * 2, size_t(4) << 20);
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_lines, want));
	lines_bytes = want;
    }
    if (pairs_need > pair_cap) {
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, d_descs);
	icb_dfree(ctx, d_ends);
	icb_dfree(ctx, d_out);
	cudaFreeHost(h_descs);
	cudaFreeHost(h_out);
	cudaFreeHost(h_ends);
	d_descs = nullptr;
	d_ends = nullptr;
	d_out = nullptr;
	h_descs = nullptr;
	h_out = nullptr;
	h_ends = nullptr;
	pair_cap = 0;
	int64_t want = std::max<int64_t>(pairs_need * 2, 4096);
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_descs, sizeof(IcbPairDesc) * want));
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_ends, sizeof(IcbAlignEnd) * want));
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_out, sizeof(icb_align_result) * want));
	ICB_CUDA(ctx, cudaMallocHost((void**)&h_descs, sizeof(IcbPairDesc) * want));
	ICB_CUDA(ctx, cudaMallocHost((void**)&h_out, sizeof(icb_align_result) * want));
	ICB_CUDA(ctx, cudaMallocHost((void**)&h_ends, sizeof(IcbAlignEnd) * want));
	pair_cap = want;
    }
    if (!d_ticket) ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_ticket, 64));
    return ICB_OK;
}

void IcbAlignWorkspace::release(icb_ctx* ctx)
{
    icb_dfree(ctx, d_trace);
    icb_dfree(ctx, d_lines);
    icb_dfree(ctx, d_descs);
    icb_dfree(ctx, d_ends);
    icb_dfree(ctx, d_out);
    icb_dfree(ctx, d_ticket);
    cudaFreeHost(h_descs);
    cudaFreeHost(h_out);
    cudaFreeHost(h_ends);
    *this = IcbAlignWorkspace();
}

int icb_align32_run(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* descs, int64_t cnt,
		    int64_t tiles, int64_t line_words, int k)
{
    if (cnt <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    int rc = ws.ensure(ctx, size_t(tiles) * 32 + 64, size_t(line_words) * 4 + 64, cnt);
    if (rc) return rc;
    ICB_CUDA(ctx, cudaMemsetAsync(ws.d_ticket, 0, 8, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.d_descs, descs, sizeof(IcbPairDesc) * cnt, cudaMemcpyHostToDevice, st));
    int64_t warps = std::min<int64_t>(cnt, int64_t(ctx->sm_count) * 16);
    unsigned grid = unsigned((warps + AL_WARPS - 1) / AL_WARPS);
    ICB_LAUNCH(ctx, "align_dp32",
	       icb_k_align32_dp<<<grid, AL_THREADS, 0, st>>>(d_seq, ws.d_descs, int(cnt), reinterpret_cast<uint4*>(ws.d_trace), ws.d_lines,
							       k, ws.d_out, ws.d_ticket, nullptr, nullptr, nullptr));
    return ICB_OK;
}

// the same over a pair queue built on the device: *d_count descriptors (at most max_pairs), results (window counts) to
// aln_out[desc.out]; the caller guarantees the workspace sizes (IcbDLoopState::tiles_cap / words_cap)
int icb_align32_run_dev(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* d_descs, const int32_t* d_count,
			int max_pairs, int k, int32_t* aln_out)
{
    if (max_pairs <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    ICB_CUDA(ctx, cudaMemsetAsync(ws.d_ticket, 0, 8, st));
    int64_t warps = std::min<int64_t>(max_pairs, int64_t(ctx->sm_count) * 16);
    unsigned grid = unsigned((warps + AL_WARPS - 1) / AL_WARPS);
    ICB_LAUNCH(ctx, "align_dp32",
	       icb_k_align32_dp<<<grid, AL_THREADS, 0, st>>>(d_seq, d_descs, 0, reinterpret_cast<uint4*>(ws.d_trace), ws.d_lines, k, nullptr,
							       ws.d_ticket, d_count, aln_out, nullptr));
    return ICB_OK;
}

namespace {

constexpr int64_t LIMIT16 = 8000;			  // align16: H + 2j + 64 must stay below 2^15
constexpr int64_t TRACE_BUDGET_TILES = (int64_t(4) << 30) / 32;  // trace arena per launch of icb_align_pairs: 4 GiB (a few thousand read pairs: every SM full)

struct Launch {
    std::vector<IcbPairDesc> descs;
    std::vector<int64_t> slot;	// caller's result index of each desc
    int64_t tiles = 0, words = 0;
    int max_m = 0;
};

// run one launch on the chosen back end and scatter its results; pairs the packed kernel refused are appended to `redo`
int run_launch(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, Launch& L, bool fast, int k, icb_align_result* results,
	       std::vector<int64_t>* redo)
{
    const int64_t cnt = int64_t(L.descs.size());
    if (cnt == 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    int rc = ws.ensure(ctx, size_t(L.tiles) * 32 + 64, size_t(L.words) * 4 + 64, cnt);
    if (rc) return rc;
    memcpy(ws.h_descs, L.descs.data(), sizeof(IcbPairDesc) * cnt);
    rc = fast ? icb_align16_run(ctx, ws, d_seq, ws.h_descs, cnt, L.tiles, L.words, L.max_m, k)
	      : icb_align32_run(ctx, ws, d_seq, ws.h_descs, cnt, L.tiles, L.words, k);
    if (rc) return rc;
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.h_out, ws.d_out, sizeof(icb_align_result) * cnt, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    for (int64_t i = 0; i < cnt; i++) {
	if (fast && ws.h_out[i].comp_len < 0) redo->push_back(L.slot[i]);
	else results[L.slot[i]] = ws.h_out[i];
    }
    return ICB_OK;
}

}  // namespace

int icb_align_pairs_on(icb_ctx* ctx, icb_readset* rs, const icb_pair* pairs, int64_t n_pairs, int kmer_size,
		       icb_align_result* results)
{
    IcbRange nvtx_range("ParasailAlign batch (icb_align_pairs)");
    if (n_pairs == 0) return ICB_OK;
    if (!ctx->align_ws) ctx->align_ws = new IcbAlignWorkspace;
    IcbAlignWorkspace& ws = *ctx->align_ws;
    auto make_desc = [&](int64_t p, IcbPairDesc& d) -> int {
	const icb_pair& q = pairs[p];
	if (q.read < 0 || q.read >= rs->n_reads || q.rep < 0 || q.rep >= rs->n_reads)
	    return ctx->fail(ICB_ERR_ARG, "pair index outside the read set");
	d.q_off = rs->h_offs[q.read];
	d.r_off = rs->h_offs[q.rep];
	d.n = int32_t(rs->h_offs[q.read + 1] - rs->h_offs[q.read]);
	d.m = int32_t(rs->h_offs[q.rep + 1] - rs->h_offs[q.rep]);
	if (d.n < 1 || d.m < 1) return ctx->fail(ICB_ERR_ARG, "empty sequence in an alignment pair");
	d.rc = q.strand == -1;
	d.open = q.gap_open;
	d.limit = q.limit;
	d.pad = 0;
	return ICB_OK;
    };
    // largest matrices first (longest-processing-time order over the SMs)
    std::vector<int64_t> fast, slow;
    std::vector<IcbPairDesc> all(n_pairs);
    for (int64_t p = 0; p < n_pairs; p++) {
	int rc = make_desc(p, all[p]);
	if (rc) return rc;
	const bool ok16 = std::max(all[p].n, all[p].m) <= LIMIT16 && all[p].open >= 1 && all[p].open <= 16;
	(ok16 ? fast : slow).push_back(p);
    }
    auto by_cells = [&](int64_t x, int64_t y) { return int64_t(all[x].n) * all[x].m > int64_t(all[y].n) * all[y].m; };
    std::sort(fast.begin(), fast.end(), by_cells);
    // ---- packed 16-bit path
    {
	size_t done = 0;
	while (done < fast.size()) {
	    Launch L;
	    while (done < fast.size()) {
		IcbPairDesc d = all[fast[done]];
		d.ncg = (d.m + 71) >> 3;
		const int64_t t = int64_t((d.n + 7) >> 3) * d.ncg;
		if (!L.descs.empty() && L.tiles + t > TRACE_BUDGET_TILES) break;
		d.trace_off = L.tiles;
		d.line_off = L.words;
		d.out = int32_t(L.descs.size());
		L.tiles += t;
		L.words += d.m + 8;
		L.max_m = std::max(L.max_m, d.m);
		L.descs.push_back(d);
		L.slot.push_back(fast[done]);
		done++;
	    }
	    int rc = run_launch(ctx, ws, rs->d_seq, L, true, kmer_size, results, &slow);
	    if (rc) return rc;
	}
    }
    // ---- 32-bit path: very long sequences, characters outside ACGT
    std::sort(slow.begin(), slow.end(), by_cells);
    {
	size_t done = 0;
	while (done < slow.size()) {
	    Launch L;
	    while (done < slow.size()) {
		IcbPairDesc d = all[slow[done]];
		d.ncg = (d.m + 7) >> 3;
		const int64_t t = int64_t((d.n + 7) >> 3) * d.ncg;
		if (!L.descs.empty() && L.tiles + t > TRACE_BUDGET_TILES) break;
		d.trace_off = L.tiles;
		d.line_off = L.words;
		d.out = int32_t(L.descs.size());
		L.tiles += t;
		L.words += 3 * int64_t(d.m) + ((d.n + 7) & ~7) + 8;
		L.descs.push_back(d);
		L.slot.push_back(slow[done]);
		done++;
	    }
	    int rc = run_launch(ctx, ws, rs->d_seq, L, false, kmer_size, results, nullptr);
	    if (rc) return rc;
	}
    }
    return ICB_OK;
}

extern "C" {

int icb_align_traceback(icb_readset* rs, const icb_pair* pair, int kmer_size, icb_align_result* result, char* comp, int64_t cap)
{
    if (!rs || !pair || !result || cap < 0 || (cap > 0 && !comp)) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (kmer_size < 1 || kmer_size > 32) return ctx->fail(ICB_ERR_ARG, "kmer_size must be in [1,32]");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (pair->read < 0 || pair->read >= rs->n_reads || pair->rep < 0 || pair->rep >= rs->n_reads)
	return ctx->fail(ICB_ERR_ARG, "pair index outside the read set");
    if (!ctx->align_ws) ctx->align_ws = new IcbAlignWorkspace;
    IcbAlignWorkspace& ws = *ctx->align_ws;
    cudaStream_t st = ctx->stream;
    IcbPairDesc d{};
    d.q_off = rs->h_offs[pair->read];
    d.r_off = rs->h_offs[pair->rep];
    d.n = int32_t(rs->h_offs[pair->read + 1] - rs->h_offs[pair->read]);
    d.m = int32_t(rs->h_offs[pair->rep + 1] - rs->h_offs[pair->rep]);
    if (d.n < 1 || d.m < 1) return ctx->fail(ICB_ERR_ARG, "empty sequence in an alignment pair");
    d.rc = pair->strand == -1;
    d.open = pair->gap_open;
    d.limit = pair->limit;
    d.ncg = (d.m + 7) >> 3;
    const int64_t tiles = int64_t((d.n + 7) >> 3) * d.ncg;
    const int64_t words = 3 * int64_t(d.m) + ((d.n + 7) & ~7) + 8;
    int rc = ws.ensure(ctx, size_t(tiles) * 32 + 64, size_t(words) * 4 + 64, 1);
    if (rc) return rc;
    uint8_t* d_rec = nullptr;
    const int64_t rec_len = int64_t(d.n) + d.m + 8;
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&d_rec, size_t(rec_len)));
    ws.h_descs[0] = d;
    ICB_CUDA(ctx, cudaMemsetAsync(ws.d_ticket, 0, 8, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.d_descs, ws.h_descs, sizeof(IcbPairDesc), cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "align_dp32",
	       icb_k_align32_dp<<<1, AL_THREADS, 0, st>>>(rs->d_seq, ws.d_descs, 1, reinterpret_cast<uint4*>(ws.d_trace), ws.d_lines,
							    kmer_size, ws.d_out, ws.d_ticket, nullptr, nullptr, d_rec));
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.h_out, ws.d_out, sizeof(icb_align_result), cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    *result = ws.h_out[0];
    const int64_t len = result->comp_len;
    if (comp && len > 0 && len <= rec_len) {
	std::vector<char> tmp;
	tmp.resize(size_t(len));
	ICB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_rec, size_t(len), cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, ctx->wait());
	const int64_t w = std::min(len, cap);
	for (int64_t i = 0; i < w; i++) comp[i] = tmp[size_t(len - 1 - i)];  // recorded in traceback order
    }
    icb_dfree(ctx, d_rec);
    return ICB_OK;
}

int icb_align_pairs(icb_readset* rs, const icb_pair* pairs, int64_t n_pairs, int kmer_size, icb_align_result* results)
{
    if (!rs || n_pairs < 0 || (n_pairs > 0 && (!pairs || !results))) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (kmer_size < 1 || kmer_size > 32) return ctx->fail(ICB_ERR_ARG, "kmer_size must be in [1,32]");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    return icb_align_pairs_on(ctx, rs, pairs, n_pairs, kmer_size, results);
}

}  // extern "C"

cudaError_t icb_residency_set_align32(const IcbResidencyLog& L) { return icb_residency_set_here(L); }
