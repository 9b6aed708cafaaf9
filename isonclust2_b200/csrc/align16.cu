// Stage 3, fast path — semi-global alignment of one (read, representative) pair by a whole CTA, in packed
// 16-bit arithmetic, followed by a warp-per-pair traceback kernel.
//
// Replaces ParasailAlign (reference cluster.cpp:408-423 -> parasail_sg_trace_scan_16), parasail_result_get_traceback
// (vendor/parasail/src/traceback_template.c:13-190) and getAlnRatio's window count (cluster.cpp:442-459); the
// recurrence, trace-bit tie-breaks and end-cell rule are the ones align.cu documents (sg_trace.c:137-152,
// sg_trace_scan_avx2_256_16.c:355-395).
//
// icb_k_align16_dp — a systolic wavefront over the CTA's 256 threads.  Thread g owns 8 consecutive query rows as two
//   packed 16-bit halves: low half rows 8g..8g+3 at column j, high half rows 8g+4..8g+7 at column j-1 (the high
//   rows need the low rows' result of the same column, so they trail by one step).  A step is therefore four
//   dependent packed cell updates = 8 cells, each update VIADDMNMX.S16x2 x2 (E, F), PRMT (score look-up),
//   VIMNMX3.S16x2 (H) plus the four trace bits.  Thread g trails thread g-1 by two columns; inside a warp the
//   hand-over is one shuffle per value, between warps a shared-memory ring read 16 steps after it is written (warps are staggered
//   by 80 steps and meet at a barrier every 16 steps).
//   Scores are stored biased, H~ = H + 2j + 64: the diagonal increment becomes {0,+4} (one byte-permute from a
//   one-hot word per reference base), every packed value stays positive (plain 32-bit adds never carry between the
//   halves) and E~ never decreases along a row.  Comparisons are between quantities of the same cell, so the trace
//   bits are those of the unbiased recurrence.  The bias bounds the operands: max(n, m) <= 8000.
//   Reads longer than 2048 rows take several passes; the bottom row of a pass goes through a global line.
//   Trace: 4 bits per cell, one 32-byte tile per (8 rows x 8 steps) written as two 16-byte stores, addressed in
//   the kernel's own skewed coordinates (u = j + 2*(rowblock mod 32) + half), which the traceback undoes.
// icb_k_align16_trace — one warp per pair walks the tiles back from the end cell (a 4x4-tile window is held across
//   the lanes) and feeds the k-bit window register.
// Roofline class: integer ALU (GCUPS); trace traffic 0.5 B/cell written once, ~64 B per traceback step read.
#include <algorithm>

#include "align.cuh"

namespace {

constexpr int A16_THREADS = 256;
constexpr int A16_WARPS = A16_THREADS / 32;
constexpr int A16_DELAY = 16;			   // steps between a warp publishing its bottom row and the next warp reading it:
						   // one CTA barrier per A16_DELAY steps (two 8-step tiles)
constexpr int A16_LAG = 64 + A16_DELAY;		   // warp-to-warp stagger in steps: thread 31 holds column s - 64 after local step s - 1
constexpr int A16_PASS_ROWS = A16_THREADS * 8;	   // 2048 query rows per pass
constexpr int A16_BIAS = 64;
constexpr int A16_RING = 64;			   // ring slots per warp (steps)
constexpr uint32_t ONE2 = 0x00010001u;
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return (uint32_t(lo) & 0xffffu) | (uint32_t(hi) << 16); }
// A,C,G,T (either case) -> 0..3, anything else -> 4
__device__ __forceinline__ int code5(uint32_t c)
{
    c &= 0xdfu;
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

struct Lane16 {
    uint32_t H[4], E[4], T[4], SEL[4];
    uint32_t hb, fb, hup_prev, Wprev;
    int rbest, rbj;  // best of the last row (owner lane only): first maximum
    int cbest, ci;   // best of the last column over this lane's rows: smallest row among equals
};

struct Pass16 {
    const uint32_t* W;	 // shared: one-hot score word per reference column, W[m] = 0
    uint2* ring;	 // shared: [warp][A16_RING] hand-over of (H, F) between warps
    uint32_t* line;	 // global: bottom row of the previous pass / of this pass
    int m, n;
    int lane, warp, pass;
    int row0;		 // first query row of this lane in this pass
    uint32_t NEGO, C2O;	 // packed -open, packed 2-open
    int top;		 // where thread 0 of the warp takes the row above from: 0 the ring, 1 the matrix edge, 2 the line
    bool write_line;	 // this thread is the last of a pass that is followed by another
    bool owner;		 // this thread holds the last query row
    int owner_rr;
};

// EDGE: some thread of the warp is outside its column range or at a boundary column in this block.
// SPECIAL: the warp is the first of the CTA (its thread 0 takes the row above from the matrix edge / the previous pass)
// or the last of a pass that is followed by another (its thread 31 writes the line).
template <int Q, bool EDGE>
__device__ __forceinline__ void step16(Lane16& S, const Pass16& P, int sl, const uint32_t* __restrict__ Wp, const uint2* __restrict__ rd,
				       uint2* __restrict__ wr)
{
    // Wp = W + sl0 - 2*lane, rd / wr = this block's 8 ring slots to read / write (all offsets below are compile-time)
    const int lane = P.lane;
    const int jl = sl - 2 * lane;
    // thread 31 publishes the values it holds after the PREVIOUS step; the next warp reads them A16_DELAY steps later
    if (lane == 31) wr[Q] = make_uint2(S.hb, S.fb);
    uint32_t x = __shfl_up_sync(FULL, S.hb, 1);
    uint32_t y = __shfl_up_sync(FULL, S.fb, 1);
    {
	uint32_t x0, y0;
	if (P.top == 0) {  // from the warp above
	    const uint2 v = rd[Q];
	    x0 = v.x;
	    y0 = v.y;
	}
	else if (P.top == 1 || jl < 0 || jl >= P.m) {  // first row of the matrix: H[-1][j] = 0, F[-1][j] = -inf
	    x0 = uint32_t(2 * jl + A16_BIAS) << 16;
	    y0 = 0x00010000u;
	}
	else {	// bottom row of the previous pass
	    const uint32_t v = __ldcg(P.line + jl);
	    x0 = v << 16;
	    y0 = v & 0xffff0000u;
	}
	if (lane == 0) {
	    x = x0;
	    y = y0;
	}
    }
    uint32_t hup = prmt(x, S.hb, 0x5432);  // low half <- the rows above (previous thread's high half), high half <- own low half
    uint32_t fup = prmt(y, S.fb, 0x5432);
    uint32_t hd = S.hup_prev;
    S.hup_prev = hup;
    const bool active = EDGE ? (jl >= 0 && jl <= P.m) : true;
    if (active) {
	const uint32_t Wlo = Wp[Q];
	const uint32_t Whi = S.Wprev;
	S.Wprev = Wlo;
#pragma unroll
	for (int r = 0; r < 4; r++) {
	    const uint32_t sc = prmt(Wlo, Whi, S.SEL[r]);
	    const uint32_t D = hd + sc;
	    const uint32_t fx = fup - ONE2;
	    const uint32_t F = __viaddmax_s16x2(hup, P.NEGO, fx);
	    const uint32_t ex = S.E[r] + ONE2;
	    const uint32_t En = __viaddmax_s16x2(S.H[r], P.C2O, ex);
	    const uint32_t Hn = __vimax3_s16x2(D, En, F);
	    // trace nibble: 8 = H != diagonal, 4 = H != F, 2 = E opened here, 1 = F opened here
	    const uint32_t bu = __vmins2(Hn - D, ONE2);
	    const uint32_t bv = __vmins2(Hn - F, ONE2);
	    const uint32_t bx = __vmins2(En - ex, ONE2);
	    const uint32_t by = __vmins2(F - fx, ONE2);
	    const uint32_t nib = ((bu * 2u + bv) * 2u + bx) * 2u + by;
	    S.T[r] = (Q == 0 && !EDGE) ? nib : S.T[r] * 16u + nib;  // 4 steps x 4 bits per half, first step in the top nibble
	    hd = S.H[r];
	    S.H[r] = Hn;
	    S.E[r] = En;
	    hup = Hn;
	    fup = F;
	}
	S.hb = hup;
	S.fb = fup;
	if (EDGE) {
	    if (jl == 0) {
		// the high half just "computed" column -1: put the left boundary back (H[i][-1] = 0, E = -inf)
#pragma unroll
		for (int r = 0; r < 4; r++) {
		    S.H[r] = (S.H[r] & 0xffffu) | (uint32_t(A16_BIAS - 2) << 16);
		    S.E[r] = (S.E[r] & 0xffffu) | 0x00010000u;
		}
		S.hb = S.H[3];
	    }
	    if (jl == P.m - 1) {
		const int unb = 2 * (P.m - 1) + A16_BIAS;
#pragma unroll
		for (int r = 0; r < 4; r++) {
		    const int v = int(S.H[r] & 0xffffu) - unb;
		    if (P.row0 + r < P.n && v > S.cbest) {
			S.cbest = v;
			S.ci = P.row0 + r;
		    }
		}
	    }
	    if (jl == P.m) {
		const int unb = 2 * (P.m - 1) + A16_BIAS;
#pragma unroll
		for (int r = 0; r < 4; r++) {
		    const int v = int(S.H[r] >> 16) - unb;
		    if (P.row0 + 4 + r < P.n && v > S.cbest) {
			S.cbest = v;
			S.ci = P.row0 + 4 + r;
		    }
		}
	    }
	}
	if (P.owner) {
	    const int r = P.owner_rr & 3;
	    const uint32_t hv = r == 0 ? S.H[0] : r == 1 ? S.H[1] : r == 2 ? S.H[2] : S.H[3];
	    const bool hi = P.owner_rr >= 4;
	    const int col = hi ? jl - 1 : jl;
	    const int v = int(hi ? (hv >> 16) : (hv & 0xffffu)) - 2 * col - A16_BIAS;
	    if (col >= 0 && col < P.m && v > S.rbest) {
		S.rbest = v;
		S.rbj = col;
	    }
	}
	if (P.write_line) {  // last thread of a pass that is not the last: its bottom row feeds the next pass
	    const int jh = jl - 1;
	    if (jh >= 0 && jh < P.m) __stcg(P.line + jh, (S.hb >> 16) | (S.fb & 0xffff0000u));
	}
    }
    else if (EDGE) {
	// keep the nibble positions of the half tile aligned with the step index
#pragma unroll
	for (int r = 0; r < 4; r++) S.T[r] *= 16u;
    }
}

// four steps (sl0 is a multiple of 4) and the 16-byte half tile they fill
template <bool EDGE>
__device__ __forceinline__ void quad16(Lane16& S, const Pass16& P, int sl0, uint4* __restrict__ half_tile, bool store)
{
    const int sg0 = sl0 + A16_LAG * P.warp;  // global step of the quad start
    const uint32_t* Wp = P.W + (sl0 - 2 * P.lane);
    const uint2* rd = P.ring + (P.warp > 0 ? (P.warp - 1) * A16_RING : 0) + (sg0 & (A16_RING - 1));
    uint2* wr = P.ring + P.warp * A16_RING + ((sg0 + A16_DELAY) & (A16_RING - 1));
    if (EDGE) {
#pragma unroll
	for (int r = 0; r < 4; r++) S.T[r] = 0;
    }
    step16<0, EDGE>(S, P, sl0 + 0, Wp, rd, wr);
    step16<1, EDGE>(S, P, sl0 + 1, Wp, rd, wr);
    step16<2, EDGE>(S, P, sl0 + 2, Wp, rd, wr);
    step16<3, EDGE>(S, P, sl0 + 3, Wp, rd, wr);
    if (store) *half_tile = make_uint4(S.T[0], S.T[1], S.T[2], S.T[3]);
}

__global__ void __launch_bounds__(A16_THREADS, 4)
icb_k_align16_dp(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, uint4* __restrict__ trace,
		 uint32_t* __restrict__ lines, IcbAlignEnd* __restrict__ ends, const int32_t* __restrict__ n_pairs_dev)
{
    if (n_pairs_dev && int(blockIdx.x) >= *n_pairs_dev) return;  // device-built pair queue (dloop.cu): the grid is its capacity
    extern __shared__ uint32_t smW[];  // m + 1 words
    __shared__ uint2 s_ring[A16_WARPS * A16_RING];
    __shared__ int s_red[A16_WARPS * 2];
    __shared__ int s_row[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const IcbPairDesc d = descs[blockIdx.x];
    const int n = d.n, m = d.m;

    // ---- stage the reference as one-hot score words; refuse anything outside ACGT (parasail scores it 0)
    int bad = 0;
    for (int j = tid; j < m; j += A16_THREADS) {
	const int c = code5(icb_ref_char(seq, d, j));
	bad |= c > 3;
	smW[j] = 4u << (8 * (c & 3));
    }
    if (tid == 0) smW[m] = 0;
    for (int i = tid; i < n; i += A16_THREADS) bad |= code5(seq[d.q_off + i]) > 3;
    if (__syncthreads_or(bad)) {
	if (tid == 0) {
	    IcbAlignEnd e;
	    e.score = 0;
	    e.ei = e.ej = 0;
	    e.status = 1;
	    ends[blockIdx.x] = e;
	}
	return;
    }

    Lane16 S;
    S.rbest = -(1 << 30);
    S.rbj = m - 1;
    S.cbest = -(1 << 30);
    S.ci = n - 1;
    Pass16 P;
    P.W = smW;
    P.ring = s_ring;
    P.line = lines + d.line_off;
    P.m = m;
    P.n = n;
    P.lane = lane;
    P.warp = warp;
    P.NEGO = pack2(-d.open, -d.open);
    P.C2O = pack2(2 - d.open, 2 - d.open);
    const int npass = (n + A16_PASS_ROWS - 1) / A16_PASS_ROWS;
    const int nloc = (m + 64 + 7) & ~7;	 // local steps: thread 31 finishes column m at step m + 62 and publishes it at m + 63
    const int last = n - 1;
    uint4* tr = trace + d.trace_off * 2;

    for (int pass = 0; pass < npass; pass++) {
	const int rows_left = n - pass * A16_PASS_ROWS;
	const int nw = min(A16_WARPS, (rows_left + 255) >> 8);
	P.pass = pass;
	P.row0 = pass * A16_PASS_ROWS + tid * 8;
	P.top = warp > 0 ? 0 : (pass == 0 ? 1 : 2);
	P.write_line = pass + 1 < npass && tid == A16_THREADS - 1;
	P.owner = (last / A16_PASS_ROWS == pass) && (((last % A16_PASS_ROWS) >> 3) == tid);
	P.owner_rr = last & 7;
	// per-thread constants: the score selectors of its 8 rows
#pragma unroll
	for (int r = 0; r < 4; r++) {
	    const int ilo = P.row0 + r, ihi = P.row0 + 4 + r;
	    const int alo = ilo < n ? (code5(seq[d.q_off + ilo]) & 3) : 0;
	    const int ahi = ihi < n ? (code5(seq[d.q_off + ihi]) & 3) : 0;
	    S.SEL[r] = uint32_t(alo) | 0x80u | (uint32_t(4 + ahi) << 8) | 0x8000u;
	    S.H[r] = pack2(A16_BIAS - 2, A16_BIAS - 2);	 // H[i][-1] = 0 at column -1
	    S.E[r] = ONE2;				 // E[i][-1] = -inf
	    S.T[r] = 0;
	}
	S.hb = pack2(A16_BIAS - 2, A16_BIAS - 2);
	S.fb = ONE2;
	S.hup_prev = pack2(A16_BIAS - 2, A16_BIAS - 2);
	S.Wprev = 0;
	const int rowblock = pass * A16_THREADS + tid;
	const bool rows_valid = P.row0 < n;
	const int nblk = (A16_LAG * (nw - 1) + nloc) >> 3;
	for (int blk = 0; blk < nblk; blk++) {
	    const int sl0 = blk * 8 - A16_LAG * warp;
	    if (warp < nw && sl0 >= 0 && sl0 < nloc) {
		const bool edge = sl0 < 63 || sl0 + 7 > m - 2;
		// the 8x8-step tile, filled half by half
		const bool store = rows_valid && sl0 + 7 >= 2 * lane && sl0 <= 2 * lane + m;
		uint4* dst = tr + (int64_t(rowblock) * d.ncg + (sl0 >> 3)) * 2;
		if (edge) {
#pragma unroll 1
		    for (int hq = 0; hq < 2; hq++) quad16<true>(S, P, sl0 + 4 * hq, dst + hq, store);
		}
		else {
#pragma unroll 1
		    for (int hq = 0; hq < 2; hq++) quad16<false>(S, P, sl0 + 4 * hq, dst + hq, store);
		}
	    }
	    // a slot written in tile b is read in tile b + A16_DELAY / 8: one barrier per A16_DELAY / 8 tiles keeps them apart
	    if ((blk + 1) % (A16_DELAY / 8) == 0) __syncthreads();
	}
	if (nblk % (A16_DELAY / 8)) __syncthreads();
    }
    // ---- end cell (sg_trace_scan_avx2_256_16.c:355-395): first maximum of the last row, then the last column:
    // strictly larger wins, the smallest row among equals; a tie with the row maximum only moves the end if
    // that maximum sits in the corner.
    {
	const int ow = (last % A16_PASS_ROWS) >> 8, ol = ((last % A16_PASS_ROWS) >> 3) & 31;
	if (warp == ow && lane == ol) {
	    s_row[0] = S.rbest;
	    s_row[1] = S.rbj;
	}
	int cb = S.cbest, ci = S.ci;
#pragma unroll
	for (int dlt = 16; dlt; dlt >>= 1) {
	    const int ov = __shfl_xor_sync(FULL, cb, dlt), oi = __shfl_xor_sync(FULL, ci, dlt);
	    if (ov > cb || (ov == cb && oi < ci)) {
		cb = ov;
		ci = oi;
	    }
	}
	if (lane == 0) {
	    s_red[2 * warp] = cb;
	    s_red[2 * warp + 1] = ci;
	}
	__syncthreads();
	if (tid == 0) {
	    for (int w = 1; w < A16_WARPS; w++) {
		const int ov = s_red[2 * w], oi = s_red[2 * w + 1];
		if (ov > cb || (ov == cb && oi < ci)) {
		    cb = ov;
		    ci = oi;
		}
	    }
	    const int best = s_row[0], bj = s_row[1];
	    IcbAlignEnd e;
	    e.score = best;
	    e.ei = n - 1;
	    e.ej = bj;
	    if (cb > best || (cb == best && bj == m - 1)) {
		e.score = cb;
		e.ei = ci;
		e.ej = m - 1;
	    }
	    e.status = 0;
	    ends[blockIdx.x] = e;
	}
    }
}

// ------------------------------------------------------------------------------------------------ traceback
// Mirrors parasail_result_get_traceback (traceback_template.c:72-175) and feeds getAlnRatio's window count
// (cluster.cpp:442-459).  One warp per pair, lane-uniform control flow.
constexpr int TB_WARPS = 4;

__global__ void __launch_bounds__(TB_WARPS * 32)
icb_k_align16_trace(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, int n_pairs,
		    const uint4* __restrict__ trace, const IcbAlignEnd* __restrict__ ends, int k,
		    icb_align_result* __restrict__ out, const int32_t* __restrict__ n_pairs_dev, int32_t* __restrict__ aln_out)
{
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * TB_WARPS + (threadIdx.x >> 5);
    if (n_pairs_dev) n_pairs = *n_pairs_dev;
    if (p >= n_pairs) return;
    const IcbPairDesc d = descs[p];
    const IcbAlignEnd en = ends[p];
    if (en.status != 0) {
	if (aln_out) {
	    if (lane == 0) aln_out[d.out] = -3;  // DL_ALN_REDO32: a character outside ACGT, the 32-bit kernel takes the pair
	    return;
	}
	if (lane == 0) {
	    icb_align_result r;
	    r.aligned = r.score = r.end_query = r.end_ref = 0;
	    r.comp_len = -1;  // marks "not computed here"
	    out[d.out] = r;
	}
	return;
    }
    const int n = d.n, m = d.m, ncg = d.ncg;
    const uint4* tr = trace + d.trace_off * 2;
    IcbWindowCounter st;
    st.init(k, d.limit);
    int i = en.ei, j = en.ej;
    // free tail (traceback_template.c:72-94)
    if (i + 1 == n) st.push_zeros(m - 1 - j);
    else st.push_zeros(n - 1 - i);
    int where = 0;  // 0 diag, 1 ins (consumes reference), 2 del (consumes query)
    // window: row blocks g0..g0+3, for row block g0+a the tiles ub[a]..ub[a]+3; lane = (a*4 + b)*2 + half-tile
    int g0 = 1 << 28;
    int ub0 = 0, ub1 = 0, ub2 = 0, ub3 = 0;
    uint4 wt = make_uint4(0, 0, 0, 0);
    uint32_t qc = 0, rc = 0;
    int qbase = 0, cbase = 1 << 28;
    const int la = lane >> 3, lb = (lane >> 1) & 3, lh = lane & 1;
    bool pend = false;
    uint32_t pa = 0, pb = 0;
    while (i >= 0 || j >= 0) {
	if (pend) {
	    st.push(pa == pb ? 1u : 0u);  // traceback.c match_char, case-insensitive
	    pend = false;
	}
	if (i < 0) {
	    st.push_zeros(j + 1);
	    break;
	}
	if (j < 0) {
	    st.push_zeros(i + 1);
	    break;
	}
	const int gg = i >> 3, rr = i & 7, half = rr >> 2, r = rr & 3;
	const int u = j + 2 * (gg & 31) + half, ug = u >> 3, q = u & 7;
	int a = gg - g0;
	int ub = a == 0 ? ub0 : a == 1 ? ub1 : a == 2 ? ub2 : ub3;
	if (a < 0 || a > 3 || ug < ub || ug > ub + 3 || j < cbase) {
	    // new window with the current cell in its bottom-right corner
	    g0 = gg - 3;
	    ub0 = ((j + 2 * ((g0 + 0) & 31) + 1) >> 3) - 3;
	    ub1 = ((j + 2 * ((g0 + 1) & 31) + 1) >> 3) - 3;
	    ub2 = ((j + 2 * ((g0 + 2) & 31) + 1) >> 3) - 3;
	    ub3 = ((j + 2 * ((g0 + 3) & 31) + 1) >> 3) - 3;
	    const int myg = g0 + la;
	    const int myu = (la == 0 ? ub0 : la == 1 ? ub1 : la == 2 ? ub2 : ub3) + lb;
	    if (myg >= 0 && myu >= 0 && myu < ncg) wt = __ldcg(tr + (int64_t(myg) * ncg + myu) * 2 + lh);
	    qbase = g0 * 8;
	    cbase = j - 31;
	    const int qi = qbase + lane, rj = cbase + lane;
	    qc = (qi >= 0 && qi < n) ? icb_up_char(seq[d.q_off + qi]) : 0u;
	    rc = (rj >= 0 && rj < m) ? icb_up_char(icb_ref_char(seq, d, rj)) : 1u;
	    a = 3;
	    ub = ub3;
	}
	const uint32_t comp = r == 0 ? wt.x : r == 1 ? wt.y : r == 2 ? wt.z : wt.w;
	const uint32_t word = __shfl_sync(FULL, comp, ((a << 2) + (ug - ub)) * 2 + (q >> 2));
	const uint32_t nib = (word >> (16 * half + 4 * (3 - (q & 3)))) & 15u;
	// nib bits: 8 = H != diagonal, 4 = H != F, 2 = E opened here (DIAG_E), 1 = F opened here (DIAG_F)
	if (where == 0) {
	    if (!(nib & 8u)) {
		pa = __shfl_sync(FULL, qc, i - qbase);
		pb = __shfl_sync(FULL, rc, j - cbase);
		pend = true;
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
    if (pend) st.push(pa == pb ? 1u : 0u);
    if (aln_out) {
	if (lane == 0) aln_out[d.out] = st.pushed >= k ? st.aligned : 0;
	return;
    }
    if (lane == 0) {
	icb_align_result r;
	r.aligned = st.pushed >= k ? st.aligned : 0;
	r.score = en.score;
	r.end_query = en.ei;
	r.end_ref = en.ej;
	r.comp_len = int(st.pushed);
	out[d.out] = r;
    }
}

}  // namespace

int icb_align16_run(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* descs, int64_t cnt,
		    int64_t tiles, int64_t line_words, int max_m, int k)
{
    if (cnt <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    int rc = ws.ensure(ctx, size_t(tiles) * 32 + 64, size_t(line_words) * 4 + 64, cnt);
    if (rc) return rc;
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.d_descs, descs, sizeof(IcbPairDesc) * cnt, cudaMemcpyHostToDevice, st));
    const size_t smem = size_t(max_m + 1) * 4;
    ICB_LAUNCH(ctx, "align_dp",
	       icb_k_align16_dp<<<unsigned(cnt), A16_THREADS, smem, st>>>(d_seq, ws.d_descs, reinterpret_cast<uint4*>(ws.d_trace),
									   reinterpret_cast<uint32_t*>(ws.d_lines), ws.d_ends, nullptr));
    ICB_LAUNCH(ctx, "align_trace",
	       icb_k_align16_trace<<<unsigned((cnt + TB_WARPS - 1) / TB_WARPS), TB_WARPS * 32, 0, st>>>(
		   d_seq, ws.d_descs, int(cnt), reinterpret_cast<const uint4*>(ws.d_trace), ws.d_ends, k, ws.d_out, nullptr, nullptr));
    return ICB_OK;
}

// the same over a pair queue built on the device: *d_count descriptors (at most max_pairs = the grid), window counts to
// aln_out[desc.out] (-3 where the packed kernel refused the pair); the caller guarantees the workspace sizes
int icb_align16_run_dev(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* d_descs, const int32_t* d_count,
			int max_pairs, int max_m, int k, int32_t* aln_out)
{
    if (max_pairs <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    const size_t smem = size_t(max_m + 1) * 4;
    ICB_LAUNCH(ctx, "align_dp",
	       icb_k_align16_dp<<<unsigned(max_pairs), A16_THREADS, smem, st>>>(d_seq, d_descs, reinterpret_cast<uint4*>(ws.d_trace),
										 reinterpret_cast<uint32_t*>(ws.d_lines), ws.d_ends, d_count));
    ICB_LAUNCH(ctx, "align_trace",
	       icb_k_align16_trace<<<unsigned((max_pairs + TB_WARPS - 1) / TB_WARPS), TB_WARPS * 32, 0, st>>>(
		   d_seq, d_descs, 0, reinterpret_cast<const uint4*>(ws.d_trace), ws.d_ends, k, nullptr, d_count, aln_out));
    return ICB_OK;
}
