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
// icb_k_align16_trace — one warp per pair walks back from the end cell a RUN at a time (the lanes look ahead along the
//   diagonal / the gap, one ballot ends the run) and feeds the k-bit window register 32 columns at a time.
// Roofline class: integer ALU (GCUPS); trace traffic 0.5 B/cell written once, one 32-byte sector per 8 lanes and run read.
#include <algorithm>
#include <cstdlib>

#include "align.cuh"

#include "residency.cuh"

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
// PLAIN: the warp is not the first of the CTA (its thread 0 takes the row above from the ring, never from the matrix edge /
// the previous pass), holds no thread with the last query row and does not write the line of a following pass: the
// per-step tests of those three cases drop out of its inner loop.
template <int Q, bool EDGE, bool PLAIN>
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
	if (PLAIN || P.top == 0) {  // from the warp above
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
	if (!PLAIN && P.owner) {
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
	if (!PLAIN && P.write_line) {  // last thread of a pass that is not the last: its bottom row feeds the next pass
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
template <bool EDGE, bool PLAIN>
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
    step16<0, EDGE, PLAIN>(S, P, sl0 + 0, Wp, rd, wr);
    step16<1, EDGE, PLAIN>(S, P, sl0 + 1, Wp, rd, wr);
    step16<2, EDGE, PLAIN>(S, P, sl0 + 2, Wp, rd, wr);
    step16<3, EDGE, PLAIN>(S, P, sl0 + 3, Wp, rd, wr);
    if (store) *half_tile = make_uint4(S.T[0], S.T[1], S.T[2], S.T[3]);
}

__global__ void __launch_bounds__(A16_THREADS, 4)
icb_k_align16_dp(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, uint4* __restrict__ trace,
		 uint32_t* __restrict__ lines, IcbAlignEnd* __restrict__ ends, const int32_t* __restrict__ n_pairs_dev, int allow_plain)
{
    if (n_pairs_dev && int(blockIdx.x) >= *n_pairs_dev) return;  // device-built pair queue (dloop.cu): the grid is its capacity
    const unsigned int res_t0 = g_icb_residency.rec ? icb_res_now() : 0u;
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
	// (warp-uniform) none of the special cases of step16 can occur in this warp during this pass
	const bool plain = allow_plain && warp > 0 && !((last / A16_PASS_ROWS == pass) && (((last % A16_PASS_ROWS) >> 8) == warp)) &&
			   !(pass + 1 < npass && warp == A16_WARPS - 1);
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
		    for (int hq = 0; hq < 2; hq++) quad16<true, false>(S, P, sl0 + 4 * hq, dst + hq, store);
		}
		else if (plain) {
#pragma unroll 1
		    for (int hq = 0; hq < 2; hq++) quad16<false, true>(S, P, sl0 + 4 * hq, dst + hq, store);
		}
		else {
#pragma unroll 1
		    for (int hq = 0; hq < 2; hq++) quad16<false, false>(S, P, sl0 + 4 * hq, dst + hq, store);
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
	    icb_res_record(RES_DP_CTA, icb_res_tag(descs), res_t0, unsigned(n) | (unsigned(m) << 16));
	}
    }
}

// ------------------------------------------------------------------------------------------------ one warp per pair
// icb_k_align16w_dp — the same recurrence, lane layout (8 rows per lane as two packed halves, lane l two columns behind
// lane l - 1) and trace tiles as icb_k_align16_dp, but ONE WARP walks the whole matrix in strips of 256 rows: no CTA
// barrier, no ring between warps, no stagger between them (the ramp is 64 steps per strip instead of 80 per warp of the
// CTA), and every warp of an SM works on its own pair, so the four schedulers carry the same load whatever the read
// length.  What the CTA kernel takes from shared memory comes through registers here: the one-hot score word of a
// column travels down the lanes with the column (one more shuffle per step), lane 0 takes it — and the bottom row of the
// strip above — from a 32-step feed the warp refills with coalesced loads one group ahead, and lane 31 leaves the strip's
// bottom row in a 32-step buffer flushed with one coalesced store.
constexpr int AW_GROUP = 32;  // steps per feed / flush group

struct PassW {
    int m, n, lane;
    int row0;		 // first query row of this lane in this strip
    uint32_t NEGO, C2O;	 // packed -open, packed 2-open
    bool owner;		 // this lane holds the last query row (last strip only)
    int owner_rr;
};

// LAST: the strip that holds the last query row (tracks the row maximum, leaves no bottom row behind)
template <int Q, bool EDGE, bool LAST>
__device__ __forceinline__ void stepw(Lane16& S, uint32_t& Whl, const PassW& P, int sl, const uint2 v, uint32_t* __restrict__ outb)
{
    // v = this step's feed entry, outb = this quad's four entries (offsets below are compile-time)
    const int lane = P.lane;
    const int jl = sl - 2 * lane;
    uint32_t x = __shfl_up_sync(FULL, S.hb, 1);
    uint32_t y = __shfl_up_sync(FULL, S.fb, 1);
    uint32_t wl = __shfl_up_sync(FULL, Whl, 1);  // the score word of the column the lane above has just finished in its high half
    if (lane == 0) {  // row above the strip (matrix edge: H = 0, F = -inf; or the bottom row of the strip before) and W[sl]
	x = v.x << 16;
	y = v.x & 0xffff0000u;
	wl = v.y;
    }
    uint32_t hup = prmt(x, S.hb, 0x5432);  // low half <- the rows above (previous lane's high half), high half <- own low half
    uint32_t fup = prmt(y, S.fb, 0x5432);
    uint32_t hd = S.hup_prev;
    S.hup_prev = hup;
    const bool active = EDGE ? (jl >= 0 && jl <= P.m) : true;
    if (active) {
	const uint32_t Wlo = wl;
	const uint32_t Whi = S.Wprev;
	S.Wprev = Wlo;
	Whl = Whi;
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
	if (LAST && P.owner) {
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
    }
    else if (EDGE) {
	// keep the nibble positions of the half tile aligned with the step index
#pragma unroll
	for (int r = 0; r < 4; r++) S.T[r] *= 16u;
    }
    // the strip's bottom row at column jl - 1 (the flush drops what lies outside the matrix)
    if (!LAST && lane == 31) outb[Q] = (S.hb >> 16) | (S.fb & 0xffff0000u);
}

template <bool EDGE, bool LAST>
__device__ __forceinline__ void quadw(Lane16& S, uint32_t& Whl, const PassW& P, int sl0, const uint2* __restrict__ feed, uint32_t* __restrict__ outb,
				      uint4* __restrict__ half_tile, bool store)
{
    if (EDGE) {
#pragma unroll
	for (int r = 0; r < 4; r++) S.T[r] = 0;
    }
    // the four feed entries are taken up front (every lane reads the same words): lane 0's row above is on the critical
    // path of each step, its load must not be
    const uint2 v0 = feed[0], v1 = feed[1], v2 = feed[2], v3 = feed[3];
    stepw<0, EDGE, LAST>(S, Whl, P, sl0 + 0, v0, outb);
    stepw<1, EDGE, LAST>(S, Whl, P, sl0 + 1, v1, outb);
    stepw<2, EDGE, LAST>(S, Whl, P, sl0 + 2, v2, outb);
    stepw<3, EDGE, LAST>(S, Whl, P, sl0 + 3, v3, outb);
    if (store) *half_tile = make_uint4(S.T[0], S.T[1], S.T[2], S.T[3]);
}

// s_prog[w] = strip * 65536 + columns of that strip's bottom row warp w has put in the line (monotone): strip s + 1, in the
// next warp, reads a column of the line only after strip s wrote it.
template <bool LAST>
__device__ __forceinline__ void stripw(Lane16& S, PassW& P, const uint8_t* __restrict__ seq, const IcbPairDesc& d, int pass, uint4* __restrict__ tr,
				       uint32_t* __restrict__ line, uint2* __restrict__ s_feed, uint32_t* __restrict__ s_out,
				       volatile int* __restrict__ s_prog, int warp, int nwarps)
{
    const int lane = P.lane, n = P.n, m = P.m;
    const int nloc = (m + 64 + 7) & ~7;	 // lane 31 finishes column m at step m + 62
    P.row0 = pass * (32 * 8) + lane * 8;
#pragma unroll
    for (int r = 0; r < 4; r++) {
	const int ilo = P.row0 + r, ihi = P.row0 + 4 + r;
	const int alo = ilo < n ? (code5(seq[d.q_off + ilo]) & 3) : 0;
	const int ahi = ihi < n ? (code5(seq[d.q_off + ihi]) & 3) : 0;
	S.SEL[r] = uint32_t(alo) | 0x80u | (uint32_t(4 + ahi) << 8) | 0x8000u;
	S.H[r] = pack2(A16_BIAS - 2, A16_BIAS - 2);  // H[i][-1] = 0 at column -1
	S.E[r] = ONE2;				     // E[i][-1] = -inf
	S.T[r] = 0;
    }
    S.hb = pack2(A16_BIAS - 2, A16_BIAS - 2);
    S.fb = ONE2;
    S.hup_prev = pack2(A16_BIAS - 2, A16_BIAS - 2);
    S.Wprev = 0;
    uint32_t Whl = 0;
    const int rowblock = pass * 32 + lane;
    const bool rows_valid = P.row0 < n;
    __syncwarp();  // this warp's strip before is done with the feed
    const int above = (warp + nwarps - 1) % nwarps;  // the warp that computes strip pass - 1
    // what lane 0 needs at step c: the row above at column c and the score word of column c; `upto` columns of the
    // strip above must be in the line
    auto load_feed = [&](int c, int upto) {
	uint2 v;
	if (pass > 0) {
	    const int need = (pass - 1) * 65536 + min(m, upto);
	    while (s_prog[above] < need) __nanosleep(64);
	    if (nwarps > 1) __threadfence_block();
	}
	v.x = (pass > 0 && c < m) ? __ldcg(line + c) : (uint32_t(2 * c + A16_BIAS) | 0x00010000u);
	v.y = c < m ? (4u << (8 * (code5(icb_ref_char(seq, d, c)) & 3))) : 0u;
	return v;
    };
    uint2 pre = load_feed(lane, AW_GROUP);
    const int nblk = nloc >> 3;
    for (int blk = 0; blk < nblk; blk++) {
	const int sl0 = blk * 8;
	if ((blk & 3) == 0) {  // a new group of 32 steps: publish its feed, fetch the next one
	    const int g = blk >> 2;
	    s_feed[(g & 1) * AW_GROUP + lane] = pre;
	    __syncwarp();
	    pre = load_feed((g + 1) * AW_GROUP + lane, (g + 2) * AW_GROUP);
	}
	const uint2* feed = s_feed + ((blk >> 2) & 1) * AW_GROUP + (sl0 & (AW_GROUP - 1));
	uint32_t* outb = s_out + (sl0 & (AW_GROUP - 1));
	const bool edge = sl0 < 63 || sl0 + 7 > m - 2;
	const bool store = rows_valid && sl0 + 7 >= 2 * lane && sl0 <= 2 * lane + m;
	uint4* dst = tr + (int64_t(rowblock) * d.ncg + blk) * 2;
	if (edge) {
#pragma unroll 1
	    for (int hq = 0; hq < 2; hq++) quadw<true, LAST>(S, Whl, P, sl0 + 4 * hq, feed + 4 * hq, outb + 4 * hq, dst + hq, store);
	}
	else {
#pragma unroll 1
	    for (int hq = 0; hq < 2; hq++) quadw<false, LAST>(S, Whl, P, sl0 + 4 * hq, feed + 4 * hq, outb + 4 * hq, dst + hq, store);
	}
	if (!LAST && ((blk & 3) == 3 || blk + 1 == nblk)) {
	    // entry e of the buffer belongs to step (sl0 & ~31) + e, i.e. to column that step - 63
	    __syncwarp();
	    const int base = sl0 & ~(AW_GROUP - 1);
	    const int c = base + lane - 63;
	    if (c >= 0 && c < m && base + lane <= sl0 + 7) __stcg(line + c, s_out[lane]);
	    if (nwarps > 1) __threadfence_block();  // the line is read by another warp of this CTA (both sides go to L2: .cg)
	    __syncwarp();
	    if (lane == 0) s_prog[warp] = pass * 65536 + (blk + 1 == nblk ? m : max(0, min(m, sl0 + 8 - 63)));
	}
    }
}

constexpr int AW_MAXW = 8;  // warps per pair, at most

// blockDim.x / 32 warps per pair: warp w computes strips w, w + W, w + 2W, ... each as soon as the strip above has its
// first columns in the line (a pipeline of warps through L2, ~160 steps apart, without any CTA barrier).  W = 1 gives the
// most pairs per second, a larger W the shorter time per pair the device loop's rounds wait for.
__global__ void __launch_bounds__(32 * AW_MAXW, 3)
icb_k_align16w_dp(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, uint4* __restrict__ trace,
		  uint32_t* __restrict__ lines, IcbAlignEnd* __restrict__ ends, const int32_t* __restrict__ n_pairs_dev)
{
    icb_res_mark(RES_DP, descs);
    if (n_pairs_dev && int(blockIdx.x) >= *n_pairs_dev) return;  // device-built pair queue (dloop.cu): the grid is its capacity
    const unsigned int res_t0 = g_icb_residency.rec ? icb_res_now() : 0u;
    __shared__ uint2 s_feed[AW_MAXW][2 * AW_GROUP];
    __shared__ uint32_t s_out[AW_MAXW][AW_GROUP];
    __shared__ int s_prog[AW_MAXW];
    __shared__ int s_red[AW_MAXW * 2];
    __shared__ int s_row[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const IcbPairDesc d = descs[blockIdx.x];
    const int n = d.n, m = d.m;
    // refuse anything outside ACGT (parasail scores it 0): the 32-bit kernel takes the pair
    int bad = 0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) bad |= code5(seq[d.r_off + j]) > 3;
    for (int i = threadIdx.x; i < n; i += blockDim.x) bad |= code5(seq[d.q_off + i]) > 3;
    if (lane == 0) s_prog[warp] = 0;
    if (__syncthreads_or(bad)) {
	if (threadIdx.x == 0) {
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
    PassW P;
    P.m = m;
    P.n = n;
    P.lane = lane;
    P.NEGO = pack2(-d.open, -d.open);
    P.C2O = pack2(2 - d.open, 2 - d.open);
    const int last = n - 1;
    P.owner = ((last & 255) >> 3) == lane;
    P.owner_rr = last & 7;
    const int npass = (n + 255) >> 8;
    uint4* tr = trace + d.trace_off * 2;
    uint32_t* line = lines + d.line_off;
    for (int pass = warp; pass < npass; pass += nwarps) {
	if (pass + 1 < npass) stripw<false>(S, P, seq, d, pass, tr, line, s_feed[warp], s_out[warp], s_prog, warp, nwarps);
	else stripw<true>(S, P, seq, d, pass, tr, line, s_feed[warp], s_out[warp], s_prog, warp, nwarps);
    }
    // ---- end cell (sg_trace_scan_avx2_256_16.c:355-395): first maximum of the last row, then the last column:
    // strictly larger wins, the smallest row among equals; a tie with the row maximum only moves the end if
    // that maximum sits in the corner.
    const int ol = (last & 255) >> 3;
    if (warp == (npass - 1) % nwarps && lane == ol) {
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
    if (threadIdx.x == 0) {
	for (int w = 1; w < nwarps; w++) {
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
	icb_res_record(RES_DP_CTA, icb_res_tag(descs), res_t0, unsigned(n) | (unsigned(m) << 16));
    }
}

// getAlnRatio's window count (cluster.cpp:442-459), fed up to 32 comparison-string columns at a time: the same sequence of
// states as IcbWindowCounter::push column by column (align.cuh), every lane holding the same state and lane p evaluating
// the window that closes at the p-th new column.
struct WarpWindowCounter {
    uint32_t reg, mask;
    int k, limit;
    long long pushed;
    int aligned;
    __device__ __forceinline__ void init(int kk, int lim)
    {
	reg = 0;
	mask = kk >= 32 ? 0xffffffffu : ((1u << kk) - 1u);
	k = kk;
	limit = lim;
	pushed = 0;
	aligned = 0;
    }
    // bits: bit t = the t-th new column is a match; cnt <= 32 columns (warp-uniform arguments)
    __device__ __forceinline__ void push_bits(uint32_t bits, int cnt, int lane)
    {
	if (cnt <= 0) return;
	const uint32_t upto = bits & ((2u << lane) - 1u);				   // columns 0..lane
	const uint32_t fresh = __brev(upto) >> (31 - lane);				   // column 0 in bit `lane`, column `lane` in bit 0
	const uint32_t w = ((lane == 31 ? 0u : (reg << (lane + 1))) | fresh) & mask;	   // the register after column `lane`
	const bool counts = lane < cnt && pushed + lane >= k && __popc(w) >= limit;
	aligned += __popc(__ballot_sync(FULL, counts));
	reg = __shfl_sync(FULL, w, cnt - 1);
	pushed += cnt;
    }
    __device__ void push_zeros(long long g, int lane)
    {
	const int first = int(g < k ? g : k);
	push_bits(0u, first, lane);
	const long long rest = g - first;
	if (rest > 0) {
	    // the register is all zero now; every further column closes an all-blank window
	    long long counted = rest;
	    if (pushed < k) counted -= (k - pushed);
	    if (counted > 0 && 0 >= limit) aligned += int(counted);
	    pushed += rest;
	}
    }
};

// the 4 trace bits of cell (i, j): icb_k_align16_dp's tile layout (row block i / 8, skewed step u = j + 2 * (rowblock mod 32) + half)
__device__ __forceinline__ uint32_t trace_nibble(const uint32_t* __restrict__ tr, int ncg, int i, int j)
{
    const int gg = i >> 3, half = (i >> 2) & 1, r = i & 3;
    const int u = j + 2 * (gg & 31) + half;
    const uint32_t word = __ldcg(tr + ((int64_t(gg) * ncg + (u >> 3)) * 2 + ((u >> 2) & 1)) * 4 + r);
    return (word >> (16 * half + 4 * (3 - (u & 3)))) & 15u;
}

// ------------------------------------------------------------------------------------------------ traceback
// Mirrors parasail_result_get_traceback (traceback_template.c:72-175) and feeds getAlnRatio's window count
// (cluster.cpp:442-459).  One warp per pair; the walk itself is serial, so the warp looks AHEAD instead: in the diagonal
// state lane t reads the trace bits (and the two characters) of cell (i - t, j - t), in a gap state those of the t-th
// cell along the gap; one ballot finds where the run ends, the whole run goes into the window count at once, and the
// cell that ended a diagonal run also yields the first step of the gap it opens.  One round of loads per run (a dozen
// columns on ONT reads) instead of one dependent shuffle chain per column.
constexpr int TB_WARPS = 4;

__global__ void __launch_bounds__(TB_WARPS * 32)
icb_k_align16_trace(const uint8_t* __restrict__ seq, const IcbPairDesc* __restrict__ descs, int n_pairs,
		    const uint4* __restrict__ trace, const IcbAlignEnd* __restrict__ ends, int k,
		    icb_align_result* __restrict__ out, const int32_t* __restrict__ n_pairs_dev, int32_t* __restrict__ aln_out)
{
    icb_res_mark(RES_TRACE, descs);
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
    const unsigned int res_t0 = g_icb_residency.rec ? icb_res_now() : 0u;
    const uint32_t* tr = reinterpret_cast<const uint32_t*>(trace + d.trace_off * 2);
    WarpWindowCounter st;
    st.init(k, d.limit);
    int i = en.ei, j = en.ej;
    // free tail (traceback_template.c:72-94)
    if (i + 1 == n) st.push_zeros(m - 1 - j, lane);
    else st.push_zeros(n - 1 - i, lane);
    int where = 0;  // 0 diag, 1 ins (consumes reference), 2 del (consumes query)
    while (i >= 0 && j >= 0) {
	// nib bits: 8 = H != diagonal, 4 = H != F, 2 = E opened here (DIAG_E), 1 = F opened here (DIAG_F)
	if (where == 0) {
	    const int ii = i - lane, jj = j - lane;
	    const bool valid = ii >= 0 && jj >= 0;
	    uint32_t nib = 8u, qc = 0u, rc = 1u;
	    if (valid) {
		nib = trace_nibble(tr, ncg, ii, jj);
		qc = icb_up_char(seq[d.q_off + ii]);
		rc = icb_up_char(icb_ref_char(seq, d, jj));
	    }
	    const uint32_t stops = __ballot_sync(FULL, (nib & 8u) != 0u);
	    const uint32_t same = __ballot_sync(FULL, qc == rc);  // traceback.c match_char, case-insensitive
	    const int run = stops ? __ffs(stops) - 1 : 32;
	    st.push_bits(run == 32 ? same : (same & ((1u << run) - 1u)), run, lane);
	    i -= run;
	    j -= run;
	    if (run < 32 && i >= 0 && j >= 0) {
		// neither diagonal nor F: INS (traceback_template.c:118-120); the gap's first column is this very cell
		const uint32_t sn = __shfl_sync(FULL, nib, run);
		st.push_bits(0u, 1, lane);
		if (sn & 4u) {
		    j--;
		    where = (sn & 2u) ? 0 : 1;
		}
		else {
		    i--;
		    where = (sn & 1u) ? 0 : 2;
		}
	    }
	}
	else {
	    const int ii = where == 1 ? i : i - lane, jj = where == 1 ? j - lane : j;
	    const bool valid = ii >= 0 && jj >= 0;
	    const uint32_t closes = where == 1 ? 2u : 1u;
	    const uint32_t nib = valid ? trace_nibble(tr, ncg, ii, jj) : closes;
	    const uint32_t stops = __ballot_sync(FULL, (nib & closes) != 0u);
	    const int first = stops ? __ffs(stops) - 1 : 32;
	    const int left = where == 1 ? j + 1 : i + 1;  // cells before the matrix edge
	    // the closing cell is itself a gap column; an edge ends the gap without one
	    const int cols = first == 32 ? 32 : (first < left ? first + 1 : first);
	    st.push_bits(0u, cols, lane);
	    if (where == 1) j -= cols;
	    else i -= cols;
	    if (first < 32) where = 0;
	}
    }
    if (i < 0 && j >= 0) st.push_zeros(j + 1, lane);
    else if (j < 0 && i >= 0) st.push_zeros(i + 1, lane);
    if (lane == 0) icb_res_record(RES_TRACE_WARP, icb_res_tag(descs), res_t0, unsigned(n) | (unsigned(m) << 16));
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

static int a16_allow_plain()
{
    static const int v = getenv("ICB_A16_NOPLAIN") ? 0 : 1;
    return v;
}
// ICB_A16_CTA=1: the CTA-per-pair wavefront kernel instead of the warp-per-pair one (same results, same trace tiles)
static bool a16_use_cta()
{
    static const bool v = getenv("ICB_A16_CTA") != nullptr;
    return v;
}
// warps per pair of the warp-pipeline kernel (ICB_A16_WARPS; latency-first contexts: ICB_A16_WARPS_LATENCY)
static int a16_warps(bool latency)
{
    static const int w = getenv("ICB_A16_WARPS") ? std::max(1, std::min(AW_MAXW, atoi(getenv("ICB_A16_WARPS")))) : 2;
    static const int wl = getenv("ICB_A16_WARPS_LATENCY") ? std::max(1, std::min(AW_MAXW, atoi(getenv("ICB_A16_WARPS_LATENCY")))) : 8;
    return latency ? wl : w;
}

int icb_align16_run(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* descs, int64_t cnt,
		    int64_t tiles, int64_t line_words, int max_m, int k)
{
    if (cnt <= 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    int rc = ws.ensure(ctx, size_t(tiles) * 32 + 64, size_t(line_words) * 4 + 64, cnt);
    if (rc) return rc;
    ICB_CUDA(ctx, cudaMemcpyAsync(ws.d_descs, descs, sizeof(IcbPairDesc) * cnt, cudaMemcpyHostToDevice, st));
    const size_t smem = size_t(max_m + 1) * 4;
    if (a16_use_cta())
	ICB_LAUNCH(ctx, "align_dp",
		   icb_k_align16_dp<<<unsigned(cnt), A16_THREADS, smem, st>>>(d_seq, ws.d_descs, reinterpret_cast<uint4*>(ws.d_trace),
									       reinterpret_cast<uint32_t*>(ws.d_lines), ws.d_ends, nullptr, a16_allow_plain()));
    else
	ICB_LAUNCH(ctx, "align_dp",
		   icb_k_align16w_dp<<<unsigned(cnt), 32 * a16_warps(false), 0, st>>>(d_seq, ws.d_descs, reinterpret_cast<uint4*>(ws.d_trace),
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
    // the worker's batch rank decides how its alignment kernels queue against those of the other batches in flight
    if (a16_use_cta())
	ICB_LAUNCH(ctx, "align_dp",
		   ICB_CUDA(ctx, icb_launch_prio(icb_k_align16_dp, dim3(unsigned(max_pairs)), dim3(A16_THREADS), smem, st, ctx->aln_priority, d_seq,
						 d_descs, reinterpret_cast<uint4*>(ws.d_trace), reinterpret_cast<uint32_t*>(ws.d_lines), ws.d_ends,
						 d_count, a16_allow_plain())));
    else
	ICB_LAUNCH(ctx, "align_dp",
		   ICB_CUDA(ctx, icb_launch_prio(icb_k_align16w_dp, dim3(unsigned(max_pairs)), dim3(32 * a16_warps(ctx->aln_latency)), 0, st, ctx->aln_priority, d_seq, d_descs,
						 reinterpret_cast<uint4*>(ws.d_trace), reinterpret_cast<uint32_t*>(ws.d_lines), ws.d_ends, d_count)));
    ICB_LAUNCH(ctx, "align_trace",
	       ICB_CUDA(ctx, icb_launch_prio(icb_k_align16_trace, dim3(unsigned((max_pairs + TB_WARPS - 1) / TB_WARPS)), dim3(TB_WARPS * 32), 0, st,
					     ctx->aln_priority, d_seq, d_descs, 0, reinterpret_cast<const uint4*>(ws.d_trace),
					     static_cast<const IcbAlignEnd*>(ws.d_ends), k, static_cast<icb_align_result*>(nullptr), d_count, aln_out)));
    return ICB_OK;
}

// ------------------------------------------------------------------------------------------------ diagnostics
cudaError_t icb_residency_set_dloop(const IcbResidencyLog& L);
cudaError_t icb_residency_set_index(const IcbResidencyLog& L);
cudaError_t icb_residency_set_align32(const IcbResidencyLog& L);
static cudaError_t residency_set_everywhere(const IcbResidencyLog& L)
{
    cudaError_t e = icb_residency_set_here(L);
    if (e == cudaSuccess) e = icb_residency_set_dloop(L);
    if (e == cudaSuccess) e = icb_residency_set_index(L);
    if (e == cudaSuccess) e = icb_residency_set_align32(L);
    return e;
}

extern "C" {

int icb_debug_residency_start(icb_ctx* ctx, int64_t capacity)
{
    if (!ctx || capacity <= 0 || capacity > (int64_t(1) << 28)) return ICB_ERR_ARG;
    IcbResidencyLog L;
    ICB_CUDA(ctx, cudaMemcpyFromSymbol(&L, g_icb_residency, sizeof L));
    if (L.rec) {
	cudaFree(L.rec);
	cudaFree(L.count);
    }
    ICB_CUDA(ctx, cudaMalloc(&L.rec, sizeof(uint4) * size_t(capacity)));
    ICB_CUDA(ctx, cudaMalloc(&L.count, sizeof(unsigned int)));
    ICB_CUDA(ctx, cudaMemset(L.count, 0, sizeof(unsigned int)));
    L.cap = unsigned(capacity);
    ICB_CUDA(ctx, residency_set_everywhere(L));
    ICB_CUDA(ctx, cudaDeviceSynchronize());
    return ICB_OK;
}

// copies the records out (at most `capacity` of them; returns how many were written through *n_records), stops and frees the log
int icb_debug_residency_dump(icb_ctx* ctx, uint32_t* records, int64_t capacity, int64_t* n_records)
{
    if (!ctx || !n_records) return ICB_ERR_ARG;
    ICB_CUDA(ctx, cudaDeviceSynchronize());
    IcbResidencyLog L;
    ICB_CUDA(ctx, cudaMemcpyFromSymbol(&L, g_icb_residency, sizeof L));
    if (!L.rec) return ctx->fail(ICB_ERR_STATE, "the residency log is not running");
    unsigned int cnt = 0;
    ICB_CUDA(ctx, cudaMemcpy(&cnt, L.count, sizeof cnt, cudaMemcpyDeviceToHost));
    const int64_t have = std::min<int64_t>(cnt, L.cap);
    const int64_t take = std::min<int64_t>(have, capacity);
    if (records && take > 0) ICB_CUDA(ctx, cudaMemcpy(records, L.rec, sizeof(uint4) * size_t(take), cudaMemcpyDeviceToHost));
    *n_records = take;
    cudaFree(L.rec);
    cudaFree(L.count);
    L = IcbResidencyLog{nullptr, nullptr, 0u};
    ICB_CUDA(ctx, residency_set_everywhere(L));
    return ICB_OK;
}

}  // extern "C"
