// Stage 1b — batched read sketching: homopolymer compression, 2-bit packing, forward and
// reverse-complement k-mer values, sliding-window minimizers (leftmost minimum, emit on change).
//
// Replaces, per read, HomopolymerCompress (reference hpc.cpp:4-32), RevComp (util.cpp:13-38),
// KmerEncodeSeq (kmer_index.cpp:5-17, kmer_index.h:29-66) and GetKmerMinimizers
// (minimizer.cpp:78-123) as driven by PrepareSortedBatch (qualscore.cpp:62-96).
//
// Design (B200): one 128-thread CTA per read, persistent over the read list.  All intermediate state of a
// read lives in shared memory (compressed bases, 2-bit codes, one byte per window holding argmin-minus-
// window-start for each strand, flag bit-masks and their prefix counts), so HBM sees each input byte once
// and each output byte once.
//   A. raw tiles of 2048 bases are staged with 128-bit loads; run ends are found with warp ballots, ranked
//      with popc prefix sums and the run maximum of the qualities is taken by a short backward walk;
//   B. k-mer values of both strands are extracted from the packed codes (funnel shifts; the reverse strand is
//      the bit-reversed complement of the same field, no second sequence is materialised), staged per 512
//      windows and reduced to the window argmin;
//   C. "argmin changed" flags are ballots, ranks are popc prefix sums, the read's slice of the minimizer
//      arena is claimed with one atomicAdd and {Min,Pos,Index} triples are written.
// Reads longer than LMAX bases take the same code with its buffers in a global scratch arena (slow path).
//
// Roofline class: HBM.  Algorithmic bytes per read: 2L (seq+qual in) + 2L_h (compressed seq+qual out)
// + 12(M_f+M_r) (triples out) + 24 (per-read scalars).
#include "icb_common.cuh"

namespace {

constexpr int SK_THREADS = 128;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int TILE_A = 2048;		    // raw bases per phase-A tile
constexpr int STAGE_A = TILE_A + 32;	    // staged bytes per array (alignment head + look-ahead)
constexpr int TILE_B = 512;		    // windows per phase-B tile
constexpr int MAX_W = 64;		    // w-k+1 supported by the staging buffer
constexpr int STAGE_B = TILE_B + MAX_W + 8; // k-mer values staged per strand
constexpr int STAGE_BYTES = (2 * STAGE_A > 8 * STAGE_B) ? 2 * STAGE_A : 8 * STAGE_B;
constexpr int LMAX_FAST = 4096;		    // longest raw read on the shared-memory path

struct SketchArgs {
    const uint8_t* seq;
    const uint8_t* qual;
    const int64_t* offs;
    const uint8_t* skip;
    int64_t n_reads;
    int k, w;
    int32_t* status;
    int32_t* hpc_len;
    uint8_t* hpc_seq;
    uint8_t* hpc_qual;
    icb_minimizer* mins;
    int64_t mins_capacity;
    int64_t* min_off;
    int32_t* nf;
    int32_t* nr;
    unsigned long long* counters;  // [0] arena cursor [1] error flags [2] list-1 length [3] ticket (stage 1) [4] list-2 length
				   // [5] ticket (stage 2) [6] ticket (stage 3)
    int32_t* long_list;		   // list 1: reads the warp kernel handed on (too long for it, or too many minimizers)
    int32_t* long_list2;	   // list 2: reads longer than the shared-memory CTA kernel takes
    int use_list;		   // CTA kernel: 0 = all reads by ticket, 1 = the reads of list 1
    uint8_t* long_scratch;	   // GLOBAL variant: per-CTA scratch
    int64_t long_scratch_stride;
    int64_t lmax;		   // capacity (bases) of the per-read buffers
    // 1 << (32 - 2k), handed to the warp kernel as DATA: a k-mer in the low 2k bits of a word (anything above it) times this
    // is the k-mer in the top 2k bits, one IMAD on the FMA pipe.  Computed in the kernel the compiler sees a power of two and
    // emits a shift and an add on the ALU pipe, the pipe this kernel is bound by.
    uint32_t key_mul;
    // The same for the "argmin changed" flags: minus_one = 0xffffffff and bit[t] = 1 << t as data, so that key - previous key and
    // flag * bit + mask are IMADs (a compare, a select and an add per window on the ALU pipe otherwise).
    uint32_t minus_one;
    uint32_t bit[16];
};

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ bool is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// A,C,G,T -> 0,1,2,3 (kmer_index.h:29-45)
__device__ __forceinline__ uint32_t code_of(uint32_t c) { return ((c >> 1) & 3u) ^ ((c >> 2) & 1u); }

// kk = min(k,16) bases starting at base position p, as a 2*kk-bit number, first base most significant
__device__ __forceinline__ uint32_t field_at(const uint32_t* codes, int p, int kk)
{
    int wi = p >> 4;
    uint64_t x = (uint64_t(codes[wi]) << 32) | codes[wi + 1];
    int sh = 64 - 2 * (p & 15) - 2 * kk;
    uint32_t v = uint32_t(x >> sh);
    return kk == 16 ? v : (v & ((1u << (2 * kk)) - 1u));
}

// reverse complement of a kk-base field
__device__ __forceinline__ uint32_t rc_field(uint32_t g, int kk)
{
    uint32_t r = __brev(~g);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return r >> (32 - 2 * kk);
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total via *total
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp /*SK_WARPS+1*/, int* total)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
	int t = __shfl_up_sync(0xffffffffu, inc, d);
	if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SK_WARPS; i++) {
	int t = s_warp[i];
	if (i < warp) base += t;
	tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

template <bool GLOBAL, typename PT>
__global__ void __launch_bounds__(SK_THREADS) icb_k_sketch(SketchArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, W = a.w - a.k + 1;
    const int kk = k < 16 ? k : 16;
    const int64_t lmax = a.lmax;

    // ---- carve the per-read buffers
    uint8_t* stage = smem_raw;				      // STAGE_BYTES
    __shared__ uint32_t s_amask[TILE_A / 32];		      // phase A run-end ballots
    __shared__ int s_apre[TILE_A / 32];
    __shared__ int s_warp[SK_WARPS + 1];
    __shared__ int s_carry[2][2];			      // [parity]{base, max}
    __shared__ long long s_base;
    __shared__ int s_next;

    uint8_t *hseq, *hqual;
    uint32_t *codes, *fmask, *rmask;
    PT *fpre, *rpre;
    {
	uint8_t* p = GLOBAL ? a.long_scratch + int64_t(blockIdx.x) * a.long_scratch_stride : smem_raw + STAGE_BYTES;
	int64_t nmask = lmax / 32 + 2;
	hseq = p;  p += lmax + 32;
	hqual = p; p += lmax + 32;
	codes = reinterpret_cast<uint32_t*>(p); p += (lmax / 16 + 2) * 4;
	fmask = reinterpret_cast<uint32_t*>(p); p += nmask * 4;
	rmask = reinterpret_cast<uint32_t*>(p); p += nmask * 4;
	fpre = reinterpret_cast<PT*>(p); p += nmask * sizeof(PT);
	rpre = reinterpret_cast<PT*>(p);
    }
    uint8_t* sA_seq = stage;
    uint8_t* sA_qual = stage + STAGE_A;
    uint32_t* sB_fv = reinterpret_cast<uint32_t*>(stage);
    uint32_t* sB_rv = sB_fv + STAGE_B;

    for (;;) {
	// ---- next read
	__syncthreads();
	if (tid == 0) {
	    if (GLOBAL) {
		unsigned long long t = atomicAdd(&a.counters[6], 1ull);
		s_next = t < a.counters[4] ? a.long_list2[t] : -1;
	    }
	    else if (a.use_list) {
		unsigned long long t = atomicAdd(&a.counters[5], 1ull);
		s_next = t < a.counters[2] ? a.long_list[t] : -1;
	    }
	    else {
		unsigned long long t = atomicAdd(&a.counters[5], 1ull);
		s_next = t < (unsigned long long)a.n_reads ? int(t) : -1;
	    }
	}
	__syncthreads();
	const int r = s_next;
	if (r < 0) break;

	const int64_t start = a.offs[r];
	const int L = int(a.offs[r + 1] - start);
	if (!GLOBAL) {
	    if (a.skip && a.skip[r]) {
		if (tid == 0) { a.status[r] = 0; a.hpc_len[r] = 0; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		continue;
	    }
	    // qualscore.cpp:62-63: compress only if raw length > 2k or >= w
	    if (!(unsigned(L) > unsigned(2 * k) || unsigned(L) >= unsigned(a.w))) {
		if (tid == 0) { a.status[r] = 2; a.hpc_len[r] = 0; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		continue;
	    }
	    if (L > lmax) {
		if (tid == 0) a.long_list2[atomicAdd(&a.counters[4], 1ull)] = r;
		continue;
	    }
	}
	const int shiftA = int(start & 15);  // hseq/hqual keep the 16-byte phase of the global slot
	const uint8_t* gseq = a.seq + (start - shiftA);
	const uint8_t* gqual = a.qual + (start - shiftA);

	// ================================================================ phase A: homopolymer compression
	int runBase = 0;
	bool bad = false;
	const int nTilesA = (L + TILE_A - 1) / TILE_A;
	for (int t = 0; t < nTilesA; t++) {
	    const int p0 = t * TILE_A;
	    // stage [p0 - shiftA, p0 - shiftA + STAGE_A) of the read's slot
	    const int needA = min(STAGE_A, shiftA + (L - p0) + 1);  // +1: look-ahead byte (slack after the buffers covers it)
	    for (int c = tid; c < (needA + 15) / 16; c += SK_THREADS) {
		reinterpret_cast<uint4*>(sA_seq)[c] = reinterpret_cast<const uint4*>(gseq + p0)[c];
		reinterpret_cast<uint4*>(sA_qual)[c] = reinterpret_cast<const uint4*>(gqual + p0)[c];
	    }
	    __syncthreads();
	    uint32_t endbits = 0;
#pragma unroll
	    for (int i = 0; i < TILE_A / SK_THREADS; i++) {
		int p = p0 + i * SK_THREADS + tid;
		int si = p - p0 + shiftA;
		bool valid = p < L;
		uint32_t c = sA_seq[si];
		uint32_t nx = (p + 1 < L) ? sA_seq[si + 1] : 0x100u;
		bool isEnd = valid && c != nx;
		uint32_t m = __ballot_sync(0xffffffffu, isEnd);
		if (lane == 0) s_amask[i * SK_WARPS + warp] = m;
		endbits |= uint32_t(isEnd) << i;
	    }
	    __syncthreads();
	    // exclusive prefix over the 64 ballot words (in position order: word j = positions 32j..32j+31)
	    if (warp == 0) {
		int c0 = __popc(s_amask[2 * lane]), c1 = __popc(s_amask[2 * lane + 1]);
		int inc = c0 + c1;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
		    int v = __shfl_up_sync(0xffffffffu, inc, d);
		    if (lane >= d) inc += v;
		}
		s_apre[2 * lane] = inc - c0 - c1;
		s_apre[2 * lane + 1] = inc - c1;
		if (lane == 31) s_warp[SK_WARPS] = inc;
	    }
	    __syncthreads();
	    const int carryBase = s_carry[(t + 1) & 1][0], carryMax = s_carry[(t + 1) & 1][1];
#pragma unroll
	    for (int i = 0; i < TILE_A / SK_THREADS; i++) {
		int p = p0 + i * SK_THREADS + tid;
		bool lastOfTile = (p == p0 + TILE_A - 1) && (p + 1 < L);
		if (!((endbits >> i) & 1u) && !lastOfTile) continue;
		int si = p - p0 + shiftA;
		uint32_t c = sA_seq[si];
		int qmax = (signed char)sA_qual[si];
		int j = si - 1;
		while (j >= shiftA && sA_seq[j] == c) {
		    int q = (signed char)sA_qual[j];
		    qmax = q > qmax ? q : qmax;  // hpc.cpp:24 compares chars (signed)
		    j--;
		}
		if (j < shiftA && t > 0 && uint32_t(carryBase) == c) qmax = carryMax > qmax ? carryMax : qmax;
		if ((endbits >> i) & 1u) {
		    int w = i * SK_WARPS + warp;
		    int o = runBase + s_apre[w] + __popc(s_amask[w] & lanemask_lt());
		    bad |= !is_acgt(c);
		    if (o < lmax) {
			hseq[shiftA + o] = uint8_t(c);
			hqual[shiftA + o] = uint8_t(qmax);
		    }
		}
		if (lastOfTile) {
		    s_carry[t & 1][0] = int(c);
		    s_carry[t & 1][1] = qmax;
		}
	    }
	    runBase += s_warp[SK_WARPS];
	    __syncthreads();
	}
	const int Lh = runBase;
	const bool sketchable = !(unsigned(Lh) < unsigned(2 * k) || unsigned(Lh) < unsigned(a.w));  // qualscore.cpp:65-66
	if (!sketchable) {
	    if (tid == 0) { a.status[r] = 3; a.hpc_len[r] = Lh; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
	    continue;
	}
	if (__syncthreads_or(bad)) {
	    // the reference's RevComp throws here (util.cpp:31-33)
	    if (tid == 0) { atomicOr(&a.counters[1], 1ull); a.status[r] = -1; a.hpc_len[r] = Lh; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
	    continue;
	}
	// ---- compressed read -> global (16-byte chunks share the phase of the global slot)
	{
	    uint8_t* gs = a.hpc_seq + (start - shiftA);
	    uint8_t* gq = a.hpc_qual + (start - shiftA);
	    const int lo = shiftA, hi = shiftA + Lh;	// valid byte range inside the phase-aligned buffers
	    const int nch = (hi + 15) >> 4;
	    for (int c = tid; c < nch; c += SK_THREADS) {
		int b0 = c << 4;
		if (b0 >= lo && b0 + 16 <= hi) {
		    reinterpret_cast<uint4*>(gs)[c] = reinterpret_cast<const uint4*>(hseq)[c];
		    reinterpret_cast<uint4*>(gq)[c] = reinterpret_cast<const uint4*>(hqual)[c];
		}
		else {
		    for (int b = b0; b < b0 + 16; b++)
			if (b >= lo && b < hi) {
			    gs[b] = hseq[b];
			    gq[b] = hqual[b];
			}
		}
	    }
	}
	// ---- pack 2-bit codes, 16 bases per word, first base in the top bits
	{
	    const int nwords = (Lh + 15) / 16 + 2;
	    for (int j = tid; j < nwords; j += SK_THREADS) {
		uint32_t wv = 0;
#pragma unroll
		for (int b = 0; b < 16; b++) {
		    int pos = j * 16 + b;
		    uint32_t c = pos < Lh ? code_of(hseq[shiftA + pos]) : 0u;
		    wv |= c << (30 - 2 * b);
		}
		codes[j] = wv;
	    }
	}
	__syncthreads();

	// ================================================================ phase B: window argmins
	// hseq / hqual are now reused: one byte per window = argmin - window start, forward / reverse strand
	uint8_t* wF = hseq;
	uint8_t* wR = hqual;
	const int nk = Lh - k;		// k-mers per strand (kmer_index.cpp:12 drops the last one)
	const int nwin = nk - W + 1;	// windows per strand; forward windows q=0..nwin-1, reverse q=1..nwin
	int Mf = 0, Mr = 0;
	if (nwin > 0) {
	    const int offF = k - kk;	// k > 16: uint32 wrap keeps the LAST 16 bases forward, the FIRST 16 reversed
	    for (int q0 = 0; q0 <= nwin; q0 += TILE_B) {
		// stage k-mer values of positions q0 .. q0+TILE_B+W-1 (clipped to nk inclusive)
		const int np = min(TILE_B + W, nk + 1 - q0);
		for (int i = tid; i < np; i += SK_THREADS) {
		    int p = q0 + i;
		    uint32_t g = field_at(codes, p, kk);
		    sB_rv[i] = rc_field(g, kk);
		    sB_fv[i] = offF ? field_at(codes, p + offF, kk) : g;
		}
		__syncthreads();
		const int nq = min(TILE_B, nwin + 1 - q0);
		for (int i = tid; i < nq; i += SK_THREADS) {
		    // forward: leftmost minimum of fv[q..q+W-1]; reverse strand: leftmost in reverse coordinates
		    // = rightmost minimum of rv[q..q+W-1]
		    uint32_t bf = sB_fv[i], br = sB_rv[i];
		    int tf = 0, tr = 0;
		    for (int t = 1; t < W; t++) {
			uint32_t vf = sB_fv[i + t], vr = sB_rv[i + t];
			if (vf < bf) { bf = vf; tf = t; }
			if (vr <= br) { br = vr; tr = t; }
		    }
		    wF[q0 + i] = uint8_t(tf);
		    wR[q0 + i] = uint8_t(tr);
		}
		__syncthreads();
	    }
	    // ---- flags: minimizer emitted for the first window and whenever the argmin position changes
	    const int nwords = (nwin + 32) / 32;  // covers q = 0..nwin
	    for (int q = tid; q < nwords * 32; q += SK_THREADS) {
		bool ff = false, fr = false;
		if (q < nwin) ff = (q == 0) || (int(wF[q]) + 1 != int(wF[q - 1]));
		if (q >= 1 && q <= nwin) fr = (q == nwin) || (int(wR[q]) != int(wR[q + 1]) + 1);
		uint32_t mf = __ballot_sync(0xffffffffu, ff), mr = __ballot_sync(0xffffffffu, fr);
		if (lane == 0) {
		    fmask[q >> 5] = mf;
		    rmask[q >> 5] = mr;
		}
	    }
	    __syncthreads();
	    for (int j0 = 0; j0 < nwords; j0 += SK_THREADS) {
		int j = j0 + tid;
		int cf = j < nwords ? __popc(fmask[j]) : 0, cr = j < nwords ? __popc(rmask[j]) : 0;
		int tot;
		int ef = block_excl_scan(cf, s_warp, &tot);
		if (j < nwords) fpre[j] = PT(Mf + ef);
		Mf += tot;
		int er = block_excl_scan(cr, s_warp, &tot);
		if (j < nwords) rpre[j] = PT(Mr + er);
		Mr += tot;
	    }
	}
	// ================================================================ phase C: claim the slice, write triples
	if (tid == 0) s_base = (long long)atomicAdd(&a.counters[0], (unsigned long long)(Mf + Mr));
	__syncthreads();
	const int64_t base = s_base;
	if (tid == 0) {
	    a.status[r] = 1;
	    a.hpc_len[r] = Lh;
	    a.nf[r] = Mf;
	    a.nr[r] = Mr;
	    a.min_off[r] = base;
	}
	if (base + Mf + Mr <= a.mins_capacity && nwin > 0) {
	    const int offF = k - kk;
	    icb_minimizer* of = a.mins + base;
	    icb_minimizer* orv = of + Mf;
	    for (int q = tid; q <= nwin; q += SK_THREADS) {
		uint32_t mf = fmask[q >> 5], mr = rmask[q >> 5];
		uint32_t bit = 1u << (q & 31), lt = bit - 1u;
		if (mf & bit) {
		    int idx = int(fpre[q >> 5]) + __popc(mf & lt);
		    int am = q + wF[q];
		    icb_minimizer m;
		    m.Min = field_at(codes, am + offF, kk);
		    m.Pos = uint32_t(am);
		    m.Index = uint32_t(idx);
		    of[idx] = m;
		}
		if (mr & bit) {
		    int asc = int(rpre[q >> 5]) + __popc(mr & lt);
		    int idx = Mr - 1 - asc;
		    int am = q + wR[q];
		    icb_minimizer m;
		    m.Min = rc_field(field_at(codes, am, kk), kk);
		    m.Pos = uint32_t(nk - am);	// position in reverse-complement coordinates
		    m.Index = uint32_t(idx);
		    orv[idx] = m;
		}
	    }
	}
    }
}

// ================================================================================================
// Warp-per-read kernel (the fast path: 2k + 5 <= 32 bits of key, W in {5, 8}, raw length <= SW_LMAX)
//
// One warp owns one read; nothing is block-synchronised.  Every lane keeps 16 consecutive bases (phase A) or 16
// consecutive windows (phase B) in registers, so the per-base work is SWAR / packed arithmetic instead of one thread
// per base:
//   A. run ends of a 16-base chunk from three SWAR operations per word, ranks by popc + one warp scan, the run
//      maximum of the qualities by a 16-step register loop (carried across lanes by a segmented warp scan);
//   B. the compressed read is packed twice at 2 bits per base — big-endian (forward k-mers are the top bits of a funnel
//      shift) and little-endian complemented (reverse-complement k-mers are the low bits) — and validated (ACGT only);
//      a window minimum is VIMNMX3 over packed keys (value << 5 | position in the lane's frame), which yields the
//      leftmost (forward) / rightmost (reverse strand) argmin directly; "argmin changed" is one compare of two keys;
//   C. flagged argmin positions go to a per-warp list in shared memory, the read's slice of the arena is claimed with
//      one atomicAdd and lane j writes minimizer j: fully coalesced 12-byte records.
constexpr int SW_WARPS = 8;
constexpr int SW_THREADS = SW_WARPS * 32;
// Two compiled shapes: reads up to 2688 bases with 64 registers -> 4 CTAs (32 warps) per SM, the default; reads up to 3072
// bases with 72 registers -> 3 CTAs per SM (ICB_SKETCH_CTAS=3).  Longer reads go to the CTA kernel either way.
constexpr int SW_TICKET = 4;		    // reads taken per ticket
template <int LMAX>
struct SwShape {
    static constexpr int HBUF = LMAX + 32;	 // bytes per compressed buffer: 16-byte phase of the global slot + slack
    static constexpr int NWORD = LMAX / 16 + 4;	 // packed words per stream
    static constexpr int LIST = HBUF / 2;	 // argmin positions (uint16) per strand that fit the reused buffers
    static constexpr int PER_WARP = 2 * HBUF + 2 * NWORD * 4;
};

__device__ __forceinline__ uint32_t sw_prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
// acc + (x != y) * bit as IMAD (x - y), VIMNMX (min with 1), IMAD: PTX, because the compiler turns the C form into a compare, a
// select and an add, all three on the ALU pipe (minus_one = 0xffffffff and bit come in as data, SketchArgs)
__device__ __forceinline__ uint32_t sw_flag_acc(uint32_t x, uint32_t y, uint32_t minus_one, uint32_t bit, uint32_t acc)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(y), "r"(minus_one), "r"(x));
    asm("min.u32 %0, %0, 1;" : "+r"(d));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(d), "r"(bit));
    return acc;
}
// bit i of the result = byte i of x is non-zero
__device__ __forceinline__ uint32_t sw_nonzero_bytes(uint32_t x)
{
    uint32_t t = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    return (t * 0x00204081u) >> 28;
}

// KC: k as a compile-time constant (0 = run-time k): which word pair a forward k-mer sits in is then known per unrolled step
template <int W, int SW_LMAX, int MIN_CTAS, int KC>
__global__ void __launch_bounds__(SW_THREADS, MIN_CTAS) icb_k_sketch_warp(SketchArgs a)
{
    constexpr int SW_HBUF = SwShape<SW_LMAX>::HBUF, SW_NWORD = SwShape<SW_LMAX>::NWORD, SW_LIST = SwShape<SW_LMAX>::LIST,
		  SW_PER_WARP = SwShape<SW_LMAX>::PER_WARP;
    extern __shared__ __align__(16) uint8_t sw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* hs = sw_smem + warp * SW_PER_WARP;
    uint8_t* hq = hs + SW_HBUF;
    uint32_t* Fs = reinterpret_cast<uint32_t*>(hq + SW_HBUF);
    uint32_t* Rs = Fs + SW_NWORD;
    const int k = KC ? KC : a.k;
    const uint32_t kmask = (1u << (2 * k)) - 1u;
    const uint32_t krmul = a.key_mul;  // 1 << (32 - 2k), opaque to the compiler (see SketchArgs)

    for (;;) {
	long long t0 = 0;
	if (lane == 0) t0 = (long long)atomicAdd(&a.counters[3], (unsigned long long)SW_TICKET);
	t0 = __shfl_sync(0xffffffffu, t0, 0);
	if (t0 >= a.n_reads) break;
	for (int tr = 0; tr < SW_TICKET; tr++) {
	    const long long r = t0 + tr;
	    if (r >= a.n_reads) break;
	    __syncwarp();
	    const int64_t start = a.offs[r];
	    const int L = int(a.offs[r + 1] - start);
	    if (a.skip && a.skip[r]) {
		if (lane == 0) { a.status[r] = 0; a.hpc_len[r] = 0; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		continue;
	    }
	    // qualscore.cpp:62-63: compress only if raw length > 2k or >= w
	    if (!(unsigned(L) > unsigned(2 * k) || unsigned(L) >= unsigned(a.w))) {
		if (lane == 0) { a.status[r] = 2; a.hpc_len[r] = 0; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		continue;
	    }
	    if (L > SW_LMAX) {
		if (lane == 0) a.long_list[atomicAdd(&a.counters[2], 1ull)] = int(r);
		continue;
	    }
	    const int shiftA = int(start & 15);  // hs / hq keep the 16-byte phase of the global slot
	    const uint8_t* gseq = a.seq + (start - shiftA);
	    const uint8_t* gqual = a.qual + (start - shiftA);

	    // ============================================================ phase A: homopolymer compression
	    int runBase = 0;
	    int tileCarry = -128;  // running quality maximum of a run left open by the previous tile
	    const int nchunks = (shiftA + L + 15) >> 4;
	    for (int c0 = 0; c0 < nchunks; c0 += 32) {
		const int c = c0 + lane;
		const int pfirst = 16 * c - shiftA;  // raw position of byte 0 of this lane's chunk
		uint4 s4 = make_uint4(0, 0, 0, 0), q4 = make_uint4(0, 0, 0, 0);
		uint32_t nbcur = 0;
		if (c < nchunks) {
		    s4 = *reinterpret_cast<const uint4*>(gseq + 16 * c);
		    q4 = *reinterpret_cast<const uint4*>(gqual + 16 * c);
		    if (lane == 31) nbcur = gseq[16 * c + 16];
		}
		// valid bytes of the chunk: raw positions 0 .. L-1
		const int lo = max(0, -pfirst), hi = min(16, L - pfirst);
		uint32_t vmask = hi > lo ? ((0xffffu >> (16 - hi)) & (0xffffu << lo)) : 0u;
		// the byte after the chunk: first byte of the next lane's chunk; lane 31 reads it (it exists: slack after the buffers)
		uint32_t nb = __shfl_down_sync(0xffffffffu, s4.x, 1);
		if (lane == 31) nb = nbcur;
		// run ends: byte i differs from byte i+1
		uint32_t e = sw_nonzero_bytes(s4.x ^ __funnelshift_r(s4.x, s4.y, 8)) |
			     (sw_nonzero_bytes(s4.y ^ __funnelshift_r(s4.y, s4.z, 8)) << 4) |
			     (sw_nonzero_bytes(s4.z ^ __funnelshift_r(s4.z, s4.w, 8)) << 8) |
			     (sw_nonzero_bytes(s4.w ^ __funnelshift_r(s4.w, nb, 8)) << 12);
		const int ilast = L - 1 - pfirst;  // the last base of the read always ends a run
		if (ilast >= 0 && ilast < 16) e |= 1u << ilast;
		e &= vmask;
		if (vmask != 0xffffu) {
		    // bytes outside the read must not contribute to a run maximum: quality -128
		    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
		    for (int i = 0; i < 4; i++) {
			m0 |= ((vmask >> i) & 1u) * (0xffu << (8 * i));
			m1 |= ((vmask >> (4 + i)) & 1u) * (0xffu << (8 * i));
			m2 |= ((vmask >> (8 + i)) & 1u) * (0xffu << (8 * i));
			m3 |= ((vmask >> (12 + i)) & 1u) * (0xffu << (8 * i));
		    }
		    q4.x = (q4.x & m0) | (0x80808080u & ~m0);
		    q4.y = (q4.y & m1) | (0x80808080u & ~m1);
		    q4.z = (q4.z & m2) | (0x80808080u & ~m2);
		    q4.w = (q4.w & m3) | (0x80808080u & ~m3);
		}
		// ranks of the run ends
		const int cnt = __popc(e);
		int inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
		    int v = __shfl_up_sync(0xffffffffu, inc, d);
		    if (lane >= d) inc += v;
		}
		uint8_t* ph = hs + (shiftA + runBase + inc - cnt);  // hq = hs + SW_HBUF
		uint8_t* const ph0 = ph;
		runBase += __shfl_sync(0xffffffffu, inc, 31);
		// 16-step register loop: running maximum of the open run, one (base, max) pair stored per run end
		int qmax = -128;
		const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
		const uint32_t qw[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
		for (int i = 0; i < 16; i++) {
		    // byte i of the quality word, sign-extended (hpc.cpp:24 compares chars): prmt with sign replication
		    const int qv = int(sw_prmt(qw[i >> 2], 0u, 0x8880u | uint32_t(i & 3) | (uint32_t(i & 3) << 4) | (uint32_t(i & 3) << 8) |
									    (uint32_t(i & 3) << 12)));
		    qmax = max(qmax, qv);
		    if (e & (1u << i)) {
			ph[0] = uint8_t(sw[i >> 2] >> (8 * (i & 3)));
			ph[SW_HBUF] = uint8_t(qmax);
			ph++;
			qmax = -128;
		    }
		}
		// carry of a run that is still open at the end of a chunk into the next chunk(s)
		const bool open = vmask != 0 && !(e & 0x8000u) && hi == 16;
		int cv = open ? qmax : -128;
		bool pass = open && e == 0;  // the whole chunk belongs to one open run
		if (lane == 0 && pass) cv = max(cv, tileCarry);
		if (__any_sync(0xffffffffu, pass)) {  // rare: a run longer than a chunk passes its maximum through whole chunks
#pragma unroll
		    for (int d = 1; d < 32; d <<= 1) {
			const int v2 = __shfl_up_sync(0xffffffffu, cv, d);
			const int p2 = __shfl_up_sync(0xffffffffu, int(pass), d);
			if (lane >= d && pass) {
			    cv = max(cv, v2);
			    pass = p2 != 0;
			}
		    }
		}
		int carryIn = __shfl_up_sync(0xffffffffu, cv, 1);
		if (lane == 0) carryIn = tileCarry;
		tileCarry = __shfl_sync(0xffffffffu, cv, 31);
		if (carryIn > -128 && e != 0) {
		    const int qv = int((signed char)ph0[SW_HBUF]);
		    if (carryIn > qv) ph0[SW_HBUF] = uint8_t(carryIn);
		}
	    }
	    __syncwarp();
	    const int Lh = runBase;
	    if (unsigned(Lh) < unsigned(2 * k) || unsigned(Lh) < unsigned(a.w)) {  // qualscore.cpp:65-66
		if (lane == 0) { a.status[r] = 3; a.hpc_len[r] = Lh; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		continue;
	    }
	    // ---- pack 2 bits per base (both streams), check ACGT; word j of a stream = compressed bases 16j .. 16j+15
	    {
		const uint32_t* hsw = reinterpret_cast<const uint32_t*>(hs);
		const int wsh = shiftA >> 2, bsh = 8 * (shiftA & 3);
		const int nwords = (Lh + 15) >> 4;
		uint32_t badacc = 0;
		for (int j = lane; j < nwords + 3; j += 32) {  // three zero words behind the read: lanes past its end read them
		    uint32_t F = 0, R = 0;
		    if (j < nwords) {
			uint32_t x[5];
#pragma unroll
			for (int i = 0; i < 5; i++) x[i] = hsw[wsh + 4 * j + i];
			uint32_t dacc = 0, dlast[4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
			    const uint32_t wv = __funnelshift_r(x[i], x[i + 1], bsh);	// bases 16j+4i .. +3, first base in the low byte
			    const uint32_t cd = ((wv >> 1) & 0x03030303u) ^ ((wv >> 2) & 0x01010101u);	// A,C,G,T -> 0..3 per byte
			    uint32_t sel = (cd | (cd >> 4)) & 0x00ff00ffu;
			    sel = (sel | (sel >> 8)) & 0xffffu;
			    dlast[i] = sw_prmt(0x54474341u, 0u, sel) ^ wv;	// "ACGT"[code] must give the base back
			    dacc |= dlast[i];
			    F |= ((cd * 0x40100401u) >> 24) << (24 - 8 * i);		   // b0<<6 | b1<<4 | b2<<2 | b3
			    R |= (((cd ^ 0x03030303u) * 0x01041040u) >> 24) << (8 * i);   // complement: b0 | b1<<2 | b2<<4 | b3<<6
			}
			if (j == nwords - 1) {	// the last word may hold stale bytes behind the read: only its first Lh - 16j bases count
			    dacc = 0;
#pragma unroll
			    for (int i = 0; i < 4; i++) {
				const int nvalid = Lh - (16 * j + 4 * i);
				dacc |= nvalid >= 4 ? dlast[i] : (nvalid <= 0 ? 0u : (dlast[i] & (0xffffffffu >> (32 - 8 * nvalid))));
			    }
			}
			badacc |= dacc;
		    }
		    Fs[j] = F;
		    Rs[j] = R;
		}
		if (__any_sync(0xffffffffu, badacc != 0)) {
		    // the reference's RevComp throws here (util.cpp:31-33)
		    if (lane == 0) { atomicOr(&a.counters[1], 1ull); a.status[r] = -1; a.hpc_len[r] = Lh; a.nf[r] = 0; a.nr[r] = 0; a.min_off[r] = 0; }
		    continue;
		}
	    }
	    // ---- compressed read -> global (16-byte chunks share the phase of the global slot)
	    {
		uint8_t* gs = a.hpc_seq + (start - shiftA);
		uint8_t* gq = a.hpc_qual + (start - shiftA);
		const int lo = shiftA, hi = shiftA + Lh;
		const int cfull0 = (lo + 15) >> 4, cfull1 = hi >> 4;  // chunks [cfull0, cfull1) lie entirely inside [lo, hi)
		for (int c = cfull0 + lane; c < cfull1; c += 32) {
		    reinterpret_cast<uint4*>(gs)[c] = reinterpret_cast<const uint4*>(hs)[c];
		    reinterpret_cast<uint4*>(gq)[c] = reinterpret_cast<const uint4*>(hq)[c];
		}
		// head [lo, 16*cfull0) and tail [16*cfull1, hi): at most 15 bytes each, one byte per lane
		const int headEnd = min(hi, cfull0 << 4);
		if (lo + lane < headEnd && lane < 16) {
		    gs[lo + lane] = hs[lo + lane];
		    gq[lo + lane] = hq[lo + lane];
		}
		const int tail0 = max(cfull1 << 4, headEnd) + (lane - 16);
		if (lane >= 16 && tail0 < hi) {
		    gs[tail0] = hs[tail0];
		    gq[tail0] = hq[tail0];
		}
	    }
	    __syncwarp();
	    // ============================================================ phase B: window minima, flags, argmin lists
	    uint16_t* posF = reinterpret_cast<uint16_t*>(hs);  // the compressed bytes are in global memory now
	    uint16_t* posR = reinterpret_cast<uint16_t*>(hq);
	    const int nk = Lh - k;	   // k-mers per strand (kmer_index.cpp:12 drops the last one)
	    const int nwin = nk - W + 1;   // forward windows q = 0 .. nwin-1, reverse windows (forward coordinates) q = 1 .. nwin
	    int Mf = 0, Mr = 0;
	    bool overflow = false;
	    if (nwin > 0) {
		uint32_t prevF = 0, prevR = 0;	// last window keys of the previous tile's lane 31 (in that lane's frame)
		for (int q00 = 0; q00 <= nwin + 1; q00 += 512) {
		    const int q0 = q00 + 16 * lane;
		    const int jw = min(q0 >> 4, SW_NWORD - 3);	// lanes past the end of the read only produce unused keys
		    const uint32_t fB = Fs[jw], fC = Fs[jw + 1], fD = Fs[jw + 2], rB = Rs[jw], rC = Rs[jw + 1], rD = Rs[jw + 2];
		    // keys of the k-mers at positions q0 + o, o = 0 .. 15 + W - 1; frame index i = o + 1 (i = 0 is position q0 - 1)
		    uint32_t kf[16 + W - 1], kr[16 + W - 1];
#pragma unroll
		    for (int o = 0; o < 16 + W - 1; o++) {
			const uint32_t z = o == 0 ? rB : (o < 16 ? __funnelshift_r(rB, rC, 2 * o) : (o == 16 ? rC : __funnelshift_r(rC, rD, 2 * (o - 16))));
			if (KC != 0) {
			    // forward k-mer into the LOW 2k bits of a word (whatever lies above it is multiplied away): one
			    // shift of the word, or one funnel shift of the pair, that holds it — known per step for a constant k
			    const uint32_t hi = o < 16 ? fB : fC, lo = o < 16 ? fC : fD;
			    const int sh = 64 - 2 * (o & 15) - 2 * KC;  // right shift of hi:lo that brings the last base to bit 0
			    const uint32_t yl = sh >= 32 ? (hi >> (sh - 32)) : __funnelshift_r(lo, hi, sh);
			    kf[o] = yl * krmul + uint32_t(o + 1);  // value in the top 2k bits, position in the low 5
			}
			else {
			    const uint32_t y = o == 0 ? fB : (o < 16 ? __funnelshift_l(fC, fB, 2 * o) : (o == 16 ? fC : __funnelshift_l(fD, fC, 2 * (o - 16))));
			    const uint32_t vf = y >> (32 - 2 * k);
			    kf[o] = vf * 32u + uint32_t(o + 1);	   // smallest key = smallest value, leftmost position
			}
			// reverse strand: leftmost there = rightmost here.  The multiplication drops the bits above the k-mer
			// (mod 2^32): value in the top 2k bits, position in the low 5 of the 32 - 2k below — one IMAD, no mask
			kr[o] = z * krmul + uint32_t(31 - (o + 1));
		    }
		    uint32_t mf[16], mr[16];
#pragma unroll
		    for (int t = 0; t < 16; t++) {
			uint32_t x = kf[t], y = kr[t];
#pragma unroll
			for (int u = 1; u + 1 < W; u += 2) {
			    x = __vimin3_u32(x, kf[t + u], kf[t + u + 1]);
			    y = __vimin3_u32(y, kr[t + u], kr[t + u + 1]);
			}
			if ((W & 1) == 0) {
			    x = min(x, kf[t + W - 1]);
			    y = min(y, kr[t + W - 1]);
			}
			mf[t] = x;
			mr[t] = y;
		    }
		    // key of the window before this lane's first one, moved into this lane's frame (16 positions further)
		    uint32_t pf = __shfl_up_sync(0xffffffffu, mf[15], 1), pr = __shfl_up_sync(0xffffffffu, mr[15], 1);
		    if (lane == 0) {
			pf = prevF;
			pr = prevR;
		    }
		    pf -= 16u;	// forward key holds +index
		    pr += 16u;	// reverse key holds 31 - index
		    prevF = __shfl_sync(0xffffffffu, mf[15], 31);
		    prevR = __shfl_sync(0xffffffffu, mr[15], 31);
		    // forward flag of window q: its argmin differs from window q-1 (always for q = 0)
		    // reverse flag, kept by the owner of window q, belongs to window q-1: its argmin differs from window q
		    // (always for q-1 = nwin)
		    uint32_t ff = 0, fr = 0;
#pragma unroll
		    for (int t = 0; t < 16; t++) {
			const uint32_t bf = t == 0 ? pf : mf[t - 1];
			const uint32_t br = t == 0 ? pr : mr[t - 1];
			// keys differ <=> argmin positions differ: difference and accumulation on the FMA pipe (SketchArgs::bit)
			ff = sw_flag_acc(mf[t], bf, a.minus_one, a.bit[t], ff);
			fr = sw_flag_acc(mr[t], br, a.minus_one, a.bit[t], fr);
		    }
		    {
			// forward: windows q0+t in [0, nwin), window 0 always; reverse (bit t <-> window q0+t-1): windows in [1, nwin],
			// window nwin always
			const int fhi = min(16, nwin - q0);
			const uint32_t fvalid = fhi > 0 ? (0xffffu >> (16 - fhi)) : 0u;
			if (q0 == 0) ff |= 1u;
			ff &= fvalid;
			const int rlo = max(0, 2 - q0), rhi = min(16, nwin + 2 - q0);
			const uint32_t rvalid = rhi > rlo ? ((0xffffu >> (16 - rhi)) & (0xffffu << rlo)) : 0u;
			const int tl = nwin + 1 - q0;
			if (tl >= 0 && tl < 16) fr |= 1u << tl;
			fr &= rvalid;
		    }
		    // ranks
		    const int cf = __popc(ff), cr = __popc(fr);
		    int inc = cf | (cr << 16);
#pragma unroll
		    for (int d = 1; d < 32; d <<= 1) {
			int v = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += v;
		    }
		    const int tot = __shfl_sync(0xffffffffu, inc, 31);
		    int rf = Mf + (inc & 0xffff) - cf, rr = Mr + (inc >> 16) - cr;
		    Mf += tot & 0xffff;
		    Mr += tot >> 16;
		    if (Mf > SW_LIST || Mr > SW_LIST) {  // (stores are predicated: nothing is written past a list)
			overflow = true;
			break;
		    }
		    // argmin positions (k-mer start, forward coordinates) of the flagged windows; running pointers, not indices:
		    // one predicated store + one predicated pointer bump per flag
		    {
			uint16_t* pF = posF + rf;
			uint16_t* pR = posR + rr;
			const uint32_t q0m1 = uint32_t(q0 - 1);
#pragma unroll
			for (int t = 0; t < 16; t++) {
			    const uint32_t key = t == 0 ? pr : mr[t - 1];
			    const uint32_t vF = q0m1 + (mf[t] & 31u);
			    const uint32_t vR = q0m1 + (~key & 31u);  // 31 - (key & 31)
			    if (ff & (1u << t)) *pF++ = uint16_t(vF);
			    if (fr & (1u << t)) *pR++ = uint16_t(vR);
			}
		    }
		}
	    }
	    if (overflow) {  // more minimizers than the shared-memory lists hold: the CTA kernel takes the read
		if (lane == 0) a.long_list[atomicAdd(&a.counters[2], 1ull)] = int(r);
		continue;
	    }
	    __syncwarp();
	    // ============================================================ phase C: claim the slice, write triples
	    long long base = 0;
	    if (lane == 0) {
		base = (long long)atomicAdd(&a.counters[0], (unsigned long long)(Mf + Mr));
		a.status[r] = 1;
		a.hpc_len[r] = Lh;
		a.nf[r] = Mf;
		a.nr[r] = Mr;
		a.min_off[r] = base;
	    }
	    base = __shfl_sync(0xffffffffu, base, 0);
	    if (base + Mf + Mr <= a.mins_capacity) {
		// running word pointers (32 records = 96 words per step) instead of 64-bit index arithmetic per record
		uint32_t* of = reinterpret_cast<uint32_t*>(a.mins + base + lane);
		const uint16_t* pp = posF + lane;
		for (int j = lane; j < Mf; j += 32, of += 96, pp += 32) {
		    const uint32_t p = *pp;
		    const uint32_t* fw = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(Fs) + ((p >> 2) & ~3u));
		    const uint32_t y = __funnelshift_l(fw[1], fw[0], 2 * (p & 15));
		    of[0] = y >> (32 - 2 * k);	// Min
		    of[1] = p;			// Pos
		    of[2] = uint32_t(j);	// Index
		}
		uint32_t* orv = reinterpret_cast<uint32_t*>(a.mins + base + Mf + (Mr - 1 - lane));
		pp = posR + lane;
		for (int j = lane; j < Mr; j += 32, orv -= 96, pp += 32) {
		    const uint32_t p = *pp;
		    const uint32_t* rw = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(Rs) + ((p >> 2) & ~3u));
		    const uint32_t z = __funnelshift_r(rw[0], rw[1], 2 * (p & 15));
		    orv[0] = z & kmask;		       // Min
		    orv[1] = uint32_t(nk) - p;	       // Pos in reverse-complement coordinates
		    orv[2] = uint32_t(Mr - 1 - j);     // Index
		}
	    }
	}
    }
}

template <typename PT>
size_t per_read_bytes(int64_t lmax)
{
    int64_t nmask = lmax / 32 + 2;
    return size_t(2 * (lmax + 32) + (lmax / 16 + 2) * 4 + 2 * nmask * 4 + 2 * nmask * sizeof(PT));
}

// canonical CSR for the download: out[offs_f[i] ...] / out_r[...] from the arena
__global__ void icb_k_gather_mins(const icb_minimizer* __restrict__ arena, const int64_t* __restrict__ min_off,
				  const int32_t* __restrict__ nf, const int32_t* __restrict__ nr,
				  const int64_t* __restrict__ offs_f, const int64_t* __restrict__ offs_r, int64_t n_reads,
				  icb_minimizer* __restrict__ out_f, icb_minimizer* __restrict__ out_r)
{
    int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(arena + min_off[r]);
    int a = nf[r], b = nr[r];
    uint32_t* df = reinterpret_cast<uint32_t*>(out_f + offs_f[r]);
    uint32_t* dr = reinterpret_cast<uint32_t*>(out_r + offs_r[r]);
    for (int i = lane; i < 3 * a; i += 32) df[i] = src[i];
    for (int i = lane; i < 3 * b; i += 32) dr[i] = src[3 * a + i];
}

}  // namespace

int icb_launch_sketch(icb_readset* rs, int k, int w, const uint8_t* d_skip)
{
    icb_ctx* ctx = rs->ctx;
    SketchArgs a;
    a.seq = rs->d_seq;
    a.qual = rs->d_qual;
    a.offs = rs->d_offs;
    a.skip = d_skip;
    a.n_reads = rs->n_reads;
    a.k = k;
    a.w = w;
    a.status = rs->d_status;
    a.hpc_len = rs->d_hpc_len;
    a.hpc_seq = rs->d_hpc_seq;
    a.hpc_qual = rs->d_hpc_qual;
    a.mins = rs->d_mins;
    a.mins_capacity = rs->mins_capacity;
    a.min_off = rs->d_min_off;
    a.nf = rs->d_nf;
    a.nr = rs->d_nr;
    a.counters = rs->d_counters;
    a.long_list = rs->d_long_list;
    a.long_list2 = rs->d_long_list + rs->n_reads + 1;
    a.use_list = 0;
    a.long_scratch = nullptr;
    a.long_scratch_stride = 0;
    a.lmax = LMAX_FAST;
    a.key_mul = (2 * k <= 31) ? (1u << (32 - 2 * k)) : 0u;  // only the warp kernel (2k + 5 <= 32) reads it
    a.minus_one = 0xffffffffu;
    for (int t = 0; t < 16; t++) a.bit[t] = 1u << t;
    if (rs->n_reads == 0) return ICB_OK;
    cudaStream_t st = ctx->stream;
    const int W = w - k + 1;

    // ---- stage 1: warp-per-read kernel (keys of 2k + 5 bits, compiled window widths)
    const bool warp_path = 2 * k + 5 <= 32 && k >= 2 && (W == 5 || W == 8);
    if (warp_path) {
	static const int want_ctas = [] {
	    const char* e = getenv("ICB_SKETCH_CTAS");
	    return e && atoi(e) == 3 ? 3 : 4;
	}();
	const size_t smem = size_t(SW_WARPS) * (want_ctas == 3 ? SwShape<3072>::PER_WARP : SwShape<2688>::PER_WARP);
	// compile-time k for the two usual sketches: k 11 / w 15 (the reference's defaults, args.h:13-16) and k 13 / w 20
	const int kc = (k == 11 && W == 5) ? 11 : (k == 13 && W == 8) ? 13 : 0;
	auto kern = want_ctas == 3 ? (W == 5 ? (kc ? icb_k_sketch_warp<5, 3072, 3, 11> : icb_k_sketch_warp<5, 3072, 3, 0>)
					     : (kc ? icb_k_sketch_warp<8, 3072, 3, 13> : icb_k_sketch_warp<8, 3072, 3, 0>))
				   : (W == 5 ? (kc ? icb_k_sketch_warp<5, 2688, 4, 11> : icb_k_sketch_warp<5, 2688, 4, 0>)
					     : (kc ? icb_k_sketch_warp<8, 2688, 4, 13> : icb_k_sketch_warp<8, 2688, 4, 0>));
	ICB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	int per_sm = 0;
	ICB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SW_THREADS, smem));
	if (per_sm < 1) per_sm = 1;
	int64_t grid = int64_t(ctx->sm_count) * per_sm;
	const int64_t need = (rs->n_reads + int64_t(SW_WARPS) * SW_TICKET - 1) / (int64_t(SW_WARPS) * SW_TICKET);
	if (grid > need) grid = need;
	ICB_LAUNCH(ctx, "sketch_reads", kern<<<unsigned(grid), SW_THREADS, smem, st>>>(a));
	// reads it handed on (longer than SW_LMAX, or too many minimizers for its lists)
	unsigned long long c[8];
	ICB_CUDA(ctx, cudaMemcpyAsync(c, rs->d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, cudaStreamSynchronize(st));
	if (c[2] == 0) return ICB_OK;
	a.use_list = 1;
    }
    // ---- stage 2: CTA-per-read kernel, buffers in shared memory (every k / w; reads up to LMAX_FAST)
    {
	size_t smem = STAGE_BYTES + per_read_bytes<uint16_t>(LMAX_FAST);
	ICB_CUDA(ctx, cudaFuncSetAttribute(icb_k_sketch<false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
	int per_sm = 0;
	ICB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icb_k_sketch<false, uint16_t>, SK_THREADS, smem));
	if (per_sm < 1) per_sm = 1;
	int64_t grid = int64_t(ctx->sm_count) * per_sm;
	if (grid > rs->n_reads) grid = rs->n_reads;
	ICB_LAUNCH(ctx, warp_path ? "sketch_reads_cta" : "sketch_reads",
		   icb_k_sketch<false, uint16_t><<<unsigned(grid), SK_THREADS, smem, st>>>(a));
    }
    // ---- stage 3: the same code with its buffers in a global scratch arena, for reads above LMAX_FAST
    if (rs->max_len > LMAX_FAST) {
	const int nblk = 64;
	int64_t lmaxL = (rs->max_len + 63) & ~int64_t(31);
	int64_t stride = (int64_t(per_read_bytes<uint32_t>(lmaxL)) + 255) & ~int64_t(255);
	if (!rs->d_long_scratch) {
	    cudaError_t e = icb_dmalloc(ctx, (void**)&rs->d_long_scratch, size_t(stride) * nblk);
	    if (e != cudaSuccess) return ctx->fail_cuda(e, "cudaMalloc long scratch", __FILE__, __LINE__);
	}
	a.long_scratch = rs->d_long_scratch;
	a.long_scratch_stride = stride;
	a.lmax = lmaxL;
	ICB_LAUNCH(ctx, "sketch_reads_long",
		   icb_k_sketch<true, uint32_t><<<nblk, SK_THREADS, STAGE_BYTES, st>>>(a));
    }
    return ICB_OK;
}

extern "C" {

int icb_sketch_reads(icb_readset* rs, int kmer_size, int window_size, const double qualtab_nomin[129],
		     const uint8_t* skip)
{
    IcbRange nvtx_range("PrepareSortedBatch sketches (icb_sketch_reads)");
    if (!rs || !qualtab_nomin) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    const int k = kmer_size, w = window_size;
    if (k < 1 || k > 31 || w < k) return ctx->fail(ICB_ERR_ARG, "need 1 <= k <= 31 and k <= w");
    if (w - k + 1 > MAX_W) return ctx->fail(ICB_ERR_ARG, "w - k + 1 above 64 is not supported");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    // per-read / per-base arrays are allocated for the read set's capacity (reads appended later: icb_readset_append_from)
    const int64_t n = std::max(rs->n_reads, rs->cap_reads), nb = std::max(rs->n_bases, rs->cap_bases);
    cudaError_t e;
#define SK_ALLOC(ptr, bytes)                                              \
    if (!(ptr) && (e = icb_dmalloc(ctx, (void**)&(ptr), (size_t)(bytes))) != cudaSuccess) \
	return ctx->fail_cuda(e, "cudaMalloc " #ptr, __FILE__, __LINE__);
    SK_ALLOC(rs->d_status, sizeof(int32_t) * (n + 1));
    SK_ALLOC(rs->d_hpc_len, sizeof(int32_t) * (n + 1));
    SK_ALLOC(rs->d_hpc_err, sizeof(double) * (n + 1));
    SK_ALLOC(rs->d_hpc_seq, nb + 256);
    SK_ALLOC(rs->d_hpc_qual, nb + 256);
    SK_ALLOC(rs->d_min_off, sizeof(int64_t) * (n + 1));
    SK_ALLOC(rs->d_nf, sizeof(int32_t) * (n + 1));
    SK_ALLOC(rs->d_nr, sizeof(int32_t) * (n + 1));
    SK_ALLOC(rs->d_long_list, sizeof(int32_t) * 2 * (n + 1));  // list 1 and list 2
    uint8_t* d_skip = nullptr;
    if (skip && rs->n_reads > 0) {
	ICB_CUDA(ctx, cudaMallocAsync((void**)&d_skip, rs->n_reads, ctx->stream));
	ICB_CUDA(ctx, cudaMemcpyAsync(d_skip, skip, rs->n_reads, cudaMemcpyHostToDevice, ctx->stream));
    }
    // minimizer arena: expected density 2/(W+1) per compressed base and strand; grow and rerun on overflow
    const int W = w - k + 1;
    int64_t want = int64_t(double(nb) * (2.0 * 2.0 / (W + 1)) * 0.95) + 1024;
    if (want > 2 * nb + 1024) want = 2 * nb + 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
	if (rs->mins_capacity < want) {
	    icb_dfree(ctx, rs->d_mins);
	    rs->d_mins = nullptr;
	    rs->mins_capacity = 0;
	    if ((e = icb_dmalloc(ctx, (void**)&rs->d_mins, sizeof(icb_minimizer) * size_t(want))) != cudaSuccess)
		return ctx->fail_cuda(e, "cudaMalloc minimizer arena", __FILE__, __LINE__);
	    rs->mins_capacity = want;
	}
	ICB_CUDA(ctx, cudaMemsetAsync(rs->d_counters, 0, sizeof(unsigned long long) * 8, ctx->stream));
	int rc = icb_launch_sketch(rs, k, w, d_skip);
	if (rc) return rc;
	unsigned long long c[8];
	ICB_CUDA(ctx, cudaMemcpyAsync(c, rs->d_counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
	ICB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (c[1] & 1ull) {
	    if (d_skip) cudaFreeAsync(d_skip, ctx->stream);
	    return ctx->fail(ICB_ERR_INVALID_BASE, "Invalid base encountered");
	}
	rs->arena_used = int64_t(c[0]);
	if (int64_t(c[0]) <= rs->mins_capacity) break;
	// exact need is known now; appended reads get the expected density of their bases on top
	want = int64_t(c[0]) + 1024 + int64_t(double(nb - rs->n_bases) * (2.0 * 2.0 / (W + 1)) * 1.1);
    }
    if (d_skip) ICB_CUDA(ctx, cudaFreeAsync(d_skip, ctx->stream));
    // CalcErrorRate on the compressed qualities (qualscore.cpp:78): FP64 chain per read, thread per read
    int rc = icb_launch_error_rate(rs, rs->d_hpc_qual, rs->d_offs, rs->d_hpc_len, rs->d_status, qualtab_nomin, rs->d_hpc_err);
    if (rc) return rc;
    rs->sketched = true;
    rs->sk_k = k;
    rs->sk_w = w;
    rs->total_f = rs->total_r = -1;
    rs->host_ready = false;
    return ICB_OK;
}

static int sketch_counts(icb_readset* rs, std::vector<int32_t>& nf, std::vector<int32_t>& nr)
{
    icb_ctx* ctx = rs->ctx;
    nf.resize(rs->n_reads);
    nr.resize(rs->n_reads);
    if (rs->n_reads) {
	ICB_CUDA(ctx, cudaMemcpyAsync(nf.data(), rs->d_nf, sizeof(int32_t) * rs->n_reads, cudaMemcpyDeviceToHost, ctx->stream));
	ICB_CUDA(ctx, cudaMemcpyAsync(nr.data(), rs->d_nr, sizeof(int32_t) * rs->n_reads, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ICB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ICB_OK;
}

int icb_sketch_sizes(icb_readset* rs, int64_t* n_minimizers, int64_t* n_rev_minimizers)
{
    if (!rs) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (!rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_sketch_sizes before icb_sketch_reads");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (rs->total_f < 0) {
	std::vector<int32_t> nf, nr;
	int rc = sketch_counts(rs, nf, nr);
	if (rc) return rc;
	int64_t a = 0, b = 0;
	for (int64_t i = 0; i < rs->n_reads; i++) {
	    a += nf[i];
	    b += nr[i];
	}
	rs->total_f = a;
	rs->total_r = b;
    }
    if (n_minimizers) *n_minimizers = rs->total_f;
    if (n_rev_minimizers) *n_rev_minimizers = rs->total_r;
    return ICB_OK;
}

int icb_sketch_download(icb_readset* rs, int32_t* status, int32_t* hpc_len, double* hpc_err, char* hpc_seq,
			char* hpc_qual, icb_minimizer* mins, int64_t* min_offs, icb_minimizer* rev_mins,
			int64_t* rev_offs)
{
    if (!rs) return ICB_ERR_ARG;
    icb_ctx* ctx = rs->ctx;
    if (!rs->sketched) return ctx->fail(ICB_ERR_STATE, "icb_sketch_download before icb_sketch_reads");
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = rs->n_reads;
    cudaStream_t st = ctx->stream;
    if (status && n) ICB_CUDA(ctx, cudaMemcpyAsync(status, rs->d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (hpc_len && n) ICB_CUDA(ctx, cudaMemcpyAsync(hpc_len, rs->d_hpc_len, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    if (hpc_err && n) ICB_CUDA(ctx, cudaMemcpyAsync(hpc_err, rs->d_hpc_err, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (hpc_seq && rs->n_bases) ICB_CUDA(ctx, cudaMemcpyAsync(hpc_seq, rs->d_hpc_seq, rs->n_bases, cudaMemcpyDeviceToHost, st));
    if (hpc_qual && rs->n_bases) ICB_CUDA(ctx, cudaMemcpyAsync(hpc_qual, rs->d_hpc_qual, rs->n_bases, cudaMemcpyDeviceToHost, st));
    if (mins || rev_mins || min_offs || rev_offs) {
	std::vector<int32_t> nf, nr;
	int rc = sketch_counts(rs, nf, nr);
	if (rc) return rc;
	std::vector<int64_t> of(n + 1), orv(n + 1);
	of[0] = orv[0] = 0;
	for (int64_t i = 0; i < n; i++) {
	    of[i + 1] = of[i] + nf[i];
	    orv[i + 1] = orv[i] + nr[i];
	}
	rs->total_f = of[n];
	rs->total_r = orv[n];
	if (min_offs) memcpy(min_offs, of.data(), sizeof(int64_t) * (n + 1));
	if (rev_offs) memcpy(rev_offs, orv.data(), sizeof(int64_t) * (n + 1));
	if ((mins || rev_mins) && n) {
	    int64_t *d_of = nullptr, *d_or = nullptr;
	    icb_minimizer *d_f = nullptr, *d_r = nullptr;
	    ICB_CUDA(ctx, cudaMallocAsync((void**)&d_of, sizeof(int64_t) * (n + 1), st));
	    ICB_CUDA(ctx, cudaMallocAsync((void**)&d_or, sizeof(int64_t) * (n + 1), st));
	    ICB_CUDA(ctx, cudaMallocAsync((void**)&d_f, sizeof(icb_minimizer) * (of[n] + 1), st));
	    ICB_CUDA(ctx, cudaMallocAsync((void**)&d_r, sizeof(icb_minimizer) * (orv[n] + 1), st));
	    ICB_CUDA(ctx, cudaMemcpyAsync(d_of, of.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
	    ICB_CUDA(ctx, cudaMemcpyAsync(d_or, orv.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
	    unsigned grid = unsigned((n * 32 + 255) / 256);
	    ICB_LAUNCH(ctx, "gather_minimizers",
		       icb_k_gather_mins<<<grid, 256, 0, st>>>(rs->d_mins, rs->d_min_off, rs->d_nf, rs->d_nr, d_of, d_or, n, d_f, d_r));
	    if (mins && of[n]) ICB_CUDA(ctx, cudaMemcpyAsync(mins, d_f, sizeof(icb_minimizer) * of[n], cudaMemcpyDeviceToHost, st));
	    if (rev_mins && orv[n]) ICB_CUDA(ctx, cudaMemcpyAsync(rev_mins, d_r, sizeof(icb_minimizer) * orv[n], cudaMemcpyDeviceToHost, st));
	    ICB_CUDA(ctx, cudaStreamSynchronize(st));
	    ICB_CUDA(ctx, cudaFreeAsync(d_of, st));
	    ICB_CUDA(ctx, cudaFreeAsync(d_or, st));
	    ICB_CUDA(ctx, cudaFreeAsync(d_f, st));
	    ICB_CUDA(ctx, cudaFreeAsync(d_r, st));
	}
    }
    ICB_CUDA(ctx, cudaStreamSynchronize(st));
    return ICB_OK;
}

}  // extern "C"
