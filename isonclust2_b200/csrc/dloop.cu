// Device-resident greedy loop: the decisions of ClusterSortedReads for one chunk of sorted reads, taken by kernels.
//
// Mirrors, per read, getBestCluster (reference cluster.cpp:530-568) = getBestClusterMapping (:355-406) on the candidate
// records of the pool-mode query (index.cu), then getBestClusterAln (:461-515) over the candidates tied at the top count,
// in the reference's candidate order (SortMinimizerHits, :622-636: iteration order of the hit map + libstdc++'s unstable
// std::sort, ref_order.h / ref_order_hd.h), consuming alignment results as the batched DP kernels deliver them; and the
// commit of ClusterSortedReads' loop body (:163-261): join, or found a cluster (AddMinimizers happens at chunk end).
//
//   icb_k_dl_init        resets the chunk's per-position records
//   icb_k_dl_decide      one CTA per undecided / invalidated read: decision under the founders committed so far
//   icb_k_dl_commit      one CTA walks the chunk in order, commits every decided read up to the first that still waits
//                        for device work (re-deciding inline the reads a founder committed in this walk invalidates)
//   icb_k_dl_gather      builds the round's work queues: alignment pair descriptors (trace space handed out in chunk
//                        order) and the reads whose FULL candidate list is needed (tied candidates -> candidate order)
//   icb_k_dl_finalize    founders of the chunk -> insert list for the main index, representative table
//
// Exactness: a decision is a pure function of (the read's candidate records, which provisional founders among them are
// committed, the alignment results consumed so far); `depmask` lists the provisional founders a read's records mention,
// and a read is re-decided whenever one of them was committed after its last decision.  Alignments requested for reads
// behind the commit frontier are speculative: their results are keyed by candidate and stay valid whatever is committed.
#include <cub/block/block_scan.cuh>

#include "dloop.cuh"
#include "ref_order_hd.h"

namespace {

typedef icb_sort_hd::Key Key;
constexpr int DL_CSM = 512;  // candidate records of one read staged in shared memory

struct DlShared {
    int cnt, n_group, req, target, ok;
};

__device__ __forceinline__ bool dl_real(const IcbDLoopView& V, int C0, int id) { return id < C0 || V.founder_cls[id - C0] >= 0; }
__device__ __forceinline__ int dl_real_id(const IcbDLoopView& V, int C0, int id) { return id < C0 ? id : V.founder_cls[id - C0]; }

__device__ __forceinline__ bool dl_dirty(const IcbDLoopView& V, int pos)
{
    const uint32_t* dep = V.depmask + int64_t(pos) * V.mw;
    const uint32_t* seen = V.seenmask + int64_t(pos) * V.mw;
    const volatile uint32_t* com = V.st->commask;
    uint32_t d = 0;
    for (int w = 0; w < V.mw; w++) d |= dep[w] & com[w] & ~seen[w];
    return d != 0;
}

// getBestClusterAln's walk (cluster.cpp:480-512) over the read's alignment plan, as far as results are available.
// One thread.
__device__ void dl_advance(const IcbDLoopView& V, int pos)
{
    const IcbQueryHeader h = V.hdr[pos];
    const int C0 = V.st->C0;
    const int n = V.plan_n[pos];
    int i = V.plan_i[pos], calls = V.aln_calls[pos];
    const int r = V.chunk_reads[pos];
    const unsigned readLen = unsigned(V.offs[r + 1] - V.offs[r]);
    uint8_t out = DL_NEW;
    while (i < n) {
	const int slot = V.plan[h.cand_off + i];
	const int a = V.aln[h.cand_off + slot];
	if (a < 0) {
	    V.need_slot[pos] = slot;
	    out = DL_NEED_PAIR;
	    break;
	}
	calls++;
	if (double(a) / double(readLen) >= V.aligned_th) {  // cluster.cpp:458, 503
	    const IcbCand c = V.cands[h.cand_off + slot];
	    V.dec_cls[pos] = dl_real_id(V, C0, c.id);
	    V.dec_strand[pos] = c.strand;
	    out = DL_JOIN;
	    break;
	}
	i++;
    }
    V.plan_i[pos] = i;
    V.aln_calls[pos] = calls;
    if (out != DL_NEED_PAIR) V.flags[pos] |= 1;	 // usedAln: the reference counts ALN_INVOKED once per read (cluster.cpp:563)
    V.outcome[pos] = out;
}

__device__ void dl_bitonic(Key* a, int NP)
{
    for (int k = 2; k <= NP; k <<= 1)
	for (int j = k >> 1; j > 0; j >>= 1) {
	    for (int i = threadIdx.x; i < (NP >> 1); i += blockDim.x) {
		const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
		const int hi = lo + j;
		const bool asc = (lo & k) == 0;
		const Key x = a[lo], y = a[hi];
		if ((x > y) == asc) {
		    a[lo] = y;
		    a[hi] = x;
		}
	    }
	    __syncthreads();
	}
}

// The members of the group Size == target of the read's candidates, in the order the reference's sorted `order` vector
// lists them (ref_order.h), as slots of the read's candidate list written to V.plan[cand_off ...]; S.n_group = how many.
//   insertion order of GetMinimizerHits (minimizer.cpp:44-76): (strand: +1 first, Index of the first hit, cluster id)
//   iteration order of libstdc++'s unordered_map: buckets in reverse order of their first insertion, newest node first
//   inside a bucket (no rehash while the number of keys stays within the bucket count)
//   std::sort by Size, descending: ref_order_hd.h
// Returns false (and raises the chunk's fallback flag) if the list does not fit the shared-memory plan or the reference's
// container would rehash.
__device__ bool dl_ref_order(const IcbDLoopView& V, int pos, int C0, const IcbCand* C, int nC, uint32_t target, DlShared& S, Key* keys,
			     uint32_t* t_id, uint32_t* t_ss, uint32_t* headbits)
{
    const int tid = threadIdx.x;
    const IcbCand* F = V.full + V.full_off[pos];
    const int nF = V.full_n[pos];
    const IcbQueryHeader h = V.hdr[pos];
    const uint32_t bkt = V.bkt[V.chunk_reads[pos]];
    if (tid == 0) {
	S.cnt = 0;
	S.n_group = 0;
    }
    __syncthreads();
    for (int i = tid; i < nF; i += blockDim.x) {
	const IcbCand c = F[i];
	if (!dl_real(V, C0, c.id)) continue;
	const int e = atomicAdd(&S.cnt, 1);
	if (e < DL_NMAX)
	    keys[e] = (Key(c.strand < 0) << 61) | (Key(uint32_t(c.tmapped) & 0x7fffu) << 46) | (Key(uint32_t(dl_real_id(V, C0, c.id))) << 15) |
		      Key(uint32_t(c.count) & 0x7fffu);
    }
    __syncthreads();
    const int n = S.cnt;
    if (n > DL_NMAX || uint32_t(n) > bkt) {
	if (tid == 0) atomicMax(&V.st->fallback, n > DL_NMAX ? DL_FB_TOO_MANY : DL_FB_REHASH);
	return false;
    }
    int NP = 2;
    while (NP < n) NP <<= 1;
    for (int i = n + tid; i < NP; i += blockDim.x) keys[i] = ~Key(0);
    __syncthreads();
    // ---- insertion rank t
    dl_bitonic(keys, NP);
    const uint32_t two64 = uint32_t((((~uint64_t(0)) % bkt) + 1) % bkt);
    for (int t = tid; t < n; t += blockDim.x) {
	const Key k1 = keys[t];
	const uint32_t id = uint32_t(k1 >> 15) & 0x7fffffffu, size = uint32_t(k1) & 0x7fffu, sb = uint32_t(k1 >> 61) & 1u;
	t_id[t] = id;
	t_ss[t] = size << 1 | sb;
	// bucket = size_t(int(id * strand)) % bkt (StrandedClsHash, minimizer.h:53-58, sign-extended)
	const uint32_t r = id % bkt;
	uint32_t b = r;
	if (sb && id != 0) b = two64 >= r ? two64 - r : two64 + bkt - r;
	keys[t] = Key(b) << 32 | Key(uint32_t(t));
    }
    for (int i = tid; i < (NP + 31) / 32; i += blockDim.x) headbits[i] = 0;
    __syncthreads();
    // ---- first insertion of every bucket
    dl_bitonic(keys, NP);
    for (int p = tid; p < n; p += blockDim.x)
	if (p == 0 || uint32_t(keys[p - 1] >> 32) != uint32_t(keys[p] >> 32)) atomicOr(&headbits[p >> 5], 1u << (p & 31));
    __syncthreads();
    for (int p = tid; p < n; p += blockDim.x) {
	if (!(headbits[p >> 5] >> (p & 31) & 1u)) continue;
	const uint32_t fs = uint32_t(keys[p]);	// the segment is sorted by t: its head is the bucket's oldest node
	int q = p;
	do {
	    const uint32_t t = uint32_t(keys[q]);
	    keys[q] = Key(0x7fffffffu - fs) << 32 | Key(0x7fffffffu - t);
	    q++;
	} while (q < n && !(headbits[q >> 5] >> (q & 31) & 1u));
    }
    __syncthreads();
    // ---- iteration order: first-seen descending, then newest first
    dl_bitonic(keys, NP);
    for (int p = tid; p < n; p += blockDim.x) {
	const uint32_t t = 0x7fffffffu - uint32_t(keys[p]);
	keys[p] = Key(t_ss[t] >> 1) << 32 | Key(t);
    }
    __syncthreads();
    // ---- std::sort, then the group in its final order -> slots of the candidate list
    if (tid == 0) {
	icb_sort_hd::introsort_group(keys, long(n), target);
	int g = 0;
	for (int p = 0; p < n; p++) {
	    if (uint32_t(keys[p] >> 32) != target) continue;
	    const uint32_t t = uint32_t(keys[p]);
	    const int id = int(t_id[t]), strand = (t_ss[t] & 1u) ? -1 : 1;
	    for (int s = 0; s < nC; s++)
		if (C[s].strand == strand && dl_real(V, C0, C[s].id) && dl_real_id(V, C0, C[s].id) == id) {
		    V.plan[h.cand_off + g++] = s;
		    break;
		}
	}
	S.n_group = g;
	atomicAdd((unsigned long long*)&V.st->order_calls, 1ull);
    }
    __syncthreads();
    return true;
}

// getBestCluster for the read at `pos` under the founders committed so far.  Whole CTA; the caller has checked that the
// read needs a decision.  Dynamic shared memory: DL_CSM candidate records, then the order emulation's arrays.
__device__ void dl_decide(const IcbDLoopView& V, int pos, unsigned char* smem)
{
    __shared__ DlShared S;
    const int tid = threadIdx.x;
    IcbCand* sC = reinterpret_cast<IcbCand*>(smem);
    Key* keys = reinterpret_cast<Key*>(smem + sizeof(IcbCand) * DL_CSM);
    uint32_t* t_id = reinterpret_cast<uint32_t*>(keys + DL_NMAX);
    uint32_t* t_ss = t_id + DL_NMAX;
    uint32_t* headbits = t_ss + DL_NMAX;
    const int C0 = V.st->C0;
    const IcbQueryHeader h = V.hdr[pos];
    const int nC = h.n_cand;
    const int r = V.chunk_reads[pos];
    const IcbCand* C = V.cands + h.cand_off;
    if (nC <= DL_CSM) {
	for (int i = tid; i < nC; i += blockDim.x) sC[i] = C[i];
	C = sC;
    }
    for (int w = tid; w < V.mw; w += blockDim.x) V.seenmask[int64_t(pos) * V.mw + w] = ((const volatile uint32_t*)V.st->commask)[w];
    __syncthreads();
    const bool mappingMode = V.mode == 0 || V.mode == 1;
    const double Lh = double(size_t(V.hpc_len[r]));
    auto passes = [&](const IcbCand& c) {
	const float mr = float(double(c.tmapped) / Lh);	 // cluster.cpp:351, 390: the ratio is narrowed to float
	return double(mr) >= V.mapped_th;
    };
    if (tid == 0) {
	atomicAdd((unsigned long long*)&V.st->decisions, 1ull);
	V.flags[pos] = 0;
	V.plan_n[pos] = 0;
	V.plan_i[pos] = 0;
	V.aln_calls[pos] = 0;
	V.dec_cls[pos] = -1;
	V.dec_strand[pos] = 0;
	int top = h.top_snap;
	for (int s = 0; s < nC; s++)
	    if (dl_real(V, C0, C[s].id)) top = max(top, C[s].count);
	S.req = 0;
	S.target = top;
	uint8_t out = DL_NONE;
	if (top == 0) out = DL_NEW;  // no hits at all (cluster.cpp:541-543)
	if (out == DL_NONE && mappingMode && !(unsigned(top) < unsigned(V.min_shared))) {
	    const int thr = int(double(unsigned(top)) * V.min_frac);  // cluster.cpp:387
	    // the reference walks the candidates by descending count and takes the first that passes: the passing
	    // candidates with the largest count decide; several of them -> the reference's order among equals is needed
	    int bestc = -1, npass = 0, pslot = -1;
	    for (int s = 0; s < nC; s++) {
		const IcbCand c = C[s];
		if (c.count < thr || c.count < bestc || !dl_real(V, C0, c.id) || !passes(c)) continue;
		if (c.count > bestc) {
		    bestc = c.count;
		    npass = 1;
		    pslot = s;
		}
		else
		    npass++;
	    }
	    if (npass == 1) {
		V.dec_cls[pos] = dl_real_id(V, C0, C[pslot].id);
		V.dec_strand[pos] = C[pslot].strand;
		out = DL_JOIN;
	    }
	    else if (npass > 1) {
		S.req = 1;
		S.target = bestc;
	    }
	}
	if (out == DL_NONE && S.req == 0) {
	    if (unsigned(top) < unsigned(V.min_shared)) out = DL_NEW;  // cluster.cpp:553-556
	    else if (V.mode == 1) out = DL_NEW;			       // fast: minimizers only
	    else {
		// getBestClusterAln: the candidates tied with the top count, in the reference's order (cluster.cpp:480-512)
		int ntied = 0, tslot = -1;
		for (int s = 0; s < nC; s++)
		    if (C[s].count == top && dl_real(V, C0, C[s].id)) {
			ntied++;
			tslot = s;
		    }
		if (ntied == 1) {
		    V.plan[h.cand_off] = tslot;
		    V.plan_n[pos] = 1;
		    S.req = 3;
		}
		else
		    S.req = 2;
	    }
	}
	V.outcome[pos] = out;
    }
    __syncthreads();
    const int req = S.req;
    if (req == 0) return;
    if (req == 3) {
	if (tid == 0) dl_advance(V, pos);
	__syncthreads();
	return;
    }
    // ---- the reference's candidate order is needed: from the read's full candidate list
    if (V.full_off[pos] < 0) {
	if (tid == 0) V.outcome[pos] = DL_NEED_FULL;
	__syncthreads();
	return;
    }
    if (!dl_ref_order(V, pos, C0, C, nC, uint32_t(S.target), S, keys, t_id, t_ss, headbits)) {
	if (tid == 0) V.outcome[pos] = DL_NEED_FULL;  // the chunk goes back to the host loop (fallback flag is up)
	__syncthreads();
	return;
    }
    if (tid == 0) {
	V.flags[pos] |= 2;  // usedTie
	const int g = S.n_group;
	if (req == 1) {
	    uint8_t out = DL_NEW;  // unreachable: the group holds the passing candidates
	    for (int i = 0; i < g; i++) {
		const IcbCand c = C[V.plan[h.cand_off + i]];
		if (passes(c)) {
		    V.dec_cls[pos] = dl_real_id(V, C0, c.id);
		    V.dec_strand[pos] = c.strand;
		    out = DL_JOIN;
		    break;
		}
	    }
	    V.outcome[pos] = out;
	}
	else {
	    V.plan_n[pos] = g;
	    dl_advance(V, pos);
	}
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_init(IcbDLoopView V, int n_chunk, int C0, int next_cls, uint32_t cand_cap,
							      unsigned long long full_cap, unsigned long long tiles_cap,
							      unsigned long long words_cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
	IcbDLoopState& s = *V.st;
	s.n_chunk = n_chunk;
	s.C0 = C0;
	s.commit_pos = 0;
	s.next_cls = next_cls;
	s.n_founders = 0;
	s.done = 0;
	s.fallback = (V.force_fallback > 0 && (C0 + n_chunk) % V.force_fallback == 0) ? DL_FB_FORCED : 0;
	s.rounds = 0;
	s.cand_cursor = 0;
	s.cand_cap = cand_cap;
	s.full_cursor = 0;
	s.full_cap = full_cap;
	s.n_pairs16 = s.n_pairs32 = s.n_full = 0;
	s.progress = 0;
	s.tiles_cap = tiles_cap;
	s.words_cap = words_cap;
	s.aln_invoked = s.aln_calls_ref = s.pairs_aligned = s.cells_aligned = s.tie_fallbacks = 0;
	s.decisions = s.order_calls = s.hits = s.cands = 0;
	for (int w = 0; w < DL_MAXW; w++) s.commask[w] = 0;
    }
    if (i < n_chunk) {
	V.founder_cls[i] = -1;
	V.outcome[i] = DL_NONE;
	V.flags[i] = 0;
	V.plan_n[i] = 0;
	V.plan_i[i] = 0;
	V.aln_calls[i] = 0;
	V.full_off[i] = -1;
	V.full_n[i] = 0;
	V.out_cls[i] = -1;
	V.out_strand[i] = 0;
	for (int w = 0; w < V.mw; w++) {
	    V.depmask[int64_t(i) * V.mw + w] = 0;
	    V.seenmask[int64_t(i) * V.mw + w] = 0;
	}
    }
}

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_decide(IcbDLoopView V)
{
    extern __shared__ __align__(16) unsigned char dl_smem[];
    const volatile IcbDLoopState* st = V.st;
    if (st->done || st->fallback) return;
    const int pos = blockIdx.x;
    if (pos < st->commit_pos || pos >= st->n_chunk) return;
    const uint8_t o = V.outcome[pos];
    if (o == DL_NONE || dl_dirty(V, pos) || (o == DL_NEED_FULL && V.full_off[pos] >= 0)) dl_decide(V, pos, dl_smem);
    else if (o == DL_NEED_PAIR) {
	if (threadIdx.x == 0) dl_advance(V, pos);
    }
}

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_commit(IcbDLoopView V)
{
    extern __shared__ __align__(16) unsigned char dl_smem[];
    IcbDLoopState* st = V.st;
    const volatile IcbDLoopState* vst = V.st;
    if (vst->done || vst->fallback) return;
    const int N = vst->n_chunk;
    int pos = vst->commit_pos;
    const int pos0 = pos;
    while (pos < N) {
	uint8_t o = *(volatile uint8_t*)&V.outcome[pos];
	if (o == DL_NONE || dl_dirty(V, pos) || (o == DL_NEED_FULL && V.full_off[pos] >= 0)) dl_decide(V, pos, dl_smem);
	else if (o == DL_NEED_PAIR) {
	    if (threadIdx.x == 0) dl_advance(V, pos);
	    __syncthreads();
	}
	if (vst->fallback) break;
	o = *(volatile uint8_t*)&V.outcome[pos];
	if (o != DL_JOIN && o != DL_NEW) break;
	if (threadIdx.x == 0) {
	    const uint8_t f = V.flags[pos];
	    if (f & 1) {
		st->aln_invoked++;
		st->aln_calls_ref += V.aln_calls[pos];
	    }
	    if (f & 2) st->tie_fallbacks++;
	    if (o == DL_NEW) {
		// new cluster (cluster.cpp:177-222); its minimizers join the main index at chunk end
		const int cls = st->next_cls++;
		V.founder_cls[pos] = cls;
		V.founder_pos[st->n_founders++] = pos;
		st->commask[pos >> 5] |= 1u << (pos & 31);
		V.out_cls[pos] = cls;
		V.out_strand[pos] = 1;
	    }
	    else {
		V.out_cls[pos] = V.dec_cls[pos];
		V.out_strand[pos] = V.dec_strand[pos];
	    }
	    __threadfence();
	}
	__syncthreads();
	pos++;
    }
    if (threadIdx.x == 0) {
	st->commit_pos = pos;
	st->rounds++;
	st->progress = pos > pos0;
	if (pos >= N) st->done = 1;
    }
}

constexpr int GA_ITEMS = DL_MAXW * 32 / DL_THREADS;  // chunk positions per thread of the gather kernel

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_gather(IcbDLoopView V, int max_pairs)
{
    typedef cub::BlockScan<long long, DL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    IcbDLoopState* st = V.st;
    if (threadIdx.x == 0) st->n_pairs16 = st->n_pairs32 = st->n_full = 0;
    if (st->done || st->fallback) return;
    const int N = st->n_chunk, C0 = st->C0, cp = st->commit_pos;
    const int k = V.k;
    // what every undecided read waits for
    int kind[GA_ITEMS];  // 0 nothing, 1 pair (16-bit kernel), 2 pair (32-bit kernel), 3 full list
    long long tiles[GA_ITEMS], words[GA_ITEMS];
    IcbPairDesc d[GA_ITEMS];
    for (int q = 0; q < GA_ITEMS; q++) {
	const int pos = cp + threadIdx.x * GA_ITEMS + q;
	kind[q] = 0;
	tiles[q] = words[q] = 0;
	if (pos >= N) continue;
	const uint8_t o = V.outcome[pos];
	if (o == DL_NEED_FULL && V.full_off[pos] < 0) {
	    kind[q] = 3;
	    tiles[q] = V.hdr[pos].n_total;
	}
	else if (o == DL_NEED_PAIR) {
	    const IcbQueryHeader h = V.hdr[pos];
	    const int slot = V.need_slot[pos];
	    const IcbCand c = V.cands[h.cand_off + slot];
	    const int r = V.chunk_reads[pos];
	    const int rep = c.id < C0 ? V.cls_founder[c.id] : V.chunk_reads[c.id - C0];
	    IcbPairDesc& p = d[q];
	    p.q_off = V.offs[r];
	    p.r_off = V.offs[rep];
	    p.n = int32_t(V.offs[r + 1] - V.offs[r]);
	    p.m = int32_t(V.offs[rep + 1] - V.offs[rep]);
	    p.rc = c.strand == -1;
	    const double e = V.raw_err[r] + V.raw_err[rep];		    // cluster.cpp:494-496: e1 (read) + e2 (representative)
	    p.open = e <= 0.01 ? 5 : (e <= 0.04 ? 4 : (e <= 0.1 ? 3 : 2));  // setGapOpen, cluster.cpp:425-440
	    const double lim = floor((1.0 - e) * double(unsigned(k)));	    // getAlnRatio, cluster.cpp:446
	    p.limit = lim < -1.0 ? -1 : (lim > double(k + 1) ? k + 1 : int(lim));
	    p.out = int32_t(h.cand_off + slot);
	    p.pad = 0;
	    const bool ok16 = max(p.n, p.m) <= 8000 && V.aln[h.cand_off + slot] != DL_ALN_REDO32;
	    if (ok16) {
		p.ncg = (p.m + 71) >> 3;
		kind[q] = 1;
		words[q] = p.m + 8;
	    }
	    else {
		p.ncg = (p.m + 7) >> 3;
		kind[q] = 2;
		words[q] = 3ll * p.m + ((p.n + 7) & ~7) + 8;
	    }
	    tiles[q] = (long long)((p.n + 7) >> 3) * p.ncg;
	}
    }
    // chunk-order prefixes: the alignment workspace and the queues are handed out from the commit frontier on, so the read
    // the loop waits for is always served; what does not fit waits for the next round
    long long v[GA_ITEMS], pre[GA_ITEMS], tot;
    long long c16[GA_ITEMS], t16[GA_ITEMS], w16[GA_ITEMS], c32[GA_ITEMS], t32[GA_ITEMS], w32[GA_ITEMS], cf[GA_ITEMS], rf[GA_ITEMS];
#define DL_SCAN(dst, expr)                                  \
    for (int q = 0; q < GA_ITEMS; q++) v[q] = (expr);      \
    Scan(tmp).ExclusiveSum(v, pre, tot);                    \
    __syncthreads();                                        \
    for (int q = 0; q < GA_ITEMS; q++) dst[q] = pre[q];
    DL_SCAN(c16, kind[q] == 1)
    DL_SCAN(t16, kind[q] == 1 ? tiles[q] : 0)
    DL_SCAN(w16, kind[q] == 1 ? words[q] : 0)
    DL_SCAN(c32, kind[q] == 2)
    DL_SCAN(t32, kind[q] == 2 ? tiles[q] : 0)
    DL_SCAN(w32, kind[q] == 2 ? words[q] : 0)
    DL_SCAN(cf, kind[q] == 3)
    DL_SCAN(rf, kind[q] == 3 ? tiles[q] : 0)
#undef DL_SCAN
    const unsigned long long full_base = st->full_cursor;
    int n16 = 0, n32 = 0, nfull = 0;
    long long cells = 0, full_used = 0;
    bool full_over = false;
    for (int q = 0; q < GA_ITEMS; q++) {
	const int pos = cp + threadIdx.x * GA_ITEMS + q;
	if (kind[q] == 1) {
	    if (c16[q] < max_pairs && (unsigned long long)(t16[q] + tiles[q]) <= st->tiles_cap &&
		(unsigned long long)(w16[q] + words[q]) <= st->words_cap) {
		d[q].trace_off = t16[q];
		d[q].line_off = w16[q];
		V.descs16[c16[q]] = d[q];
		n16++;
		cells += (long long)d[q].n * d[q].m;
	    }
	}
	else if (kind[q] == 2) {
	    if (c32[q] < max_pairs && (unsigned long long)(t32[q] + tiles[q]) <= st->tiles_cap &&
		(unsigned long long)(w32[q] + words[q]) <= st->words_cap) {
		d[q].trace_off = t32[q];
		d[q].line_off = w32[q];
		V.descs32[c32[q]] = d[q];
		n32++;
		cells += (long long)d[q].n * d[q].m;
	    }
	}
	else if (kind[q] == 3) {
	    if (full_base + (unsigned long long)(rf[q] + tiles[q]) <= st->full_cap) {
		V.full_list[cf[q]] = pos;
		V.full_list_off[cf[q]] = int64_t(full_base) + int64_t(rf[q]);
		nfull++;
		full_used = max(full_used, rf[q] + tiles[q]);
	    }
	    else
		full_over = true;
	}
    }
    // accepted entries form a prefix of each queue (the capacity tests are monotone along the prefix sums)
    if (n16) atomicAdd(&st->n_pairs16, n16);
    if (n32) atomicAdd(&st->n_pairs32, n32);
    if (nfull) atomicAdd(&st->n_full, nfull);
    if (cells) atomicAdd((unsigned long long*)&st->cells_aligned, (unsigned long long)cells);
    if (n16 + n32) atomicAdd((unsigned long long*)&st->pairs_aligned, (unsigned long long)(n16 + n32));
    if (full_used) atomicMax(&st->full_cursor, full_base + (unsigned long long)full_used);
    if (full_over) atomicMax(&st->fallback, DL_FB_FULLPOOL);
    __syncthreads();
    if (threadIdx.x == 0 && !st->progress && st->n_pairs16 + st->n_pairs32 + st->n_full == 0 && !st->fallback)
	st->fallback = DL_FB_STUCK;  // nothing committed and nothing to wait for: never expected; the host loop takes the chunk
}

// chunk end: the founders' insert list for the main index (AddMinimizers, cluster.cpp:178-180), their representative
// reads, and the chunk's query totals
__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_finalize(IcbDLoopView V, int* n_ins)
{
    IcbDLoopState* st = V.st;
    if (threadIdx.x == 0) *n_ins = 0;
    if (!st->done || st->fallback) return;
    const int nfo = st->n_founders;
    for (int i = threadIdx.x; i < nfo; i += blockDim.x) {
	const int pos = V.founder_pos[i];
	const int cls = V.founder_cls[pos];
	V.ins_reads[i] = V.chunk_reads[pos];
	V.ins_ids[i] = cls;
	V.cls_founder[cls] = V.chunk_reads[pos];
    }
    long long hits = 0, cands = 0;
    for (int i = threadIdx.x; i < st->n_chunk; i += blockDim.x) {
	hits += V.hdr[i].hits;
	cands += V.hdr[i].n_cand;
    }
    if (hits) atomicAdd((unsigned long long*)&st->hits, (unsigned long long)hits);
    if (cands) atomicAdd((unsigned long long*)&st->cands, (unsigned long long)cands);
    __syncthreads();
    if (threadIdx.x == 0) {
	*n_ins = nfo;
	st->done = 2;  // finalized
    }
}

}  // namespace

size_t icb_dl_smem_bytes() { return sizeof(IcbCand) * DL_CSM + sizeof(Key) * DL_NMAX + 2 * sizeof(uint32_t) * DL_NMAX + DL_NMAX / 8 + 64; }

static cudaError_t dl_attr_once()
{
    static cudaError_t err = [] {
	cudaError_t e = cudaFuncSetAttribute(icb_k_dl_decide, cudaFuncAttributeMaxDynamicSharedMemorySize, int(icb_dl_smem_bytes()));
	if (e == cudaSuccess) e = cudaFuncSetAttribute(icb_k_dl_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, int(icb_dl_smem_bytes()));
	return e;
    }();
    return err;
}

int icb_dl_init(icb_ctx* ctx, const IcbDLoopView& V, int n_chunk, int C0, uint32_t cand_cap, unsigned long long full_cap,
		unsigned long long tiles_cap, unsigned long long words_cap)
{
    ICB_CUDA(ctx, dl_attr_once());
    ICB_LAUNCH(ctx, "dl_init",
	       icb_k_dl_init<<<(n_chunk + DL_THREADS - 1) / DL_THREADS, DL_THREADS, 0, ctx->stream>>>(V, n_chunk, C0, C0, cand_cap, full_cap,
												       tiles_cap, words_cap));
    return ICB_OK;
}

int icb_dl_decide(icb_ctx* ctx, const IcbDLoopView& V, int n_chunk)
{
    ICB_LAUNCH(ctx, "dl_decide", icb_k_dl_decide<<<n_chunk, DL_THREADS, icb_dl_smem_bytes(), ctx->stream>>>(V));
    return ICB_OK;
}

int icb_dl_commit(icb_ctx* ctx, const IcbDLoopView& V)
{
    ICB_LAUNCH(ctx, "dl_commit", icb_k_dl_commit<<<1, DL_THREADS, icb_dl_smem_bytes(), ctx->stream>>>(V));
    return ICB_OK;
}

int icb_dl_gather(icb_ctx* ctx, const IcbDLoopView& V, int max_pairs)
{
    ICB_LAUNCH(ctx, "dl_gather", icb_k_dl_gather<<<1, DL_THREADS, 0, ctx->stream>>>(V, max_pairs));
    return ICB_OK;
}

int icb_dl_finalize(icb_ctx* ctx, const IcbDLoopView& V, int* d_n_ins)
{
    ICB_LAUNCH(ctx, "dl_finalize", icb_k_dl_finalize<<<1, DL_THREADS, 0, ctx->stream>>>(V, d_n_ins));
    return ICB_OK;
}
