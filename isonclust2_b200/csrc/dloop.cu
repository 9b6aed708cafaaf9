// Device-resident greedy loop: the decisions of ClusterSortedReads for one chunk of sorted reads, taken by kernels.
//
// Mirrors, per read, getBestCluster (reference cluster.cpp:530-568) = getBestClusterMapping (:355-406) on the candidate
// records of the pool-mode query (index.cu), then getBestClusterAln (:461-515) over the candidates tied at the top count,
// in the reference's candidate order (SortMinimizerHits, :622-636: iteration order of the hit map + libstdc++'s unstable
// std::sort, ref_order.h / ref_order_hd.h), consuming alignment results as the batched DP kernels deliver them; and the
// commit of ClusterSortedReads' loop body (:163-261): join, or found a cluster (AddMinimizers happens at chunk end).
//
//   icb_k_dl_init        resets the chunk's per-position records
//   icb_k_dl_decide      one warp per undecided / invalidated read: decision under the founders committed so far; reads
//                        whose decision needs the reference's candidate ORDER are queued for
//   icb_k_dl_order       one CTA per queued read: insertion order, hash-map iteration order and std::sort of the read's
//                        full candidate list in shared memory (two shared-memory plans: usual and large lists)
//   icb_k_dl_commit      one warp walks the chunk in order, commits every decided read up to the first that still waits
//                        for device work or for its candidate order (re-deciding inline what a new founder invalidated)
//   icb_k_dl_gather      builds the round's work queues: alignment pair descriptors (trace space handed out in chunk
//                        order) and the reads whose FULL candidate list is needed (tied candidates -> candidate order)
//   icb_k_dl_finalize    founders of the chunk -> insert list for the main index, representative table
//
// Exactness: a decision is a pure function of (the read's candidate records, which provisional founders among them are
// committed, the alignment results consumed so far); `depmask` lists the provisional founders a read's records mention,
// and a read is re-decided whenever one of them was committed after its last decision.  Alignments requested for reads
// behind the commit frontier are speculative: their results are keyed by candidate and stay valid whatever is committed.
#include <cub/block/block_scan.cuh>

#include <algorithm>

#include "residency.cuh"
#include "dloop.cuh"
#include "ref_order_hd.h"

namespace {

typedef icb_sort_hd::Key Key;
constexpr int DL_CSM = 512;  // candidate records of one read staged in shared memory by the order kernel
constexpr int DL_CW = 64;    // ... and by a warp of the decision kernels
constexpr int DL_WARPS = DL_THREADS / 32;
constexpr int DL_ORDER_GRID = 96;  // CTAs of a candidate-order launch (each loops over the queue)

struct DlShared {
    int cnt, n_group;
};

// Which provisional founders count, and under which cluster id.  EXACT: the founders committed so far (what the commit walk
// decides under).  Otherwise the PREDICTED view: also the reads behind the commit frontier whose current decision is "new
// cluster", with the ids they get if every prediction before them holds.  Decisions behind the frontier are taken under
// the predicted view, so that the alignments they ask for are the ones they will need once the predictions have come true;
// a decision only counts when the commit walk finds that the view it was taken under agrees with the committed founders.
__device__ __forceinline__ bool dl_real(const IcbDLoopView& V, int C0, int id, bool exact)
{
    return id < C0 || (exact ? V.founder_cls[id - C0] : V.view_cls[id - C0]) >= 0;
}
__device__ __forceinline__ int dl_real_id(const IcbDLoopView& V, int C0, int id, bool exact)
{
    return id < C0 ? id : (exact ? V.founder_cls[id - C0] : V.view_cls[id - C0]);
}

// Is the read's decision stale under the current view?  It is if the view it was taken under (seenmask) differs from the
// current one in a founder the read's candidate records mention (depmask), or — for a read that has chunk founders among
// its candidates at all — in ANY founder before it: cluster ids of chunk founders count the founders before them, and
// the reference's candidate order depends on the ids.
__device__ __forceinline__ bool dl_dirty(const IcbDLoopView& V, int pos, bool exact)
{
    const uint32_t* dep = V.depmask + int64_t(pos) * V.mw;
    const uint32_t* seen = V.seenmask + int64_t(pos) * V.mw;
    const volatile uint32_t* cur = exact ? V.st->commask : V.st->viewmask;
    uint32_t d1 = 0, d2 = 0, has = 0;
    const int wl = pos >> 5;
    for (int w = 0; w <= wl && w < V.mw; w++) {
	const uint32_t before = w < wl ? 0xffffffffu : ((1u << (pos & 31)) - 1u);
	const uint32_t c = cur[w], s = seen[w];
	const uint32_t diff = (c ^ s) & before;
	d1 |= diff & dep[w];
	d2 |= diff;
	has |= dep[w] & (c | s) & before;
    }
    return d1 != 0 || (has != 0 && d2 != 0);
}

// the mapping test of getBestClusterMapping for one candidate: the ratio is narrowed to float (cluster.cpp:351, 390)
__device__ __forceinline__ bool dl_passes(const IcbDLoopView& V, int r, const IcbCand& c)
{
    const float mr = float(double(c.tmapped) / double(size_t(V.hpc_len[r])));
    return double(mr) >= V.mapped_th;
}

// getBestClusterAln's walk (cluster.cpp:480-512) over the read's alignment plan, as far as results are available.
// One thread.
__device__ void dl_advance(const IcbDLoopView& V, int pos, bool exact)
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
	    // the alignment the walk waits for, and — ahead of its result — the next tied candidates' (the reference only aligns
	    // them if this one fails: speculative work, but one round trip instead of one per candidate)
	    int cnt = 0;
	    for (int j = i; j < n && cnt < V.spec_pairs; j++) {
		const int sj = V.plan[h.cand_off + j];
		if (V.aln[h.cand_off + sj] < 0) V.need_slot[pos * DL_SPEC + cnt++] = sj;
	    }
	    V.need_n[pos] = cnt;
	    out = DL_NEED_PAIR;
	    break;
	}
	calls++;
	if (double(a) / double(readLen) >= V.aligned_th) {  // cluster.cpp:458, 503
	    const IcbCand c = V.cands[h.cand_off + slot];
	    V.dec_cls[pos] = dl_real_id(V, C0, c.id, exact);
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

// libstdc++'s __introsort_loop (ref_order_hd.h: restricted to the ranges that hold the group Size == target) by the whole
// CTA.  What is sequential in the library — the unguarded Hoare partition — has a closed form: with L the positions
// (ascending) of the elements that stop the upward scan (not greater than the pivot) and R the positions (descending) of
// those that stop the downward scan (not smaller), the partition swaps exactly the pairs (L[i], R[i]) while L[i] < R[i]
// (i < s), and returns min(L[s], R[s-1]): both scans only ever see untouched elements between the last swapped pair, and
// the first swapped-in element beyond it stops them.  Ranks come from two block scans, the swaps are independent.
// Thread 0 keeps the library's control flow (median of three, depth budget, heap sort when it is used up).
__device__ void dl_introsort_cta(Key* keys, int n, uint32_t target, uint16_t* Lpos, uint16_t* Rpos)
{
    typedef cub::BlockScan<int, DL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int st_first[64], st_last[64], st_depth[64];
    __shared__ int s_sp, s_f, s_l, s_d, s_done, s_cnt;
    const int tid = threadIdx.x;
    if (tid == 0) {
	int lg = 0;
	for (unsigned v = unsigned(n); v > 1; v >>= 1) lg++;
	s_f = 0;
	s_l = n;
	s_d = 2 * lg;
	s_sp = 0;
	s_done = 0;
    }
    __syncthreads();
    for (;;) {
	const int first = s_f, last = s_l, depth = s_d;
	int descend = last - first > 16;
	if (descend) {
	    int h = 0;
	    for (int i = first + tid; i < last; i += DL_THREADS) h |= uint32_t(keys[i] >> 32) == target;
	    descend = __syncthreads_or(h);
	}
	if (descend && depth == 0) {
	    if (tid == 0) icb_sort_hd::heap_sort(keys + first, keys + last);
	    descend = 0;
	}
	if (descend) {
	    if (tid == 0) {
		icb_sort_hd::median_to_first(keys + first, keys + first + 1, keys + first + (last - first) / 2, keys + last - 1);
		s_cnt = 0;
	    }
	    __syncthreads();
	    const Key pivot = keys[first];
	    const int lo = first + 1;
	    const int per = (last - lo + DL_THREADS - 1) / DL_THREADS;
	    const int a = min(last, lo + tid * per), b = min(last, a + per);
	    int cL = 0, cR = 0;
	    for (int i = a; i < b; i++) {
		const Key k = keys[i];
		cL += !icb_sort_hd::gt(k, pivot);
		cR += !icb_sort_hd::gt(pivot, k);
	    }
	    int exL, totL, exR, totR;
	    Scan(tmp).ExclusiveSum(cL, exL, totL);
	    __syncthreads();
	    Scan(tmp).ExclusiveSum(cR, exR, totR);
	    __syncthreads();
	    int j = exL;
	    for (int i = a; i < b; i++)
		if (!icb_sort_hd::gt(keys[i], pivot)) Lpos[j++] = uint16_t(i);
	    j = totR - exR - cR;  // stoppers of the downward scan to the right of this thread's piece
	    for (int i = b - 1; i >= a; i--)
		if (!icb_sort_hd::gt(pivot, keys[i])) Rpos[j++] = uint16_t(i);
	    __syncthreads();
	    const int m = min(totL, totR);
	    int c = 0;
	    for (int i = tid; i < m; i += DL_THREADS) c += Lpos[i] < Rpos[i];
	    if (c) atomicAdd(&s_cnt, c);
	    __syncthreads();
	    const int s = s_cnt;  // L ascending, R descending: the pairs with L[i] < R[i] are a prefix
	    for (int i = tid; i < s; i += DL_THREADS) {
		const int x = Lpos[i], y = Rpos[i];
		const Key t = keys[x];
		keys[x] = keys[y];
		keys[y] = t;
	    }
	    __syncthreads();
	    if (tid == 0) {
		const int c1 = s < totL ? int(Lpos[s]) : 0x7fffffff, c2 = s > 0 ? int(Rpos[s - 1]) : 0x7fffffff;
		const int cut = min(c1, c2);
		// the library recurses into [cut, last) and goes on with [first, cut): disjoint ranges, any order
		st_first[s_sp] = first;
		st_last[s_sp] = cut;
		st_depth[s_sp] = depth - 1;
		s_sp++;
		s_f = cut;
		s_d = depth - 1;
	    }
	    __syncthreads();
	    continue;
	}
	__syncthreads();
	if (tid == 0) {
	    if (s_sp == 0) s_done = 1;
	    else {
		s_sp--;
		s_f = st_first[s_sp];
		s_l = st_last[s_sp];
		s_d = st_depth[s_sp];
	    }
	}
	__syncthreads();
	if (s_done) break;
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
template <int NMAX>
__device__ bool dl_ref_order(const IcbDLoopView& V, int pos, int C0, const IcbCand* C, int nC, uint32_t target, DlShared& S, Key* keys,
			     uint32_t* t_id, uint16_t* w16a, uint16_t* w16b, uint32_t* headbits)
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
	if (!dl_real(V, C0, c.id, false)) continue;
	const int e = atomicAdd(&S.cnt, 1);
	if (e < NMAX)
	    keys[e] = (Key(c.strand < 0) << 61) | (Key(uint32_t(c.tmapped) & 0x7fffu) << 46) | (Key(uint32_t(dl_real_id(V, C0, c.id, false))) << 15) |
		      Key(uint32_t(c.count) & 0x7fffu);
    }
    __syncthreads();
    const int n = S.cnt;
    if (n > NMAX || uint32_t(n) > bkt) {
	if (tid == 0) atomicMax(&V.st->fallback, n > NMAX ? DL_FB_TOO_MANY : DL_FB_REHASH);
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
	t_id[t] = id | sb << 31;
	w16a[t] = uint16_t(size);
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
	keys[p] = Key(w16a[t]) << 32 | Key(t);
    }
    __syncthreads();
    // ---- std::sort, then the group in its final order -> slots of the candidate list
    dl_introsort_cta(keys, n, target, w16a, w16b);
    if (tid == 0) {
	int g = 0;
	for (int p = 0; p < n; p++) {
	    if (uint32_t(keys[p] >> 32) != target) continue;
	    const uint32_t t = uint32_t(keys[p]);
	    const int id = int(t_id[t] & 0x7fffffffu), strand = (t_id[t] >> 31) ? -1 : 1;
	    for (int s = 0; s < nC; s++)
		if (C[s].strand == strand && dl_real(V, C0, C[s].id, false) && dl_real_id(V, C0, C[s].id, false) == id) {
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

// getBestCluster for the read at `pos` under the founders committed so far, by ONE WARP (lane 0 takes the decision on the
// candidate records the lanes staged).  A decision that needs the reference's candidate order is queued for the order
// kernel (may_queue) or left open (the commit walk stops there).  sC: DL_CW records of shared memory of this warp.
__device__ void dl_decide_light(const IcbDLoopView& V, int pos, IcbCand* sC, bool may_queue, bool exact)
{
    const int lane = threadIdx.x & 31;
    const int C0 = V.st->C0;
    const IcbQueryHeader h = V.hdr[pos];
    const int nC = h.n_cand;
    const int r = V.chunk_reads[pos];
    const IcbCand* C = V.cands + h.cand_off;
    if (nC <= DL_CW) {
	for (int i = lane; i < nC; i += 32) sC[i] = C[i];
	C = sC;
    }
    for (int w = lane; w < V.mw; w += 32)
	V.seenmask[int64_t(pos) * V.mw + w] = ((const volatile uint32_t*)(exact ? V.st->commask : V.st->viewmask))[w];
    __syncwarp();
    if (lane == 0) {
	const bool mappingMode = V.mode == 0 || V.mode == 1;
	atomicAdd((unsigned long long*)&V.st->decisions, 1ull);
	V.flags[pos] = 0;
	V.plan_n[pos] = 0;
	V.plan_i[pos] = 0;
	V.aln_calls[pos] = 0;
	V.dec_cls[pos] = -1;
	V.dec_strand[pos] = 0;
	int top = h.top_snap;
	for (int s = 0; s < nC; s++)
	    if (dl_real(V, C0, C[s].id, exact)) top = max(top, C[s].count);
	int req = 0, target = top;
	uint8_t out = DL_NONE;
	if (top == 0) out = DL_NEW;  // no hits at all (cluster.cpp:541-543)
	if (out == DL_NONE && mappingMode && !(unsigned(top) < unsigned(V.min_shared))) {
	    const int thr = int(double(unsigned(top)) * V.min_frac);  // cluster.cpp:387
	    // the reference walks the candidates by descending count and takes the first that passes: the passing
	    // candidates with the largest count decide; several of them -> the reference's order among equals is needed
	    int bestc = -1, npass = 0, pslot = -1;
	    for (int s = 0; s < nC; s++) {
		const IcbCand c = C[s];
		if (c.count < thr || c.count < bestc || !dl_real(V, C0, c.id, exact) || !dl_passes(V, r, c)) continue;
		if (c.count > bestc) {
		    bestc = c.count;
		    npass = 1;
		    pslot = s;
		}
		else
		    npass++;
	    }
	    if (npass == 1) {
		V.dec_cls[pos] = dl_real_id(V, C0, C[pslot].id, exact);
		V.dec_strand[pos] = C[pslot].strand;
		out = DL_JOIN;
	    }
	    else if (npass > 1) {
		req = 1;
		target = bestc;
	    }
	}
	if (out == DL_NONE && req == 0) {
	    if (unsigned(top) < unsigned(V.min_shared)) out = DL_NEW;  // cluster.cpp:553-556
	    else if (V.mode == 1) out = DL_NEW;			       // fast: minimizers only
	    else {
		// getBestClusterAln: the candidates tied with the top count, in the reference's order (cluster.cpp:480-512)
		int ntied = 0, tslot = -1;
		for (int s = 0; s < nC; s++)
		    if (C[s].count == top && dl_real(V, C0, C[s].id, exact)) {
			ntied++;
			tslot = s;
		    }
		if (ntied == 1) {
		    V.plan[h.cand_off] = tslot;
		    V.plan_n[pos] = 1;
		    req = 3;
		}
		else
		    req = 2;
	    }
	}
	if (req == 3) dl_advance(V, pos, exact);
	else {
	    if (req) {
		// the reference's candidate order is needed: from the read's full candidate list
		// (every founder committed before a read with tied candidates changes its order: far behind the frontier the
		// order is not worth computing yet — the read stays undecided until the frontier comes within spec_window)
		if (V.full_off[pos] < 0) out = DL_NEED_FULL;
		else if (!may_queue || pos - V.st->commit_pos > V.spec_window) out = DL_NONE;
		else {
		    const int n = V.full_n[pos];
		    V.ord_req[pos] = uint8_t(req);
		    V.ord_target[pos] = target;
		    out = DL_NEED_ORDER;
		    if (n <= DL_NMAX_S) V.order_list_s[atomicAdd(&V.st->n_order_s, 1)] = pos;
		    else if (n <= DL_NMAX_L) V.order_list_l[atomicAdd(&V.st->n_order_l, 1)] = pos;
		    else {
			atomicMax(&V.st->fallback, DL_FB_TOO_MANY);
			out = DL_NEED_FULL;
		    }
		}
	    }
	    V.outcome[pos] = out;
	}
    }
    __syncwarp();
}

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_init(IcbDLoopView V, int n_chunk, int C0, int next_cls, uint32_t cand_cap,
							      unsigned long long full_cap, unsigned long long tiles_cap,
							      unsigned long long words_cap)
{
    icb_res_mark(RES_INIT, V.descs16);
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
	s.n_order_s = s.n_order_l = 0;
	s.progress = 0;
	s.tiles_cap = tiles_cap;
	s.words_cap = words_cap;
	s.aln_invoked = s.aln_calls_ref = s.pairs_aligned = s.cells_aligned = s.tie_fallbacks = 0;
	s.decisions = s.order_calls = s.hits = s.cands = 0;
	s.stop_full = s.stop_pair_first = s.stop_pair_again = 0;
	for (int w = 0; w < DL_MAXW; w++) s.commask[w] = s.viewmask[w] = 0;
    }
    if (i < n_chunk) {
	V.founder_cls[i] = -1;
	V.view_cls[i] = -1;
	V.outcome[i] = DL_NONE;
	V.flags[i] = 0;
	V.plan_n[i] = 0;
	V.plan_i[i] = 0;
	V.aln_calls[i] = 0;
	V.full_off[i] = -1;
	V.full_n[i] = 0;
	V.req_count[i] = 0;
	V.out_cls[i] = -1;
	V.out_strand[i] = 0;
	for (int w = 0; w < V.mw; w++) {
	    V.depmask[int64_t(i) * V.mw + w] = 0;
	    V.seenmask[int64_t(i) * V.mw + w] = 0;
	}
    }
}

// one warp per chunk position
__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_decide(IcbDLoopView V)
{
    icb_res_mark(RES_DECIDE, V.descs16);
    __shared__ IcbCand s_c[DL_WARPS][DL_CW];
    const volatile IcbDLoopState* st = V.st;
    if (st->done || st->fallback) return;
    const int warp = threadIdx.x >> 5;
    const int pos = blockIdx.x * DL_WARPS + warp;
    if (pos < st->commit_pos || pos >= st->n_chunk) return;
    const uint8_t o = V.outcome[pos];
    if (o == DL_NONE || o == DL_NEED_ORDER || dl_dirty(V, pos, false) || (o == DL_NEED_FULL && V.full_off[pos] >= 0))
	dl_decide_light(V, pos, s_c[warp], true, false);
    else if (o == DL_NEED_PAIR) {
	if ((threadIdx.x & 31) == 0) dl_advance(V, pos, false);
    }
}

// the reads the decision kernel queued: candidate order of the group, then the rest of the decision
template <int NMAX>
__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_order(IcbDLoopView V, const int32_t* __restrict__ list, const int32_t* __restrict__ n_list)
{
    icb_res_mark(NMAX == DL_NMAX_S ? RES_ORDER : RES_ORDER_L, V.descs16);
    extern __shared__ __align__(16) unsigned char dl_smem[];
    __shared__ DlShared S;
    const volatile IcbDLoopState* st = V.st;
    if (st->done || st->fallback) return;
    const int tid = threadIdx.x;
    IcbCand* sC = reinterpret_cast<IcbCand*>(dl_smem);
    Key* keys = reinterpret_cast<Key*>(dl_smem + sizeof(IcbCand) * DL_CSM);
    uint32_t* t_id = reinterpret_cast<uint32_t*>(keys + NMAX);
    uint16_t* w16a = reinterpret_cast<uint16_t*>(t_id + NMAX);
    uint16_t* w16b = w16a + NMAX;
    uint32_t* headbits = reinterpret_cast<uint32_t*>(w16b + NMAX);
    const int n_items = *n_list;
    const int C0 = st->C0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
	const int pos = list[item];
	const IcbQueryHeader h = V.hdr[pos];
	const int nC = h.n_cand;
	const int r = V.chunk_reads[pos];
	const IcbCand* C = V.cands + h.cand_off;
	__syncthreads();  // the previous item's arrays are done with
	if (nC <= DL_CSM) {
	    for (int i = tid; i < nC; i += blockDim.x) sC[i] = C[i];
	    C = sC;
	}
	__syncthreads();
	const int req = V.ord_req[pos];
	const bool ok = dl_ref_order<NMAX>(V, pos, C0, C, nC, uint32_t(V.ord_target[pos]), S, keys, t_id, w16a, w16b, headbits);
	if (tid == 0) {
	    if (!ok) V.outcome[pos] = DL_NEED_FULL;  // the chunk goes back to the host loop (the fallback flag is up)
	    else {
		V.flags[pos] |= 2;  // usedTie
		const int g = S.n_group;
		if (req == 1) {
		    uint8_t out = DL_NEW;  // unreachable: the group holds the passing candidates
		    for (int i = 0; i < g; i++) {
			const IcbCand c = C[V.plan[h.cand_off + i]];
			if (dl_passes(V, r, c)) {
			    V.dec_cls[pos] = dl_real_id(V, C0, c.id, false);
			    V.dec_strand[pos] = c.strand;
			    out = DL_JOIN;
			    break;
			}
		    }
		    V.outcome[pos] = out;
		}
		else {
		    V.plan_n[pos] = g;
		    V.plan_i[pos] = 0;
		    V.aln_calls[pos] = 0;
		    dl_advance(V, pos, false);
		}
	    }
	}
    }
}

// one warp walks the chunk from the commit frontier
__global__ void __launch_bounds__(32) icb_k_dl_commit(IcbDLoopView V)
{
    icb_res_mark(RES_COMMIT, V.descs16);
    __shared__ IcbCand s_c[DL_CW];
    IcbDLoopState* st = V.st;
    const volatile IcbDLoopState* vst = V.st;
    if (vst->done || vst->fallback) return;
    const int N = vst->n_chunk;
    int pos = vst->commit_pos;
    const int pos0 = pos;
    if (threadIdx.x == 0) st->n_order_s = st->n_order_l = 0;  // the queues of the decision kernel that follows
    while (pos < N) {
	uint8_t o = *(volatile uint8_t*)&V.outcome[pos];
	if (o == DL_NONE || dl_dirty(V, pos, true) || (o == DL_NEED_FULL && V.full_off[pos] >= 0)) dl_decide_light(V, pos, s_c, false, true);
	else if (o == DL_NEED_PAIR) {
	    if (threadIdx.x == 0) dl_advance(V, pos, true);
	    __syncwarp();
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
	    __threadfence_block();  // lane 0 -> the other lanes of this warp (the next read's decision); later kernels see it at the kernel boundary
	}
	__syncwarp();
	pos++;
    }
    if (threadIdx.x == 0) {
	if (pos > pos0) st->progress = 1;
	st->commit_pos = pos;
	if (pos >= N) st->done = 1;
    }
    __syncwarp();
    // ---- the predicted view for the decisions behind the frontier: committed founders + the undecided reads whose decision
    // currently is "new cluster", numbered as the commit walk will number them if the predictions hold
    {
	const int lane = threadIdx.x;
	int next = *(volatile int32_t*)&st->next_cls;
	for (int w = 0; w < V.mw; w++) {
	    const int p = w * 32 + lane;
	    const bool pred = p >= pos && p < N && *(volatile uint8_t*)&V.outcome[p] == DL_NEW;
	    const uint32_t pw = __ballot_sync(0xffffffffu, pred);
	    if (p < N) V.view_cls[p] = p < pos ? V.founder_cls[p] : (pred ? next + __popc(pw & ((1u << lane) - 1u)) : -1);
	    if (lane == 0) st->viewmask[w] = st->commask[w] | pw;
	    next += __popc(pw);
	}
    }
}

constexpr int GA_ITEMS = DL_MAXW * 32 / DL_THREADS;  // chunk positions per thread of the gather kernel

__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_gather(IcbDLoopView V, int max_pairs)
{
    icb_res_mark(RES_GATHER, V.descs16);
    typedef cub::BlockScan<long long, DL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    IcbDLoopState* st = V.st;
    if (threadIdx.x == 0) st->n_pairs16 = st->n_pairs32 = st->n_full = 0;
    if (st->done || st->fallback) return;
    const int N = st->n_chunk, C0 = st->C0, cp = st->commit_pos;
    const int k = V.k;
    // what every undecided read waits for
    constexpr int NI = GA_ITEMS * DL_SPEC;
    int kind[NI];  // 0 nothing, 1 pair (16-bit kernel), 2 pair (32-bit kernel), 3 full list
    long long tiles[NI], words[NI];
    IcbPairDesc d[NI];
    for (int e = 0; e < NI; e++) {
	kind[e] = 0;
	tiles[e] = words[e] = 0;
    }
    for (int q = 0; q < GA_ITEMS; q++) {
	const int pos = cp + threadIdx.x * GA_ITEMS + q;
	if (pos >= N) continue;
	const uint8_t o = V.outcome[pos];
	if (o == DL_NEED_FULL && V.full_off[pos] < 0) {
	    kind[q * DL_SPEC] = 3;
	    tiles[q * DL_SPEC] = V.hdr[pos].n_total;
	}
	else if (o == DL_NEED_PAIR) {
	    const IcbQueryHeader h = V.hdr[pos];
	    const int r = V.chunk_reads[pos];
	    const int nn = min(V.need_n[pos], DL_SPEC);
	    for (int sp = 0; sp < nn; sp++) {
		const int e = q * DL_SPEC + sp;
		const int slot = V.need_slot[pos * DL_SPEC + sp];
		const IcbCand c = V.cands[h.cand_off + slot];
		const int rep = c.id < C0 ? V.cls_founder[c.id] : V.chunk_reads[c.id - C0];
		IcbPairDesc& p = d[e];
		p.q_off = V.offs[r];
		p.r_off = V.offs[rep];
		p.n = int32_t(V.offs[r + 1] - V.offs[r]);
		p.m = int32_t(V.offs[rep + 1] - V.offs[rep]);
		p.rc = c.strand == -1;
		const double er = V.raw_err[r] + V.raw_err[rep];		      // cluster.cpp:494-496: e1 (read) + e2 (representative)
		p.open = er <= 0.01 ? 5 : (er <= 0.04 ? 4 : (er <= 0.1 ? 3 : 2));  // setGapOpen, cluster.cpp:425-440
		const double lim = floor((1.0 - er) * double(unsigned(k)));	      // getAlnRatio, cluster.cpp:446
		p.limit = lim < -1.0 ? -1 : (lim > double(k + 1) ? k + 1 : int(lim));
		p.out = int32_t(h.cand_off + slot);
		p.pad = 0;
		const bool ok16 = max(p.n, p.m) <= 8000 && V.aln[h.cand_off + slot] != DL_ALN_REDO32;
		if (ok16) {
		    p.ncg = (p.m + 71) >> 3;
		    kind[e] = 1;
		    words[e] = p.m + 8;
		}
		else {
		    p.ncg = (p.m + 7) >> 3;
		    kind[e] = 2;
		    words[e] = 3ll * p.m + ((p.n + 7) & ~7) + 8;
		}
		tiles[e] = (long long)((p.n + 7) >> 3) * p.ncg;
	    }
	}
    }
    // chunk-order prefixes: the alignment workspace and the queues are handed out from the commit frontier on, so the read
    // the loop waits for is always served; what does not fit waits for the next round
    long long v[NI], pre[NI], tot;
    long long c16[NI], t16[NI], w16[NI], c32[NI], t32[NI], w32[NI], cf[NI], rf[NI];
#define DL_SCAN(dst, expr)                            \
    for (int q = 0; q < NI; q++) v[q] = (expr);      \
    Scan(tmp).ExclusiveSum(v, pre, tot);              \
    __syncthreads();                                  \
    for (int q = 0; q < NI; q++) dst[q] = pre[q];
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
    for (int q = 0; q < NI; q++) {
	const int pos = cp + threadIdx.x * GA_ITEMS + q / DL_SPEC;
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
    for (int q = 0; q < GA_ITEMS; q++) {
	const int pos = cp + threadIdx.x * GA_ITEMS + q;
	if (pos < N && kind[q * DL_SPEC] == 1) {
	    if (pos == cp) (V.req_count[pos] ? st->stop_pair_again : st->stop_pair_first)++;
	    V.req_count[pos]++;
	}
	if (pos == cp && pos < N && kind[q * DL_SPEC] == 3) st->stop_full++;
    }
    if (threadIdx.x == 0) {
	st->rounds++;
	// the read at the frontier waits for work that was not queued (it does not fit the workspace on its own): never
	// expected with the capacities the host guarantees; the host loop takes the chunk
	const uint8_t o = V.outcome[cp];
	if (st->n_pairs16 + st->n_pairs32 + st->n_full == 0 && (o == DL_NEED_PAIR || o == DL_NEED_FULL || o == DL_NEED_ORDER) && !st->fallback)
	    st->fallback = DL_FB_STUCK;
	st->progress = 0;
    }
}

// chunk end: the founders' insert list for the main index (AddMinimizers, cluster.cpp:178-180), their representative
// reads, and the chunk's query totals
__global__ void __launch_bounds__(DL_THREADS) icb_k_dl_finalize(IcbDLoopView V, int* n_ins)
{
    icb_res_mark(RES_FINALIZE, V.descs16);
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

template <int NMAX>
static size_t order_smem_bytes()
{
    return sizeof(IcbCand) * DL_CSM + sizeof(Key) * NMAX + 2 * sizeof(uint32_t) * NMAX + NMAX / 8 + 64;
}

static cudaError_t dl_attr_once()
{
    static cudaError_t err = [] {
	cudaError_t e = cudaFuncSetAttribute(icb_k_dl_order<DL_NMAX_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(order_smem_bytes<DL_NMAX_S>()));
	if (e == cudaSuccess)
	    e = cudaFuncSetAttribute(icb_k_dl_order<DL_NMAX_L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(order_smem_bytes<DL_NMAX_L>()));
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

// decisions of every read that needs one, then the candidate orders they queued
int icb_dl_decide(icb_ctx* ctx, const IcbDLoopView& V, int n_chunk)
{
    const int hp = ctx->hi_priority;
    ICB_LAUNCH(ctx, "dl_decide", ICB_CUDA(ctx, icb_launch_prio(icb_k_dl_decide, dim3((n_chunk + DL_WARPS - 1) / DL_WARPS), dim3(DL_THREADS), 0,
								 ctx->stream, hp, V)));
    const int grid = std::min(n_chunk, DL_ORDER_GRID);
    ICB_LAUNCH(ctx, "dl_order", ICB_CUDA(ctx, icb_launch_prio(icb_k_dl_order<DL_NMAX_S>, dim3(grid), dim3(DL_THREADS), order_smem_bytes<DL_NMAX_S>(),
								ctx->stream, hp, V, (const int32_t*)V.order_list_s, (const int32_t*)&V.st->n_order_s)));
    ICB_LAUNCH(ctx, "dl_order_large",
	       ICB_CUDA(ctx, icb_launch_prio(icb_k_dl_order<DL_NMAX_L>, dim3(std::min(grid, 16)), dim3(DL_THREADS), order_smem_bytes<DL_NMAX_L>(),
					     ctx->stream, hp, V, (const int32_t*)V.order_list_l, (const int32_t*)&V.st->n_order_l)));
    return ICB_OK;
}

int icb_dl_commit(icb_ctx* ctx, const IcbDLoopView& V)
{
    ICB_LAUNCH(ctx, "dl_commit", ICB_CUDA(ctx, icb_launch_prio(icb_k_dl_commit, dim3(1), dim3(32), 0, ctx->stream, ctx->hi_priority, V)));
    return ICB_OK;
}

int icb_dl_gather(icb_ctx* ctx, const IcbDLoopView& V, int max_pairs)
{
    ICB_LAUNCH(ctx, "dl_gather", ICB_CUDA(ctx, icb_launch_prio(icb_k_dl_gather, dim3(1), dim3(DL_THREADS), 0, ctx->stream, ctx->hi_priority, V, max_pairs)));
    return ICB_OK;
}

int icb_dl_finalize(icb_ctx* ctx, const IcbDLoopView& V, int* d_n_ins)
{
    ICB_LAUNCH(ctx, "dl_finalize", icb_k_dl_finalize<<<1, DL_THREADS, 0, ctx->stream>>>(V, d_n_ins));
    return ICB_OK;
}

cudaError_t icb_residency_set_dloop(const IcbResidencyLog& L) { return icb_residency_set_here(L); }
