// Stage 2 kernels — index build (AddMinimizers, reference minimizer.cpp:31-42) and the per-chunk candidate
// query that replaces GetMinimizerHits / ConsolidateMinimizerHits (minimizer.cpp:44-76, cluster.cpp:609-615),
// the hit counts of SortMinimizerHits (cluster.cpp:622-636) and the integer part of getMappedRatio
// (cluster.cpp:324-353) for the candidates getBestClusterMapping can reach (cluster.cpp:355-406).
//
// Exactness across the greedy loop's sequential dependency: a chunk of sorted reads is scored against the
// index SNAPSHOT at chunk start (clusters < C0) plus a CHUNK index that holds every read of the chunk as a
// provisional founder (id = C0 + position); a read only sees provisional founders that precede it.  The host
// loop keeps the ones that really became founders.  Hits against a cluster are independent of all other
// clusters, so dropping provisional ones is exact.
//
// One 128-thread CTA per read.  Shared-minimizer counts are dense 16-bit counters in shared memory (one
// 32-bit word per cluster: low half forward strand, high half reverse), bumped with shared-memory atomics
// while threads walk the posting chains of the read's minimizers.  The mapped length needs each candidate's
// hits in minimizer order: a bitmap over the read's minimizer ranks per relevant candidate, filled in a second
// walk, then scanned by one thread per candidate.  Roofline class: L2/HBM latency-bound gathers (posting
// chains), reported as hits/s; bytes per read = 12(M_f+M_r) + 8(M_f+M_r) + 8H + 16C.
#include <mutex>

#include "residency.cuh"
#include "dloop.cuh"
#include "index.cuh"

namespace {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t NIL = 0xFFFFFFFFu;
constexpr uint32_t DEAD = 0xFFFFFFFFu;	// posting node whose cluster left the list (UpdateMinDB's erase)
constexpr int Q_THREADS = 128;
constexpr int INS_SET = 2048;  // shared-memory value set of icb_k_index_insert (sketches of up to 1024 minimizers)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t find_slot(const IcbIndexView& ix, uint32_t key)
{
    if (key == EMPTY) return ix.mask + 1;
    uint32_t s = hash32(key) & ix.mask;
    for (;;) {
	uint32_t k = ix.keys[s];
	if (k == key) return s;
	if (k == EMPTY) return NIL;
	s = (s + 1) & ix.mask;
    }
}

__device__ __forceinline__ uint32_t claim_slot(const IcbIndexView& ix, uint32_t key)
{
    if (key == EMPTY) return ix.mask + 1;
    uint32_t s = hash32(key) & ix.mask;
    for (;;) {
	uint32_t prev = atomicCAS(&ix.keys[s], EMPTY, key);
	if (prev == EMPTY || prev == key) return s;
	s = (s + 1) & ix.mask;
    }
}

template <class F>
__device__ __forceinline__ void for_each_posting(const IcbIndexView& ix, uint32_t key, F f)
{
    uint32_t s = find_slot(ix, key);
    if (s == NIL) return;
    uint32_t n = ix.heads[s];
    while (n != NIL) {
	uint2 nd = ix.nodes[n];
	if (nd.x != DEAD) f(nd.x);
	n = nd.y;
    }
}

// one CTA per founder: one node per DISTINCT forward minimizer value (the reference's `cls > back()` test keeps
// a cluster once per posting list).  Node slots are claimed once per 128 minimizers (one atomic per CTA tile,
// not per node: a single global counter would serialise in L2).
__global__ void __launch_bounds__(Q_THREADS)
icb_k_index_insert(IcbIndexView ix, const icb_minimizer* __restrict__ arena, const int64_t* __restrict__ min_off,
		   const int32_t* __restrict__ nf, const int32_t* __restrict__ reads, const int32_t* __restrict__ ids, int n,
		   const int32_t* __restrict__ n_dev)
{
    __shared__ int s_w[Q_THREADS / 32];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_set[INS_SET];	 // values of this founder seen so far (open addressing), for sketches that fit
    int b = blockIdx.x;
    if (n_dev) n = *n_dev;
    if (b >= n) return;
    const int r = reads[b];
    const uint32_t id = uint32_t(ids[b]);
    const int M = nf[r];
    const icb_minimizer* m = arena + min_off[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool use_set = 2 * M <= INS_SET;  // load factor <= 1/2; which duplicate keeps the node does not matter
    if (use_set) {
	for (int i = threadIdx.x; i < INS_SET; i += Q_THREADS) s_set[i] = EMPTY;
	__syncthreads();
    }
    for (int i0 = 0; i0 < M; i0 += Q_THREADS) {
	const int i = i0 + threadIdx.x;
	uint32_t v = 0;
	bool fresh = false;
	if (i < M) {
	    v = m[i].Min;
	    fresh = true;
	    if (use_set && v != EMPTY) {
		uint32_t s = hash32(v) & (INS_SET - 1);
		for (;;) {
		    const uint32_t prev = atomicCAS(&s_set[s], EMPTY, v);
		    if (prev == EMPTY) break;
		    if (prev == v) {
			fresh = false;
			break;
		    }
		    s = (s + 1) & (INS_SET - 1);
		}
	    }
	    else {
		for (int j = 0; j < i; j++)
		    if (m[j].Min == v) {
			fresh = false;
			break;
		    }
	    }
	}
	const uint32_t bal = __ballot_sync(0xffffffffu, fresh);
	if (lane == 0) s_w[warp] = __popc(bal);
	__syncthreads();
	int before = 0, total = 0;
	for (int w = 0; w < Q_THREADS / 32; w++) {
	    if (w < warp) before += s_w[w];
	    total += s_w[w];
	}
	if (threadIdx.x == 0) s_base = atomicAdd(ix.node_count, uint32_t(total));
	__syncthreads();
	if (fresh) {
	    const uint32_t nid = s_base + uint32_t(before + __popc(bal & ((1u << lane) - 1u)));
	    const uint32_t s = claim_slot(ix, v);
	    const uint32_t old = atomicExch(&ix.heads[s], nid);
	    ix.nodes[nid] = make_uint2(id, old);
	}
	__syncthreads();
    }
}

// UpdateMinDB (reference minimizer.cpp:124-160) for one cluster whose representative changed from read `r_old` to
// read `r_new` (the reference calls it after a consensus update, cluster.cpp:287-291): with oldSet / newSet the distinct
// FORWARD minimizer values of the two sketches, `cls` leaves the posting list of every value in oldSet \ newSet and joins
// the list of every value in newSet \ oldSet.  Posting order is irrelevant here: the query only counts per cluster, and
// the reference keeps its lists sorted (std::set round trip / std::sort), i.e. order carries no information there either.
// Emptied lists stay as keys, as in the reference (the erase is commented out, minimizer.cpp:150-152).
// One CTA; both value lists are staged in shared memory (dynamic: nf[r_old] + nf[r_new] words).
__global__ void __launch_bounds__(Q_THREADS)
icb_k_index_update(IcbIndexView ix, const icb_minimizer* __restrict__ arena, const int64_t* __restrict__ min_off,
		   const int32_t* __restrict__ nf, int r_old, int r_new, uint32_t cls)
{
    extern __shared__ uint32_t s_vals[];
    const int Mo = nf[r_old], Mn = nf[r_new];
    uint32_t* vo = s_vals;
    uint32_t* vn = s_vals + Mo;
    const icb_minimizer* mo = arena + min_off[r_old];
    const icb_minimizer* mn = arena + min_off[r_new];
    for (int i = threadIdx.x; i < Mo; i += Q_THREADS) vo[i] = mo[i].Min;
    for (int i = threadIdx.x; i < Mn; i += Q_THREADS) vn[i] = mn[i].Min;
    __syncthreads();
    // toDel = oldSet \ newSet: first occurrence of each old value that the new sketch does not hold
    for (int i = threadIdx.x; i < Mo; i += Q_THREADS) {
	const uint32_t v = vo[i];
	bool act = true;
	for (int j = 0; j < i && act; j++) act = vo[j] != v;
	for (int j = 0; j < Mn && act; j++) act = vn[j] != v;
	if (!act) continue;
	const uint32_t s = find_slot(ix, v);
	if (s == NIL) continue;
	for (uint32_t n = ix.heads[s]; n != NIL;) {
	    const uint2 nd = ix.nodes[n];
	    if (nd.x == cls) ix.nodes[n].x = DEAD;  // every copy: tmp.erase(best) on a set
	    n = nd.y;
	}
    }
    // toIns = newSet \ oldSet
    for (int i = threadIdx.x; i < Mn; i += Q_THREADS) {
	const uint32_t v = vn[i];
	bool act = true;
	for (int j = 0; j < i && act; j++) act = vn[j] != v;
	for (int j = 0; j < Mo && act; j++) act = vo[j] != v;
	if (!act) continue;
	const uint32_t nid = atomicAdd(ix.node_count, 1u);
	const uint32_t s = claim_slot(ix, v);
	const uint32_t old = atomicExch(&ix.heads[s], nid);
	ix.nodes[nid] = make_uint2(cls, old);
    }
}

__global__ void icb_k_index_rehash(IcbIndexView oldv, IcbIndexView newv)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > oldv.mask + 1) return;
    uint32_t k = s == oldv.mask + 1 ? EMPTY : oldv.keys[s];
    uint32_t h = oldv.heads[s];
    if (s <= oldv.mask && k == EMPTY) return;
    if (h == NIL) return;
    uint32_t t = claim_slot(newv, k);
    newv.heads[t] = h;
}

// error class of p_emp_prob.cpp:66-94: round(e, 2) (util.cpp:6-10) clamped to [0.01, 0.15] -> 0..14
__global__ void icb_k_classify(const double* __restrict__ err, const int32_t* __restrict__ status, int64_t n, uint8_t* __restrict__ cls)
{
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    if (status[i] == 1) {
	double r = round(__dmul_rn(err[i], 100.0));  // std::round: half away from zero
	c = r > 15.0 ? 15 : (r < 1.0 ? 1 : int(r));
	c -= 1;
    }
    cls[i] = uint8_t(c);
}

template <bool SMEM_CNT>
__device__ __forceinline__ void icb_query_body(const IcbQueryArgs& a);

template <bool SMEM_CNT>
__global__ void __launch_bounds__(Q_THREADS) icb_k_query(IcbQueryArgs a)
{
    icb_query_body<SMEM_CNT>(a);
}

// the same with the arguments in device memory: the launch is part of a captured graph (dloop: one graph per round), what
// changes from chunk to chunk (snapshot size, index views, scratch) is rewritten by a small copy instead of a new capture
template <bool SMEM_CNT>
__global__ void __launch_bounds__(Q_THREADS) icb_k_query_ptr(const IcbQueryArgs* __restrict__ ap)
{
    const IcbQueryArgs a = *ap;
    icb_query_body<SMEM_CNT>(a);
}

template <bool SMEM_CNT>
__device__ __forceinline__ void icb_query_body(const IcbQueryArgs& a)
{
    if (g_icb_residency.rec && blockIdx.x == 0 && threadIdx.x == 0)
	icb_res_record(a.mode == 2 ? RES_QUERY_POOL : RES_QUERY_FULL, a.res_tag, icb_res_now(), 0u);
    extern __shared__ uint32_t sm[];
    __shared__ int s_red[Q_THREADS / 32];
    __shared__ int s_red2[Q_THREADS / 32];
    __shared__ int s_red3[Q_THREADS / 32];
    __shared__ int s_n;
    __shared__ uint32_t s_off;
    __shared__ unsigned long long s_off64;
    __shared__ int s_ok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ctot = a.C0 + a.n_chunk;
    uint32_t* cnt = SMEM_CNT ? sm : a.g_scratch + int64_t(blockIdx.x) * a.g_stride;
    uint32_t* list = SMEM_CNT ? sm + Ctot : sm;		 // 2 words per record: (id<<1|half), count
    uint32_t* bm = list + 2 * a.cap;
    const int bmWords = a.smem_words - (SMEM_CNT ? Ctot : 0) - 2 * a.cap;
    uint32_t* first = a.g_scratch + int64_t(blockIdx.x) * a.g_stride + (SMEM_CNT ? 0 : Ctot);  // mode 1 only

    const int n_items = a.dl_n_list ? *a.dl_n_list : (a.list ? a.n_list : a.n_chunk);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
	const int pos = a.list ? a.list[item] : item;
	const int r = a.chunk_reads[pos];
	const int nf = a.nf[r], nr = a.nr[r];
	const icb_minimizer* mins = a.arena + a.min_off[r];
	__syncthreads();
	// first-hit table: the full-list mode, and the pool mode of the device loop (which emits the full list right away for
	// reads whose best candidates are tied, see below)
	const bool want_first = a.mode == 1 || (a.mode == 2 && a.dl_full_pool != nullptr);
	for (int i = tid; i < Ctot; i += Q_THREADS) cnt[i] = 0;
	if (want_first)
	    for (int i = tid; i < 2 * Ctot; i += Q_THREADS) first[i] = 0xFFFFFFFFu;
	if (tid == 0) s_n = 0;
	__syncthreads();
	// ---- walk 1: shared-minimizer counts per (cluster, strand)
	for (int t = tid; t < nf + nr; t += Q_THREADS) {
	    const icb_minimizer m = mins[t];
	    const int half = t >= nf;
	    const uint32_t inc = 1u << (16 * half);
	    for_each_posting(a.main_idx, m.Min, [&](uint32_t id) {
		atomicAdd(&cnt[id], inc);
		if (want_first) atomicMin(&first[2 * id + half], m.Index);
	    });
	    for_each_posting(a.chunk_idx, m.Min, [&](uint32_t id) {
		if (int(id) - a.C0 < pos) {
		    atomicAdd(&cnt[id], inc);
		    if (want_first) atomicMin(&first[2 * id + half], m.Index);
		}
	    });
	}
	__syncthreads();
	// ---- top count among snapshot clusters, number of live keys
	int top = 0, live = 0, hits = 0;
	for (int i = tid; i < Ctot; i += Q_THREADS) {
	    uint32_t v = cnt[i];
	    int lo = v & 0xffff, hi = v >> 16;
	    live += (lo > 0) + (hi > 0);
	    hits += lo + hi;
	    if (i < a.C0) top = max(top, max(lo, hi));
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) {
	    top = max(top, __shfl_xor_sync(0xffffffffu, top, d));
	    live += __shfl_xor_sync(0xffffffffu, live, d);
	    hits += __shfl_xor_sync(0xffffffffu, hits, d);
	}
	if (lane == 0) {
	    s_red[warp] = top;
	    s_red2[warp] = live;
	    s_red3[warp] = hits;
	}
	__syncthreads();
	top = 0;
	live = 0;
	hits = 0;
	for (int i = 0; i < Q_THREADS / 32; i++) {
	    top = max(top, s_red[i]);
	    live += s_red2[i];
	    hits += s_red3[i];
	}
	IcbQueryHeader* hdr = a.hdr + item;
	// every (cluster, strand) with a hit as a record {id, strand, count, first hit}, in ascending id order (the host's counting
	// sort relies on it): thread t owns a contiguous id range, its output offset is the exclusive prefix of the live keys
	// before that range.  Leaves the number of records in s_n.
	auto emit_full = [&](IcbCand* out) {
	    const int per = (Ctot + Q_THREADS - 1) / Q_THREADS;
	    const int lo = min(Ctot, tid * per), hi = min(Ctot, lo + per);
	    int mine = 0;
	    for (int i = lo; i < hi; i++) {
		const uint32_t v = cnt[i];
		mine += ((v & 0xffffu) != 0) + ((v >> 16) != 0);
	    }
	    int inc = mine;
#pragma unroll
	    for (int d = 1; d < 32; d <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += t;
	    }
	    __syncthreads();  // s_red was read above
	    if (lane == 31) s_red[warp] = inc;
	    __syncthreads();
	    int slot = inc - mine;
	    for (int w = 0; w < warp; w++) slot += s_red[w];
	    for (int i = lo; i < hi; i++) {
		const uint32_t v = cnt[i];
		for (int half = 0; half < 2; half++) {
		    const int c = half ? int(v >> 16) : int(v & 0xffff);
		    if (c > 0) {
			IcbCand cd;
			cd.id = i;
			cd.strand = half ? -1 : 1;
			cd.count = c;
			cd.tmapped = int(first[2 * i + half]);
			out[slot++] = cd;
			// the candidate ORDER of a read depends on every real candidate of its full list
			if (a.dl_depmask && i >= a.C0)
			    atomicOr(&a.dl_depmask[int64_t(pos) * a.dl_mw + ((i - a.C0) >> 5)], 1u << ((i - a.C0) & 31));
		    }
		}
	    }
	    __syncthreads();
	    if (tid == Q_THREADS - 1) s_n = slot;
	    __syncthreads();
	};
	if (a.mode == 1) {
	    emit_full(a.cands + a.full_off[item]);
	    if (tid == 0) {
		hdr->top_snap = top;
		hdr->n_cand = s_n;
		hdr->overflow = 0;
		hdr->n_total = live;
		hdr->n_rel_all = 0;
		hdr->hits = hits;
		hdr->cand_off = 0;
		if (a.dl_full_off) {
		    a.dl_full_n[pos] = s_n;
		    a.dl_full_off[pos] = a.full_off[item];
		}
	    }
	    continue;
	}
	if (a.mode == 2 && a.dl_full_pool) {
	    // ---- reads whose best candidates are tied will need the reference's candidate ORDER, hence their full candidate
	    // list: it is emitted now, while the counts are at hand (the device loop would otherwise come back for it a round
	    // trip later).  Best = over snapshot and provisional candidates alike; a tie the committed founders only reveal later
	    // still takes the separate full-list query.
	    int ta = 0;
	    for (int i = tid; i < Ctot; i += Q_THREADS) {
		const uint32_t v = cnt[i];
		ta = max(ta, max(int(v & 0xffff), int(v >> 16)));
	    }
#pragma unroll
	    for (int d = 16; d; d >>= 1) ta = max(ta, __shfl_xor_sync(0xffffffffu, ta, d));
	    __syncthreads();
	    if (lane == 0) s_red[warp] = ta;
	    __syncthreads();
	    ta = 0;
	    for (int i = 0; i < Q_THREADS / 32; i++) ta = max(ta, s_red[i]);
	    int nt = 0;
	    for (int i = tid; i < Ctot; i += Q_THREADS) {
		const uint32_t v = cnt[i];
		nt += (int(v & 0xffff) == ta) + (int(v >> 16) == ta);
	    }
#pragma unroll
	    for (int d = 16; d; d >>= 1) nt += __shfl_xor_sync(0xffffffffu, nt, d);
	    __syncthreads();
	    if (lane == 0) s_red[warp] = nt;
	    __syncthreads();
	    if (tid == 0) {
		int n2 = 0;
		for (int i = 0; i < Q_THREADS / 32; i++) n2 += s_red[i];
		s_ok = 0;
		if (ta > 0 && n2 >= 2 && !(unsigned(ta) < unsigned(a.min_shared))) {
		    const unsigned long long off = atomicAdd(&a.dl_state->full_cursor, (unsigned long long)live);
		    if (off + (unsigned long long)live <= a.dl_state->full_cap) {
			s_ok = 1;
			s_off64 = off;
		    }
		}
	    }
	    __syncthreads();
	    if (s_ok) {
		emit_full(a.dl_full_pool + s_off64);
		if (tid == 0) {
		    a.dl_full_n[pos] = s_n;
		    a.dl_full_off[pos] = int64_t(s_off64);
		    s_n = 0;
		}
	    }
	    __syncthreads();
	}
	if (a.mode == 2) {
	    // ---- pool mode (device-resident loop): the relevant candidates of mode 0, written as records into the chunk's
	    // candidate pool at a range reserved with one atomic per read
	    const int tbase2 = max(top, a.min_shared);
	    int thr2 = __double2int_rz(__dmul_rn(double(tbase2), a.min_frac));
	    if (thr2 < 1) thr2 = 1;
	    int mine = 0;
	    for (int i = tid; i < Ctot; i += Q_THREADS) {
		const uint32_t v = cnt[i];
		mine += (int(v & 0xffff) >= thr2) + (int(v >> 16) >= thr2);
	    }
#pragma unroll
	    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
	    __syncthreads();  // s_red was read above
	    if (lane == 0) s_red[warp] = mine;
	    __syncthreads();
	    if (tid == 0) {
		int nAll = 0;
		for (int i = 0; i < Q_THREADS / 32; i++) nAll += s_red[i];
		const uint32_t off = atomicAdd(&a.dl_state->cand_cursor, uint32_t(nAll));
		const bool ok = nAll <= 32767 && uint64_t(off) + uint64_t(nAll) <= uint64_t(a.dl_state->cand_cap);
		if (!ok) atomicMax(&a.dl_state->fallback, nAll > 32767 ? DL_FB_TOO_MANY : DL_FB_POOL);
		s_off = off;
		s_ok = ok;
		s_red2[0] = nAll;
	    }
	    __syncthreads();
	    const uint32_t off = s_off;
	    const int nAll = s_red2[0];
	    if (!s_ok) {
		if (tid == 0) {
		    hdr->top_snap = top;
		    hdr->n_cand = 0;
		    hdr->overflow = 1;
		    hdr->n_total = live;
		    hdr->n_rel_all = nAll;
		    hdr->hits = hits;
		    hdr->cand_off = 0;
		}
		continue;
	    }
	    for (int i = tid; i < Ctot; i += Q_THREADS) {
		const uint32_t v = cnt[i];
		uint32_t nv = v;
		for (int half = 0; half < 2; half++) {
		    const int c = half ? int(v >> 16) : int(v & 0xffff);
		    if (c >= thr2) {
			const int slot = atomicAdd(&s_n, 1);
			IcbCand cd;
			cd.id = i;
			cd.strand = half ? -1 : 1;
			cd.count = c;
			cd.tmapped = 0;
			a.dl_pool[off + slot] = cd;
			a.dl_pool_aln[off + slot] = DL_ALN_UNKNOWN;
			const uint32_t tag = 0x8000u | uint32_t(slot);
			nv = half ? ((nv & 0x0000ffffu) | (tag << 16)) : ((nv & 0xffff0000u) | tag);
			if (i >= a.C0) atomicOr(&a.dl_depmask[int64_t(pos) * a.dl_mw + ((i - a.C0) >> 5)], 1u << ((i - a.C0) & 31));
		    }
		}
		if (nv != v) cnt[i] = nv;
	    }
	    __syncthreads();
	    const int nRel2 = s_n;
	    const int bw2 = (max(nf, nr) + 31) >> 5;
	    const int perRound2 = max(1, bmWords / max(bw2, 1));
	    const int myClass2 = a.read_class[r];
	    const int Lh2 = a.hpc_len[r];
	    for (int base = 0; base < nRel2; base += perRound2) {
		const int cntRound = min(perRound2, nRel2 - base);
		for (int i = tid; i < cntRound * bw2; i += Q_THREADS) bm[i] = 0;
		__syncthreads();
		for (int t = tid; t < nf + nr; t += Q_THREADS) {
		    const icb_minimizer m = mins[t];
		    const int half = t >= nf;
		    auto mark = [&](uint32_t id) {
			const uint32_t v = cnt[id];
			const uint32_t h = half ? (v >> 16) : (v & 0xffffu);
			if (h & 0x8000u) {
			    const int slot = int(h & 0x7fffu) - base;
			    if (slot >= 0 && slot < cntRound) atomicOr(&bm[slot * bw2 + (m.Index >> 5)], 1u << (m.Index & 31));
			}
		    };
		    for_each_posting(a.main_idx, m.Min, mark);
		    for_each_posting(a.chunk_idx, m.Min, [&](uint32_t id) {
			if (int(id) - a.C0 < pos) mark(id);
		    });
		}
		__syncthreads();
		for (int q = tid; q < cntRound; q += Q_THREADS) {
		    IcbCand* cd = a.dl_pool + off + base + q;
		    const int id = cd->id, half = cd->strand < 0;
		    const int M = half ? nr : nf;
		    const icb_minimizer* mm = mins + (half ? nf : 0);
		    const int repRead = id < a.C0 ? a.cls_founder[id] : a.chunk_reads[id - a.C0];
		    const int nm = a.nmax[int(a.read_class[repRead]) * 15 + myClass2];
		    const uint32_t* b = bm + q * bw2;
		    long long T = 0;
		    int prev = -1;
		    for (int wd = 0; wd < bw2; wd++) {
			uint32_t x = b[wd];
			while (x) {
			    const int bit = __ffs(x) - 1;
			    x &= x - 1;
			    const int idx = wd * 32 + bit;
			    if (prev < 0) {
				if (idx <= nm) T += mm[idx].Pos;
			    }
			    else if (idx - prev - 1 <= nm) T += (long long)mm[idx].Pos - mm[prev].Pos;
			    prev = idx;
			}
		    }
		    if (prev >= 0 && M - prev - 1 <= nm) T += (long long)Lh2 - mm[prev].Pos;
		    cd->tmapped = int(T);
		}
		__syncthreads();
	    }
	    if (tid == 0) {
		hdr->top_snap = top;
		hdr->n_cand = nRel2;
		hdr->overflow = 0;
		hdr->n_total = live;
		hdr->n_rel_all = nAll;
		hdr->hits = hits;
		hdr->cand_off = off;
	    }
	    continue;
    }
	// ---- mode 0: candidates the mapping / alignment tests can reach.  The real top count may still be raised
	// by a provisional founder, never lowered, so int(max(top, minShared) * minFraction) is a safe lower bound
	// of the cut getBestClusterMapping applies (cluster.cpp:387).
	const int tbase = max(top, a.min_shared);
	int thr = __double2int_rz(__dmul_rn(double(tbase), a.min_frac));
	if (thr < 1) thr = 1;
	for (int i = tid; i < Ctot; i += Q_THREADS) {
	    uint32_t v = cnt[i];
	    uint32_t nv = v;
	    for (int half = 0; half < 2; half++) {
		int c = half ? int(v >> 16) : int(v & 0xffff);
		if (c >= thr) {
		    int slot = atomicAdd(&s_n, 1);
		    if (slot < a.cap) {
			list[2 * slot] = uint32_t(i) << 1 | uint32_t(half);
			list[2 * slot + 1] = uint32_t(c);
			uint32_t tag = 0x8000u | uint32_t(slot);
			nv = half ? ((nv & 0x0000ffffu) | (tag << 16)) : ((nv & 0xffff0000u) | tag);
		    }
		}
	    }
	    if (nv != v) cnt[i] = nv;
	}
	__syncthreads();
	const int nRelAll = s_n;
	const int nRel = min(nRelAll, a.cap);
	const int bw = (max(nf, nr) + 31) >> 5;
	const int perRound = max(1, bmWords / max(bw, 1));
	const int myClass = a.read_class[r];
	const int Lh = a.hpc_len[r];
	for (int base = 0; base < nRel; base += perRound) {
	    const int cntRound = min(perRound, nRel - base);
	    for (int i = tid; i < cntRound * bw; i += Q_THREADS) bm[i] = 0;
	    __syncthreads();
	    // ---- walk 2: which of the read's minimizers hit each relevant candidate
	    for (int t = tid; t < nf + nr; t += Q_THREADS) {
		const icb_minimizer m = mins[t];
		const int half = t >= nf;
		auto mark = [&](uint32_t id) {
		    uint32_t v = cnt[id];
		    uint32_t h = half ? (v >> 16) : (v & 0xffffu);
		    if (h & 0x8000u) {
			int slot = int(h & 0x7fffu) - base;
			if (slot >= 0 && slot < cntRound) atomicOr(&bm[slot * bw + (m.Index >> 5)], 1u << (m.Index & 31));
		    }
		};
		for_each_posting(a.main_idx, m.Min, mark);
		for_each_posting(a.chunk_idx, m.Min, [&](uint32_t id) {
		    if (int(id) - a.C0 < pos) mark(id);
		});
	    }
	    __syncthreads();
	    // ---- totalMapped (cluster.cpp:333-351) in integers: a gap of g missed minimizers is bridged iff
	    // pow(pError, g) >= minProbNoHits  <=>  g <= nmax[class(rep), class(read)]
	    if (tid < cntRound) {
		const int slot = base + tid;
		const uint32_t e = list[2 * slot];
		const int id = int(e >> 1), half = int(e & 1);
		const int M = half ? nr : nf;
		const icb_minimizer* mm = mins + (half ? nf : 0);
		const int repRead = id < a.C0 ? a.cls_founder[id] : a.chunk_reads[id - a.C0];
		const int nm = a.nmax[int(a.read_class[repRead]) * 15 + myClass];
		const uint32_t* b = bm + tid * bw;
		long long T = 0;
		int prev = -1;
		for (int wd = 0; wd < bw; wd++) {
		    uint32_t x = b[wd];
		    while (x) {
			int bit = __ffs(x) - 1;
			x &= x - 1;
			int idx = wd * 32 + bit;
			if (prev < 0) {
			    if (idx <= nm) T += mm[idx].Pos;
			}
			else if (idx - prev - 1 <= nm) T += (long long)mm[idx].Pos - mm[prev].Pos;
			prev = idx;
		    }
		}
		if (prev >= 0 && M - prev - 1 <= nm) T += (long long)Lh - mm[prev].Pos;
		IcbCand cd;
		cd.id = id;
		cd.strand = half ? -1 : 1;
		cd.count = int(list[2 * slot + 1]);
		cd.tmapped = int(T);
		a.cands[int64_t(item) * a.cap + slot] = cd;
	    }
	    __syncthreads();
	}
	if (tid == 0) {
	    hdr->top_snap = top;
	    hdr->n_cand = nRel;
	    hdr->overflow = nRelAll > a.cap;
	    hdr->n_total = live;
	    hdr->n_rel_all = nRelAll;
	    hdr->hits = hits;
	    hdr->cand_off = 0;
	}
    }
}

}  // namespace

// ---------------------------------------------------------------- host side of the index

int IcbIndex::create(icb_ctx* c, uint32_t cap_pow2, int64_t node_cap)
{
    ctx = c;
    capacity = cap_pow2;
    node_capacity = node_cap;
    nodes_upper = 0;
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&v.keys, sizeof(uint32_t) * (size_t(capacity) + 1)));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&v.heads, sizeof(uint32_t) * (size_t(capacity) + 1)));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&v.nodes, sizeof(uint2) * size_t(node_capacity)));
    ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&v.node_count, sizeof(uint32_t) * 2));
    v.mask = capacity - 1;
    return clear();
}

int IcbIndex::clear()
{
    nodes_upper = 0;
    ICB_CUDA(ctx, cudaMemsetAsync(v.keys, 0xff, sizeof(uint32_t) * (size_t(capacity) + 1), ctx->stream));
    ICB_CUDA(ctx, cudaMemsetAsync(v.heads, 0xff, sizeof(uint32_t) * (size_t(capacity) + 1), ctx->stream));
    ICB_CUDA(ctx, cudaMemsetAsync(v.node_count, 0, sizeof(uint32_t) * 2, ctx->stream));
    return ICB_OK;
}

int IcbIndex::reserve(int64_t extra)
{
    const int64_t need = nodes_upper + extra;
    if (need > node_capacity) {
	int64_t nc = node_capacity;
	while (nc < need) nc *= 2;
	uint2* nn = nullptr;
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&nn, sizeof(uint2) * size_t(nc)));
	ICB_CUDA(ctx, cudaMemcpyAsync(nn, v.nodes, sizeof(uint2) * size_t(nodes_upper), cudaMemcpyDeviceToDevice, ctx->stream));
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, v.nodes);
	v.nodes = nn;
	node_capacity = nc;
    }
    if (need * 2 > int64_t(capacity)) {	 // distinct keys <= nodes: keep the load factor under one half
	uint64_t ncap = capacity;
	while (int64_t(ncap) < need * 2) ncap *= 2;
	if (ncap > (1ull << 31)) return ctx->fail(ICB_ERR_NOMEM, "minimizer index too large");
	IcbIndexView nv = v;
	nv.mask = uint32_t(ncap - 1);
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&nv.keys, sizeof(uint32_t) * (size_t(ncap) + 1)));
	ICB_CUDA(ctx, icb_dmalloc(ctx, (void**)&nv.heads, sizeof(uint32_t) * (size_t(ncap) + 1)));
	ICB_CUDA(ctx, cudaMemsetAsync(nv.keys, 0xff, sizeof(uint32_t) * (size_t(ncap) + 1), ctx->stream));
	ICB_CUDA(ctx, cudaMemsetAsync(nv.heads, 0xff, sizeof(uint32_t) * (size_t(ncap) + 1), ctx->stream));
	ICB_LAUNCH(ctx, "index_rehash", icb_k_index_rehash<<<unsigned((size_t(capacity) + 1 + 255) / 256), 256, 0, ctx->stream>>>(v, nv));
	ICB_CUDA(ctx, ctx->wait());
	icb_dfree(ctx, v.keys);
	icb_dfree(ctx, v.heads);
	v = nv;
	capacity = uint32_t(ncap);
    }
    return ICB_OK;
}

void IcbIndex::destroy()
{
    icb_dfree(ctx, v.keys);
    icb_dfree(ctx, v.heads);
    icb_dfree(ctx, v.nodes);
    icb_dfree(ctx, v.node_count);
    v = IcbIndexView{};
}

int icb_index_insert(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, const int32_t* d_reads, const int32_t* d_ids, int n)
{
    if (n <= 0) return ICB_OK;
    ICB_LAUNCH(ctx, "index_insert",
	       icb_k_index_insert<<<n, Q_THREADS, 0, ctx->stream>>>(idx.v, rs->d_mins, rs->d_min_off, rs->d_nf, d_reads, d_ids, n, nullptr));
    return ICB_OK;
}

int icb_index_insert_dev(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, const int32_t* d_reads, const int32_t* d_ids, const int32_t* d_n,
			 int n_max)
{
    if (n_max <= 0) return ICB_OK;
    ICB_LAUNCH(ctx, "index_insert",
	       icb_k_index_insert<<<n_max, Q_THREADS, 0, ctx->stream>>>(idx.v, rs->d_mins, rs->d_min_off, rs->d_nf, d_reads, d_ids, 0, d_n));
    return ICB_OK;
}

int icb_index_update(icb_ctx* ctx, IcbIndex& idx, const icb_readset* rs, int r_old, int r_new, int cls, int nf_old, int nf_new)
{
    const size_t smem = sizeof(uint32_t) * size_t(nf_old + nf_new);
    if (smem > 48 * 1024) {
	static std::once_flag once;
	static cudaError_t once_err = cudaSuccess;
	std::call_once(once, [] { once_err = cudaFuncSetAttribute(icb_k_index_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
	ICB_CUDA(ctx, once_err);
	if (smem > 200 * 1024) return ctx->fail(ICB_ERR_NOMEM, "index update: sketches too large for one CTA");
    }
    ICB_LAUNCH(ctx, "index_update",
	       icb_k_index_update<<<1, Q_THREADS, smem, ctx->stream>>>(idx.v, rs->d_mins, rs->d_min_off, rs->d_nf, r_old, r_new, uint32_t(cls)));
    return ICB_OK;
}

int icb_query_launch(icb_ctx* ctx, const IcbQueryArgs& a, int n_blocks, size_t smem_bytes)
{
    if (n_blocks <= 0) return ICB_OK;
    // the opt-in shared-memory ceiling is per function and process-wide: raise it once to the largest plan the host
    // side ever makes (cluster.cu caps the plan at 200 KiB) so concurrent worker threads never lower it under each other
    static std::once_flag once;
    static cudaError_t once_err = cudaSuccess;
    std::call_once(once, [] {
	once_err = cudaFuncSetAttribute(icb_k_query<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	if (once_err == cudaSuccess)
	    once_err = cudaFuncSetAttribute(icb_k_query<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    ICB_CUDA(ctx, once_err);
    if (a.smem_cnt) {
	ICB_LAUNCH(ctx, a.mode == 0 ? "index_query" : (a.mode == 1 ? "index_query_full" : "index_query_pool"),
		   icb_k_query<true><<<n_blocks, Q_THREADS, smem_bytes, ctx->stream>>>(a));
    }
    else {
	ICB_LAUNCH(ctx, a.mode == 0 ? "index_query" : (a.mode == 1 ? "index_query_full" : "index_query_pool"),
		   icb_k_query<false><<<n_blocks, Q_THREADS, smem_bytes, ctx->stream>>>(a));
    }
    return ICB_OK;
}

int icb_query_launch_ptr(icb_ctx* ctx, const IcbQueryArgs* d_args, bool smem_cnt, int n_blocks, size_t smem_bytes)
{
    if (n_blocks <= 0) return ICB_OK;
    static std::once_flag once;
    static cudaError_t once_err = cudaSuccess;
    std::call_once(once, [] {
	once_err = cudaFuncSetAttribute(icb_k_query_ptr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	if (once_err == cudaSuccess)
	    once_err = cudaFuncSetAttribute(icb_k_query_ptr<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    ICB_CUDA(ctx, once_err);
    if (smem_cnt) {
	ICB_LAUNCH(ctx, "index_query_full",
		   ICB_CUDA(ctx, icb_launch_prio(icb_k_query_ptr<true>, dim3(n_blocks), dim3(Q_THREADS), smem_bytes, ctx->stream, ctx->hi_priority, d_args)));
    }
    else {
	ICB_LAUNCH(ctx, "index_query_full",
		   ICB_CUDA(ctx, icb_launch_prio(icb_k_query_ptr<false>, dim3(n_blocks), dim3(Q_THREADS), smem_bytes, ctx->stream, ctx->hi_priority, d_args)));
    }
    return ICB_OK;
}

int icb_classify_errors(icb_ctx* ctx, const double* d_err, const int32_t* d_status, int64_t n, uint8_t* d_class)
{
    if (n <= 0) return ICB_OK;
    ICB_LAUNCH(ctx, "classify_errors", icb_k_classify<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(d_err, d_status, n, d_class));
    return ICB_OK;
}

cudaError_t icb_residency_set_index(const IcbResidencyLog& L) { return icb_residency_set_here(L); }
