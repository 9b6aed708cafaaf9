// Per-primitive face of the C ABI: one entry point per reference function of the path, for a host
// that keeps its own ClusterSortedReads loop and its own containers, and for the C++ header shim
// (include/isonclust2_compat/) the reference's unit test compiles against.  The batched hot path (icb_sketch_reads,
// icb_cluster_reads) does the same arithmetic fused and resident; these entry points take and return HOST arrays and
// run one small launch each — they exist for drop-in granularity, not for throughput.
//
//   icb_hpc_compress      HomopolymerCompress          hpc.cpp:4-32
//   icb_kmer_encode       KmerEncodeSeq                kmer_index.cpp:5-17 (the last k-mer is dropped, as the reference does)
//   icb_kmer_minimizers   GetKmerMinimizers            minimizer.cpp:78-123
//   icb_calc_error_rate   CalcErrorRate                qualscore.cpp:147-154
//   icb_index_*           MinimizerDB + AddMinimizers / UpdateMinDB / GetMinimizerHits   minimizer.cpp:31-76, 124-160
//   icb_mapped_length     getMappedRatio's totalMapped cluster.cpp:324-353 (integer form)
//   icb_count_windows     getAlnRatio's window count   cluster.cpp:442-459
#include <cub/block/block_scan.cuh>

#include <algorithm>
#include <vector>

#include "index.cuh"

namespace {

constexpr int P_THREADS = 256;

// device scratch of one call: freed (stream-ordered) when the helper goes out of scope
struct Scratch {
    icb_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(icb_ctx* c) : ctx(c) {}
    ~Scratch()
    {
	for (void* p : ptrs) icb_dfree(ctx, p);
    }
    template <class T>
    cudaError_t get(T** p, size_t n)
    {
	cudaError_t e = icb_dmalloc(ctx, (void**)p, sizeof(T) * (n ? n : 1));
	if (e == cudaSuccess) ptrs.push_back(*p);
	return e;
    }
};

// one CTA per read: run starts are ranked with a block scan per 256-base tile; the thread of a run start walks its run
// for the largest quality (signed char comparison, hpc.cpp:24)
__global__ void __launch_bounds__(P_THREADS)
icb_k_hpc(const char* __restrict__ seq, const char* __restrict__ qual, const int64_t* __restrict__ offs, char* __restrict__ out_seq,
	  char* __restrict__ out_qual, int32_t* __restrict__ out_len)
{
    typedef cub::BlockScan<int, P_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_carry;
    const int64_t a = offs[blockIdx.x], n = offs[blockIdx.x + 1] - a;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < n; t0 += P_THREADS) {
	const int64_t i = t0 + threadIdx.x;
	const int start = i < n && (i == 0 || seq[a + i] != seq[a + i - 1]);
	int rank, total;
	Scan(tmp).ExclusiveSum(start, rank, total);
	const int base = s_carry;
	if (start) {
	    const char c = seq[a + i];
	    char q = qual[a + i];
	    for (int64_t j = i + 1; j < n && seq[a + j] == c; j++)
		if (q < qual[a + j]) q = qual[a + j];
	    out_seq[a + base + rank] = c;
	    out_qual[a + base + rank] = q;
	}
	__syncthreads();
	if (threadIdx.x == 0) s_carry = base + total;
	__syncthreads();
    }
    if (threadIdx.x == 0) out_len[blockIdx.x] = s_carry;
}

// BaseToNumber (kmer_index.h): A C G T -> 0..3, anything else -> unsigned(-1); KmerToIndex: 4 * prefix + base in
// unsigned arithmetic
__global__ void icb_k_kmer_encode(const char* __restrict__ seq, int64_t n_out, int k, uint32_t* __restrict__ out)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    uint32_t v = 0;
    for (int j = 0; j < k; j++) {
	const char c = seq[i + j];
	const uint32_t b = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0xffffffffu;
	v = 4u * v + b;
    }
    out[i] = v;
}

// leftmost minimum of every window of W consecutive k-mers
__global__ void icb_k_window_argmin(const uint32_t* __restrict__ kmers, int64_t n_win, int W, int32_t* __restrict__ am)
{
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n_win) return;
    int64_t best = j;
    for (int64_t t = j + 1; t < j + W; t++)
	if (kmers[t] < kmers[best]) best = t;
    am[j] = int32_t(best);
}

// a minimizer is emitted for window 0 and whenever the argmin position changes; Index = rank (one CTA, scan with carry)
__global__ void __launch_bounds__(P_THREADS)
icb_k_emit_minimizers(const uint32_t* __restrict__ kmers, const int32_t* __restrict__ am, int64_t n_win, icb_minimizer* __restrict__ out,
		      int64_t* __restrict__ n_out)
{
    typedef cub::BlockScan<int, P_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < n_win; t0 += P_THREADS) {
	const int64_t j = t0 + threadIdx.x;
	const int emit = j < n_win && (j == 0 || am[j] != am[j - 1]);
	int rank, total;
	Scan(tmp).ExclusiveSum(emit, rank, total);
	const int base = s_carry;
	if (emit) {
	    icb_minimizer m;
	    m.Min = kmers[am[j]];
	    m.Pos = uint32_t(am[j]);
	    m.Index = uint32_t(base + rank);
	    out[base + rank] = m;
	}
	__syncthreads();
	if (threadIdx.x == 0) s_carry = base + total;
	__syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_carry;
}

// CalcErrorRate: the sum runs left to right in double (the order is the spec), one thread per string
__global__ void icb_k_error_rate(const char* __restrict__ qual, const int64_t* __restrict__ offs, int64_t n, const double* __restrict__ tab,
				 double* __restrict__ out)
{
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int64_t a = offs[r], b = offs[r + 1];
    double s = 0.0;
    for (int64_t i = a; i < b; i++) s = __dadd_rn(s, tab[int(static_cast<unsigned char>(qual[i]))]);
    out[r] = __ddiv_rn(s, double(uint64_t(b - a)));
}

// GetMinimizerHits: every (read minimizer, cluster of its posting list) pair
__global__ void icb_k_query_hits(IcbIndexView ix, const icb_minimizer* __restrict__ mins, int n_f, int n_all, icb_hit* __restrict__ out,
				 unsigned long long cap, unsigned long long* __restrict__ cursor)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_all) return;
    const icb_minimizer m = mins[t];
    const uint32_t EMPTY = 0xffffffffu;
    uint32_t s;
    if (m.Min == EMPTY) s = ix.mask + 1;
    else {
	uint32_t x = m.Min;
	x ^= x >> 16;
	x *= 0x85ebca6bu;
	x ^= x >> 13;
	x *= 0xc2b2ae35u;
	x ^= x >> 16;
	s = x & ix.mask;
	for (;;) {
	    const uint32_t k = ix.keys[s];
	    if (k == m.Min) break;
	    if (k == EMPTY) return;
	    s = (s + 1) & ix.mask;
	}
    }
    for (uint32_t n = ix.heads[s]; n != 0xffffffffu;) {
	const uint2 nd = ix.nodes[n];
	if (nd.x != 0xffffffffu) {
	    const unsigned long long at = atomicAdd(cursor, 1ull);
	    if (at < cap) {
		icb_hit h;
		h.cls = nd.x;
		h.strand = t >= n_f ? -1 : 1;
		h.pos = m.Pos;
		h.index = m.Index;
		out[at] = h;
	    }
	}
	n = nd.y;
    }
}

// totalMapped of getMappedRatio in integers: a run of g missed minimizers is bridged iff g <= nmax
__global__ void icb_k_mapped_length(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ idx, int64_t n_hits, int64_t n_mins,
				    int64_t hpc_len, int32_t nmax, int64_t* __restrict__ out)
{
    if (threadIdx.x || blockIdx.x) return;
    long long t = 0;
    if ((long long)idx[0] <= nmax) t += pos[0];
    for (int64_t i = 0; i + 1 < n_hits; i++)
	if ((long long)idx[i + 1] - (long long)idx[i] - 1 <= nmax) t += (long long)pos[i + 1] - (long long)pos[i];
    if (n_mins - (long long)idx[n_hits - 1] - 1 <= nmax) t += hpc_len - (long long)pos[n_hits - 1];
    *out = t;
}

// windows of k consecutive columns (all but the last one) of a comparison string holding at least `limit` match marks
__global__ void icb_k_count_windows(const char* __restrict__ comp, int64_t len, int k, int limit, unsigned long long* __restrict__ out)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i + k >= len) return;  // window i = [i, i + k) is tested while i + k != len (cluster.cpp:447-455)
    int nm = 0;
    for (int j = 0; j < k; j++) nm += comp[i + j] == '|';
    if (nm >= limit) atomicAdd(out, 1ull);
}

}  // namespace

struct icb_index {
    icb_ctx* ctx = nullptr;
    IcbIndex idx;
};

extern "C" {

int icb_hpc_compress(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n, char* hpc_seq, char* hpc_qual,
		     int32_t* hpc_len)
{
    if (!ctx || n < 0 || (n > 0 && (!seq || !qual || !offs || !hpc_seq || !hpc_qual || !hpc_len))) return ICB_ERR_ARG;
    if (n == 0) return ICB_OK;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nb = offs[n] - offs[0];
    if (offs[0] != 0) return ctx->fail(ICB_ERR_ARG, "offs[0] must be 0");
    Scratch S(ctx);
    char *d_s, *d_q, *d_os, *d_oq;
    int64_t* d_o;
    int32_t* d_l;
    ICB_CUDA(ctx, S.get(&d_s, nb));
    ICB_CUDA(ctx, S.get(&d_q, nb));
    ICB_CUDA(ctx, S.get(&d_os, nb));
    ICB_CUDA(ctx, S.get(&d_oq, nb));
    ICB_CUDA(ctx, S.get(&d_o, n + 1));
    ICB_CUDA(ctx, S.get(&d_l, n));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_s, seq, nb, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_q, qual, nb, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_o, offs, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "hpc_compress", icb_k_hpc<<<unsigned(n), P_THREADS, 0, st>>>(d_s, d_q, d_o, d_os, d_oq, d_l));
    ICB_CUDA(ctx, cudaMemcpyAsync(hpc_seq, d_os, nb, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(hpc_qual, d_oq, nb, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(hpc_len, d_l, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_kmer_encode(icb_ctx* ctx, const char* seq, int64_t len, int kmer_size, uint32_t* out, int64_t* n_out)
{
    if (!ctx || !n_out || len < 0 || kmer_size < 1 || kmer_size > 32 || (len > 0 && !seq)) return ICB_ERR_ARG;
    const int64_t n = len >= kmer_size ? len - kmer_size : 0;  // i = 0 .. len-k-1: the reference drops the last k-mer
    *n_out = n;
    if (n == 0) return ICB_OK;
    if (!out) return ICB_ERR_ARG;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Scratch S(ctx);
    char* d_s;
    uint32_t* d_k;
    ICB_CUDA(ctx, S.get(&d_s, len));
    ICB_CUDA(ctx, S.get(&d_k, n));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_s, seq, len, cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "kmer_encode", icb_k_kmer_encode<<<unsigned((n + 255) / 256), 256, 0, st>>>(d_s, n, kmer_size, d_k));
    ICB_CUDA(ctx, cudaMemcpyAsync(out, d_k, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_kmer_minimizers(icb_ctx* ctx, const uint32_t* kmers, int64_t n, int kmer_size, int window_size, icb_minimizer* out, int64_t* n_out)
{
    if (!ctx || !n_out || n < 0 || (n > 0 && !kmers)) return ICB_ERR_ARG;
    const int64_t W = int64_t(window_size) - kmer_size + 1;
    if (W < 1) return ctx->fail(ICB_ERR_ARG, "window_size must be at least kmer_size");
    *n_out = 0;
    const int64_t n_win = n - W + 1;
    if (n_win <= 0) return ICB_OK;
    if (!out) return ICB_ERR_ARG;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Scratch S(ctx);
    uint32_t* d_k;
    int32_t* d_am;
    icb_minimizer* d_m;
    int64_t* d_n;
    ICB_CUDA(ctx, S.get(&d_k, n));
    ICB_CUDA(ctx, S.get(&d_am, n_win));
    ICB_CUDA(ctx, S.get(&d_m, n_win));
    ICB_CUDA(ctx, S.get(&d_n, 1));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_k, kmers, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "window_argmin", icb_k_window_argmin<<<unsigned((n_win + 255) / 256), 256, 0, st>>>(d_k, n_win, int(W), d_am));
    ICB_LAUNCH(ctx, "emit_minimizers", icb_k_emit_minimizers<<<1, P_THREADS, 0, st>>>(d_k, d_am, n_win, d_m, d_n));
    int64_t cnt = 0;
    ICB_CUDA(ctx, cudaMemcpyAsync(&cnt, d_n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    if (cnt > 0) {
	ICB_CUDA(ctx, cudaMemcpyAsync(out, d_m, sizeof(icb_minimizer) * cnt, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, ctx->wait());
    }
    *n_out = cnt;
    return ICB_OK;
}

int icb_calc_error_rate(icb_ctx* ctx, const char* qual, const int64_t* offs, int64_t n, const double tab[129], double* out)
{
    if (!ctx || n < 0 || !tab || (n > 0 && (!qual || !offs || !out))) return ICB_ERR_ARG;
    if (n == 0) return ICB_OK;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nb = offs[n];
    for (int64_t i = 0; i < nb; i++)
	if (static_cast<unsigned char>(qual[i]) > 128) return ctx->fail(ICB_ERR_ARG, "quality character outside the 129-entry table");
    Scratch S(ctx);
    char* d_q;
    int64_t* d_o;
    double *d_t, *d_e;
    ICB_CUDA(ctx, S.get(&d_q, nb));
    ICB_CUDA(ctx, S.get(&d_o, n + 1));
    ICB_CUDA(ctx, S.get(&d_t, 129));
    ICB_CUDA(ctx, S.get(&d_e, n));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_q, qual, nb, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_o, offs, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_t, tab, sizeof(double) * 129, cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "calc_error_rate", icb_k_error_rate<<<unsigned((n + 127) / 128), 128, 0, st>>>(d_q, d_o, n, d_t, d_e));
    ICB_CUDA(ctx, cudaMemcpyAsync(out, d_e, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_index_create(icb_ctx* ctx, icb_index** out)
{
    if (!ctx || !out) return ICB_ERR_ARG;
    *out = nullptr;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    icb_index* ix = new icb_index;
    ix->ctx = ctx;
    int rc = ix->idx.create(ctx, 1u << 14, 1 << 14);
    if (rc) {
	delete ix;
	return rc;
    }
    ICB_CUDA(ctx, ctx->wait());
    *out = ix;
    return ICB_OK;
}

void icb_index_destroy(icb_index* ix)
{
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    cudaStreamSynchronize(ix->ctx->stream);
    ix->idx.destroy();
    delete ix;
}

// the index kernels read sketches from a read-set arena: a one- or two-sketch arena is staged for the call
static int stage_sketches(icb_ctx* ctx, Scratch& S, icb_readset& fake, const icb_minimizer* a, int64_t na, const icb_minimizer* b, int64_t nb)
{
    cudaStream_t st = ctx->stream;
    int64_t h_off[2] = {0, na};
    int32_t h_nf[2] = {int32_t(na), int32_t(nb)};
    ICB_CUDA(ctx, S.get(&fake.d_mins, size_t(na + nb)));
    ICB_CUDA(ctx, S.get(&fake.d_min_off, 2));
    ICB_CUDA(ctx, S.get(&fake.d_nf, 2));
    if (na) ICB_CUDA(ctx, cudaMemcpyAsync(fake.d_mins, a, sizeof(icb_minimizer) * na, cudaMemcpyHostToDevice, st));
    if (nb) ICB_CUDA(ctx, cudaMemcpyAsync(fake.d_mins + na, b, sizeof(icb_minimizer) * nb, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpy(fake.d_min_off, h_off, sizeof h_off, cudaMemcpyHostToDevice));
    ICB_CUDA(ctx, cudaMemcpy(fake.d_nf, h_nf, sizeof h_nf, cudaMemcpyHostToDevice));
    return ICB_OK;
}

int icb_index_add(icb_index* ix, uint32_t cls, const icb_minimizer* mins, int64_t n)
{
    if (!ix || n < 0 || (n > 0 && !mins)) return ICB_ERR_ARG;
    if (n == 0) return ICB_OK;
    icb_ctx* ctx = ix->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    icb_readset fake;
    int rc = stage_sketches(ctx, S, fake, mins, n, nullptr, 0);
    if (rc) return rc;
    int32_t h[2] = {0, int32_t(cls)};
    int32_t* d_ri;
    ICB_CUDA(ctx, S.get(&d_ri, 2));
    ICB_CUDA(ctx, cudaMemcpy(d_ri, h, sizeof h, cudaMemcpyHostToDevice));
    rc = ix->idx.reserve(n);
    if (rc) return rc;
    rc = icb_index_insert(ctx, ix->idx, &fake, d_ri, d_ri + 1, 1);
    if (rc) return rc;
    ix->idx.nodes_upper += n;
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_index_update_cluster(icb_index* ix, uint32_t cls, const icb_minimizer* old_mins, int64_t n_old, const icb_minimizer* new_mins,
			     int64_t n_new)
{
    if (!ix || n_old < 0 || n_new < 0 || (n_old > 0 && !old_mins) || (n_new > 0 && !new_mins)) return ICB_ERR_ARG;
    icb_ctx* ctx = ix->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch S(ctx);
    icb_readset fake;
    int rc = stage_sketches(ctx, S, fake, old_mins, n_old, new_mins, n_new);
    if (rc) return rc;
    rc = ix->idx.reserve(n_new);
    if (rc) return rc;
    rc = icb_index_update(ctx, ix->idx, &fake, 0, 1, int(cls), int(n_old), int(n_new));
    if (rc) return rc;
    ix->idx.nodes_upper += n_new;
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_index_query_hits(icb_index* ix, const icb_minimizer* mins, int64_t n, const icb_minimizer* rev_mins, int64_t n_rev, icb_hit* out,
			 int64_t cap, int64_t* n_out)
{
    if (!ix || !n_out || n < 0 || n_rev < 0 || cap < 0 || (n > 0 && !mins) || (n_rev > 0 && !rev_mins) || (cap > 0 && !out)) return ICB_ERR_ARG;
    *n_out = 0;
    if (n + n_rev == 0) return ICB_OK;
    icb_ctx* ctx = ix->ctx;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Scratch S(ctx);
    icb_minimizer* d_m;
    icb_hit* d_h;
    unsigned long long* d_c;
    ICB_CUDA(ctx, S.get(&d_m, size_t(n + n_rev)));
    ICB_CUDA(ctx, S.get(&d_h, size_t(cap)));
    ICB_CUDA(ctx, S.get(&d_c, 1));
    if (n) ICB_CUDA(ctx, cudaMemcpyAsync(d_m, mins, sizeof(icb_minimizer) * n, cudaMemcpyHostToDevice, st));
    if (n_rev) ICB_CUDA(ctx, cudaMemcpyAsync(d_m + n, rev_mins, sizeof(icb_minimizer) * n_rev, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemsetAsync(d_c, 0, 8, st));
    ICB_LAUNCH(ctx, "index_query_hits",
	       icb_k_query_hits<<<unsigned((n + n_rev + 127) / 128), 128, 0, st>>>(ix->idx.v, d_m, int(n), int(n + n_rev), d_h,
										    (unsigned long long)cap, d_c));
    unsigned long long cnt = 0;
    ICB_CUDA(ctx, cudaMemcpyAsync(&cnt, d_c, 8, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    *n_out = int64_t(cnt);
    if (int64_t(cnt) > cap) return ICB_OK;  // the caller asks again with room for *n_out hits
    if (cnt) {
	ICB_CUDA(ctx, cudaMemcpyAsync(out, d_h, sizeof(icb_hit) * cnt, cudaMemcpyDeviceToHost, st));
	ICB_CUDA(ctx, ctx->wait());
	// the order GetMinimizerHits appends them in (minimizer.cpp:44-76): forward minimizers by Index, each posting list by
	// ascending cluster id, then the reverse ones likewise
	std::sort(out, out + cnt, [](const icb_hit& a, const icb_hit& b) {
	    if (a.strand != b.strand) return a.strand > b.strand;
	    if (a.index != b.index) return a.index < b.index;
	    return a.cls < b.cls;
	});
    }
    return ICB_OK;
}

int icb_mapped_length(icb_ctx* ctx, const uint32_t* hit_pos, const uint32_t* hit_index, int64_t n_hits, int64_t n_minimizers, int64_t hpc_len,
		      int32_t n_max, int64_t* total_mapped)
{
    if (!ctx || !total_mapped || n_hits < 1 || !hit_pos || !hit_index) return ICB_ERR_ARG;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Scratch S(ctx);
    uint32_t *d_p, *d_i;
    int64_t* d_o;
    ICB_CUDA(ctx, S.get(&d_p, n_hits));
    ICB_CUDA(ctx, S.get(&d_i, n_hits));
    ICB_CUDA(ctx, S.get(&d_o, 1));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_p, hit_pos, 4 * n_hits, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_i, hit_index, 4 * n_hits, cudaMemcpyHostToDevice, st));
    ICB_LAUNCH(ctx, "mapped_length", icb_k_mapped_length<<<1, 32, 0, st>>>(d_p, d_i, n_hits, n_minimizers, hpc_len, n_max, d_o));
    ICB_CUDA(ctx, cudaMemcpyAsync(total_mapped, d_o, 8, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    return ICB_OK;
}

int icb_count_windows(icb_ctx* ctx, const char* comp, int64_t len, int kmer_size, int limit, int64_t* aligned)
{
    if (!ctx || !aligned || len < 0 || kmer_size < 1 || (len > 0 && !comp)) return ICB_ERR_ARG;
    *aligned = 0;
    if (len <= kmer_size) return ICB_OK;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Scratch S(ctx);
    char* d_c;
    unsigned long long* d_o;
    ICB_CUDA(ctx, S.get(&d_c, len));
    ICB_CUDA(ctx, S.get(&d_o, 1));
    ICB_CUDA(ctx, cudaMemcpyAsync(d_c, comp, len, cudaMemcpyHostToDevice, st));
    ICB_CUDA(ctx, cudaMemsetAsync(d_o, 0, 8, st));
    ICB_LAUNCH(ctx, "count_windows", icb_k_count_windows<<<unsigned((len + 255) / 256), 256, 0, st>>>(d_c, len, kmer_size, limit, d_o));
    unsigned long long v = 0;
    ICB_CUDA(ctx, cudaMemcpyAsync(&v, d_o, 8, cudaMemcpyDeviceToHost, st));
    ICB_CUDA(ctx, ctx->wait());
    *aligned = int64_t(v);
    return ICB_OK;
}

}  // extern "C"
