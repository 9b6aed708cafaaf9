// Internal definitions shared by the CUDA translation units of libisonclust2_b200.so.
#pragma once
#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "isonclust2_b200.h"

#define ICB_SM_COUNT_FALLBACK 148

struct icb_kernel_timer {
    int64_t launches = 0;
    double total_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct IcbAlignWorkspace;

struct icb_ctx {
    int device = 0;
    IcbAlignWorkspace* align_ws = nullptr;  // stage 3 buffers (align.cuh), created on first use
    int sm_count = ICB_SM_COUNT_FALLBACK;
    int hi_priority = 0;  // greatest launch priority of the device (cudaDeviceGetStreamPriorityRange)
    int lo_priority = 0;  // least
    int aln_priority = 0; // launch priority of this context's alignment kernels in the device loop (a worker's batch rank sets it)
    bool aln_latency = false;  // device loop: this context's pairs go to the CTA-per-pair kernel (shortest time per pair) instead of
			       // the warp-per-pair one (most pairs per second)
    cudaStream_t stream = nullptr;
    std::string last_error;
    bool timing = false;
    int64_t launch_count = 0;
    std::map<std::string, icb_kernel_timer> timers;
    std::vector<cudaEvent_t> event_pool;
    // icb_cluster_batches keeps its workers (child context + clusterer with all its buffers) for the next call
    std::vector<icb_ctx*> pool_ctx;
    std::vector<icb_clusterer*> pool_cl;
    // worker contexts sleep in wait() (blocking-sync event) instead of spinning: more batches are kept in flight than
    // the host has cores, and a spinning waiter would take a core from a worker that has decisions to make
    cudaEvent_t trace_base = nullptr;  // ICB_TRACE=<file>: launch intervals relative to this event are appended to the file
    bool blocking_wait = false;
    cudaEvent_t wait_event = nullptr;

    // host waits until everything queued on `stream` has finished
    cudaError_t wait();
    int fail_cuda(cudaError_t e, const char* what, const char* file, int line);
    int fail(int code, const std::string& msg);
    cudaEvent_t get_event();
    void timer_begin(const char* name);
    void timer_end(const char* name);
    void timers_collect();
};

// Device buffers of the clustering loop come from the stream-ordered allocator (the device's default memory pool with the
// release threshold raised, icb_ctx_create): growing or dropping one worker's buffer is ordered on that worker's stream
// only, where cudaFree would wait for the whole device and stall every other worker's launches meanwhile.
inline cudaError_t icb_dmalloc(icb_ctx* ctx, void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream); }
inline cudaError_t icb_dfree(icb_ctx* ctx, void* p) { return p ? cudaFreeAsync(p, ctx->stream) : cudaSuccess; }

#define ICB_CUDA(ctx, call)                                                   \
    do {                                                                      \
	cudaError_t e__ = (call);                                             \
	if (e__ != cudaSuccess) return (ctx)->fail_cuda(e__, #call, __FILE__, __LINE__); \
    } while (0)

// NVTX range over a host-side phase of the path (header-only NVTX 3: a no-op unless a tool is attached); the names are the
// reference's function names, so a timeline reads like its call stack.
struct IcbRange {
    explicit IcbRange(const char* name) { nvtxRangePushA(name); }
    ~IcbRange() { nvtxRangePop(); }
    IcbRange(const IcbRange&) = delete;
    IcbRange& operator=(const IcbRange&) = delete;
};

// RAII-free launch bracket: ICB_LAUNCH(ctx, "name", kernel<<<...>>>(...));
#define ICB_LAUNCH(ctx, name, ...)                                            \
    do {                                                                      \
	(ctx)->launch_count++;                                                \
	if ((ctx)->timing) (ctx)->timer_begin(name);                          \
	__VA_ARGS__;                                                          \
	if ((ctx)->timing) (ctx)->timer_end(name);                            \
	ICB_CUDA(ctx, cudaGetLastError());                                    \
    } while (0)

// A launch with an execution priority (also recorded in a captured graph node).  The device loop's small kernels (commit /
// decide / order / gather / full-list query) run at the device's greatest priority: the alignment kernels of the other
// batches in flight fill every SM's register file, and a pending small CTA must not queue behind their backlog.
template <class... KArgs, class... Args>
inline cudaError_t icb_launch_prio(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int prio, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = prio;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct icb_readset {
    icb_ctx* ctx = nullptr;
    int64_t n_reads = 0;
    int64_t n_bases = 0;
    int64_t max_len = 0;
    // room behind the uploaded reads for reads appended later (icb_readset_append_from: consensus representatives); every
    // per-read / per-base array is allocated for cap_reads / cap_bases
    int64_t cap_reads = 0, cap_bases = 0;
    int64_t arena_used = 0;	  // minimizer arena entries in use (after icb_sketch_reads)
    std::vector<int64_t> h_offs;  // host copy of the offsets
    // inputs
    uint8_t* d_seq = nullptr;
    uint8_t* d_qual = nullptr;
    int64_t* d_offs = nullptr;
    // stage 1a
    bool scored = false;
    double* d_score = nullptr;
    double* d_err = nullptr;
    // stage 1b
    bool sketched = false;
    int sk_k = 0, sk_w = 0;
    int32_t* d_status = nullptr;
    int32_t* d_hpc_len = nullptr;
    double* d_hpc_err = nullptr;
    uint8_t* d_hpc_seq = nullptr;
    uint8_t* d_hpc_qual = nullptr;
    icb_minimizer* d_mins = nullptr;  // arena: per read fwd then rev, located by d_min_off
    int64_t mins_capacity = 0;
    int64_t* d_min_off = nullptr;  // arena offset of read i (fwd minimizers first)
    int32_t* d_nf = nullptr;	   // forward count
    int32_t* d_nr = nullptr;	   // reverse count
    unsigned long long* d_counters = nullptr;  // [0] arena cursor, [1] error flags, [2] long-read list cursor
    int32_t* d_long_list = nullptr;
    uint8_t* d_long_scratch = nullptr;
    int64_t total_f = 0, total_r = 0;
    // host mirror of the per-read scalars the greedy loop reads, and the device error classes (filled once,
    // by the first clusterer created on this read set)
    bool host_ready = false;
    std::vector<int32_t> h_status, h_hpc_len, h_nf, h_nr;
    std::vector<double> h_score, h_err;
    int max_m = 0;
    uint8_t* d_class = nullptr;
    std::vector<uint32_t> h_bkt;  // per read: bucket count of the reference's hit map (libstdc++ _M_next_bkt(20 * (nf + nr)))
    uint32_t* d_bkt = nullptr;
};

// a worker's context: same device, own stream and timers; destroy folds timers / launch count into the parent
icb_ctx* icb_ctx_create_child(icb_ctx* parent);
void icb_ctx_fold_child(icb_ctx* parent, icb_ctx* child);  // timers / launch count -> parent, child's reset
void icb_ctx_destroy_child(icb_ctx* parent, icb_ctx* child);
void icb_worker_pool_release(icb_ctx* ctx);  // cluster.cu
int icb_readset_prepare_host(icb_readset* rs);

// stage launchers (defined in score.cu / sketch.cu)
int icb_launch_score(icb_readset* rs, int k, const double* tab, const double* tab_nomin);
int icb_launch_error_rate(icb_readset* rs, const uint8_t* d_qual, const int64_t* d_offs, const int32_t* d_len,
			  const int32_t* d_status, const double* tab_nomin, double* d_out);
int icb_launch_sketch(icb_readset* rs, int k, int w, const uint8_t* d_skip);
