// Context, error reporting, kernel timers and the device-resident read set.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "align.cuh"

int icb_ctx::fail_cuda(cudaError_t e, const char* what, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what, file, line);
    last_error = buf;
    return ICB_ERR_CUDA;
}

int icb_ctx::fail(int code, const std::string& msg)
{
    last_error = msg;
    return code;
}

cudaError_t icb_ctx::wait()
{
    if (!blocking_wait) return cudaStreamSynchronize(stream);
    cudaError_t e;
    if (!wait_event && (e = cudaEventCreateWithFlags(&wait_event, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(wait_event, stream)) != cudaSuccess) return e;
    return cudaEventSynchronize(wait_event);
}

cudaEvent_t icb_ctx::get_event()
{
    if (!event_pool.empty()) {
	cudaEvent_t e = event_pool.back();
	event_pool.pop_back();
	return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void icb_ctx::timer_begin(const char* name)
{
    auto& t = timers[name];
    cudaEvent_t a = get_event(), b = get_event();
    cudaEventRecord(a, stream);
    t.pending.emplace_back(a, b);
}

void icb_ctx::timer_end(const char* name)
{
    auto& t = timers[name];
    cudaEventRecord(t.pending.back().second, stream);
}

// development aid: ICB_TRACE=<file> appends one line "kernel stream start_ms end_ms" per timed launch (times relative to the
// moment timing was switched on), from which the overlap of the worker streams can be reconstructed
static std::mutex g_trace_mutex;

void icb_ctx::timers_collect()
{
    cudaStreamSynchronize(stream);
    FILE* tf = nullptr;
    const char* tpath = trace_base ? getenv("ICB_TRACE") : nullptr;
    std::unique_lock<std::mutex> lock(g_trace_mutex, std::defer_lock);
    if (tpath && *tpath) {
	lock.lock();
	tf = fopen(tpath, "a");
    }
    for (auto& kv : timers) {
	for (auto& p : kv.second.pending) {
	    float ms = 0.f;
	    cudaEventElapsedTime(&ms, p.first, p.second);
	    if (tf) {
		float a = 0.f, b = 0.f;
		cudaEventElapsedTime(&a, trace_base, p.first);
		cudaEventElapsedTime(&b, trace_base, p.second);
		fprintf(tf, "%s %p %.4f %.4f\n", kv.first.c_str(), (void*)stream, a, b);
	    }
	    kv.second.launches++;
	    kv.second.total_ms += ms;
	    event_pool.push_back(p.first);
	    event_pool.push_back(p.second);
	}
	kv.second.pending.clear();
    }
    if (tf) fclose(tf);
}

icb_ctx* icb_ctx_create_child(icb_ctx* parent)
{
    auto ctx = new icb_ctx;
    ctx->device = parent->device;
    ctx->sm_count = parent->sm_count;
    ctx->hi_priority = parent->hi_priority;
    ctx->lo_priority = parent->lo_priority;
    ctx->aln_priority = parent->lo_priority;
    ctx->timing = parent->timing;
    ctx->trace_base = parent->trace_base;
    ctx->blocking_wait = true;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
	delete ctx;
	return nullptr;
    }
    return ctx;
}

void icb_ctx_fold_child(icb_ctx* parent, icb_ctx* child)
{
    child->timers_collect();
    for (auto& kv : child->timers) {
	auto& t = parent->timers[kv.first];
	t.launches += kv.second.launches;
	t.total_ms += kv.second.total_ms;
    }
    child->timers.clear();
    parent->launch_count += child->launch_count;
    child->launch_count = 0;
}

void icb_ctx_destroy_child(icb_ctx* parent, icb_ctx* child)
{
    if (!child) return;
    cudaSetDevice(child->device);
    icb_ctx_fold_child(parent, child);
    for (auto e : child->event_pool) cudaEventDestroy(e);
    if (child->wait_event) cudaEventDestroy(child->wait_event);
    if (child->align_ws) {
	child->align_ws->release(child);
	delete child->align_ws;
    }
    cudaStreamSynchronize(child->stream);
    cudaStreamDestroy(child->stream);
    delete child;
}

extern "C" {

int icb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ICB_SEGV_TRACE=1: print the native backtrace of a crashing thread (addresses resolve with addr2line on the shipped .so)
static void icb_segv_handler(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    static const char msg[] = "[icb] fatal signal, native backtrace:\n";
    if (write(2, msg, sizeof msg - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

int icb_ctx_create(int device, icb_ctx** out)
{
    if (!out) return ICB_ERR_ARG;
    *out = nullptr;
    if (getenv("ICB_SEGV_TRACE")) {
	signal(SIGSEGV, icb_segv_handler);
	signal(SIGABRT, icb_segv_handler);
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return ICB_ERR_CUDA;  // no CPU fallback: fail loudly
    if (device < 0 || device >= n) return ICB_ERR_ARG;
    // many worker streams: without this they share 8 hardware queues and wait behind each other's kernels (only effective
    // if the process has not initialised CUDA yet; bench.py and the CLI also set it)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    auto ctx = new icb_ctx;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
	delete ctx;
	return ICB_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
	int least = 0, greatest = 0;
	if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) {
	    ctx->hi_priority = greatest;
	    ctx->lo_priority = ctx->aln_priority = least;
	}
    }
    {
	// memory returned to the stream-ordered pool stays cached in it (no trimming at synchronisation points)
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
	    uint64_t keep = UINT64_MAX;
	    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	}
    }
    *out = ctx;
    return ICB_OK;
}

void icb_ctx_destroy(icb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->timers_collect();
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
    icb_worker_pool_release(ctx);
    if (ctx->align_ws) {
	ctx->align_ws->release(ctx);
	delete ctx->align_ws;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    {
	// hand the memory the pool cached for this context's read sets and workers back to the device
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    delete ctx;
}

int icb_ctx_sync(icb_ctx* ctx)
{
    ICB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ICB_OK;
}

void* icb_ctx_stream(icb_ctx* ctx) { return (void*)ctx->stream; }
const char* icb_last_error(icb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

const char* icb_strerror(int code)
{
    switch (code) {
	case ICB_OK: return "ok";
	case ICB_ERR_CUDA: return "CUDA runtime error (no usable GPU or a failed call)";
	case ICB_ERR_ARG: return "invalid argument";
	case ICB_ERR_INVALID_BASE: return "Invalid base encountered";
	case ICB_ERR_NO_PEMP: return "Empirical probability lookup failure!";
	case ICB_ERR_STATE: return "call order violated";
	case ICB_ERR_NOMEM: return "out of memory";
    }
    return "unknown error";
}

int icb_ctx_timing(icb_ctx* ctx, int enable)
{
    ctx->timers_collect();
    ctx->timing = enable != 0;
    if (ctx->timing && getenv("ICB_TRACE") && !ctx->trace_base) {
	ICB_CUDA(ctx, cudaEventCreate(&ctx->trace_base));
	ICB_CUDA(ctx, cudaEventRecord(ctx->trace_base, ctx->stream));
	ICB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ICB_OK;
}

int icb_ctx_timing_reset(icb_ctx* ctx)
{
    ctx->timers_collect();
    ctx->timers.clear();
    return ICB_OK;
}

int icb_ctx_kernel_time(icb_ctx* ctx, const char* name, int64_t* launches, double* total_ms)
{
    ctx->timers_collect();
    auto it = ctx->timers.find(name);
    if (it == ctx->timers.end()) {
	if (launches) *launches = 0;
	if (total_ms) *total_ms = 0.0;
	return ICB_OK;
    }
    if (launches) *launches = it->second.launches;
    if (total_ms) *total_ms = it->second.total_ms;
    return ICB_OK;
}

int64_t icb_ctx_launch_count(icb_ctx* ctx) { return ctx->launch_count; }

// ---------------------------------------------------------------- read set

static const int64_t PAD = 256;	 // slack after the byte buffers so 16-byte tile loads may overrun

int icb_readset_create(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n_reads,
		       icb_readset** out)
{
    return icb_readset_create_reserve(ctx, seq, qual, offs, n_reads, 0, 0, 0, out);
}

int icb_readset_create_reserve(icb_ctx* ctx, const char* seq, const char* qual, const int64_t* offs, int64_t n_reads, int64_t extra_reads,
			       int64_t extra_bases, int64_t extra_max_len, icb_readset** out)
{
    IcbRange nvtx_range("upload reads (icb_readset_create)");
    if (!ctx || !out || n_reads < 0 || (n_reads > 0 && (!seq || !qual || !offs)) || extra_reads < 0 || extra_bases < 0 || extra_max_len < 0)
	return ICB_ERR_ARG;
    *out = nullptr;
    ICB_CUDA(ctx, cudaSetDevice(ctx->device));
    auto rs = new icb_readset;
    rs->ctx = ctx;
    rs->n_reads = n_reads;
    rs->h_offs.reserve(size_t(n_reads + extra_reads + 1));
    rs->h_offs.assign(offs, offs + n_reads + 1);
    if (n_reads == 0) rs->h_offs.assign(1, 0);
    for (int64_t i = 0; i < n_reads; i++) {
	int64_t l = offs[i + 1] - offs[i];
	if (l < 0 || l > 0x7fffff00LL) {
	    delete rs;
	    return ctx->fail(ICB_ERR_ARG, "offsets must be non-decreasing and reads shorter than 2^31");
	}
	if (l > rs->max_len) rs->max_len = l;
    }
    rs->n_bases = rs->h_offs[n_reads] - rs->h_offs[0];
    if (rs->h_offs[0] != 0) {
	delete rs;
	return ctx->fail(ICB_ERR_ARG, "offs[0] must be 0");
    }
    rs->cap_reads = n_reads + extra_reads;
    rs->cap_bases = rs->n_bases + extra_bases;
    if (extra_reads > 0 && extra_max_len > rs->max_len) rs->max_len = extra_max_len;  // alignment workspaces are sized by it
    const int64_t nb = rs->n_bases, cb = rs->cap_bases, cr = rs->cap_reads;
    cudaError_t e;
#define RS_ALLOC(ptr, bytes)                                           \
    if ((e = icb_dmalloc(ctx, (void**)&(ptr), (size_t)(bytes))) != cudaSuccess) { \
	icb_readset_destroy(rs);                                       \
	return ctx->fail_cuda(e, "cudaMalloc " #ptr, __FILE__, __LINE__);  \
    }
    RS_ALLOC(rs->d_seq, cb + PAD);
    RS_ALLOC(rs->d_qual, cb + PAD);
    RS_ALLOC(rs->d_offs, sizeof(int64_t) * (cr + 1));
    RS_ALLOC(rs->d_score, sizeof(double) * (cr + 1));
    RS_ALLOC(rs->d_err, sizeof(double) * (cr + 1));
    RS_ALLOC(rs->d_counters, sizeof(unsigned long long) * 8);
#undef RS_ALLOC
    ICB_CUDA(ctx, cudaMemsetAsync(rs->d_seq + nb, 0, size_t(cb - nb + PAD), ctx->stream));
    ICB_CUDA(ctx, cudaMemsetAsync(rs->d_qual + nb, 0, size_t(cb - nb + PAD), ctx->stream));
    if (nb > 0) {
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_seq, seq, nb, cudaMemcpyHostToDevice, ctx->stream));
	ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_qual, qual, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    ICB_CUDA(ctx, cudaMemcpyAsync(rs->d_offs, rs->h_offs.data(), sizeof(int64_t) * (n_reads + 1), cudaMemcpyHostToDevice,
				  ctx->stream));
    *out = rs;
    return ICB_OK;
}

void icb_readset_destroy(icb_readset* rs)
{
    if (!rs) return;
    cudaSetDevice(rs->ctx->device);
    cudaStreamSynchronize(rs->ctx->stream);
    // stream-ordered frees (the device pool keeps the memory for the next read set: a pipeline that uploads one read set
    // after the other pays the driver's map / unmap of gigabytes only once)
    icb_ctx* ctx = rs->ctx;
    for (void* q : {(void*)rs->d_seq, (void*)rs->d_qual, (void*)rs->d_offs, (void*)rs->d_score, (void*)rs->d_err, (void*)rs->d_status,
		    (void*)rs->d_hpc_len, (void*)rs->d_hpc_err, (void*)rs->d_hpc_seq, (void*)rs->d_hpc_qual, (void*)rs->d_mins,
		    (void*)rs->d_min_off, (void*)rs->d_nf, (void*)rs->d_nr, (void*)rs->d_counters, (void*)rs->d_long_list,
		    (void*)rs->d_long_scratch, (void*)rs->d_class, (void*)rs->d_bkt})
	icb_dfree(ctx, q);
    delete rs;
}

int64_t icb_readset_n_reads(const icb_readset* rs) { return rs->n_reads; }
int64_t icb_readset_n_bases(const icb_readset* rs) { return rs->n_bases; }

}  // extern "C"
