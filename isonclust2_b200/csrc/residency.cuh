// Diagnostics (icb_debug_residency_start / _dump, include/isonclust2_b200.h): while the log runs, the kernels of the device
// loop leave 16-byte records {kind | SM id << 4 | tag << 12, start, end, payload}, times in units of 1.024 us of the device's
// global timer.  kind 1 = one CTA of the packed alignment kernel (payload n | m << 16), 2 = one warp of its traceback,
// 3.. = a MARK at the start of a kernel of a round (start = end = the time its first CTA began; enum IcbResKind), from which
// tools/round_timeline.py rebuilds where the rounds of one batch spend their time; tag = bits of the batch's pair-queue
// address (one value per batch in flight).  Every translation unit keeps its own copy of the log descriptor (no
// relocatable device code in this build); icb_residency_set_<unit> installs it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct IcbResidencyLog {
    uint4* rec;
    unsigned int* count;
    unsigned int cap;
};
enum IcbResKind { RES_DP_CTA = 1, RES_TRACE_WARP = 2, RES_COMMIT = 3, RES_DECIDE = 4, RES_ORDER = 5, RES_ORDER_L = 6, RES_GATHER = 7,
		  RES_QUERY_FULL = 8, RES_DP = 9, RES_TRACE = 10, RES_DP32 = 11, RES_QUERY_POOL = 12, RES_FINALIZE = 13, RES_INIT = 14 };

static __device__ IcbResidencyLog g_icb_residency = {nullptr, nullptr, 0u};

static inline cudaError_t icb_residency_set_here(const IcbResidencyLog& L) { return cudaMemcpyToSymbol(g_icb_residency, &L, sizeof L); }

__device__ __forceinline__ unsigned int icb_res_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return static_cast<unsigned int>(t >> 10);
}
__device__ __forceinline__ unsigned int icb_res_tag(const void* p) { return (unsigned int)((unsigned long long)p >> 8) & 0xfffffu; }
__device__ __forceinline__ void icb_res_record(int kind, unsigned int tag, unsigned int t0, unsigned int payload)
{
    const IcbResidencyLog L = g_icb_residency;
    if (!L.rec) return;
    unsigned int smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    const unsigned int at = atomicAdd(L.count, 1u);
    if (at < L.cap) L.rec[at] = make_uint4(unsigned(kind) | ((smid & 0xffu) << 4) | (tag << 12), t0, icb_res_now(), payload);
}
// a mark at the start of a kernel: thread 0 of block 0
__device__ __forceinline__ void icb_res_mark(int kind, const void* tag_ptr)
{
    if (g_icb_residency.rec && blockIdx.x == 0 && threadIdx.x == 0) icb_res_record(kind, icb_res_tag(tag_ptr), icb_res_now(), 0u);
}
