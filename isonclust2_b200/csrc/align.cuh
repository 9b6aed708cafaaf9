// Stage 3 internals shared by the two alignment back ends (align16.cu: packed 16-bit CTA-per-pair wavefront,
// align.cu: 32-bit warp-per-pair kernel for very long or non-ACGT sequences) and their host dispatcher.
#pragma once
#include "icb_common.cuh"

// one alignment problem, as both back ends see it
struct IcbPairDesc {
    int64_t q_off, r_off;  // byte offsets of read / representative in the sequence buffer
    int32_t n, m;	   // lengths (query rows, reference columns)
    int32_t rc;		   // representative is taken reverse-complemented
    int32_t open;	   // gap open
    int32_t limit;	   // a window counts if matches >= limit
    int32_t out;	   // slot in the caller's result array
    int64_t trace_off;	   // in 32-byte tiles
    int64_t line_off;	   // in 32-bit words: boundary lines
    int32_t ncg;	   // align16: tiles per row block
    int32_t pad;
};

// end cell of the DP, handed from the DP kernel to the traceback kernel
struct IcbAlignEnd {
    int32_t score, ei, ej;
    int32_t status;  // 0 ok, 1 the packed kernel refused the pair (a character outside ACGT): rerun on the 32-bit path
};

// growable device / pinned-host buffers of one context's alignment calls (no allocation in steady state)
struct IcbAlignWorkspace {
    uint8_t* d_trace = nullptr;
    size_t trace_bytes = 0;
    int* d_lines = nullptr;
    size_t lines_bytes = 0;
    IcbPairDesc* d_descs = nullptr;
    IcbAlignEnd* d_ends = nullptr;
    icb_align_result* d_out = nullptr;
    unsigned long long* d_ticket = nullptr;
    IcbPairDesc* h_descs = nullptr;  // pinned
    icb_align_result* h_out = nullptr;
    IcbAlignEnd* h_ends = nullptr;
    int64_t pair_cap = 0;
    int ensure(icb_ctx* ctx, size_t trace_need, size_t lines_need, int64_t pairs_need);
    void release(icb_ctx* ctx);
};

// upper-cased character of the reference string as parasail sees it (RevComp applied when rc; util.cpp:13-38)
__device__ __forceinline__ uint32_t icb_ref_char(const uint8_t* __restrict__ seq, const IcbPairDesc& d, int j)
{
    if (!d.rc) return seq[d.r_off + j];
    uint32_t c = seq[d.r_off + (d.m - 1 - j)];
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}
__device__ __forceinline__ uint32_t icb_up_char(uint32_t c) { return (c >= 'a' && c <= 'z') ? c - 32u : c; }

// getAlnRatio's window count (cluster.cpp:442-459) fed one comparison-string column at a time, in traceback
// (reverse) order — the count is symmetric under reversal except for which end window is skipped: the
// reference never tests the LAST window of the forward string, i.e. the FIRST one completed here.
struct IcbWindowCounter {
    uint32_t reg, mask;
    int k, limit;
    long long pushed;
    int aligned;
    uint8_t* rec;  // icb_align_traceback only: the comparison string, one byte per column, in traceback (reverse) order
    __device__ __forceinline__ void init(int kk, int lim, uint8_t* record = nullptr)
    {
	rec = record;
	reg = 0;
	mask = kk >= 32 ? 0xffffffffu : ((1u << kk) - 1u);
	k = kk;
	limit = lim;
	pushed = 0;
	aligned = 0;
    }
    __device__ __forceinline__ void push(uint32_t bit)
    {
	reg = ((reg << 1) | bit) & mask;
	if (pushed >= k && __popc(reg) >= limit) aligned++;
	if (rec) rec[pushed] = bit ? uint8_t('|') : uint8_t(' ');
	pushed++;
    }
    __device__ void push_zeros(long long g)
    {
	long long first = g < k ? g : k;
	for (long long t = 0; t < first; t++) push(0);
	long long rest = g - first;
	if (rest > 0) {
	    // the register is all zero now; every further column closes an all-blank window
	    long long counted = rest;
	    if (pushed < k) counted -= (k - pushed);
	    if (counted > 0 && 0 >= limit) aligned += int(counted);
	    if (rec)
		for (long long t = 0; t < rest; t++) rec[pushed + t] = uint8_t(' ');
	    pushed += rest;
	}
    }
};

// back ends (host launchers): descs are in pinned/host memory, results land in ws.h_out[desc.out]
int icb_align16_run(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* descs, int64_t cnt,
		    int64_t tiles, int64_t line_words, int max_m, int k);
int icb_align32_run(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* descs, int64_t cnt,
		    int64_t tiles, int64_t line_words, int k);
// device-built pair queues (dloop.cu): *d_count descriptors in device memory, window counts to aln_out[desc.out]
int icb_align16_run_dev(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* d_descs, const int32_t* d_count,
			int max_pairs, int max_m, int k, int32_t* aln_out);
int icb_align32_run_dev(icb_ctx* ctx, IcbAlignWorkspace& ws, const uint8_t* d_seq, const IcbPairDesc* d_descs, const int32_t* d_count,
			int max_pairs, int k, int32_t* aln_out);
// the dispatcher behind icb_align_pairs, on an explicit context (stream)
int icb_align_pairs_on(icb_ctx* ctx, icb_readset* rs, const icb_pair* pairs, int64_t n_pairs, int kmer_size,
		       icb_align_result* results);
