/* TEST INFRASTRUCTURE — oracle build only.
 *
 * parasail's own sg_dispatch.c (vendor/parasail/src/sg_dispatch.c:6842-6858)
 * routes parasail_sg_trace_scan_16/_32 through function pointers that pick
 * AVX2, then SSE4.1, then SSE2, then the scalar implementation.  Linking that
 * file drags in every one of parasail's ~1000 generated kernels, so the
 * oracle build compiles only the sg_trace_scan kernels the reference reaches
 * (cluster.cpp:413-419) and keeps the same preference order here. */
#include "config.h"
#include "parasail.h"
#include "parasail/cpuid.h"

#define DISPATCH(BITS)                                                         \
    parasail_result_t* parasail_sg_trace_scan_##BITS(                          \
	const char* const restrict s1, const int s1Len,                        \
	const char* const restrict s2, const int s2Len, const int open,        \
	const int gap, const parasail_matrix_t* matrix)                        \
    {                                                                          \
	if (parasail_can_use_avx2())                                           \
	    return parasail_sg_trace_scan_avx2_256_##BITS(s1, s1Len, s2,       \
							  s2Len, open, gap,    \
							  matrix);             \
	if (parasail_can_use_sse41())                                          \
	    return parasail_sg_trace_scan_sse41_128_##BITS(s1, s1Len, s2,      \
							   s2Len, open, gap,   \
							   matrix);            \
	if (parasail_can_use_sse2())                                           \
	    return parasail_sg_trace_scan_sse2_128_##BITS(s1, s1Len, s2,       \
							  s2Len, open, gap,    \
							  matrix);             \
	return parasail_sg_trace_scan(s1, s1Len, s2, s2Len, open, gap,         \
				      matrix);                                 \
    }

DISPATCH(16)
DISPATCH(32)
