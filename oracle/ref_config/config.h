/* Hand-written stand-in for the header parasail's configure step would generate
 * (vendor/parasail/cmake/config.h.in).  x86-64 Linux, gcc >= 9, AVX2 capable
 * compiler.  Only used to compile the REFERENCE oracle (oracle/_ref); test
 * infrastructure, not product code. */
#ifndef ORACLE_REF_PARASAIL_CONFIG_H
#define ORACLE_REF_PARASAIL_CONFIG_H
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_ALIGNED_ALLOC 1
#define HAVE_MEMALIGN 1
#define HAVE_MALLOC_H 1
#define SIZEOF_INT 4
#define HAVE_XGETBV 1
#define HAVE_SSE2 1
#define HAVE_SSE2_MM_SET1_EPI64X 1
#define HAVE_SSE2_MM_SET_EPI64X 1
#define HAVE_SSE41 1
#define HAVE_SSE41_MM_INSERT_EPI64 1
#define HAVE_SSE41_MM_EXTRACT_EPI64 1
#define HAVE_AVX2 1
#define HAVE_AVX2_MM256_SET1_EPI64X 1
#define HAVE_AVX2_MM256_SET_EPI64X 1
#define HAVE_AVX2_MM256_INSERT_EPI64 1
#define HAVE_AVX2_MM256_INSERT_EPI32 1
#define HAVE_AVX2_MM256_INSERT_EPI16 1
#define HAVE_AVX2_MM256_INSERT_EPI8 1
#define HAVE_AVX2_MM256_EXTRACT_EPI64 1
#define HAVE_AVX2_MM256_EXTRACT_EPI32 1
#define HAVE_AVX2_MM256_EXTRACT_EPI16 1
#define HAVE_AVX2_MM256_EXTRACT_EPI8 1
#define HAVE_ALTIVEC 0
#define HAVE_NEON 0
#define HAVE_ZLIB 1
#define WORDS_BIGENDIAN 0
#define HAVE_SYS_MMAN_H 1
#define HAVE_UNISTD_H 1
#define HAVE_GETOPT 1
#define HAVE_POLL 1
#define HAVE___BUILTIN_CPU_INIT 1
#define HAVE___BUILTIN_CPU_SUPPORTS 1
#define HAVE_GETSYSTEMTIMEASFILETIME 0
#define HAVE_CLOCK_GET_TIME 0
#define HAVE_CLOCK_GETTIME_MONOTONIC 1
#define HAVE_CLOCK_GETTIME_REALTIME 1
#define INT64_LITERAL_SUFFIX_I64 0
#define INT64_LITERAL_SUFFIX_LL 1
#ifndef __cplusplus
#define restrict __restrict
#endif
#endif
