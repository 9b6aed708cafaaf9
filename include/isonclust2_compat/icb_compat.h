// isonclust2_compat — the reference's C++ header face (hpc.h, kmer_index.h, minimizer.h, qualscore.h, cluster.h,
// p_emp_prob.h) implemented over the C ABI of libisonclust2_b200.so.  Header-only: put this directory on
// the include path in place of the reference's src/ and link -lisonclust2_b200; the reference's own unit test
// (test/isONclust2_test.cpp) compiles unchanged against it (tests/test_compat_gpu.py).  Every function that is arithmetic
// of the hot path calls the device through the per-primitive entry points of include/isonclust2_b200.h; what the
// reference defines through host libm / libstdc++ (pow thresholds, the hit map's iteration order, std::sort) stays host
// code here, exactly as the reference has it.
#ifndef ISONCLUST2_COMPAT_H_INCLUDED
#define ISONCLUST2_COMPAT_H_INCLUDED

#include <stdexcept>
#include <string>

#include "isonclust2_b200.h"

namespace icb_compat {

// one lazily created context on device $ICB_COMPAT_DEVICE (default 0) for the single-read calls of this face
inline icb_ctx* ctx()
{
    static icb_ctx* c = [] {
	icb_ctx* x = nullptr;
	const char* d = getenv("ICB_COMPAT_DEVICE");
	const int rc = icb_ctx_create(d ? atoi(d) : 0, &x);
	if (rc != ICB_OK) throw std::runtime_error(std::string("isonclust2_compat: no CUDA context: ") + icb_strerror(rc));
	return x;
    }();
    return c;
}

inline void check(int rc, const char* what)
{
    if (rc != ICB_OK) throw std::runtime_error(std::string("isonclust2_compat: ") + what + ": " + icb_last_error(ctx()) + " (" + icb_strerror(rc) + ")");
}

}  // namespace icb_compat

#endif
