// p_emp_prob.h of the reference's header face (reference src/p_emp_prob.h:16-19, p_emp_prob.cpp:22-94).
#ifndef P_EMP_PROB_H_INCLUDED
#define P_EMP_PROB_H_INCLUDED

#include <cmath>
#include <stdexcept>

#include "icb_compat.h"

// the 15 x 15 grid of empirical shared-minimizer probabilities for one (k, w): error classes 0.01 .. 0.15
struct MinSharedMap {
    double p[225];
};

inline MinSharedMap InitMinSharedMap(int kmerSize, int windowSize)
{
    MinSharedMap m;
    const int rc = icb_load_p_shared(nullptr, kmerSize, windowSize, m.p);
    if (rc != ICB_OK) throw std::runtime_error("isonclust2_compat: InitMinSharedMap: cannot read the empirical probability table");
    return m;
}

// errors are rounded to two decimals (util.cpp:6-10) and clamped to [0.01, 0.15] before the lookup
inline double GetPMinShared(double e1, double e2, const MinSharedMap& msm)
{
    auto cls = [](double e) {
	double r = std::round(e * 100) / 100;
	if (r > 0.15) r = 0.15;
	if (r < 0.01) r = 0.01;
	int c = int(std::lround(r * 100.0)) - 1;
	return c < 0 ? 0 : (c > 14 ? 14 : c);
    };
    const double v = msm.p[cls(e1) * 15 + cls(e2)];
    if (!(v == v)) throw std::invalid_argument("Empirical probability lookup failure!");
    return v;
}

#endif
