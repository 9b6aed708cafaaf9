// Host-side pieces of the hot path that the reference evaluates with the host's libm / libstdc++ and that
// must therefore stay on the host to remain bit-exact: the 10^(-Q/10) tables, the -10*log10(err) quality
// filter and the stable sort by score.  No device work here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "isonclust2_b200.h"

extern "C" {

// InitQualTab / InitQualTabNomin (reference qualscore.cpp:156-180)
void icb_init_qual_tab(double tab[129], int nomin)
{
    for (int i = 0; i < 129; i++) tab[i] = 0.0;
    for (int i = 33; i <= 128; i++) {
	double t = std::pow(10, -((i - 33) / 10.0));
	if (!nomin && t > 0.79433) t = 0.79433;
	tab[i] = t;
    }
}

// the read filter of PrepareSortedBatch / ClusterSortedReads (qualscore.cpp:56, cluster.cpp:157)
int icb_quality_filter(const double* err, int64_t n, double min_qual, uint8_t* filtered)
{
    if (n < 0 || (n > 0 && (!err || !filtered))) return ICB_ERR_ARG;
    for (int64_t i = 0; i < n; i++) filtered[i] = ((-10 * std::log10(err[i])) <= min_qual) ? 1 : 0;
    return ICB_OK;
}

// SortByQualScores (qualscore.cpp:138-145): std::stable_sort, a before b iff score(a) > score(b)
int icb_sort_by_score(const double* score, int64_t n, int32_t* order)
{
    if (n < 0 || n > 0x7fffffff || (n > 0 && (!score || !order))) return ICB_ERR_ARG;
    std::iota(order, order + n, 0);
    std::stable_sort(order, order + n, [score](int32_t a, int32_t b) { return score[a] > score[b]; });
    return ICB_OK;
}

}  // extern "C"
