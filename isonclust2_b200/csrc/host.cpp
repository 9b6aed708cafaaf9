// Host-side pieces of the hot path that the reference evaluates with the host's libm / libstdc++ and that
// must therefore stay on the host to remain bit-exact: the 10^(-Q/10) tables, the -10*log10(err) quality
// filter and the stable sort by score.  No device work here.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "helpers.h"
#include "isonclust2_b200.h"

namespace {
// fn(a, b) over [0, n) in pieces of at least `grain`, on up to 16 host threads
template <class F>
void parallel_ranges(int64_t n, int64_t grain, F fn)
{
    const int T = int(std::min<int64_t>(std::min<int64_t>(16, icb_host_cores()), (n + grain - 1) / grain));
    if (T <= 1) {
	fn(0, n);
	return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([=] { fn(n * t / T, n * (t + 1) / T); });
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

// InitMinSharedMap (p_emp_prob.cpp:22-47) for (k, w): rows with the same k and |w_row - w| <= 2, later rows overwriting
// earlier ones, symmetric in (e1, e2); the table ships as data/p_emp_prob.bin (tools/make_p_emp_table.py)
int icb_load_p_shared(const char* path, int kmer_size, int window_size, double p_shared[225])
{
    if (!p_shared) return ICB_ERR_ARG;
    for (int i = 0; i < 225; i++) p_shared[i] = std::numeric_limits<double>::quiet_NaN();
    std::string file;
    if (path) file = path;
    else if (const char* env = getenv("ICB_PEMP_TABLE")) file = env;
    else {
	Dl_info info;
	if (!dladdr(reinterpret_cast<void*>(&icb_load_p_shared), &info) || !info.dli_fname) return ICB_ERR_STATE;
	std::string so(info.dli_fname);
	const size_t s = so.rfind('/');
	file = (s == std::string::npos ? std::string(".") : so.substr(0, s)) + "/data/p_emp_prob.bin";
    }
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) return ICB_ERR_ARG;
    char magic[8];
    uint32_t n = 0;
    struct __attribute__((packed)) Row {
	uint8_t k, w, c1, c2;
	double p;
    };
    std::vector<Row> rows;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "ICBPEMP1", 8) == 0 && fread(&n, 4, 1, f) == 1;
    if (ok) {
	try {
	    rows.resize(n);
	}
	catch (...) {
	    fclose(f);
	    return ICB_ERR_NOMEM;
	}
	ok = fread(rows.data(), sizeof(Row), n, f) == n;
    }
    fclose(f);
    if (!ok) return ICB_ERR_ARG;
    for (const Row& r : rows) {
	if (int(r.k) != kmer_size || std::abs(int(r.w) - window_size) > 2) continue;
	if (r.c1 < 1 || r.c1 > 15 || r.c2 < 1 || r.c2 > 15) continue;
	p_shared[(r.c1 - 1) * 15 + (r.c2 - 1)] = r.p;
	p_shared[(r.c2 - 1) * 15 + (r.c1 - 1)] = r.p;
    }
    return ICB_OK;
}

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
static int quality_filter_impl(const double* err, int64_t n, double min_qual, uint8_t* filtered)
{
    if (n < 0 || (n > 0 && (!err || !filtered))) return ICB_ERR_ARG;
    parallel_ranges(n, 1 << 16, [&](int64_t a, int64_t b) {
	for (int64_t i = a; i < b; i++) filtered[i] = ((-10 * std::log10(err[i])) <= min_qual) ? 1 : 0;
    });
    return ICB_OK;
}

// SortByQualScores (qualscore.cpp:138-145): std::stable_sort, a before b iff score(a) > score(b).  The stable order by a
// strict weak ordering is unique — (score descending, index ascending) — so sorted runs merged stably give the same
// permutation as one std::stable_sort; with a NaN among the scores the comparator is no strict weak ordering and the
// single library call is kept.
static int sort_by_score_impl(const double* score, int64_t n, int32_t* order)
{
    if (n < 0 || n > 0x7fffffff || (n > 0 && (!score || !order))) return ICB_ERR_ARG;
    std::iota(order, order + n, 0);
    auto cmp = [score](int32_t a, int32_t b) { return score[a] > score[b]; };
    bool nan = false;
    for (int64_t i = 0; i < n; i++) nan |= score[i] != score[i];
    const int T = nan || n < (1 << 16) ? 1 : std::min<int>(16, icb_host_cores());
    if (T <= 1) {
	std::stable_sort(order, order + n, cmp);
	return ICB_OK;
    }
    std::vector<int64_t> cut(T + 1);
    for (int t = 0; t <= T; t++) cut[t] = n * t / T;
    {
	std::vector<std::thread> th;
	for (int t = 0; t < T; t++) th.emplace_back([&, t] { std::stable_sort(order + cut[t], order + cut[t + 1], cmp); });
	for (auto& x : th) x.join();
    }
    std::vector<int32_t> tmp(static_cast<size_t>(n));
    int32_t *src = order, *dst = tmp.data();
    for (int width = 1; width < T; width *= 2) {
	std::vector<std::thread> th;
	for (int t = 0; t < T; t += 2 * width) {
	    const int64_t a = cut[t], m = cut[std::min(t + width, T)], b = cut[std::min(t + 2 * width, T)];
	    th.emplace_back([=] { std::merge(src + a, src + m, src + m, src + b, dst + a, cmp); });  // ties: the left run first
	}
	for (auto& x : th) x.join();
	std::swap(src, dst);
    }
    if (src != order) std::copy(src, src + n, order);
    return ICB_OK;
}

// the C ABI never throws: thread creation (std::system_error) or allocation failures inside become error codes
int icb_quality_filter(const double* err, int64_t n, double min_qual, uint8_t* filtered)
{
    try {
	return quality_filter_impl(err, n, min_qual, filtered);
    }
    catch (const std::bad_alloc&) {
	return ICB_ERR_NOMEM;
    }
    catch (...) {
	return ICB_ERR_STATE;
    }
}

int icb_sort_by_score(const double* score, int64_t n, int32_t* order)
{
    try {
	return sort_by_score_impl(score, n, order);
    }
    catch (const std::bad_alloc&) {
	return ICB_ERR_NOMEM;
    }
    catch (...) {
	return ICB_ERR_STATE;
    }
}

}  // extern "C"
