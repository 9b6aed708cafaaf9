// Host test of csrc/ref_order.h: the computed iteration order + sort must equal what the real libstdc++ container and
// std::sort give for the reference's construction (minimizer.cpp:44-76, cluster.cpp:622-636).
#include <cstdio>
#include <random>

#include "ref_order.h"

struct Cand {
    int32_t id, strand, count, tmapped;
};

int main()
{
    std::mt19937 rng(12345);
    IcbRefOrderScratch S;
    std::vector<IcbSortedCand> got;
    int bad = 0;
    for (int it = 0; it < 3000; it++) {
	const int ncls = 1 + int(rng() % (it % 7 == 0 ? 9000 : 700));
	const int nf = 1 + int(rng() % 600), nr = 1 + int(rng() % 600);
	const size_t hint = (it % 11 == 0) ? size_t(1 + rng() % 50) : size_t(20) * size_t(nf + nr);
	std::vector<Cand> v;
	for (int id = 0; id < ncls; id++)
	    for (int s = 0; s < 2; s++)
		if (rng() % 3) v.push_back(Cand{id, s ? -1 : 1, 1 + int(rng() % 6), int(rng() % (s ? nr : nf))});
	// the reference: insertion in (strand, first hit, id) order into the real container, then std::sort
	std::vector<Cand> ins(v);
	std::stable_sort(ins.begin(), ins.end(), [](const Cand& a, const Cand& b) {
	    if (a.strand != b.strand) return a.strand > b.strand;
	    return a.tmapped < b.tmapped;
	});
	std::unordered_map<std::pair<int, int>, unsigned, IcbPairHash> m(hint);
	for (auto& c : ins) m[std::make_pair(c.id, c.strand)] = unsigned(c.count);
	std::vector<IcbSortedCand> want;
	for (auto& kv : m) want.push_back(IcbSortedCand{kv.second, unsigned(kv.first.first), kv.first.second});
	std::sort(want.begin(), want.end(), [](const IcbSortedCand& a, const IcbSortedCand& b) { return a.Size > b.Size; });
	icb_reference_order(v.data(), int(v.size()), hint, S, got);
	bool same = got.size() == want.size();
	for (size_t i = 0; same && i < got.size(); i++)
	    same = got[i].Size == want[i].Size && got[i].Cls == want[i].Cls && got[i].Strand == want[i].Strand;
	if (!same) {
	    bad++;
	    if (bad < 5) fprintf(stderr, "case %d differs (n=%zu hint=%zu)\n", it, v.size(), hint);
	}
    }
    printf("%s: %d of 3000 cases differ\n", bad ? "FAIL" : "ok", bad);
    return bad ? 1 : 0;
}
