// Host test of csrc/ref_order.h: the computed iteration order + sort must equal what the real libstdc++ container and
// std::sort give for the reference's construction (minimizer.cpp:44-76, cluster.cpp:622-636).
#include <cstdio>
#include <random>

#include "ref_order.h"
#include "ref_order_hd.h"

struct Cand {
    int32_t id, strand, count, tmapped;
};

int main()
{
    std::mt19937 rng(12345);
    IcbRefOrderScratch S;
    std::vector<IcbSortedCand> got, grp;
    int bad = 0, bad_group = 0, bad_hd = 0;
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
	// the group-only variant: for every count value, the members of that group in the order of `want`
	for (unsigned target = 1; target <= 6; target++) {
	    if (it % 3 && target != 6 && target != unsigned(1 + it % 6)) continue;
	    icb_reference_order_group(v.data(), int(v.size()), hint, target, S, grp);
	    size_t q = 0;
	    bool ok = true;
	    for (auto& w : want)
		if (w.Size == target) {
		    ok = ok && q < grp.size() && grp[q].Size == w.Size && grp[q].Cls == w.Cls && grp[q].Strand == w.Strand;
		    q++;
		}
	    ok = ok && q == grp.size();
	    // the same from a list the caller keeps in insertion order (what csrc/cluster.cu does), also after dropping elements
	    {
		std::vector<int32_t> perm;
		icb_insertion_order(v.data(), int(v.size()), S, perm);
		std::vector<Cand> pv;
		for (int32_t i : perm) pv.push_back(v[i]);
		std::vector<IcbSortedCand> grp2;
		icb_reference_order_group<true>(pv.data(), int(pv.size()), hint, target, S, grp2);
		ok = ok && grp2.size() == grp.size();
		for (size_t i = 0; ok && i < grp.size(); i++) ok = grp2[i].Cls == grp[i].Cls && grp2[i].Strand == grp[i].Strand && grp2[i].Size == grp[i].Size;
		if (it % 5 == 0) {
		    std::vector<Cand> sub_id, sub_ins;
		    for (auto& c : v)
			if (c.id % 3) sub_id.push_back(c);
		    for (auto& c : pv)
			if (c.id % 3) sub_ins.push_back(c);
		    std::vector<IcbSortedCand> a, b;
		    icb_reference_order_group(sub_id.data(), int(sub_id.size()), hint, target, S, a);
		    icb_reference_order_group<true>(sub_ins.data(), int(sub_ins.size()), hint, target, S, b);
		    ok = ok && a.size() == b.size();
		    for (size_t i = 0; ok && i < a.size(); i++) ok = a[i].Cls == b[i].Cls && a[i].Strand == b[i].Strand;
		}
	    }
	    if (!ok) {
		bad_group++;
		if (bad_group < 5) fprintf(stderr, "case %d group %u differs (n=%zu hint=%zu)\n", it, target, v.size(), hint);
	    }
	}
    }
    // skewed count distributions (few large values, like real candidate lists), adversarial patterns for the depth budget
    for (int it = 0; it < 4000; it++) {
	const int n = 1 + int(rng() % (it % 5 == 0 ? 6000 : 1500));
	const int nf = 300 + int(rng() % 300), nr = 300 + int(rng() % 300);
	const size_t hint = size_t(20) * size_t(nf + nr);
	std::vector<Cand> v;
	unsigned top = 0;
	for (int id = 0; v.size() < size_t(n); id++)
	    for (int s = 0; s < 2; s++)
		if (rng() % 2) {
		    int cnt = 1;
		    if (it % 4 == 3) cnt = 1 + (id % 7 == 0 ? int(rng() % 3) : 0) + (it % 8 == 7 ? id % 50 : 0);
		    else while (cnt < 9 && rng() % 3 == 0) cnt++;
		    top = std::max(top, unsigned(cnt));
		    v.push_back(Cand{id, s ? -1 : 1, cnt, int(rng() % (s ? nr : nf))});
		}
	icb_reference_order(v.data(), int(v.size()), hint, S, got);
	for (unsigned target : {top, top > 1 ? top - 1 : 1u, 1u}) {
	    icb_reference_order_group(v.data(), int(v.size()), hint, target, S, grp);
	    size_t q = 0;
	    bool ok = true;
	    for (auto& w : got)
		if (w.Size == target) {
		    ok = ok && q < grp.size() && grp[q].Cls == w.Cls && grp[q].Strand == w.Strand;
		    q++;
		}
	    ok = ok && q == grp.size();
	    if (!ok) {
		bad_group++;
		if (bad_group < 5) fprintf(stderr, "skewed case %d group %u differs (n=%zu)\n", it, target, v.size());
	    }
	}
    }
    // Musser's median-of-three killer (drives std::sort into its heap-sort fallback), with and without ties: the pruned
    // introsort loop against the real std::sort on the same keys
    for (int it = 0; it < 300; it++) {
	using namespace icb_introsort;
	int n = (200 + int(rng() % 3000)) & ~1;
	std::vector<uint32_t> sz(n);
	const int h = n / 2;
	for (int i = 1; i <= h; i++) {
	    if (i & 1) {
		sz[i - 1] = i;
		sz[i] = h + i;
	    }
	    sz[h + i - 1] = 2 * i;
	}
	const int coarse = 1 + (it % 4 == 0 ? 0 : int(rng() % 8));
	std::vector<Key> a(n), b;
	for (int i = 0; i < n; i++) {
	    uint32_t v = sz[i] / coarse + 1;
	    if (it & 1) v = n + 2 - v;
	    a[i] = uint64_t(v) << 32 | uint32_t(i);
	}
	b = a;
	std::sort(b.begin(), b.end(), [](Key x, Key y) { return gt(x, y); });
	for (int q = 0; q < 4; q++) {
	    const uint32_t target = q == 0 ? uint32_t(b[0] >> 32) : uint32_t(b[rng() % n] >> 32);
	    std::vector<Key> c = a;
	    long lg = 0;
	    for (size_t v = n; v > 1; v >>= 1) lg++;
	    loop(c.data(), c.data() + n, 2 * lg, target);
	    std::vector<uint32_t> g1, g2;
	    for (auto x : c)
		if (uint32_t(x >> 32) == target) g1.push_back(uint32_t(x));
	    for (auto x : b)
		if (uint32_t(x >> 32) == target) g2.push_back(uint32_t(x));
	    if (g1 != g2) bad_group++;
	    // the library-free restatement that also runs on the device (ref_order_hd.h)
	    std::vector<Key> e = a;
	    icb_sort_hd::introsort_group(e.data(), long(n), target);
	    std::vector<uint32_t> g3;
	    for (auto x : e)
		if (uint32_t(x >> 32) == target) g3.push_back(uint32_t(x));
	    if (g3 != g2) bad_hd++;
	}
    }
    // ref_order_hd.h against std::sort on random keys with heavy ties (what candidate lists look like) and on all-distinct keys
    for (int it = 0; it < 3000; it++) {
	const int n = 1 + int(rng() % (it % 6 == 0 ? 9000 : 2500));
	const uint32_t span = it % 3 == 0 ? 6u : (it % 3 == 1 ? 40u : 100000u);
	std::vector<uint64_t> a(n), b;
	for (int i = 0; i < n; i++) {
	    uint32_t v = 1 + rng() % span;
	    if (it % 4 == 1) while (v > 1 && rng() % 3) v--;  // skewed towards small counts
	    a[i] = uint64_t(v) << 32 | uint32_t(i);
	}
	b = a;
	std::sort(b.begin(), b.end(), [](uint64_t x, uint64_t y) { return (x >> 32) > (y >> 32); });
	for (int q = 0; q < 3; q++) {
	    const uint32_t target = q == 0 ? uint32_t(b[0] >> 32) : uint32_t(b[rng() % n] >> 32);
	    std::vector<uint64_t> e = a;
	    icb_sort_hd::introsort_group(e.data(), long(n), target);
	    std::vector<uint32_t> g2, g3;
	    for (auto x : b)
		if (uint32_t(x >> 32) == target) g2.push_back(uint32_t(x));
	    for (auto x : e)
		if (uint32_t(x >> 32) == target) g3.push_back(uint32_t(x));
	    if (g3 != g2) bad_hd++;
	}
    }
    printf("%s: %d of 3000 cases differ, %d group queries differ, %d host/device restatement queries differ\n",
	   (bad || bad_group || bad_hd) ? "FAIL" : "ok", bad, bad_group, bad_hd);
    return (bad || bad_group || bad_hd) ? 1 : 0;
}
