// poa.cpp — partial-order alignment consensus on the HOST: the piece of the consensus path (`sort -c <positive>`) that
// north_star keeps on the host.  It restates, bit for bit, what the reference gets from its vendored spoa 4.0.9 when it is
// built with AVX2 (vendor/spoa/CMakeLists.txt:37,44-50: -march=native), as driven by consensus.cpp:15-22 (Align +
// AddAlignment), :88-94 (GenerateConsensus) and :128-137 (Clear + re-add), engine parameters main.cpp:284-324
// (m 4, n -8, g -8, e -4, q -20, c -1: convex gaps):
//
//   PoaGraph            spoa::Graph: AddAlignment graph.cpp:159-247, TopologicalSort :249-300, TraverseHeaviestBundle /
//                       BranchCompletion :441-528, Clear :648-657, save / load graph.hpp:188-276 (the `.cer` ConsGs blobs)
//   poa_align<T, N>     SimdAlignmentEngine::Align / Initialize / Convex, simd_alignment_engine_implementation.hpp:518-780,
//                       1586-2131.  The SIMD engine is not the textbook recurrence: the horizontal gap states E / Q are
//                       prefix maxima taken block by block (N = 16 lanes of int16, or 8 of int32 when the worst-case score
//                       does not fit int16, :718-776), a block's E only sees the cells of its own block BEFORE the
//                       horizontal states were folded in, arithmetic wraps in the lane type, the row maximum of a sink
//                       row starts at 0, and the backtrack compares in int.  All of that decides ties, so it is emulated
//                       lane by lane: the per-predecessor pass as plain loops the compiler vectorises, the block pass on
//                       GCC vectors (PoaVec).  Host-only file, compiled by g++ (not through nvcc).
//   linear / affine gap modes (other engine parameters) are not built: the reference hard-codes the convex parameters.
//
// No CUDA in this file; it is part of the library so that the CLI and the tests reach it through the C ABI (icb_poa_*).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "isonclust2_b200.h"

namespace {

typedef std::vector<std::pair<int32_t, int32_t>> PoaAlignment;  // (node id or -1, sequence position or -1)

struct PoaGraph {
    struct Node {
	uint32_t code;
	std::vector<uint32_t> in, out;	// edge indices, in the order spoa appends them
	std::vector<uint32_t> aligned;	// node ids
    };
    struct Edge {
	uint32_t tail, head;
	std::vector<uint32_t> labels;
	int64_t weight;
    };
    uint32_t num_codes = 0;
    int32_t coder[256], decoder[256];
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    std::vector<uint32_t> sequences;  // first node of every sequence
    std::vector<uint32_t> rank_to_node;
    std::vector<uint32_t> consensus;

    PoaGraph() { clear(); }

    void clear()  // graph.cpp:648-657
    {
	num_codes = 0;
	for (int i = 0; i < 256; i++) coder[i] = decoder[i] = -1;
	nodes.clear();
	edges.clear();
	sequences.clear();
	rank_to_node.clear();
	consensus.clear();
    }

    uint32_t add_node(uint32_t code)
    {
	nodes.emplace_back();
	nodes.back().code = code;
	return uint32_t(nodes.size() - 1);
    }

    void add_edge(uint32_t tail, uint32_t head, uint32_t weight)  // graph.cpp:80-91
    {
	for (uint32_t e : nodes[tail].out)
	    if (edges[e].head == head) {
		edges[e].labels.push_back(uint32_t(sequences.size()));
		edges[e].weight += weight;
		return;
	    }
	Edge ed;
	ed.tail = tail;
	ed.head = head;
	ed.labels.assign(1, uint32_t(sequences.size()));
	ed.weight = weight;
	edges.push_back(std::move(ed));
	const uint32_t id = uint32_t(edges.size() - 1);
	nodes[tail].out.push_back(id);
	nodes[head].in.push_back(id);
    }

    // graph.cpp:93-110; returns the first node or -1
    int64_t add_sequence(const unsigned char* s, const std::vector<uint32_t>& w, uint32_t begin, uint32_t end)
    {
	if (begin == end) return -1;
	int64_t prev = -1;
	for (uint32_t i = begin; i < end; i++) {
	    const uint32_t cur = add_node(uint32_t(coder[s[i]]));
	    if (prev >= 0) add_edge(uint32_t(prev), cur, w[i - 1] + w[i]);
	    prev = cur;
	}
	return int64_t(nodes.size()) - int64_t(end - begin);
    }

    void add_alignment(const PoaAlignment& aln, const unsigned char* s, uint32_t len, uint32_t weight)	// graph.cpp:119-247
    {
	if (len == 0) return;
	std::vector<uint32_t> w(len, weight);
	for (uint32_t i = 0; i < len; i++)
	    if (coder[s[i]] == -1) {
		coder[s[i]] = int32_t(num_codes);
		decoder[num_codes++] = s[i];
	    }
	if (aln.empty()) {
	    sequences.push_back(uint32_t(add_sequence(s, w, 0, len)));
	    topological_sort();
	    return;
	}
	std::vector<uint32_t> valid;
	for (const auto& it : aln)
	    if (it.second != -1) {
		if (it.second < 0 || it.second >= int32_t(len)) throw std::invalid_argument("invalid alignment");
		valid.push_back(uint32_t(it.second));
	    }
	if (valid.empty()) throw std::invalid_argument("missing sequence in alignment");
	int64_t begin = add_sequence(s, w, 0, valid.front());
	int64_t prev = begin >= 0 ? int64_t(nodes.size()) - 1 : -1;
	const int64_t last = add_sequence(s, w, valid.back() + 1, len);
	for (const auto& it : aln) {
	    if (it.second == -1) continue;
	    const uint32_t code = uint32_t(coder[s[it.second]]);
	    int64_t cur = -1;
	    if (it.first == -1) {
		cur = add_node(code);
	    }
	    else {
		const uint32_t jt = uint32_t(it.first);
		if (nodes[jt].code == code) {
		    cur = jt;
		}
		else {
		    for (uint32_t kt : nodes[jt].aligned)
			if (nodes[kt].code == code) {
			    cur = kt;
			    break;
			}
		    if (cur < 0) {
			cur = add_node(code);
			const std::vector<uint32_t> al = nodes[jt].aligned;  // add_node may have moved the vector
			for (uint32_t kt : al) {
			    nodes[kt].aligned.push_back(uint32_t(cur));
			    nodes[size_t(cur)].aligned.push_back(kt);
			}
			nodes[jt].aligned.push_back(uint32_t(cur));
			nodes[size_t(cur)].aligned.push_back(jt);
		    }
		}
	    }
	    if (begin < 0) begin = cur;
	    if (prev >= 0) add_edge(uint32_t(prev), uint32_t(cur), w[uint32_t(it.second) - 1] + w[uint32_t(it.second)]);
	    prev = cur;
	}
	if (last >= 0) add_edge(uint32_t(prev), uint32_t(last), w[valid.back()] + w[valid.back() + 1]);
	sequences.push_back(uint32_t(begin));
	topological_sort();
    }

    void topological_sort()  // graph.cpp:249-300
    {
	rank_to_node.clear();
	const size_t n = nodes.size();
	std::vector<uint8_t> marks(n, 0), ignored(n, 0);
	std::vector<uint32_t> stack;
	for (uint32_t s = 0; s < n; s++) {
	    if (marks[s] != 0) continue;
	    stack.push_back(s);
	    while (!stack.empty()) {
		const uint32_t cur = stack.back();
		bool valid = true;
		if (marks[cur] != 2) {
		    for (uint32_t e : nodes[cur].in)
			if (marks[edges[e].tail] != 2) {
			    stack.push_back(edges[e].tail);
			    valid = false;
			}
		    if (!ignored[cur])
			for (uint32_t a : nodes[cur].aligned)
			    if (marks[a] != 2) {
				stack.push_back(a);
				ignored[a] = 1;
				valid = false;
			    }
		    if (valid) {
			marks[cur] = 2;
			if (!ignored[cur]) {
			    rank_to_node.push_back(cur);
			    for (uint32_t a : nodes[cur].aligned) rank_to_node.push_back(a);
			}
		    }
		    else {
			marks[cur] = 1;
		    }
		}
		if (valid) stack.pop_back();
	    }
	}
    }

    int64_t branch_completion(uint32_t rank, std::vector<int64_t>& scores, std::vector<int64_t>& pred)	// graph.cpp:488-528
    {
	const uint32_t start = rank_to_node[rank];
	for (uint32_t e : nodes[start].out)
	    for (uint32_t f : nodes[edges[e].head].in)
		if (edges[f].tail != start) scores[edges[f].tail] = -1;
	int64_t mx = -1;
	for (size_t i = rank + 1; i < rank_to_node.size(); i++) {
	    const uint32_t it = rank_to_node[i];
	    scores[it] = -1;
	    pred[it] = -1;
	    for (uint32_t e : nodes[it].in) {
		const Edge& jt = edges[e];
		if (scores[jt.tail] == -1) continue;
		if (scores[it] < jt.weight || (scores[it] == jt.weight && scores[size_t(pred[it])] <= scores[jt.tail])) {
		    scores[it] = jt.weight;
		    pred[it] = jt.tail;
		}
	    }
	    if (pred[it] >= 0) scores[it] += scores[size_t(pred[it])];
	    if (mx < 0 || scores[size_t(mx)] < scores[it]) mx = it;
	}
	return mx;
    }

    void traverse_heaviest_bundle()  // graph.cpp:441-486
    {
	if (rank_to_node.empty()) return;
	std::vector<int64_t> pred(nodes.size(), -1), scores(nodes.size(), -1);
	int64_t mx = -1;
	for (uint32_t it : rank_to_node) {
	    for (uint32_t e : nodes[it].in) {
		const Edge& jt = edges[e];
		if (scores[it] < jt.weight || (scores[it] == jt.weight && scores[size_t(pred[it])] <= scores[jt.tail])) {
		    scores[it] = jt.weight;
		    pred[it] = jt.tail;
		}
	    }
	    if (pred[it] >= 0) scores[it] += scores[size_t(pred[it])];
	    if (mx < 0 || scores[size_t(mx)] < scores[it]) mx = it;
	}
	if (!nodes[size_t(mx)].out.empty()) {
	    std::vector<uint32_t> rank_of(nodes.size(), 0);
	    for (uint32_t i = 0; i < rank_to_node.size(); i++) rank_of[rank_to_node[i]] = i;
	    while (!nodes[size_t(mx)].out.empty()) {
		mx = branch_completion(rank_of[size_t(mx)], scores, pred);
		if (mx < 0) throw std::runtime_error("branch completion left the graph");  // spoa dereferences a null here
	    }
	}
	consensus.clear();
	while (pred[size_t(mx)] >= 0) {
	    consensus.push_back(uint32_t(mx));
	    mx = pred[size_t(mx)];
	}
	consensus.push_back(uint32_t(mx));
	std::reverse(consensus.begin(), consensus.end());
    }

    std::string generate_consensus()  // graph.cpp:368-375
    {
	traverse_heaviest_bundle();
	std::string dst;
	for (uint32_t it : consensus) dst += char(decoder[nodes[it].code]);
	return dst;
    }

    // ---- graph.hpp:188-230 through cereal's binary archive (native little-endian; unique_ptr = one validity byte)
    void save(std::string& out) const
    {
	out.clear();
	auto put = [&](const void* p, size_t n) { out.append(reinterpret_cast<const char*>(p), n); };
	auto p32 = [&](uint32_t v) { put(&v, 4); };
	auto p64 = [&](uint64_t v) { put(&v, 8); };
	size_t nl = 0, na = 0;
	for (const auto& e : edges) nl += e.labels.size();
	for (const auto& n : nodes) na += n.aligned.size();
	out.reserve(4 + 2 * 1032 + 8 + nodes.size() * 9 + 8 + edges.size() * 17 + nl * 4 + 8 + sequences.size() * 4 + 8 + edges.size() * 8 + 8 +
		    na * 4 + 8 + rank_to_node.size() * 4 + 8 + consensus.size() * 4);
	p32(num_codes);
	p64(256);
	put(coder, 1024);
	p64(256);
	put(decoder, 1024);
	p64(nodes.size());
	for (uint32_t i = 0; i < nodes.size(); i++) {
	    out.push_back(char(1));
	    p32(i);
	    p32(nodes[i].code);
	}
	p64(edges.size());
	for (const auto& e : edges) {
	    out.push_back(char(1));
	    p64(e.labels.size());
	    if (!e.labels.empty()) put(e.labels.data(), e.labels.size() * 4);
	    p64(uint64_t(e.weight));
	}
	p64(sequences.size());
	if (!sequences.empty()) put(sequences.data(), sequences.size() * 4);
	p64(edges.size());
	for (const auto& e : edges) {
	    p32(e.tail);
	    p32(e.head);
	}
	uint64_t npairs = 0;
	for (uint32_t i = 0; i < nodes.size(); i++)
	    for (uint32_t a : nodes[i].aligned)
		if (i < a) npairs++;
	p64(npairs);
	for (uint32_t i = 0; i < nodes.size(); i++)
	    for (uint32_t a : nodes[i].aligned)
		if (i < a) {
		    p32(i);
		    p32(a);
		}
	p64(rank_to_node.size());
	if (!rank_to_node.empty()) put(rank_to_node.data(), rank_to_node.size() * 4);
	p64(consensus.size());
	if (!consensus.empty()) put(consensus.data(), consensus.size() * 4);
    }

    void load(const unsigned char* p, size_t n)	 // graph.hpp:232-276
    {
	clear();
	size_t o = 0;
	auto need = [&](size_t k) {
	    if (k > n - o) throw std::runtime_error("truncated graph");
	};
	auto get = [&](void* d, size_t k) {
	    need(k);
	    memcpy(d, p + o, k);
	    o += k;
	};
	auto g32 = [&]() {
	    uint32_t v;
	    get(&v, 4);
	    return v;
	};
	auto g64 = [&]() {
	    uint64_t v;
	    get(&v, 8);
	    return v;
	};
	auto count = [&](size_t elt) {
	    const uint64_t c = g64();
	    if (c > (n - o) / elt) throw std::runtime_error("truncated graph");
	    return size_t(c);
	};
	num_codes = g32();
	if (count(4) != 256) throw std::runtime_error("graph coder is not 256 entries");
	get(coder, 1024);
	if (count(4) != 256) throw std::runtime_error("graph decoder is not 256 entries");
	get(decoder, 1024);
	const size_t nn = count(1);
	nodes.resize(nn);
	for (size_t i = 0; i < nn; i++) {
	    uint8_t valid;
	    get(&valid, 1);
	    if (!valid) throw std::runtime_error("graph with a null node");
	    const uint32_t id = g32();
	    if (id != i) throw std::runtime_error("graph node ids are not 0..n-1");
	    nodes[i].code = g32();
	}
	const size_t ne = count(1);
	edges.resize(ne);
	for (size_t i = 0; i < ne; i++) {
	    uint8_t valid;
	    get(&valid, 1);
	    if (!valid) throw std::runtime_error("graph with a null edge");
	    const size_t nl = count(4);
	    edges[i].labels.resize(nl);
	    if (nl) get(edges[i].labels.data(), nl * 4);
	    edges[i].weight = int64_t(g64());
	}
	const size_t ns = count(4);
	sequences.resize(ns);
	if (ns) get(sequences.data(), ns * 4);
	const size_t nc = count(8);
	if (nc != ne) throw std::runtime_error("graph connections do not match its edges");
	for (size_t i = 0; i < nc; i++) {
	    edges[i].tail = g32();
	    edges[i].head = g32();
	    if (edges[i].tail >= nn || edges[i].head >= nn) throw std::runtime_error("graph edge leaves the node table");
	    nodes[edges[i].tail].out.push_back(uint32_t(i));
	    nodes[edges[i].head].in.push_back(uint32_t(i));
	}
	const size_t npairs = count(8);
	for (size_t i = 0; i < npairs; i++) {
	    const uint32_t a = g32(), b = g32();
	    if (a >= nn || b >= nn) throw std::runtime_error("graph aligned pair leaves the node table");
	    nodes[a].aligned.push_back(b);
	    nodes[b].aligned.push_back(a);
	}
	const size_t nr = count(4);
	rank_to_node.resize(nr);
	if (nr) get(rank_to_node.data(), nr * 4);
	const size_t ncs = count(4);
	consensus.resize(ncs);
	if (ncs) get(consensus.data(), ncs * 4);
	for (uint32_t v : sequences)
	    if (v >= nn) throw std::runtime_error("graph sequence leaves the node table");
	for (uint32_t v : rank_to_node)
	    if (v >= nn) throw std::runtime_error("graph rank leaves the node table");
	for (uint32_t v : consensus)
	    if (v >= nn) throw std::runtime_error("graph consensus leaves the node table");
	if (o != n) throw std::runtime_error("bytes behind the graph");
	// the alignment walks rows in rank order and reads the rows of a node's predecessors: the stored order must be a
	// permutation of the nodes with every edge pointing forward (spoa trusts the file; a damaged one must not index outside
	// the matrices here)
	if (nr != nn) throw std::runtime_error("graph ranks are not a permutation of its nodes");
	std::vector<uint32_t> rank_of(nn, 0xffffffffu);
	for (size_t i = 0; i < nr; i++) {
	    if (rank_of[rank_to_node[i]] != 0xffffffffu) throw std::runtime_error("graph ranks are not a permutation of its nodes");
	    rank_of[rank_to_node[i]] = uint32_t(i);
	}
	for (const auto& e : edges)
	    if (rank_of[e.tail] >= rank_of[e.head]) throw std::runtime_error("graph edge points backwards in the stored order");
	for (uint32_t c = 0; c < num_codes && c < 256; c++)
	    if (decoder[c] < 0 || decoder[c] > 255) throw std::runtime_error("graph decoder entry outside a byte");
	if (num_codes > 256) throw std::runtime_error("graph with more than 256 codes");
	for (const auto& nd : nodes)
	    if (nd.code >= num_codes) throw std::runtime_error("graph node with an unknown code");
    }
};

struct PoaParams {
    int type;  // 0 kSW (local), 1 kNW (global), 2 kOV (semi-global): main.cpp:292-303
    int m, n, g, e, q, c;
};

// AlignmentEngine::WorstCaseAlignmentScore (alignment_engine.cpp:100-110)
int64_t worst_case_score(const PoaParams& P, int64_t i, int64_t j)
{
    auto gap = [&](int64_t len) -> int64_t { return len == 0 ? 0 : std::min<int64_t>(P.g + (len - 1) * P.e, P.q + (len - 1) * P.c); };
    return std::min(-1 * (P.m * std::min(i, j) + gap(std::llabs(i - j))), gap(i) + gap(j));
}

// N lanes of T as one GCC vector; up<SH>(v, fill): lane k takes v[k - SH], the lanes below SH take fill[0]
template <typename T, int N>
struct PoaVec {
    typedef T V __attribute__((vector_size(sizeof(T) * N)));
    typedef typename std::make_unsigned<T>::type U;
    typedef U VU __attribute__((vector_size(sizeof(T) * N)));
    static inline __attribute__((always_inline)) V add(V x, V y) { return (V)((VU)x + (VU)y); }  // wraps like paddw / paddd
    static inline __attribute__((always_inline)) V sub(V x, V y) { return (V)((VU)x - (VU)y); }
    static inline __attribute__((always_inline)) V max(V x, V y) { return x > y ? x : y; }
    static inline __attribute__((always_inline)) V splat(T x) { return x - V{}; }
    template <int SH, size_t... K>
    static inline __attribute__((always_inline)) V up_impl(V v, V fill, std::index_sequence<K...>)
    {
	return __builtin_shufflevector(v, fill, (int(K) >= SH ? int(K) - SH : N)...);
    }
    template <int SH>
    static inline __attribute__((always_inline)) V up(V v, V fill)
    {
	return up_impl<SH>(v, fill, std::make_index_sequence<N>{});
    }
};

// work arrays kept between calls (SimdAlignmentEngine::Realloc, :424-516); every cell read by a call is written by it first
template <typename T>
struct PoaWork {
    std::vector<T> H, F, O, E, Q, profile;
    std::vector<int32_t> fc, fcE, fcQ;
    std::vector<uint32_t> rank_of;
};

template <typename T, int N, int LOGN>
__attribute__((always_inline)) inline PoaAlignment poa_align(const unsigned char* seq, uint32_t len, const PoaGraph& G, const PoaParams& P, PoaWork<T>& Wk, int32_t* score_out)
{
    const T kNegInf = T(std::numeric_limits<T>::min() + 1024);
    const size_t W = (size_t(len) + N - 1) / N;	 // matrix_width in vectors
    const size_t S = W * N;			 // lanes per row
    const size_t rows = G.nodes.size() + 1;
    const auto& rank_to_node = G.rank_to_node;
    typedef typename std::make_unsigned<T>::type U;
    auto wrap = [](int64_t v) { return T(v); };				       // lane arithmetic wraps (paddw / paddd)
    auto add = [](T a, T b) -> T { return T(U(U(a) + U(b))); };		       // the same in the lane's own width

    if (Wk.H.size() < rows * S) {
	const size_t want = rows * S + rows * S / 4;
	Wk.H.resize(want);
	Wk.F.resize(want);
	Wk.O.resize(want);
	Wk.E.resize(want);
	Wk.Q.resize(want);
    }
    if (Wk.profile.size() < size_t(G.num_codes) * S) Wk.profile.resize(size_t(G.num_codes) * S);
    Wk.fc.assign(rows, 0);
    Wk.fcE.assign(rows, 0);
    Wk.fcQ.assign(rows, 0);
    Wk.rank_of.assign(G.nodes.size(), 0);
    T* H = Wk.H.data();
    T* F = Wk.F.data();
    T* O = Wk.O.data();
    T* E = Wk.E.data();
    T* Q = Wk.Q.data();

    // ---- Initialize (:518-700)
    const int padding = -1 * std::max(std::max(std::abs(P.m), std::abs(P.n)), std::max(std::abs(P.g), std::abs(P.q)));
    for (uint32_t c = 0; c < G.num_codes; c++) {
	const int ch = G.decoder[c];
	T* pr = Wk.profile.data() + size_t(c) * S;
	for (size_t j = 0; j < S; j++) pr[j] = j < len ? T(ch == int(seq[j]) ? P.m : P.n) : T(padding);
    }
    for (uint32_t i = 0; i < rank_to_node.size(); i++) Wk.rank_of[rank_to_node[i]] = i;
    for (size_t j = 0; j < S; j++) {
	O[j] = kNegInf;
	F[j] = kNegInf;
	Q[j] = wrap(int64_t(P.q) + int64_t(j) * P.c);
	E[j] = wrap(int64_t(P.g) + int64_t(j) * P.e);
    }
    auto first_column_gap = [&](std::vector<int32_t>& fcg, int open, int ext) {
	fcg[0] = 0;
	for (size_t i = 1; i < rows; i++) {
	    const auto& in = G.nodes[rank_to_node[i - 1]].in;
	    int32_t penalty = in.empty() ? open - ext : int32_t(kNegInf);
	    for (uint32_t e : in) penalty = std::max(penalty, fcg[Wk.rank_of[G.edges[e].tail] + 1]);
	    fcg[i] = penalty + ext;
	}
    };
    first_column_gap(Wk.fcQ, P.q, P.c);
    first_column_gap(Wk.fcE, P.g, P.e);
    if (P.type == 0) {
	for (size_t j = 0; j < S; j++) H[j] = 0;
    }
    else {
	for (size_t j = 0; j < S; j++) H[j] = std::max(E[j], Q[j]);
	if (P.type == 1)
	    for (size_t i = 0; i < rows; i++) Wk.fc[i] = std::max(Wk.fcE[i], Wk.fcQ[i]);
    }
    const int32_t* fc = Wk.fc.data();

    // ---- the matrix, row by row in topological order (:1640-1790)
    T max_score = P.type == 0 ? T(0) : kNegInf;
    int32_t max_i = -1, max_j = -1;
    const T gme = T(P.g - P.e), qmc = T(P.q - P.c), ve = T(P.e), vc = T(P.c);
    T pen_e[LOGN], pen_c[LOGN];
    pen_e[0] = ve;
    pen_c[0] = vc;
    for (int s = 1; s < LOGN; s++) {
	pen_e[s] = wrap(int64_t(pen_e[s - 1]) + pen_e[s - 1]);
	pen_c[s] = wrap(int64_t(pen_c[s - 1]) + pen_c[s - 1]);
    }
    std::vector<uint32_t> preds;
    for (uint32_t rk = 0; rk < rank_to_node.size(); rk++) {
	const PoaGraph::Node& node = G.nodes[rank_to_node[rk]];
	const T* pr = Wk.profile.data() + size_t(node.code) * S;
	const size_t i = size_t(rk) + 1;
	T* Hr = H + i * S;
	T* Fr = F + i * S;
	T* Or = O + i * S;
	T* Er = E + i * S;
	T* Qr = Q + i * S;
	preds.clear();
	if (node.in.empty()) preds.push_back(0);
	for (uint32_t e : node.in) preds.push_back(Wk.rank_of[G.edges[e].tail] + 1);
	for (size_t p = 0; p < preds.size(); p++) {
	    const T* __restrict Hp = H + size_t(preds[p]) * S;
	    const T* __restrict Fp = F + size_t(preds[p]) * S;
	    const T* __restrict Op = O + size_t(preds[p]) * S;
	    T* __restrict Hw = Hr;
	    T* __restrict Fw = Fr;
	    T* __restrict Ow = Or;
	    const T x0 = T(fc[preds[p]]);
	    if (p == 0) {
		for (size_t j = 0; j < S; j++) {
		    Fw[j] = add(std::max(add(Hp[j], gme), Fp[j]), ve);
		    Ow[j] = add(std::max(add(Hp[j], qmc), Op[j]), vc);
		}
		Hw[0] = add(x0, pr[0]);
		for (size_t j = 1; j < S; j++) Hw[j] = add(Hp[j - 1], pr[j]);
	    }
	    else {
		for (size_t j = 0; j < S; j++) {
		    Fw[j] = std::max(Fw[j], add(std::max(add(Hp[j], gme), Fp[j]), ve));
		    Ow[j] = std::max(Ow[j], add(std::max(add(Hp[j], qmc), Op[j]), vc));
		}
		Hw[0] = std::max(Hw[0], add(x0, pr[0]));
		for (size_t j = 1; j < S; j++) Hw[j] = std::max(Hw[j], add(Hp[j - 1], pr[j]));
	    }
	}
	// One register of the reference = one vector here (PoaVec below: GCC vector extensions with constant shuffles, the
	// compiler picks the instructions — 2 x SSE2 for a plain x86-64 build).
	typedef PoaVec<T, N> PV;
	typedef typename PV::V V;
	const V vninf = PV::splat(kNegInf), vgme = PV::splat(gme), vqmc = PV::splat(qmc), vve = PV::splat(ve), vvc = PV::splat(vc), vzero = V{};
	V vpe[LOGN], vpc[LOGN];
	for (int sidx = 0; sidx < LOGN; sidx++) {
	    vpe[sidx] = PV::splat(pen_e[sidx]);
	    vpc[sidx] = PV::splat(pen_c[sidx]);
	}
	T xl = T(fc[i]), yl = T(fc[i]);
	V best = vzero;
	for (size_t b = 0; b < W; b++) {
	    V h, fb, ob;
	    memcpy(&h, Hr + b * N, sizeof h);
	    memcpy(&fb, Fr + b * N, sizeof fb);
	    memcpy(&ob, Or + b * N, sizeof ob);
	    h = PV::max(h, PV::max(fb, ob));
	    V a = PV::add(PV::add(PV::template up<1>(h, PV::splat(xl)), vgme), vve);
	    V q = PV::add(PV::add(PV::template up<1>(h, PV::splat(yl)), vqmc), vvc);
	    // _mmxxx_prefix_max (:107-114, :142-148): lanes move up by 1, 2, 4 (, 8), the vacated ones read -infinity
	    a = PV::max(a, PV::template up<1>(PV::add(a, vpe[0]), vninf));
	    q = PV::max(q, PV::template up<1>(PV::add(q, vpc[0]), vninf));
	    a = PV::max(a, PV::template up<2>(PV::add(a, vpe[1]), vninf));
	    q = PV::max(q, PV::template up<2>(PV::add(q, vpc[1]), vninf));
	    a = PV::max(a, PV::template up<4>(PV::add(a, vpe[2]), vninf));
	    q = PV::max(q, PV::template up<4>(PV::add(q, vpc[2]), vninf));
	    if (LOGN > 3) {
		a = PV::max(a, PV::template up<(N > 8 ? 8 : 1)>(PV::add(a, vpe[LOGN - 1]), vninf));
		q = PV::max(q, PV::template up<(N > 8 ? 8 : 1)>(PV::add(q, vpc[LOGN - 1]), vninf));
	    }
	    h = PV::max(h, PV::max(a, q));
	    const V xs = PV::max(h, PV::sub(a, vgme)), ys = PV::max(h, PV::sub(q, vqmc));
	    xl = xs[N - 1];
	    yl = ys[N - 1];
	    if (P.type == 0) h = PV::max(h, vzero);
	    best = PV::max(best, h);
	    memcpy(Hr + b * N, &h, sizeof h);
	    memcpy(Er + b * N, &a, sizeof a);
	    memcpy(Qr + b * N, &q, sizeof q);
	}
	if (P.type == 0 || (P.type == 2 && node.out.empty())) {
	    T row_max = 0;  // _mmxxx_max_value starts at 0 (:254)
	    for (int k = 0; k < N; k++) row_max = std::max(row_max, T(best[k]));
	    if (max_score < row_max) {
		max_score = row_max;
		max_i = int32_t(i);
	    }
	}
	else if (P.type == 1 && node.out.empty()) {
	    const T v = Hr[len - 1];
	    if (max_score < v) {
		max_score = v;
		max_i = int32_t(i);
	    }
	}
    }
    if (max_i == -1) return PoaAlignment();
    if (score_out) *score_out = max_score;
    if (P.type == 1) {
	max_j = int32_t(len) - 1;
    }
    else {
	const T* Hr = H + size_t(max_i) * S;
	for (size_t j = 0; j < S; j++)
	    if (Hr[j] == max_score) {
		max_j = int32_t(j);
		break;
	    }
	// A best row without a single non-negative cell: its maximum was taken as 0 (above), no cell holds it, the reference's
	// column search returns -1 and its backtrack stops before the first step — an empty alignment, and AddAlignment then
	// adds the sequence as an unconnected chain.
	if (max_j < 0) return PoaAlignment();
    }

    // ---- backtrack (:1793-2090); cells are read where the reference's lane buffers were loaded from
    PoaAlignment aln;
    int32_t i = max_i, j = max_j, prev_i = 0, prev_j = 0;
    auto cell = [&](const T* M, int32_t r, int32_t col) -> int { return int(M[size_t(r) * S + size_t(col)]); };
    while (true) {
	if (j == -1 || i == 0) break;
	const PoaGraph::Node& node = G.nodes[rank_to_node[size_t(i) - 1]];
	preds.clear();
	if (node.in.empty()) preds.push_back(0);
	for (uint32_t e : node.in) preds.push_back(Wk.rank_of[G.edges[e].tail] + 1);
	const int Hc = cell(H, i, j);
	if (P.type == 0 && Hc == 0) break;
	const int prof = int(Wk.profile[size_t(node.code) * S + size_t(j)]);
	bool found = false, extend_left = false, extend_up = false;
	for (uint32_t p : preds) {
	    const int hd = j == 0 ? int(T(fc[p])) : cell(H, int32_t(p), j - 1);
	    if (Hc == hd + prof) {
		prev_i = int32_t(p);
		prev_j = j - 1;
		found = true;
		break;
	    }
	}
	if (!found) {
	    for (uint32_t p : preds) {
		const int32_t pi = int32_t(p);
		if ((extend_up = (Hc == cell(F, pi, j) + P.e)) || Hc == cell(H, pi, j) + P.g || (extend_up = (Hc == cell(O, pi, j) + P.c)) ||
		    Hc == cell(H, pi, j) + P.q) {
		    prev_i = pi;
		    prev_j = j;
		    found = true;
		    break;
		}
	    }
	}
	if (!found) {
	    const int El = j == 0 ? int(T(fc[i])) : cell(E, i, j - 1);
	    const int Hl = j == 0 ? int(T(fc[i])) : cell(H, i, j - 1);
	    const int Ql = j == 0 ? int(T(fc[i])) : cell(Q, i, j - 1);
	    if ((extend_left = (Hc == El + P.e)) || Hc == Hl + P.g || (extend_left = (Hc == Ql + P.c)) || Hc == Hl + P.q) {
		prev_i = i;
		prev_j = j - 1;
		found = true;
	    }
	}
	aln.emplace_back(i == prev_i ? -1 : int32_t(rank_to_node[size_t(i) - 1]), j == prev_j ? -1 : j);
	i = prev_i;
	j = prev_j;
	if (extend_left) {
	    if (j < 0) break;  // the reference reads outside its matrices here
	    while (true) {
		aln.emplace_back(-1, j);
		--j;
		if (j == -1 || cell(E, i, j) + P.e != cell(E, i, j + 1) || cell(Q, i, j) + P.c != cell(Q, i, j + 1)) break;
	    }
	}
	else if (extend_up) {
	    if (i == 0) break;	// same remark
	    while (true) {
		const PoaGraph::Node& nd = G.nodes[rank_to_node[size_t(i) - 1]];
		bool stop = true;
		prev_i = 0;
		for (uint32_t e : nd.in) {
		    const int32_t pi = int32_t(Wk.rank_of[G.edges[e].tail] + 1);
		    if (cell(F, i, j) == cell(F, pi, j) + P.e || cell(O, i, j) == cell(O, pi, j) + P.c) {
			prev_i = pi;
			stop = false;
			break;
		    }
		}
		if (stop)
		    for (uint32_t e : nd.in) {
			const int32_t pi = int32_t(Wk.rank_of[G.edges[e].tail] + 1);
			if (cell(F, i, j) == cell(H, pi, j) + P.g || cell(O, i, j) == cell(H, pi, j) + P.q) {
			    prev_i = pi;
			    break;
			}
		    }
		aln.emplace_back(int32_t(rank_to_node[size_t(i) - 1]), -1);
		i = prev_i;
		if (stop || i == 0) break;
	    }
	}
    }
    if (P.type == 1) {	// :2093-2115
	while (i == 0 && j != -1) {
	    aln.emplace_back(-1, j);
	    --j;
	}
	while (i != 0 && j == -1) {
	    aln.emplace_back(int32_t(rank_to_node[size_t(i) - 1]), -1);
	    const auto& in = G.nodes[rank_to_node[size_t(i) - 1]].in;
	    if (in.empty()) {
		i = 0;
	    }
	    else {
		// the reference leaves i unchanged (and never returns) when no predecessor continues either gap chain
		bool moved = false;
		for (uint32_t e : in) {
		    const int32_t pi = int32_t(Wk.rank_of[G.edges[e].tail] + 1);
		    if (Wk.fcE[size_t(i)] == Wk.fcE[size_t(pi)] + P.e || Wk.fcQ[size_t(i)] == Wk.fcQ[size_t(pi)] + P.c) {
			i = pi;
			moved = true;
			break;
		    }
		}
		if (!moved) throw std::runtime_error("global alignment: the first column has no predecessor");
	    }
	}
    }
    std::reverse(aln.begin(), aln.end());
    return aln;
}

}  // namespace

struct icb_poa_graph {
    PoaGraph g;
    PoaWork<int16_t> w16;
    PoaWork<int32_t> w32;
    std::string scratch;
};

namespace {

// SimdAlignmentEngine::Align (:703-780)
// The engine twice: compiled for AVX2 and for the build's baseline, picked when the library is loaded (GCC function
// multi-versioning) — the lane shuffles of the block pass are one instruction each with AVX2 and a dozen with SSE2.
__attribute__((target_clones("avx2", "default"))) PoaAlignment poa_align_i16(const unsigned char* seq, uint32_t len, const PoaGraph& G, const PoaParams& P, PoaWork<int16_t>& Wk, int32_t* score)
{
    return poa_align<int16_t, 16, 4>(seq, len, G, P, Wk, score);
}
__attribute__((target_clones("avx2", "default"))) PoaAlignment poa_align_i32(const unsigned char* seq, uint32_t len, const PoaGraph& G, const PoaParams& P, PoaWork<int32_t>& Wk, int32_t* score)
{
    return poa_align<int32_t, 8, 3>(seq, len, G, P, Wk, score);
}

PoaAlignment poa_align_any(icb_poa_graph* h, const unsigned char* seq, uint32_t len, const PoaParams& P, int32_t* score)
{
    if (h->g.nodes.empty() || len == 0) return PoaAlignment();
    const int64_t worst = worst_case_score(P, int64_t(len) + 8, int64_t(h->g.nodes.size()));
    if (worst < int64_t(std::numeric_limits<int32_t>::min()) + 1024) throw std::invalid_argument("possible overflow");
    if (worst < int64_t(std::numeric_limits<int16_t>::min()) + 1024) return poa_align_i32(seq, len, h->g, P, h->w32, score);
    return poa_align_i16(seq, len, h->g, P, h->w16, score);
}

bool poa_params(int algo, PoaParams& P)
{
    if (algo < 0 || algo > 2) return false;
    P.type = algo;
    P.m = 4;  // main.cpp:284-289
    P.n = -8;
    P.g = -8;
    P.e = -4;
    P.q = -20;
    P.c = -1;
    return true;
}

}  // namespace

#define ICB_POA_GUARD(body)                 \
    try {                                   \
	body                                \
    }                                       \
    catch (const std::bad_alloc&) {         \
	return ICB_ERR_NOMEM;               \
    }                                       \
    catch (const std::invalid_argument&) {  \
	return ICB_ERR_ARG;                 \
    }                                       \
    catch (...) {                           \
	return ICB_ERR_STATE;               \
    }

extern "C" {

int icb_poa_graph_create(icb_poa_graph** graph)
{
    if (!graph) return ICB_ERR_ARG;
    ICB_POA_GUARD(*graph = new icb_poa_graph(); return ICB_OK;)
}

void icb_poa_graph_destroy(icb_poa_graph* graph) { delete graph; }

int icb_poa_graph_load(const void* bytes, int64_t n, icb_poa_graph** graph)
{
    if (!graph || !bytes || n < 0) return ICB_ERR_ARG;
    icb_poa_graph* h = nullptr;
    try {
	h = new icb_poa_graph();
	h->g.load(reinterpret_cast<const unsigned char*>(bytes), size_t(n));
	*graph = h;
	return ICB_OK;
    }
    catch (const std::bad_alloc&) {
	delete h;
	return ICB_ERR_NOMEM;
    }
    catch (...) {
	delete h;
	return ICB_ERR_ARG;
    }
}

int icb_poa_graph_save(icb_poa_graph* graph, void* out, int64_t cap, int64_t* n)
{
    if (!graph || !n) return ICB_ERR_ARG;
    ICB_POA_GUARD(graph->g.save(graph->scratch); *n = int64_t(graph->scratch.size());
		  if (out && cap >= *n) memcpy(out, graph->scratch.data(), graph->scratch.size()); return ICB_OK;)
}

int icb_poa_graph_sizes(const icb_poa_graph* graph, int64_t* n_sequences, int64_t* n_nodes, int64_t* n_edges)
{
    if (!graph) return ICB_ERR_ARG;
    if (n_sequences) *n_sequences = int64_t(graph->g.sequences.size());
    if (n_nodes) *n_nodes = int64_t(graph->g.nodes.size());
    if (n_edges) *n_edges = int64_t(graph->g.edges.size());
    return ICB_OK;
}

int icb_poa_graph_clear(icb_poa_graph* graph)
{
    if (!graph) return ICB_ERR_ARG;
    graph->g.clear();
    return ICB_OK;
}

int icb_poa_align(icb_poa_graph* graph, const char* seq, int64_t len, int algo, int32_t* node_ids, int32_t* positions, int64_t cap,
		  int64_t* n_out, int32_t* score)
{
    PoaParams P;
    if (!graph || !seq || len < 0 || len > 0x7fffffff || !n_out || !poa_params(algo, P)) return ICB_ERR_ARG;
    ICB_POA_GUARD(const PoaAlignment a = poa_align_any(graph, reinterpret_cast<const unsigned char*>(seq), uint32_t(len), P, score);
		  *n_out = int64_t(a.size()); if (node_ids && positions && cap >= *n_out) for (size_t i = 0; i < a.size(); i++) {
		      node_ids[i] = a[i].first;
		      positions[i] = a[i].second;
		  } return ICB_OK;)
}

int icb_poa_graph_add(icb_poa_graph* graph, const char* seq, int64_t len, uint32_t weight, int algo)
{
    PoaParams P;
    if (!graph || !seq || len < 0 || len > 0x7fffffff || !poa_params(algo, P)) return ICB_ERR_ARG;
    ICB_POA_GUARD(const unsigned char* s = reinterpret_cast<const unsigned char*>(seq);
		  const PoaAlignment a = poa_align_any(graph, s, uint32_t(len), P, nullptr); graph->g.add_alignment(a, s, uint32_t(len), weight);
		  return ICB_OK;)
}

int icb_poa_graph_consensus(icb_poa_graph* graph, char* out, int64_t cap, int64_t* len)
{
    if (!graph || !len) return ICB_ERR_ARG;
    ICB_POA_GUARD(const std::string c = graph->g.generate_consensus(); *len = int64_t(c.size());
		  if (out && cap >= *len) memcpy(out, c.data(), c.size()); return ICB_OK;)
}

}  // extern "C"
