// cer.cpp — see cer.hpp.  Pure host code; built with -ffp-contract=off like every host file that touches doubles
// (SURVEY App. C#13).
#include "cer.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <set>

namespace icb {

bool same_sort_args(const CmdArgs& a, const CmdArgs& b)
{
    return a.Verbose == b.Verbose && a.Debug == b.Debug && a.KmerSize == b.KmerSize && a.BatchSize == b.BatchSize &&
	   a.WindowSize == b.WindowSize && a.MinShared == b.MinShared && a.MinQual == b.MinQual &&
	   a.MappedThreshold == b.MappedThreshold && a.MinFraction == b.MinFraction;
}

// ---------------------------------------------------------------- Writer / Reader
Writer::Writer(const std::string& path) : f_(fopen(path.c_str(), "wb")), buf_(nullptr), len_(0), cap_(size_t(8) << 20), total_(0), path_(path)
{
    if (!f_) throw std::runtime_error("Failed to open " + path + "!");
    buf_ = static_cast<char*>(malloc(cap_));
}
Writer::~Writer()
{
    if (f_) {
	try {
	    close();
	}
	catch (...) {
	}
    }
    free(buf_);
}
void Writer::put(const void* p, size_t n)
{
    if (n && fwrite(p, 1, n, f_) != n) throw std::runtime_error("write failed: " + path_);
    total_ += n;
}
void Writer::flush()
{
    put(buf_, len_);
    len_ = 0;
}
void Writer::close()
{
    if (!f_) return;
    flush();
    FILE* f = f_;
    f_ = nullptr;
    if (fclose(f) != 0) throw std::runtime_error("close failed: " + path_);
}

Reader::Reader(const std::string& path) : o_(0)
{
    FILE* f = fopen(path.c_str(), "rb");
    // the reference opens an ifstream and lets cereal fail on the first read
    if (!f) throw std::runtime_error("Failed to read 4 bytes from input stream! Read 0");
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    d_.resize(n > 0 ? size_t(n) : 0);
    if (n > 0 && fread(&d_[0], 1, size_t(n), f) != size_t(n)) {
	fclose(f);
	throw std::runtime_error("read failed: " + path);
    }
    fclose(f);
}
std::string Reader::str()
{
    uint64_t n = u64();
    need(n);
    std::string s(d_.data() + o_, size_t(n));
    o_ += n;
    return s;
}

// ---------------------------------------------------------------- pieces
void write_args(Writer& w, const CmdArgs& a)
{
    w.u8(a.Verbose);
    w.u8(a.Debug);
    w.str(a.InFastq);
    w.i32(a.KmerSize);
    w.i32(a.BatchSize);
    w.i32(a.BatchMaxSeq);
    w.i32(a.WindowSize);
    w.i32(a.MinShared);
    w.i32(a.ConsMinSize);
    w.i32(a.ConsMaxSize);
    w.i32(a.ConsPeriod);
    w.i32(a.MinClsSize);
    w.f64(a.MinQual);
    w.f64(a.MappedThreshold);
    w.f64(a.AlignedThreshold);
    w.f64(a.MinFraction);
    w.f64(a.MinProbNoHits);
    w.str(a.BatchOutFolder);
    w.i32(a.Mode);
}
void read_args(Reader& r, CmdArgs& a)
{
    a.Verbose = r.u8() != 0;
    a.Debug = r.u8() != 0;
    a.InFastq = r.str();
    a.KmerSize = r.i32();
    a.BatchSize = r.i32();
    a.BatchMaxSeq = r.i32();
    a.WindowSize = r.i32();
    a.MinShared = r.i32();
    a.ConsMinSize = r.i32();
    a.ConsMaxSize = r.i32();
    a.ConsPeriod = r.i32();
    a.MinClsSize = r.i32();
    a.MinQual = r.f64();
    a.MappedThreshold = r.f64();
    a.AlignedThreshold = r.f64();
    a.MinFraction = r.f64();
    a.MinProbNoHits = r.f64();
    a.BatchOutFolder = r.str();
    a.Mode = r.i32();
}

static inline void write_seq(Writer& w, const SeqRec& s)
{
    w.str(s.name);
    w.str(s.seq);
    w.str(s.qual);
    w.f64(s.score);
    w.f64(s.err);
}
static inline void write_useq(Writer& w, const std::unique_ptr<SeqRec>& s)
{
    if (!s) {
	w.u8(0);
	return;
    }
    w.u8(1);
    write_seq(w, *s);
}
static inline void write_mins(Writer& w, const std::vector<Minimizer>& m)
{
    w.u64(m.size());
    w.raw(m.data(), m.size() * sizeof(Minimizer));
}
static std::unique_ptr<SeqRec> read_useq(Reader& r)
{
    if (!r.u8()) return nullptr;
    std::unique_ptr<SeqRec> s(new SeqRec);
    s->name = r.str();
    s->seq = r.str();
    s->qual = r.str();
    s->score = r.f64();
    s->err = r.f64();
    return s;
}
static void read_mins(Reader& r, std::vector<Minimizer>& m)
{
    uint64_t n = r.u64();
    if (n > (r.size() - r.pos()) / sizeof(Minimizer)) throw std::runtime_error("Failed to read minimizers: truncated input");
    m.resize(size_t(n));
    r.raw(m.data(), size_t(n) * sizeof(Minimizer));
}

// cereal's shared_ptr protocol (cereal/types/memory.hpp, cereal.hpp registerSharedPointer): u32 id, 0 = null; the first
// occurrence of an address has the MSB set and is followed by the object; ids count from 1 over ALL shared_ptr types
struct PtrIds {
    std::unordered_map<const void*, uint32_t> seen;
    uint32_t next = 1;
    // returns true when the object has to follow
    bool put(Writer& w, const void* p)
    {
	if (!p) {
	    w.u32(0);
	    return false;
	}
	auto it = seen.find(p);
	if (it != seen.end()) {
	    w.u32(it->second);
	    return false;
	}
	uint32_t id = next++;
	seen.emplace(p, id);
	w.u32(id | 0x80000000u);
	return true;
    }
};

static void write_header(Writer& w, const Batch& b)
{
    w.i32(b.BatchNr);
    w.u64(b.BatchStart);
    w.u64(b.BatchEnd);
    w.u64(b.BatchBases);
    w.i32(b.TotalReads);
    w.i32(b.NrCls);
    write_args(w, b.SortArgs);
    w.str(b.LeftLeaf);
    w.str(b.RightLeaf);
    w.i32(b.Depth);
}

void save_batch(const Batch& b, const std::string& path)
{
    Writer w(path);
    write_header(w, b);
    w.u64(b.MinDB.size());
    for (const auto& kv : b.MinDB) {  // container iteration order
	w.u32(kv.first);
	w.u64(kv.second.size());
	w.raw(kv.second.data(), kv.second.size() * sizeof(unsigned));
    }
    PtrIds ids;
    w.u64(b.Cls.size());
    for (const auto& c : b.Cls) {
	if (!ids.put(w, c.get())) continue;
	w.u64(c->size());
	for (const auto& p : *c) {
	    if (!ids.put(w, p.get())) continue;
	    write_useq(w, p->raw);
	    write_useq(w, p->hpc);
	    write_mins(w, p->mins);
	    write_mins(w, p->rev);
	    w.i32(p->strand);
	    w.str(p->id);
	}
    }
    w.u64(b.ConsGs.size());
    for (size_t i = 0; i < b.ConsGs.size(); i++) {
	if (!b.ConsGsValid[i]) {
	    w.u8(0);
	    continue;
	}
	w.u8(1);
	w.raw(b.ConsGs[i].data(), b.ConsGs[i].size());
    }
    w.close();
}

// length of one serialized spoa::Graph starting at the reader's position (graph.hpp:219-229)
static void skip_graph(Reader& r)
{
    auto arith_vec = [&](size_t elt) {
	uint64_t n = r.u64();
	if (n > (r.size() - r.pos()) / elt) throw std::runtime_error("Failed to read graph: truncated input");
	r.skip(size_t(n) * elt);
    };
    r.u32();	   // num_codes_
    arith_vec(4);  // coder_
    arith_vec(4);  // decoder_
    uint64_t nn = r.u64();
    for (uint64_t i = 0; i < nn; i++)
	if (r.u8()) r.skip(8);	// Node{id, code}
    uint64_t ne = r.u64();
    for (uint64_t i = 0; i < ne; i++)
	if (r.u8()) {
	    arith_vec(4);  // labels
	    r.skip(8);	   // weight
	}
    arith_vec(4);  // sequences
    arith_vec(8);  // connections
    arith_vec(8);  // aligned_nodes
    arith_vec(4);  // rank_to_node_id
    arith_vec(4);  // consensus
}

std::unique_ptr<Batch> load_batch(const std::string& path)
{
    Reader r(path);
    std::unique_ptr<Batch> b(new Batch);
    b->BatchNr = r.i32();
    b->BatchStart = r.u64();
    b->BatchEnd = r.u64();
    b->BatchBases = r.u64();
    b->TotalReads = r.i32();
    b->NrCls = r.i32();
    read_args(r, b->SortArgs);
    b->LeftLeaf = r.str();
    b->RightLeaf = r.str();
    b->Depth = r.i32();
    {  // cereal/types/concepts/pair_associative_container.hpp:49-71 on a default-constructed map
	uint64_t n = r.u64();
	b->MinDB.clear();
	auto hint = b->MinDB.begin();
	for (uint64_t i = 0; i < n; i++) {
	    unsigned key = r.u32();
	    uint64_t m = r.u64();
	    if (m > (r.size() - r.pos()) / 4) throw std::runtime_error("Failed to read postings: truncated input");
	    std::vector<unsigned> v(static_cast<size_t>(m));
	    r.raw(v.data(), size_t(m) * 4);
	    hint = b->MinDB.emplace_hint(hint, std::move(key), std::move(v));
	}
    }
    std::unordered_map<uint32_t, std::shared_ptr<Cluster>> cls_by_id;
    std::unordered_map<uint32_t, std::shared_ptr<ProcSeq>> seq_by_id;
    uint64_t nc = r.u64();
    b->Cls.reserve(size_t(nc));
    for (uint64_t i = 0; i < nc; i++) {
	uint32_t id = r.u32();
	if (id == 0) {
	    b->Cls.emplace_back(nullptr);
	    continue;
	}
	if (!(id & 0x80000000u)) {
	    auto it = cls_by_id.find(id);
	    if (it == cls_by_id.end()) throw std::runtime_error("Error while trying to deserialize a smart pointer. Could not find id " + std::to_string(id));
	    b->Cls.emplace_back(it->second);
	    continue;
	}
	auto c = std::make_shared<Cluster>();
	cls_by_id.emplace(id & 0x7fffffffu, c);
	uint64_t m = r.u64();
	c->reserve(size_t(m));
	for (uint64_t j = 0; j < m; j++) {
	    uint32_t pid = r.u32();
	    if (pid == 0) {
		c->emplace_back(nullptr);
		continue;
	    }
	    if (!(pid & 0x80000000u)) {
		auto it = seq_by_id.find(pid);
		if (it == seq_by_id.end()) throw std::runtime_error("Error while trying to deserialize a smart pointer. Could not find id " + std::to_string(pid));
		c->emplace_back(it->second);
		continue;
	    }
	    auto p = std::make_shared<ProcSeq>();
	    seq_by_id.emplace(pid & 0x7fffffffu, p);
	    p->raw = read_useq(r);
	    p->hpc = read_useq(r);
	    read_mins(r, p->mins);
	    read_mins(r, p->rev);
	    p->strand = r.i32();
	    p->id = r.str();
	    c->emplace_back(std::move(p));
	}
	b->Cls.emplace_back(std::move(c));
    }
    uint64_t ng = r.u64();
    b->ConsGs.resize(size_t(ng));
    b->ConsGsValid.assign(size_t(ng), 0);
    for (uint64_t i = 0; i < ng; i++) {
	if (!r.u8()) continue;
	b->ConsGsValid[i] = 1;
	size_t p0 = r.pos();
	skip_graph(r);
	b->ConsGs[i].assign(r.data() + p0, r.pos() - p0);
    }
    return b;
}

void chain_graph(const std::string& seq, std::string& out)
{
    // Graph::AddAlignment with an empty alignment (graph.cpp:155-179): codes in order of first appearance, one node per
    // base, one edge per adjacent pair with label 0 (sequences_.size() at the time) and weight 1+1, topological order = ids
    const uint32_t n = uint32_t(seq.size());
    int32_t coder[256], decoder[256];
    for (int i = 0; i < 256; i++) coder[i] = decoder[i] = -1;
    uint32_t num_codes = 0;
    for (uint32_t i = 0; i < n; i++) {
	unsigned char c = static_cast<unsigned char>(seq[i]);
	if (coder[c] == -1) {
	    coder[c] = int32_t(num_codes);
	    decoder[num_codes++] = c;
	}
    }
    const uint32_t ne = n ? n - 1 : 0;
    size_t total = 4 + 2 * (8 + 1024) + 8 + size_t(n) * 9 + 8 + size_t(ne) * 25 + (8 + (n ? 4 : 0)) + 8 + size_t(ne) * 8 + 8 + 8 + size_t(n) * 4 + 8;
    out.resize(total);
    char* p = &out[0];
    auto put = [&](const void* v, size_t k) {
	memcpy(p, v, k);
	p += k;
    };
    auto p32 = [&](uint32_t v) { put(&v, 4); };
    auto p64 = [&](uint64_t v) { put(&v, 8); };
    p32(num_codes);
    p64(256);
    put(coder, 1024);
    p64(256);
    put(decoder, 1024);
    p64(n);
    for (uint32_t i = 0; i < n; i++) {
	*p++ = 1;
	p32(i);
	p32(uint32_t(coder[static_cast<unsigned char>(seq[i])]));
    }
    p64(ne);
    for (uint32_t i = 0; i < ne; i++) {
	*p++ = 1;
	p64(1);
	p32(0);
	p64(2);
    }
    if (n) {  // sequences_: the first node of the one sequence
	p64(1);
	p32(0);
    }
    else {
	p64(0);
    }
    p64(ne);
    for (uint32_t i = 0; i < ne; i++) {
	p32(i);
	p32(i + 1);
    }
    p64(0);  // aligned_nodes
    p64(n);
    for (uint32_t i = 0; i < n; i++) p32(i);
    p64(0);  // consensus
    out.resize(size_t(p - &out[0]));
}

void save_sorted_idx(const std::string& fastq_path, const std::string& path)
{
    Writer w(path);
    w.str(fastq_path);
    w.close();
}

// ---------------------------------------------------------------- sort: streaming batch writer
void save_sorted_batch(const std::string& path, int batch_nr, uint64_t start, uint64_t end, uint64_t bases, const CmdArgs& args,
		       const FlatReads& R, const int32_t* order)
{
    Writer w(path);
    Batch hdr;
    hdr.BatchNr = batch_nr;
    hdr.BatchStart = start;
    hdr.BatchEnd = end;
    hdr.BatchBases = bases;
    hdr.TotalReads = 0;  // never set by `sort` (serialize.h:29; SURVEY App. C#6: observed 0)
    hdr.NrCls = int32_t(end - start + 1);
    hdr.SortArgs = args;
    hdr.Depth = -1;
    write_header(w, hdr);
    w.u64(0);  // MinDB: empty
    const uint64_t n = end - start + 1;
    w.u64(n);
    uint32_t id = 1;
    for (uint64_t s = start; s <= end; s++) {
	const int32_t r = order[s];
	const std::string& name = R.names[r];
	w.u32((id++) | 0x80000000u);  // shared_ptr<Cluster>
	w.u64(1);
	w.u32((id++) | 0x80000000u);  // shared_ptr<ProcSeq>
	const int64_t o = R.offs[r], len = R.offs[r + 1] - o;
	const int st = R.status[r];
	if (st == 1) {	// qualscore.cpp:75-89
	    w.u8(1);
	    w.str(name);
	    w.str(R.seq + o, size_t(len));
	    w.str(R.qual + o, size_t(len));
	    w.f64(R.score[r]);
	    w.f64(R.err[r]);
	    const int64_t hl = R.hpc_len[r];
	    w.u8(1);
	    w.str(name);
	    w.str(R.hpc_seq + o, size_t(hl));
	    w.str(R.hpc_qual + o, size_t(hl));
	    w.f64(R.score[r]);	// HomopolymerCompress copies the score (hpc.cpp:30)
	    w.f64(R.hpc_err[r]);
	    const int64_t nf = R.min_offs[r + 1] - R.min_offs[r], nr = R.rev_offs[r + 1] - R.rev_offs[r];
	    w.u64(uint64_t(nf));
	    w.raw(R.mins + R.min_offs[r], size_t(nf) * sizeof(Minimizer));
	    w.u64(uint64_t(nr));
	    w.raw(R.rev_mins + R.rev_offs[r], size_t(nr) * sizeof(Minimizer));
	    w.i32(1);
	    w.str(name);
	}
	else if (st == 2) {  // qualscore.cpp:91-96: raw read kept with score -1, no sketch
	    w.u8(1);
	    w.str(name);
	    w.str(R.seq + o, size_t(len));
	    w.str(R.qual + o, size_t(len));
	    w.f64(-1.0);
	    w.f64(R.err[r]);
	    w.u8(0);
	    w.u64(0);
	    w.u64(0);
	    w.i32(0);
	    w.str(name);
	}
	else {	// filtered (qualscore.cpp:56-61) or compressed read too short (:65-73)
	    w.u8(0);
	    w.u8(0);
	    w.u64(0);
	    w.u64(0);
	    w.i32(0);
	    w.str(name);
	}
    }
    w.u64(0);  // ConsGs: empty
    w.close();
}

// ---------------------------------------------------------------- cluster bookkeeping
void add_minimizers(const std::vector<Minimizer>& mins, unsigned cls, MinimizerDB& db)
{
    for (const auto& m : mins) {
	auto v = db.find(m.Min);
	if (v == db.end()) {
	    db[m.Min] = std::vector<unsigned>{cls};
	}
	else if (v->second.size() == 0 || cls > v->second.back()) {
	    v->second.emplace_back(cls);
	}
    }
}

std::unique_ptr<Batch> create_pseudo_batch(const Batch& in)
{
    std::unique_ptr<Batch> nb(new Batch);
    nb->BatchNr = -in.BatchNr;
    nb->BatchStart = in.BatchStart;
    nb->BatchEnd = in.BatchEnd;
    nb->BatchBases = 0;
    nb->TotalReads = 0;
    nb->SortArgs = in.SortArgs;
    nb->Depth = -1;
    nb->Cls = in.Cls;
    nb->NrCls = int32_t(nb->Cls.size());
    return nb;
}

static void inconsistent(const Batch& b)
{
    std::cerr << "Inconsistent batch state: NrCluster " << b.NrCls << " vs " << b.Cls.size() << std::endl;
    exit(1);
}
int nr_clusters(const Batch& b)
{
    if (b.NrCls != int(b.Cls.size())) inconsistent(b);
    int count = 0;
    for (const auto& c : b.Cls)
	if (c && c->at(0)->raw && c->at(0)->raw->score > -1) count++;
    return count;
}
int nr_nontrivial_clusters(const Batch& b)
{
    if (b.NrCls != int(b.Cls.size())) inconsistent(b);
    int count = 0;
    for (const auto& c : b.Cls)
	if (c && c->at(0)->raw && c->at(0)->raw->score > -1 && c->size() > 2) count++;
    return count;
}

bool right_is_eligible(const Batch& left, const Batch& right, size_t i, unsigned* size_filtered)
{
    const auto& reads = right.Cls;
    const CmdArgs& args = left.SortArgs;
    if (!reads[i] || reads[i]->size() == 0) return false;  // cluster.cpp:117-119
    const int minClsSize = left.SortArgs.MinClsSize;
    if (right.Depth > 0 && minClsSize > 1 && int(reads[i]->size() - 1) < minClsSize) {	// :120-124
	if (size_filtered) (*size_filtered)++;
	return false;
    }
    const auto& read = reads[i]->at(0);
    if (!read || !read->raw) return false;  // :126-128
    SeqRec& seq = *read->raw;
    if (seq.score < 0) return false;  // :145-147
    // the reference dereferences HpcSeq before these tests (:130-131); a batch written by `sort` never has Raw without Hpc
    // unless the score is negative, which returned above
    if (seq.seq.length() < unsigned(2 * args.KmerSize) || !read->hpc || read->hpc->seq.length() < unsigned(2 * args.KmerSize)) {  // :148-155
	seq.score = -1.0;
	return false;
    }
    if ((-10 * log10(seq.err)) <= args.MinQual) {  // :157-160
	seq.score = -1.0;
	return false;
    }
    return true;
}

void apply_decision(Batch& left, Batch& right, size_t i, int best, int strand, bool seq_purge)
{
    auto& cls = left.Cls;
    auto& reads = right.Cls;
    if (best == -2) return;
    if (best == -1) {  // cluster.cpp:177-223
	const unsigned newId = unsigned(cls.size());
	const size_t nrReads = reads[i]->size();
	add_minimizers(reads[i]->at(0)->mins, newId, left.MinDB);
	if (nrReads == 1) {
	    const auto& rep = reads[i]->at(0);
	    auto nrep = std::make_shared<ProcSeq>();
	    nrep->raw.reset(new SeqRec(*rep->raw));
	    nrep->hpc.reset(new SeqRec(*rep->hpc));
	    nrep->mins = rep->mins;
	    nrep->rev = rep->rev;
	    nrep->strand = rep->strand;
	    nrep->id = rep->id;
	    const std::string repName = "rep_" + std::to_string(left.BatchNr) + "_" + std::to_string(newId);
	    nrep->raw->name = repName;
	    nrep->hpc->name = repName;
	    reads[i]->insert(reads[i]->begin(), 1, std::move(nrep));
	}
	left.ConsGs.emplace_back();
	left.ConsGsValid.push_back(1);
	chain_graph(reads[i]->at(0)->raw->seq, left.ConsGs.back());
	cls.emplace_back(reads[i]);
	if (nrReads == 1 && cls[newId]->size() != 2) {
	    std::cerr << "Inconsistent initial cluster size " << cls[newId]->size() << " for read: " << cls[newId]->at(0)->raw->name
		      << std::endl;
	    std::cerr << "Aborting clustering!" << std::endl;
	    exit(1);
	}
	left.NrCls++;
	if (right.ConsGs.size() > 0 && right.ConsGsValid[i]) {
	    right.ConsGsValid[i] = 0;
	    right.ConsGs[i].clear();
	}
    }
    else {  // cluster.cpp:224-261 (what follows :263 is the caller's: consensus)
	auto& dst = *cls[size_t(best)];
	dst.reserve(dst.size() + reads[i]->size());
	for (size_t j = 0; j < reads[i]->size(); j++) {
	    auto& s = reads[i]->at(j);
	    if (!s) {
		std::cerr << "Null pointer at position " << j << " in read array " << i << " of size: " << reads[i]->size() << std::endl;
		exit(1);
	    }
	    if (strand == -1) {
		if (s->strand == 1)
		    s->strand = -1;
		else if (s->strand == -1)
		    s->strand = 1;
		else
		    throw std::runtime_error("Invalid match strand!");
	    }
	    s->mins.clear();
	    s->mins.shrink_to_fit();
	    s->rev.clear();
	    s->rev.shrink_to_fit();
	    if (!seq_purge) {
		s->raw.reset();
		s->hpc.reset();
	    }
	}
	size_t first = reads[i]->size() > 1 ? 1 : 0;
	for (size_t j = first; j < reads[i]->size(); j++) dst.emplace_back(std::move(reads[i]->at(j)));
    }
}

void finish_decisions(Batch& left, Batch& right)  // cluster.cpp:319-321
{
    left.Depth++;
    left.BatchEnd = right.BatchEnd;
    left.BatchBases = left.BatchBases + right.BatchBases;
}

void apply_decisions(Batch& left, Batch& right, const std::vector<int32_t>& best_of, const std::vector<int32_t>& strand_of, bool seq_purge)
{
    left.ConsGs.reserve(left.Cls.size());
    for (size_t i = 0; i < right.Cls.size(); i++) apply_decision(left, right, i, best_of[i], strand_of[i], seq_purge);
    finish_decisions(left, right);
}

// UpdateMinDB (minimizer.cpp:124-160) on the batch's own container: operator[] creates the entries the reference creates,
// so the map's iteration order — the order of the `.cer` — stays the reference's
void update_min_db(unsigned best, const std::vector<Minimizer>& old_mins, const std::vector<Minimizer>& new_mins, MinimizerDB& db)
{
    std::set<unsigned> oldSet, newSet, toIns, toDel;
    for (const auto& m : old_mins) oldSet.insert(m.Min);
    for (const auto& m : new_mins) newSet.insert(m.Min);
    std::set_difference(oldSet.begin(), oldSet.end(), newSet.begin(), newSet.end(), std::inserter(toDel, toDel.begin()));
    std::set_difference(newSet.begin(), newSet.end(), oldSet.begin(), oldSet.end(), std::inserter(toIns, toIns.begin()));
    for (unsigned m : toDel) {
	auto& mins = db[m];
	std::set<unsigned> tmp(mins.begin(), mins.end());
	tmp.erase(best);
	mins.assign(tmp.begin(), tmp.end());
    }
    for (unsigned m : toIns) {
	auto& tv = db[m];
	tv.push_back(best);
	std::sort(tv.begin(), tv.end());
    }
}

}  // namespace icb
