// cer.hpp — host-side batch model and the `.cer` byte format of isONclust2 (cereal BinaryOutputArchive, native
// little-endian, no header).  Written from the format description in SURVEY.md App. B; nothing here uses cereal.
//
// What each piece mirrors in the reference (paths relative to the reference tree):
//   CmdArgs          args.h:9-37 (field order of serialize(), :29-36)
//   SeqRec           Seq, seq.h:18-98 (serialize :63-67: name, seq, qual, score, errorRate)
//   ProcSeq          cluster_data.h:14-27 (RawSeq, HpcSeq, Mins, RevMins, MatchStrand, Id)
//   Batch            serialize.h:24-43 (BatchNr .. ConsGs)
//   MinimizerDB      minimizer.h:36-41, 60-61 — a REAL std::unordered_map with the identity hash: the file stores the map
//                    in container iteration order, so the same container with the same operation
//                    history (reserve 10^6 when empty, cluster.cpp:92-94; emplace_hint on load) is the exact route
//   graph blobs      spoa::Graph::save, vendor/spoa/include/spoa/graph.hpp:188-230; with consensus off every graph is the
//                    chain of one sequence (cluster.cpp:200-204), written by chain_graph(); loaded graphs pass through
#ifndef ICB_CER_HPP
#define ICB_CER_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace icb {

enum ClsMode { Sahlin = 0, Fast = 1, Furious = 2, None = 3 };  // args.h:7

struct CmdArgs {
    bool Verbose = false, Debug = false;
    std::string InFastq;
    int KmerSize = 11, BatchSize = 50000, BatchMaxSeq = 30000, WindowSize = 15, MinShared = 5, ConsMinSize = 50, ConsMaxSize = -150,
	ConsPeriod = 500, MinClsSize = 3;
    double MinQual = 7.0, MappedThreshold = 0.65, AlignedThreshold = 0.2, MinFraction = 0.8, MinProbNoHits = 0.1;
    std::string BatchOutFolder = "isONclust2_batches";
    int Mode = Sahlin;
};
// operator== of args.cpp:426-462 (InFastq, BatchOutFolder, BatchMaxSeq and Mode are not compared)
bool same_sort_args(const CmdArgs& a, const CmdArgs& b);

struct SeqRec {
    std::string name, seq, qual;
    double score = 0.0, err = 0.0;
};
struct Minimizer {
    uint32_t Min, Pos, Index;
};
struct ProcSeq {
    std::unique_ptr<SeqRec> raw, hpc;
    std::vector<Minimizer> mins, rev;
    int32_t strand = 0;
    std::string id;
};
typedef std::vector<std::shared_ptr<ProcSeq>> Cluster;
typedef std::vector<std::shared_ptr<Cluster>> Clusters;

struct UnsignedHash {
    inline std::size_t operator()(const unsigned& u) const { return std::size_t(u); }
};
typedef std::unordered_map<unsigned, std::vector<unsigned>, UnsignedHash> MinimizerDB;
static const std::size_t MIN_DB_RESERVE = 1000000;  // minimizer.h:14

struct Batch {
    int32_t BatchNr = 0;
    uint64_t BatchStart = 0, BatchEnd = 0, BatchBases = 0;
    int32_t TotalReads = 0, NrCls = 0;
    CmdArgs SortArgs;
    std::string LeftLeaf, RightLeaf;
    int32_t Depth = 0;
    MinimizerDB MinDB;
    Clusters Cls;
    std::vector<std::string> ConsGs;   // serialized spoa::Graph per cluster (without the unique_ptr validity byte)
    std::vector<uint8_t> ConsGsValid;  // unique_ptr validity per graph
};

// ---------------------------------------------------------------- byte sink / source
class Writer {
public:
    explicit Writer(const std::string& path);
    ~Writer();
    void close();
    inline void raw(const void* p, size_t n)
    {
	if (n > cap_ - len_) {
	    flush();
	    if (n > cap_) {
		put(p, n);
		return;
	    }
	}
	memcpy(buf_ + len_, p, n);
	len_ += n;
    }
    inline void u8(uint8_t v) { raw(&v, 1); }
    inline void i32(int32_t v) { raw(&v, 4); }
    inline void u32(uint32_t v) { raw(&v, 4); }
    inline void u64(uint64_t v) { raw(&v, 8); }
    inline void f64(double v) { raw(&v, 8); }
    inline void str(const char* p, size_t n)
    {
	u64(n);
	raw(p, n);
    }
    inline void str(const std::string& s) { str(s.data(), s.size()); }
    uint64_t bytes() const { return total_ + len_; }

private:
    void flush();
    void put(const void* p, size_t n);
    FILE* f_;
    char* buf_;
    size_t len_, cap_;
    uint64_t total_;
    std::string path_;
};

class Reader {
public:
    explicit Reader(const std::string& path);  // whole file in memory
    inline void raw(void* p, size_t n)
    {
	need(n);
	memcpy(p, d_.data() + o_, n);
	o_ += n;
    }
    inline uint8_t u8()
    {
	uint8_t v;
	raw(&v, 1);
	return v;
    }
    inline int32_t i32()
    {
	int32_t v;
	raw(&v, 4);
	return v;
    }
    inline uint32_t u32()
    {
	uint32_t v;
	raw(&v, 4);
	return v;
    }
    inline uint64_t u64()
    {
	uint64_t v;
	raw(&v, 8);
	return v;
    }
    inline double f64()
    {
	double v;
	raw(&v, 8);
	return v;
    }
    std::string str();
    void skip(size_t n)
    {
	need(n);
	o_ += n;
    }
    size_t pos() const { return o_; }
    size_t size() const { return d_.size(); }
    const char* data() const { return d_.data(); }
    bool eof() const { return o_ == d_.size(); }

private:
    inline void need(size_t n)
    {
	// cereal: "Failed to read N bytes from input stream! Read M" (binary.hpp) -> LoadBatch prints and exit(1)
	if (n > d_.size() - o_)
	    throw std::runtime_error("Failed to read " + std::to_string(n) + " bytes from input stream! Read " +
				     std::to_string(d_.size() - o_));
    }
    std::string d_;
    size_t o_;
};

// ---------------------------------------------------------------- whole batches
void write_args(Writer& w, const CmdArgs& a);
void read_args(Reader& r, CmdArgs& a);
// SaveBatch / LoadBatch (serialize.cpp:5-27)
void save_batch(const Batch& b, const std::string& path);
std::unique_ptr<Batch> load_batch(const std::string& path);
// the spoa graph of one sequence added with weight 1 to an empty graph (cluster.cpp:200-204, consensus.cpp:15-22)
void chain_graph(const std::string& seq, std::string& out);
// SortedIdx (output.h:15-23): one string
void save_sorted_idx(const std::string& fastq_path, const std::string& path);

// ---------------------------------------------------------------- streaming writer for `sort`
// Emits the bytes SaveBatch produces for the Batch PrepareSortedBatch builds (qualscore.cpp:39-105) straight from the
// flat arrays the device stages return — no per-read objects.  Slot s of the batch is read order[s].
struct FlatReads {
    const char* seq;
    const char* qual;
    const int64_t* offs;
    const std::string* names;
    const double* score;
    const double* err;
    // sketch download
    const int32_t* status;
    const int32_t* hpc_len;
    const double* hpc_err;
    const char* hpc_seq;
    const char* hpc_qual;
    const Minimizer* mins;
    const int64_t* min_offs;
    const Minimizer* rev_mins;
    const int64_t* rev_offs;
};
void save_sorted_batch(const std::string& path, int batch_nr, uint64_t start, uint64_t end, uint64_t bases, const CmdArgs& args,
		       const FlatReads& R, const int32_t* order);

// ---------------------------------------------------------------- ClusterSortedReads' bookkeeping (cluster.cpp:67-322, consensus off)
// Eligibility of right cluster i for getBestCluster (the skips of cluster.cpp:116-160): evaluated on the host from the
// loaded batch; returns false if the read is skipped.
bool right_is_eligible(const Batch& left, const Batch& right, size_t i, unsigned* size_filtered);
// Applies the per-read decisions to the batches exactly as the reference's loop does: best[i] = cluster id read i joins,
// -1 = founds a new cluster, -2 = skipped; strand[i] = matched strand of a join.  seq_purge = `-j` (cluster.cpp:249-252).
void apply_decisions(Batch& left, Batch& right, const std::vector<int32_t>& best, const std::vector<int32_t>& strand, bool seq_purge);
// one read of that loop / the three assignments behind it (cluster.cpp:319-321): for a caller that interleaves its own work
// between the reads (consensus, cluster.cpp:263-309)
void apply_decision(Batch& left, Batch& right, size_t i, int best, int strand, bool seq_purge);
void finish_decisions(Batch& left, Batch& right);
// UpdateMinDB (minimizer.cpp:124-160)
void update_min_db(unsigned best, const std::vector<Minimizer>& old_mins, const std::vector<Minimizer>& new_mins, MinimizerDB& db);
// AddMinimizers (minimizer.cpp:31-42)
void add_minimizers(const std::vector<Minimizer>& mins, unsigned cls, MinimizerDB& db);
// CreatePseudoBatch (serialize.cpp:29-43)
std::unique_ptr<Batch> create_pseudo_batch(const Batch& in);
// Batch::NrClusters / NrNontrivialClusters (serialize.h:45-79)
int nr_clusters(const Batch& b);
int nr_nontrivial_clusters(const Batch& b);

}  // namespace icb
#endif
