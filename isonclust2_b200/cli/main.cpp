// isONclust2 (B200 build) — the `sort` / `cluster` / `dump` / `info` subcommands of the reference's CLI (main.cpp, args.cpp) on top
// of the C ABI in include/isonclust2_b200.h.  Same flags, same output files, byte-identical `.cer` batches; the
// data-parallel work (quality scores, sketches, candidate queries, alignments) runs on the device, the host keeps the
// FASTQ / `.cer` I/O and the order-dependent bookkeeping (cer.cpp).  There is no CPU fallback: without a CUDA device every
// subcommand that computes exits 1 (`dump` and `info` only read files and need no device).
//
// Reference call sites mirrored here: mainSort main.cpp:74-202, mainCluster main.cpp:238-392, mainInfo main.cpp:394-409,
// ParseArgsSort args.cpp:11-151, ParseArgsCluster args.cpp:154-259, SequencesPToFastq / WriteScores output.cpp:30-91.
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "cer.hpp"
#include "isonclust2_b200.h"

using namespace icb;
using std::cerr;
using std::endl;

#ifndef ICB_VERSION
#define ICB_VERSION "v2.4-b200"
#endif

static bool VERBOSE = false;

// ---------------------------------------------------------------- help (args.cpp:318-424)
static void print_help_sort()
{
    std::cout << "\nsort - sort reads and write out batches:\n"
		 "\t-B --batch-size        Batch size in kilobases (default: 50000)\n"
		 "\t-M --batch-max-seq     Maximum number of sequences per batch (default: 3000).\n"
		 "\t-k --kmer-size         Kmer size (default: 11).\n"
		 "\t-w --window-size       Window size (default: 15).\n"
		 "\t-m --min-shared        Minimum number of minimizers shared between read and cluster (default: 5).\n"
		 "\t-q --min-qual          Minimum average quality value (default: 7.0).\n"
		 "\t-x --mode  Clustering mode:\n"
		 "\t           * sahlin (default): use minimizers first, alignment second \n"
		 "\t           * fast: use minimizers only\n"
		 "\t           * furious: always use alignment\n"
		 "\t-g --low-cons-size     Use all sequences for consensus below this size (default: 20).\n"
		 "\t-c --max-cons-size     Maximum number of sequences used for consensus (default: 150).\n"
		 "\t-P --cons-period       Do not recalculate consensus after this many seuqences added (default: 500).\n"
		 "\t-r --mapped-threshold  Minmum mapped fraction of read to be \tincluded in cluster (default: 0.65).\n"
		 "\t-a --aligned-threshold Minimum aligned fraction of read to be included in cluster (default: 0.2).\n"
		 "\t-f --min-fraction      Minimum fraction of minimizers shared compared to best hit, in order to continue "
		 "mapping (default: 0.8).\n"
		 "\t-p --min-prob-no-hits  Minimum probability for i consecutive \tminimizers to be different between read and "
		 "representative (default: 0.1)\n"
		 "\t-F --min-cls-size      Skip clusters smaller than this in the left batch (default: 3).\n"
		 "\t-o --outfolder         Output folder (default: \t./isONclust2_batches).\n"
		 "\t-h --help              Print help.\n"
		 "\t-v --verbose           Verbose output.\n"
		 "\t-d --debug             Print debug info.\n"
		 "\t[positional argument]  Input fastq file (required).\n";
}
static void print_help_cluster()
{
    std::cout << "\ncluster - cluster and/or merge batches:\n"
		 "\t-l --left-batch        Left input batch (mandatory).\n"
		 "\t-r --right-batch       Right input batch (optional).\n"
		 "\t-o --outfile           Output batch.\n"
		 "\t-x --mode  Clustering mode:\n"
		 "\t           * sahlin (default): use minimizers first, alignment second \n"
		 "\t           * fast: use minimizers only\n"
		 "\t           * furious: use alignment only\n"
		 "\t-A --spoa-algo  spoa alignment algorithm:\n"
		 "\t           * 0 (default): local\n"
		 "\t           * 1 : global\n"
		 "\t           * 2 : semi-global\n"
		 "\t-z --min-purge         Purge minimizer database from output batch.\n"
		 "\t-j --keep-seq          Do not purge non-representative sequences from output batches.\n"
		 "\t-F --min-cls-size      Skip clusters smaller than this in the left batch.\n"
		 "\t-v --verbose           Verbose output.\n"
		 "\t-Q --quiet             Supress progress bar.\n"
		 "\t-d --debug             Print debug info.\n"
		 "\t-h --help              Print help.\n";
}
static void print_help_info()
{
    std::cout << "\ninfo:\n"
		 "\t-h --help              Print help.\n"
		 "\t[positional argument]  Input serialized batch (required).\n";
}
static void print_help_dump()
{
    std::cout << "\ndump - dump clustered batch:\n"
		 "\t-o --outdir            Output directory.\n"
		 "\t-i --index             Index of sorted reads.\n"
		 "\t-v --verbose           Verbose output.\n"
		 "\t-d --debug             Print debug info.\n"
		 "\t-h --help              Print help.\n";
}
static void print_help()
{
    std::cout << "isONclust2 version: " << ICB_VERSION << endl;
    std::cout << "Available subcommands: sort, cluster, dump, info, help, version" << endl;
    print_help_sort();
    print_help_cluster();
    print_help_dump();
    print_help_info();
    std::cout << "\nhelp - print help message\n";
    std::cout << "\nversion - print version\n";
}

// ---------------------------------------------------------------- argument parsing
static bool parse_mode(const char* s, int* mode)
{
    std::string m = s ? s : "";
    if (m == "sahlin")
	*mode = Sahlin;
    else if (m == "fast")
	*mode = Fast;
    else if (m == "furious")
	*mode = Furious;
    else {
	cerr << "Illegal clustering mode: " << m << endl;
	print_help_sort();
	return false;
    }
    return true;
}

static CmdArgs parse_args_sort(int argc, char** argv)
{
    static const struct option longopts[] = {
	{"version", no_argument, 0, 'V'},	   {"verbose", no_argument, 0, 'v'},
	{"debug", no_argument, 0, 'd'},		   {"mode", optional_argument, 0, 'x'},
	{"help", no_argument, 0, 'h'},		   {"kmer-size", required_argument, 0, 'k'},
	{"window-size", required_argument, 0, 'w'}, {"min-shared", required_argument, 0, 'm'},
	{"mapped-threshold", required_argument, 0, 'r'}, {"aligned-threshold", required_argument, 0, 'a'},
	{"min-fraction", required_argument, 0, 'f'}, {"min-prob-no-hits", required_argument, 0, 'p'},
	{"min-qual", required_argument, 0, 'q'},	   {"min-cls-size", required_argument, 0, 'F'},
	{"low-cons-size", required_argument, 0, 'g'}, {"max-cons-size", required_argument, 0, 'c'},
	{"cons-period", required_argument, 0, 'P'}, {"outfolder", required_argument, 0, 'o'},
	{"batch-size", required_argument, 0, 'B'},  {"batch-max-seq", required_argument, 0, 'M'},
	{0, 0, 0, 0},
    };
    CmdArgs res;
    argc--;
    argv++;  // getopt sees "sort" as the program name
    int iarg = 0, index;
    while ((iarg = getopt_long(argc, argv, "k:w:dhvo:m:r:a:f:p:q:B:x:g:c:M:P:F:", longopts, &index)) != -1) {
	switch (iarg) {
	    case 'h': print_help_sort(); exit(0);
	    case 'k': res.KmerSize = atoi(optarg); break;
	    case 'm': res.MinShared = atoi(optarg); break;
	    case 'r': res.MappedThreshold = atof(optarg); break;
	    case 'a': res.AlignedThreshold = atof(optarg); break;
	    case 'f': res.MinFraction = atof(optarg); break;
	    case 'p': res.MinProbNoHits = atof(optarg); break;
	    case 'q': res.MinQual = atof(optarg); break;
	    case 'w': res.WindowSize = atoi(optarg); break;
	    case 'B': res.BatchSize = atoi(optarg); break;
	    case 'M': res.BatchMaxSeq = atoi(optarg); break;
	    case 'P': res.ConsPeriod = atoi(optarg); break;
	    case 'g': res.ConsMinSize = atoi(optarg); break;
	    case 'c': res.ConsMaxSize = atoi(optarg); break;
	    case 'F': res.MinClsSize = atoi(optarg); break;
	    case 'o': res.BatchOutFolder = optarg; break;
	    case 'v': res.Verbose = true; break;
	    case 'd': res.Debug = true; break;
	    case 'x':
		if (!parse_mode(optarg, &res.Mode)) exit(1);
		break;
	    default: break;
	}
    }
    if (argc - optind != 1) {
	cerr << "Please specify one input fastq file!" << endl;
	exit(1);
    }
    if (res.KmerSize > 31) {
	cerr << "Maximum supported kmer size is 31!" << endl;
	exit(1);
    }
    if (res.KmerSize < 10) {
	cerr << "Minimum supported kmer size is 10!" << endl;
	exit(1);
    }
    if (res.KmerSize > res.WindowSize) {
	cerr << "Kmer size cannot be larger than the window size!" << endl;
	exit(1);
    }
    res.InFastq = argv[optind];
    return res;
}

struct CmdArgsCluster {
    bool Verbose = false, Quiet = false, Debug = false, MinPurge = false, SeqPurge = false;
    int MinClsSize = -1;
    std::string LeftCereal, RightCereal, OutCereal;
    int Mode = None;
    int SpoaAlgo = 2;
};

static CmdArgsCluster parse_args_cluster(int argc, char** argv)
{
    static const struct option longopts[] = {
	{"quiet", no_argument, 0, 'Q'},		 {"version", no_argument, 0, 'V'},
	{"verbose", no_argument, 0, 'v'},	 {"min-purge", no_argument, 0, 'z'},
	{"min-cls-size", required_argument, 0, 'F'}, {"keep-seq", no_argument, 0, 'j'},
	{"debug", no_argument, 0, 'd'},		 {"spoa-algo", optional_argument, 0, 'A'},
	{"mode", optional_argument, 0, 'x'},	 {"help", no_argument, 0, 'h'},
	{"outfile", required_argument, 0, 'o'},	 {"left-batch", required_argument, 0, 'l'},
	{"right-batch", optional_argument, 0, 'r'}, {0, 0, 0, 0},
    };
    CmdArgsCluster res;
    argc--;
    argv++;
    int iarg = 0, index;
    while ((iarg = getopt_long(argc, argv, "Vdhvo:l:r:Qx:A:zjF:", longopts, &index)) != -1) {
	switch (iarg) {
	    case 'h': print_help_cluster(); exit(0);
	    case 'o': res.OutCereal = optarg; break;
	    case 'l': res.LeftCereal = optarg; break;
	    case 'r': res.RightCereal = optarg ? optarg : ""; break;
	    case 'v': res.Verbose = true; break;
	    case 'Q': res.Quiet = true; break;
	    case 'd': res.Debug = true; break;
	    case 'A': res.SpoaAlgo = optarg ? atoi(optarg) : res.SpoaAlgo; break;
	    case 'z': res.MinPurge = true; break;
	    case 'j': res.SeqPurge = true; break;
	    case 'F': res.MinClsSize = atoi(optarg); break;
	    case 'x':
		if (!parse_mode(optarg, &res.Mode)) exit(1);
		break;
	    default: break;
	}
    }
    if (res.LeftCereal == "") {
	cerr << "Specifying left input batch is mandatory!" << endl;
	print_help();
	exit(1);
    }
    if (res.OutCereal == "") {
	cerr << "Specifying output batch file is mandatory!" << endl;
	print_help();
	exit(1);
    }
    return res;
}

// ---------------------------------------------------------------- small helpers
static int dir_exists(const std::string& path)
{
    struct stat info;
    if (stat(path.c_str(), &info) != 0) return 0;
    return (info.st_mode & S_IFDIR) ? 1 : 0;
}
static void create_outdir(const std::string& d)  // output.cpp:110-123
{
    if (dir_exists(d)) {
	cerr << "Warning: reusing existing output directory: " << d << endl;
	return;
    }
    int result = mkdir(d.c_str(), 0755);
    if (result != 0) {
	cerr << "Failed to create output directory!"
	     << "Error: " << result << endl;
	exit(1);
    }
}
static FILE* create_file(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "wb");
    if (!f) {
	cerr << "Failed to open " + p + "!" << endl;
	exit(1);
    }
    setvbuf(f, nullptr, _IOFBF, 1 << 22);
    return f;
}

static void die_icb(icb_ctx* ctx, int rc, const char* what)
{
    const char* msg = ctx ? icb_last_error(ctx) : "";
    if (rc == ICB_ERR_INVALID_BASE) {
	// the reference: RevComp throws a std::string (util.cpp:31-33), nobody catches it -> terminate
	cerr << "terminate called after throwing an instance of 'std::string' (a read holds a base outside ACGT)" << endl;
	abort();
    }
    cerr << "isONclust2 (B200): " << what << " failed: [" << rc << "] " << ((msg && *msg) ? msg : icb_strerror(rc)) << endl;
    exit(1);
}
#define ICB_DO(ctx, call)                   \
    do {                                    \
	int rc__ = (call);                  \
	if (rc__) die_icb(ctx, rc__, #call); \
    } while (0)

static icb_ctx* open_device()
{
    const char* dev = getenv("ICB_DEVICE");
    icb_ctx* ctx = nullptr;
    int rc = icb_ctx_create(dev ? atoi(dev) : 0, &ctx);
    if (rc) {
	cerr << "isONclust2 (B200): no usable CUDA device (" << icb_strerror(rc) << "); this build has no CPU path." << endl;
	exit(1);
    }
    return ctx;
}

// ---------------------------------------------------------------- FASTQ (bioparser FastqParser semantics, fastq_parser.hpp:27-118:
// names cut at the first white space, multi-line sequences/qualities joined, gzip or plain input through zlib)
struct Fastq {
    std::vector<std::string> names;
    std::string seq, qual;
    std::vector<int64_t> offs{0};
};

static void invalid_fastq()
{
    cerr << "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  [bioparser::FastqParser] error: invalid file format" << endl;
    abort();
}

static void parse_fastq(const std::string& path, Fastq& fq)
{
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) {
	cerr << "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  [bioparser::Parser::Create] error: unable to open file " << path << endl;
	abort();
    }
    gzbuffer(f, 1 << 20);
    std::vector<char> buf(size_t(4) << 20);
    enum { NAME, DATA, COMMENT, QUALITY } st = NAME;
    std::string name;
    size_t data0 = fq.seq.size(), qual0 = fq.qual.size();
    bool any = false;
    auto finish = [&]() {
	size_t nl = 0;
	while (nl < name.size() && !isspace(static_cast<unsigned char>(name[nl]))) nl++;
	const size_t dl = fq.seq.size() - data0, ql = fq.qual.size() - qual0;
	if (nl <= 1 || name[0] != '@' || dl == 0 || ql == 0 || dl != ql) invalid_fastq();
	fq.names.emplace_back(name.data() + 1, nl - 1);
	fq.offs.push_back(int64_t(fq.seq.size()));
	name.clear();
	data0 = fq.seq.size();
	qual0 = fq.qual.size();
	any = false;
	st = NAME;
    };
    for (;;) {
	int n = gzread(f, buf.data(), unsigned(buf.size()));
	if (n <= 0) break;
	for (int i = 0; i < n; i++) {
	    const char c = buf[i];
	    if (c == '\n') {
		if (st == NAME) {
		    st = DATA;
		}
		else if (st == COMMENT) {
		    st = QUALITY;
		}
		else if (st == QUALITY && fq.qual.size() - qual0 >= fq.seq.size() - data0) {
		    finish();
		}
		continue;
	    }
	    any = true;
	    switch (st) {
		case NAME: name.push_back(c); break;
		case DATA:
		    if (c == '+')
			st = COMMENT;
		    else if (!isspace(static_cast<unsigned char>(c)))
			fq.seq.push_back(c);
		    break;
		case COMMENT: break;
		case QUALITY:
		    if (!isspace(static_cast<unsigned char>(c))) fq.qual.push_back(c);
		    break;
	    }
	}
    }
    gzclose(f);
    if (any) {
	if (st != QUALITY) invalid_fastq();
	finish();
    }
}

// ---------------------------------------------------------------- sort (main.cpp:74-202)
static int main_sort(int argc, char** argv)
{
    CmdArgs args = parse_args_sort(argc, argv);
    VERBOSE = args.Verbose;
    if (VERBOSE) {
	cerr << "isONclust2 version: " << ICB_VERSION << endl;
	cerr << "Batches output directory: " << args.BatchOutFolder << endl;
	cerr << "Minimum batch size: " << args.BatchSize << " kilobases" << endl;
	cerr << "Kmer size: " << args.KmerSize << endl;
	cerr << "Window size: " << args.WindowSize << endl;
	cerr << "Consensus period: " << args.ConsPeriod << endl;
	cerr << "Minimum cluster size for consensus: " << args.ConsMinSize << endl;
	cerr << "Maximum cluster size for consensus: " << args.ConsMaxSize << endl;
	cerr << "Minimum average quality: " << args.MinQual << endl;
	cerr << "Minimum shared minimizers: " << args.MinShared << endl;
	cerr << "Minimum fraction of top minimizer hit: " << args.MinFraction << endl;
	cerr << "Mapping threshold: " << args.MappedThreshold << endl;
	cerr << "Alignment threshold: " << args.AlignedThreshold << endl;
	cerr << "Minimum probability no hit: " << args.MinProbNoHits << endl;
	cerr << "Minimum cluster size in left batches: " << args.MinClsSize << endl;
	cerr << "Debug output: " << (args.Debug ? "on" : "off") << endl;
    }
    const std::string batchDir = args.BatchOutFolder + "/batches";
    create_outdir(args.BatchOutFolder);
    create_outdir(batchDir);

    Fastq fq;
    parse_fastq(args.InFastq, fq);
    const int64_t n = int64_t(fq.names.size());
    if (VERBOSE) cerr << "Parsed " << n << " sequences." << endl;

    // FillQualScores + SortByQualScores
    double qualTab[129], qualTabNomin[129];
    icb_init_qual_tab(qualTab, 0);
    icb_init_qual_tab(qualTabNomin, 1);
    icb_ctx* ctx = open_device();
    icb_readset* rs = nullptr;
    ICB_DO(ctx, icb_readset_create(ctx, fq.seq.data(), fq.qual.data(), fq.offs.data(), n, &rs));
    const size_t un = static_cast<size_t>(n);
    std::vector<double> score(un, 0.0), err(un, 0.0);
    std::vector<int32_t> order(un, 0);
    if (n) {
	ICB_DO(ctx, icb_score_reads(rs, args.KmerSize, qualTab, qualTabNomin));
	ICB_DO(ctx, icb_scores_download(rs, score.data(), err.data()));
	ICB_DO(ctx, icb_sort_by_score(score.data(), n, order.data()));
    }
    if (VERBOSE) cerr << "Finished sorting sequences." << endl;

    // SequencesPToFastq + WriteScores (output.cpp:30-91)
    const std::string sortedFastq = args.BatchOutFolder + "/sorted_reads.fastq";
    {
	FILE* out = create_file(sortedFastq);
	FILE* tsv = create_file(args.BatchOutFolder + "/sorted_reads_idx.tsv");
	fputs("Id\tPos\n", tsv);
	unsigned long long seeker = 0;
	for (int64_t s = 0; s < n; s++) {
	    const int32_t r = order[s];
	    if (score[r] < 0) continue;
	    const std::string& name = fq.names[r];
	    const int64_t o = fq.offs[r], len = fq.offs[r + 1] - o;
	    fprintf(tsv, "%s\t%llu\n", name.c_str(), seeker);
	    fputc('@', out);
	    fwrite(name.data(), 1, name.size(), out);
	    fputc('\n', out);
	    fwrite(fq.seq.data() + o, 1, size_t(len), out);
	    fputs("\n+\n", out);
	    fwrite(fq.qual.data() + o, 1, size_t(len), out);
	    fputc('\n', out);
	    seeker += unsigned(name.length() + size_t(len) + size_t(len) + 6);	// WriteFastqRecord returns unsigned
	}
	fclose(out);
	fclose(tsv);
	save_sorted_idx(sortedFastq, args.BatchOutFolder + "/sorted_reads_idx.cer");
    }
    if (VERBOSE) cerr << "Sorted sequences written to: " << sortedFastq << endl;
    const std::string scoresTsv = args.BatchOutFolder + "/scores.tsv";
    {
	FILE* out = create_file(scoresTsv);
	for (int64_t s = 0; s < n; s++) fprintf(out, "%s\t%g\n", fq.names[order[s]].c_str(), score[order[s]]);  // default ostream precision
	fclose(out);
    }
    if (VERBOSE) {
	cerr << "Scores written to: " << scoresTsv << endl;
	cerr << "Preparing batches:" << endl;
    }

    // PrepareSortedBatch for every read at once (the per-read work does not depend on the batch it lands in)
    std::vector<uint8_t> skip(un, 0);
    std::vector<int32_t> status(un, 0), hpc_len(un, 0);
    std::vector<double> hpc_err(un, 0.0);
    std::vector<char> hpc_seq, hpc_qual;
    std::vector<Minimizer> mins, rev;
    std::vector<int64_t> min_offs(un + 1, 0), rev_offs(un + 1, 0);
    if (n) {
	ICB_DO(ctx, icb_quality_filter(err.data(), n, args.MinQual, skip.data()));
	ICB_DO(ctx, icb_sketch_reads(rs, args.KmerSize, args.WindowSize, qualTabNomin, skip.data()));
	int64_t nf = 0, nr = 0;
	ICB_DO(ctx, icb_sketch_sizes(rs, &nf, &nr));
	hpc_seq.resize(fq.seq.size() + 1);
	hpc_qual.resize(fq.seq.size() + 1);
	mins.resize(size_t(nf) + 1);
	rev.resize(size_t(nr) + 1);
	ICB_DO(ctx, icb_sketch_download(rs, status.data(), hpc_len.data(), hpc_err.data(), hpc_seq.data(), hpc_qual.data(),
					reinterpret_cast<icb_minimizer*>(mins.data()), min_offs.data(),
					reinterpret_cast<icb_minimizer*>(rev.data()), rev_offs.data()));
    }
    icb_readset_destroy(rs);
    icb_ctx_destroy(ctx);

    FlatReads R{fq.seq.data(), fq.qual.data(), fq.offs.data(), fq.names.data(), score.data(), err.data(),
		status.data(), hpc_len.data(), hpc_err.data(), hpc_seq.data(), hpc_qual.data(), mins.data(),
		min_offs.data(), rev.data(), rev_offs.data()};
    // the batch windows of main.cpp:149-199
    unsigned long batchBases = 0;
    int batchSeqs = 0, nrBatches = 0, batchStart = 0;
    int64_t i = 0;
    auto emit = [&](int64_t last) {
	const std::string outFile = batchDir + "/isONbatch_" + std::to_string(nrBatches) + ".cer";
	save_sorted_batch(outFile, nrBatches, uint64_t(batchStart), uint64_t(last), batchBases, args, R, order.data());
	if (VERBOSE)
	    cerr << "\tWritten batch " << nrBatches << " with " << (i - batchStart + 1) << " sequences and "
		 << int(double(batchBases) / 1000.0) << " kilobases." << endl;
    };
    for (i = 0; i < n; i++) {
	batchBases += static_cast<unsigned long>(fq.offs[order[i] + 1] - fq.offs[order[i]]);
	batchSeqs++;
	if ((args.BatchSize > 0) &&
	    ((batchBases > static_cast<unsigned long>(args.BatchSize * 1000)) || ((args.BatchMaxSeq > 0) && batchSeqs >= args.BatchMaxSeq))) {
	    emit(i);
	    batchBases = 0;
	    batchSeqs = 0;
	    batchStart = int(i + 1);
	    nrBatches++;
	}
    }
    if (batchStart < int(n)) emit(n - 1);
    return 0;
}

// ---------------------------------------------------------------- info (main.cpp:394-422)
static void print_batch_info(const Batch& b)
{
    cerr << "\tBatch number: " << b.BatchNr << endl;
    cerr << "\tBatch range: [" << b.BatchStart << "," << b.BatchEnd << "]" << endl;
    cerr << "\tDepth: " << b.Depth << endl;
    cerr << "\tNr sequences: " << b.BatchEnd - b.BatchStart + 1 << endl;
    cerr << "\tNr bases: " << b.BatchBases << endl;
    cerr << "\tNr clusters: " << nr_clusters(b) << endl;
    cerr << "\tNr nontrivial clusters: " << nr_nontrivial_clusters(b) << endl;
    cerr << "\tMinimizers in database: " << int(b.MinDB.size()) << endl;
}

static std::unique_ptr<Batch> load_or_die(const std::string& path)
{
    try {
	return load_batch(path);
    }
    catch (const std::runtime_error& e) {  // serialize.cpp:19-24
	cerr << "Failed to load batch " << path << ":" << e.what() << endl;
	exit(1);
    }
}

static int main_info(int argc, char** argv)
{
    if (argc <= 2 || std::string(argv[2]) == "-h") {
	print_help_info();
	exit(0);
    }
    auto b = load_or_die(argv[2]);
    cerr << "Loaded batch from " << argv[2] << ":" << endl;
    print_batch_info(*b);
    return 0;
}

// ---------------------------------------------------------------- dump (mainDump main.cpp:204-236, dumpBatchInfo / dumpClusters
// main.cpp:414-457, WriteClusters output.cpp:151-275, ParseArgsDump args.cpp:260-311).  Host only: it reads the final
// `.cer` and the sorted FASTQ that `sort` wrote, no device work.
static std::string rev_comp_or_die(const std::string& s)  // RevComp util.cpp:13-38
{
    std::string o(s.rbegin(), s.rend());
    for (char& c : o) {
	switch (c) {
	    case 'A': c = 'T'; break;
	    case 'C': c = 'G'; break;
	    case 'G': c = 'C'; break;
	    case 'T': c = 'A'; break;
	    default:
		// the reference throws a std::string nobody catches -> terminate
		cerr << "terminate called after throwing an instance of 'std::string' (Invalid base encountered: " << c << ")" << endl;
		abort();
	}
    }
    return o;
}

static int main_dump(int argc, char** argv)
{
    static const struct option longopts[] = {
	{"verbose", no_argument, 0, 'v'},      {"debug", no_argument, 0, 'd'},	      {"help", no_argument, 0, 'h'},
	{"outdir", required_argument, 0, 'o'}, {"index", required_argument, 0, 'i'}, {0, 0, 0, 0},
    };
    std::string outDir, index;
    bool verbose = false;
    int sargc = argc - 1;
    char** sargv = argv + 1;
    int iarg = 0, li = 0;
    while (iarg != -1) {
	iarg = getopt_long(sargc, sargv, "dhvo:i:", longopts, &li);
	switch (iarg) {
	    case 'h': print_help_dump(); exit(0);
	    case 'o': outDir = optarg; break;
	    case 'i': index = optarg; break;
	    case 'v': verbose = true; break;
	    case 'd': break;
	}
    }
    if (outDir.empty()) {
	cerr << "Specifying output directory is mandatory!" << endl;
	print_help();
	exit(1);
    }
    if (optind >= sargc) {	// the reference dereferences a null pointer here
	cerr << "isONclust2 (B200): dump needs the clustered batch as positional argument" << endl;
	exit(1);
    }
    VERBOSE = verbose;
    const std::string inCereal = sargv[optind];
    if (VERBOSE) cerr << "Loading batch..." << endl;
    auto b = load_or_die(inCereal);
    if (VERBOSE) {
	cerr << "Loaded batch from " << inCereal << ":" << endl;
	print_batch_info(*b);
    }
    b->MinDB = MinimizerDB(0, UnsignedHash());
    b->ConsGs.clear();
    b->ConsGsValid.clear();
    create_outdir(outDir);
    {  // dumpBatchInfo, after the MinDB was dropped (main.cpp:217-219)
	FILE* f = create_file(outDir + "/batch_info.tsv");
	fprintf(f, "Name\tValue\nBatchNumber\t%d\nBatchStart\t%llu\nBatchEnd\t%llu\nDepth\t%d\nNrBases\t%llu\nNrClusters\t%d\nNrNontrivialCls\t%d\nMinDBsize\t%d\n",
		b->BatchNr, (unsigned long long)b->BatchStart, (unsigned long long)b->BatchEnd, b->Depth, (unsigned long long)b->BatchBases,
		nr_clusters(*b), nr_nontrivial_clusters(*b), int(b->MinDB.size()));
	fclose(f);
    }
    std::string sortedFastq;
    try {  // LoadIndex output.cpp:94-108
	Reader r(index);
	sortedFastq = r.str();
    }
    catch (const std::runtime_error& e) {
	cerr << "Failed to load index " << index << ":" << e.what() << endl;
	exit(1);
    }
    if (VERBOSE) {
	cerr << "Loaded index from " << index << ":" << endl;
	cerr << "Sorting clusters by size." << endl;
    }
    if (!b->Cls.empty()) {
	for (size_t i = 0; i < b->Cls.size(); i++) {
	    if (!b->Cls[i] || b->Cls[i]->empty() || !b->Cls[i]->at(0) || !b->Cls[i]->at(0)->raw) {
		cerr << "Null pointer instead of cluster rep sequence at index: " << i << endl;
		exit(1);
	    }
	}
	// SortClustersBySize cluster.cpp:570-580: std::sort (unstable) — the same algorithm on the same comparisons
	std::sort(b->Cls.begin(), b->Cls.end(), [](const std::shared_ptr<Cluster> x, const std::shared_ptr<Cluster> y) {
	    if (x->size() == y->size()) return x->at(0)->raw->score > y->at(0)->raw->score;
	    return x->size() > y->size();
	});
	// dumpClusters main.cpp:430-457
	struct IdInfo {
	    unsigned Cls;
	    int Strand;
	};
	std::unordered_map<std::string, IdInfo> idToCls;
	FILE* info = create_file(outDir + "/clusters_info.tsv");
	create_outdir(outDir + "/cluster_fastq");
	fputs("ClusterId\tSize\n", info);
	for (size_t i = 0; i < b->Cls.size(); i++) {
	    fprintf(info, "%zu\t%zu\n", i, b->Cls[i]->size() - 1);
	    for (auto& cc : *b->Cls[i]) idToCls[cc->id] = IdInfo{unsigned(i), cc->strand};
	}
	fclose(info);
	// WriteClusters output.cpp:151-275
	FILE* infq = fopen(sortedFastq.c_str(), "rb");
	if (!infq) {
	    cerr << "Failed to open " + sortedFastq + "!" << endl;
	    exit(1);
	}
	FILE* tsv = create_file(outDir + "/clusters.tsv");
	FILE* cons = create_file(outDir + "/cluster_cons.fq");
	fputs("ClusterId\tStrand\tRead\n", tsv);
	if (VERBOSE) cerr << "Writing out cluster information:" << endl;
	for (size_t i = 0; i < b->Cls.size(); i++) {
	    const ProcSeq& rep = *b->Cls[i]->at(0);
	    const SeqRec& s = *rep.raw;
	    if (s.score < 0) continue;
	    std::string seq = s.seq;
	    if (rep.strand == -1) seq = rev_comp_or_die(seq);
	    // the quality line is the stored one even on the minus strand (the reference's FIXME, output.cpp:200)
	    fprintf(cons, "@cluster_%zu origin=%s:%d length=%zu size=%zu\n", i, s.name.c_str(), rep.strand, seq.size(), b->Cls[i]->size() - 1);
	    fwrite(seq.data(), 1, seq.size(), cons);
	    fputs("\n+\n", cons);
	    fwrite(s.qual.data(), 1, s.qual.size(), cons);
	    fputc('\n', cons);
	}
	fclose(cons);
	if (VERBOSE) cerr << endl;
	b->Cls = Clusters();
	if (VERBOSE) cerr << "Loading fastq records:" << endl;
	// cluster id -> its records in file order; the reference keeps them in an unordered_map and writes one file per key
	std::unordered_map<unsigned, std::vector<std::string>> seqCache;
	{
	    char* line = nullptr;
	    size_t cap = 0;
	    std::string rec[4];
	    for (;;) {
		// GetRecords output.cpp:128-141: four getline calls; a header read that hits EOF ends the file
		ssize_t n = getline(&line, &cap, infq);
		if (n <= 0) break;
		const bool hadNl = line[n - 1] == '\n';
		rec[0].assign(line, size_t(hadNl ? n - 1 : n));
		if (!hadNl) break;  // std::getline sets eofbit when no newline ends the last line
		for (int q = 1; q < 4; q++) {
		    n = getline(&line, &cap, infq);
		    if (n <= 0) rec[q].clear();
		    else rec[q].assign(line, size_t(line[n - 1] == '\n' ? n - 1 : n));
		}
		const std::string readId = rec[0].empty() ? std::string() : rec[0].substr(1);
		auto v = idToCls.find(readId);
		if (v == idToCls.end()) continue;
		if (v->second.Strand == -1) {
		    rec[1] = rev_comp_or_die(rec[1]);
		    std::reverse(rec[3].begin(), rec[3].end());
		}
		fprintf(tsv, "%u\t%d\t%s\n", v->second.Cls, v->second.Strand, readId.c_str());
		std::string out;
		out.reserve(rec[0].size() + rec[1].size() + rec[2].size() + rec[3].size() + 4);
		for (int q = 0; q < 4; q++) {
		    out += rec[q];
		    out += '\n';
		}
		seqCache[v->second.Cls].push_back(std::move(out));
	    }
	    free(line);
	}
	fclose(infq);
	fclose(tsv);
	if (VERBOSE) cerr << endl << "Writing out cluster fastqs:" << endl;
	for (auto& c : seqCache) {
	    FILE* f = fopen((outDir + "/cluster_fastq/" + std::to_string(c.first) + ".fq").c_str(), "wb");
	    if (!f) {
		cerr << "Failed to open " + outDir + "/cluster_fastq/" + std::to_string(c.first) + ".fq!" << endl;
		exit(1);
	    }
	    for (auto& r : c.second) fwrite(r.data(), 1, r.size(), f);
	    fclose(f);
	}
	if (VERBOSE) cerr << endl;
    }
    if (VERBOSE) cerr << "Dump complete." << endl;
    return 0;
}

// ---------------------------------------------------------------- p_shared grid (InitMinSharedMap, p_emp_prob.cpp:22-47)
static std::string exe_dir()
{
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    std::string p(buf);
    size_t s = p.rfind('/');
    return s == std::string::npos ? "." : p.substr(0, s);
}

static void load_p_shared(int k, int w, double out[225])
{
    for (int i = 0; i < 225; i++) out[i] = std::numeric_limits<double>::quiet_NaN();
    const char* env = getenv("ICB_PEMP_TABLE");
    std::string path = env ? env : exe_dir() + "/../data/p_emp_prob.bin";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
	cerr << "isONclust2 (B200): cannot open the empirical probability table " << path << " (set ICB_PEMP_TABLE)" << endl;
	exit(1);
    }
    char magic[8];
    uint32_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "ICBPEMP1", 8) != 0 || fread(&n, 4, 1, f) != 1) {
	cerr << "isONclust2 (B200): bad table file " << path << endl;
	exit(1);
    }
    struct __attribute__((packed)) Row {
	uint8_t k, w, c1, c2;
	double p;
    };
    std::vector<Row> rows(n);
    if (fread(rows.data(), sizeof(Row), n, f) != n) {
	cerr << "isONclust2 (B200): truncated table file " << path << endl;
	exit(1);
    }
    fclose(f);
    for (const Row& r : rows) {	 // file order: later rows overwrite
	if (int(r.k) != k || std::abs(int(r.w) - w) > 2) continue;
	if (r.c1 < 1 || r.c1 > 15 || r.c2 < 1 || r.c2 > 15) continue;
	out[(r.c1 - 1) * 15 + (r.c2 - 1)] = r.p;
	out[(r.c2 - 1) * 15 + (r.c1 - 1)] = r.p;
    }
}

// ---------------------------------------------------------------- cluster (main.cpp:238-392 + cluster.cpp:67-322)
// ---------------------------------------------------------------- cluster with consensus (`sort -c <positive>`)
// cluster.cpp:263-309 + consensus.cpp:34-137.  The partial-order alignment and the heaviest-bundle consensus stay on the host
// (icb_poa_*, bit-exact to the reference's spoa); what a new consensus means for the clustering runs on the device: the
// consensus is sketched (icb_sketch_reads on a one-read set), appended to the read set with the reference's weighted error
// rates (icb_readset_append_from) and made the cluster's representative — UpdateMinDB on the device index and the
// sequence later reads are aligned with (icb_clusterer_update_representative).
// A representative that changes under read t invalidates every decision taken for later reads, so the device clusters
// the right reads in sub-batches that END where the first change can happen: a join only yields a consensus once the
// cluster's graph holds consMinSize sequences, so `need` = consMinSize - sequences joins are safe for every cluster.
struct ConsensusRun {
    Batch& left;
    Batch& right;
    const std::vector<size_t>& eligible;
    size_t nl;
    icb_ctx* ctx;
    icb_readset* rs;
    icb_clusterer* cl;
    const int32_t* ids;
    const CmdArgs& args;
    int algo;
    bool seqPurge;
    std::vector<icb_poa_graph*> graphs;	 // live graph per left cluster (parsed from / written back to left.ConsGs)
    std::vector<int32_t> nseq;		 // sequences in the cluster's graph, -1 = not known yet (blob not parsed)
    std::vector<uint8_t> blocked;	 // leaf: the cluster outgrew ConsPeriod (cluster.cpp:267-271), for good
    unsigned consInvoked;
    double qualTab[129], qualTabNomin[129];
};

static icb_poa_graph* cons_graph(ConsensusRun& R, size_t c)
{
    if (R.graphs.size() < R.left.ConsGs.size()) R.graphs.resize(R.left.ConsGs.size(), nullptr);
    if (!R.graphs[c]) {
	if (c >= R.left.ConsGs.size() || !R.left.ConsGsValid[c]) {
	    cerr << "isONclust2 (B200): cluster " << c << " has no consensus graph in the left batch" << endl;
	    exit(1);
	}
	int rc = icb_poa_graph_load(R.left.ConsGs[c].data(), int64_t(R.left.ConsGs[c].size()), &R.graphs[c]);
	if (rc) {
	    cerr << "isONclust2 (B200): cannot read the consensus graph of cluster " << c << endl;
	    exit(1);
	}
    }
    return R.graphs[c];
}

static int64_t graph_sequences(icb_poa_graph* g)
{
    int64_t n = 0;
    icb_poa_graph_sizes(g, &n, nullptr, nullptr);
    return n;
}

// UpdateClusterConsensus (consensus.cpp:34-126) and what cluster.cpp:289-303 does with its result
static void consensus_step(ConsensusRun& R, size_t i, int best, int matchStrand, const std::string& readSeq, double readRawErr, double readHpcErr)
{
    const CmdArgs& args = R.args;
    Batch& left = R.left;
    Batch& right = R.right;
    Cluster& cls = *left.Cls[size_t(best)];
    const bool leaf = left.Depth == -1;
    if (leaf && args.ConsPeriod > 0 && int(cls.size()) > args.ConsPeriod) {  // cluster.cpp:267-271
	R.blocked[size_t(best)] = 1;
	return;
    }
    icb_poa_graph* g = cons_graph(R, size_t(best));
    const bool haveRight = right.ConsGs.size() > 0 && right.ConsGsValid[i];
    const std::string consName = "cons_" + std::to_string(left.BatchNr) + "_" + std::to_string(i);
    const int consMinSize = leaf ? args.ConsMinSize : 2;
    const int k = args.KmerSize, w = args.WindowSize;

    const size_t leftSize = size_t(graph_sequences(g));
    size_t rightSize = 1;
    if (haveRight) {
	icb_poa_graph* rg = nullptr;
	if (icb_poa_graph_load(right.ConsGs[i].data(), int64_t(right.ConsGs[i].size()), &rg)) {
	    cerr << "isONclust2 (B200): cannot read the consensus graph of right cluster " << i << endl;
	    exit(1);
	}
	rightSize = size_t(graph_sequences(rg));
	icb_poa_graph_destroy(rg);
    }
    // consensus.cpp:47-49 calls RevComp(rs) for a read matched on the reverse strand and DROPS the result (RevComp returns a
    // new string, util.cpp:13-38): the read goes into the graph as it is, after the same validity check of its bases
    if (matchStrand == -1) (void)rev_comp_or_die(readSeq);
    const std::string& rs = readSeq;
    ProcSeq& rep = *cls.at(0);
    const double hpcErr = (rep.hpc->err * double(leftSize) + readHpcErr * double(rightSize)) / double(leftSize + rightSize);
    const double rawErr = (rep.raw->err * double(leftSize) + readRawErr * double(rightSize)) / double(leftSize + rightSize);
    int rc = icb_poa_graph_add(g, rs.data(), int64_t(rs.size()), uint32_t(haveRight ? rightSize : 1), R.algo);
    if (rc) {
	cerr << "isONclust2 (B200): partial-order alignment failed [" << rc << "] for " << consName << endl;
	exit(1);
    }
    R.nseq[size_t(best)] = int32_t(leftSize + 1);
    auto drop_right_graph = [&]() {  // cluster.cpp:305-308
	if (right.ConsGs.size() > 0 && right.ConsGsValid[i]) {
	    right.ConsGsValid[i] = 0;
	    right.ConsGs[i].clear();
	}
    };
    if (int(leftSize + 1) < consMinSize) {
	drop_right_graph();
	return;
    }
    int64_t clen = 0;
    icb_poa_graph_consensus(g, nullptr, 0, &clen);
    std::string cons(size_t(clen), '\0');
    if (icb_poa_graph_consensus(g, &cons[0], clen, &clen)) {
	cerr << "isONclust2 (B200): consensus generation failed for " << consName << endl;
	exit(1);
    }
    if (!(cons.length() > unsigned(2 * k) || cons.length() >= unsigned(w))) {
	cerr << "isONclust2 (B200): the consensus " << consName << " is shorter than 2k and the window (" << cons.length() << " bases); not supported" << endl;
	exit(1);
    }
    const std::vector<Minimizer> oldMins = rep.mins;
    rep.raw->seq = cons;
    rep.raw->name = consName;
    rep.raw->err = rawErr;
    rep.raw->score = rawErr * double(cons.length());
    const char fixedQualRaw = std::to_string(int(-10 * log10(rawErr)) + 33)[0];	 // consensus.cpp:98-99: the first character
    rep.raw->qual = std::string(cons.length(), fixedQualRaw);

    // ---- the consensus on the device: sketch, append, new representative
    icb_readset* one = nullptr;
    const int64_t o2[2] = {0, int64_t(cons.size())};
    ICB_DO(R.ctx, icb_readset_create(R.ctx, cons.data(), rep.raw->qual.data(), o2, 1, &one));
    ICB_DO(R.ctx, icb_score_reads(one, k, R.qualTab, R.qualTabNomin));
    ICB_DO(R.ctx, icb_sketch_reads(one, k, w, R.qualTabNomin, nullptr));
    int32_t status = 0, hpcLen = 0;
    int64_t nf = 0, nr = 0, mo[2] = {0, 0}, ro[2] = {0, 0};
    ICB_DO(R.ctx, icb_sketch_sizes(one, &nf, &nr));
    std::string hs(cons.size() + 1, '\0'), hq(cons.size() + 1, '\0');
    std::vector<Minimizer> mins(size_t(nf) + 1), rev(size_t(nr) + 1);
    double devHpcErr = 0.0;
    ICB_DO(R.ctx, icb_sketch_download(one, &status, &hpcLen, &devHpcErr, &hs[0], &hq[0], reinterpret_cast<icb_minimizer*>(mins.data()), mo,
				      reinterpret_cast<icb_minimizer*>(rev.data()), ro));
    if (status != 1) {
	// consensus.cpp:110-116 keeps such a representative with score -1; it can never be matched again
	cerr << "isONclust2 (B200): the consensus " << consName << " is too short after homopolymer compression (" << hpcLen << " bases); not supported" << endl;
	exit(1);
    }
    hs.resize(size_t(hpcLen));
    hq.resize(size_t(hpcLen));
    mins.resize(size_t(nf));
    rev.resize(size_t(nr));
    std::unique_ptr<SeqRec> hpc(new SeqRec);  // HomopolymerCompressObj (hpc.cpp:34-62): name and qualities of the raw consensus
    hpc->name = consName;
    hpc->seq = hs;
    hpc->qual = hq;
    hpc->err = hpcErr;
    hpc->score = hpcErr * double(hs.length());
    rep.hpc = std::move(hpc);
    rep.mins = mins;
    rep.rev = rev;
    int64_t idx = -1;
    ICB_DO(R.ctx, icb_readset_append_from(R.rs, one, 0, rawErr, rep.raw->score, hpcErr, &idx));
    ICB_DO(R.ctx, icb_clusterer_update_representative(R.cl, int32_t(best), int32_t(idx)));
    icb_readset_destroy(one);
    R.consInvoked++;
    update_min_db(unsigned(best), oldMins, rep.mins, left.MinDB);

    if (int(leftSize + 1) > args.ConsMaxSize) {	 // ConsPurge (consensus.cpp:128-137)
	const uint32_t wgt = uint32_t(leftSize + 1);
	icb_poa_graph_clear(g);
	if (icb_poa_graph_add(g, cons.data(), int64_t(cons.size()), wgt, R.algo)) {
	    cerr << "isONclust2 (B200): cannot rebuild the consensus graph of " << consName << endl;
	    exit(1);
	}
	R.nseq[size_t(best)] = 1;
    }
    drop_right_graph();
}

static void cluster_with_consensus(ConsensusRun& R, std::vector<int32_t>& best, std::vector<int32_t>& strand)
{
    Batch& left = R.left;
    Batch& right = R.right;
    const CmdArgs& args = R.args;
    const bool leaf = left.Depth == -1;
    const int consMinSize = leaf ? args.ConsMinSize : 2;
    const int64_t ne = int64_t(R.eligible.size());
    R.graphs.assign(left.ConsGs.size(), nullptr);
    R.nseq.assign(left.Cls.size(), -1);
    R.blocked.assign(left.Cls.size(), 0);
    left.ConsGs.reserve(left.Cls.size() + size_t(ne));
    // joins that are safe for every cluster: a cluster whose graph size is unknown (a loaded left batch) may change at once
    auto safe_joins = [&]() -> int64_t {
	int64_t d = std::max(1, consMinSize - 1) + 1;  // a cluster founded inside the sub-batch: its founder + need joins
	for (size_t c = 0; c < R.nseq.size(); c++) {
	    if (R.blocked[c]) continue;
	    const int64_t need = R.nseq[c] < 0 ? 1 : std::max<int64_t>(1, int64_t(consMinSize) - R.nseq[c]);
	    if (need < d) d = need;
	    if (d == 1) break;
	}
	return d;
    };
    std::vector<int32_t> cluster_of, strand_of;
    int32_t ncls = int32_t(R.nl);
    int64_t pos = 0;
    while (pos < ne) {
	const int64_t nsub = std::min<int64_t>(std::min<int64_t>(safe_joins(), ne - pos), 4096);
	cluster_of.assign(size_t(nsub), -1);
	strand_of.assign(size_t(nsub), 0);
	ICB_DO(R.ctx, icb_cluster_reads(R.cl, R.ids + R.nl + pos, nsub, cluster_of.data(), strand_of.data()));
	for (int64_t t = 0; t < nsub; t++) {
	    const size_t i = R.eligible[size_t(pos + t)];
	    const int32_t c = cluster_of[size_t(t)];
	    if (c < 0) {
		cerr << "isONclust2 (B200): device and host disagree on whether " << right.Cls[i]->at(0)->id << " is clustered" << endl;
		exit(1);
	    }
	    const ProcSeq& r0 = *right.Cls[i]->at(0);
	    const std::string readSeq = r0.raw->seq;  // cluster.cpp:170-175
	    const double readRawErr = r0.raw->err, readHpcErr = r0.hpc->err;
	    if (c == ncls) {
		best[i] = -1;
		ncls++;
		apply_decision(left, right, i, -1, 0, R.seqPurge);
		R.nseq.push_back(1);
		R.blocked.push_back(0);
	    }
	    else {
		best[i] = c;
		strand[i] = strand_of[size_t(t)];
		apply_decision(left, right, i, c, strand[i], R.seqPurge);
		consensus_step(R, i, c, strand[i], readSeq, readRawErr, readHpcErr);
	    }
	}
	pos += nsub;
    }
    finish_decisions(left, right);
    for (size_t c = 0; c < R.graphs.size(); c++)
	if (R.graphs[c]) {
	    int64_t nb = 0;
	    icb_poa_graph_save(R.graphs[c], nullptr, 0, &nb);
	    left.ConsGs[c].resize(size_t(nb));
	    if (icb_poa_graph_save(R.graphs[c], &left.ConsGs[c][0], nb, &nb)) {
		cerr << "isONclust2 (B200): cannot serialize the consensus graph of cluster " << c << endl;
		exit(1);
	    }
	    icb_poa_graph_destroy(R.graphs[c]);
	    R.graphs[c] = nullptr;
	}
}

static int main_cluster(int argc, char** argv)
{
    CmdArgsCluster cmd = parse_args_cluster(argc, argv);
    VERBOSE = cmd.Verbose;
    const bool SINGLE = cmd.RightCereal == "";
    auto left = load_or_die(cmd.LeftCereal);
    if (VERBOSE) {
	cerr << "Loaded input batch from " << cmd.LeftCereal << ":" << endl;
	print_batch_info(*left);
    }
    std::unique_ptr<Batch> right;
    if (!SINGLE) {
	right = load_or_die(cmd.RightCereal);
	cerr << "Loaded input batch from " << cmd.RightCereal << ":" << endl;
	right->MinDB = MinimizerDB(0, UnsignedHash());
	print_batch_info(*right);
    }
    else {
	right = create_pseudo_batch(*left);
	if (VERBOSE) {
	    cerr << "Created pseudo-batch for single clustering:" << endl;
	    print_batch_info(*right);
	    cerr << "Resetting input clusters." << endl;
	}
	left->Cls.clear();
	if (left->Depth > 0) left->Depth = -left->Depth;
	left->NrCls = 0;
	left->MinDB = MinimizerDB(MIN_DB_RESERVE, UnsignedHash());
    }
    if (left->SortArgs.Mode != cmd.Mode) left->SortArgs.Mode = cmd.Mode;
    if (right->SortArgs.Mode != cmd.Mode) right->SortArgs.Mode = cmd.Mode;
    if (VERBOSE && left->SortArgs.ConsMaxSize > 0)  // main.cpp:304-320
	cerr << "Generating consensus using spoa algorithm: " << (cmd.SpoaAlgo == 0 ? "local" : cmd.SpoaAlgo == 1 ? "global" : cmd.SpoaAlgo == 2 ? "semi-global" : "")
	     << endl;
    if (cmd.MinClsSize > 0) left->SortArgs.MinClsSize = cmd.MinClsSize;
    if (VERBOSE) {
	cerr << "Clustering mode: ";
	switch (left->SortArgs.Mode) {
	    case None: cerr << "Invalid clustering mode: " << int(None) << endl; exit(1);
	    case Sahlin: cerr << "sahlin"; break;
	    case Fast: cerr << "fast"; break;
	    case Furious: cerr << "furious"; break;
	}
	cerr << endl;
    }
    // ---- ClusterSortedReads: the checks of cluster.cpp:70-95
    if (!same_sort_args(left->SortArgs, right->SortArgs)) {
	cerr << "The left and right batches have been sorted with different parameters! " << endl;
	cerr << "Refusing to carry on with clustering as results would not make sense! " << endl;
	exit(1);
    }
    const CmdArgs args = left->SortArgs;
    if (right->Depth > 0 && right->BatchStart != (left->BatchEnd + 1)) {
	cerr << "Trying to merge non-consecutive batches! Giving up!" << endl;
	exit(1);
    }
    if (left->Depth > 0 && right->Depth > left->Depth) {
	cerr << "The left input batch must have higher depth!" << endl;
	exit(1);
    }
    const bool CONS = args.ConsMaxSize > 0;  // cluster.cpp:263-265
    unsigned consInvoked = 0;
    if (!SINGLE && left->Cls.size() > 0 && left->MinDB.size() == 0) {
	// The reference queries the left batch's MinDB AS LOADED: a left batch written with -z (--min-purge) has none, its
	// clusters are invisible to the merge and every right read founds a new cluster.  This build seeds the device index from
	// the representatives instead, which would find them — a different output.  Refuse rather than differ.
	cerr << "isONclust2 (B200): the left batch has " << left->Cls.size() << " clusters but an empty minimizer database (written with "
	     << "-z / --min-purge?); merging into it is not supported — re-create it without -z." << endl;
	exit(1);
    }
    if (left->MinDB.size() == 0) left->MinDB = MinimizerDB(MIN_DB_RESERVE, UnsignedHash());
    right->MinDB = MinimizerDB(0, UnsignedHash());

    // ---- the reads the device sees: representatives of the existing (left) clusters, then the eligible right reads
    const size_t nl = left->Cls.size(), nright = right->Cls.size();
    unsigned sizeFiltered = 0;
    std::vector<int32_t> best(nright, -2), strand(nright, 0);
    std::vector<size_t> eligible;
    for (size_t i = 0; i < nright; i++)
	if (right_is_eligible(*left, *right, i, &sizeFiltered)) eligible.push_back(i);

    if (args.Mode == None || eligible.empty()) {
	// getBestCluster with Mode None returns (-1, 0) for every read (cluster.cpp:545-567): every eligible read founds a cluster
	for (size_t i : eligible) best[i] = -1;
    }
    else {
	std::string seq, qual;
	std::vector<int64_t> offs{0};
	auto push = [&](const SeqRec& s) {
	    seq += s.seq;
	    qual += s.qual;
	    offs.push_back(int64_t(seq.size()));
	};
	for (size_t c = 0; c < nl; c++) {
	    const auto& cl = left->Cls[c];
	    if (!cl || cl->empty() || !cl->at(0) || !cl->at(0)->raw || !cl->at(0)->hpc) {
		cerr << "isONclust2 (B200): left cluster " << c << " has no representative sequence" << endl;
		exit(1);
	    }
	    push(*cl->at(0)->raw);
	}
	for (size_t i : eligible) push(*right->Cls[i]->at(0)->raw);
	const int64_t n = int64_t(offs.size()) - 1;

	double qualTab[129], qualTabNomin[129];
	icb_init_qual_tab(qualTab, 0);
	icb_init_qual_tab(qualTabNomin, 1);
	icb_ctx* ctx = open_device();
	icb_readset* rs = nullptr;
	if (CONS) {
	    // room for the consensus representatives (one per join at most, cluster.cpp:263-303).  A consensus is about as long
	    // as its cluster's transcript, but reads matched on the reverse strand enter the graph unturned (consensus_step)
	    // and the semi-global alignment then hangs them on an end, so it can be several reads long: four times the right
	    // batch's bases in all, none longer than eight times the longest read
	    int64_t maxLen = 0, rightBases = 0;
	    for (int64_t r = 0; r < n; r++) maxLen = std::max(maxLen, offs[size_t(r) + 1] - offs[size_t(r)]);
	    rightBases = offs[size_t(n)] - offs[nl];
	    ICB_DO(ctx, icb_readset_create_reserve(ctx, seq.data(), qual.data(), offs.data(), n, int64_t(eligible.size()), 4 * rightBases + (1 << 20),
						   8 * maxLen + 1024, &rs));
	}
	else {
	    ICB_DO(ctx, icb_readset_create(ctx, seq.data(), qual.data(), offs.data(), n, &rs));
	}
	ICB_DO(ctx, icb_score_reads(rs, args.KmerSize, qualTab, qualTabNomin));
	ICB_DO(ctx, icb_sketch_reads(rs, args.KmerSize, args.WindowSize, qualTabNomin, nullptr));
	if (CONS) {
	    // A consensus representative carries error rates that are weighted means (consensus.cpp:57-63), not functions of
	    // its quality string: cluster with the values stored in the batches, as the reference does.
	    std::vector<double> rawErr(static_cast<size_t>(n)), rawScore(static_cast<size_t>(n)), hpcErr(static_cast<size_t>(n));
	    for (int64_t r = 0; r < n; r++) {
		const ProcSeq& p = size_t(r) < nl ? *left->Cls[size_t(r)]->at(0) : *right->Cls[eligible[size_t(r) - nl]]->at(0);
		rawErr[size_t(r)] = p.raw->err;
		rawScore[size_t(r)] = p.raw->score;
		hpcErr[size_t(r)] = p.hpc->err;
	    }
	    ICB_DO(ctx, icb_readset_set_errors(rs, rawErr.data(), rawScore.data(), hpcErr.data()));
	}
	{  // the sketches in the batch files must be the ones this k / w produce (they are what the reference would query)
	    const size_t un = static_cast<size_t>(n);
	    std::vector<int32_t> status(un, 0);
	    std::vector<int64_t> mo(un + 1, 0), ro(un + 1, 0);
	    int64_t nf = 0, nr = 0;
	    ICB_DO(ctx, icb_sketch_sizes(rs, &nf, &nr));
	    std::vector<Minimizer> mins(static_cast<size_t>(nf) + 1), rev(static_cast<size_t>(nr) + 1);
	    ICB_DO(ctx, icb_sketch_download(rs, status.data(), nullptr, nullptr, nullptr, nullptr, reinterpret_cast<icb_minimizer*>(mins.data()),
					    mo.data(), reinterpret_cast<icb_minimizer*>(rev.data()), ro.data()));
	    for (int64_t r = 0; r < n; r++) {
		const ProcSeq& p = size_t(r) < nl ? *left->Cls[size_t(r)]->at(0) : *right->Cls[eligible[size_t(r) - nl]]->at(0);
		const bool same = status[r] == 1 && size_t(mo[r + 1] - mo[r]) == p.mins.size() && size_t(ro[r + 1] - ro[r]) == p.rev.size() &&
				  (p.mins.empty() || memcmp(p.mins.data(), mins.data() + mo[r], p.mins.size() * sizeof(Minimizer)) == 0) &&
				  (p.rev.empty() || memcmp(p.rev.data(), rev.data() + ro[r], p.rev.size() * sizeof(Minimizer)) == 0);
		if (!same) {
		    cerr << "isONclust2 (B200): the stored minimizers of " << p.id << " are not those of k=" << args.KmerSize
			 << " w=" << args.WindowSize << " on its stored sequence; refusing to cluster an inconsistent batch." << endl;
		    exit(1);
		}
	    }
	}
	icb_cluster_params P;
	memset(&P, 0, sizeof(P));
	P.kmer_size = args.KmerSize;
	P.window_size = args.WindowSize;
	P.mode = args.Mode;
	P.min_shared = args.MinShared;
	P.mapped_threshold = args.MappedThreshold;
	P.aligned_threshold = args.AlignedThreshold;
	P.min_fraction = args.MinFraction;
	P.min_prob_no_hits = args.MinProbNoHits;
	P.min_qual = args.MinQual;
	load_p_shared(args.KmerSize, args.WindowSize, P.p_shared);
	const char* chunk = getenv("ICB_CHUNK");
	P.chunk_size = chunk ? atoi(chunk) : 0;
	icb_clusterer* cl = nullptr;
	{
	    int rc = icb_clusterer_create(rs, &P, &cl);
	    if (rc == ICB_ERR_NO_PEMP) {  // GetPMinShared throws std::invalid_argument (p_emp_prob.cpp:88), uncaught in the reference
		cerr << "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  Empirical probability lookup failure!" << endl;
		abort();
	    }
	    if (rc) die_icb(ctx, rc, "icb_clusterer_create");
	}
	std::vector<int32_t> ids(static_cast<size_t>(n), 0);
	for (int64_t r = 0; r < n; r++) ids[size_t(r)] = int32_t(r);
	if (nl) ICB_DO(ctx, icb_clusterer_seed(cl, ids.data(), int64_t(nl)));
	const int64_t ne = int64_t(eligible.size());
	std::vector<int32_t> cluster_of(static_cast<size_t>(ne), -1), strand_of(static_cast<size_t>(ne), 0);
	if (CONS) {
	    ConsensusRun run{*left, *right, eligible, nl, ctx, rs, cl, ids.data(), args, cmd.SpoaAlgo, cmd.SeqPurge, {}, {}, {}, 0};
	    memcpy(run.qualTab, qualTab, sizeof qualTab);
	    memcpy(run.qualTabNomin, qualTabNomin, sizeof qualTabNomin);
	    cluster_with_consensus(run, best, strand);
	    consInvoked = run.consInvoked;
	}
	else {
	    ICB_DO(ctx, icb_cluster_reads(cl, ids.data() + nl, ne, cluster_of.data(), strand_of.data()));
	}
	int32_t ncls = int32_t(nl);
	for (int64_t t = 0; !CONS && t < ne; t++) {
	    const int32_t c = cluster_of[size_t(t)];
	    if (c < 0) {
		cerr << "isONclust2 (B200): device and host disagree on whether " << right->Cls[eligible[size_t(t)]]->at(0)->id << " is clustered" << endl;
		exit(1);
	    }
	    if (c == ncls) {
		best[eligible[size_t(t)]] = -1;
		ncls++;
	    }
	    else {
		best[eligible[size_t(t)]] = c;
		strand[eligible[size_t(t)]] = strand_of[size_t(t)];
	    }
	}
	if (VERBOSE) {
	    icb_cluster_stats st;
	    icb_clusterer_stats(cl, &st);
	    cerr << "Finished clustering!" << endl;
	    cerr << "Alignment invocation count: " << st.aln_invoked << " (" << (nright ? 100.0 * double(st.aln_invoked) / double(nright) : 0.0) << "%)" << endl;
	    cerr << "Consensus invocation count: " << consInvoked << " (" << (consInvoked && nright ? double(consInvoked) / double(nright) * 100 : 0.0)
		 << "%)" << endl;
	}
	icb_clusterer_destroy(cl);
	icb_readset_destroy(rs);
	icb_ctx_destroy(ctx);
    }
    if (!(CONS && args.Mode != None && !eligible.empty())) apply_decisions(*left, *right, best, strand, cmd.SeqPurge);
    if (VERBOSE) {
	std::cout << std::endl;	 // cluster.cpp:312
	if (left->SortArgs.MinClsSize > 1)
	    cerr << "Filtered out " << sizeFiltered << " input clusters smaller than " << left->SortArgs.MinClsSize << "." << endl;
	unsigned count = 0;
	for (auto& c : left->Cls)
	    if (c->size() > 1) count++;
	cerr << "Number of clusters larger than 1: " << count << endl;
	cerr << "Output batch statistics:" << endl;
	print_batch_info(*left);
    }
    left->LeftLeaf = cmd.LeftCereal;
    left->RightLeaf = cmd.RightCereal;
    if (cmd.MinPurge) {
	cerr << "Purging minimizer database in output batch!" << endl;
	left->MinDB = MinimizerDB(0, UnsignedHash());
    }
    save_batch(*left, cmd.OutCereal);
    if (VERBOSE) cerr << "Output batch written to: " << cmd.OutCereal << endl;
    return 0;
}

// ---------------------------------------------------------------- tree: every batch of a sort + the whole merge tree in one process
// What the reference's pipeline runs as one `cluster` job per batch file and one `cluster -l A -r B` job per merge
// (README.md:105-117), on ONE device context: the reads of all batches are uploaded, scored and sketched once, every
// node's index and representatives stay resident between levels (icb_cluster_tree_traced), and the batches of the nodes —
// cluster lists in the reference's order, MinDB, consensus chains — are rebuilt on the host from the decisions with the same
// code that serves `cluster` (apply_decisions).  The root is written as the `.cer` the last merge job would write.
struct CmdArgsTree {
    bool Verbose = false, MinPurge = false, SeqPurge = false;
    int MinClsSize = -1, Workers = 0;
    std::string BatchDir, OutCereal;
    int Mode = None;
};

static void print_help_tree()
{
    cerr << "Usage: isONclust2 tree [flags] -o root.cer <directory with the isONbatch_<i>.cer files of `sort`>\n"
	 << "Clusters every sorted batch and merges them as a tree (what `cluster` does per file and per pair of files), in one process.\n"
	 << "\t-x --mode\t\tClustering mode: sahlin, fast, furious.\n"
	 << "\t-o --outfile\t\tOutput batch (the root of the tree).\n"
	 << "\t-F --min-cls-size\tMinimum cluster size of merged right batches.\n"
	 << "\t-w --workers\t\tBatches / merges in flight (default: library default).\n"
	 << "\t-z --min-purge\t\tPurge the minimizer database of the output batch.\n"
	 << "\t-j --keep-seq\t\tAs for cluster.\n"
	 << "\t-v --verbose\t\tVerbose output.\n";
}

static CmdArgsTree parse_args_tree(int argc, char** argv)
{
    static const struct option longopts[] = {
	{"verbose", no_argument, 0, 'v'},	   {"min-purge", no_argument, 0, 'z'},	 {"min-cls-size", required_argument, 0, 'F'},
	{"keep-seq", no_argument, 0, 'j'},	   {"mode", required_argument, 0, 'x'},	 {"help", no_argument, 0, 'h'},
	{"outfile", required_argument, 0, 'o'},	   {"workers", required_argument, 0, 'w'}, {0, 0, 0, 0},
    };
    CmdArgsTree res;
    argc--;
    argv++;
    int iarg = 0, index;
    while ((iarg = getopt_long(argc, argv, "hvo:x:zjF:w:", longopts, &index)) != -1) {
	switch (iarg) {
	    case 'h': print_help_tree(); exit(0);
	    case 'o': res.OutCereal = optarg; break;
	    case 'v': res.Verbose = true; break;
	    case 'z': res.MinPurge = true; break;
	    case 'j': res.SeqPurge = true; break;
	    case 'F': res.MinClsSize = atoi(optarg); break;
	    case 'w': res.Workers = atoi(optarg); break;
	    case 'x':
		if (!parse_mode(optarg, &res.Mode)) exit(1);
		break;
	    default: break;
	}
    }
    if (optind >= argc) {
	cerr << "The directory of the sorted batches is mandatory!" << endl;
	print_help_tree();
	exit(1);
    }
    res.BatchDir = argv[optind];
    if (res.OutCereal == "") {
	cerr << "Specifying output batch file is mandatory!" << endl;
	print_help_tree();
	exit(1);
    }
    return res;
}

static void fill_cluster_params(const CmdArgs& args, icb_cluster_params& P)
{
    memset(&P, 0, sizeof(P));
    P.kmer_size = args.KmerSize;
    P.window_size = args.WindowSize;
    P.mode = args.Mode;
    P.min_shared = args.MinShared;
    P.mapped_threshold = args.MappedThreshold;
    P.aligned_threshold = args.AlignedThreshold;
    P.min_fraction = args.MinFraction;
    P.min_prob_no_hits = args.MinProbNoHits;
    P.min_qual = args.MinQual;
    load_p_shared(args.KmerSize, args.WindowSize, P.p_shared);
    const char* chunk = getenv("ICB_CHUNK");
    P.chunk_size = chunk ? atoi(chunk) : 0;
}

// the per-batch preparation of main_cluster's single-batch case (main.cpp:286-303 of the reference)
static std::unique_ptr<Batch> reset_for_single(Batch& left, int mode, int min_cls_size)
{
    auto right = create_pseudo_batch(left);
    left.Cls.clear();
    if (left.Depth > 0) left.Depth = -left.Depth;
    left.NrCls = 0;
    left.MinDB = MinimizerDB(MIN_DB_RESERVE, UnsignedHash());
    if (left.SortArgs.Mode != mode) left.SortArgs.Mode = mode;
    if (right->SortArgs.Mode != mode) right->SortArgs.Mode = mode;
    if (min_cls_size > 0) left.SortArgs.MinClsSize = min_cls_size;
    right->MinDB = MinimizerDB(0, UnsignedHash());
    return right;
}

static int main_tree(int argc, char** argv)
{
    CmdArgsTree cmd = parse_args_tree(argc, argv);
    VERBOSE = cmd.Verbose;
    // ---- the sorted batches, in order
    std::vector<std::unique_ptr<Batch>> nodes;
    std::vector<std::string> names;
    for (int b = 0;; b++) {
	const std::string path = cmd.BatchDir + "/isONbatch_" + std::to_string(b) + ".cer";
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) break;
	fclose(f);
	nodes.push_back(load_or_die(path));
	names.push_back(path);
    }
    const int nb = int(nodes.size());
    if (nb == 0) {
	cerr << "No sorted batches (isONbatch_0.cer ...) in " << cmd.BatchDir << endl;
	return 1;
    }
    for (int b = 0; b < nb; b++) {
	Batch& B = *nodes[size_t(b)];
	if (B.SortArgs.ConsMaxSize > 0) {
	    cerr << "isONclust2 (B200): consensus generation (-c " << B.SortArgs.ConsMaxSize
		 << ") is not part of this build; sort with the default -c (disabled)." << endl;
	    return 1;
	}
	CmdArgs a = B.SortArgs, z = nodes[0]->SortArgs;
	a.Mode = z.Mode = cmd.Mode;
	if (!same_sort_args(a, z)) {
	    cerr << "The batches have been sorted with different parameters! " << endl;
	    return 1;
	}
    }
    // ---- leaves: what `cluster -l isONbatch_b.cer` prepares; the reads the device sees are the eligible ones of every batch
    std::vector<std::unique_ptr<Batch>> rights(static_cast<size_t>(nb));
    std::vector<std::vector<size_t>> eligible(static_cast<size_t>(nb));
    std::string seq, qual;
    std::vector<int64_t> offs{0}, bounds{0};
    for (int b = 0; b < nb; b++) {
	Batch& L = *nodes[size_t(b)];
	rights[size_t(b)] = reset_for_single(L, cmd.Mode, cmd.MinClsSize);
	Batch& R = *rights[size_t(b)];
	unsigned sf = 0;
	for (size_t i = 0; i < R.Cls.size(); i++)
	    if (right_is_eligible(L, R, i, &sf)) {
		eligible[size_t(b)].push_back(i);
		const SeqRec& s = *R.Cls[i]->at(0)->raw;
		seq += s.seq;
		qual += s.qual;
		offs.push_back(int64_t(seq.size()));
	    }
	bounds.push_back(int64_t(offs.size()) - 1);
    }
    const CmdArgs args = nodes[0]->SortArgs;
    const int64_t n = int64_t(offs.size()) - 1;
    const int min_cls = args.MinClsSize;
    std::vector<int32_t> leaf_c, leaf_s;
    struct Job {
	int32_t left, right;
	std::vector<int32_t> cls, strand;
    };
    std::vector<Job> jobs;
    if (args.Mode == None || n == 0) {
	cerr << "isONclust2 (B200): tree needs a clustering mode (-x) and at least one eligible read" << endl;
	return 1;
    }
    {
	double qualTab[129], qualTabNomin[129];
	icb_init_qual_tab(qualTab, 0);
	icb_init_qual_tab(qualTabNomin, 1);
	icb_ctx* ctx = open_device();
	icb_readset* rs = nullptr;
	ICB_DO(ctx, icb_readset_create(ctx, seq.data(), qual.data(), offs.data(), n, &rs));
	ICB_DO(ctx, icb_score_reads(rs, args.KmerSize, qualTab, qualTabNomin));
	ICB_DO(ctx, icb_sketch_reads(rs, args.KmerSize, args.WindowSize, qualTabNomin, nullptr));
	{  // the sketches in the batch files must be the ones this k / w produce
	    const size_t un = static_cast<size_t>(n);
	    std::vector<int32_t> status(un, 0);
	    std::vector<int64_t> mo(un + 1, 0), ro(un + 1, 0);
	    int64_t nf = 0, nr = 0;
	    ICB_DO(ctx, icb_sketch_sizes(rs, &nf, &nr));
	    std::vector<Minimizer> mins(static_cast<size_t>(nf) + 1), rev(static_cast<size_t>(nr) + 1);
	    ICB_DO(ctx, icb_sketch_download(rs, status.data(), nullptr, nullptr, nullptr, nullptr, reinterpret_cast<icb_minimizer*>(mins.data()),
					    mo.data(), reinterpret_cast<icb_minimizer*>(rev.data()), ro.data()));
	    int64_t r = 0;
	    for (int b = 0; b < nb; b++)
		for (size_t i : eligible[size_t(b)]) {
		    const ProcSeq& p = *rights[size_t(b)]->Cls[i]->at(0);
		    const bool same = status[r] == 1 && size_t(mo[r + 1] - mo[r]) == p.mins.size() && size_t(ro[r + 1] - ro[r]) == p.rev.size() &&
				      (p.mins.empty() || memcmp(p.mins.data(), mins.data() + mo[r], p.mins.size() * sizeof(Minimizer)) == 0) &&
				      (p.rev.empty() || memcmp(p.rev.data(), rev.data() + ro[r], p.rev.size() * sizeof(Minimizer)) == 0);
		    if (!same) {
			cerr << "isONclust2 (B200): the stored minimizers of " << p.id << " are not those of k=" << args.KmerSize
			     << " w=" << args.WindowSize << " on its stored sequence; refusing to cluster an inconsistent batch." << endl;
			exit(1);
		    }
		    r++;
		}
	}
	icb_cluster_params P;
	fill_cluster_params(args, P);
	std::vector<int32_t> order(static_cast<size_t>(n)), cluster_of(static_cast<size_t>(n), -1), strand_of(static_cast<size_t>(n), 0);
	for (int64_t r = 0; r < n; r++) order[size_t(r)] = int32_t(r);
	std::vector<int64_t> ncl(static_cast<size_t>(nb), 0);
	icb_cluster_stats st;
	icb_tree_trace* tr = nullptr;
	{
	    int rc = icb_cluster_tree_traced(rs, &P, order.data(), bounds.data(), nb, cmd.Workers, min_cls, cluster_of.data(), strand_of.data(),
					     ncl.data(), &st, &tr);
	    if (rc == ICB_ERR_NO_PEMP) {
		cerr << "terminate called after throwing an instance of 'std::invalid_argument'\n  what():  Empirical probability lookup failure!" << endl;
		abort();
	    }
	    if (rc) die_icb(ctx, rc, "icb_cluster_tree_traced");
	}
	const int32_t *lc = nullptr, *ls = nullptr;
	ICB_DO(ctx, icb_tree_trace_leaves(tr, &lc, &ls));
	leaf_c.assign(lc, lc + n);
	leaf_s.assign(ls, ls + n);
	for (int32_t j = 0; j < icb_tree_trace_jobs(tr); j++) {
	    Job J;
	    int64_t nr = 0;
	    const int32_t *jc = nullptr, *js = nullptr;
	    ICB_DO(ctx, icb_tree_trace_job(tr, j, &J.left, &J.right, &nr, &jc, &js));
	    J.cls.assign(jc, jc + nr);
	    J.strand.assign(js, js + nr);
	    jobs.push_back(std::move(J));
	}
	if (VERBOSE) {
	    cerr << "Finished clustering " << nb << " batches and " << jobs.size() << " merges!" << endl;
	    cerr << "Alignment invocation count: " << st.aln_invoked << endl;
	}
	icb_tree_trace_destroy(tr);
	icb_readset_destroy(rs);
	icb_ctx_destroy(ctx);
    }
    // A node the file-based tree writes and reads back is not the node it had in memory: the MinDB of a loaded batch is an
    // unordered_map refilled in file order, and its iteration order decides the byte order of the next file.  The round trip
    // is part of the reference's behaviour, so every rebuilt node takes it too (through a scratch file next to the output).
    const std::string scratch = cmd.OutCereal + ".node.tmp";
    auto round_trip = [&](std::unique_ptr<Batch>& node) {
	save_batch(*node, scratch);
	node = load_or_die(scratch);
	remove(scratch.c_str());
    };
    // ---- rebuild the leaves
    for (int b = 0; b < nb; b++) {
	Batch& L = *nodes[size_t(b)];
	Batch& R = *rights[size_t(b)];
	std::vector<int32_t> best(R.Cls.size(), -2), strand(R.Cls.size(), 0);
	int32_t ncls = 0;
	int64_t t = bounds[size_t(b)];
	for (size_t i : eligible[size_t(b)]) {
	    const int32_t c = leaf_c[size_t(t)];
	    if (c < 0) {
		cerr << "isONclust2 (B200): device and host disagree on whether " << R.Cls[i]->at(0)->id << " is clustered" << endl;
		exit(1);
	    }
	    if (c == ncls) {
		best[i] = -1;
		ncls++;
	    }
	    else {
		best[i] = c;
		strand[i] = leaf_s[size_t(t)];
	    }
	    t++;
	}
	apply_decisions(L, R, best, strand, cmd.SeqPurge);
	L.LeftLeaf = names[size_t(b)];
	L.RightLeaf = "";
	rights[size_t(b)].reset();
	if (nb > 1) round_trip(nodes[size_t(b)]);
    }
    // ---- replay the merges in the order a sequential driver runs them (level by level, left to right)
    for (const Job& J : jobs) {
	Batch& L = *nodes[size_t(J.left)];
	if (!nodes[size_t(J.right)]) {
	    cerr << "isONclust2 (B200): merge tree trace names a right node that was already merged" << endl;
	    exit(1);
	}
	Batch& R = *nodes[size_t(J.right)];
	R.MinDB = MinimizerDB(0, UnsignedHash());
	if (cmd.MinClsSize > 0) L.SortArgs.MinClsSize = cmd.MinClsSize;
	if (R.Depth > 0 && R.BatchStart != (L.BatchEnd + 1)) {
	    cerr << "Trying to merge non-consecutive batches! Giving up!" << endl;
	    exit(1);
	}
	if (L.Depth > 0 && R.Depth > L.Depth) {
	    cerr << "The left input batch must have higher depth!" << endl;
	    exit(1);
	}
	if (J.cls.size() != R.Cls.size()) {
	    cerr << "isONclust2 (B200): merge of batches " << J.left << " and " << J.right << ": " << J.cls.size() << " decisions for " << R.Cls.size()
		 << " clusters" << endl;
	    exit(1);
	}
	unsigned sf = 0;
	std::vector<int32_t> best(R.Cls.size(), -2), strand(R.Cls.size(), 0);
	int32_t ncls = int32_t(L.Cls.size());
	for (size_t i = 0; i < R.Cls.size(); i++) {
	    const bool ok = right_is_eligible(L, R, i, &sf);
	    const int32_t c = J.cls[i];
	    if (ok != (c >= 0)) {
		cerr << "isONclust2 (B200): device and host disagree on whether cluster " << i << " of the right batch takes part in the merge" << endl;
		exit(1);
	    }
	    if (!ok) continue;
	    if (c == ncls) {
		best[i] = -1;
		ncls++;
	    }
	    else {
		best[i] = c;
		strand[i] = J.strand[i];
	    }
	}
	apply_decisions(L, R, best, strand, cmd.SeqPurge);
	L.LeftLeaf = "tree:" + std::to_string(J.left);
	L.RightLeaf = "tree:" + std::to_string(J.right);
	nodes[size_t(J.right)].reset();
	if (&J != &jobs.back()) round_trip(nodes[size_t(J.left)]);
    }
    Batch& root = *nodes[0];
    for (int b = 1; b < nb; b++)
	if (nodes[size_t(b)]) {
	    cerr << "isONclust2 (B200): batch " << b << " was never merged" << endl;
	    exit(1);
	}
    if (VERBOSE) {
	cerr << "Output batch statistics:" << endl;
	print_batch_info(root);
    }
    if (cmd.MinPurge) {
	cerr << "Purging minimizer database in output batch!" << endl;
	root.MinDB = MinimizerDB(0, UnsignedHash());
    }
    save_batch(root, cmd.OutCereal);
    if (VERBOSE) cerr << "Output batch written to: " << cmd.OutCereal << endl;
    return 0;
}

int main(int argc, char** argv)
{
    if (argc == 1) {
	print_help();
	exit(0);
    }
    const std::string sub = argv[1];
    try {
	if (sub == "sort") return main_sort(argc, argv);
	if (sub == "cluster") return main_cluster(argc, argv);
	if (sub == "tree") return main_tree(argc, argv);
	if (sub == "info") return main_info(argc, argv);
	if (sub == "version") {
	    std::cout << "isONclust2 version: " << ICB_VERSION << endl;
	    return 0;
	}
	if (sub == "help" || sub == "-h" || sub == "--help") {
	    print_help();
	    return 0;
	}
	if (sub == "dump") return main_dump(argc, argv);
    }
    catch (const std::exception& e) {
	cerr << "isONclust2 (B200): " << e.what() << endl;
	return 1;
    }
    cerr << "Invalid subcommand: " << sub << endl;
    cerr << "Valid subcommands are: sort, cluster, dump, info, version, help" << endl;
    cerr << "Inkove \"isONclust2 help\" for more details." << endl;
    return 1;
}
