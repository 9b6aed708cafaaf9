// Test tool for isonclust2_b200/cli/cer.{hpp,cpp} (host code only, no device):
//   cer_tool roundtrip <in.cer> <out.cer>
//       load_batch + save_batch: the output must equal the input byte for byte (sorted, clustered and merged batches)
//   cer_tool apply <left.cer> <right.cer|-> <decisions.txt> <out.cer> <mode> <left_name> <right_name> [keep_seq]
//       ClusterSortedReads' bookkeeping with the per-read decisions read from a text file ("best strand" per line for
//       every right cluster, best = -2 skipped / -1 new cluster / id joined) instead of the device: the output must
//       equal what the reference's `cluster` wrote for the same decisions
//   cer_tool eligible <left.cer> <right.cer|->
//       prints 0/1 per right cluster (the skips of cluster.cpp:116-160)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "cer.hpp"

using namespace icb;

static void prepare(std::unique_ptr<Batch>& left, std::unique_ptr<Batch>& right, const char* right_path)
{
    if (std::string(right_path) == "-") {  // main.cpp:256-275
	right = create_pseudo_batch(*left);
	left->Cls.clear();
	if (left->Depth > 0) left->Depth = -left->Depth;
	left->NrCls = 0;
	left->MinDB = MinimizerDB(MIN_DB_RESERVE, UnsignedHash());
    }
    else {
	right = load_batch(right_path);
	right->MinDB = MinimizerDB(0, UnsignedHash());
    }
    if (left->MinDB.size() == 0) left->MinDB = MinimizerDB(MIN_DB_RESERVE, UnsignedHash());
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    std::string cmd = argv[1];
    try {
	if (cmd == "roundtrip" && argc == 4) {
	    auto b = load_batch(argv[2]);
	    save_batch(*b, argv[3]);
	    return 0;
	}
	if (cmd == "eligible" && argc == 4) {
	    auto left = load_batch(argv[2]);
	    std::unique_ptr<Batch> right;
	    prepare(left, right, argv[3]);
	    unsigned sf = 0;
	    for (size_t i = 0; i < right->Cls.size(); i++) printf("%d\n", right_is_eligible(*left, *right, i, &sf) ? 1 : 0);
	    return 0;
	}
	if (cmd == "apply" && argc >= 9) {
	    auto left = load_batch(argv[2]);
	    std::unique_ptr<Batch> right;
	    prepare(left, right, argv[3]);
	    int mode = atoi(argv[6]);
	    left->SortArgs.Mode = mode;
	    right->SortArgs.Mode = mode;
	    std::vector<int32_t> best, strand;
	    std::ifstream in(argv[4]);
	    int b, s;
	    while (in >> b >> s) {
		best.push_back(b);
		strand.push_back(s);
	    }
	    if (best.size() != right->Cls.size()) {
		fprintf(stderr, "decisions: %zu lines for %zu right clusters\n", best.size(), right->Cls.size());
		return 1;
	    }
	    unsigned sf = 0;
	    for (size_t i = 0; i < right->Cls.size(); i++) {
		bool ok = right_is_eligible(*left, *right, i, &sf);
		if (ok != (best[i] != -2)) {
		    fprintf(stderr, "eligibility mismatch at %zu: host %d decisions %d\n", i, int(ok), best[i]);
		    return 1;
		}
	    }
	    apply_decisions(*left, *right, best, strand, argc > 9 && atoi(argv[9]) != 0);
	    left->LeftLeaf = argv[7];
	    left->RightLeaf = argv[8];
	    save_batch(*left, argv[5]);
	    return 0;
	}
    }
    catch (const std::exception& e) {
	fprintf(stderr, "cer_tool: %s\n", e.what());
	return 1;
    }
    fprintf(stderr, "usage: cer_tool roundtrip|apply|eligible ...\n");
    return 2;
}
