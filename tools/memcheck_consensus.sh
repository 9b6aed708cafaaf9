#!/bin/bash
# compute-sanitizer memcheck over one consensus-on `cluster` job of the CLI (tests/consensus_case.py's input): the appended
# representatives, the per-consensus one-read sketches and the index updates.  Output: gpurun_out/<tag>_memcheck_consensus.log
set -e
TAG=${1:-r02}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W"
python - <<PY
import sys; sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import consensus_case as cc
cc.write_input(".")
PY
B=$ROOT/isonclust2_b200/bin/isONclust2
$B sort -B 50000 -M 80 -c 12 -g 4 -o srt reads.fq > /dev/null 2>&1
mkdir -p $ROOT/gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 $B cluster -x sahlin -l srt/batches/isONbatch_0.cer -o c0.cer \
    > $ROOT/gpurun_out/${TAG}_memcheck_consensus.log 2>&1 || echo "memcheck exit $?" >> $ROOT/gpurun_out/${TAG}_memcheck_consensus.log
tail -4 $ROOT/gpurun_out/${TAG}_memcheck_consensus.log
python - <<PY
import sys, json; sys.path.insert(0, "$ROOT/tests")
import consensus_case as cc
g = json.load(open("$ROOT/tests/golden/consensus_v1.json"))["digests"]["c0.cer"]
print("c0.cer under memcheck equals the reference's:", cc.digest("c0.cer") == g)
PY
