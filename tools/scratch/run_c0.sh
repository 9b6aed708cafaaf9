set -e
mkdir -p /tmp/cc && cd /tmp/cc
python - <<'PY'
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import consensus_case as cc
cc.write_input('.')
PY
B=/root/repo/isonclust2_b200/bin/isONclust2
$B sort -B 50000 -M 80 -c 12 -g 4 -o srt reads.fq > /dev/null 2>&1
$B cluster -v -x sahlin -l srt/batches/isONbatch_0.cer -o c0.cer > c0.log 2>&1 || true
tail -5 c0.log
cp c0.cer c0.log /root/repo/gpurun_out/
