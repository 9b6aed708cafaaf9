#!/bin/bash
# Dev tool: one full evidence pass on the B200 box (run through gpurun from the repo root).
#   tools/gpu_round.sh <tag> [tests] [bench] [ref] [launches] [ncu_sketch] [ncu_align] [ncu_query] [ncu_order]
# (ncu passes launch the device loop's kernels one by one, ICB_NO_GRAPH: the same kernels, individually visible)
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
tag=${1:-run}; shift
what=${*:-tests bench ref launches ncu_sketch ncu_align}
O=gpurun_out; mkdir -p $O
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${tag}_smi.txt 2>&1
nproc >> $O/${tag}_smi.txt
if has tests; then
    timeout 900 python -m pytest tests -m gpu -x -q > $O/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/${tag}_pytest_gpu.log
    tail -3 $O/${tag}_pytest_gpu.log
fi
if has bench; then
    timeout 1200 python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench exit $?"
    tail -c 600 $O/${tag}_bench.json
fi
if has ref; then
    timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err; echo "ref exit $?"
    tail -c 400 $O/${tag}_bench_reference.json
fi
if has launches; then
    ICB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/${tag}_launches.csv \
        python bench.py --reads 20000 --steps 1 --warmup 3 --quick > $O/${tag}_launches_bench.log 2>&1
    echo "launches exit $?"
fi
if has ncu_sketch; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:icb_k_sketch_warp -s 1 -c 1 -f -o $O/${tag}_sketch \
        python tools/bench_sketch.py 200000 2 > $O/${tag}_ncu_sketch.log 2>&1
    ncu -i $O/${tag}_sketch.ncu-rep --page raw --csv > $O/${tag}_sketch_raw.csv 2>/dev/null
    echo "ncu sketch exit $?"
fi
if has ncu_align; then
    # the saturated launch (all pairs of 3000 reads in one call), bracketed by cudaProfilerStart/Stop inside the tool
    timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:icb_k_align16w?_dp -c 1 -f -o $O/${tag}_align16 \
        python tools/bench_align.py 3000 128 1 prof > $O/${tag}_ncu_align.log 2>&1
    ncu -i $O/${tag}_align16.ncu-rep --page raw --csv > $O/${tag}_align16_raw.csv 2>/dev/null
    echo "ncu align exit $?"
fi
if has ncu_order; then
    ICB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icb_k_dl_order -s 40 -c 3 -f -o $O/${tag}_order \
        python bench.py --reads 30000 --steps 1 --warmup 3 --quick > $O/${tag}_ncu_order.log 2>&1
    ncu -i $O/${tag}_order.ncu-rep --page raw --csv > $O/${tag}_order_raw.csv 2>/dev/null
    echo "ncu order exit $?"
fi
if has ncu_query; then
    ICB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icb_k_query -s 60 -c 4 -f -o $O/${tag}_query \
        python bench.py --reads 20000 --steps 1 --warmup 3 --quick > $O/${tag}_ncu_query.log 2>&1
    ncu -i $O/${tag}_query.ncu-rep --page raw --csv > $O/${tag}_query_raw.csv 2>/dev/null
    echo "ncu query exit $?"
fi
ls -la $O | tail -30
