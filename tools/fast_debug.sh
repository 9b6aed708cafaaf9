# finds the smallest fast-mode run that fails and repeats it under compute-sanitizer
for n in 60000 120000 200000; do
  if ! timeout 300 python bench.py --quick --steps 1 --warmup 3 --mode fast --reads $n > gpurun_out/fast_$n.log 2>&1; then
    echo "fails at $n"; tail -3 gpurun_out/fast_$n.log | cut -c1-300
    timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --quick --steps 1 --warmup 3 --mode fast --reads $n > gpurun_out/fast_sanitizer_$n.log 2>&1
    grep -n "Invalid\|at 0x\|by thread\|========= *in\|Saved host\|icb_k" gpurun_out/fast_sanitizer_$n.log | head -40
    break
  else echo "ok at $n"; tail -1 gpurun_out/fast_$n.log | cut -c1-200; fi
done
