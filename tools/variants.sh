# tuning runs of the device loop's knobs (quick bench lines; not the bench contract): variants.sh "ENV=.. [ENV=..] -- bench flags" ...
run() { echo "== $*"; envs=(); flags=(); seen=0; for a in "$@"; do if [ "$a" = "--" ]; then seen=1; elif [ $seen = 0 ]; then envs+=("$a"); else flags+=("$a"); fi; done
  env "${envs[@]}" timeout 400 python bench.py --steps 2 --warmup 3 --quick "${flags[@]}" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['cluster_stats']; print(round(d['value']), round(d['ms_per_step']), 'rounds',s['rounds'],'order',s['order_calls'],'pairs',s['pairs_aligned'],'cells',s['cells_aligned'],'launches',d['gpu_launches'], 'stops full/first/again', s['dl_stop_full'], s['dl_stop_pair_first'], s['dl_stop_pair_again'], d['result_digest'][:12])
"; }
for v in "$@"; do run $v; done
