#!/bin/bash
# BVH round, third pass: whole GPU suite, loop-shape x driver A/B, bench lines, crossover, ncu capture of the winner.
O=gpurun_out/bvh3; mkdir -p $O gpurun_out/jd_bvh
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_all.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu_all.log
tail -4 $O/pytest_gpu_all.log
for wl in bvh_spheres169_1080p bvh_mixed74_1080p; do
  for ww in 0 1; do for sched in 0 1; do
    PT_BVH_WHILE_WHILE=$ww PT_SCHED=$sched timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > $O/ab_${wl}_ww${ww}_sched${sched}.json 2> $O/ab_${wl}_ww${ww}_sched${sched}.err
  done; done
  PT_NO_UNROLL=1 timeout 300 python bench.py --workload $wl --bvh-min 0 --steps 8 --warmup 3 --no-cpu-baseline > $O/ab_${wl}_scan_rolled.json 2> $O/ab_${wl}_scan_rolled.err
done
for f in $O/ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9), d['config']['closest_hit'], 'e2e %.3f'%(d['e2e']['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
