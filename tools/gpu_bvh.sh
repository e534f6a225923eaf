#!/bin/bash
# BVH round: parity tests, scan-vs-tree bench lines, crossover sweep.   gpurun -- bash tools/gpu_bvh.sh
O=gpurun_out/bvh; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bvh or capacity" > $O/pytest_bvh.log 2>&1; echo "pytest rc $?" >> $O/pytest_bvh.log
tail -5 $O/pytest_bvh.log
for wl in bvh_spheres169_1080p bvh_mixed74_1080p; do
  for bm in 0 12; do
    timeout 300 python bench.py --workload $wl --bvh-min $bm --steps 8 --warmup 3 --no-cpu-baseline > $O/${wl}_bm${bm}.json 2> $O/${wl}_bm${bm}.err
  done
  PT_NO_UNROLL=1 timeout 300 python bench.py --workload $wl --bvh-min 0 --steps 8 --warmup 3 --no-cpu-baseline > $O/${wl}_bm0_rolled.json 2> $O/${wl}_bm0_rolled.err
  timeout 300 python bench.py --workload $wl --bvh-min 12 --jit 1 --steps 8 --warmup 3 --no-cpu-baseline > $O/${wl}_bm12_jit1.json 2> $O/${wl}_bm12_jit1.err
done
timeout 600 python tools/bvh_crossover.py > $O/crossover.jsonl 2> $O/crossover.err
cat $O/crossover.jsonl
for f in $O/bvh_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9), d['config']['closest_hit'], 'e2e %.3f'%(d['e2e']['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
