#!/bin/bash
# BVH round, second pass: parity with the far-origin inflation, bench lines (with CPU baseline), ncu capture.
O=gpurun_out/bvh2; mkdir -p $O gpurun_out/jd_bvh
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_all.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu_all.log
tail -4 $O/pytest_gpu_all.log
for wl in bvh_spheres169_1080p bvh_mixed74_1080p; do
  timeout 300 python bench.py --workload $wl --steps 16 --warmup 3 > $O/bench_${wl}.json 2> $O/bench_${wl}.err
  PT_NO_UNROLL=1 timeout 300 python bench.py --workload $wl --bvh-min 0 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_${wl}_scan_rolled.json 2> $O/bench_${wl}_scan_rolled.err
done
timeout 600 python tools/bvh_crossover.py > $O/crossover.jsonl 2> $O/crossover.err
PT_JIT_DUMP=gpurun_out/jd_bvh timeout 600 ncu --set full --import-source on --clock-control none -k regex:pt_render -s 3 -c 1 -f -o $O/prof_bvh_spheres169 python bench.py --workload bvh_spheres169_1080p --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cp gpurun_out/jd_bvh/pt_render_jit_0.cubin $O/bvh_spheres169_kernel.cubin 2>/dev/null
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9), d['config']['closest_hit'], 'e2e %.3f'%(d['e2e']['value']/1e9), 'frac', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
cat $O/crossover.jsonl
