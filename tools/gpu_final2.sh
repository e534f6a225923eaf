#!/bin/bash
# Final measurement set, second edition (after BVH, pt_multi, the staging ring): whole GPU suite, smoke, bench lines of
# every workload with roofline + CPU baseline, the reference arm, launch list of the default bench.
O=gpurun_out/final2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_cfg2_reference.json 2> $O/bench.err
python bench.py > $O/bench_cfg2_scene1_1080p.json 2>> $O/bench.err
for cfg in "cfg1_scene0_512 --spf 64 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 16" "cfg4a_scene10_menger_1080p_pl32 --steps 8" "cfg4b_scene8_terrain_1080p_pl32 --steps 8" "cfg5_scene10_4k --steps 8" "bvh_spheres169_1080p --steps 16" "bvh_mixed74_1080p --steps 16"; do
  set -- $cfg
  python bench.py --workload $cfg > $O/bench_$1.json 2>> $O/bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
nproc > $O/host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/host.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    t=open(sys.argv[1]).read(); d=json.loads(t)
    print(sys.argv[1].split('/')[-1], 'lines', t.count(chr(10)), '%.3f Gs/s'%(d['value']/1e9), 'e2e %.3f'%(d['e2e']['value']/1e9), 'frac', round(d.get('roofline',{}).get('frac') or 0,3), 'cpu %.2f Ms/s'%((d.get('cpu_baseline',{}).get('value') or 0)/1e6), d.get('clocks'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
