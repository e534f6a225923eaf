#!/bin/bash
# Final measurement set of round 1 (v2s default for SDF / BVH-heavy scenes; steps of 64 spp), CPU baseline on every workload:
# smoke, whole GPU suite, bench lines of every workload, reference arm, launch list, ncu --set full of the v2s kernel (cfg3).
O=gpurun_out/final4; mkdir -p $O
CPU=${PT_FINAL_CPU-}
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_cfg2_reference.json 2> $O/bench.err
python bench.py > $O/bench_cfg2_scene1_1080p.json 2>> $O/bench.err
for cfg in "cfg1_scene0_512 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 8" "cfg4a_scene10_menger_1080p_pl32 --steps 6" "cfg4b_scene8_terrain_1080p_pl32 --steps 6" "cfg5_scene10_4k --steps 4" "bvh_spheres169_1080p --steps 8" "bvh_mixed74_1080p --steps 8"; do
  set -- $cfg
  python bench.py --workload $cfg $CPU > $O/bench_$1.json 2>> $O/bench.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
mkdir -p gpurun_out/jd_final4
PT_JIT_DUMP=gpurun_out/jd_final4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pt_render_jit -s 3 -c 1 -f -o $O/ncu_v2s_cfg3 python bench.py --workload cfg3_scene9_mandelbulb_1080p --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_cfg3.log 2>&1
ncu -i $O/ncu_v2s_cfg3.ncu-rep --page raw --csv > $O/ncu_full_v2s_cfg3_raw.csv 2>/dev/null
cp gpurun_out/jd_final4/pt_render_jit_0.cubin $O/cfg3_kernel.cubin 2>/dev/null
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
nproc > $O/host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/host.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    t=open(sys.argv[1]).read(); d=json.loads(t)
    print(sys.argv[1].split('/')[-1], 'lines', t.count(chr(10)), '%.3f Gs/s'%(d['value']/1e9), 'e2e %.3f'%(d['e2e']['value']/1e9), 'frac', round(d.get('roofline',{}).get('frac') or 0,3), 'cpu %.2f Ms/s'%(((d.get('cpu_baseline') or {}).get('value') or 0)/1e6), d.get('clocks'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
