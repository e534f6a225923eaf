#!/bin/bash
# driver v3 (flat loop + gated regeneration) vs v1/v2 on the analytic workloads; BVH workloads with the FMA slab test.
O=gpurun_out/v3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v3 or bvh" > $O/pytest_v3.log 2>&1; echo "pytest rc $?" >> $O/pytest_v3.log
tail -4 $O/pytest_v3.log
for wl in cfg2_scene1_1080p "cfg1_scene0_512 --spf 64"; do
  set -- $wl
  PT_SCHED=0 timeout 300 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu-baseline > $O/${1}_v1.json 2> $O/${1}_v1.err
  PT_SCHED=1 timeout 300 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu-baseline > $O/${1}_v2.json 2> $O/${1}_v2.err
  for T in 1 4 8 16 24 32; do
    PT_SCHED=3 PT_REGEN_T=$T timeout 300 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu-baseline > $O/${1}_v3_T$T.json 2> $O/${1}_v3_T$T.err
  done
  for mb in 5 7; do
    PT_SCHED=3 PT_REGEN_T=16 PT_MIN_BLOCKS=$mb timeout 300 python bench.py --workload $wl --steps 16 --warmup 3 --no-cpu-baseline > $O/${1}_v3_T16_mb$mb.json 2> $O/${1}_v3_T16_mb$mb.err
  done
done
for wl in bvh_spheres169_1080p bvh_mixed74_1080p; do
  for sched in 0 1 3; do
    PT_SCHED=$sched timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > $O/${wl}_sched$sched.json 2> $O/${wl}_sched$sched.err
  done
done
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32; do
  PT_SCHED=3 timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_v3.json 2> $O/${wl}_v3.err
done
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9), 'e2e %.3f'%(d['e2e']['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
