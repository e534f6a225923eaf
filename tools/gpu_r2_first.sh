#!/bin/bash
# First GPU session of round 2 (prepared at the end of round 1, when the GPU budget was spent): device parity of the
# v3s driver (verified on the host emulator only so far), then its A/B against v1 on the scenes without SDFs, and the
# 1 -> 8 GPU scaling with the v2s defaults if 8 GPUs are given.   usage: gpurun --timeout 600 -- 'bash tools/gpu_r2_first.sh'
O=gpurun_out/r2_first; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v3s or v3_driver" > $O/pytest_v3s.log 2>&1; echo "pytest rc $?" >> $O/pytest_v3s.log
tail -4 $O/pytest_v3s.log
B="--warmup 3 --no-cpu-baseline --steps 8"
for wl in cfg2_scene1_1080p cfg1_scene0_512 bvh_spheres169_1080p; do
  PT_SCHED=0 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v1.json 2> $O/${wl}_v1.err
  for T in 12 16 20 24; do
    PT_SCHED=7 PT_REGEN_T=$T timeout 300 python bench.py --workload $wl $B > $O/${wl}_v3s_T$T.json 2> $O/${wl}_v3s_T$T.err
  done
  PT_SCHED=7 PT_MIN_BLOCKS=5 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v3s_T16_mb5.json 2> $O/${wl}_v3s_mb5.err
done
# fast-mode statistical gate with v3s as the driver of the run-time compiled kernels
PT_SCHED=7 PT_JIT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fast_mode_statistically or full_size_properties_cfg2" > $O/pytest_v3s_fast.log 2>&1; echo "pytest rc $?" >> $O/pytest_v3s_fast.log
tail -3 $O/pytest_v3s_fast.log
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
