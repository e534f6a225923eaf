#!/bin/bash
# driver v2s (in-warp sample stealing, PT_SCHED=5) and the by-value cyclide call: parity, then A/B against v2 / v1.
O=gpurun_out/steal; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "v2s or sdf_scenes or all_shots or v2d or bvh_strict" > $O/pytest_v2s.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2s.log
tail -8 $O/pytest_v2s.log
B="--steps 6 --warmup 3 --no-cpu-baseline"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  PT_SCHED=1 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2.json 2> $O/${wl}_v2.err
  PT_SCHED=5 PT_STEAL_S=16 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2s_S16.json 2> $O/${wl}_v2s_S16.err
  PT_SCHED=5 PT_STEAL_S=8 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2s_S8.json 2> $O/${wl}_v2s_S8.err
done
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32; do
  for TR in "16 16" "4 16" "8 32"; do set -- $TR
    PT_SCHED=5 PT_FEED_T=$1 PT_SDF_REPS=$2 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2s_S16_T$1_R$2.json 2> $O/${wl}_v2s_S16_T$1_R$2.err
  done
done
for wl in cfg2_scene1_1080p cfg1_scene0_512 bvh_mixed74_1080p; do
  timeout 300 python bench.py --workload $wl $B > $O/${wl}_default.json 2> $O/${wl}_default.err
  PT_SCHED=5 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2s_S16.json 2> $O/${wl}_v2s_S16.err
  PT_SCHED=1 timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2.json 2> $O/${wl}_v2.err
done
PT_SCHED=5 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_v2s.log 2>&1
PT_SCHED=1 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_v2.log 2>&1
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
