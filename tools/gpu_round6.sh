#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --warmup 3"
for cfg in "cfg2_scene1_1080p --steps 16" "cfg1_scene0_512 --spf 64 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 4" "cfg4a_scene10_menger_1080p_pl32 --steps 4" "cfg4b_scene8_terrain_1080p_pl32 --steps 3"; do
  set -- $cfg
  PT_SCHED=0 PT_MIN_BLOCKS=6 $B --workload $cfg > gpurun_out/f_$1_v1_mb6.json 2>/dev/null
  PT_SCHED=0 PT_MIN_BLOCKS=6 PT_NO_UNROLL=1 $B --workload $cfg > gpurun_out/f_$1_v1_mb6_nu.json 2>/dev/null
  PT_SCHED=1 PT_MIN_BLOCKS=4 PT_FEED_T=33 PT_SDF_REPS=8 $B --workload $cfg > gpurun_out/f_$1_v2_mb4.json 2>/dev/null
  PT_SCHED=1 PT_MIN_BLOCKS=6 PT_FEED_T=33 PT_SDF_REPS=8 $B --workload $cfg > gpurun_out/f_$1_v2_mb6.json 2>/dev/null
  PT_SCHED=1 PT_MIN_BLOCKS=6 PT_FEED_T=33 PT_SDF_REPS=8 PT_NO_UNROLL=1 $B --workload $cfg > gpurun_out/f_$1_v2_mb6_nu.json 2>/dev/null
  PT_SCHED=1 PT_MIN_BLOCKS=5 PT_FEED_T=33 PT_SDF_REPS=16 PT_NO_UNROLL=1 $B --workload $cfg > gpurun_out/f_$1_v2_mb5_nu_r16.json 2>/dev/null
done
