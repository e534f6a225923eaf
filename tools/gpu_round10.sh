#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --warmup 2"
for cfg in "cfg2_scene1_1080p --steps 8" "cfg1_scene0_512 --spf 64 --steps 8" "cfg3_scene9_mandelbulb_1080p --steps 3" "cfg4a_scene10_menger_1080p_pl32 --steps 3" "cfg4b_scene8_terrain_1080p_pl32 --steps 3" "cfg5_scene10_4k --steps 2"; do
  set -- $cfg
  PT_SCHED=0 $B --workload $cfg > gpurun_out/g_$1_v1.json 2>/dev/null
  for T in 4 8 12 16 33; do for R in 4 8 16; do
    PT_SCHED=1 PT_FEED_T=$T PT_SDF_REPS=$R $B --workload $cfg > gpurun_out/g_$1_v2_T${T}_R${R}.json 2>/dev/null
  done; done
  for mb in 4 5; do PT_SCHED=1 PT_MIN_BLOCKS=$mb $B --workload $cfg > gpurun_out/g_$1_v2_mb$mb.json 2>/dev/null; done
done
