#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline"
run() { # name, env..., -- args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for T in 1 4 8 12 16 24 33; do
 for R in 2 4 8; do
  for wl in cfg3_scene9_mandelbulb_1080p cfg4b_scene8_terrain_1080p_pl32 cfg4a_scene10_menger_1080p_pl32; do
    PT_FEED_T=$T PT_SDF_REPS=$R $B --workload $wl --steps 3 --warmup 3 > gpurun_out/sweep_${wl}_T${T}_R${R}.json 2>/dev/null
  done
 done
done
for mb in 3 4 5 6; do
  PT_SCHED=0 PT_MIN_BLOCKS=$mb $B --steps 16 > gpurun_out/v1_cfg2_minblocks$mb.json 2>/dev/null
  PT_SCHED=1 PT_MIN_BLOCKS=$mb $B --steps 16 > gpurun_out/v2_cfg2_minblocks$mb.json 2>/dev/null
done
