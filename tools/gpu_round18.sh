#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --warmup 2 --steps 8"
for wl in "cfg2_scene1_1080p" "cfg1_scene0_512 --spf 64"; do
  set -- $wl
  for mb in 6 7; do for nu in 0 1; do
    PT_MIN_BLOCKS=$mb PT_NO_UNROLL=$nu $B --workload $wl > gpurun_out/u_$1_mb${mb}_nu$nu.json 2>/dev/null
  done; done
done
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  python bench.py --no-cpu-baseline --warmup 2 --steps 4 --workload $wl > gpurun_out/u_${wl}_default.json 2>/dev/null
  PT_MIN_BLOCKS=8 python bench.py --no-cpu-baseline --warmup 2 --steps 4 --workload $wl > gpurun_out/u_${wl}_mb8.json 2>/dev/null
done
