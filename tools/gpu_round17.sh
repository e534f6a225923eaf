#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --warmup 2 --steps 4"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg5_scene10_4k; do
  $B --workload $wl > gpurun_out/t_${wl}_base.json 2>/dev/null
  PT_NO_UNROLL=1 $B --workload $wl > gpurun_out/t_${wl}_nu.json 2>/dev/null
  PT_MIN_BLOCKS=7 $B --workload $wl > gpurun_out/t_${wl}_mb7.json 2>/dev/null
  PT_MIN_BLOCKS=5 $B --workload $wl > gpurun_out/t_${wl}_mb5.json 2>/dev/null
  PT_NO_UNROLL=1 PT_MIN_BLOCKS=7 $B --workload $wl > gpurun_out/t_${wl}_nu_mb7.json 2>/dev/null
done
