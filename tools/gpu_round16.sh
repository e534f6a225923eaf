#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --warmup 2 --steps 4"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  $B --workload $wl > gpurun_out/s_${wl}_coop1.json 2>/dev/null
  PT_COOP_NORMALS=0 $B --workload $wl > gpurun_out/s_${wl}_coop0.json 2>/dev/null
  PT_SDF_REPS=8 $B --workload $wl > gpurun_out/s_${wl}_coop1_R8.json 2>/dev/null
  PT_SDF_REPS=32 $B --workload $wl > gpurun_out/s_${wl}_coop1_R32.json 2>/dev/null
done
