#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --warmup 2 --steps 4"
for wl in cfg2_scene1_1080p cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  PT_SCHED=1 $B --workload $wl > gpurun_out/q_${wl}_v2.json 2>/dev/null
done
