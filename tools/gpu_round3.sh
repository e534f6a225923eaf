#!/bin/bash
mkdir -p gpurun_out
python tools/sched_stats.py cfg2_scene1_1080p cfg3_scene9_mandelbulb_1080p cfg4b_scene8_terrain_1080p_pl32 cfg1_scene0_512 > gpurun_out/sched_stats.log 2>&1
# source-level profile of the v1 driver, static fast kernel (has line info for real files)
PT_SCHED=0 ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v1_static_cfg2 python bench.py --jit 0 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v1_static.log 2>&1
ls gpurun_out
