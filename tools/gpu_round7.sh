#!/bin/bash
mkdir -p gpurun_out/jd_terrain gpurun_out/jd_menger gpurun_out/jd_mandel
B="python bench.py --no-cpu-baseline --warmup 3 --steps 1"
PT_JIT_DUMP=gpurun_out/jd_terrain ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v2_terrain $B --workload cfg4b_scene8_terrain_1080p_pl32 > gpurun_out/ncu_t.log 2>&1
PT_JIT_DUMP=gpurun_out/jd_menger ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v2_menger $B --workload cfg4a_scene10_menger_1080p_pl32 >> gpurun_out/ncu_t.log 2>&1
PT_JIT_DUMP=gpurun_out/jd_mandel ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v2_mandel $B --workload cfg3_scene9_mandelbulb_1080p >> gpurun_out/ncu_t.log 2>&1
python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > gpurun_out/sched_stats2.log 2>&1
