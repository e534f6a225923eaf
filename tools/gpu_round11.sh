#!/bin/bash
mkdir -p gpurun_out/jd_cfg2 gpurun_out/jd_cfg3
B="python bench.py --no-cpu-baseline --warmup 3 --steps 1"
PT_JIT_DUMP=gpurun_out/jd_cfg2 ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_final_cfg2 $B > gpurun_out/ncu_f.log 2>&1
PT_JIT_DUMP=gpurun_out/jd_cfg3 ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_final_cfg3 $B --workload cfg3_scene9_mandelbulb_1080p >> gpurun_out/ncu_f.log 2>&1
for R in 16 32 64; do PT_SDF_REPS=$R python bench.py --no-cpu-baseline --warmup 2 --steps 3 --workload cfg4a_scene10_menger_1080p_pl32 > gpurun_out/h_menger_R$R.json 2>/dev/null; PT_SDF_REPS=$R python bench.py --no-cpu-baseline --warmup 2 --steps 3 --workload cfg4b_scene8_terrain_1080p_pl32 > gpurun_out/h_terrain_R$R.json 2>/dev/null; PT_SDF_REPS=$R python bench.py --no-cpu-baseline --warmup 2 --steps 3 --workload cfg3_scene9_mandelbulb_1080p > gpurun_out/h_mandel_R$R.json 2>/dev/null; done
