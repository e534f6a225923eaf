#!/bin/bash
mkdir -p gpurun_out/jitdump
B="python bench.py --no-cpu-baseline"
PT_SCHED=0 $B --steps 16 > gpurun_out/e_v1_cfg2_fastmath.json 2>/dev/null
for mb in 6 7 8; do
  PT_SCHED=0 PT_MIN_BLOCKS=$mb $B --steps 16 > gpurun_out/e_v1_cfg2_mb$mb.json 2>/dev/null
  PT_SCHED=0 PT_MIN_BLOCKS=$mb $B --workload cfg3_scene9_mandelbulb_1080p --steps 4 > gpurun_out/e_v1_cfg3_mb$mb.json 2>/dev/null
done
PT_SCHED=0 PT_NO_UNROLL=1 $B --steps 16 > gpurun_out/e_v1_cfg2_nounroll.json 2>/dev/null
PT_SCHED=0 PT_NO_UNROLL=1 PT_MIN_BLOCKS=6 $B --steps 16 > gpurun_out/e_v1_cfg2_nounroll_mb6.json 2>/dev/null
PT_SCHED=0 $B --workload cfg3_scene9_mandelbulb_1080p --steps 4 > gpurun_out/e_v1_cfg3.json 2>/dev/null
PT_SCHED=0 $B --workload cfg4a_scene10_menger_1080p_pl32 --steps 4 > gpurun_out/e_v1_cfg4a.json 2>/dev/null
PT_SCHED=0 $B --workload cfg4b_scene8_terrain_1080p_pl32 --steps 4 > gpurun_out/e_v1_cfg4b.json 2>/dev/null
PT_SCHED=1 PT_JIT_DUMP=gpurun_out/jitdump ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v2_cfg2 $B --steps 1 --warmup 3 > gpurun_out/ncu_v2.log 2>&1
ls gpurun_out/jitdump
