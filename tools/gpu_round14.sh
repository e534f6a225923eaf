#!/bin/bash
mkdir -p gpurun_out/jd_v2cfg2
PT_SCHED=1 PT_JIT_DUMP=gpurun_out/jd_v2cfg2 ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_v2_cfg2b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
PT_STATS=1 PT_SCHED=1 python tools/sched_stats.py cfg2_scene1_1080p > gpurun_out/sched_stats3.log 2>&1
