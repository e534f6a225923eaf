#!/bin/bash
O=gpurun_out/v2d_ncu; mkdir -p $O gpurun_out/jd_v2d
PT_SCHED=4 PT_JIT_DUMP=gpurun_out/jd_v2d timeout 600 ncu --set full --import-source on --clock-control none -k regex:pt_render -s 2 -c 1 -f -o $O/prof_v2d_cfg3 python bench.py --workload cfg3_scene9_mandelbulb_1080p --steps 1 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
cp gpurun_out/jd_v2d/pt_render_jit_0.cubin $O/v2d_cfg3.cubin
ls -la $O
