#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline"
for cfg in "cfg2_scene1_1080p --steps 32" "cfg1_scene0_512 --spf 64 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 8" "cfg4a_scene10_menger_1080p_pl32 --steps 8" "cfg4b_scene8_terrain_1080p_pl32 --steps 4"; do
  set -- $cfg
  $B --workload $cfg > gpurun_out/v2_$1.json 2> gpurun_out/v2_$1.err
  PT_SCHED=0 $B --workload $cfg > gpurun_out/v1_$1.json 2>> gpurun_out/v2_$1.err
done
for reps in 1 2 8; do PT_SDF_REPS=$reps $B --workload cfg3_scene9_mandelbulb_1080p --steps 8 > gpurun_out/v2_cfg3_reps$reps.json 2>/dev/null; done
M="smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,launch__registers_per_thread"
for wl in cfg2_scene1_1080p cfg3_scene9_mandelbulb_1080p cfg4b_scene8_terrain_1080p_pl32; do
ncu --metrics $M --clock-control none -k regex:pt_render -s 3 -c 1 --csv --log-file gpurun_out/lanes_v2_$wl.csv $B --workload $wl --steps 1 --warmup 3 > /dev/null 2>&1
done
ls gpurun_out
