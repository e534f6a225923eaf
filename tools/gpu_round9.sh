#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu-baseline --warmup 2"
for cfg in "cfg2_scene1_1080p --steps 8" "cfg1_scene0_512 --spf 64 --steps 8" "cfg3_scene9_mandelbulb_1080p --steps 4" "cfg4a_scene10_menger_1080p_pl32 --steps 3" "cfg4b_scene8_terrain_1080p_pl32 --steps 3" "cfg5_scene10_4k --steps 3"; do
  set -- $cfg
  $B --workload $cfg > gpurun_out/w_$1_mega.json 2>gpurun_out/w_$1.err
  $B --pipeline wavefront --workload $cfg > gpurun_out/w_$1_wf.json 2>>gpurun_out/w_$1.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_wf_cfg3.csv $B --pipeline wavefront --workload cfg3_scene9_mandelbulb_1080p --steps 1 --warmup 1 > /dev/null 2>&1
