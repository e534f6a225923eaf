#!/bin/bash
# One GPU session: tests, bench lines for every workload, launch list + one full ncu capture of the top kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_cfg2_reference.json 2>> gpurun_out/bench_cfg2.err
for cfg in "cfg1_scene0_512 --spf 64 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 16" "cfg4a_scene10_menger_1080p_pl32 --steps 16" "cfg4b_scene8_terrain_1080p_pl32 --steps 8"; do
  set -- $cfg
  python bench.py --workload $cfg > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
done
# variants on cfg2: generic static kernel, strict mode
python bench.py --jit 0 --steps 16 --no-cpu-baseline > gpurun_out/bench_cfg2_jit0.json 2>> gpurun_out/bench_cfg2.err
python bench.py --jit 2 --mode strict --steps 8 --no-cpu-baseline > gpurun_out/bench_cfg2_strict.json 2>> gpurun_out/bench_cfg2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pt_render -s 3 -c 1 -f -o gpurun_out/prof_cfg3 python bench.py --workload cfg3_scene9_mandelbulb_1080p --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
