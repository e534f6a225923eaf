#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 32 > gpurun_out/v_cfg2.json 2>gpurun_out/v.err
python bench.py --no-cpu-baseline --steps 8 --workload cfg5_scene10_4k > gpurun_out/v_cfg5.json 2>>gpurun_out/v.err
tail -2 gpurun_out/v.err
