#!/bin/bash
mkdir -p gpurun_out
python tools/rmse_vs_time.py cfg1_scene0_512 cfg2_scene1_1080p cfg3_scene9_mandelbulb_1080p > gpurun_out/rmse_vs_time.jsonl 2> gpurun_out/rmse.err
python tools/rmse_vs_time.py cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 --ref-spp 16384 --max-spp 1024 >> gpurun_out/rmse_vs_time.jsonl 2>> gpurun_out/rmse.err
tail -3 gpurun_out/rmse.err
