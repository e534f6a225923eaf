#!/bin/bash
mkdir -p gpurun_out
PT_SCHED=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sdf_scenes or long_paths or smoke or fast_mode_statistically" 2>&1 | tail -5 > gpurun_out/pytest_v2p.log
cat gpurun_out/pytest_v2p.log
B="python bench.py --no-cpu-baseline --warmup 2 --steps 3"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  PT_SCHED=1 $B --workload $wl > gpurun_out/p_${wl}_v2.json 2>/dev/null
  for pm in 8 16 24 32; do for T in 8 16; do
    PT_SCHED=2 PT_POOL_MIN=$pm PT_FEED_T=$T timeout 300 $B --workload $wl > gpurun_out/p_${wl}_v2p_pm${pm}_T$T.json 2>/dev/null
  done; done
done
