#!/bin/bash
# SDF-phase scheduling: population histogram, PT_SDF_MIN x PT_SDF_EXIT sweep on the SDF workloads (v2 driver).
O=gpurun_out/sdfsched; mkdir -p $O
timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_base.log 2>&1
PT_SDF_MIN=16 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_min16.log 2>&1
cat $O/stats_base.log $O/stats_min16.log
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  for m in 0 8 12 16 24; do for x in 0 4; do
    PT_SDF_MIN=$m PT_SDF_EXIT=$x timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_min${m}_exit${x}.json 2> $O/${wl}_min${m}_exit${x}.err
  done; done
  for m in 16 24; do
    PT_SDF_MIN=$m PT_FEED_T=16 timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_min${m}_T16.json 2> $O/${wl}_min${m}_T16.err
    PT_SDF_MIN=$m PT_SDF_REPS=32 timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_min${m}_R32.json 2> $O/${wl}_min${m}_R32.err
  done
done
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
