#!/bin/bash
# driver v2d (two pixels per lane) vs v2 on the SDF workloads: parity, scheduling statistics, T x R sweep.
O=gpurun_out/v2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v2d" > $O/pytest_v2d.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2d.log
tail -6 $O/pytest_v2d.log
PT_SCHED=4 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_v2d.log 2>&1
cat $O/stats_v2d.log
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  PT_SCHED=1 timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_v2.json 2> $O/${wl}_v2.err
  for T in 4 8 16; do for R in 8 16 32; do
    PT_SCHED=4 PT_FEED_T=$T PT_SDF_REPS=$R timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_v2d_T${T}_R${R}.json 2> $O/${wl}_v2d_T${T}_R${R}.err
  done; done
  PT_SCHED=4 PT_MIN_BLOCKS=4 timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_v2d_mb4.json 2> $O/${wl}_v2d_mb4.err
done
for wl in cfg2_scene1_1080p bvh_mixed74_1080p; do
  PT_SCHED=4 timeout 300 python bench.py --workload $wl --steps 8 --warmup 2 --no-cpu-baseline > $O/${wl}_v2d.json 2> $O/${wl}_v2d.err
done
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
