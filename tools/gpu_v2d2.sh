#!/bin/bash
# v2d with batched swaps: PT_SWAP_MIN x PT_FEED_T on the SDF workloads.
O=gpurun_out/v2d2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v2d" > $O/pytest_v2d.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2d.log
tail -3 $O/pytest_v2d.log
PT_SCHED=4 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 > $O/stats_v2d.log 2>&1
cat $O/stats_v2d.log
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  for S in 4 8 16 24; do for T in 4 8; do
    PT_SCHED=4 PT_SWAP_MIN=$S PT_FEED_T=$T timeout 300 python bench.py --workload $wl --steps 4 --warmup 2 --no-cpu-baseline > $O/${wl}_S${S}_T${T}.json 2> $O/${wl}_S${S}_T${T}.err
  done; done
done
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
