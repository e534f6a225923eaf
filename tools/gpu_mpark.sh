#!/bin/bash
# march parking (PT_MPARK=1) on top of v2s: parity (strict bit-exact), then A/B over stack size / batch threshold.
O=gpurun_out/mpark; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v2s" > $O/pytest_v2s.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2s.log
tail -6 $O/pytest_v2s.log
B="--warmup 3 --no-cpu-baseline --steps 4"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  timeout 300 python bench.py --workload $wl $B > $O/${wl}_v2s.json 2> $O/${wl}_v2s.err
  for CM in "24 12" "24 6" "24 20" "16 8" "32 16"; do set -- $CM
    mb=6; [ $1 -gt 24 ] && mb=5
    PT_MPARK=1 PT_MPARK_CAP=$1 PT_MPARK_MIN=$2 PT_MIN_BLOCKS=$mb timeout 300 python bench.py --workload $wl $B > $O/${wl}_mpark_cap$1_min$2.json 2> $O/${wl}_mpark_cap$1_min$2.err
  done
  PT_MPARK=1 PT_FEED_T=4 timeout 300 python bench.py --workload $wl $B > $O/${wl}_mpark_cap24_min12_T4.json 2> $O/${wl}_mpark_T4.err
  PT_MPARK=1 PT_FEED_T=12 timeout 300 python bench.py --workload $wl $B > $O/${wl}_mpark_cap24_min12_T12.json 2> $O/${wl}_mpark_T12.err
done
PT_MPARK=1 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_mpark.log 2>&1
timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_v2s.log 2>&1
cat $O/stats_mpark.log
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
