#!/bin/bash
# march parking, second edition: a batch is popped only when PT_MPARK_MIN paths wait AND as many lanes are free.
O=gpurun_out/mpark2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "march_parking" > $O/pytest_mpark.log 2>&1; echo "pytest rc $?" >> $O/pytest_mpark.log
tail -4 $O/pytest_mpark.log
B="--warmup 2 --no-cpu-baseline --steps 3"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32; do
  for CMR in "24 12 16" "24 16 16" "24 20 16" "32 16 16" "32 24 16" "24 16 32" "32 24 32"; do set -- $CMR
    mb=6; [ $1 -gt 24 ] && mb=5
    PT_MPARK=1 PT_MPARK_CAP=$1 PT_MPARK_MIN=$2 PT_SDF_REPS=$3 PT_MIN_BLOCKS=$mb timeout 300 python bench.py --workload $wl $B > $O/${wl}_mpark_cap$1_min$2_R$3.json 2> $O/${wl}_mpark_cap$1_min$2_R$3.err
  done
done
PT_MPARK=1 PT_MPARK_MIN=16 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_mpark_min16.log 2>&1
cat $O/stats_mpark_min16.log
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
