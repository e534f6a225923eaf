#!/bin/bash
# march parking, third edition: full march state on the stack, unfinished rays are re-parked after every SDF execution.
O=gpurun_out/mpark3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "march_parking or v2s_driver" > $O/pytest_mpark.log 2>&1; echo "pytest rc $?" >> $O/pytest_mpark.log
tail -4 $O/pytest_mpark.log
B="--warmup 2 --no-cpu-baseline --steps 3"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32; do
  for CMR in "20 8 16" "20 12 16" "20 16 16" "32 16 16" "32 24 16" "20 12 8"; do set -- $CMR
    mb=6; [ $1 -gt 20 ] && mb=5
    PT_MPARK=1 PT_MPARK_CAP=$1 PT_MPARK_MIN=$2 PT_SDF_REPS=$3 PT_MIN_BLOCKS=$mb timeout 300 python bench.py --workload $wl $B > $O/${wl}_mpark_cap$1_min$2_R$3.json 2> $O/${wl}_mpark_cap$1_min$2_R$3.err
  done
done
PT_MPARK=1 timeout 300 python bench.py --workload cfg5_scene10_4k $B > $O/cfg5_scene10_4k_mpark_cap20_min12_R16.json 2> $O/cfg5_mpark.err
PT_MPARK=1 timeout 300 python tools/sched_stats.py cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 > $O/stats_mpark.log 2>&1
cat $O/stats_mpark.log
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
