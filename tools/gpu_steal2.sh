#!/bin/bash
# v2s in fast mode with the whole dispatch pooled (PT_STEAL_S=0): parity, pool size (spf per dispatch), occupancy, feeder threshold.
O=gpurun_out/steal2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "v2s" > $O/pytest_v2s.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2s.log
tail -8 $O/pytest_v2s.log
B="--warmup 3 --no-cpu-baseline"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  PT_SCHED=5 timeout 300 python bench.py --workload $wl $B --steps 8 --spf 16 > $O/${wl}_v2s_S0_spf16.json 2> $O/${wl}_v2s_S0_spf16.err
  PT_SCHED=5 timeout 300 python bench.py --workload $wl $B --steps 4 --spf 32 > $O/${wl}_v2s_S0_spf32.json 2> $O/${wl}_v2s_S0_spf32.err
  PT_SCHED=5 timeout 300 python bench.py --workload $wl $B --steps 4 --spf 64 > $O/${wl}_v2s_S0_spf64.json 2> $O/${wl}_v2s_S0_spf64.err
  PT_SCHED=1 timeout 300 python bench.py --workload $wl $B --steps 4 --spf 64 > $O/${wl}_v2_spf64.json 2> $O/${wl}_v2_spf64.err
  PT_SCHED=5 PT_MIN_BLOCKS=5 timeout 300 python bench.py --workload $wl $B --steps 8 --spf 16 > $O/${wl}_v2s_S0_spf16_mb5.json 2> $O/${wl}_v2s_S0_spf16_mb5.err
  PT_SCHED=5 PT_FEED_T=4 timeout 300 python bench.py --workload $wl $B --steps 8 --spf 16 > $O/${wl}_v2s_S0_spf16_T4.json 2> $O/${wl}_v2s_S0_spf16_T4.err
  PT_SCHED=5 PT_FEED_T=6 PT_SDF_REPS=12 timeout 300 python bench.py --workload $wl $B --steps 8 --spf 16 > $O/${wl}_v2s_S0_spf16_T6_R12.json 2> $O/${wl}_v2s_S0_spf16_T6_R12.err
done
for wl in cfg2_scene1_1080p cfg1_scene0_512 bvh_mixed74_1080p; do
  PT_SCHED=5 timeout 300 python bench.py --workload $wl $B --steps 8 > $O/${wl}_v2s_S0.json 2> $O/${wl}_v2s_S0.err
done
PT_SCHED=0 timeout 300 python bench.py --workload cfg2_scene1_1080p $B --steps 4 --spf 64 > $O/cfg2_scene1_1080p_v1_spf64.json 2> $O/cfg2_v1_spf64.err
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
