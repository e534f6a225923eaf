#!/bin/bash
# v2sp (PT_SCHED=6: persistent warps streaming over tiles) against v2s: parity, then spf 1 / 4 / 16 / 64 and 2 / 4 slots.
O=gpurun_out/steal3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "v2s" > $O/pytest_v2s.log 2>&1; echo "pytest rc $?" >> $O/pytest_v2s.log
tail -8 $O/pytest_v2s.log
grep "pooled vs" $O/pytest_v2s.log
B="--warmup 3 --no-cpu-baseline"
for wl in cfg3_scene9_mandelbulb_1080p cfg4a_scene10_menger_1080p_pl32 cfg4b_scene8_terrain_1080p_pl32 cfg5_scene10_4k; do
  for spf in 4 16 64; do
    st=$(( 256 / spf )); [ $st -gt 16 ] && st=16
    PT_SCHED=5 timeout 300 python bench.py --workload $wl $B --steps $st --spf $spf > $O/${wl}_v2s_spf$spf.json 2> $O/${wl}_v2s_spf$spf.err
    PT_SCHED=6 timeout 300 python bench.py --workload $wl $B --steps $st --spf $spf > $O/${wl}_v2sp_spf$spf.json 2> $O/${wl}_v2sp_spf$spf.err
  done
  PT_SCHED=6 PT_TILE_SLOTS=2 timeout 300 python bench.py --workload $wl $B --steps 4 --spf 64 > $O/${wl}_v2sp_slots2_spf64.json 2> $O/${wl}_v2sp_slots2_spf64.err
  PT_SCHED=6 PT_TILE_SLOTS=8 timeout 300 python bench.py --workload $wl $B --steps 16 --spf 4 > $O/${wl}_v2sp_slots8_spf4.json 2> $O/${wl}_v2sp_slots8_spf4.err
done
for wl in cfg2_scene1_1080p cfg1_scene0_512 bvh_mixed74_1080p; do
  PT_SCHED=6 timeout 300 python bench.py --workload $wl $B --steps 8 > $O/${wl}_v2sp_spf64.json 2> $O/${wl}_v2sp_spf64.err
done
PT_SCHED=6 timeout 300 python bench.py --workload cfg2_scene1_1080p $B --steps 16 --spf 4 > $O/cfg2_scene1_1080p_v2sp_spf4.json 2> $O/cfg2_v2sp_spf4.err
PT_SCHED=0 timeout 300 python bench.py --workload cfg2_scene1_1080p $B --steps 16 --spf 4 > $O/cfg2_scene1_1080p_v1_spf4.json 2> $O/cfg2_v1_spf4.err
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s'%(d['value']/1e9))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
