#!/bin/bash
# time to fixed relRMSE with the v2s defaults (second half of BASELINE.json's metric); one JSON line per workload
O=gpurun_out/rmse2; mkdir -p $O
timeout 60 python tools/rmse_vs_time.py cfg1_scene0_512 cfg2_scene1_1080p > $O/rmse_a.jsonl 2> $O/rmse_a.err
timeout 75 python tools/rmse_vs_time.py cfg3_scene9_mandelbulb_1080p > $O/rmse_b.jsonl 2> $O/rmse_b.err
timeout 50 python tools/rmse_vs_time.py --ref-spp 16384 --max-spp 2048 cfg4a_scene10_menger_1080p_pl32 > $O/rmse_c.jsonl 2> $O/rmse_c.err
timeout 70 python tools/rmse_vs_time.py --ref-spp 16384 --max-spp 2048 cfg4b_scene8_terrain_1080p_pl32 > $O/rmse_d.jsonl 2> $O/rmse_d.err
cat $O/rmse_?.jsonl > $O/rmse_vs_time.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/rmse2/rmse_vs_time.jsonl'):
    d=json.loads(l); print(d['workload'], 'ref', d['reference']['spp'], '%.1fs'%d['reference']['seconds'], d['time_to_rel_rmse'])
PY
