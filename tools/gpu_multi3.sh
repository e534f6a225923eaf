#!/bin/bash
# 2 GPUs with the v2s default and 64-sample steps: multi-GPU tests, pt_multi CLI at 1 and 2 GPUs, bench.py torchrun N=2.
O=gpurun_out/multi3; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.csv
timeout 600 python -m pytest tests -m gpu -q -k "multi" > $O/pytest_multi.log 2>&1; echo "pytest rc $?" >> $O/pytest_multi.log
tail -4 $O/pytest_multi.log
for g in 1 2; do
  timeout 300 pathtracer_b200/lib/pt_render --scene scenes/scene10.json --width 3840 --height 2160 --spp 256 --spf 64 --fast --jit 2 --gpus $g > $O/cli_scene10_4k_gpus$g.json 2> $O/cli_scene10_4k_gpus$g.err
done
cat $O/cli_*.json
for wl in cfg2_scene1_1080p cfg5_scene10_4k; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_${wl}_n2.json 2> $O/bench_${wl}_n2.err
  timeout 300 python bench.py --gpus 1 --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
done
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'n_gpus', d['n_gpus'], '%.3f Gs/s'%(d['value']/1e9), 'reduce_ms', d.get('reduce_ms'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
