#!/bin/bash
# pt_multi (single host thread, NCCL reduce) on 2 GPUs: tests + the CLI at 1 and 2 GPUs, next to bench.py's torchrun path.
O=gpurun_out/multi2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "native_multi" > $O/pytest_multi.log 2>&1; echo "pytest rc $?" >> $O/pytest_multi.log
tail -5 $O/pytest_multi.log
for g in 1 2; do
  timeout 300 pathtracer_b200/lib/pt_render --scene scenes/scene1.json --width 1920 --height 1080 --spp 1024 --spf 16 --fast --jit 2 --gpus $g > $O/cli_scene1_gpus$g.json 2> $O/cli_scene1_gpus$g.err
  timeout 300 pathtracer_b200/lib/pt_render --scene scenes/scene10.json --width 3840 --height 2160 --spp 256 --spf 16 --fast --jit 2 --gpus $g > $O/cli_scene10_4k_gpus$g.json 2> $O/cli_scene10_4k_gpus$g.err
done
cat $O/cli_*.json $O/cli_*.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 3 --no-cpu-baseline > $O/bench_cfg2_n2.json 2> $O/bench_cfg2_n2.err
python - <<'PY'
import json
t=open('gpurun_out/multi2/bench_cfg2_n2.json').read(); print('stdout lines:', t.count(chr(10))); d=json.loads(t); print('bench.py torchrun n=2: %.2f Gs/s reduce %.3f ms'%(d['value']/1e9, d['reduce_ms']))
PY
