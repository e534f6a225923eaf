#!/bin/bash
# multi-GPU session (gpurun --gpus N): 2-GPU parity test, then bench at N = 1, 2 (and 4, 8 when present)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" > gpurun_out/multi.log
python -m pytest tests/test_multi_gpu.py -m gpu -x -q >> gpurun_out/multi.log 2>&1
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  for wl in cfg2_scene1_1080p cfg5_scene10_4k; do
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${wl}_n$n.json 2>> gpurun_out/multi.log
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload $wl --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${wl}_n$n.json 2>> gpurun_out/multi.log
    fi
  done
done
tail -5 gpurun_out/multi.log
