#!/bin/bash
# Final measurement set of the round (shipped defaults): tests, smoke, bench lines for every workload with roofline and
# CPU baseline, the reference arm, launch list + one full ncu capture of the default bench's kernel.
mkdir -p gpurun_out/final gpurun_out/jd_final
O=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_cfg2_reference.json 2> $O/bench.err
python bench.py > $O/bench_cfg2_scene1_1080p.json 2>> $O/bench.err
for cfg in "cfg1_scene0_512 --spf 64 --steps 16" "cfg3_scene9_mandelbulb_1080p --steps 16" "cfg4a_scene10_menger_1080p_pl32 --steps 8" "cfg4b_scene8_terrain_1080p_pl32 --steps 8" "cfg5_scene10_4k --steps 8"; do
  set -- $cfg
  python bench.py --workload $cfg > $O/bench_$1.json 2>> $O/bench.err
done
python bench.py --mode strict --steps 8 --no-cpu-baseline > $O/bench_cfg2_strict.json 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
PT_JIT_DUMP=gpurun_out/jd_final ncu --set full --clock-control none -k regex:pt_render -s 3 -c 1 -f -o $O/prof_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cp gpurun_out/jd_final/pt_render_jit_0.cubin $O/cfg2_kernel.cubin 2>/dev/null
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
nproc > $O/host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/host.txt
tail -3 $O/pytest_gpu.log
