#!/bin/bash
# One parametrised GPU session script (replaces round 1's 43 one-off tools/gpu_*.sh).
#   usage: gpurun --timeout T -- 'bash tools/gpu_session.sh <out-name> <step> [<step> ...]'
# Steps (each writes under gpurun_out/<out-name>/):
#   smoke            __graft_entry__.smoke()
#   tests            the whole -m gpu suite
#   tests:<expr>     -m gpu suite restricted by -k <expr>
#   bench            default bench line (headline + per-workload) and the reference arm
#   ab:<workload>:<name>:<opts>   bench.py --workload <workload> --no-per-workload --no-cpu-baseline --no-rmse with
#                    --opt k=v for every k=v in the comma separated <opts>   (A/B of tuning options)
#   launches         ncu launch list (gpu__time_duration) of the default bench command
#   ncu:<workload>[:<opts>]   ncu --set full of the workload's render kernel + summary json + per-function profile
#   stats:<workload>[:<opts>] scheduling statistics of the in-warp drivers
#   sanitizer:<tool>[:<opts>] compute-sanitizer --tool <memcheck|racecheck|initcheck> over a small parity subset, one run per
#                    driver and per generation / resolve kernel configuration (or only for the comma separated <opts>)
#   wavefront        ncu of the wavefront pipeline's kernels on cfg3 (queue traffic, L2 hit rate, sectors per request)
#   rmse             tools/rmse_vs_time.py with a strict reference at 1/4 resolution
O=gpurun_out/$1; shift; mkdir -p $O
optflags() { local f=""; IFS=',' read -ra KV <<< "$1"; for kv in "${KV[@]}"; do [ -n "$kv" ] && f="$f --opt $kv"; done; echo $f; }
for step in "$@"; do
  IFS=':' read -r kind a b c <<< "$step"
  case $kind in
    smoke) python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log;;
    tests) if [ -n "$a" ]; then timeout 2400 python -m pytest tests -m gpu -q -s -k "$a" > $O/pytest_gpu_k.log 2>&1; echo "pytest -k rc=$?" >> $O/pytest_gpu_k.log; tail -4 $O/pytest_gpu_k.log
           else timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log; fi;;
    bench) python bench.py --impl reference --steps 4 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
           python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
           python - $O/bench_default.json $O/bench_reference.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=json.load(open(sys.argv[2]))
print('headline %s: %.3f Gs/s  e2e %.3f  frac %.3f (peak %.1f TF)  ref arm %.2f Ms/s (%s, %d cores)  e2e/ref %.0fx' % (d['config']['workload'], d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac'], d['roofline']['peak'], r['value']/1e6, r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], d['e2e']['value']/r['value']))
for k,v in d['config'].get('per_workload',{}).items(): print('  %-36s %.3f Gs/s e2e %.3f frac %s sched %s' % (k, v['value']/1e9, v['e2e']/1e9, v['roofline_frac'] and round(v['roofline_frac'],3), v['driver_sched']))
print('  time_to_rel_rmse', json.dumps(d.get('time_to_rel_rmse',{}).get('targets')), 'spread', d.get('time_to_rel_rmse',{}).get('fit_spread'))
print('  clocks', d.get('clocks'))
PY
           ;;
    ab) python bench.py --workload $a --no-per-workload --no-cpu-baseline --no-rmse --steps 8 --warmup 3 $(optflags "$c") > $O/ab_${a}_${b}.json 2> $O/ab_${a}_${b}.err
        python - $O/ab_${a}_${b}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.3f Gs/s  e2e %.3f  sched %s  frac %s' % (d['value']/1e9, d['e2e']['value']/1e9, d['config']['driver_sched'], round(d.get('roofline',{}).get('frac') or 0,3)))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
        ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-rmse > $O/launches_bench.log 2>&1; grep -c pt_render $O/launches_default.csv;;
    ncu) mkdir -p $O/jd_$a; tag=${a%%_*}
         PT_JIT_DUMP=$O/jd_$a timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_render_jit -s 3 -c 1 -f -o $O/ncu_$tag python bench.py --workload $a --no-per-workload --no-cpu-baseline --no-rmse --steps 2 --warmup 3 $(optflags "$b") > $O/ncu_$tag.log 2>&1
         ncu -i $O/ncu_$tag.ncu-rep --page raw --csv > $O/ncu_raw_$tag.csv 2>/dev/null
         read W H <<< $(python -c "from bench import WORKLOADS as w; print(w['$a'][1], w['$a'][2])")
         python tools/ncu_summary.py $O/ncu_raw_$tag.csv $a "pt_render_jit $b" $((W*H*64)) > $O/ncu_summary_$tag.json 2>$O/ncu_summary_$tag.err
         cub=$(ls $O/jd_$a/*.cubin 2>/dev/null | head -1); [ -n "$cub" ] && python tools/ncu_by_line.py $O/ncu_$tag.ncu-rep $cub pt_render_jit > $O/by_function_$tag.txt 2>&1
         rm -f $O/ncu_$tag.ncu-rep.tmp; head -c 1200 $O/ncu_summary_$tag.json;;
    stats) python tools/sched_stats.py $a $(echo $b | tr ',' ' ') > $O/stats_${a}_$(echo $b | tr ',=' '__').txt 2>&1; cat $O/stats_${a}_*.txt | tail -8;;
    sanitizer) for drv in ${b:-"sched=0" "sched=7" "sched=5" "sched=8" "sched=5,pregen=1,resolve=1,pregen_max_mb=1" "sched=7,pregen=1,pregen_max_mb=1"}; do
                 tag=$(echo $drv | tr ',=' '__')
                 timeout 1500 compute-sanitizer --tool $a --error-exitcode 1 python tools/sanitizer_subset.py $(echo $drv | tr ',' ' ') > $O/sanitizer_${a}_$tag.log 2>&1
                 echo "sanitizer $a $drv rc=$?" | tee -a $O/sanitizer_summary.txt; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/sanitizer_${a}_$tag.log | tail -1 | tee -a $O/sanitizer_summary.txt
               done;;
    wavefront) timeout 900 ncu --set full --clock-control none -k regex:pt_wf_ -s 40 -c 24 -f -o $O/ncu_wavefront python bench.py --workload cfg3_scene9_mandelbulb_1080p --pipeline wavefront --no-per-workload --no-cpu-baseline --no-rmse --steps 1 --warmup 3 --spf 16 > $O/ncu_wavefront.log 2>&1
               ncu -i $O/ncu_wavefront.ncu-rep --page raw --csv > $O/ncu_wavefront_raw.csv 2>/dev/null
               python tools/wavefront_summary.py $O/ncu_wavefront_raw.csv > $O/wavefront_summary.txt 2>&1; cat $O/wavefront_summary.txt;;
    rmse) python tools/rmse_vs_time.py cfg1_scene0_512 cfg2_scene1_1080p cfg3_scene9_mandelbulb_1080p cfg5_scene10_4k --scale 4 --ref-spp 16384 --max-spp 1024 > $O/rmse_vs_time.jsonl 2> $O/rmse.err; wc -l $O/rmse_vs_time.jsonl;;
    *) echo "unknown step $step";;
  esac
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv
nproc > $O/host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/host.txt
