#!/usr/bin/env python3
"""ncu raw page (csv) of one launch -> the summary json bench.py reads `roofline.traffic` from.
usage: ncu_summary.py raw.csv workload 'kernel description' samples_per_launch > profiles/<set>/ncu_summary_<cfg>.json"""
import csv, json, sys
raw, wl, kernel, spl = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(raw)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names, units, vals = rows[h], rows[h + 1], rows[h + 2]
d = {n: (v, u) for n, u, v in zip(names, units, vals)}
keep = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__icc_request_hit_rate.pct', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
out = {k: list(d[k]) for k in keep if k in d}
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tr = sum(float(d[k][0].replace(',', '')) * scale[d[k][1]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
out.update({'traffic_bytes_per_launch': tr, 'workload': wl, 'kernel': kernel, 'samples_per_launch': spl})
print(json.dumps(out, indent=1))
