#!/usr/bin/env python3
"""ncu raw page (csv) of the wavefront pipeline's kernels -> per-kernel table of what north_star asks about the path-state
queues: time, achieved DRAM GB/s, L2 hit rate, L1 sectors per request (coalescing), lanes per instruction.
usage: wavefront_summary.py ncu_wavefront_raw.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names, units = rows[h], rows[h + 1]
ix = {n: i for i, n in enumerate(names)}
want = {'t_us': 'gpu__time_duration.sum', 'dram_rd': 'dram__bytes_read.sum', 'dram_wr': 'dram__bytes_write.sum',
        'l2_hit': 'lts__t_sector_hit_rate.pct', 'l1_hit': 'l1tex__t_sector_hit_rate.pct',
        'ld_sectors': 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'ld_requests': 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'st_sectors': 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'st_requests': 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lanes': 'smsp__thread_inst_executed_per_inst_executed.ratio', 'issue': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'}
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}
agg = collections.OrderedDict()
for r in rows[h + 2:]:
    if len(r) < len(names):
        continue
    k = r[ix['Kernel Name']]
    d = agg.setdefault(k, collections.defaultdict(float))
    d['n'] += 1
    for key, metric in want.items():
        if metric in ix and r[ix[metric]] not in ('', 'n/a'):
            v = float(r[ix[metric]].replace(',', '')) * scale.get(units[ix[metric]], 1.0)
            d[key] += v
print('%-14s %4s %10s %10s %9s %8s %8s %9s %9s %7s %7s %7s' % ('kernel', 'n', 'time us', 'DRAM MB', 'DRAM GB/s', 'L2 hit%', 'L1 hit%', 'ld sec/rq', 'st sec/rq', 'lanes', 'issue%', 'DRAM%'))
tot_t = tot_b = 0.0
for k, d in agg.items():
    n = d['n']
    b = d['dram_rd'] + d['dram_wr']
    tot_t += d['t_us']; tot_b += b
    print('%-14s %4d %10.1f %10.2f %9.1f %8.1f %8.1f %9.2f %9.2f %7.2f %7.1f %7.1f' % (
        k[:14], n, d['t_us'], b / 1e6, b / max(d['t_us'], 1e-9) / 1e3, d['l2_hit'] / n, d['l1_hit'] / n,
        d['ld_sectors'] / max(d['ld_requests'], 1), d['st_sectors'] / max(d['st_requests'], 1), d['lanes'] / n, d['issue'] / n, d['dram_pct'] / n))
print('total: %.1f us in the captured launches, %.2f MB of DRAM traffic, %.1f GB/s average' % (tot_t, tot_b / 1e6, tot_b / max(tot_t, 1e-9) / 1e3))
