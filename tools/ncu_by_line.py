#!/usr/bin/env python3
"""Joins an ncu report's per-SASS-instruction counters with nvdisasm's line info and prints executed warp
instructions per source function / line.  usage: ncu_by_line.py report.ncu-rep file.cubin kernel_name [header]"""
import collections, csv, io, re, subprocess, sys

rep, cubin, kernel = sys.argv[1:4]
header = sys.argv[4] if len(sys.argv) > 4 else 'pathtracer_b200/csrc/pt_kernel.cuh'
sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
line_of, cur, inside = {}, None, False
for ln in sass.split('\n'):
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        inside = (m.group(1) == kernel)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and inside:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
funcs = []
src = open(header).read().split('\n')
for i, l in enumerate(src, 1):
    m = re.match(r'(?:PT_DEV(?:_NOINLINE)?|__device__ __forceinline__)\s+[\w:<>&\s\*]*?(\w+)\(', l)
    if m:
        funcs.append((i, m.group(1)))
def func_at(line):
    name = '?'
    for i, n in funcs:
        if i <= line:
            name = n
    return name
by_line, by_func, lanes_func = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
base = None
for r in rows[2:]:
    try:
        addr = int(r[ix['Address']], 16) if not r[ix['Address']].isdigit() else int(r[ix['Address']])
        w = int(r[ix['Instructions Executed']]); t = int(r[ix['Thread Instructions Executed']])
    except (ValueError, KeyError, IndexError):
        continue
    if base is None:
        base = addr
    loc = line_of.get(addr - base) or line_of.get(addr)
    if not loc:
        continue
    f = func_at(loc[1]) if loc[0].endswith('.cuh') else loc[0]
    by_line[(loc, )] += w
    by_func[f] += w
    lanes_func[f] += t
    tot += w
print('total warp instructions %.3e' % tot)
for f, w in by_func.most_common(40):
    print('%-34s %6.2f%%  avg lanes %5.1f' % (f, 100.0 * w / tot, lanes_func[f] / max(w, 1)))
print('--- top lines')
for (loc,), w in by_line.most_common(25):
    print('%s:%d %5.2f%%  %s' % (loc[0], loc[1], 100.0 * w / tot, src[loc[1] - 1].strip()[:110] if loc[0].endswith('.cuh') else ''))
