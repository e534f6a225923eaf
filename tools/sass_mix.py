#!/usr/bin/env python3
"""Executed-instruction mix of one profiled kernel: which SASS opcodes, and for the costly ones which source lines.
usage: python tools/sass_mix.py <ncu --page source --csv file> <cubin of the same build> [kernel] [opcodes,to,break,down]
(the csv: `ncu -i rep.ncu-rep --page source --csv`; the cubin: PT_JIT_DUMP of the same run, for nvdisasm -g line info)"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, cubin = sys.argv[1:3]
kernel = sys.argv[3] if len(sys.argv) > 3 else 'pt_render_jit'
detail = sys.argv[4].split(',') if len(sys.argv) > 4 else ['MOV', 'FSEL', 'FSETP', 'ISETP', 'BRA']
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
cur, inside, amap = None, False, {}
for ln in sass.split('\n'):
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        inside = (m.group(1) == kernel)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S+)', ln)
    if inside and m and cur:
        amap[int(m.group(1), 16)] = cur
ops, lanes, by = collections.Counter(), collections.Counter(), {op: collections.Counter() for op in detail}
tot, base = 0.0, None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
    if not m:
        continue
    n = float(r[ix['Instructions Executed']] or 0)
    op = m.group(2).split('.')[0]
    a = int(r[ix['Address']], 16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']])
    base = a if base is None else base
    ops[op] += n
    lanes[op] += float(r[ix['Thread Instructions Executed']] or 0)
    tot += n
    if op in by:
        by[op][amap.get(a - base, ('?', 0))] += n
print('total warp instructions %.4g' % tot)
for op, n in ops.most_common(40):
    print('%-10s %6.2f%%  avg lanes %4.1f' % (op, 100 * n / tot, lanes[op] / max(n, 1)))
cache = {}
def text(f, l):
    for d in ('pathtracer_b200/csrc/', 'include/'):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            cache.setdefault(f, open(p).read().split('\n'))
            return cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ''
    return ''
for op in detail:
    print('== %s by source line' % op)
    for (f, l), n in by[op].most_common(16):
        print('  %5.2f%%  %s:%d  %s' % (100 * n / tot, f, l, text(f, l)))
