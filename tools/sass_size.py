#!/usr/bin/env python3
"""Static SASS size per source function: sass_size.py file.cubin kernel [header]"""
import collections, re, subprocess, sys
cubin, kernel = sys.argv[1:3]
header = sys.argv[3] if len(sys.argv) > 3 else 'pathtracer_b200/csrc/pt_kernel.cuh'
sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
funcs = []
for i, l in enumerate(open(header).read().split('\n'), 1):
    m = re.match(r'(?:PT_DEV(?:_NOINLINE)?|__device__ __forceinline__|static __device__ __forceinline__)\s+[\w:<>&\s\*]*?(\w+)\(', l)
    if m:
        funcs.append((i, m.group(1)))
def func_at(line):
    name = '?'
    for i, n in funcs:
        if i <= line:
            name = n
    return name
cnt = collections.Counter(); cur = None; inside = False; total = 0
for ln in sass.split('\n'):
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m: inside = (m.group(1) == kernel)
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if inside and re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', ln) and cur:
        cnt[func_at(cur[1]) if cur[0].endswith('.cuh') else cur[0]] += 1; total += 1
print('total SASS instructions %d = %.1f KB' % (total, total * 16 / 1024))
for f, n in cnt.most_common(30): print('  %-34s %5d' % (f, n))
