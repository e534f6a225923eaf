#!/usr/bin/env python3
"""Analysis script (not a test): the in-warp drivers' scheduling statistics (PT_STATS: executions and lanes served per
phase) WITHOUT a GPU, from the host SIMT emulator at the full frame size of a BASELINE config on a sample of its
16x8-pixel blocks.  The counterpart of tools/sched_stats.py (which needs a device); strict arithmetic, so the paths are
the oracle's.  Knobs are given as NAME=VALUE defines.
usage: python tests/simt_stats.py scene9 [spf=8] [pl=5] [PT_SCHED=5 PT_STEAL_S=0 PT_FEED_T=8 ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pathtracer_b200 as pt  # noqa: E402
from test_simt_emulation import build_emulator, scene_inputs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'scene9'
opts = dict(a.split('=', 1) for a in sys.argv[2:])
spf, pl = int(opts.pop('spf', 8)), int(opts.pop('pl', 5))
W, H = int(opts.pop('W', 1920)), int(opts.pop('H', 1080))
defs = {'PT_SCHED': 5, 'PT_STEAL_S': 0, 'PT_STATS': 1}
defs.update({k: int(v) for k, v in opts.items()})
ubo, p, src, raw = scene_inputs(name, W, H, spf, pl)
L = build_emulator(pt, defs, src, raw)
L.simt_select_blocks.argtypes = [C.c_int] * 4
L.simt_stats.argtypes = [C.c_void_p, C.c_int]
img = np.zeros((H, W, 4), dtype=np.float32)
out = (C.c_ulonglong * 16)()
L.simt_stats(out, 1)
blocks = 0
gx, gy = (W + 15) // 16, (H + 7) // 8
for by in range(gy // 16, gy, gy // 8):          # 8 bands x 8 columns of 2x2 blocks: 256 blocks, 1024 warps
    for bx in range(gx // 16, gx, gx // 8):
        L.simt_select_blocks(bx, by, 2, 2)
        q = np.ascontiguousarray(p)
        assert L.simt_dispatch(ubo.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), 2, 0, spf, img.ctypes.data_as(C.c_void_p), 2) == 0
        blocks += 4
L.simt_stats(out, 0)
samples = blocks * 128 * spf
tot = sum(out[2 * i] for i in range(4))
print('%s %dx%d pathLength %d, %d samples per pixel, %d blocks, defines %s' % (name, W, H, pl, spf, blocks, defs))
for i, n in enumerate(['NEW', 'ISECT', 'SDF', 'SHADE']):
    ex, ln = out[2 * i], out[2 * i + 1]
    print('  %-6s executions %10d (%5.1f%%)  avg lanes %5.2f  per sample-warp %.2f' % (n, ex, 100.0 * ex / max(tot, 1), ln / max(ex, 1), ex / (samples / 32)))
hist = [out[8 + b] for b in range(8)]
if sum(hist):
    print('  SDF executions by participants (1-4, 5-8, ... 29-32): ' + ' '.join('%.1f%%' % (100.0 * h / sum(hist)) for h in hist))
# a coarse issue-slot model, calibrated on the ncu capture of v2s on cfg3 (profiles/r01_final4: 266 warp instructions per
# sample, 42 % of them in the SDF phase): warp instructions per 32 samples = sum over phases of executions x cost
R = defs.get('PT_SDF_REPS', 16)
COST = {'NEW': 700.0, 'ISECT': 450.0, 'SHADE': 550.0, 'SDF': R * 211.0}
per = {n: out[2 * i] / (samples / 32) for i, n in enumerate(['NEW', 'ISECT', 'SDF', 'SHADE'])}
model = sum(per[n] * COST[n] for n in per) + sum(per.values()) * 58.0
print('  model: %.0f warp instructions per 32 samples (calibration point: 8500 for v2s, T 8, R 16 on scene9)' % model)

