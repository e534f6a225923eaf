#!/usr/bin/env python3
"""Debug: per-phase scheduling statistics of the v2 driver (kernel built with PT_STATS=1)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PT_STATS'] = '1'
import numpy as np
import pathtracer_b200 as pt
from bench import WORKLOADS
names = ['NEW', 'ISECT', 'SDF', 'SHADE']
for wl in sys.argv[1:] or ['cfg2_scene1_1080p']:
    scene, W, H, spp, pl, _, _ = WORKLOADS[wl]
    sc = pt.Scene.load('scenes/%s.json' % scene)
    r = pt.Renderer(mode=pt.MODE_FAST, jit=2)
    r.set_scene(sc.pack_ubo(), sc.sdf_sources)
    r.resize(W, H)
    p = sc.pack_params(1, W, H, 8, pl)
    out = (C.c_ulonglong * 16)()
    L = pt.lib()
    L.pt_debug_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.pt_debug_stats(r._ctx, out, 1)
    r.dispatch(p)
    r.sync()
    L.pt_debug_stats(r._ctx, out, 1)
    tot = sum(out[2 * i] for i in range(4))
    print(wl, 'samples', W * H * 8)
    for i, n in enumerate(names):
        ex, ln = out[2 * i], out[2 * i + 1]
        print('  %-6s executions %12d (%5.1f%%)  avg lanes %5.2f  per sample-warp %.2f' % (n, ex, 100.0 * ex / max(tot, 1), ln / max(ex, 1), ex / (W * H * 8 / 32)))
    hist = [out[8 + b] for b in range(8)]
    if sum(hist):
        print('  SDF executions by lanes waiting (1-4, 5-8, ... 29-32): ' + ' '.join('%.1f%%' % (100.0 * h / sum(hist)) for h in hist))
    r.close()
