#!/usr/bin/env python3
"""Per-phase scheduling statistics of the in-warp drivers on the device (kernel built with option stats = 1).
usage: python tools/sched_stats.py [workload ...] [key=value ...]   e.g.  cfg5_scene10_4k sched=8 pool_min=24
tests/simt_stats.py gives the same numbers without a GPU (host emulator)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracer_b200 as pt
from bench import WORKLOADS, scene_file
names = ['NEW', 'ISECT', 'SDF', 'SHADE']
opts = {k: int(v) for k, v in (a.split('=', 1) for a in sys.argv[1:] if '=' in a)}
spf = opts.pop('spf', 64)
opts['stats'] = 1
for wl in [a for a in sys.argv[1:] if '=' not in a] or ['cfg5_scene10_4k']:
    scene, W, H, spp, pl, _, _ = WORKLOADS[wl]
    sc = pt.Scene.load(scene_file(scene))
    r = pt.Renderer(mode=pt.MODE_FAST, jit=2, options=opts)
    r.set_scene(sc.pack_ubo(), sc.sdf_sources)
    r.resize(W, H)
    p = sc.pack_params(1, W, H, spf, pl)
    r.debug_stats(True)
    r.dispatch(p)
    r.sync()
    out = r.debug_stats(True)
    tot = sum(out[2 * i] for i in range(4))
    print(wl, 'samples', W * H * spf, 'options', opts, 'driver', r.get_option('sched'))
    for i, n in enumerate(names):
        ex, ln = out[2 * i], out[2 * i + 1]
        print('  %-6s executions %12d (%5.1f%%)  avg lanes %5.2f  per sample-warp %.2f' % (n, ex, 100.0 * ex / max(tot, 1), ln / max(ex, 1), ex / (W * H * spf / 32)))
    hist = [out[8 + b] for b in range(8)]
    if sum(hist):
        print('  SDF executions by participants (1-4, 5-8, ... 29-32): ' + ' '.join('%.1f%%' % (100.0 * h / sum(hist)) for h in hist))
    r.close()
