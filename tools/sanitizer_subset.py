#!/usr/bin/env python3
"""A small parity subset for compute-sanitizer (memcheck / racecheck / initcheck run the kernels 10-100x slower, so the
whole -m gpu suite is out of reach): one driver per invocation (`sched=8` etc.), strict and fast builds, analytic and
SDF scenes, ragged sizes, several dispatches, the sum mode and -- for the default run -- the wavefront pipeline.
Strict results are still compared with the oracle bit for bit, so a sanitizer-clean run is also a correct one.
usage: compute-sanitizer --tool racecheck python tools/sanitizer_subset.py sched=8 [wavefront]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pathtracer_b200 as pt  # noqa: E402
from oracle import oracle  # noqa: E402

opts = {k: int(v) for k, v in (a.split('=', 1) for a in sys.argv[1:] if '=' in a)}
wavefront = 'wavefront' in sys.argv[1:]
CASES = [('scene0', 40, 24, 4, 2, 5), ('scene1', 33, 17, 6, 3, 5), ('scene10', 40, 24, 4, 2, 5), ('scene9', 33, 17, 2, 2, 5), ('scene8', 32, 16, 2, 2, 5)]
if opts.get('pregen_max_mb'):  # a frame whose records outgrow 1 MiB: eight bands over two buffers and two streams
    CASES.append(('scene0', 200, 64, 16, 16, 5))
bad = 0
for mode in (pt.MODE_STRICT, pt.MODE_FAST):
    for name, w, h, spp, spf, pl in CASES:
        sc = pt.Scene.load(os.path.join(ROOT, 'scenes', name + '.json'))
        ubo = sc.pack_ubo()
        p = sc.pack_params(1, w, h, spf, pl)
        r = pt.Renderer(device=0, mode=mode, jit=2, options=opts, pipeline=pt.PIPE_WAVEFRONT if wavefront else pt.PIPE_MEGAKERNEL)
        r.set_scene(ubo, sc.sdf_sources)
        r.resize(w, h)
        r.render(p, spp, spf)
        got = r.read_xyz()
        r.clear()
        r.dispatch_sum(p, 7, spf)
        r.finalize(p, spf)
        got2 = r.read_xyz()
        r.close()
        ok = np.isfinite(got).all() and np.isfinite(got2).all()
        if mode == pt.MODE_STRICT:
            ref = oracle.Oracle(ubo, [s.decode() for s in sc.sdf_sources]).render(p, spp, spf)
            ok = ok and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        print('%s mode %d %s: %s' % (name, mode, opts, 'ok' if ok else 'MISMATCH'), flush=True)
        bad += 0 if ok else 1
sys.exit(1 if bad else 0)
