#!/usr/bin/env python3
"""Device-timed samples/s of ANY scene file in the reference's schema (bench.py covers the BASELINE configs only).
usage: python tools/scene_rate.py scene.json [W H] [spf=64 steps=6 pl=5 mode=fast] [key=value tuning options]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pathtracer_b200 as pt
args = [a for a in sys.argv[1:] if '=' not in a]
kv = dict(a.split('=', 1) for a in sys.argv[1:] if '=' in a)
path = args[0]
W, H = (int(args[1]), int(args[2])) if len(args) >= 3 else (1920, 1080)
spf, steps, pl = int(kv.pop('spf', 64)), int(kv.pop('steps', 6)), int(kv.pop('pl', 5))
mode = pt.MODE_STRICT if kv.pop('mode', 'fast') == 'strict' else pt.MODE_FAST
sc = pt.Scene.load(path)
r = pt.Renderer(mode=mode, jit=2, options={k: int(v) for k, v in kv.items()})
r.set_scene(sc.pack_ubo(), sc.sdf_sources)
r.resize(W, H)
p = sc.pack_params(1, W, H, spf, pl)
for _ in range(3):
    r.dispatch(p)
r.sync(); r.kernel_time()
for _ in range(steps):
    r.dispatch(p)
ms, n = r.kernel_time()
print(json.dumps({'scene': path, 'width': W, 'height': H, 'spf': spf, 'path_length': pl, 'options': kv, 'driver': r.get_option('sched'),
                  'gsamples_per_s': W * H * spf * steps / (ms * 1e-3) / 1e9, 'ms_per_step': ms / steps}))
r.close()
