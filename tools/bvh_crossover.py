#!/usr/bin/env python3
"""Where the BVH starts to pay: sphere carpets of n spheres (tools/make_synthetic_scenes.py) at 1920x1080, fast mode,
count-specialised kernels, closest hit through the reference's scan vs through the tree.  Prints one JSON line per n
with the device time of each (CUDA events on the library's stream, after a warm-up dispatch).  GPU only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import pathtracer_b200 as pt  # noqa: E402
from make_synthetic_scenes import many_sphere_scene  # noqa: E402

W, H, SPF, STEPS = 1920, 1080, 16, 4


def rate(scene_text, bvh_min, options=None):
    sc = pt.Scene.parse(scene_text)
    r = pt.Renderer(device=0, mode=pt.MODE_FAST, jit=2, options=options)
    r.set_bvh(bvh_min)
    r.set_scene(sc.pack_ubo())
    active = r.bvh_active
    r.resize(W, H)
    p = sc.pack_params(1, W, H, SPF, 5)
    r.dispatch(p)
    r.sync()
    r.kernel_time()
    for _ in range(STEPS):
        r.dispatch(p)
    ms, _n = r.kernel_time()
    r.close()
    return W * H * SPF * STEPS / (ms * 1e-3), active


def main():
    for n in [int(a) for a in sys.argv[1:]] or [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 169]:
        text = json.dumps(many_sphere_scene(n))
        scan, a0 = rate(text, 0)
        scan_rolled, _ = rate(text, 0, {'no_unroll': 1})
        tree, a1 = rate(text, 2)
        assert not a0 and a1
        print(json.dumps({'spheres': n, 'scan_gsamples_s': scan / 1e9, 'scan_rolled_gsamples_s': scan_rolled / 1e9,
                          'tree_gsamples_s': tree / 1e9, 'tree_over_best_scan': tree / max(scan, scan_rolled)}), flush=True)


if __name__ == '__main__':
    main()
