#!/usr/bin/env python3
"""Synthetic scenes in the reference's JSON schema for what its shipped scenes never exercise: primitive counts near the
uniform block's capacity (1024 object floats, host:39), where the closest-hit search goes through the BVH
(SURVEY.md section 8f-3).  Deterministic; `python tools/make_synthetic_scenes.py` rewrites scenes_synthetic/*.json.
The camera, plane, materials and lights are scene0's."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import pack  # noqa: E402


def scene_path(name):
    return os.path.join(ROOT, 'scenes', name + '.json')


def many_sphere_scene(n, seed=5, duplicates=0):
    """Synthetic scene near the uniform block's capacity (170 spheres): a 13-wide carpet of small spheres under one
    emitter.  `duplicates` spheres are exact copies of earlier ones, so their hits tie bit for bit."""
    base = pack.load_scene(scene_path('scene0'))
    rng = np.random.default_rng(seed)
    spheres = [{'position': [0.0, 6.0, -2.0], 'radius': 1.5, 'materialID': 1, 'lightID': 1}]
    for i in range(n - 1 - duplicates):
        x, z = (i % 13) - 6.0, (i // 13) - 6.0
        spheres.append({'position': [x * 0.9, 0.3 + 0.2 * float(rng.random()), z * 0.9], 'radius': 0.3,
                        'materialID': 1 + i % 3, 'lightID': 0})
    for i in range(duplicates):
        spheres.append(dict(spheres[1 + 7 * i]))
    return {'camera': base['camera'], 'sphere': spheres, 'plane': base['plane'], 'material': base['material'],
            'light': base['light']}


def mixed_scene(seed=11):
    """Spheres of very different sizes, rotated boxes, lenses and cyclides scattered above the ground plane:
    30*6 + 5 + 20*11 + 12*12 + 12*16 = 741 of the 1024 object floats."""
    base = pack.load_scene(scene_path('scene0'))
    rng = np.random.default_rng(seed)
    pos = lambda: [float(rng.uniform(-6.0, 6.0)), float(rng.uniform(0.5, 8.0)), float(rng.uniform(-6.0, 6.0))]
    rot = lambda: [float(v) for v in rng.uniform(0.0, 360.0, 3)]
    spheres = [{'position': [0.0, 14.0, 0.0], 'radius': 2.0, 'materialID': 1, 'lightID': 1}]
    for i in range(29):
        spheres.append({'position': pos(), 'radius': float(10.0 ** rng.uniform(-1.7, 0.3)), 'materialID': 1 + i % 3, 'lightID': 0})
    boxes = [{'position': pos(), 'rotation': rot(), 'size': [float(v) for v in rng.uniform(0.2, 2.5, 3)],
              'materialID': 1 + i % 3, 'lightID': 0} for i in range(20)]
    lenses = [{'position': pos(), 'rotation': rot(), 'radius': float(rng.uniform(0.3, 1.0)), 'focalLength': float(rng.uniform(1.0, 3.0)),
               'thickness': float(rng.uniform(0.0, 0.6)), 'isConverging': bool(i % 2), 'materialID': 1 + i % 3, 'lightID': 0}
              for i in range(12)]
    cyclides = [{'position': pos(), 'rotation': rot(), 'scale': [float(v) for v in rng.uniform(0.5, 1.5, 3)],
                 'a': 1.0, 'b': 0.98, 'c': 0.2, 'd': 0.5, 'boundingRadius': 1.8, 'materialID': 1 + i % 3, 'lightID': 0}
                for i in range(12)]
    camera = json.loads(json.dumps(base['camera']))  # same view directions, from 2.5x as far: the whole cluster in frame
    camera['position'] = [[2.5 * float(v) for v in shot] for shot in camera['position']]
    return {'camera': camera, 'sphere': spheres, 'plane': base['plane'], 'box': boxes, 'lens': lenses,
            'cyclide': cyclides, 'material': base['material'], 'light': base['light']}



def main():
    out = os.path.join(ROOT, 'scenes_synthetic')
    os.makedirs(out, exist_ok=True)
    for name, scene in (('spheres169', many_sphere_scene(169)), ('mixed74', mixed_scene())):
        with open(os.path.join(out, name + '.json'), 'w') as f:
            json.dump(scene, f, indent=1)
            f.write('\n')


if __name__ == '__main__':
    main()
