#!/usr/bin/env python3
"""Analysis script (not a test): how evenly is the work of a warp's 8x4 pixel tile spread over its 32 lanes when every
lane runs its own pixel's samples (drivers v1 / v2)?  Uses the oracle's per-sample counters at the FULL frame size of
the BASELINE configs on four 32-row bands.  Output committed as profiles/r01_steal/lane_balance.txt; the motivation
for the sample-stealing driver v2s (pt_kernel.cuh).
usage: python tests/lane_balance.py [spf]"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle, pack  # noqa: E402

SPF = int(sys.argv[1]) if len(sys.argv) > 1 else 16
BANDS = ((200, 232), (500, 532), (700, 732), (900, 932))
# relative cost of one sample in units of "one feeder-phase execution": camera + rays + shading + SDF evaluations
WEIGHTS = {'sdf evaluations only': (1.0, 0.0, 0.0, 0.0), 'camera 1 + ray 1 + bounce 0.6 + SDF eval 0.5': (0.5, 1.0, 0.6, 1.0)}

print('lane balance of an 8x4 tile, %d samples per pixel per dispatch, 1920x1080, rows %s' % (SPF, BANDS))
for name, pl in (('scene1', 5), ('scene9', 5), ('scene10', 32), ('scene8', 32)):
    o, scene = oracle.from_scene_file(os.path.join(ROOT, 'scenes', name + '.json'), count=True)
    p = pack.pack_params(scene, shot=1, width=1920, height=1080, spf=1, path_length=pl)
    cm = np.concatenate([o.cost_map(p, SPF, y0, y1) for (y0, y1) in BANDS], 0).astype(np.float64)
    for label, w in WEIGHTS.items():
        cost = w[0] * cm[..., 0] + w[1] * cm[..., 1] + w[2] * cm[..., 2] + w[3]
        if cost.sum() == 0:
            continue
        hh, ww, s = cost.shape
        tiles = cost.reshape(hh // 4, 4, ww // 8, 8, s).transpose(0, 2, 1, 3, 4).reshape(-1, 32, s)
        lane = tiles.sum(2)
        own = lane.mean(1).sum() / lane.max(1).sum()           # a warp lives as long as its busiest lane
        per_sample = tiles.mean(1).sum() / tiles.max(1).sum()  # v1: the warp waits for the longest path of every sample
        # stealing: the pool drains at 32 lanes; the tail is bounded by the most expensive single sample
        steal = tiles.sum((1, 2)).sum() / (np.maximum(tiles.sum((1, 2)) / 32.0, tiles.max((1, 2))) * 32.0).sum()
        print('%-8s pathLength %-2d %-46s mean cost %7.2f | mean lane / busiest lane: own pixel %.3f, per-sample sync %.3f, pooled (lower bound) %.3f'
              % (name, pl, label, cost.mean(), own, per_sample, steal))
