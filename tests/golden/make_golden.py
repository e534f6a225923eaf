#!/usr/bin/env python3
"""Generates tests/golden/*.npz: XYZ buffers of the CPU oracle (strict semantics) for small renders of the shipped
scenes, plus RNG / seed vectors.  The reference itself cannot run in this image (no Vulkan, no glslang: SURVEY.md
section 0-3) and ships no golden data, so these pin OUR canonical interpretation (oracle/oracle.cpp): the CPU tests
check the oracle still reproduces them, the GPU tests check the CUDA kernels (megakernel and wavefront) do.
Run from the repo root:  python tests/golden/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, pack

CASES = [('scene0', 1, 32, 24, 2, 1, 5), ('scene1', 1, 32, 24, 2, 2, 5), ('scene2', 1, 32, 24, 1, 1, 5), ('scene9', 1, 32, 24, 2, 2, 5),
         ('scene10', 2, 32, 24, 1, 1, 32), ('scene8', 1, 32, 24, 1, 1, 32), ('scene3', 1, 32, 24, 1, 1, 5), ('scene7', 1, 32, 24, 1, 1, 5)]

def main():
    out = os.path.join(ROOT, 'tests', 'golden')
    for name, shot, w, h, spp, spf, pl in CASES:
        o, scene = oracle.from_scene_file(os.path.join(ROOT, 'scenes', name + '.json'))
        p = pack.pack_params(scene, shot, w, h, spf, pl)
        img = o.render(p, spp, spf)
        np.savez_compressed(os.path.join(out, '%s_%dx%d_%dspp.npz' % (name, w, h, spp)), xyz=img,
                            meta=np.array([shot, w, h, spp, spf, pl], dtype=np.int32))
        print(name, img[..., 1].mean())
    L = oracle.lib()
    seeds = np.concatenate([np.arange(64, dtype=np.uint32), np.array([0xFFFFFFFF, 0x12345678, 0x3FFF], dtype=np.uint32)])
    np.savez_compressed(os.path.join(out, 'pcg32.npz'), seed=seeds, out=np.array([L.oracle_pcg32(int(s)) for s in seeds], dtype=np.uint32))

if __name__ == '__main__':
    main()
