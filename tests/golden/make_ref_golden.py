#!/usr/bin/env python3
"""Generates tests/golden/ref_*.npz from the REFERENCE ITSELF: oracle/_ref = the reference's src/pathtracer.cpp (loader,
packer, push constants) and src/shader.comp (over the vendored glm) compiled for the CPU by oracle/ref_build.py.
These are outputs of the reference's own code run in this container (/root/reference must be present); they travel as
small committed fixtures so that the oracle (CPU) and the CUDA kernels (GPU box, where /root/reference does not exist)
can be compared with them.  Contents per scene: the uniform block and push constants the reference packs, per-sample
XYZ images of 1-sample dispatches (frame = k + 1) and a multi-dispatch running-mean image.
Also: PCG32 / GenerateSeed vectors and leaf-function vectors (WaveToXYZ, SampleWavelengths, BK7, Emit, SPD).
Run from the repo root:  python tests/golden/make_ref_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref

CASES = [('scene0', 1, 32, 24, 4, 5), ('scene1', 1, 32, 24, 4, 5), ('scene2', 1, 24, 16, 3, 5), ('scene3', 1, 24, 16, 3, 5),
         ('scene9', 1, 32, 24, 4, 5), ('scene10', 1, 32, 24, 4, 5), ('scene10', 2, 24, 16, 2, 32), ('scene8', 1, 32, 24, 3, 5), ('scene7', 1, 24, 16, 2, 5)]


def main():
    out = os.path.join(ROOT, 'tests', 'golden')
    for name, shot, w, h, n, pl in CASES:
        rs = ref.RefScene(os.path.join(ROOT, 'scenes', name + '.json'), shot)
        ubo = rs.ubo()
        samples = np.zeros((n, h, w, 3), dtype=np.float32)
        for k in range(n):
            img = np.zeros((h, w, 4), dtype=np.float32)
            rs.dispatch(rs.push(w, h, k + 1, 1, 1, pl), img)
            samples[k] = img[..., :3]
        mean = rs.render(w, h, 2 * n, 2, pl)          # n dispatches of 2 samples: Accumulate()'s running mean
        push = rs.push(w, h, 2, 2, 2, pl)
        np.savez_compressed(os.path.join(out, 'ref_%s_shot%d_%dx%d_pl%d.npz' % (name, shot, w, h, pl)), ubo=ubo, push=push, samples=samples,
                            mean=mean, meta=np.array([shot, w, h, n, pl], dtype=np.int32))
        print(name, shot, samples.mean(), mean[..., :3].mean())
    sh = ref.RefShader(os.path.join(ROOT, 'scenes', 'scene0.json'))
    rng = np.random.default_rng(1)
    seeds = np.concatenate([np.arange(256, dtype=np.uint32), rng.integers(0, 2**32, 4096, dtype=np.uint64).astype(np.uint32),
                            np.array([0xFFFFFFFF, 0x12345678, 0xFFFFFF80], dtype=np.uint32)])
    rs = ref.RefScene(os.path.join(ROOT, 'scenes', 'scene0.json'))
    push = rs.push(1920, 1080, 64, 64, 8)
    gxyk = np.stack([rng.integers(0, 1920, 2000), rng.integers(0, 1080, 2000), rng.integers(0, 8, 2000)], 1).astype(np.int32)
    waves = rng.uniform(360, 799.9, 512).astype(np.float32)
    ubo = rs.ubo()
    l4 = rng.uniform(390, 720, (256, 4)).astype(np.float32)
    temps = rng.uniform(1500, 12000, 256).astype(np.float32)
    np.savez_compressed(os.path.join(out, 'ref_leaves.npz'), seed=seeds, pcg=sh.pcg32_n(seeds), push=push, gxyk=gxyk,
                        gseed=sh.generate_seed_n(push, gxyk), waves=waves,
                        wave_xyz=np.array([sh.wave_to_xyz(ubo, float(x)) for x in waves]),
                        sample_wl=np.array([sh.sample_wavelengths(float(x)) for x in waves]),
                        bk7=np.array([sh.bk7(float(x)) for x in waves], dtype=np.float32), l4=l4, temps=temps,
                        emit=np.array([sh.emit(l, float(t), 7.5) for l, t in zip(l4, temps)]),
                        spd=np.array([sh.spd(l, 550.0, 6.0, i & 1) for i, l in enumerate(l4)]))


if __name__ == '__main__':
    main()
