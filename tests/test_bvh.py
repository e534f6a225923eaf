"""BVH (pathtracer_b200/csrc/pt_bvh.h, pt_bvh.cpp): host-side checks, no GPU.

tests/bvh_check.cpp is compiled against the library's own builder and traversal template and compares, ray by ray,
the closest hit found through the tree with the reference's in-order scan (shader.comp:862-934): t and object index
must be bit-identical, ties included.  The GPU side of the same claim is tests/test_gpu_parity.py::test_bvh_*."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path
from oracle import pack

sys.path.insert(0, os.path.join(ROOT, 'tools'))
from make_synthetic_scenes import many_sphere_scene, mixed_scene  # noqa: E402

CSRC = os.path.join(ROOT, 'pathtracer_b200', 'csrc')


def build_checker(out, flags=()):
    cmd = ['g++', '-O2', '-std=c++17', '-mfma', '-ffp-contract=off', '-fno-fast-math', *flags, '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
           os.path.join(ROOT, 'tests', 'bvh_check.cpp'), os.path.join(CSRC, 'pt_bvh.cpp'), os.path.join(CSRC, 'pt_prepare.cpp'), '-o', out]
    subprocess.run(cmd, check=True)
    return out


@pytest.fixture(scope='module')
def checker(tmp_path_factory):
    return build_checker(str(tmp_path_factory.mktemp('bvh') / 'bvh_check'))


def run_checker(checker, tmp_path, scene, rays, seed, cam):
    ubo = pack.pack_ubo(scene)
    path = str(tmp_path / 'ubo.bin')
    ubo.astype(np.float32).tofile(path)
    r = subprocess.run([checker, path, str(rays), str(seed)] + [repr(float(c)) for c in cam], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    words = r.stdout.split()
    return {words[i]: float(words[i + 1]) for i in range(0, len(words), 2)}


@pytest.mark.parametrize('n,dups', [(169, 0), (169, 12), (12, 0), (2, 0)])
def test_tree_equals_scan_on_sphere_carpets(checker, tmp_path, n, dups):
    scene = many_sphere_scene(n, duplicates=dups)
    cam = scene['camera']['position'][0]
    res = run_checker(checker, tmp_path, scene, 300000, 7 + n, cam)
    assert res['mismatches'] == 0 and res['prims'] == n
    if n >= 169:
        assert res['visits_near'] < 6.0, 'the tree should test a handful of the %d spheres per ray, not %.1f' % (n, res['visits_near'])


def test_tree_equals_scan_on_a_mixed_scene(checker, tmp_path):
    scene = mixed_scene()
    res = run_checker(checker, tmp_path, scene, 400000, 3, [25.0, 12.0, -18.0])
    assert res['mismatches'] == 0 and res['prims'] == 30 + 20 + 12     # the 12 cyclides stay outside the tree


def test_while_while_loop_shape_finds_the_same_hits(tmp_path):
    """pt_bvh_traverse's alternative loop shape (PT_BVH_WHILE_WHILE, an A/B knob of the kernels)."""
    ww = build_checker(str(tmp_path / 'bvh_check_ww'), ['-DPT_BVH_WHILE_WHILE=1'])
    res = run_checker(ww, tmp_path, mixed_scene(), 300000, 5, [25.0, 12.0, -18.0])
    assert res['mismatches'] == 0


@pytest.mark.parametrize('scene_fn,w,h,spp', [(lambda: many_sphere_scene(169), 160, 120, 4), (lambda: mixed_scene(), 160, 120, 4),
                                               (lambda: mixed_scene(23), 128, 96, 4), (lambda: many_sphere_scene(40, duplicates=6), 128, 96, 2)])
def test_oracle_through_the_tree_renders_the_same_image(scene_fn, w, h, spp):
    """liboracle_bvh.so = the oracle's own sphere / box / lens / cyclide routines at the leaves of the PRODUCT's tree
    (pt_bvh.cpp + pt_bvh_traverse): whole renders, every camera shot, must equal the plain oracle's in-order scan bit
    for bit.  This is the check that found the cyclide solver's mid-air roots (pt_bvh.h) and the far-origin noise."""
    from oracle import oracle
    scene = scene_fn()
    ubo = pack.pack_ubo(scene)
    for shot in range(1, int(scene['camera']['numShots']) + 1):
        p = pack.pack_params(scene, shot, w, h, spp, 5)
        scan = oracle.Oracle(ubo).render(p, spp, spp)
        tree = oracle.Oracle(ubo, bvh=True).render(p, spp, spp)
        diff = scan.view(np.uint32) != tree.view(np.uint32)
        assert not diff.any(), '%d floats differ, first at %s' % (int(diff.sum()), np.argwhere(diff)[0].tolist())
        assert np.isfinite(scan).all() and scan[..., 1].mean() > 0


def test_shipped_scenes_are_not_affected():
    """Every shipped scene has fewer bounded primitives than the default threshold: they keep the reference's scan."""
    for i in range(11):
        sc = pack.load_scene(scene_path('scene%d' % i))
        n = sum(len(sc.get(k, [])) for k in ('sphere', 'box', 'lens'))
        assert n < 12, (i, n)


def test_bvh_kernel_compiles_in_every_mode(ptlib):
    """NVRTC needs no GPU: the BVH variant of the megakernel and of the wavefront kernels builds for sm_100a."""
    import ctypes as C
    from pathtracer_b200 import api
    L = ptlib.lib()
    L.pt_kernel_compile_check.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
    ubo = ptlib.Scene.parse(json.dumps(mixed_scene())).pack_ubo()
    for mode in (4, 5, 7):          # bit 0 fast, bit 1 wavefront kernels too, bit 2 BVH
        for bake in (0, 1):
            rc = L.pt_kernel_compile_check(ubo.ctypes.data_as(C.c_void_p), api._c_strings([]), 0, mode, bake)
            assert rc == 0, L.pt_last_error(None)
