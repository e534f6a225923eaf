"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py from the CPU oracle): the oracle must
keep reproducing them bit for bit (CPU), and so must the CUDA kernels in strict mode (GPU, both pipelines)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, scene_path
from oracle import oracle, pack

FILES = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'scene*.npz')))


def case(path):
    d = np.load(path)
    shot, w, h, spp, spf, pl = (int(v) for v in d['meta'])
    return os.path.basename(path).split('_')[0], shot, w, h, spp, spf, pl, d['xyz']


def test_fixtures_exist():
    assert len(FILES) >= 8


def test_pcg32_fixture():
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'pcg32.npz'))
    L = oracle.lib()
    assert [L.oracle_pcg32(int(s)) for s in d['seed']] == [int(v) for v in d['out']]


@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(path):
    name, shot, w, h, spp, spf, pl, want = case(path)
    o, scene = oracle.from_scene_file(scene_path(name))
    got = o.render(pack.pack_params(scene, shot, w, h, spf, pl), spp, spf)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize('pipeline', [0, 1], ids=['megakernel', 'wavefront'])
@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_reproduces_golden(ptlib, path, pipeline):
    name, shot, w, h, spp, spf, pl, want = case(path)
    sc = ptlib.Scene.load(scene_path(name))
    r = ptlib.Renderer(device=0, mode=ptlib.MODE_STRICT, pipeline=pipeline)
    r.set_scene(sc.pack_ubo(), sc.sdf_sources)
    r.resize(w, h)
    r.render(sc.pack_params(shot, w, h, spf, pl), spp, spf)
    got = r.read_xyz()
    r.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
