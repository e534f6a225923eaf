"""The product's kernel source on a host SIMT emulator (tests/simt/cuda_shim.h): driver logic checked WITHOUT a GPU.

pathtracer_b200/csrc/pt_kernel.cuh is compiled, unmodified, for the host: one thread per CUDA thread, barriers for
__syncthreads / __ballot_sync / __shfl_sync / __syncwarp, statics for shared memory.  In strict mode the arithmetic is
the oracle's, so every driver -- the nested loops of v1, the flat loop with the sample pool (v3s), the phase machine with
the sample pool (v2s: table rounds) and with the pool of parked marching paths (v2m: full, tiny and overflowing pools)
-- must reproduce the oracle bit for bit, and the whole-dispatch pools (PT_STEAL_S = 0) up to fp32 summation order.
What this does not cover: the device compiler, fast-math builds, and
timing -- the `-m gpu` parity tests remain the gate for those.  TEST INFRASTRUCTURE: a checker, not a rendering path."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, scene_path
from oracle import oracle, pack, sdf_build

SIMT = os.path.join(ROOT, 'tests', 'simt')
BUILD = os.path.join(SIMT, '_build')
CSRC = os.path.join(ROOT, 'pathtracer_b200', 'csrc')
FLAGS = ['-std=c++20', '-O1', '-fPIC', '-ffp-contract=off', '-fno-fast-math', '-mfma', '-mavx2', '-Wno-unknown-pragmas',
         '-I' + SIMT, '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]


def build_emulator(ptlib, defines, sdf_sources=(), sdf_raw=None):
    """g++ the emulated kernel for one (driver, knobs, SDF unit); cached by content."""
    from pathtracer_b200 import api
    os.makedirs(BUILD, exist_ok=True)
    defs = dict(defines)
    objs = []
    unit = ''
    if sdf_sources:
        unit = api.sdf_translate(list(sdf_sources), sdf_raw)
        defs['PT_HAS_SDF'] = 1
    srcs = [os.path.join(SIMT, 'simt_main.cpp'), os.path.join(SIMT, 'cuda_shim.h'), os.path.join(CSRC, 'pt_kernel.cuh'), os.path.join(CSRC, 'pt_driver_v2m.cuh'),
            os.path.join(CSRC, 'pt_prepare.cpp'), os.path.join(CSRC, 'pt_dev_scene.h'), os.path.join(ROOT, 'include', 'pt_math.h')]
    h = hashlib.sha1((repr(sorted(defs.items())) + unit + ''.join(open(f).read() for f in srcs)).encode()).hexdigest()[:16]
    so = os.path.join(BUILD, 'simt_%s.so' % h)
    if not os.path.exists(so):
        if unit:
            cpp = os.path.join(BUILD, 'sdf_%s.cpp' % h)
            open(cpp, 'w').write(unit)
            obj = cpp[:-4] + '.o'
            subprocess.run([sdf_build.CXX, *[f for f in sdf_build.CXXFLAGS if f != '-shared'], '-c', '-o', obj, cpp], check=True)
            objs.append(obj)
        ph = hashlib.sha1(''.join(open(f).read() for f in srcs[4:]).encode()).hexdigest()[:16]
        prep = os.path.join(BUILD, 'pt_prepare_%s.o' % ph)   # the product's host-side preparation, compiled once
        if not os.path.exists(prep):
            subprocess.run(['g++', *FLAGS, '-c', os.path.join(CSRC, 'pt_prepare.cpp'), '-o', prep + '.tmp.o'], check=True)
            os.replace(prep + '.tmp.o', prep)
        cmd = ['g++', *FLAGS, *['-D%s=%s' % kv for kv in sorted(defs.items())], os.path.join(SIMT, 'simt_main.cpp'),
               prep, *objs, '-shared', '-o', so, '-lpthread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(so)
    L.simt_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return L


def emulate(L, ubo, p, spp, spf, ctas=2):
    """Renderer.render's bookkeeping (host:4042-4048) over the emulated pt_dispatch."""
    q = np.array(p, copy=True)
    W, H = int(np.ravel(q['resolution'])[0]), int(np.ravel(q['resolution'])[1])
    img = np.zeros((H, W, 4), dtype=np.float32)
    for j in range(1, spp // spf + 1):
        q['frame'] = j * spf
        q['currentSamples'] = j * spf
        q['samplesPerFrame'] = spf
        qq = np.ascontiguousarray(q)
        assert L.simt_dispatch(ubo.ctypes.data_as(C.c_void_p), qq.ctypes.data_as(C.c_void_p), 0, 0, 0,
                               img.ctypes.data_as(C.c_void_p), ctas) == 0
    return img


def scene_inputs(name, w, h, spf, pl):
    path = scene_path(name) if not name.startswith('synthetic/') else os.path.join(ROOT, 'scenes_synthetic', name.split('/', 1)[1] + '.json')
    scene = pack.load_scene(path)
    ubo = pack.pack_ubo(scene)
    src = pack.sdf_sources(scene)
    return ubo, pack.pack_params(scene, 1, w, h, spf, pl), src, ubo[pack.OFF_SDF:pack.OFF_SDF + 6 * len(src)]


DRIVERS = {
    'v1': {'PT_SCHED': 0},
    'v2s_table16': {'PT_SCHED': 5, 'PT_STEAL_S': 16}, 'v2s_table2': {'PT_SCHED': 5, 'PT_STEAL_S': 2},
    'v3s_table16': {'PT_SCHED': 7, 'PT_STEAL_S': 16}, 'v3s_table3': {'PT_SCHED': 7, 'PT_STEAL_S': 3, 'PT_REGEN_T': 4},
    # v2m: the pool of parked paths at its default size, with two slots (rays mostly march in their lanes), with a
    # threshold no warp reaches (the phase only runs when the feeders are dry) and with a greedy one
    # v2s with the box / lens / cyclide tests as a phase of their own (immediately / once eight lanes wait)
    'v2s_heavy1': {'PT_SCHED': 5, 'PT_STEAL_S': 16, 'PT_HEAVY_MIN': 1}, 'v2s_heavy8': {'PT_SCHED': 5, 'PT_STEAL_S': 4, 'PT_HEAVY_MIN': 8},
    # camera rays from the generation kernel's records (option pregen), in bands of two CTA rows (emulate()'s ctas)
    'v2s_pregen': {'PT_SCHED': 5, 'PT_STEAL_S': 16, 'PT_PREGEN': 1}, 'v3s_pregen': {'PT_SCHED': 7, 'PT_STEAL_S': 3, 'PT_REGEN_T': 4, 'PT_PREGEN': 1},
    # ... and radiance -> XYZ plus the per-pixel sums in the resolve kernel (option resolve): one round, no table, sample order
    'v2s_resolve': {'PT_SCHED': 5, 'PT_STEAL_S': 0, 'PT_PREGEN': 1, 'PT_RESOLVE': 1}, 'v3s_resolve': {'PT_SCHED': 7, 'PT_STEAL_S': 0, 'PT_REGEN_T': 4, 'PT_PREGEN': 1, 'PT_RESOLVE': 1},
    'v2m': {'PT_SCHED': 8, 'PT_STEAL_S': 4, 'PT_POOL_CAP': 32, 'PT_POOL_MIN': 24},
    'v2m_tiny_pool': {'PT_SCHED': 8, 'PT_STEAL_S': 3, 'PT_POOL_CAP': 2, 'PT_POOL_MIN': 2},
    'v2m_lazy': {'PT_SCHED': 8, 'PT_STEAL_S': 8, 'PT_POOL_CAP': 7, 'PT_POOL_MIN': 64},
    'v2m_greedy': {'PT_SCHED': 8, 'PT_STEAL_S': 2, 'PT_POOL_CAP': 32, 'PT_POOL_MIN': 1, 'PT_SDF_REPS': 3},
}
V2M = [d for d in sorted(DRIVERS) if d.startswith('v2m')]


# scenes without SDFs share one build per driver (v2m is v2s there); of the SDF scenes, scene10 (menger, pathLength 32)
# runs under every driver, scene9 / scene8 / scene3 under the phase machines (each SDF unit is its own build)
CASES = [(d, 'scene0', 48, 32, 10, 5, 5) for d in sorted(DRIVERS) if d not in V2M]
CASES += [(d, 'scene1', 50, 37, 3, 3, 5) for d in sorted(DRIVERS) if d not in V2M]
CASES += [(d, 'scene10', 40, 24, 4, 2, 32) for d in sorted(DRIVERS)]
CASES += [(d, n, w, h, spp, spf, 5) for d in ['v2s_table16', 'v2s_heavy8'] + V2M
          for (n, w, h, spp, spf) in (('scene9', 33, 17, 5, 5), ('scene8', 32, 16, 2, 2), ('scene3', 24, 16, 3, 3))]


@pytest.mark.parametrize('driver,name,w,h,spp,spf,pl', CASES)
def test_emulated_strict_kernel_equals_the_oracle(ptlib, driver, name, w, h, spp, spf, pl):
    """Every driver, strict arithmetic, ragged frame sizes, several dispatches and (v2s) several rounds per dispatch:
    the emulated kernel writes the oracle's bits."""
    ubo, p, src, raw = scene_inputs(name, w, h, spf, pl)
    L = build_emulator(ptlib, DRIVERS[driver], src, raw)
    got = emulate(L, ubo, p, spp, spf)
    ref = oracle.Oracle(ubo, src).render(p, spp, spf)
    g, r = got.view(np.uint32), ref.view(np.uint32)
    assert np.array_equal(g, r), '%s on %s: %d floats differ' % (driver, name, int((g != r).sum()))


@pytest.mark.parametrize('name,w,h,spp,spf,pl', [('scene0', 61, 43, 6, 3, 5), ('scene1', 50, 37, 16, 16, 5), ('scene10', 40, 24, 4, 4, 32)])
def test_emulated_v3s_whole_dispatch_pool(ptlib, name, w, h, spp, spf, pl):
    """v3s with PT_STEAL_S = 0 (the fast-mode configuration: one pool per dispatch, sums in shared memory in schedule
    order) under strict arithmetic: equal to the oracle up to fp32 summation order on every pixel."""
    ubo, p, src, raw = scene_inputs(name, w, h, spf, pl)
    L = build_emulator(ptlib, {'PT_SCHED': 7, 'PT_STEAL_S': 0, 'PT_REGEN_T': 8}, src, raw)
    got = emulate(L, ubo, p, spp, spf)
    ref = oracle.Oracle(ubo, src).render(p, spp, spf)
    scale = float(ref[..., :3].max())
    assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6 * scale) and (got[..., 3] == 1.0).all()


@pytest.mark.parametrize('name,w,h,spp,spf,pl', [('scene10', 40, 24, 8, 8, 32), ('scene9', 33, 17, 6, 6, 5), ('scene8', 32, 16, 3, 3, 5)])
def test_emulated_v2m_whole_dispatch_pool(ptlib, name, w, h, spp, spf, pl):
    """v2m as the fast build configures it (PT_STEAL_S = 0: one pool per dispatch, sums in schedule order) under strict
    arithmetic: equal to the oracle up to fp32 summation order on every pixel -- no path is lost or run twice while it
    moves between lanes and slots."""
    ubo, p, src, raw = scene_inputs(name, w, h, spf, pl)
    L = build_emulator(ptlib, {'PT_SCHED': 8, 'PT_STEAL_S': 0}, src, raw)
    got = emulate(L, ubo, p, spp, spf)
    ref = oracle.Oracle(ubo, src).render(p, spp, spf)
    scale = float(ref[..., :3].max())
    assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6 * scale) and (got[..., 3] == 1.0).all()


EDGE = [('scene0', 5, 3, 1, 1, 5), ('scene1', 17, 9, 33, 33, 5), ('scene1', 8, 4, 2, 2, 0), ('scene2', 31, 7, 34, 17, 2), ('scene0', 16, 8, 1, 1, 100)]


@pytest.mark.parametrize('name,w,h,spp,spf,pl', EDGE)
def test_emulated_drivers_at_the_edges(ptlib, name, w, h, spp, spf, pl):
    """Frames smaller than a tile, one sample, 33 samples per dispatch (three table rounds, the last of one sample),
    pathLength 0 (a sample is finished before it starts) and 100: table drivers give the oracle's bits, the pooled ones
    (v2s / v3s with PT_STEAL_S = 0) its image up to summation order."""
    ubo, p, src, raw = scene_inputs(name, w, h, spf, pl)
    ref = oracle.Oracle(ubo, src).render(p, spp, spf)
    for driver in ('v1', 'v2s_table16', 'v3s_table3', 'v2s_pregen', 'v3s_pregen', 'v2s_resolve', 'v3s_resolve'):
        got = emulate(build_emulator(ptlib, DRIVERS[driver], src, raw), ubo, p, spp, spf)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), driver
    scale = float(ref[..., :3].max()) or 1.0
    for defs in ({'PT_SCHED': 5, 'PT_STEAL_S': 0}, {'PT_SCHED': 7, 'PT_STEAL_S': 0}, {'PT_SCHED': 7, 'PT_STEAL_S': 0, 'PT_PREGEN': 1}):
        got = emulate(build_emulator(ptlib, defs, src, raw), ubo, p, spp, spf, 3)
        assert np.allclose(got[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6 * scale) and (got[..., 3] == 1.0).all(), defs


@pytest.mark.parametrize('driver', ['v1', 'v2s_table2', 'v3s_table3'])
def test_emulated_more_than_32_sdfs(ptlib, driver):
    """40 SDFs (scenes_synthetic/sdf40.json): the bounding-box search fills set2 as well as set1 (the reference declares
    set1..set4 and its generated dispatcher lines read them, but SearchSDF only ever writes set1: shader.comp:732-738).
    Kernel (PT_SDF_WORDS = 2) and oracle agree bit for bit; SDFs 33..40 really are in the image."""
    ubo, p, src, raw = scene_inputs('synthetic/sdf40', 40, 24, 2, 5)
    assert len(src) == 40
    defs = dict(DRIVERS[driver])
    defs['PT_SDF_WORDS'] = 2
    got = emulate(build_emulator(ptlib, defs, src, raw), ubo, p, 2, 2)
    o = oracle.Oracle(ubo, src)
    ref = o.render(p, 2, 2)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # the same scene with the last eight SDFs removed renders differently: they are reachable
    ubo32 = ubo.copy()
    ubo32[5] = 32
    ref32 = oracle.Oracle(ubo32, src[:32]).render(p, 2, 2)
    assert not np.array_equal(ref32.view(np.uint32), ref.view(np.uint32))


def baked_counts(ubo, has_sdf):
    """The defines jit policy 2 adds (pt_jit.cpp): primitive counts as compile-time constants, rolled loops with SDFs."""
    n = [int(ubo[i]) for i in range(6)]
    d = {'PT_N_SPHERES_CONST': n[0], 'PT_N_PLANES_CONST': n[1], 'PT_N_BOXES_CONST': n[2], 'PT_N_LENSES_CONST': n[3],
         'PT_N_CYCLIDES_CONST': n[4], 'PT_N_SDF_CONST': n[5]}
    if has_sdf:
        d['PT_NO_UNROLL'] = 1
    return d


@pytest.mark.parametrize('driver,name,w,h,spp,spf,pl', [('v1', 'scene0', 48, 32, 4, 2, 5), ('v3s_table16', 'scene1', 70, 45, 40, 20, 5),
                                                       ('v2s_table16', 'scene10', 40, 24, 4, 2, 5), ('v3s_table3', 'scene2', 33, 17, 9, 9, 5),
                                                       ('v2m', 'scene10', 40, 24, 4, 2, 5), ('v2m_tiny_pool', 'scene9', 33, 17, 3, 3, 5)])
def test_emulated_scene_specialised_kernels(ptlib, driver, name, w, h, spp, spf, pl):
    """jit policy 2 (what bench.py uses): the primitive counts baked in as constants, offsets folded, loops unrolled (or
    rolled, with SDFs) -- other code from the same source, same bits."""
    ubo, p, src, raw = scene_inputs(name, w, h, spf, pl)
    defs = dict(DRIVERS[driver])
    defs.update(baked_counts(ubo, bool(src)))
    got = emulate(build_emulator(ptlib, defs, src, raw), ubo, p, spp, spf)
    ref = oracle.Oracle(ubo, src).render(p, spp, spf)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
