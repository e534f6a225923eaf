"""Surface extensions (mirror / glossy / dielectric surfaces; include/pt_abi.h pt_surface_ext).

NOT reference behaviour -- shader.comp shades every surface as a Lambertian (1075-1091) and the reference's TODO.md:2
lists specular / glossy materials as future work -- so there is nothing of the reference's to compare with: PARITY
UNPINNED for this file.  What can be checked is checked:
  * off by default: no shipped scene carries an extension, an all-reference table changes nothing, and the kernels of
    such scenes are built without the extension code;
  * the oracle's statement of the three lobes against the physics they claim (law of reflection, Snell's law, the
    Fresnel equations in float64, total internal reflection, BK7 dispersion, sample / eval consistency and the energy
    bound of the GGX lobe);
  * the CUDA kernel source against that oracle, bit for bit in strict mode: here on the host SIMT emulator, under
    `-m gpu` on the device through the C ABI, for every driver and the wavefront pipeline; fast mode is gated for bias.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, SCENES, scene_path
from oracle import oracle, pack

EXT_SCENES = ['surfaces_ext', 'surfaces_ext_sdf']
MIRROR, GLOSSY, DIELECTRIC = 1, 2, 3


def ext_path(name):
    return os.path.join(ROOT, 'scenes_synthetic', name + '.json')


def ext_inputs(name, w, h, spf, pl):
    scene = pack.load_scene(ext_path(name))
    ubo = pack.pack_ubo(scene)
    src = pack.sdf_sources(scene)
    return scene, ubo, pack.pack_params(scene, 1, w, h, spf, pl), src, pack.surface_ext(scene)


def ext_table(*rows):
    t = np.zeros(len(rows), dtype=pack.SURFACE_EXT_DTYPE)
    for i, (b, r, n) in enumerate(rows):
        t[i]['bsdf'], t[i]['roughness'], t[i]['ior'] = b, r, n
    return t


# ---- off by default ----------------------------------------------------------------------------------------------------
def test_no_shipped_scene_has_extensions(ptlib):
    for name in SCENES:
        assert pack.surface_ext(pack.load_scene(scene_path(name))).size == 0
        assert ptlib.Scene.load(scene_path(name)).surface_ext().size == 0


@pytest.mark.parametrize('name', ['scene1', 'scene10'])
def test_all_reference_table_changes_nothing(name):
    scene = pack.load_scene(scene_path(name))
    ubo, src, p = pack.pack_ubo(scene), pack.sdf_sources(scene), pack.pack_params(scene, 1, 40, 24, 2, 5)
    ref = oracle.Oracle(ubo, src).render(p, 4, 2)
    same = oracle.Oracle(ubo, src, surface_ext=ext_table((0, 0.5, 1.5), (0, 0, 0), (0, 0, 0))).render(p, 4, 2)
    assert np.array_equal(ref.view(np.uint32), same.view(np.uint32))


def test_scene_schema_round_trip(ptlib, tmp_path):
    """The optional material keys survive load -> save -> load in the product's loader, agree with the oracle's own
    parser, and leave the reference's uniform block untouched (the extension travels beside it, not in it)."""
    for name in EXT_SCENES:
        sc = ptlib.Scene.load(ext_path(name))
        t = sc.surface_ext()
        assert np.array_equal(t, pack.surface_ext(pack.load_scene(ext_path(name)))) and (t['bsdf'] != 0).any()
        again = ptlib.Scene.parse(sc.to_json())
        assert np.array_equal(again.surface_ext(), t) and np.array_equal(again.pack_ubo(), sc.pack_ubo())
        stripped = json.load(open(ext_path(name)))
        for m in stripped['material']:
            for k in ('bsdf', 'roughness', 'ior'):
                m.pop(k, None)
        q = tmp_path / 'stripped.json'
        q.write_text(json.dumps(stripped))
        assert np.array_equal(ptlib.Scene.load(str(q)).pack_ubo(), sc.pack_ubo())
    with pytest.raises(ptlib.PtError):
        ptlib.Scene.parse(open(ext_path('surfaces_ext')).read().replace('"mirror"', '"chrome"'))


def test_kernels_without_extensions_do_not_contain_them(ptlib):
    """The JIT defines PT_EXT_BSDF only for scenes that use an extension: the translation unit of a reference scene is
    the one it was before the extension existed, so its cubin -- size, registers, arithmetic -- cannot have changed."""
    sc = ptlib.Scene.load(scene_path('scene1'))
    base = ptlib.kernel_compile_check(sc.pack_ubo(), sc.sdf_sources, ptlib.MODE_FAST, True)
    ext = ptlib.kernel_compile_check(sc.pack_ubo(), sc.sdf_sources, ptlib.MODE_FAST, True, surface_ext=True)
    assert 'Used' in base and 'Used' in ext
    src = open(os.path.join(ROOT, 'pathtracer_b200', 'csrc', 'pt_jit.cpp')).read()
    assert 'if (opt.surface_ext) src += "#define PT_EXT_BSDF 1\\n";' in src


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('options,wavefront', [({'sched': 0}, False), ({'sched': 5}, False), ({'sched': 7}, False), ({'sched': 8}, False), ({}, True)])
def test_extension_kernels_build_for_every_driver(ptlib, mode, options, wavefront):
    sc = ptlib.Scene.load(ext_path('surfaces_ext_sdf'))
    log = ptlib.kernel_compile_check(sc.pack_ubo(), sc.sdf_sources, mode, True, options, wavefront=wavefront, surface_ext=True)
    assert 'Used' in log and ' 0 bytes spill stores' in log or 'spill' in log


# ---- the oracle's lobes against the physics -------------------------------------------------------------------------------
def _sample(ext_row, d, n, R=(1, 1, 1, 1), l=(450.0, 520.0, 600.0, 680.0), seed=1, inside=0):
    L = oracle.lib()
    L.oracle_surface_ext_sample.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.c_void_p]
    t = ext_table(ext_row)
    d, n = np.asarray(d, np.float32), np.asarray(n, np.float32)
    R, l = np.asarray(R, np.float32), np.asarray(l, np.float32)
    s, ins, out = C.c_uint32(seed), C.c_int(inside), np.zeros(8, np.float32)
    dead = L.oracle_surface_ext_sample(t.ctypes.data, d.ctypes.data, n.ctypes.data, R.ctypes.data, l.ctypes.data, C.byref(s), C.byref(ins),
                                       out.ctypes.data)
    return dict(dir=out[:3].astype(np.float64), weight=out[3:7].astype(np.float64), pdf=float(out[7]), dead=bool(dead), seed=s.value,
                inside=ins.value)


def _unit(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def _incident(theta_deg, n=(0.0, 0.0, 1.0)):
    th = np.radians(theta_deg)
    return np.array([np.sin(th), 0.0, -np.cos(th)])


def test_mirror_obeys_the_law_of_reflection():
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = _unit(rng.normal(size=3))
        d = _unit(rng.normal(size=3))
        if d @ n > 0:
            d = -d
        R = rng.uniform(0, 1, 4)
        r = _sample((MIRROR, 0, 0), d, n, R)
        want = d - 2 * (d @ n) * n
        assert np.allclose(r['dir'], want, atol=2e-6) and abs(np.linalg.norm(r['dir']) - 1) < 1e-5
        assert np.allclose(r['weight'], R.astype(np.float32)) and r['pdf'] == 0 and not r['dead'] and r['seed'] == 1  # no draw


def fresnel64(cosi, n1, n2):
    sin2t = (n1 / n2) ** 2 * (1 - cosi ** 2)
    if sin2t >= 1:
        return 1.0
    cost = np.sqrt(1 - sin2t)
    rs = (n1 * cosi - n2 * cost) / (n1 * cosi + n2 * cost)
    rp = (n2 * cosi - n1 * cost) / (n2 * cosi + n1 * cost)
    return 0.5 * (rs * rs + rp * rp)


@pytest.mark.parametrize('theta,inside,ior', [(0, 0, 1.5), (35, 0, 1.5), (70, 0, 1.33), (85, 0, 1.5), (20, 1, 1.5), (40, 1, 1.5), (41.5, 1, 1.5), (60, 1, 1.5)])
def test_dielectric_follows_snell_and_fresnel(theta, inside, ior):
    n = np.array([0.0, 0.0, 1.0])
    d = _incident(theta)
    n1, n2 = (ior, 1.0) if inside else (1.0, ior)
    F = fresnel64(np.cos(np.radians(theta)), n1, n2)
    N, refl = 4000, 0
    for seed in range(1, N + 1):
        r = _sample((DIELECTRIC, 0, ior), d, n, R=(0.9, 0.5, 0.2, 1.0), seed=seed, inside=inside)
        assert not r['dead'] and r['pdf'] == 0 and r['seed'] != seed  # exactly one draw, a delta lobe
        o = r['dir']
        assert abs(np.linalg.norm(o) - 1) < 1e-5 and abs(o[1]) < 1e-6  # stays in the plane of incidence
        if o[2] > 0:  # reflected
            refl += 1
            assert np.allclose(o, d - 2 * (d @ n) * n, atol=2e-6) and r['inside'] == inside
            assert np.allclose(r['weight'], 1.0)
        else:         # refracted: n1 sin(i) = n2 sin(t); the path changes sides; tinted by the material's spectrum
            assert abs(n1 * np.sin(np.radians(theta)) - n2 * np.hypot(o[0], o[1])) < 3e-6 and o[0] >= 0
            assert r['inside'] == 1 - inside
            assert np.allclose(r['weight'], np.float32([0.9, 0.5, 0.2, 1.0]))
    if F >= 1.0:
        assert refl == N  # total internal reflection beyond asin(1/1.5) = 41.8 degrees
    else:
        assert abs(refl / N - F) < 4 * np.sqrt(F * (1 - F) / N) + 1e-3, (refl / N, F)
    if theta == 0:
        assert abs(F - ((ior - 1) / (ior + 1)) ** 2) < 1e-12


def test_dielectric_bk7_disperses_with_the_hero_wavelength():
    """ior = 0: BK7 by the reference's own Sellmeier fit (shader.comp:1064-1073) at the hero wavelength l.w, the wavelength
    the reference's camera lens refracts the whole bundle with (TracePathLens)."""
    n, d = np.array([0.0, 0.0, 1.0]), _incident(50)
    angles = []
    for hero in (400.0, 550.0, 700.0):
        lam = hero * 1e-3
        n_bk7 = np.sqrt(1 + 1.03961212 * lam ** 2 / (lam ** 2 - 6.00069867e-3) + 0.231792344 * lam ** 2 / (lam ** 2 - 2.00179144e-2)
                        + 1.01046945 * lam ** 2 / (lam ** 2 - 1.03560653e2))
        got = None
        for seed in range(1, 50):
            r = _sample((DIELECTRIC, 0, 0.0), d, n, l=(450.0, 500.0, 600.0, hero), seed=seed)
            if r['dir'][2] < 0:
                got = np.hypot(r['dir'][0], r['dir'][1])
                break
        assert got is not None and abs(np.sin(np.radians(50)) - n_bk7 * got) < 3e-6
        angles.append(got)
    assert angles[0] < angles[1] < angles[2]  # blue bends more: smaller sine of the refraction angle


@pytest.mark.parametrize('roughness,theta', [(0.1, 20), (0.35, 40), (0.6, 10), (0.8, 65), (1.0, 30)])
def test_glossy_sample_matches_eval_and_conserves_energy(roughness, theta):
    """weight == f(i, o) cos / pdf for the sampled direction (sampling and evaluation state the same lobe), the pdf
    integrates to at most one over the hemisphere, and a white surface never reflects more than it receives."""
    L = oracle.lib()
    L.oracle_ggx_eval.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_void_p, C.c_void_p]
    n = np.array([0.0, 0.0, 1.0], np.float32)
    d = _incident(theta).astype(np.float32)
    i = (-d).astype(np.float32)
    R = np.array([1.0, 0.8, 0.3, 0.04], np.float32)
    ws, alive, inv_pdf = [], 0, []
    N = 6000
    for seed in range(1, N + 1):
        r = _sample((GLOSSY, roughness, 0), d, n, R=R, seed=seed)
        if r['dead']:
            assert np.all(r['weight'] == 0)
            ws.append(np.zeros(4))
            continue
        alive += 1
        o = r['dir'].astype(np.float32)
        assert o[2] > 0 and abs(np.linalg.norm(o) - 1) < 1e-5 and r['pdf'] > 0
        f = np.zeros(4, np.float32)
        L.oracle_ggx_eval(i.ctypes.data, o.ctypes.data, n.ctypes.data, roughness, R.ctypes.data, f.ctypes.data)
        # D(h) has the cancellation 1 - (n.h)^2 (1 - alpha^2) at its peak: in fp32 the half vector rebuilt from (i, o) and
        # the sampled one agree to ~1e-7, D to ~1e-3 at alpha = 0.01, to ~1e-5 at alpha >= 0.1
        assert np.allclose(f.astype(np.float64) * o[2] / r['pdf'], r['weight'], rtol=5e-3 if roughness <= 0.1 else 5e-4, atol=1e-6)
        ws.append(r['weight'])
    ws = np.array(ws)
    albedo = ws.mean(axis=0)
    err = 4 * ws.std(axis=0) / np.sqrt(N)
    assert (albedo <= 1.0 + err).all() and albedo[0] > 0.25  # single scattering loses energy (to ~1/3 at alpha = 1), never gains
    assert albedo[0] >= albedo[1] >= albedo[2] >= albedo[3]  # F0 orders the channels
    if roughness <= 0.1:  # nearly a mirror: the lobe hugs the reflection direction, Fresnel of a white conductor is 1
        assert albedo[0] > 0.95


def test_glossy_tends_to_the_mirror():
    n, d = np.array([0.0, 0.0, 1.0]), _incident(30)
    want = d - 2 * (d @ n) * n
    devs = [np.degrees(np.arccos(np.clip(_sample((GLOSSY, 0.02, 0), d, n, seed=s)['dir'] @ want, -1, 1))) for s in range(1, 200)]
    assert np.median(devs) < 0.1


# ---- the kernel source against the oracle, on the host SIMT emulator -------------------------------------------------------
from test_simt_emulation import DRIVERS, build_emulator, emulate  # noqa: E402


@pytest.mark.parametrize('driver,name,w,h,spp,spf,pl', [
    ('v1', 'surfaces_ext', 48, 32, 6, 3, 8), ('v3s_table3', 'surfaces_ext', 48, 32, 6, 3, 8), ('v2s_table16', 'surfaces_ext', 40, 24, 4, 4, 12),
    ('v1', 'surfaces_ext_sdf', 32, 20, 2, 2, 8), ('v2s_table2', 'surfaces_ext_sdf', 32, 20, 4, 2, 8), ('v2m', 'surfaces_ext_sdf', 32, 20, 4, 4, 8),
    # with the generation / resolve kernels, the configuration the JIT picks for this scene (it has a cyclide)
    ('v2s_pregen', 'surfaces_ext_sdf', 32, 20, 4, 2, 8), ('v2s_resolve', 'surfaces_ext_sdf', 32, 20, 4, 4, 8), ('v3s_resolve', 'surfaces_ext', 48, 32, 6, 3, 8)])
def test_emulated_strict_kernel_equals_the_oracle_with_extensions(ptlib, driver, name, w, h, spp, spf, pl):
    scene, ubo, p, src, table = ext_inputs(name, w, h, spf, pl)
    defs = dict(DRIVERS[driver])
    defs['PT_EXT_BSDF'] = 1
    L = build_emulator(ptlib, defs, src, ubo[pack.OFF_SDF:pack.OFF_SDF + 6 * len(src)])
    L.simt_set_surface_ext.argtypes = [C.c_void_p, C.c_int]
    L.simt_set_surface_ext(table.ctypes.data, int(table.size))
    got = emulate(L, ubo, p, spp, spf)
    ref = oracle.Oracle(ubo, src, surface_ext=table).render(p, spp, spf)
    plain = oracle.Oracle(ubo, src).render(p, spp, spf)
    assert not np.array_equal(ref, plain)  # the extensions are really in the picture
    assert np.isfinite(ref).all()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), '%d floats differ' % int((got != ref).sum())


# ---- on the device, through the C ABI ------------------------------------------------------------------------------------
def gpu_render_ext(ptlib, name, w, h, spp, spf, pl, mode, options=None, pipeline=0, table='scene', jit=2):
    sc = ptlib.Scene.load(ext_path(name))
    ubo, p = sc.pack_ubo(), sc.pack_params(1, w, h, spf, pl)
    r = ptlib.Renderer(device=0, mode=mode, jit=jit, pipeline=pipeline, options=options or {})
    t = sc.surface_ext() if isinstance(table, str) else table
    r.set_surface_ext(t)
    r.set_scene(ubo, sc.sdf_sources)
    r.resize(w, h)
    r.render(p, spp, spf)
    got = r.read_xyz()
    r.close()
    return got, ubo, p, [s.decode() for s in sc.sdf_sources], t


@pytest.mark.gpu
@pytest.mark.parametrize('options,pipeline,name,w,h,spp,spf,pl', [
    ({'sched': 0}, 0, 'surfaces_ext', 96, 64, 8, 4, 8), ({'sched': 7, 'steal_s': 16}, 0, 'surfaces_ext', 96, 64, 8, 4, 8),
    ({'sched': 5, 'steal_s': 16}, 0, 'surfaces_ext', 70, 45, 6, 3, 32), ({}, 0, 'surfaces_ext', 160, 90, 4, 2, 5),
    ({'sched': 5, 'steal_s': 16}, 0, 'surfaces_ext_sdf', 96, 64, 4, 2, 8), ({'sched': 8}, 0, 'surfaces_ext_sdf', 70, 45, 6, 3, 12),
    ({'sched': 0}, 0, 'surfaces_ext_sdf', 64, 48, 2, 2, 8), ({}, 1, 'surfaces_ext', 96, 64, 8, 4, 8), ({}, 1, 'surfaces_ext_sdf', 64, 48, 4, 2, 8)])
def test_cuda_strict_bit_exact_with_extensions(ptlib, options, pipeline, name, w, h, spp, spf, pl):
    got, ubo, p, src, t = gpu_render_ext(ptlib, name, w, h, spp, spf, pl, ptlib.MODE_STRICT, options, pipeline)
    ref = oracle.Oracle(ubo, src, surface_ext=t).render(p, spp, spf)
    assert np.isfinite(got).all()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), '%d floats differ' % int((got != ref).sum())


@pytest.mark.gpu
def test_wavefront_material_sorted_shading(ptlib):
    """The wavefront pipeline groups SHADE's path rays by lobe (option wf_sort; off by default: it measured slower): the same image bit for bit with and without the sort, strict and fast -- only the order in which paths
    are shaded changes -- and equal to the oracle's in strict mode."""
    import time
    for name, w, h, spp, spf, pl in (('surfaces_ext', 160, 90, 8, 4, 8), ('surfaces_ext_sdf', 96, 64, 4, 2, 8)):
        for mode in (ptlib.MODE_STRICT, ptlib.MODE_FAST):
            images, secs = [], []
            for sort in (1, 0):
                t0 = time.perf_counter()
                got, ubo, p, src, t = gpu_render_ext(ptlib, name, w, h, spp, spf, pl, mode, {'wf_sort': sort}, pipeline=1)
                secs.append(time.perf_counter() - t0)
                images.append(got)
            assert np.array_equal(images[0].view(np.uint32), images[1].view(np.uint32)), (name, mode)
            if mode == ptlib.MODE_STRICT:
                ref = oracle.Oracle(ubo, src, surface_ext=t).render(p, spp, spf)
                assert np.array_equal(images[0].view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_cuda_reference_shading_returns_when_the_table_is_cleared(ptlib):
    """set_surface_ext(None) after an extended scene: the next set_scene builds the reference kernel again, and an
    all-reference table equals no table."""
    sc = ptlib.Scene.load(ext_path('surfaces_ext'))
    ubo, p = sc.pack_ubo(), sc.pack_params(1, 96, 64, 4, 5)
    ref = oracle.Oracle(ubo, []).render(p, 8, 4)
    r = ptlib.Renderer(device=0, mode=ptlib.MODE_STRICT, jit=2)
    images = []
    for t in (sc.surface_ext(), None, ext_table((0, 0.3, 1.5), (0, 0, 0))):
        r.set_surface_ext(t)
        r.set_scene(ubo, sc.sdf_sources)
        r.resize(96, 64)
        r.clear()
        r.render(p, 8, 4)
        images.append(r.read_xyz())
    assert not np.array_equal(images[0], ref)
    assert np.array_equal(images[1].view(np.uint32), ref.view(np.uint32)) and np.array_equal(images[2].view(np.uint32), ref.view(np.uint32))
    # argument checking, and the static kernels refuse rather than ignore the table
    for bad in (ext_table((4, 0, 0)), ext_table((2, 1.5, 0)), ext_table((3, 0, 0.5)), np.zeros(65, dtype=pack.SURFACE_EXT_DTYPE)):
        with pytest.raises(ptlib.PtError):
            r.set_surface_ext(bad)
    r.set_jit(0)
    r.set_surface_ext(sc.surface_ext())
    with pytest.raises(ptlib.PtError):
        r.set_scene(ubo, sc.sdf_sources)
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize('name,pl', [('surfaces_ext', 8), ('surfaces_ext_sdf', 8)])
def test_cuda_fast_mode_unbiased_with_extensions(ptlib, name, pl):
    """The benchmarked build against strict on the extended scenes: image means within 3 sigma of the Monte-Carlo error
    and within 0.5 % (the gate of tests/test_gpu_parity.py::test_fast_mode_is_unbiased_against_strict)."""
    w, h, spp, spf = 96, 64, 4096, 256
    sc = ptlib.Scene.load(ext_path(name))
    ubo, p, t = sc.pack_ubo(), sc.pack_params(1, w, h, spf, pl), sc.surface_ext()
    out = []
    for mode, first in ((ptlib.MODE_STRICT, 0), (ptlib.MODE_FAST, 1 << 20)):
        r = ptlib.Renderer(device=0, mode=mode, jit=2)
        r.set_surface_ext(t)
        r.set_scene(ubo, sc.sdf_sources)
        r.resize(w, h)
        means = []
        for j in range(spp // spf):
            r.clear()
            r.dispatch_sum(p, first + j * spf, spf)
            r.finalize(p, spf)
            img = r.read_xyz()
            assert np.isfinite(img).all()
            means.append(img[..., :3].astype(np.float64).mean(axis=(0, 1)))
        r.close()
        means = np.array(means)
        out.append((means.mean(axis=0), means.std(axis=0, ddof=1) / np.sqrt(len(means))))
    (ms, ss), (mf, sf) = out
    z = np.abs(mf - ms) / np.sqrt(ss ** 2 + sf ** 2)
    rel = np.abs(mf / ms - 1.0)
    print('%s: strict %s fast %s z %s rel %s' % (name, ms, mf, np.round(z, 2), np.round(rel, 5)))
    assert (z < 3.0).all() or (rel < 0.001).all(), (z, rel)
    assert (rel < 0.005).all(), rel
