"""Known-answer tests that pin the CPU oracle.  The reference ships no tests or golden vectors
("parity unpinned"); these are first-principles checks: exact integer vectors for the PCG hash, exact float values
for functions built from + - * / sqrt only, analytic checks for the rest."""
import hashlib

import numpy as np
import pytest

from conftest import scene_path, SCENES
from oracle import oracle, pack

PCG_KAT = {0x00000000: 0x07bb2fe2, 0x00000001: 0xa8beea3c, 0x00000002: 0x7a7ecc88, 0x00000003: 0x7f0ef6bc,
           0x0000003f: 0xa59978e9, 0x00000040: 0x85e59170, 0x000003ff: 0x21cccfa5, 0x00003fff: 0x78d103d9,
           0x0000ffff: 0x07d6f4f8, 0xffffffff: 0xe62a4902, 0x12345678: 0x995312e1}


def py_pcg32(s):
    state = (s * 747796405 + 2891336453) & 0xFFFFFFFF
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return ((word >> 22) ^ word) & 0xFFFFFFFF


def test_pcg32_vectors():
    L = oracle.lib()
    for s, want in PCG_KAT.items():
        assert L.oracle_pcg32(s) == want
        assert py_pcg32(s) == want


def _params(w, h, spf, frame):
    p = np.zeros((), dtype=pack.PARAMS_DTYPE)
    p['resolution'] = (w, h)
    p['samplesPerFrame'] = spf
    p['frame'] = frame
    p['currentSamples'] = frame
    return p


@pytest.mark.parametrize('gid,k,spf,seed,states,floats', [
    ((0, 0), 0, 1, 0x07bf2fe2, [0x3a20ffe8, 0xc12aaa0d, 0x2e2aba37, 0x18b7d9d6, 0xd7b1adea],
     [0.22706604, 0.75455725, 0.18033947, 0.09655534, 0.84255493]),
    ((255, 256), 0, 1, 0x07bd30e1, [0x7ad45138, 0xa7dc0c74, 0x3710a129, 0x8a56e437, 0x0aa413b4], None),
    ((511, 511), 63, 64, 0xa5997ce8, [0x3f042acf, 0x0a7fac30, 0xa9c3fb33, 0xdcfd5027, 0x90df27c7], None),
])
def test_seed_and_first_draws(gid, k, spf, seed, states, floats):
    import ctypes as C
    L = oracle.lib()
    p = _params(512, 512, spf, spf)
    got = L.oracle_generate_seed(p.ctypes.data_as(C.c_void_p), gid[0], gid[1], k)
    assert got == seed
    s = C.c_uint32(got)
    for i, st in enumerate(states):
        f = L.oracle_random_float(C.byref(s))
        assert s.value == st
        assert f == np.float32(st) / np.float32(4294967296.0)
        if floats:
            assert abs(f - floats[i]) < 1e-7


def test_random_float_can_be_one():
    assert np.float32(0xFFFFFFFF) == np.float32(4294967296.0)
    assert np.float32(np.uint32(0xFFFFFF80)) / np.float32(4294967296.0) == np.float32(1.0)


def test_cie_table_pin():
    t = pack.cie_table()
    assert hashlib.md5(t.astype('<f4').tobytes()).hexdigest() == '534a779bb345832aa54e456e53c9b76f'
    rows = t.reshape(441, 3).astype(np.float64)
    assert np.allclose(rows.sum(axis=0), [106.86534529, 106.85687253, 106.89225085], atol=1e-6)
    assert rows[555 - 360, 1] == 1.0
    assert np.allclose(rows[550 - 360], [0.4334499, 0.9949501, 0.00875], atol=1e-7)


def test_wave_to_xyz():
    import ctypes as C
    L = oracle.lib()
    ubo = pack.pack_ubo(pack.load_scene(scene_path('scene0')))
    out = np.zeros(3, dtype=np.float32)
    L.oracle_wave_to_xyz(ubo.ctypes.data_as(C.c_void_p), 550.5, out.ctypes.data_as(C.c_void_p))
    assert np.allclose(out, [0.4411226, 0.9958304, 0.0083926], atol=2e-7)
    rows = pack.cie_table().reshape(441, 3)
    want = rows[190] * np.float32(0.5) + rows[191] * np.float32(0.5)
    assert np.array_equal(out, want)
    for w in (359.9, 800.5):
        L.oracle_wave_to_xyz(ubo.ctypes.data_as(C.c_void_p), w, out.ctypes.data_as(C.c_void_p))
        assert not out.any()


@pytest.mark.parametrize('lh,want', [(500.0, (582.5, 665.0, 417.5, 500.0)), (360.0, (442.5, 525.0, 607.5, 690.0)),
                                     (800.0, (552.5, 635.0, 717.5, 470.0))])
def test_sample_wavelengths(lh, want):
    L = oracle.lib()
    out = np.zeros(4, dtype=np.float32)
    L.oracle_sample_wavelengths(lh, out.ctypes.data)
    assert tuple(out) == want


def test_bk7_index():
    L = oracle.lib()
    assert abs(L.oracle_bk7(400.0) - 1.5308485) < 2e-6
    assert abs(L.oracle_bk7(587.56) - 1.5168) < 2e-5
    assert abs(L.oracle_bk7(700.0) - 1.513064) < 2e-6


def test_emit_at_wien_peak_equals_luminosity():
    L = oracle.lib()
    for T, lum in ((5500.0, 1.0), (3000.0, 20.0), (6500.0, 0.5)):
        peak_nm = 2.8977729e6 / T
        l4 = np.array([peak_nm] * 4, dtype=np.float32)
        out = np.zeros(4, dtype=np.float32)
        L.oracle_emit(l4.ctypes.data, T, lum, out.ctypes.data)
        assert np.allclose(out, lum, rtol=2e-4)
    # against float64 Planck at an off-peak wavelength
    T, lam = 5500.0, 450.0
    l4 = np.array([lam] * 4, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    L.oracle_emit(l4.ctypes.data, T, 1.0, out.ctypes.data)
    lm = lam * 1e-9
    want = (1.1910429724e-16 * lm ** -5 / (np.exp(0.014387768775 / (lm * T)) - 1)) / (4.0956746759e-6 * T ** 5)
    assert np.allclose(out, want, rtol=1e-4)


def test_spd():
    L = oracle.lib()
    l4 = np.array([550.0, 600.0, 500.0, 450.0], dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    L.oracle_spd(l4.ctypes.data, 550.0, 10.0, 0, out.ctypes.data)
    assert out[0] == 1.0
    want = np.exp(-(((l4.astype(np.float64) - 550.0) / 200.0) ** 2))
    assert np.allclose(out, want, rtol=1e-5)
    inv = np.zeros(4, dtype=np.float32)
    L.oracle_spd(l4.ctypes.data, 550.0, 10.0, 1, inv.ctypes.data)
    assert np.allclose(inv, 1.0 - out, atol=1e-7)


def test_rotation_matrix_is_orthonormal_and_matches_closed_form():
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(50):
        deg = rng.uniform(-360, 360, 3).astype(np.float32)
        m = np.zeros(9, dtype=np.float32)
        L.oracle_rotation_matrix(deg.ctypes.data, m.ctypes.data)
        M = m.reshape(3, 3).T.astype(np.float64)  # columns -> math matrix
        assert np.allclose(M @ M.T, np.eye(3), atol=1e-5)
        a = np.deg2rad(deg.astype(np.float64))
        sx, sy, sz, cx, cy, cz = *np.sin(a), *np.cos(a)
        mX = np.array([[1, 0, 0], [0, cx, sx], [0, -sx, cx]])
        mY = np.array([[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]])
        mZ = np.array([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1]])
        assert np.allclose(M, mX @ mY @ mZ, atol=2e-5)


def test_closed_form_intersections():
    scene = {'camera': {}, 'sphere': [{'position': [0, 0, 5], 'radius': 1.0, 'materialID': 2, 'lightID': 0}],
             'plane': [{'position': [0, -1, 0], 'materialID': 1, 'lightID': 0}],
             'box': [{'position': [4, 0, 0], 'rotation': [0, 0, 0], 'size': [2, 2, 2], 'materialID': 3, 'lightID': 0}],
             'material': [{'reflection': {'peakWavelength': 550, 'sigma': 10, 'isInvert': False}}] * 3, 'light': []}
    o = oracle.Oracle(pack.pack_ubo(scene))
    t, out = o.intersect([0, 0, 0], [0, 0, 1])
    assert abs(t - 4.0) < 1e-6 and np.allclose(out[:3], [0, 0, -1]) and out[3] == 1.0 and out[4] == -1.0
    t, out = o.intersect([0, 0, 0], [0, -1, 0])
    assert abs(t - 1.0) < 1e-6 and np.allclose(out[:3], [0, 1, 0]) and out[3] == 0.0
    t, out = o.intersect([0, 0, 0], [1, 0, 0])
    assert abs(t - 3.0) < 1e-6 and np.allclose(out[:3], [-1, 0, 0]) and out[3] == 2.0
    t, out = o.intersect([0, 0, 0], [0, 1, 0])
    assert t == np.float32(1e5)  # miss
    t, out = o.intersect([0, 0, 5], [0, 0, 1])  # from inside the sphere: normal flips
    assert abs(t - 1.0) < 1e-6 and np.allclose(out[:3], [0, 0, -1])


def test_quartic_solver_on_known_roots():
    L = oracle.lib()
    roots = np.array([1.0, 2.5, -3.0, 4.0])
    coef = np.poly(roots).astype(np.float32)  # a..e
    r = np.zeros(4, dtype=np.float32)
    real = np.zeros(4, dtype=np.int32)
    L.oracle_solve_quartic(coef.ctypes.data, r.ctypes.data, real.ctypes.data)
    assert real.all()
    assert np.allclose(np.sort(r), np.sort(roots), atol=2e-3)


def test_scene0_pack_known_values():
    ubo = pack.pack_ubo(pack.load_scene(scene_path('scene0')))
    assert list(ubo[:7]) == [3, 1, 1, 1, 1, 0, 1]
    assert list(ubo[7:13]) == [0, 1, 0, 1, 1, 0]
    assert ubo[pack.OFF_LID] == 2
    cyc = 7 + 18 + 5 + 11 + 12
    assert ubo[cyc + 13] == np.float32(2.25)  # 36 * 0.0625


def test_white_furnace_like_energy_bound():
    """A grey Lambertian sphere lit by a large emitter: every pixel stays finite and non-negative."""
    o, sc = oracle.from_scene_file(scene_path('scene0'))
    p = pack.pack_params(sc, 1, 48, 32, 8, 5)
    img = o.render(p, 8, 8)
    assert np.isfinite(img).all()
    assert (img[..., 1] >= 0).all() and img[..., 3].min() == 1.0
    assert 0.01 < img[..., 1].mean() < 10.0


def test_running_mean_equals_single_dispatch_statistics():
    """Accumulate(): 4 dispatches of 2 samples == mean of the same 8 sample indices (up to fp32 rounding)."""
    o, sc = oracle.from_scene_file(scene_path('scene1'))
    p = pack.pack_params(sc, 1, 32, 24, 2, 5)
    a = o.render(p, 8, 2)
    p8 = pack.pack_params(sc, 1, 32, 24, 8, 5)
    b = o.render(p8, 8, 8)
    assert np.allclose(a, b, rtol=2e-5, atol=1e-7)


def test_per_sample_entry_matches_dispatch():
    o, sc = oracle.from_scene_file(scene_path('scene0'))
    p = pack.pack_params(sc, 1, 16, 16, 4, 5)
    img = o.render(p, 4, 4)
    s = o.samples(p, 5, 7, 0, 4)
    acc = np.zeros(3, dtype=np.float32)
    for k in range(4):
        acc = acc + s[k]
    acc = acc / np.float32(4)
    expo = np.float32(p['apertureSize']) * np.float32(p['apertureSize']) * np.float32(int(p['ISO']))
    assert np.array_equal((acc * expo).astype(np.float32), img[7, 5, :3])


@pytest.mark.parametrize('name', SCENES)
def test_all_scenes_render_finite(name):
    o, sc = oracle.from_scene_file(scene_path(name))
    p = pack.pack_params(sc, 1, 24, 16, 1, 5)
    img = o.render(p, 1, 1)
    assert np.isfinite(img).all()
    assert img[..., :3].max() > 0.0


def test_spectral_estimator_normalisation():
    """Camera inside a huge emitting sphere: every path ends on the emitter at bounce 0, so the pixel value is
    exposure * mean_k[ Emit(lambda_k) * CMF(lambda_k) ] * 330.  The hero wavelength is drawn on [360, 800] but wrapped
    into [390, 720) (shader.comp:1469, 971-974): lambda_k - 390 = (u + 82.5 k) mod 330 with u uniform on
    [-30, 410], so its density is 2/440 on an arc of 110 nm starting at (82.5 k - 30) mod 330 and 1/440 elsewhere,
    while the estimator multiplies by 330 and averages the four.  The expectation is therefore the integral of
    Emit * CMF * 330/440 * (1 + arcs(lambda)/4): a quirk that is part of the reference's observable behaviour."""
    import ctypes as C
    T, lum = 5500.0, 2.0
    scene = {'camera': pack.load_scene(scene_path('scene0'))['camera'],
             'sphere': [{'position': [0, 0, 0], 'radius': 500.0, 'materialID': 1, 'lightID': 1}],
             'material': [{'reflection': {'peakWavelength': 550.0, 'sigma': 30.0, 'isInvert': False}}],
             'light': [{'emission': {'temperature': T, 'luminosity': lum}}]}
    ubo = pack.pack_ubo(scene)
    p = pack.pack_params(scene, 1, 48, 32, 128, 5)
    img = oracle.Oracle(ubo).render(p, 128, 128)
    expo = float(p['apertureSize']) ** 2 * int(p['ISO'])
    got = img[..., :3].astype(np.float64).mean(axis=(0, 1)) / expo
    lam = np.arange(390.0, 720.0, 0.05)
    L = oracle.lib()
    emit = np.zeros(4, dtype=np.float32)
    cmf = np.zeros(3, dtype=np.float32)
    acc = np.zeros(3)
    for l in lam[::20]:                       # 1 nm steps are plenty for smooth integrands
        l4 = np.array([l] * 4, dtype=np.float32)
        L.oracle_emit(l4.ctypes.data, T, lum, emit.ctypes.data)
        L.oracle_wave_to_xyz(ubo.ctypes.data_as(C.c_void_p), float(l), cmf.ctypes.data_as(C.c_void_p))
        x = l - 390.0
        arcs = sum(1 for k in (1, 2, 3, 4) if ((x - (82.5 * k - 30.0)) % 330.0) <= 110.0)
        acc += float(emit[0]) * cmf.astype(np.float64) * (330.0 / 440.0) * (1.0 + arcs / 4.0)
    want = acc * 1.0                           # sum over 1 nm bins
    assert np.allclose(got, want, rtol=0.015), (got, want)   # ~2e5 samples: Monte-Carlo error well below 1 %


# ---- the samplers against their closed forms ("analytic / statistical") --------------------------------
def _samples(kind, n, param=0.0, normal=None, seed=12345):
    import ctypes as C
    L = oracle.lib()
    L.oracle_sample.argtypes = [C.c_int, C.c_uint, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p]
    out = np.zeros((n, 3), dtype=np.float32)
    nn = None if normal is None else np.ascontiguousarray(normal, dtype=np.float32)
    L.oracle_sample(kind, seed, n, param, None if nn is None else nn.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.astype(np.float64)


def test_uniform_disk_and_sphere_samplers():
    """SampleUniformUnitDisk (shader.comp:976-982): inside the unit disk, uniform in area (E[r^2] = 1/2, mean 0).
    SampleUniformUnitSphere (984-997): unit length, mean 0, E[z^2] = 1/3, z uniform on [-1, 1]."""
    n = 400000
    d = _samples(0, n)
    r2 = (d[:, :2] ** 2).sum(1)
    assert r2.max() <= 1.0 + 1e-6 and np.all(d[:, 2] == 0.0)
    assert abs(r2.mean() - 0.5) < 3e-3 and np.abs(d[:, :2].mean(0)).max() < 4e-3
    s = _samples(1, n)
    assert np.abs(np.linalg.norm(s, axis=1) - 1.0).max() < 1e-5
    assert np.abs(s.mean(0)).max() < 4e-3
    assert np.abs((s ** 2).mean(0) - 1.0 / 3.0).max() < 3e-3
    hist, _ = np.histogram(s[:, 2], bins=10, range=(-1.0, 1.0))
    assert np.abs(hist / n - 0.1).max() < 3e-3


def test_cosine_hemisphere_sampler_matches_its_pdf():
    """SampleCosineDirectionHemisphere (999-1003): normalize(n + uniform sphere) is cosine-weighted about n, the pdf
    CosineDirectionPDF = cos/pi (1005-1010): E[cos] = 2/3, E[cos^2] = 1/2, nothing below the horizon, symmetric about n."""
    n = 400000
    nrm = np.array([0.36, -0.48, 0.8])
    v = _samples(2, n, normal=nrm)
    c = v @ nrm
    assert np.abs(np.linalg.norm(v, axis=1) - 1.0).max() < 1e-5
    assert c.min() > -1e-6
    assert abs(c.mean() - 2.0 / 3.0) < 2e-3 and abs((c ** 2).mean() - 0.5) < 2e-3
    tangential = v - np.outer(c, nrm)
    assert np.abs(tangential.mean(0)).max() < 4e-3


@pytest.mark.parametrize('cos_max', [0.95, 0.6, 0.1])
def test_cosine_cone_sampler_matches_its_pdf(cos_max):
    """SampleCosineUnitCone (1012-1023) stays inside the cone of half-angle theta_max and is cosine-weighted there:
    CosineUnitConePDF = cos / (pi sin^2 theta_max) (1025-1028) integrates to 1 over the cone, and the sample mean of
    cos equals the pdf's E[cos] = (2/3)(1 - cos^3 theta_max) / sin^2 theta_max.  ToWorld (1113-1119) carries the cone
    onto an arbitrary axis without changing the angles."""
    import ctypes as C
    n = 400000
    v = _samples(3, n, param=cos_max)
    assert np.abs(np.linalg.norm(v, axis=1) - 1.0).max() < 1e-5
    assert v[:, 2].min() >= cos_max - 2e-4   # fp32 rounding of the half-angle construction near the rim
    sin2 = 1.0 - cos_max ** 2
    want = (2.0 / 3.0) * (1.0 - cos_max ** 3) / sin2
    assert abs(v[:, 2].mean() - want) < 2e-3
    assert np.abs(v[:, :2].mean(0)).max() < 4e-3
    # the pdf integrates to one: midpoint rule over cos in [cos_max, 1] of pdf(cos) * 2 pi
    L = oracle.lib()
    L.oracle_cone_pdf.restype = C.c_float
    L.oracle_cone_pdf.argtypes = [C.c_float, C.c_float]
    m = 2000
    cs = cos_max + (np.arange(m) + 0.5) * (1.0 - cos_max) / m
    integral = sum(L.oracle_cone_pdf(float(c), float(cos_max)) for c in cs) * 2.0 * np.pi * (1.0 - cos_max) / m
    assert abs(integral - 1.0) < 1e-3
    # histogram of cos against the pdf
    hist, edges = np.histogram(v[:, 2], bins=8, range=(cos_max, 1.0))
    lo, hi = edges[:-1], edges[1:]
    expect = (hi ** 2 - lo ** 2) / sin2
    assert np.abs(hist / n - expect).max() < 4e-3
    axis = np.array([-0.6, 0.0, 0.8])
    w = _samples(4, n, param=cos_max, normal=axis)
    assert np.allclose(w @ axis, v[:, 2], atol=2e-6)


def test_orthonormal_basis_and_mis_weight():
    """OrthonormalBasis (1093-1103, Duff et al. without the sign trick, special-cased at n.z < -0.9999999): b1, b2, n
    orthonormal for random n and at the pole.  MISPowerHeuristicsBeta2 (1293-1296): w(a, b) + w(b, a) = 1."""
    import ctypes as C
    L = oracle.lib()
    L.oracle_orthonormal_basis.argtypes = [C.c_void_p, C.c_void_p]
    L.oracle_mis_weight.restype = C.c_float
    L.oracle_mis_weight.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    ns = rng.normal(size=(2000, 3))
    ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    ns = np.vstack([ns, [[0.0, 0.0, 1.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]]]).astype(np.float32)
    for nvec in ns:
        b = np.zeros(6, dtype=np.float32)
        L.oracle_orthonormal_basis(np.ascontiguousarray(nvec).ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
        b1, b2, n64 = b[:3].astype(np.float64), b[3:].astype(np.float64), nvec.astype(np.float64)
        tol = 2e-3 if nvec[2] < -0.99 else 2e-5   # 1 / (1 + n.z) loses digits towards the south pole
        assert abs(np.dot(b1, b1) - 1.0) < tol and abs(np.dot(b2, b2) - 1.0) < tol
        assert abs(np.dot(b1, b2)) < tol and abs(np.dot(b1, n64)) < tol and abs(np.dot(b2, n64)) < tol
    for a, b in ((0.3, 0.7), (1e-3, 5.0), (2.0, 2.0)):
        assert abs(L.oracle_mis_weight(a, b) + L.oracle_mis_weight(b, a) - 1.0) < 1e-6
    assert L.oracle_mis_weight(1.0, 0.0) == 1.0


@pytest.mark.parametrize('a,h,R', [(0.0, 3.0, 1.0), (2.0, 3.0, 1.0), (4.0, 2.0, 1.0), (1.0, 5.0, 0.5)])
def test_direct_lighting_matches_the_closed_form(a, h, R):
    """A Lambertian plane under a spherical black-body emitter, nothing else: the radiance leaving the point x0 = (a, 0, 0)
    is direct light only, and for a sphere of radius R wholly above the horizon at distance d it is exactly
        L_o(lambda) = rho(lambda) * L_e(lambda) * (R / d)^2 * cos(theta_c)
    (irradiance of a uniform sphere: pi L (R/d)^2 cos theta_c; BRDF rho / pi).  TracePath with pathLength 2 combines the
    light sample (cone towards the emitter's bounding sphere, shader.comp:1298-1343) and the BSDF sample that hits the
    emitter (1359-1364) by the power heuristic, with Russian roulette on both.  The reference weights each pair of
    samples with pdfs of two different directions and does not compensate the roulette on the light sample, so it is
    not exactly unbiased; measured here: within 0.6 % of the closed form (2 M paths per case).  The gate is 2 %: a
    missing pi, cosine, 1/pdf or a wrong cone would show as tens of percent."""
    import ctypes as C
    L = oracle.lib()
    L.oracle_trace_path.argtypes = [C.c_void_p] * 5 + [C.c_uint, C.c_size_t, C.c_void_p]
    L.oracle_emit.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.oracle_spd.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
    T, lum = 5500.0, 3.0
    scene = {'camera': pack.load_scene(scene_path('scene0'))['camera'],
             'sphere': [{'position': [0, h, 0], 'radius': R, 'materialID': 1, 'lightID': 1}],
             'plane': [{'position': [0, 0, 0], 'materialID': 1, 'lightID': 0}],
             'material': [{'reflection': {'peakWavelength': 550.0, 'sigma': 6.0, 'isInvert': False}}],
             'light': [{'emission': {'temperature': T, 'luminosity': lum}}]}
    ubo = pack.pack_ubo(scene)
    p = np.ascontiguousarray(pack.pack_params(scene, 1, 64, 64, 1, 2))
    o = np.array([a, 1.0, 0.0], dtype=np.float32)
    d = np.array([0.0, -1.0, 0.0], dtype=np.float32)
    l4 = np.array([550.0, 500.0, 600.0, 450.0], dtype=np.float32)
    mean = np.zeros(4, dtype=np.float64)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.oracle_trace_path(vp(ubo), vp(p), vp(o), vp(d), vp(l4), 777, 500000, vp(mean))
    emit, rho = np.zeros(4, dtype=np.float32), np.zeros(4, dtype=np.float32)
    L.oracle_emit(vp(l4), T, lum, vp(emit))
    L.oracle_spd(vp(l4), 550.0, 6.0, 0, vp(rho))
    dist = float(np.hypot(a, h))
    want = rho.astype(np.float64) * emit.astype(np.float64) * (R / dist) ** 2 * (h / dist)
    assert np.all(want > 0) and np.all(np.abs(mean / want - 1.0) < 0.02), (mean, want)


def test_sphere_tracing_of_a_unit_sphere_sdf_matches_the_analytic_hit():
    """SphereTracing (shader.comp:779-860) over `length(p) - 1` in a bounding box: every ray that meets the analytic
    sphere is reported as a hit and no other; the hit distance is the marched t minus the 1e-3 back-off (853), the march
    stopping where |sdf| < 1e-4 (a little earlier along grazing rays): analytic t - 6e-3 <= hit <= analytic t - 0.9e-3;
    the central-difference normal (721-730, eps 1e-4 in fp32) is the radial direction within 5e-3."""
    src = 'float sdf(in vec3 p) { return length(p) - 1.0; }\nfloat sdfmaterial(in vec3 p) { return 0.0; }\n'
    ubo = np.zeros(pack.UBO_FLOATS, dtype=np.float32)
    ubo[5] = 1
    c = np.array([0.5, 0.25, 4.0])
    ubo[pack.OFF_SDF:pack.OFF_SDF + 6] = [c[0], c[1], c[2], 2.4, 2.4, 2.4]
    o = oracle.Oracle(ubo, [src])
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(600):
        org = rng.normal(size=3) * 0.3
        d = c + rng.normal(size=3) * 0.6 - org
        d /= np.linalg.norm(d)
        t, out = o.intersect(org.astype(np.float32), d.astype(np.float32))
        oc = org - c
        b, cc = float(np.dot(oc, d)), float(np.dot(oc, oc)) - 1.0
        disc = b * b - cc
        if disc <= 1e-4:            # a miss or a graze within the march's tolerance: either verdict is acceptable
            continue
        ta = -b - np.sqrt(disc)
        assert t < 1e5, 'missed a sphere the ray crosses'
        assert -6e-3 <= t - ta <= -0.9e-3, (t, ta)
        n = org + ta * d - c
        assert np.abs(out[:3] - n).max() < 5e-3 and out[4] == -1.0
        hits += 1
    assert hits > 300
    t, _ = o.intersect(np.array([0, 0, 0], np.float32), np.array([0, 1, 0], np.float32))
    assert t == np.float32(1e5)   # past the box: no hit


@pytest.mark.parametrize('a,b,c,d,scale', [(2.0, 2.0, 0.0, 0.6, 0.5),            # a torus: a = b = R, c = 0, d = r
                                           (3.36, -3.17, -1.06, -1.5, 0.25)])    # scene0's cyclide
def test_dupin_cyclide_hits_match_a_float64_root_finder(a, b, c, d, scale):
    """DupinCyclide (shader.comp:633-679) + SolveQuartic / SolveCubic (474-541): for rays shot at the surface from
    outside its bounding sphere, the first crossing of the implicit (x^2+y^2+z^2+b^2-d^2)^2 = 4((ax-cd)^2+(by)^2) -- in
    the frame after translation, scale and the shader's .xzy swap -- found by sign changes and bisection in float64,
    agrees with the fp32 closed-form solver within 2e-3 (median error 1e-5), with no false hit and no miss."""
    pos = np.array([0.0, 1.0, -3.0])
    scene = {'camera': {}, 'cyclide': [{'position': list(pos), 'rotation': [0, 0, 0], 'scale': [scale] * 3, 'a': a, 'b': b, 'c': c, 'd': d,
                                        'boundingRadius': abs(a) + abs(d) + abs(c) + 1.0, 'materialID': 1, 'lightID': 0}],
             'material': [{'reflection': {'peakWavelength': 550, 'sigma': 10, 'isInvert': False}}], 'light': []}
    o = oracle.Oracle(pack.pack_ubo(scene))
    rng = np.random.default_rng(1)
    extent = scale * (abs(a) + abs(d) + abs(c))

    def implicit(p):  # p: (..., 3) world points
        q = (p - pos) / scale
        x, y, z = q[..., 0], q[..., 2], q[..., 1]
        return (x * x + y * y + z * z + b * b - d * d) ** 2 - 4.0 * ((a * x - c * d) ** 2 + (b * y) ** 2)

    hits = errs = 0
    worst = []
    for _ in range(160):
        org = rng.normal(size=3)
        org = pos + org / np.linalg.norm(org) * extent * 2.0
        dr = pos + rng.normal(size=3) * scale * abs(a) * 0.7 - org
        dr /= np.linalg.norm(dr)
        t, _ = o.intersect(org.astype(np.float32), dr.astype(np.float32))
        ts = np.linspace(0.0, extent * 4.5, 6001)
        v = implicit(org + ts[:, None] * dr)
        idx = np.where(np.sign(v[:-1]) * np.sign(v[1:]) < 0)[0]
        if len(idx) == 0:
            if np.abs(v).min() > 1e-3 * np.abs(v).max():   # a clear miss (not a graze between two samples)
                assert t >= 1e5, 'false hit'
            continue
        lo, hi = ts[idx[0]], ts[idx[0] + 1]
        flo = implicit(org + lo * dr)
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            fm = implicit(org + mid * dr)
            if np.sign(fm) == np.sign(flo):
                lo, flo = mid, fm
            else:
                hi = mid
        assert t < 1e5, 'missed the surface'
        worst.append(abs(t - 0.5 * (lo + hi)))
        hits += 1
    assert hits > 50 and max(worst) < 2e-3 and float(np.median(worst)) < 1e-4, (hits, max(worst))


def test_lens_caps_match_the_closed_form():
    """LensIntersection / SphereSliceIntersection (shader.comp:366-448): a converging lens is two caps of spheres of radius
    2f whose rims (radius R) sit thickness/2 either side of the centre, optical axis = local x.  For a ray parallel to
    the axis at height y the hit is on the sphere centred 2f behind the cap's vertex: x = x_c - sqrt(4f^2 - y^2); normal
    radial, flipped towards the ray for hits from inside; beyond the rim and edge-on there is no hit."""
    f, R, th = 1.0, 1.2, 0.2
    scene = {'camera': {}, 'lens': [{'position': [5.0, 0.0, 0.0], 'rotation': [0, 0, 0], 'radius': R, 'focalLength': f, 'thickness': th,
                                     'isConverging': True, 'materialID': 1, 'lightID': 0}],
             'material': [{'reflection': {'peakWavelength': 550, 'sigma': 10, 'isInvert': False}}], 'light': []}
    o = oracle.Oracle(pack.pack_ubo(scene))
    h = 2 * f - np.sqrt(4 * f * f - R * R)            # cap height
    xc = 5.0 - (h + th / 2) + 2 * f                   # centre of the front cap's sphere
    for y in (0.0, 0.5, 1.0, 1.19):
        t, out = o.intersect(np.array([0, y, 0], np.float32), np.array([1, 0, 0], np.float32))
        x = xc - np.sqrt(4 * f * f - y * y)
        assert abs(t - x) < 2e-6 and np.allclose(out[:3], [(x - xc) / (2 * f), y / (2 * f), 0.0], atol=1e-6)
    assert o.intersect(np.array([0, 1.21, 0], np.float32), np.array([1, 0, 0], np.float32))[0] == np.float32(1e5)
    assert o.intersect(np.array([5, -5, 0], np.float32), np.array([0, 1, 0], np.float32))[0] == np.float32(1e5)
    xb = 5.0 + (h + th / 2) - 2 * f                    # centre of the back cap's sphere
    t, out = o.intersect(np.array([10, 0.5, 0], np.float32), np.array([-1, 0, 0], np.float32))
    assert abs(t - (10 - (xb + np.sqrt(4 - 0.25)))) < 2e-6 and out[0] > 0
    t, out = o.intersect(np.array([5, 0.5, 0], np.float32), np.array([1, 0, 0], np.float32))   # from between the caps
    assert abs(t - (xb + np.sqrt(4 - 0.25) - 5)) < 2e-6 and out[0] < 0 and out[1] < 0


def test_camera_lens_tracing_matches_float64_physics():
    """TracePathLens (shader.comp:1409-1444) for scene1's camera: the BK7 lens sits lensDistance along the camera's forward
    axis with its optical axis ON that axis (rotation (0, 90 - yaw, pitch)); a sensor ray meets the front cap (sphere of
    radius 2f), refracts by Snell with n(lambda) from the Sellmeier formula, meets the back cap from inside and refracts
    out -- with the reference's quirk that the exit index is evaluated at lambda / n (1440).  An independent float64
    computation from that description agrees with the oracle to 1e-5 on 200 random rays and wavelengths."""
    import ctypes as C
    L = oracle.lib()
    L.oracle_lens_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    scene = pack.load_scene(scene_path('scene1'))
    ubo = pack.pack_ubo(scene)
    p = np.ascontiguousarray(pack.pack_params(scene, 1, 64, 64, 1, 5))
    g = lambda k: float(np.ravel(p[k])[0])  # noqa: E731
    vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731

    def bk7(l_nm):
        l2 = (l_nm * 1e-3) ** 2
        return np.sqrt(1 + 1.03961212 * l2 / (l2 - 6.00069867e-3) + 0.231792344 * l2 / (l2 - 2.00179144e-2) + 1.01046945 * l2 / (l2 - 1.03560653e2))

    def refract(i, n, eta):
        ndi = np.dot(n, i)
        return eta * i - (eta * ndi + np.sqrt(1 - eta * eta * (1 - ndi * ndi))) * n

    cam = np.array([g('cameraPosX'), g('cameraPosY'), g('cameraPosZ')])
    f, R, th, D = g('lensFocalLength'), g('lensRadius'), g('lensThickness'), g('lensDistance')
    out = np.zeros(9, dtype=np.float32)
    L.oracle_lens_ray(vp(ubo), vp(p), vp(cam.astype(np.float32)), vp(np.array([0, 0, 1], np.float32)), 550.0, vp(out))
    fwd = out[6:9].astype(np.float64)
    assert abs(np.linalg.norm(fwd) - 1.0) < 1e-6
    u = np.cross(fwd, [0, 1, 0])
    u /= np.linalg.norm(u)
    v = np.cross(fwd, u)
    h = 2 * f - np.sqrt(4 * f * f - R * R)
    lens = cam + fwd * D
    c1, c2 = lens + fwd * (-(h + th / 2) + 2 * f), lens + fwd * ((h + th / 2) - 2 * f)
    rng = np.random.default_rng(2)
    for _ in range(200):
        so = (cam + (u * rng.uniform(-1, 1) + v * rng.uniform(-1, 1)) * 0.004).astype(np.float32)
        ap = cam + fwd * g('apertureDist') + (u * rng.uniform(-1, 1) + v * rng.uniform(-1, 1)) * 0.001
        d = ap - so
        d = (d / np.linalg.norm(d)).astype(np.float32)
        lam = float(rng.uniform(380, 780))
        L.oracle_lens_ray(vp(ubo), vp(p), vp(so), vp(d), lam, vp(out))
        o64, d64 = so.astype(np.float64), d.astype(np.float64)
        oc = o64 - c1
        b, cc = np.dot(oc, d64), np.dot(oc, oc) - 4 * f * f
        p1 = o64 + (-b - np.sqrt(b * b - cc)) * d64                     # front cap, from outside
        n = bk7(lam)
        d1 = refract(d64, (p1 - c1) / (2 * f), 1.0 / n)
        oc = p1 - c2
        b, cc = np.dot(oc, d1), np.dot(oc, oc) - 4 * f * f
        p2 = p1 + (-b + np.sqrt(b * b - cc)) * d1                       # back cap, from inside
        d2 = refract(d1, -(p2 - c2) / (2 * f), bk7(lam / n))
        assert np.abs(out[:3] - p2).max() < 1e-5 and np.abs(out[3:6] - d2).max() < 1e-5


def test_rotated_box_and_its_bounding_sphere_cull_match_a_float64_slab_test():
    """BoxIntersection (shader.comp:337-364) behind the BoundingSphere cull (263-276, 887-897): for random rays and
    random rotations the hit distance equals a float64 slab test in the box's frame (local = v * (mX mY mZ), the
    matrix pinned above) and the normal is the face normal turned back to world space; the cull never rejects a ray
    the slab test accepts, and rays that miss the box miss."""
    rng = np.random.default_rng(9)
    checked = 0
    for _ in range(12):
        deg = rng.uniform(-180, 180, 3)
        size = rng.uniform(0.5, 2.5, 3)
        pos = rng.uniform(-1, 1, 3) + np.array([0, 0, 6.0])
        scene = {'camera': {}, 'box': [{'position': list(pos), 'rotation': list(deg), 'size': list(size), 'materialID': 1, 'lightID': 0}],
                 'material': [{'reflection': {'peakWavelength': 550, 'sigma': 10, 'isInvert': False}}], 'light': []}
        o = oracle.Oracle(pack.pack_ubo(scene))
        a = np.deg2rad(np.float32(deg).astype(np.float64))
        sx, sy, sz, cx, cy, cz = *np.sin(a), *np.cos(a)
        M = (np.array([[1, 0, 0], [0, cx, sx], [0, -sx, cx]]) @ np.array([[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]]) @
             np.array([[cz, sz, 0], [-sz, cz, 0], [0, 0, 1]]))
        pos32, size32 = np.float32(pos).astype(np.float64), np.float32(size).astype(np.float64)
        for _ in range(60):
            org = rng.normal(size=3) * 0.5
            d = pos + rng.normal(size=3) * 1.2 - org
            d = np.float32(d / np.linalg.norm(d))
            org = np.float32(org)
            t, out = o.intersect(org, d)
            lo, ld = M.T @ (org.astype(np.float64) - pos32), M.T @ d.astype(np.float64)   # v * M == M^T v
            with np.errstate(divide='ignore'):
                ta, tb = (-0.5 * size32 - lo) / ld, (0.5 * size32 - lo) / ld
            t1, t2 = np.minimum(ta, tb).max(), np.maximum(ta, tb).min()
            if t1 > t2 + 1e-4 or t2 < 1e-3:
                assert t == np.float32(1e5)
                continue
            if t1 > t2 - 1e-4:
                continue   # a graze: either verdict
            want = t2 if t1 < 0 else t1
            assert abs(t - want) < 2e-5 * max(1.0, want), (t, want)
            q = np.abs((lo + ld * want) / size32)
            face = int(np.argmax(q))
            nl = np.zeros(3)
            nl[face] = -np.sign(ld[face])
            assert np.allclose(out[:3], M @ nl, atol=2e-5)
            checked += 1
    assert checked > 150


def test_accumulate_branches_against_their_formulas():
    """Accumulate (shader.comp:1492-1507).  Static branch: out = ((n - 1) in + out) / n with n = currentSamples / spf
    (integer division).  Interactive branch (currentSamples == spf and frame > spf, i.e. after a reset while frames keep
    counting): out = (1 - w) out + w in with w = 2^(-8 / (FPS persistence)).  Both checked against float64 on the
    oracle's own per-sample values."""
    o, sc = oracle.from_scene_file(scene_path('scene1'))
    W, H, spf = 24, 16, 2
    p = pack.pack_params(sc, 1, W, H, spf, 5)
    expo = float(np.ravel(p['apertureSize'])[0]) ** 2 * int(np.ravel(p['ISO'])[0])

    def frame_value(first):  # Rendering(): mean of spf Scene() calls times the exposure, for every pixel
        out = np.zeros((H, W, 3))
        for y in range(H):
            for x in range(W):
                out[y, x] = o.samples(p, x, y, first, spf).astype(np.float64).mean(0) * expo
        return out

    img = np.zeros((H, W, 4), dtype=np.float32)
    o.dispatch(p, img)                                   # frame = spf, currentSamples = spf: n = 1
    f0 = frame_value(0)
    assert np.allclose(img[..., :3], f0, rtol=2e-6, atol=1e-9)
    q = p.copy()
    q['frame'] = 2 * spf
    q['currentSamples'] = 2 * spf                        # n = 2: running mean of the two frames
    o.dispatch(q, img)
    f1 = frame_value(spf)
    assert np.allclose(img[..., :3], 0.5 * (f0 + f1), rtol=3e-6, atol=1e-9)
    r = p.copy()
    r['frame'] = 5 * spf
    r['currentSamples'] = spf                            # reset while frames keep counting: the EMA branch
    r['FPS'] = 30.0
    before = img[..., :3].astype(np.float64)
    o.dispatch(r, img)
    f4 = frame_value(4 * spf)
    w = 2.0 ** (-8.0 / (30.0 * float(np.ravel(r['persistence'])[0])))
    assert 0.0 < w < 1.0
    assert np.allclose(img[..., :3], (1.0 - w) * f4 + w * before, rtol=5e-6, atol=1e-9)
    assert (img[..., 3] == 1.0).all()


def test_search_sdf_merges_overlapping_boxes_and_searches_again_after_leaving_them():
    """SearchSDF (shader.comp:732-777) + the re-search of SphereTracing (826-840).  Three unit-sphere SDFs: boxes A and B
    overlap along the ray (their intervals merge, both bits set), box C lies beyond a gap.  A ray that crosses A's box
    beside its sphere must still find B's sphere inside the merged interval; a ray that leaves A's box without a hit must
    search again and find C; hits are the analytic ones minus the march's 1e-3 back-off."""
    src = 'float sdf(in vec3 p) { return length(p) - 1.0; }\nfloat sdfmaterial(in vec3 p) { return 0.0; }\n'
    ubo = np.zeros(pack.UBO_FLOATS, dtype=np.float32)
    ubo[5] = 3
    for i, c in enumerate([(0, 0, 4), (0.6, 0, 6.2), (-1.15, 0, 10)]):
        ubo[pack.OFF_SDF + 6 * i:pack.OFF_SDF + 6 * i + 6] = [c[0], c[1], c[2], 2.4, 2.4, 2.4]
    o = oracle.Oracle(ubo, [src] * 3)
    cases = [(1.1, 6.2 - np.sqrt(1 - 0.5 ** 2)),       # through A's box beside sphere A, into sphere B (merged interval)
             (1.19, 6.2 - np.sqrt(1 - 0.59 ** 2)),
             (-1.15, 9.0),                              # through A's box, out, across the gap, into sphere C (re-search)
             (-1.19, 10 - np.sqrt(1 - 0.04 ** 2)),
             (0.0, 3.0)]                                # sphere A itself
    for x, want in cases:
        t, out = o.intersect(np.array([x, 0, 0], np.float32), np.array([0, 0, 1], np.float32))
        assert -2.5e-3 <= t - want <= -0.9e-3, (x, t, want)
        assert out[2] < 0 and out[4] == -1.0
    assert o.intersect(np.array([3.0, 0, 0], np.float32), np.array([0, 0, 1], np.float32))[0] == np.float32(1e5)
