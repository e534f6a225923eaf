"""Pins the oracle and the host-side packers to the REFERENCE ITSELF: oracle/_ref is the reference's own
src/pathtracer.cpp (data contract, verbatim functions) and src/shader.comp (mechanically rewritten over the vendored glm)
compiled for the CPU by oracle/ref_build.py.  Integer and byte work must be bit-exact; float work is compared within the
tolerances written here, because GLSL arithmetic is only defined to a tolerance: oracle.cpp evaluates transcendentals
with include/pt_math.h and normalize as v/length, _ref uses glibc libm and glm's v*inversesqrt.

Skipped only when neither /root/reference nor prebuilt oracle/_ref objects exist.  No GPU needed.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, SCENES, scene_path
from oracle import oracle, pack, ref, ref_build

pytestmark = pytest.mark.skipif(not ref.available(), reason='no /root/reference and no prebuilt oracle/_ref')

SDF_SCENES = ['scene3', 'scene4', 'scene5', 'scene6', 'scene7', 'scene8', 'scene9', 'scene10']
GRADED = ['scene0', 'scene1', 'scene9', 'scene10', 'scene8']
SYNTHETIC = sorted(n[:-5] for n in os.listdir(os.path.join(ROOT, 'scenes_synthetic')) if n.endswith('.json'))


def params_of(push):
    return np.frombuffer(bytes(push), dtype=pack.PARAMS_DTYPE)[0]


@pytest.fixture(scope='module')
def shader0():
    return ref.RefShader(scene_path('scene0'))


# ---------------------------------------------------------------------------------------------------------------
# the recipe really took the functions SURVEY.md / DESIGN.md cite
# ---------------------------------------------------------------------------------------------------------------

@pytest.mark.skipif(not ref_build.have_reference(), reason='needs the reference tree')
def test_extracted_ranges_are_the_cited_ones():
    ref_build.build_host()
    where = json.load(open(os.path.join(ref_build.OUT, 'ref_host.lines.json')))
    assert where['App::InsertSDF'] == [2004, 2054]
    assert where['App::UpdateFromJSON'] == [2576, 2722]
    assert where['App::UpdateToJSON'] == [2724, 2858]
    assert where['App::UpdateUniformBuffer'] == [3642, 3811]
    assert where['App::UpdatePushConstant'] == [3813, 3834]
    assert where['struct UniformBufferObject'] == [187, 195]
    assert where['struct PushConstantValues'] == [197, 218]
    assert where['App::SaveRender pixel loop'][0] == 3501
    assert where['CIEXYZ1931'] == [400, 842]


@pytest.mark.skipif(not ref_build.have_reference(), reason='needs the reference tree')
def test_insert_sdf_needs_crlf_and_emits_material_line_first():
    """the reference's offsets only work on a CRLF file; the dispatcher lines it generates are
    `material` before `distance` for every SDF, SDF 1 first."""
    text = ref_build.inserted_shader(open(scene_path('scene10')).read())
    i = text.index('float SDFMATERIAL(')
    body = text[i:text.index('return sdfmaterial;', i)]
    lines = [l.strip() for l in body.split('\n') if l.strip().startswith('if ((set')]
    assert len(lines) == 2 and 'minMaterial(sdf, SDF1(p - vec3(sdfs[0], sdfs[1], sdfs[2]))' in lines[0]
    assert lines[1] == 'if ((set1 & 1) == 1) sdf = min(sdf, SDF1(p - vec3(sdfs[0], sdfs[1], sdfs[2])));'
    assert 'float SDF1(in vec3 p)' in text and 'float SDF1MATERIAL(in vec3 p)' in text


# ---------------------------------------------------------------------------------------------------------------
# data contract: loader + packers, three ways, bit for bit
# ---------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('name', SCENES)
def test_packers_equal_the_reference(ptlib, name):
    """UpdateFromJSON + UpdateUniformBuffer + UpdatePushConstant of the reference == pt_scene_* == oracle/pack.py."""
    path = scene_path(name)
    sc = ptlib.Scene.load(path)
    js = pack.load_scene(path)
    for shot in range(1, sc.num_shots + 1):
        rs = ref.RefScene(path, shot)
        assert rs.num_shots() == sc.num_shots
        u_ref = rs.ubo()
        assert np.array_equal(u_ref.view(np.uint32), sc.pack_ubo().view(np.uint32))
        assert np.array_equal(u_ref.view(np.uint32), pack.pack_ubo(js).view(np.uint32))
        assert rs.sdf_sources() == [s.decode() for s in sc.sdf_sources] == pack.sdf_sources(js)
        # the offscreen loop's third dispatch of 3 samples (host:4042-4048), frame time 1/60 s
        push = params_of(rs.push(640, 360, 9, 9, 3, path_length=7, frame_time=float(np.float32(1) / np.float32(60))))
        mine = np.array(sc.pack_params(shot, 640, 360, 3, 7), copy=True)
        mine['frame'] = 9
        mine['currentSamples'] = 9
        theirs = np.array(pack.pack_params(js, shot, 640, 360, 3, 7, dispatch=3), copy=True)
        for f in pack.PARAMS_DTYPE.names:
            if f == 'FPS':
                assert abs(float(push[f]) - 60.0) < 1e-4
                continue
            assert np.array_equal(push[f], np.ravel(mine)[0][f]), f
            assert np.array_equal(push[f], theirs[f]), f
        rs.close()


@pytest.mark.parametrize('name', SYNTHETIC)
def test_packers_equal_the_reference_at_capacity(ptlib, name):
    path = os.path.join(ROOT, 'scenes_synthetic', name + '.json')
    rs = ref.RefScene(path)
    theirs, mine = rs.ubo().view(np.uint32).copy(), ptlib.Scene.load(path).pack_ubo().view(np.uint32).copy()
    js = pack.load_scene(path)
    if max(len(js.get('lens', [])), len(js.get('cyclide', []))) > len(js.get('plane', [])):
        # host:3704,3728 read planes[i].lightID with i past the end of `planes`: undefined behaviour in the reference
        # (heap garbage decides which lenses / cyclides get sampled).  The product and oracle/pack.py define it as
        # "not registered" (DESIGN.md section 3); everything else must still be identical.
        for u in (theirs, mine):
            u[6] = 0
            u[pack.OFF_LID:pack.OFF_LID + pack.MAX_LIGHTIDS] = 0
    assert np.array_equal(theirs, mine)
    rs.close()


def test_light_registration_quirk_is_the_reference_behaviour(ptlib, tmp_path):
    """host:3704,3728 read planes[i].lightID for lenses and cyclides."""
    s = pack.load_scene(scene_path('scene0'))
    s['lens'][0]['lightID'] = 1
    s['plane'][0]['lightID'] = 1
    p = tmp_path / 'quirk.json'
    p.write_text(json.dumps(s))
    rs = ref.RefScene(str(p))
    u = rs.ubo()
    assert u[6] == 4 and list(u[pack.OFF_LID:pack.OFF_LID + 4]) == [2, 3, 5, 6]
    assert np.array_equal(u.view(np.uint32), ptlib.Scene.parse(json.dumps(s)).pack_ubo().view(np.uint32))


def test_cie_table_is_the_reference_table(ptlib):
    t = ref.cie_table()
    assert np.array_equal(t.view(np.uint32), pack.cie_table().view(np.uint32))
    assert np.array_equal(t.view(np.uint32), ptlib.api.cie1931_table().view(np.uint32))


@pytest.mark.parametrize('name', SCENES)
def test_scene_save_equals_update_to_json(ptlib, name):
    """pt_scene_to_json vs the reference's UpdateToJSON + RoundDecimal(1e5) (host:2724-2858, 935-943): same values.
    (Key order and number formatting are nlohmann's; the parsed documents are compared.)"""
    rs = ref.RefScene(scene_path(name))
    theirs = json.loads(rs.to_json())
    mine = json.loads(ptlib.Scene.load(scene_path(name)).to_json())

    def close(a, b, where):
        if isinstance(a, dict):
            assert isinstance(b, dict) and set(a) == set(b), where
            for k in a:
                close(a[k], b[k], where + '/' + k)
        elif isinstance(a, list):
            assert isinstance(b, list) and len(a) == len(b), where
            for i, (x, y) in enumerate(zip(a, b)):
                close(x, y, '%s[%d]' % (where, i))
        elif isinstance(a, str) or isinstance(a, bool):
            assert a == b, where
        else:
            assert abs(float(a) - float(b)) <= 1e-9 * max(1.0, abs(float(a))), (where, a, b)
    close(theirs, mine, name)


def test_display_transform_bytes_equal_save_render(ptlib, tmp_path):
    """pt_write_ppm vs the reference's SaveRender loop + SavePPM (host:3501-3512, 918-933): every byte, all four
    tonemaps, on values that cover black, negatives after the matrix, > 1, huge, NaN."""
    import ctypes as C
    rng = np.random.default_rng(11)
    h, w = 64, 96
    img = np.ones((h, w, 4), dtype=np.float32)
    img[..., :3] = (rng.random((h, w, 3)) ** 3 * 3).astype(np.float32)
    img[0, :8, :3] = 0.0
    img[1, :8, :3] = [[0.0, 0.0, 1.0]] * 8      # negative red after XYZ -> RGB
    img[2, :8, :3] = 1e6
    img[3, :4, 0] = np.nan
    img[4, :8, :3] = rng.random((8, 3)).astype(np.float32) * 1e-4
    L = ptlib.lib()
    for tm in (0, 1, 2, 3):
        theirs = ref.save_render_pixels(img, tm)
        path = str(tmp_path / ('t%d.ppm' % tm)).encode()
        assert L.pt_write_ppm(path, img.ctypes.data_as(C.c_void_p), w, h, tm) == 0
        raw = open(path, 'rb').read()
        rpath = str(tmp_path / ('r%d.ppm' % tm)).encode()
        assert ref.host().ref_save_ppm(rpath, w, h, theirs.ctypes.data_as(C.c_void_p)) == 0
        assert raw == open(rpath, 'rb').read(), 'tonemap %d' % tm


# ---------------------------------------------------------------------------------------------------------------
# integer work: bit-exact
# ---------------------------------------------------------------------------------------------------------------

def test_pcg32_and_seeds_bit_exact(shader0):
    rng = np.random.default_rng(5)
    seeds = np.concatenate([np.arange(0, 4096, dtype=np.uint32), np.array([0xFFFFFFFF, 0x12345678], dtype=np.uint32),
                            rng.integers(0, 2**32, 200000, dtype=np.uint64).astype(np.uint32)])
    theirs = shader0.pcg32_n(seeds)
    L = oracle.lib()
    mine = np.array([L.oracle_pcg32(int(s)) for s in seeds[:6000]], dtype=np.uint32)
    assert np.array_equal(theirs[:6000], mine)
    # the whole batch against a numpy restatement of shader.comp:937-941
    st = seeds.astype(np.uint64) * 747796405 + 2891336453 & 0xFFFFFFFF
    word = (((st >> ((st >> 28) + 4)) ^ st) * 277803737) & 0xFFFFFFFF
    assert np.array_equal(theirs, ((word >> 22) ^ word).astype(np.uint32))
    # SURVEY App. E known answers
    assert shader0.pcg32(0) == 0x07bb2fe2 and shader0.pcg32(0xFFFFFFFF) == 0xe62a4902
    # Scene()'s seed for random (pixel, sample) triples at three frame sizes and two dispatch shapes
    rs = ref.RefScene(scene_path('scene0'))
    for (w, h, frame, spf) in ((512, 512, 1, 1), (1920, 1080, 64, 8), (3840, 2160, 16384, 64)):
        push = rs.push(w, h, frame, frame, spf)
        g = np.stack([rng.integers(0, w, 3000), rng.integers(0, h, 3000), rng.integers(0, spf, 3000)], 1).astype(np.int32)
        theirs = shader0.generate_seed_n(push, g)
        p = params_of(push)
        mine = np.array([L.oracle_generate_seed(p.tobytes(), int(a), int(b), int(c)) for a, b, c in g[:1500]],
                        dtype=np.uint32)
        assert np.array_equal(theirs[:1500], mine)
    # float(seed) / 0xFFFFFFFFu
    import ctypes as C
    for s in (0, 1, 0x3a20ffe8, 0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF):
        f, nxt = shader0.random_float(s)
        c = C.c_uint32(s)
        assert oracle.lib().oracle_random_float(C.byref(c)) == f and c.value == nxt


# ---------------------------------------------------------------------------------------------------------------
# leaf functions
# ---------------------------------------------------------------------------------------------------------------

def test_exact_leaves_bit_equal(shader0):
    """Functions built from + - * / sqrt floor only: same bits in glm and in the oracle."""
    import ctypes as C
    L = oracle.lib()
    ubo = ref.RefScene(scene_path('scene0')).ubo()
    rng = np.random.default_rng(2)
    for wave in np.concatenate([[360.0, 555.0, 550.5, 799.999, 800.0, 359.9, 801.0],
                                rng.uniform(360, 800, 3000)]).astype(np.float32):
        mine = np.zeros(3, np.float32)
        L.oracle_wave_to_xyz(ubo.ctypes.data_as(C.c_void_p), float(wave), mine.ctypes.data_as(C.c_void_p))
        if wave < 800.0:    # at exactly 800 the shader reads past the table (App. C-16; unreachable: lambda < 720)
            assert np.array_equal(shader0.wave_to_xyz(ubo, float(wave)).view(np.uint32), mine.view(np.uint32)), wave
    for lh in rng.uniform(360, 800, 3000).astype(np.float32):
        mine = np.zeros(4, np.float32)
        L.oracle_sample_wavelengths(float(lh), mine.ctypes.data_as(C.c_void_p))
        assert np.array_equal(shader0.sample_wavelengths(float(lh)).view(np.uint32), mine.view(np.uint32))
    for l in rng.uniform(300, 900, 3000).astype(np.float32):
        assert shader0.bk7(float(l)) == L.oracle_bk7(float(l))
    for a, b in rng.uniform(0, 3, (2000, 2)).astype(np.float32):
        L.oracle_mis_weight.restype = C.c_float
        L.oracle_mis_weight.argtypes = [C.c_float, C.c_float]
        assert shader0.mis_weight(float(a), float(b)) == L.oracle_mis_weight(float(a), float(b))


def test_transcendental_leaves_within_tolerance(shader0):
    """RotationMatrix, Emit, SpectralPowerDistribution, the samplers: libm vs pt_math.h."""
    import ctypes as C
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for ang in rng.uniform(-360, 360, (500, 3)).astype(np.float32):
        mine = np.zeros(9, np.float32)
        L.oracle_rotation_matrix(ang.ctypes.data_as(C.c_void_p), mine.ctypes.data_as(C.c_void_p))
        assert np.abs(shader0.rotation_matrix(ang).ravel() - mine).max() < 2e-6
    for _ in range(500):
        l4 = rng.uniform(390, 720, 4).astype(np.float32)
        T, lum = float(rng.uniform(1500, 12000)), float(rng.uniform(0.1, 50))
        mine = np.zeros(4, np.float32)
        L.oracle_emit(l4.ctypes.data_as(C.c_void_p), T, lum, mine.ctypes.data_as(C.c_void_p))
        assert np.allclose(shader0.emit(l4, T, lum), mine, rtol=2e-5, atol=0)
        peak, sig, inv = float(rng.uniform(400, 700)), float(rng.uniform(3, 12)), int(rng.integers(0, 2))
        L.oracle_spd(l4.ctypes.data_as(C.c_void_p), peak, sig, inv, mine.ctypes.data_as(C.c_void_p))
        assert np.allclose(shader0.spd(l4, peak, sig, inv), mine, rtol=2e-5, atol=1e-7)
    nrm = np.array([0.36, -0.48, 0.8], dtype=np.float32)
    for kind, param in ((0, 0.0), (1, 0.0), (2, 0.0), (3, 0.9), (4, 0.5)):
        mine = np.zeros((4000, 3), np.float32)
        L.oracle_sample(kind, 12345, 4000, C.c_float(param), nrm.ctypes.data_as(C.c_void_p), mine.ctypes.data_as(C.c_void_p))
        theirs = shader0.sample(kind, 12345, 4000, param, nrm)
        assert np.abs(theirs - mine).max() < 5e-6, kind     # same RNG stream, same draws, same order
    b1, b2 = shader0.orthonormal_basis(nrm)
    mine = np.zeros(6, np.float32)
    L.oracle_orthonormal_basis(nrm.ctypes.data_as(C.c_void_p), mine.ctypes.data_as(C.c_void_p))
    assert np.abs(np.concatenate([b1, b2]) - mine).max() < 1e-6


def test_accumulate_both_branches(shader0):
    """Accumulate (shader.comp:1492-1507): running mean and the EMA branch, against the oracle's dispatch semantics."""
    rs = ref.RefScene(scene_path('scene0'))
    a, b = np.array([0.25, 0.5, 0.75], np.float32), np.array([1.0, 2.0, 4.0], np.float32)
    # static branch, n = currentSamples / spf = 12 / 4 = 3  ->  ((n-1)*in + out)/n
    got = shader0.accumulate(rs.push(8, 8, 12, 12, 4), a, b)
    assert np.array_equal(got, ((np.float32(2) * a + b) / np.float32(3)).astype(np.float32))
    # EMA branch: currentSamples == spf and frame > spf; weight = 2^(-8/(FPS*persistence))
    push = rs.push(8, 8, 40, 4, 4, frame_time=1.0 / 30.0, persistence=0.5)
    got = shader0.accumulate(push, a, b)
    w = np.float32(2.0 ** (-8.0 / (float(params_of(push)['FPS']) * 0.5)))
    assert np.allclose(got, (1 - w) * b + w * a, rtol=1e-6)


def camera_rays(js, n, rng, spread=0.6):
    cam = np.array(js['camera']['position'][0], dtype=np.float32)
    yaw, pitch = np.radians(js['camera']['angle'][0])
    fwd = np.array([np.sin(yaw) * np.cos(pitch), np.sin(pitch), np.cos(yaw) * np.cos(pitch)])
    d = fwd[None, :] + rng.normal(0, spread, (n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = cam[None, :] + rng.normal(0, 0.3, (n, 3))
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.mark.parametrize('name', ['scene0', 'scene1', 'scene2', 'scene3', 'scene7', 'scene9', 'scene10', 'scene8'])
def test_intersection_on_random_rays(name):
    """Intersection() (shader.comp:862-934: every primitive type + sphere tracing) on rays around the camera and on
    secondary rays started from the hits.  Same object (material / light id) on >= 99.5 % of the rays; on those, hit
    distance within 1e-4 relative on >= 99 %.  What is left are grazing hits and the quartic / march thresholds."""
    import ctypes as C
    path = scene_path(name)
    rs = ref.RefScene(path)
    js = pack.load_scene(path)
    ubo = rs.ubo()
    O = oracle.Oracle(ubo, pack.sdf_sources(js))
    rng = np.random.default_rng(7)
    n = 3000 if name in SDF_SCENES else 6000
    rays = camera_rays(js, n, rng)
    push = rs.push(64, 64, 1, 1, 1)
    theirs = rs.shader().intersect(ubo, push, rays)
    # secondary rays: from the reference's hit points, uniformly random directions
    hit = theirs[:, 0] < 1e5
    o2 = rays[hit, :3] + rays[hit, 3:] * theirs[hit, :1] + theirs[hit, 1:4] * 1e-3
    d2 = rng.normal(0, 1, o2.shape)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    rays = np.concatenate([rays, np.concatenate([o2, d2], 1).astype(np.float32)])
    theirs = rs.shader().intersect(ubo, push, rays)
    mine = np.zeros_like(theirs)
    for i, r in enumerate(rays):
        t, out = O.intersect(r[:3], r[3:])
        mine[i, 0] = t
        mine[i, 1:] = out
    miss_t, miss_m = theirs[:, 0] >= 1e5, mine[:, 0] >= 1e5
    same_obj = (miss_t == miss_m) & (miss_t | ((np.abs(theirs[:, 4] - mine[:, 4]) < 1e-3) & (theirs[:, 5] == mine[:, 5])))
    assert same_obj.mean() >= 0.995, same_obj.mean()
    both = same_obj & ~miss_t
    rel = np.abs(theirs[both, 0] - mine[both, 0]) / np.maximum(np.abs(mine[both, 0]), 1e-6)
    print('%s: same object %.4f; t within 1e-4: %.4f, within 1e-3: %.4f, within 1e-2: %.4f' % (
        name, same_obj.mean(), (rel <= 1e-4).mean(), (rel <= 1e-3).mean(), (rel <= 1e-2).mean()))
    if name == 'scene8':
        # the terrain is not a distance bound (slopes > 1 from the high-frequency sines), so the over-relaxed march
        # overshoots and back-steps chaotically: an ulp in one SDF value can move the accepted point by a step
        assert (rel <= 1e-4).mean() >= 0.95 and (rel <= 1e-2).mean() >= 0.995
    else:
        assert (rel <= 1e-4).mean() >= 0.99, (rel <= 1e-4).mean()
    assert hit.mean() > 0.3      # the test really hits the scene
    ndot = np.sum(theirs[both, 1:4] * mine[both, 1:4], axis=1)
    if name != 'scene8':         # terrain normals are central differences of 22 sines at eps 1e-4: noise by design
        assert (ndot > 0.999).mean() >= 0.98, (ndot > 0.999).mean()


@pytest.mark.parametrize('name', SDF_SCENES)
def test_sdf_dispatchers_agree(name):
    """SDF() / SDFMATERIAL() as the reference's InsertSDF builds them vs the oracle's own translation of the snippets."""
    path = scene_path(name)
    rs = ref.RefScene(path)
    js = pack.load_scene(path)
    ubo = rs.ubo()
    O = oracle.Oracle(ubo, pack.sdf_sources(js))
    rng = np.random.default_rng(13)
    pos, size = np.array(js['sdf'][0]['position']), np.array(js['sdf'][0]['boundingSize'])
    pts = (pos + (rng.random((20000, 3)) - 0.5) * size * 1.2).astype(np.float32)
    d_t, m_t = rs.shader().sdf_eval(ubo, pts)
    d_m, m_m = O.sdf_eval(pts)
    assert np.abs(d_t - d_m).max() <= 1e-6 * max(1.0, np.abs(d_m).max())   # measured: <= 1e-6 absolute (terrain, |d| <= 10)
    assert (np.abs(m_t - m_m) < 1e-3).mean() > 0.999


def test_camera_lens_agrees():
    """TracePathLens (shader.comp:1409-1444): two refractions through the BK7 lens."""
    import ctypes as C
    path = scene_path('scene1')
    rs = ref.RefScene(path)
    ubo = rs.ubo()
    push = rs.push(1920, 1080, 1, 1, 1)
    p = params_of(push)
    L = oracle.lib()
    rng = np.random.default_rng(17)
    worst, checked = 0.0, 0
    for _ in range(400):
        out = np.zeros(9, np.float32)
        cam = np.array([p['cameraPosX'], p['cameraPosY'], p['cameraPosZ']], np.float32)
        # a sensor point and a direction toward the aperture, in world space, through the oracle's own camera frame
        o = cam + rng.normal(0, 0.002, 3).astype(np.float32)
        lh = float(rng.uniform(360, 800))
        L.oracle_lens_ray(ubo.ctypes.data_as(C.c_void_p), p.tobytes(), o.ctypes.data_as(C.c_void_p),
                          np.zeros(3, np.float32).ctypes.data_as(C.c_void_p), C.c_float(lh), out.ctypes.data_as(C.c_void_p))
        fwd = out[6:9].copy()
        d = fwd + rng.normal(0, 0.02, 3).astype(np.float32)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        L.oracle_lens_ray(ubo.ctypes.data_as(C.c_void_p), p.tobytes(), o.ctypes.data_as(C.c_void_p),
                          d.ctypes.data_as(C.c_void_p), C.c_float(lh), out.ctypes.data_as(C.c_void_p))
        o_t, d_t = rs.shader().lens_ray(ubo, push, lh, o, d, fwd)
        if np.all(np.isfinite(out[:6])) and np.all(np.isfinite(d_t)):
            worst = max(worst, float(np.abs(o_t - out[:3]).max()), float(np.abs(d_t - out[3:6]).max()))
    assert worst < 5e-5, worst


# ---------------------------------------------------------------------------------------------------------------
# whole samples and images
# ---------------------------------------------------------------------------------------------------------------

# measured agreement per scene (fraction of samples whose XYZ agrees to 1e-4 / 1e-3 relative), with margin; the rest
# are paths that forked at a threshold (hit / miss at a silhouette, roulette, 1e-4 march epsilon) after an ulp-level
# difference: cyclide quartics (scene0/2/3/7) and numerical SDF normals (terrain above all) amplify such differences.
PER_SAMPLE = {'scene0': (0.97, 0.98), 'scene1': (0.995, 0.997), 'scene2': (0.94, 0.96), 'scene9': (0.96, 0.97),
              'scene10': (0.975, 0.98), 'scene8': (0.50, 0.82), 'scene3': (0.93, 0.975)}


@pytest.mark.parametrize('name', sorted(PER_SAMPLE))
def test_per_sample_xyz_against_the_reference_shader(name):
    """Scene() of the reference shader vs oracle.cpp, sample by sample (1-sample dispatches, frame = k + 1), same seeds:
    black/non-black agree, the stated fractions agree numerically, and the images agree far below Monte-Carlo noise."""
    path = scene_path(name)
    rs = ref.RefScene(path)
    js = pack.load_scene(path)
    ubo = rs.ubo()
    O = oracle.Oracle(ubo, pack.sdf_sources(js))
    W, H, n = 64, 48, 6
    A = np.zeros((n, H, W, 3), np.float32)
    B = np.zeros_like(A)
    for k in range(n):
        push = rs.push(W, H, k + 1, 1, 1, 5)
        a = np.zeros((H, W, 4), np.float32)
        b = np.zeros((H, W, 4), np.float32)
        rs.dispatch(push, a)
        O.dispatch(params_of(push), b)
        assert np.all(a[..., 3] == 1.0) and np.all(b[..., 3] == 1.0)
        A[k], B[k] = a[..., :3], b[..., :3]
    za, zb = np.abs(A).max(axis=3) == 0, np.abs(B).max(axis=3) == 0
    rel = np.abs(A - B).max(axis=3) / np.maximum(np.abs(B).max(axis=3), 1e-20)
    ok4 = ((rel <= 1e-4) | (za & zb)).mean()
    ok3 = ((rel <= 1e-3) | (za & zb)).mean()
    print('%s: %.4f of samples within 1e-4, %.4f within 1e-3, forks %.4f' % (name, ok4, ok3, 1 - ok3))
    assert ok4 >= PER_SAMPLE[name][0] and ok3 >= PER_SAMPLE[name][1], (ok4, ok3)
    assert (za != zb).mean() < 0.02
    ia, ib = A.mean(axis=0), B.mean(axis=0)
    # the images: difference of the means against the sample-to-sample spread (a fork moves one sample)
    sigma = B.std(axis=0).mean() / np.sqrt(n)
    assert np.abs(ia.mean(axis=(0, 1)) - ib.mean(axis=(0, 1))).max() <= 0.02 * ib.mean() + 0.0
    assert np.sqrt(np.mean((ia - ib) ** 2)) < 0.35 * sigma * np.sqrt(3) + 1e-6


@pytest.mark.parametrize('name', GRADED)
def test_image_relrmse_below_the_monte_carlo_floor(name):
    """32x24 at 64 spp in dispatches of 8: relRMSE(reference shader, oracle) on the same sample indices is a small
    fraction of the relRMSE between two oracle renders with disjoint indices (the Monte-Carlo floor)."""
    path = scene_path(name)
    rs = ref.RefScene(path)
    js = pack.load_scene(path)
    O = oracle.Oracle(rs.ubo(), pack.sdf_sources(js))
    W, H, spp, spf = 32, 24, 64, 8
    theirs = rs.render(W, H, spp, spf)
    p = np.array(pack.pack_params(js, 1, W, H, spf, 5), copy=True)
    mine = O.render(p, spp, spf)
    other = np.zeros((H, W, 4), np.float32)
    for j in range(1, spp // spf + 1):      # disjoint sample indices, same bookkeeping
        q = np.array(p, copy=True)
        q['frame'] = (1 << 20) + j * spf
        q['currentSamples'] = j * spf
        O.dispatch(q, other)

    def relrmse(i, r):
        eps = (0.01 * r[..., :3].mean()) ** 2
        return float(np.sqrt(np.mean((i[..., :3] - r[..., :3]) ** 2 / (r[..., :3] ** 2 + eps))))
    floor = relrmse(other, mine)
    d = relrmse(theirs, mine)
    print('%s: relRMSE(ref, oracle) %.4f, Monte-Carlo floor %.4f' % (name, d, floor))
    assert d < 0.25 * floor, (d, floor)
    assert abs(float(theirs[..., 1].mean()) / float(mine[..., 1].mean()) - 1) < 0.01
