"""The SDF plug-in front end (pt_sdf_front.cpp) on the CPU: the text it feeds NVRTC also compiles with g++, so its
output can be compared bit for bit with the oracle's independent translation (oracle/sdf_build.py); and NVRTC
compiles the whole kernel with the shipped snippets (needs no GPU)."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, scene_path
from oracle import oracle, pack, sdf_build

SDF_SCENES = ['scene3', 'scene4', 'scene6', 'scene8', 'scene9', 'scene10']


def build_product_unit(ptlib, sources, raw):
    from pathtracer_b200 import api
    text = api.sdf_translate(sources, raw)
    tag = hashlib.sha1(text.encode()).hexdigest()[:16]
    os.makedirs(sdf_build.BUILD, exist_ok=True)
    so = os.path.join(sdf_build.BUILD, 'prod_%s.so' % tag)
    if not os.path.exists(so):
        cpp = so[:-3] + '.cpp'
        open(cpp, 'w').write(text)
        subprocess.run([sdf_build.CXX, *sdf_build.CXXFLAGS, '-o', so, cpp], check=True)
    lib = C.CDLL(so)
    for f in (lib.pt_sdf_dispatch, lib.pt_sdfmaterial_dispatch):
        f.restype = C.c_float
        f.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]

    class Unit:  # the dispatchers take the shader's four masks; the tests below mostly use the first
        @staticmethod
        def pt_sdf_dispatch(x, y, z, set1, set2=0, set3=0, set4=0):
            return lib.pt_sdf_dispatch(x, y, z, set1, set2, set3, set4)

        @staticmethod
        def pt_sdfmaterial_dispatch(x, y, z, set1, set2=0, set3=0, set4=0):
            return lib.pt_sdfmaterial_dispatch(x, y, z, set1, set2, set3, set4)
    return Unit, text


@pytest.mark.parametrize('name', SDF_SCENES)
def test_front_end_matches_oracle_translation(ptlib, name):
    scene = pack.load_scene(scene_path(name))
    ubo = pack.pack_ubo(scene)
    src = pack.sdf_sources(scene)
    L, text = build_product_unit(ptlib, src, ubo[pack.OFF_SDF:pack.OFF_SDF + 6 * len(src)])
    assert 'SDF1' in text and 'SDF1MATERIAL' in text
    o = oracle.Oracle(ubo, src)
    rng = np.random.default_rng(7)
    pos = ubo[pack.OFF_SDF:pack.OFF_SDF + 3]
    size = ubo[pack.OFF_SDF + 3:pack.OFF_SDF + 6]
    pts = (pos + (rng.random((4000, 3)) - 0.5) * size * 1.2).astype(np.float32)
    d_ref, m_ref = o.sdf_eval(pts, 1)
    d = np.array([L.pt_sdf_dispatch(*map(float, p), 1) for p in pts], dtype=np.float32)
    m = np.array([L.pt_sdfmaterial_dispatch(*map(float, p), 1) for p in pts], dtype=np.float32)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))
    assert np.array_equal(m.view(np.uint32), m_ref.view(np.uint32))
    assert np.isfinite(d).all()
    # set1 == 0 evaluates nothing: MAXDIST and material 0
    assert L.pt_sdf_dispatch(0.0, 0.0, 0.0, 0) == np.float32(1e5)
    assert L.pt_sdfmaterial_dispatch(0.0, 0.0, 0.0, 0) == 0.0


def test_shipped_sdf_files_load_unchanged(ptlib):
    """sdfs/*.glsl (what the reference's "Change SDF" button loads, host:3474-3480) go through both translators."""
    for f in ('blob', 'mandelbulb', 'menger', 'terrain'):
        src = open(os.path.join(ROOT, 'sdfs', f + '.glsl')).read()
        raw = np.array([0, 0, 0, 4, 4, 4], dtype=np.float32)
        L, _ = build_product_unit(ptlib, [src], raw)
        ubo = np.zeros(pack.UBO_FLOATS, dtype=np.float32)
        ubo[5] = 1
        ubo[pack.OFF_SDF:pack.OFF_SDF + 6] = raw
        o = oracle.Oracle(ubo, [src])
        pts = (np.random.default_rng(3).random((500, 3)).astype(np.float32) - 0.5) * 3
        d_ref, m_ref = o.sdf_eval(pts, 1)
        d = np.array([L.pt_sdf_dispatch(*map(float, p), 1) for p in pts], dtype=np.float32)
        assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))


def test_two_sdfs_dispatch_order_and_masks(ptlib):
    """SDF() takes the min over set bits; SDFMATERIAL() returns the material of the nearest (host:2023-2051)."""
    a = 'float sdf(in vec3 p) { return length(p) - 1.0; }\nfloat sdfmaterial(in vec3 p) { return 2.0; }\n'
    b = 'float sdf(in vec3 p) { return length(p) - 0.5; }\nfloat sdfmaterial(in vec3 p) { return 5.0; }\n'
    raw = np.array([0, 0, 0, 2, 2, 2, 3, 0, 0, 1, 1, 1], dtype=np.float32)
    L, text = build_product_unit(ptlib, [a, b], raw)
    body = text[text.index('float SDFMATERIAL('):]
    assert body.index('SDF1MATERIAL(p') < body.index('sdf = min(sdf, SDF1(') < body.index('SDF2MATERIAL(p')  # host:2049-2050
    assert abs(L.pt_sdf_dispatch(0, 0, 0, 1) + 1.0) < 1e-7
    assert abs(L.pt_sdf_dispatch(0, 0, 0, 2) - 2.5) < 1e-6
    assert abs(L.pt_sdf_dispatch(0, 0, 0, 3) + 1.0) < 1e-7
    assert L.pt_sdfmaterial_dispatch(0, 0, 0, 3) == 2.0
    assert L.pt_sdfmaterial_dispatch(3, 0, 0, 3) == 5.0
    ubo = np.zeros(pack.UBO_FLOATS, dtype=np.float32)
    ubo[5] = 2
    ubo[pack.OFF_SDF:pack.OFF_SDF + 12] = raw
    o = oracle.Oracle(ubo, [a, b])
    pts = (np.random.default_rng(5).random((300, 3)).astype(np.float32) - 0.3) * 5
    for mask in (1, 2, 3):
        d_ref, m_ref = o.sdf_eval(pts, mask)
        d = np.array([L.pt_sdf_dispatch(*map(float, p), mask) for p in pts], dtype=np.float32)
        m = np.array([L.pt_sdfmaterial_dispatch(*map(float, p), mask) for p in pts], dtype=np.float32)
        assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)) and np.array_equal(m, m_ref)


def test_more_than_32_sdfs_use_the_other_masks(ptlib):
    """InsertSDF's dispatcher line for SDF i reads mask set<i/32 + 1>, bit i % 32 (host:2012, 2029-2033); the front end
    and the oracle's translator both reproduce that for up to 128 snippets, and 129 are refused."""
    from pathtracer_b200 import api
    n = 70
    srcs = ['float sdf(in vec3 p) { return length(p) - %.2f; }\nfloat sdfmaterial(in vec3 p) { return %d.0; }\n' % (0.1 + 0.01 * i, i % 5) for i in range(n)]
    raw = np.zeros(6 * n, dtype=np.float32)
    for i in range(n):
        raw[6 * i:6 * i + 6] = [0.3 * i, 0.0, 0.0, 1.0, 1.0, 1.0]
    L, text = build_product_unit(ptlib, srcs, raw)
    assert 'if ((set2 & 1u) == 1u) sdf = min(sdf, SDF33(' in text and 'if ((set3 & 32u) == 32u) sdf = min(sdf, SDF70(' in text
    ubo = np.zeros(pack.UBO_FLOATS, dtype=np.float32)
    ubo[5] = n
    ubo[pack.OFF_SDF:pack.OFF_SDF + 6 * n] = raw
    o = oracle.Oracle(ubo, srcs)
    pts = (np.random.default_rng(2).random((200, 3)).astype(np.float32)) * np.array([21.0, 1.0, 1.0], dtype=np.float32)
    for words in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 32, 0), (0xFFFFFFFF, 0xFFFFFFFF, 0x3F, 0), (0, 0x80000000, 1, 0)):
        d_ref, m_ref = o.sdf_eval(pts, words)
        d = np.array([L.pt_sdf_dispatch(*map(float, q), *words) for q in pts], dtype=np.float32)
        m = np.array([L.pt_sdfmaterial_dispatch(*map(float, q), *words) for q in pts], dtype=np.float32)
        assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)) and np.array_equal(m, m_ref)
    assert abs(L.pt_sdf_dispatch(9.6, 0.0, 0.0, 0, 1, 0, 0) + 0.42) < 1e-6      # SDF 33 sits at x = 9.6, radius 0.42
    with pytest.raises(ptlib.PtError):
        api.sdf_translate(srcs[:1] * 129, np.zeros(6 * 129, dtype=np.float32))


def test_rewrite_rules(ptlib):
    from pathtracer_b200 import api
    src = ('float helper(inout vec3 q, out float w, in float s) { w = 2.; q.xy = q.yx; return 1e-3 * s + .5 + 3; }\n'
           '// a comment with sdf in it would break the reference too\n'
           'float sdf(in vec3 p) { float w; vec3 q = p.zyx; return helper(q, w, 1.5f) + length(p.xz) - 1.0e+0; }\n'
           'float sdfmaterial(in vec3 p) { return 0.0; }\n')
    # the reference renames the FIRST "sdf" substring (host:2015): here that is inside the comment -> a compile error
    # there; our front end reproduces the rename and then drops the comment, so sdf() is left un-renamed -> error
    text = api.sdf_translate([src.replace('// a comment with sdf in it would break the reference too\n', '')])
    assert 'vec3& q' in text and 'float& w' in text and 'float s' in text and ' in ' not in text.split('snippet 1')[1]
    assert '2.f' in text and '1e-3f' in text and '.5f' in text and '1.5f' in text and '1.0e+0f' in text and '+ 3;' in text
    assert 'p.sw3<2,1,0>()' in text and 'p.sw2<0,2>()' in text and 'q.lsw2<0,1>() = q.sw2<1,0>()' in text


def test_errors_are_reported(ptlib):
    from pathtracer_b200 import api
    with pytest.raises(ptlib.PtError) as e:
        api.sdf_translate(['float distance_only(in vec3 p) { return 0.0; }'])
    assert e.value.code == -2
    with pytest.raises(ptlib.PtError) as e:
        api.sdf_compile_check(['float sdf(in vec3 p) { return undefined_function(p); }\nfloat sdfmaterial(in vec3 p) { return 0.0; }'])
    assert e.value.code == -2 and 'undefined_function' in str(e.value)


@pytest.mark.parametrize('name', ['scene9', 'scene10', 'scene8', 'scene3'])
@pytest.mark.parametrize('mode', [0, 1])
def test_nvrtc_builds_the_kernel_with_shipped_snippets(ptlib, name, mode):
    from pathtracer_b200 import api
    scene = pack.load_scene(scene_path(name))
    ubo = pack.pack_ubo(scene)
    src = pack.sdf_sources(scene)
    api.sdf_compile_check(src, ubo[pack.OFF_SDF:pack.OFF_SDF + 6 * len(src)], mode)


@pytest.mark.parametrize('name,mode,options', [
    ('scene10', 1, {'sched': 5}), ('scene10', 1, {'sched': 8}), ('scene10', 1, {'sched': 7}), ('scene10', 1, {'sched': 0}),
    ('scene10', 1, {'sched': 8, 'pool_cap': 16, 'pool_min': 12, 'min_blocks': 4}), ('scene10', 0, {'sched': 5}), ('scene10', 0, {'sched': 8}),
    ('scene10', 0, {'sched': 8, 'steal_s': 16}),   # the pool shrinks until table + pool fit the 48 KB of static shared memory
    ('scene9', 1, {}), ('scene9', 1, {'sched': 8}), ('scene8', 1, {'sched': 5, 'steal_s': 8}), ('scene8', 1, {'sched': 8}), ('scene3', 1, {'sched': 8}),
    ('scene1', 1, {'sched': 0}), ('scene1', 1, {'sched': 5}), ('scene1', 1, {'sched': 7}), ('scene1', 0, {'sched': 7}), ('scene1', 1, {'sched': 8}),
    ('scene0', 1, {'sched': 7}), ('scene2', 1, {'sched': 7}), ('scene10', 1, {'stats': 1, 'sched': 8}),
    ('scene10', 1, {'sched': 5, 'pregen': 1}), ('scene10', 0, {'sched': 5, 'pregen': 1}), ('scene1', 1, {'sched': 7, 'pregen': 1}), ('scene8', 1, {'pregen': 1})])
def test_every_driver_of_the_megakernel_builds(ptlib, name, mode, options):
    """The scene-specialised kernel (jit policy 2) compiles with NVRTC for sm_100a under every driver / option the A/B
    measurements use (no GPU needed), stays inside the 48 KB of static shared memory and the register budget of its
    launch bounds, and spills at most a few words."""
    import re
    scene = pack.load_scene(scene_path(name))
    ubo = pack.pack_ubo(scene)
    log = ptlib.kernel_compile_check(ubo, pack.sdf_sources(scene), mode, True, options)
    m = re.search(r"Compiling entry function 'pt_render_jit'.*?Used (\d+) registers.*?(\d+) bytes smem", log, re.S)
    assert m, log
    regs, smem = int(m.group(1)), int(m.group(2))
    assert smem <= 48 * 1024
    assert regs <= 128
    spills = [int(x) for x in re.findall(r'(\d+) bytes spill stores', log)]
    assert max(spills) <= (256 if options.get('sched') == 8 else 192), log  # v2m (not a default) carries a pool's worth of state


def test_generation_kernel_is_built_where_it_pays(ptlib):
    """Option pregen, auto: scenes with a cyclide (their kernel outgrows the instruction cache with the camera inline) get
    the generation kernel pt_gen_jit next to the render kernel, which then is some 260 instructions shorter; the others,
    and the drivers that do not pool samples, do not -- unless asked."""
    import re

    def entries(name, options):
        scene = pack.load_scene(scene_path(name))
        log = ptlib.kernel_compile_check(pack.pack_ubo(scene), pack.sdf_sources(scene), 1, True, options)
        return set(re.findall(r"Compiling entry function '(\w+)'", log))
    assert 'pt_gen_jit' in entries('scene10', {}) and 'pt_gen_jit' in entries('scene0', {}) and 'pt_gen_jit' in entries('scene9', {})
    assert 'pt_gen_jit' not in entries('scene1', {}) and 'pt_gen_jit' not in entries('scene8', {})
    assert 'pt_gen_jit' in entries('scene1', {'pregen': 1}) and 'pt_gen_jit' not in entries('scene10', {'pregen': 0})
    assert 'pt_gen_jit' not in entries('scene10', {'sched': 0, 'pregen': 1}) and 'pt_gen_jit' not in entries('scene10', {'sched': 8, 'pregen': 1})


def test_unknown_option_is_an_error(ptlib):
    scene = pack.load_scene(scene_path('scene1'))
    with pytest.raises(ptlib.PtError):
        ptlib.kernel_compile_check(pack.pack_ubo(scene), [], 1, True, {'sched': 6})
    with pytest.raises(ptlib.PtError):
        ptlib.kernel_compile_check(pack.pack_ubo(scene), [], 1, True, {'nonsense': 1})
