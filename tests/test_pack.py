"""The C++ scene loader / packer behind the C ABI vs the independent numpy restatement in oracle/pack.py."""
import json

import numpy as np
import pytest

from conftest import SCENES, scene_path
from oracle import pack


@pytest.mark.parametrize('name', SCENES)
def test_ubo_and_params_bit_identical(ptlib, name):
    sc = ptlib.Scene.load(scene_path(name))
    ref_scene = pack.load_scene(scene_path(name))
    ubo, ref = sc.pack_ubo(), pack.pack_ubo(ref_scene)
    assert ubo.shape == (4097,) and ubo.nbytes == 16388
    assert np.array_equal(ubo.view(np.uint32), ref.view(np.uint32))
    for shot in range(1, sc.num_shots + 1):
        p = sc.pack_params(shot, 640, 360, 3, 7)
        q = pack.pack_params(ref_scene, shot, 640, 360, 3, 7)
        assert p.tobytes() == q.tobytes()
    assert [s.decode() for s in sc.sdf_sources] == pack.sdf_sources(ref_scene)


def test_params_layout_is_the_push_constant_block(ptlib):
    d = ptlib.PARAMS_DTYPE
    offs = {n: d.fields[n][1] for n in d.names}
    assert d.itemsize == 88
    assert offs == {'resolution': 0, 'frame': 8, 'currentSamples': 12, 'samplesPerFrame': 16, 'FPS': 20,
                    'persistence': 24, 'pathLength': 28, 'cameraAngle': 32, 'cameraPosX': 40, 'cameraPosY': 44,
                    'cameraPosZ': 48, 'ISO': 52, 'cameraSize': 56, 'apertureSize': 60, 'apertureDist': 64,
                    'lensRadius': 68, 'lensFocalLength': 72, 'lensThickness': 76, 'lensDistance': 80, 'tonemap': 84}


def test_light_registration_quirk(ptlib):
    """host:3704,3728: lenses and cyclides are registered as sampled lights iff planes[i].lightID > 0."""
    s = pack.load_scene(scene_path('scene0'))
    s['lens'][0]['lightID'] = 1          # emits when hit ...
    ubo = ptlib.Scene.parse(json.dumps(s)).pack_ubo()
    assert ubo[6] == 1                   # ... but is not sampled: plane 0 has lightID 0
    s['plane'][0]['lightID'] = 1
    ubo = ptlib.Scene.parse(json.dumps(s)).pack_ubo()
    ref = pack.pack_ubo(s)
    assert np.array_equal(ubo.view(np.uint32), ref.view(np.uint32))
    assert ubo[6] == 4                   # sphere 2, plane 0, lens 0 and cyclide 0 (both keyed on plane 0)
    assert list(ubo[pack.OFF_LID:pack.OFF_LID + 4]) == [2, 3, 5, 6]


def test_missing_arrays_are_legal(ptlib):
    ubo = ptlib.Scene.load(scene_path('scene8')).pack_ubo()   # no plane / box / lens / cyclide keys
    assert list(ubo[:7]) == [2, 0, 0, 0, 0, 1, 1]


def test_bad_scene_is_an_error(ptlib):
    with pytest.raises(ptlib.PtError):
        ptlib.Scene.parse('{"camera": {"numShots": 3, "position": [], "angle": []}}')
    with pytest.raises(ptlib.PtError):
        ptlib.Scene.parse('{not json')
    with pytest.raises(ptlib.PtError):
        ptlib.Scene.load('/nonexistent/scene.json')
    with pytest.raises(ptlib.PtError):
        ptlib.Scene.load(scene_path('scene0')).pack_params(shot=9)


def test_cie_table_exported(ptlib):
    from pathtracer_b200 import api
    assert np.array_equal(api.cie1931_table(), pack.cie_table())


@pytest.mark.parametrize('name', SCENES)
def test_scene_json_round_trip(ptlib, name, tmp_path):
    """UpdateToJSON twin (host:2724-2858): load -> save -> load gives the same uniform block (the shipped files are
    already rounded to 1e-5, the precision SaveScene writes) and the same key order as the shipped file."""
    sc = ptlib.Scene.load(scene_path(name))
    text = sc.to_json()
    again = ptlib.Scene.parse(text)
    assert np.array_equal(sc.pack_ubo().view(np.uint32), again.pack_ubo().view(np.uint32))
    for shot in range(1, sc.num_shots + 1):
        assert sc.pack_params(shot, 64, 64, 1, 5).tobytes() == again.pack_params(shot, 64, 64, 1, 5).tobytes()
    assert [s for s in sc.sdf_sources] == [s for s in again.sdf_sources]
    ours, ref = json.loads(text), json.load(open(scene_path(name)))
    assert list(ours.keys()) == [k for k in ref.keys() if ref[k] != []]
    assert list(ours['camera'].keys()) == list(ref['camera'].keys())
    out = tmp_path / 'scene.json'
    sc.save(out)
    assert out.read_text() == text


def test_round_decimal_semantics(ptlib):
    s = pack.load_scene(scene_path('scene0'))
    s['sphere'][0]['radius'] = 1.234567
    s['sphere'][1]['position'] = [-0.000004, -1.999996, 2.5]
    out = json.loads(ptlib.Scene.parse(json.dumps(s)).to_json())
    assert out['sphere'][0]['radius'] == 1.23457
    assert out['sphere'][1]['position'] == [0.0, -2.0, 2.5]


def test_capacity_is_the_uniform_block(ptlib):
    """Any scene the reference's 1024-float objects[] can describe is accepted (170 spheres, or 93 boxes, ...); the
    kernel for it builds (NVRTC, no GPU needed)."""
    import ctypes as C
    from pathtracer_b200 import api
    base = pack.load_scene(scene_path('scene1'))
    L = ptlib.lib()
    L.pt_kernel_compile_check.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int]
    for key, n in (('sphere', 170), ('box', 93), ('plane', 204)):
        scene = {'camera': base['camera'], key: [dict(base[key][0]) for _ in range(n)], 'material': base['material'], 'light': base['light']}
        ubo = ptlib.Scene.parse(json.dumps(scene)).pack_ubo()
        assert np.array_equal(ubo.view(np.uint32), pack.pack_ubo(scene).view(np.uint32))
        assert L.pt_kernel_compile_check(ubo.ctypes.data_as(C.c_void_p), api._c_strings([]), 0, 1, 0) == 0, L.pt_last_error(None)
    bad = pack.pack_ubo(base)
    bad[0] = 400.0   # a hand-made block claiming more spheres than objects[] holds
    assert L.pt_kernel_compile_check(bad.ctypes.data_as(C.c_void_p), api._c_strings([]), 0, 1, 0) == -1


def test_loader_survives_mutated_scene_files(ptlib):
    """Scene files come from outside: byte flips, truncations, deleted chunks and absurd nesting of a shipped scene are
    either parsed (and then pack, save and report their extensions without incident) or refused with PT_ERR_IO /
    PT_ERR_ARG -- the loader never crashes the host (the reference throws from nlohmann::json, host:893-908)."""
    import random
    rnd = random.Random(7)
    base = open(scene_path('scene10')).read()
    parsed = refused = 0
    for it in range(800):
        s = list(base)
        mode = it % 4
        if mode == 0:
            for _ in range(rnd.randint(1, 8)):
                s[rnd.randrange(len(s))] = chr(rnd.choice([0x20, 0x22, 0x2c, 0x5b, 0x5d, 0x7b, 0x7d, 0x30, 0x2d, 0x65, 0x2e, 0x5c, 0x0a, 0x01, 0xe9]))
        elif mode == 1:
            s = s[:rnd.randrange(len(s))]
        elif mode == 2:
            i = rnd.randrange(len(s))
            del s[i:i + rnd.randint(1, 200)]
        else:
            i = rnd.randrange(len(s))
            s[i:i] = list(rnd.choice(['[', '{"a":']) * rnd.randint(1, 5000))
        try:
            sc = ptlib.Scene.parse(''.join(s))
            sc.pack_ubo(), sc.to_json(), sc.surface_ext(), sc.pack_params(1, 64, 64, 1, 5)
            parsed += 1
        except ptlib.PtError as e:
            assert e.code in (-1, -4), e
            refused += 1
    assert parsed > 50 and refused > 300, (parsed, refused)
