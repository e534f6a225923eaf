"""numpy restatement of the reference's scene loader and uniform/push-constant packer -- TEST INFRASTRUCTURE.

Follows src/pathtracer.cpp of the reference: UpdateFromJSON (host:2576-2722), UpdateUniformBuffer (host:3642-3811,
including the light-registration quirk of host:3704/3728 that tests planes[i].lightID for lenses and cyclides),
UpdatePushConstant (host:3813-3834) and the offscreen frame bookkeeping (host:4042-4048).
Independent of pathtracer_b200/csrc/pt_scene.cpp; tests/test_pack.py compares the two bit for bit.
"""
import json
import re
import numpy as np

MAX_OBJECTS, MAX_SDFS, MAX_MATERIALS, MAX_LIGHTS, MAX_LIGHTIDS, CIE_SIZE = 1024, 768, 783, 128, 64, 1323
OFF_OBJ, OFF_SDF = 7, 7 + 1024
OFF_MAT = OFF_SDF + MAX_SDFS
OFF_LGT = OFF_MAT + MAX_MATERIALS
OFF_LID = OFF_LGT + MAX_LIGHTS
OFF_CIE = OFF_LID + MAX_LIGHTIDS
UBO_FLOATS = OFF_CIE + CIE_SIZE  # 4097

PARAMS_DTYPE = np.dtype([
    ('resolution', '<i4', (2,)), ('frame', '<i4'), ('currentSamples', '<i4'), ('samplesPerFrame', '<i4'),
    ('FPS', '<f4'), ('persistence', '<f4'), ('pathLength', '<i4'), ('cameraAngle', '<f4', (2,)),
    ('cameraPosX', '<f4'), ('cameraPosY', '<f4'), ('cameraPosZ', '<f4'), ('ISO', '<i4'), ('cameraSize', '<f4'),
    ('apertureSize', '<f4'), ('apertureDist', '<f4'), ('lensRadius', '<f4'), ('lensFocalLength', '<f4'),
    ('lensThickness', '<f4'), ('lensDistance', '<f4'), ('tonemap', '<i4')])
assert PARAMS_DTYPE.itemsize == 88


def cie_table(path=None):
    """The 1323 floats of include/pt_cie1931.inc (public CIE 1931 2-degree data; see tools/extract_cie.py)."""
    import os
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'include', 'pt_cie1931.inc')
    txt = open(path).read()
    txt = txt[txt.index('*/') + 2:]
    vals = [float(t[:-1]) for t in re.findall(r'[-+0-9.eE]+f', txt)]
    assert len(vals) == CIE_SIZE
    return np.array(vals, dtype=np.float32)


def load_scene(path):
    with open(path) as f:
        return json.load(f)


def f32(x):
    return np.float32(x)


def pack_ubo(scene):
    """UpdateUniformBuffer (host:3642-3811). Returns float32[4097]."""
    sph, pla, box = scene.get('sphere', []), scene.get('plane', []), scene.get('box', [])
    lens, cyc, sdf = scene.get('lens', []), scene.get('cyclide', []), scene.get('sdf', [])
    mats, lights = scene.get('material', []), scene.get('light', [])
    objs, light_ids = [], []
    for i, s in enumerate(sph):
        objs += [*s['position'], s['radius'], float(int(s['materialID'])), float(int(s['lightID']))]
        if int(s['lightID']) > 0:
            light_ids.append(i)
    for i, p in enumerate(pla):
        objs += [*p['position'], float(int(p['materialID'])), float(int(p['lightID']))]
        if int(p['lightID']) > 0:
            light_ids.append(len(sph) + i)
    for i, b in enumerate(box):
        objs += [*b['position'], *b['rotation'], *b['size'], float(int(b['materialID'])), float(int(b['lightID']))]
        if int(b['lightID']) > 0:
            light_ids.append(len(sph) + len(pla) + i)
    for i, l in enumerate(lens):
        objs += [*l['position'], *l['rotation'], l['radius'], l['focalLength'], l['thickness'],
                 1.0 if l['isConverging'] else 0.0, float(int(l['materialID'])), float(int(l['lightID']))]
        # host:3704 tests planes[i].lightID (sic).  Reading past the end of `planes` is undefined in the
        # reference; defined here as "not registered".
        if i < len(pla) and int(pla[i]['lightID']) > 0:
            light_ids.append(len(sph) + len(pla) + len(box) + i)
    for i, c in enumerate(cyc):
        sc = [f32(v) for v in c['scale']]
        m = max(max(sc[0], sc[1]), sc[2])
        brad = f32(c['boundingRadius'])
        packed_brad = f32(f32(brad * brad) * f32(m * m))  # host:3723-3725, fp32 arithmetic
        objs += [*c['position'], *c['rotation'], *c['scale'], c['a'], c['b'], c['c'], c['d'], packed_brad,
                 float(int(c['materialID'])), float(int(c['lightID']))]
        if i < len(pla) and int(pla[i]['lightID']) > 0:  # host:3728 (sic)
            light_ids.append(len(sph) + len(pla) + len(box) + len(lens) + i)
    ubo = np.zeros(UBO_FLOATS, dtype=np.float32)
    ubo[0:7] = [len(sph), len(pla), len(box), len(lens), len(cyc), len(sdf), len(light_ids)]

    def put(off, cap, vals):
        v = np.array(vals, dtype=np.float64).astype(np.float32)[:cap]
        ubo[off:off + len(v)] = v

    put(OFF_OBJ, MAX_OBJECTS, objs)
    put(OFF_SDF, MAX_SDFS, [v for s in sdf for v in (*s['position'], *s['boundingSize'])])
    put(OFF_MAT, MAX_MATERIALS, [v for m in mats for v in (m['reflection']['peakWavelength'], m['reflection']['sigma'],
                                                           1.0 if m['reflection']['isInvert'] else 0.0)])
    put(OFF_LGT, MAX_LIGHTS, [v for l in lights for v in (l['emission']['temperature'], l['emission']['luminosity'])])
    put(OFF_LID, MAX_LIGHTIDS, [float(v) for v in light_ids])
    ubo[OFF_CIE:] = cie_table()
    return ubo


def pack_params(scene, shot=1, width=512, height=512, spf=1, path_length=5, dispatch=1, tonemap=3):
    """UpdatePushConstant (host:3813-3834) for offscreen dispatch number `dispatch` (1-based, host:4042-4048)."""
    cam = scene['camera']
    p = np.zeros((), dtype=PARAMS_DTYPE)
    p['resolution'] = (width, height)
    p['frame'] = dispatch * spf
    p['currentSamples'] = dispatch * spf
    p['samplesPerFrame'] = spf
    p['FPS'] = 60.0
    p['persistence'] = 0.0625  # host:1174
    p['pathLength'] = path_length
    ang = cam['angle'][shot - 1]
    p['cameraAngle'] = (-f32(ang[1]), f32(ang[0]))  # host:3821
    pos = cam['position'][shot - 1]
    p['cameraPosX'], p['cameraPosY'], p['cameraPosZ'] = pos
    p['ISO'] = int(cam['ISO'])
    p['cameraSize'] = cam['size']
    p['apertureSize'] = cam['apertureSize']
    p['apertureDist'] = cam['apertureDistance']
    p['lensRadius'] = cam['lensRadius']
    p['lensFocalLength'] = cam['lensFocalLength']
    p['lensThickness'] = cam['lensThickness']
    p['lensDistance'] = cam['lensDistance']
    p['tonemap'] = tonemap
    return p


SURFACE_EXT_DTYPE = np.dtype([('bsdf', '<i4'), ('roughness', '<f4'), ('ior', '<f4'), ('pad', '<f4')])  # pt_surface_ext


def surface_ext(scene):
    """The optional material keys "bsdf" / "roughness" / "ior" (an extension of the reference's schema; no
    shipped scene has them) as a pt_surface_ext table: entry i extends material i; empty when no material has a "bsdf"."""
    names = {'reference': 0, 'diffuse': 0, 'mirror': 1, 'glossy': 2, 'dielectric': 3}
    mats = scene.get('material', [])
    n = max([i + 1 for i, m in enumerate(mats) if names[m.get('bsdf', 'reference')] != 0], default=0)
    t = np.zeros(n, dtype=SURFACE_EXT_DTYPE)
    for i in range(n):
        t[i]['bsdf'] = names[mats[i].get('bsdf', 'reference')]
        t[i]['roughness'] = mats[i].get('roughness', 0.0)
        t[i]['ior'] = mats[i].get('ior', 0.0)
    return t


def sdf_sources(scene):
    return [s['glsl'] for s in scene.get('sdf', [])]
