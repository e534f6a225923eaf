"""Host-side logic behind the C ABI that needs no GPU: the JSON reader/writer, error reporting, the bench contract."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import ROOT, scene_path
from oracle import pack


def minimal_scene(**extra):
    s = {'camera': {'numShots': 1, 'position': [[0, 1, -5]], 'angle': [[0, 0]], 'ISO': 100, 'size': 0.05, 'apertureSize': 0.002,
                    'apertureDistance': 0.049, 'lensRadius': 0.01, 'lensFocalLength': 0.03, 'lensThickness': 0.0, 'lensDistance': 0.05}}
    s.update(extra)
    return s


def test_json_escapes_and_unicode(ptlib):
    glsl = 'float sdf(in vec3 p) { return length(p) - 1.0; } // "quoted" \\ back\tslash é 中\nfloat sdfmaterial(in vec3 p) { return 0.0; }\r\n'
    s = minimal_scene(sdf=[{'position': [0, 0, 0], 'boundingSize': [2, 2, 2], 'glsl': glsl}])
    for text in (json.dumps(s), json.dumps(s, ensure_ascii=False), json.dumps(s, indent=3)):
        sc = ptlib.Scene.parse(text)
        assert sc.sdf_sources[0].decode() == glsl
        assert json.loads(sc.to_json())['sdf'][0]['glsl'] == glsl
    surrogate = json.dumps(minimal_scene(sdf=[{'position': [0, 0, 0], 'boundingSize': [1, 1, 1], 'glsl': 'sdf sdfmaterial \U0001F600'}]))
    assert '\\ud83d' in surrogate
    assert ptlib.Scene.parse(surrogate).sdf_sources[0].decode().endswith('\U0001F600')


@pytest.mark.parametrize('bad', ['', '{', '{"camera": }', '{"camera": {"numShots": 1,}}', '[1, 2', '{"a": 01}', '{"a": "\\x"}',
                                 '{"camera": {"numShots": 1}} trailing', '{"camera": nul}', '"just a string"'])
def test_malformed_json_is_reported(ptlib, bad):
    with pytest.raises(ptlib.PtError) as e:
        ptlib.Scene.parse(bad)
    assert e.value.code == -4


def test_number_forms(ptlib):
    text = json.dumps(minimal_scene()).replace('"ISO": 100', '"ISO": 1.6e3').replace('"size": 0.05', '"size": 5E-2')
    p = ptlib.Scene.parse(text).pack_params(1, 8, 8, 1, 5)
    assert int(p['ISO']) == 1600 and float(p['cameraSize']) == np.float32(0.05)


finite = st.floats(min_value=-1e4, max_value=1e4, allow_nan=False, allow_infinity=False, width=32)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(finite, finite, finite, st.floats(min_value=0.0625, max_value=50, width=32), st.integers(0, 5), st.integers(0, 3)),
                min_size=0, max_size=20))
def test_random_sphere_scenes_pack_like_the_numpy_restatement(spheres):
    import pathtracer_b200 as pt
    s = minimal_scene(sphere=[{'position': [x, y, z], 'radius': r, 'materialID': m, 'lightID': l} for x, y, z, r, m, l in spheres],
                      material=[{'reflection': {'peakWavelength': 550.0, 'sigma': 30.0, 'isInvert': False}}],
                      light=[{'emission': {'temperature': 5500.0, 'luminosity': 10.0}}])
    sc = pt.Scene.parse(json.dumps(s))
    assert np.array_equal(sc.pack_ubo().view(np.uint32), pack.pack_ubo(s).view(np.uint32))
    # save -> load keeps everything that survives the reference's 1e-5 rounding
    again = json.loads(sc.to_json())
    for a, b in zip(again.get('sphere', []), s['sphere']):
        assert np.allclose(a['position'], b['position'], atol=6e-6, rtol=1e-6) and a['materialID'] == b['materialID']


@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(min_value=2**-20, max_value=2**20, allow_nan=False, width=32), min_size=1, max_size=6))
def test_front_end_suffixes_every_float_literal(values):
    from pathtracer_b200 import api
    lits = [repr(float(v)) for v in values]
    body = ' + '.join(lits)
    src = 'float sdf(in vec3 p) { return p.x * (%s); }\nfloat sdfmaterial(in vec3 p) { return 0.0; }\n' % body
    text = api.sdf_translate([src])
    snippet = text[text.index('snippet 1'):text.index('shader.comp:706-711')]
    for lit in lits:
        assert lit + 'f' in snippet
    import re
    assert not re.search(r'\d\.\d+(?![\dfeE])', snippet.split('*/', 1)[1].replace('p.x', ''))


def test_bench_reference_arm_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--workload', 'cfg1_scene0_512'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().split('\n')[-1])
    assert d['impl'] == 'reference' and d['value'] > 0 and d['unit'] == 'samples/s' and d['higher_is_better'] is True
    from oracle import ref
    assert d['cpu_baseline']['kind'] == ('reference' if ref.available() else 'port')   # oracle/_ref when it was built
    assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))                   # all host cores ...
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'] == 'cfg1_scene0_512'


def test_bench_reference_arm_ignores_torchrun_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; round 1's reference arm then ran on one core and every N >= 2
    ratio of the scaling record was inflated by the core count.  The arm sizes itself from the affinity mask instead."""
    env = dict(os.environ, OMP_NUM_THREADS='1', RANK='0', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--workload', 'cfg1_scene0_512', '--gpus', '2'], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().split('\n')[-1])
    assert d['cpu_baseline']['cores'] == len(os.sched_getaffinity(0)) and 'warning' not in d
    # the other ranks print nothing and exit 0
    env['RANK'] = '1'
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--gpus', '2'],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_bench_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_bench_clock_sampler_nvml_path_and_fallback(monkeypatch):
    """bench.py samples SM clock / power / throttle reasons during the timed region through NVML (a fake module here);
    only samples inside [mark_begin, mark_end] count, reason bits map to the contract's names, and when NVML is
    unusable the sampler falls back (nvidia-smi; absent in this container: an empty but well-formed result)."""
    import time
    import types
    sys.path.insert(0, ROOT)
    import bench

    state = {'clock': 1965, 'mask': 0}
    fake = types.ModuleType('pynvml')
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: ('handle', i)
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: state['clock']
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: state['mask']
    fake.nvmlDeviceGetPowerUsage = lambda h: 312800
    monkeypatch.setitem(sys.modules, 'pynvml', fake)
    s = bench.ClockSampler(0)
    s.start()
    assert s.source == 'nvml'
    state['clock'], state['mask'] = 1200, 0x8            # before the window: must not be reported
    time.sleep(0.05)
    state['clock'], state['mask'] = 1965, 0x4            # sw_power_cap inside the window
    time.sleep(0.03)
    s.mark_begin()
    time.sleep(0.12)
    s.mark_end()
    state['clock'], state['mask'] = 1000, 0x40
    time.sleep(0.03)
    out = s.stop()
    assert out['source'] == 'nvml' and out['samples_in_timed_region'] >= 5 and out['samples'] == out['samples_in_timed_region']
    assert out['sm_mhz'] == 1965.0 and out['sm_max_mhz'] == 1965.0 and out['reasons'] == ['sw_power_cap']
    assert abs(out['power_w_max'] - 312.8) < 1e-6

    def broken():
        raise RuntimeError('no driver')
    fake.nvmlInit = broken
    s = bench.ClockSampler(0)
    s.start()
    s.mark_begin(); s.mark_end()
    out = s.stop()
    assert out['source'] in (None, 'nvidia-smi') and 'sm_mhz' in out and 'reasons' in out and 'samples' in out
