"""The drop-in boundary: libpt_cuda.so loads without a GPU and exports every symbol include/pt_abi.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pt_abi.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pt_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(ptlib):
    lib = ctypes.CDLL(ptlib.library_path())
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


LAYOUT_TU = r'''
#include <stddef.h>
#include "pt_abi.h"
/* the uniform block: 4097 tightly packed floats in the reference's order (shader.comp:19-27 == host:187-195) */
static_assert(sizeof(pt_ubo) == 16388, "uniform block");
static_assert(offsetof(pt_ubo, numObjects) == 0 && offsetof(pt_ubo, objects) == 4 * 7, "ubo head");
static_assert(offsetof(pt_ubo, sdfs) == 4 * 1031 && offsetof(pt_ubo, materials) == 4 * 1799, "ubo sdfs / materials");
static_assert(offsetof(pt_ubo, lights) == 4 * 2582 && offsetof(pt_ubo, lightIDs) == 4 * 2710, "ubo lights");
static_assert(offsetof(pt_ubo, CIEXYZ1931) == 4 * 2774, "ubo CIE table");
/* the push-constant block, byte offsets as SURVEY.md section 8a lists them (shader.comp:31-52 == host:197-218) */
static_assert(sizeof(pt_params) == 88, "push constants");
static_assert(offsetof(pt_params, resolution) == 0 && offsetof(pt_params, frame) == 8 && offsetof(pt_params, currentSamples) == 12, "");
static_assert(offsetof(pt_params, samplesPerFrame) == 16 && offsetof(pt_params, FPS) == 20 && offsetof(pt_params, persistence) == 24, "");
static_assert(offsetof(pt_params, pathLength) == 28 && offsetof(pt_params, cameraAngle) == 32 && offsetof(pt_params, cameraPosX) == 40, "");
static_assert(offsetof(pt_params, cameraPosY) == 44 && offsetof(pt_params, cameraPosZ) == 48 && offsetof(pt_params, ISO) == 52, "");
static_assert(offsetof(pt_params, cameraSize) == 56 && offsetof(pt_params, apertureSize) == 60 && offsetof(pt_params, apertureDist) == 64, "");
static_assert(offsetof(pt_params, lensRadius) == 68 && offsetof(pt_params, lensFocalLength) == 72 && offsetof(pt_params, lensThickness) == 76, "");
static_assert(offsetof(pt_params, lensDistance) == 80 && offsetof(pt_params, tonemap) == 84, "");
int main(void) { return 0; }
'''


def test_struct_layout_through_a_compiler(tmp_path):
    """sizeof / offsetof of the two boundary blocks, checked by gcc (C11) and g++ (C++17) on a real translation unit."""
    import subprocess
    src = tmp_path / 'layout.c'
    src.write_text(LAYOUT_TU.replace('static_assert', '_Static_assert'))
    subprocess.run(['gcc', '-std=c11', '-fsyntax-only', '-I' + os.path.join(ROOT, 'include'), str(src)], check=True)
    srcpp = tmp_path / 'layout.cpp'
    srcpp.write_text(LAYOUT_TU)
    subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-I' + os.path.join(ROOT, 'include'), str(srcpp)], check=True)
    assert 7 + 1024 + 768 + 783 + 128 + 64 + 1323 == 4097


def test_integration_md_binding_compiles(tmp_path):
    """The reference-side binding shown in INTEGRATION.md, through g++: its static_asserts against the reference's own
    struct definitions (taken from /root/reference when present, else from stand-ins with the same members), and every
    pt_* call of the snippet against the header's prototypes."""
    import subprocess
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'```cpp\n(.*?)```', text, flags=re.S)
    assert blocks, 'INTEGRATION.md has no ```cpp block'
    ref_structs = None
    ref_host = os.path.join(ROOT, 'oracle', '_ref', 'ref_host.cpp')
    if os.path.exists(ref_host):
        t = open(ref_host).read()
        a, b = t.index('#define MAX_OBJECTS_SIZE'), t.index('const float CIEXYZ1931[1323]')
        ref_structs = '#include <glm/glm.hpp>\n#include <string>\n' + t[a:b]
    if ref_structs is None or not os.path.exists('/root/reference/includes/glm/glm.hpp'):
        ref_structs = '''
struct UniformBufferObject { float numObjects[7]; float packedObjects[1024]; float packedSdfs[768]; float packedMaterials[783];
                             float packedLights[128]; float packedLightIDs[64]; float CIEXYZ1931[1323]; };
struct PushConstantValues { int resolution[2]; int frame, currentSamples, samplesPerFrame; float FPS, persistence; int pathLength;
                            float cameraAngle[2]; float cameraPosX, cameraPosY, cameraPosZ; int ISO; float cameraSize, apertureSize,
                            apertureDist, lensRadius, lensFocalLength, lensThickness, lensDistance; int tonemap; };
'''
    # the members of the reference's App the snippet touches (host:1088-1089, 1129, 1146, 1157-1191), as a base class
    base = '''
struct sdf_stub { std::string glsl; };
struct AppBase {
    int W = 1280, H = 720, tonemap = 3;
    UniformBufferObject ubo;
    PushConstantValues pushConstant;
    std::vector<sdf_stub> sdfs;
    std::string renderDir;
    void UpdateUniformBuffer() {}
    void UpdatePushConstant() {}
};
'''
    body = '\n'.join(blocks)
    assert 'class App {' in body
    head, tail = body.split('class App {', 1)
    tu = ('#include <stdexcept>\n#include <string>\n#include <vector>\n#include <cstring>\n' + ref_structs + head + base +
          'class App : public AppBase {' + tail + '\nint main() { return 0; }\n')
    src = tmp_path / 'binding.cpp'
    src.write_text(tu)
    r = subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-I' + os.path.join(ROOT, 'include'), '-I/root/reference/includes', str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_no_cpu_fallback(ptlib):
    """Without a CUDA device the context refuses to exist (this test only asserts that on a GPU-less box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(ptlib.PtError) as e:
        ptlib.Renderer()
    assert e.value.code == -5


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under pathtracer_b200/ or include/ may import, include or link it."""
    bad = []
    for base in ('pathtracer_b200', 'include', 'examples'):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if 'lib' in dp.split(os.sep):
                continue
            for f in fn:
                if f.endswith(('.py', '.cpp', '.cu', '.cuh', '.h', 'Makefile')):
                    t = open(os.path.join(dp, f), errors='replace').read()
                    if re.search(r'(import\s+oracle|from\s+oracle|oracle/|liboracle)', t) and 'TEST' not in f:
                        # comments that merely mention the oracle as the parity gate are fine
                        for line in t.split('\n'):
                            if re.search(r'(import\s+oracle|from\s+oracle|#include\s*"[^"]*oracle|liboracle)', line):
                                bad.append((f, line))
    assert not bad, bad


def test_pfm_and_ppm_writers(ptlib, tmp_path):
    """pt_write_pfm / pt_read_pfm round trip (bottom-up rows) and pt_write_ppm = SaveRender + SavePPM (host:3491-3518,
    918-933) with the display transform of shader.frag:31-93, checked against a numpy restatement."""
    import ctypes as C
    import numpy as np
    L = ptlib.lib()
    rng = np.random.default_rng(4)
    img = np.ones((5, 7, 4), dtype=np.float32)
    img[..., :3] = rng.random((5, 7, 3)).astype(np.float32) * 2
    pfm = str(tmp_path / 'a.pfm').encode()
    assert L.pt_write_pfm(pfm, img.ctypes.data_as(C.c_void_p), 7, 5, 0) == 0
    raw = open(pfm, 'rb').read()
    assert raw.startswith(b'PF\n7 5\n-1.0\n')
    body = np.frombuffer(raw[len(b'PF\n7 5\n-1.0\n'):], dtype='<f4').reshape(5, 7, 3)
    assert np.array_equal(body[::-1], img[..., :3])          # PFM stores the bottom scanline first
    back = np.zeros_like(img)
    assert L.pt_read_pfm(pfm, back.ctypes.data_as(C.c_void_p), 7, 5) == 0 and np.array_equal(back, img)
    ppm = str(tmp_path / 'a.ppm').encode()
    for tm in (0, 1, 2, 3):
        assert L.pt_write_ppm(ppm, img.ctypes.data_as(C.c_void_p), 7, 5, tm) == 0
        raw = open(ppm, 'rb').read()
        assert raw.startswith(b'P6\n7\n5\n255\n')
        got = np.frombuffer(raw[len(b'P6\n7\n5\n255\n'):], dtype=np.uint8).reshape(5, 7, 3).astype(int)
        xyz = img[..., :3].astype(np.float64)
        e2d = np.array([[0.9531874, -0.0265906, 0.0238731], [-0.0382467, 1.0288406, 0.0094060], [0.0026068, -0.0030332, 1.0892565]])
        x2r = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])
        rgb = np.maximum(xyz @ e2d.T @ x2r.T, 0)
        with np.errstate(divide='ignore'):
            rgb = [rgb, rgb / (1 + rgb), rgb * (2.51 * rgb + 0.03) / (rgb * (2.43 * rgb + 0.59) + 0.14), np.exp(-0.25 / rgb)][tm]
        rgb = np.clip(rgb, 0, 1)
        srgb = np.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * rgb ** (1 / 2.4) - 0.055)
        assert np.abs(got - (srgb * 255).astype(int)).max() <= 1


def test_exr_writer(ptlib, tmp_path):
    """pt_write_exr: parsed back by a 30-line reader here (header attributes, offset table, scanline blocks) and, for
    the R,G,B flavour, by OpenCV's OpenEXR codec."""
    import ctypes as C
    import struct
    import numpy as np
    L = ptlib.lib()
    rng = np.random.default_rng(9)
    W, H = 9, 6
    img = np.ones((H, W, 4), dtype=np.float32)
    img[..., :3] = rng.random((H, W, 3)).astype(np.float32) * 3

    def parse(path):
        b = open(path, 'rb').read()
        assert struct.unpack_from('<II', b, 0) == (20000630, 2)
        pos, attrs = 8, {}
        while b[pos] != 0:
            e = b.index(b'\0', pos); name = b[pos:e].decode(); pos = e + 1
            e = b.index(b'\0', pos); typ = b[pos:e].decode(); pos = e + 1
            n = struct.unpack_from('<I', b, pos)[0]; pos += 4
            attrs[name] = (typ, b[pos:pos + n]); pos += n
        pos += 1
        assert attrs['compression'][1] == b'\0' and attrs['lineOrder'][1] == b'\0'
        assert struct.unpack('<4i', attrs['dataWindow'][1]) == (0, 0, W - 1, H - 1)
        names = [attrs['channels'][1][i * 18:i * 18 + 1].decode() for i in range(3)]
        offs = struct.unpack_from('<%dQ' % H, b, pos)
        out = np.zeros((H, W, 3), dtype=np.float32)
        for y in range(H):
            yy, nbytes = struct.unpack_from('<ii', b, offs[y])
            assert yy == y and nbytes == W * 12
            rows = np.frombuffer(b, dtype='<f4', count=W * 3, offset=offs[y] + 8).reshape(3, W)
            out[y] = rows.T
        return names, out

    p1 = str(tmp_path / 'xyz.exr')
    assert L.pt_write_exr(p1.encode(), img.ctypes.data_as(C.c_void_p), W, H, 0) == 0
    names, data = parse(p1)
    assert names == ['X', 'Y', 'Z'] and np.array_equal(data, img[..., :3])
    p2 = str(tmp_path / 'rgb.exr')
    assert L.pt_write_exr(p2.encode(), img.ctypes.data_as(C.c_void_p), W, H, 1) == 0
    names, data = parse(p2)
    assert names == ['B', 'G', 'R']
    e2d = np.array([[0.9531874, -0.0265906, 0.0238731], [-0.0382467, 1.0288406, 0.0094060], [0.0026068, -0.0030332, 1.0892565]])
    x2r = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])
    rgb = img[..., :3].astype(np.float64) @ e2d.T @ x2r.T
    assert np.allclose(data[..., ::-1], rgb, rtol=1e-5, atol=1e-6)
    os.environ['OPENCV_IO_ENABLE_OPENEXR'] = '1'
    try:
        import cv2
        cv = cv2.imread(p2, cv2.IMREAD_UNCHANGED)
    except Exception:
        cv = None
    if cv is not None:
        assert cv.shape == (H, W, 3) and np.allclose(cv[..., ::-1], rgb, rtol=1e-5, atol=1e-6)
