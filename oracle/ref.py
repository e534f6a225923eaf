"""ctypes front-end of oracle/_ref (the reference's own sources compiled for the CPU, see oracle/ref_build.py)
-- TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import ctypes as C
import json
import os
import numpy as np

from . import ref_build

_host = None
_shaders = {}


def available():
    """True when oracle/_ref can be used: prebuilt objects present, or the reference tree is there to build from."""
    return os.path.exists(ref_build.host_so()) or ref_build.have_reference()


def host():
    global _host
    if _host is None:
        L = C.CDLL(ref_build.build_host())
        L.ref_last_error.restype = C.c_char_p
        L.ref_app_new.restype = C.c_void_p
        L.ref_app_free.argtypes = [C.c_void_p]
        L.ref_load_scene.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_load_scene_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_get_ubo.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_num_sdfs.argtypes = [C.c_void_p]
        L.ref_num_shots.argtypes = [C.c_void_p]
        L.ref_sdf_glsl.argtypes = [C.c_void_p, C.c_int]
        L.ref_sdf_glsl.restype = C.c_char_p
        L.ref_get_push.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   C.c_int, C.c_int, C.c_void_p]
        L.ref_to_json.restype = C.c_long
        L.ref_to_json.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.ref_save_render_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_save_ppm.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_round_decimal.restype = C.c_double
        L.ref_round_decimal.argtypes = [C.c_double, C.c_double]
        for n in ('ref_host_spd', 'ref_host_blackbody', 'ref_host_blackbody_peak'):
            getattr(L, n).restype = C.c_float
        L.ref_host_spd.argtypes = [C.c_float] * 4
        L.ref_host_blackbody.argtypes = [C.c_float] * 2
        L.ref_host_blackbody_peak.argtypes = [C.c_float]
        L.ref_cie_table.restype = C.POINTER(C.c_float)
        _host = L
    return _host


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """A scene file loaded by the reference's own UpdateFromJSON / UpdateUniformBuffer (host:2576-2722, 3642-3811)."""

    def __init__(self, path, shot=1):
        self.L = host()
        self.path = path
        self.app = self.L.ref_app_new()
        if self.L.ref_load_scene(self.app, path.encode(), shot) != 0:
            raise RuntimeError('reference loader: ' + self.L.ref_last_error().decode())
        self._shader = None

    def close(self):
        if self.app:
            self.L.ref_app_free(self.app)
            self.app = None

    def __del__(self):
        self.close()

    def ubo(self):
        out = np.zeros(4097, dtype=np.float32)
        self.L.ref_get_ubo(self.app, _p(out))
        return out

    def sdf_sources(self):
        return [self.L.ref_sdf_glsl(self.app, i).decode() for i in range(self.L.ref_num_sdfs(self.app))]

    def num_shots(self):
        return self.L.ref_num_shots(self.app)

    def push(self, width, height, frame, current_samples, spf, path_length=5, frame_time=0.0166, persistence=0.0625,
             tonemap=3):
        """UpdatePushConstant (host:3813-3834) -> the 88-byte block as uint8[88]."""
        out = np.zeros(88, dtype=np.uint8)
        self.L.ref_get_push(self.app, width, height, frame, current_samples, spf, frame_time, persistence,
                            path_length, tonemap, _p(out))
        return out

    def to_json(self):
        cap = 1 << 22
        buf = C.create_string_buffer(cap)
        n = self.L.ref_to_json(self.app, buf, cap)
        if n < 0 or n >= cap:
            raise RuntimeError('reference UpdateToJSON: ' + self.L.ref_last_error().decode())
        return buf.value.decode()

    def shader(self):
        if self._shader is None:
            self._shader = RefShader(self.path)
        return self._shader

    def dispatch(self, push, image, row_start=0, row_step=1, row_end=None, threads=0):
        """One vkCmdDispatch of the reference shader on `image` (H, W, 4) float32, read-modify-write."""
        self.shader().dispatch(self.ubo(), push, image, row_start, row_step, row_end, threads)

    def render(self, width, height, spp, spf=1, path_length=5, first_sample=0, threads=0, row_start=0, row_step=1):
        """The offscreen MainLoop's bookkeeping (host:4042-4048) around `spp/spf` dispatches on a fresh image."""
        img = np.zeros((height, width, 4), dtype=np.float32)
        ubo = self.ubo()
        n = (spp + spf - 1) // spf
        for j in range(1, n + 1):
            push = self.push(width, height, first_sample + j * spf, j * spf, spf, path_length)
            self.shader().dispatch(ubo, push, img, row_start, row_step, None, threads)
        return img


class RefShader:
    """oracle/_ref/ref_shader_<tag>.so for one scene file: src/shader.comp with that scene's SDF snippets."""

    def __init__(self, scene_path):
        key = os.path.abspath(scene_path)
        if key not in _shaders:
            L = C.CDLL(ref_build.build_shader(scene_path))
            L.ref_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            L.ref_pcg32.restype = C.c_uint32
            L.ref_pcg32.argtypes = [C.c_uint32]
            L.ref_random_float.restype = C.c_float
            L.ref_random_float.argtypes = [C.c_void_p]
            L.ref_generate_seed.restype = C.c_uint32
            L.ref_generate_seed.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int]
            L.ref_wave_to_xyz.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
            L.ref_sample_wavelengths.argtypes = [C.c_float, C.c_void_p]
            L.ref_bk7.restype = C.c_float
            L.ref_bk7.argtypes = [C.c_float]
            L.ref_rotation_matrix.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_emit.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
            L.ref_spd.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
            L.ref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_sdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
            L.ref_lens_ray.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
            L.ref_pcg32_n.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.ref_generate_seed_n.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_sample.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
            L.ref_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            L.ref_solve_quartic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_cone_pdf.restype = C.c_float
            L.ref_cone_pdf.argtypes = [C.c_float, C.c_float]
            L.ref_mis_weight.restype = C.c_float
            L.ref_mis_weight.argtypes = [C.c_float, C.c_float]
            L.ref_orthonormal_basis.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_trace_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
            L.ref_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            _shaders[key] = L
        self.L = _shaders[key]

    def dispatch(self, ubo, push, image, row_start=0, row_step=1, row_end=None, threads=0):
        assert image.dtype == np.float32 and image.flags.c_contiguous and image.shape[2] == 4
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        push = np.ascontiguousarray(push)
        assert ubo.size == 4097 and push.nbytes == 88
        h = image.shape[0] if row_end is None else row_end
        self.L.ref_dispatch(_p(ubo), _p(push), _p(image), row_start, h, row_step, threads)

    def pcg32(self, seed):
        return int(self.L.ref_pcg32(seed & 0xFFFFFFFF))

    def pcg32_n(self, seeds):
        a = np.ascontiguousarray(seeds, dtype=np.uint32)
        out = np.zeros_like(a)
        self.L.ref_pcg32_n(_p(a), a.size, _p(out))
        return out

    def generate_seed_n(self, push, gxyk):
        """gxyk (n, 3) int32 of (gid.x, gid.y, k) -> uint32 seeds of Scene()."""
        a = np.ascontiguousarray(gxyk, dtype=np.int32)
        push = np.ascontiguousarray(push)
        out = np.zeros(len(a), dtype=np.uint32)
        self.L.ref_generate_seed_n(_p(push), _p(a), len(a), _p(out))
        return out

    def sample(self, kind, seed, n, param=0.0, normal=None):
        out = np.zeros((n, 3), dtype=np.float32)
        nn = None if normal is None else np.asarray(normal, dtype=np.float32)
        self.L.ref_sample(kind, seed & 0xFFFFFFFF, n, param, None if nn is None else _p(nn), _p(out))
        return out

    def visible(self, ubo, push, origin, direction, light_object):
        od = np.concatenate([np.asarray(origin, np.float32), np.asarray(direction, np.float32)])
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        push = np.ascontiguousarray(push)
        return bool(self.L.ref_visible(_p(ubo), _p(push), _p(od), int(light_object)))

    def solve_quartic(self, coef5):
        c = np.asarray(coef5, dtype=np.float32)
        r = np.zeros(4, dtype=np.float32)
        b = np.zeros(4, dtype=np.int32)
        self.L.ref_solve_quartic(_p(c), _p(r), _p(b))
        return r, b.astype(bool)

    def cone_pdf(self, c, cmax):
        return float(self.L.ref_cone_pdf(c, cmax))

    def mis_weight(self, a, b):
        return float(self.L.ref_mis_weight(a, b))

    def orthonormal_basis(self, n):
        a = np.asarray(n, dtype=np.float32)
        out = np.zeros(6, dtype=np.float32)
        self.L.ref_orthonormal_basis(_p(a), _p(out))
        return out[:3].copy(), out[3:].copy()

    def trace_path(self, ubo, push, origin, direction, l4, seed, n):
        od = np.concatenate([np.asarray(origin, np.float32), np.asarray(direction, np.float32)])
        l = np.asarray(l4, dtype=np.float32)
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        push = np.ascontiguousarray(push)
        out = np.zeros(4, dtype=np.float64)
        self.L.ref_trace_path(_p(ubo), _p(push), _p(od), _p(l), seed & 0xFFFFFFFF, n, _p(out))
        return out

    def accumulate(self, push, in3, out3):
        a = np.asarray(in3, dtype=np.float32)
        o = np.array(out3, dtype=np.float32)
        push = np.ascontiguousarray(push)
        self.L.ref_accumulate(_p(push), _p(a), _p(o))
        return o

    def random_float(self, seed):
        s = C.c_uint32(seed & 0xFFFFFFFF)
        f = self.L.ref_random_float(C.byref(s))
        return float(f), int(s.value)

    def generate_seed(self, push, x, y, k):
        push = np.ascontiguousarray(push)
        return int(self.L.ref_generate_seed(_p(push), x, y, k))

    def wave_to_xyz(self, ubo, wave):
        out = np.zeros(3, dtype=np.float32)
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        self.L.ref_wave_to_xyz(_p(ubo), wave, _p(out))
        return out

    def sample_wavelengths(self, lh):
        out = np.zeros(4, dtype=np.float32)
        self.L.ref_sample_wavelengths(lh, _p(out))
        return out

    def bk7(self, l):
        return float(self.L.ref_bk7(l))

    def rotation_matrix(self, angle):
        a = np.asarray(angle, dtype=np.float32)
        out = np.zeros(9, dtype=np.float32)
        self.L.ref_rotation_matrix(_p(a), _p(out))
        return out.reshape(3, 3)  # [column][row], as glm / GLSL store it

    def emit(self, l4, temperature, luminosity):
        a = np.asarray(l4, dtype=np.float32)
        out = np.zeros(4, dtype=np.float32)
        self.L.ref_emit(_p(a), temperature, luminosity, _p(out))
        return out

    def spd(self, l4, peak, sigma, invert):
        a = np.asarray(l4, dtype=np.float32)
        out = np.zeros(4, dtype=np.float32)
        self.L.ref_spd(_p(a), peak, sigma, int(invert), _p(out))
        return out

    def intersect(self, ubo, push, rays):
        """rays (n, 6) origin+dir -> (n, 6) {t, normal, materialID, lightID} of the shader's Intersection()."""
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        push = np.ascontiguousarray(push)
        out = np.zeros((len(rays), 6), dtype=np.float32)
        self.L.ref_intersect(_p(ubo), _p(push), _p(rays), len(rays), _p(out))
        return out

    def sdf_eval(self, ubo, points, set1=1):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        d = np.zeros(len(pts), dtype=np.float32)
        m = np.zeros(len(pts), dtype=np.float32)
        self.L.ref_sdf_eval(_p(ubo), _p(pts), len(pts), set1, _p(d), _p(m))
        return d, m

    def lens_ray(self, ubo, push, l, origin, direction, forward):
        od = np.concatenate([np.asarray(origin, np.float32), np.asarray(direction, np.float32)])
        fw = np.asarray(forward, dtype=np.float32)
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        push = np.ascontiguousarray(push)
        self.L.ref_lens_ray(_p(ubo), _p(push), l, _p(od), _p(fw))
        return od[:3].copy(), od[3:].copy()


def save_render_pixels(xyz_rgba, tonemap=3):
    """SaveRender's per-pixel loop (host:3501-3512) on an (H, W, 4) float32 image -> (H, W, 3) uint8."""
    L = host()
    img = np.ascontiguousarray(xyz_rgba, dtype=np.float32)
    h, w = img.shape[:2]
    out = np.zeros((h, w, 3), dtype=np.uint8)
    app = L.ref_app_new()
    L.ref_save_render_pixels(app, _p(img), w, h, tonemap, _p(out))
    L.ref_app_free(app)
    return out


def cie_table():
    return np.ctypeslib.as_array(host().ref_cie_table(), shape=(1323,)).copy()
