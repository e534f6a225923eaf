"""ctypes front-end of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE (see oracle/oracle.cpp)."""
import ctypes as C
import os
import subprocess
import numpy as np

from . import pack, sdf_build

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build():
    subprocess.run(['make', '-s', '-C', HERE], check=True)


def lib(count=False, bvh=False):
    name = 'liboracle_bvh.so' if bvh else ('liboracle_count.so' if count else 'liboracle.so')
    if name not in _libs:
        path = os.path.join(HERE, '_build', name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_pcg32.restype = C.c_uint32
        L.oracle_pcg32.argtypes = [C.c_uint32]
        L.oracle_generate_seed.restype = C.c_uint32
        L.oracle_random_float.restype = C.c_float
        L.oracle_bk7.restype = C.c_float
        L.oracle_bk7.argtypes = [C.c_float]
        L.oracle_intersect.restype = C.c_float
        L.oracle_counter_names.restype = C.c_char_p
        L.oracle_sample_wavelengths.argtypes = [C.c_float, C.c_void_p]
        L.oracle_wave_to_xyz.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        L.oracle_emit.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.oracle_spd.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.oracle_sdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_sdf_eval4.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_dispatch_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_dispatch_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_load_sdf.argtypes = [C.c_char_p]
        L.oracle_set_surface_ext.argtypes = [C.c_void_p, C.c_int]
        if bvh:
            L.oracle_bvh_build.argtypes = [C.c_void_p]
        _libs[name] = L
    return _libs[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One scene bound to the CPU oracle. `ubo` float32[4097], `sdf_sources` list of GLSL strings."""

    def __init__(self, ubo, sdf_sources=(), count=False, threads=0, bvh=False, surface_ext=None):
        """bvh=True: liboracle_bvh.so, the checker that routes the closest-hit search through the PRODUCT's BVH
        (pt_bvh.cpp) with the oracle's primitives at the leaves; must render what the plain oracle renders."""
        self.L = lib(count, bvh)
        self.bvh = bvh
        self.count = count
        self.ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        assert self.ubo.size == pack.UBO_FLOATS
        self.sdf_so = sdf_build.build(list(sdf_sources))
        self.threads = threads
        # surface extensions (pt_surface_ext table; NOT reference behaviour, parity unpinned): None = the reference's shading
        self.surface_ext = np.ascontiguousarray(surface_ext if surface_ext is not None else [], dtype=pack.SURFACE_EXT_DTYPE)

    def _bind(self):
        if self.L.oracle_set_surface_ext(_p(self.surface_ext) if self.surface_ext.size else None, int(self.surface_ext.size)) != 0:
            raise RuntimeError('oracle: bad surface extension table')
        if self.L.oracle_load_sdf(self.sdf_so.encode()) != 0:
            raise RuntimeError('oracle: cannot load SDF dispatchers %s' % self.sdf_so)
        self.L.oracle_set_threads(int(self.threads))
        if self.bvh and self.L.oracle_bvh_build(_p(self.ubo)) <= 0:
            raise RuntimeError('oracle: cannot build the BVH for this scene')

    def dispatch(self, params, image, row_start=0, row_step=1):
        """One vkCmdDispatch: image (H,W,4) float32 is read-modify-written. Returns counters dict if count.
        row_start/row_step restrict it to a strided subset of rows (bounded CPU-baseline samples)."""
        self._bind()
        assert image.dtype == np.float32 and image.flags.c_contiguous
        params = np.ascontiguousarray(params)
        n = self.L.oracle_num_counters()
        cnt = np.zeros(n, dtype=np.uint64)
        rc = self.L.oracle_dispatch_rows(_p(self.ubo), _p(params), _p(image), _p(cnt), row_start, row_step)
        if rc != 0:
            raise RuntimeError('oracle_dispatch failed: %d' % rc)
        if self.count:
            return dict(zip(self.L.oracle_counter_names().decode().split(','), (int(v) for v in cnt)))
        return None

    def render(self, params, total_samples, spf):
        """Offscreen MainLoop bookkeeping (host:4042-4048): returns the (H,W,4) running-mean image."""
        p = np.array(params, copy=True)
        W, H = int(p['resolution'][0]), int(p['resolution'][1])
        img = np.zeros((H, W, 4), dtype=np.float32)
        for j in range(1, total_samples // spf + 1):
            p['frame'] = j * spf
            p['currentSamples'] = j * spf
            p['samplesPerFrame'] = spf
            self.dispatch(p, img)
        return img

    def samples(self, params, gx, gy, first, n):
        self._bind()
        params = np.ascontiguousarray(params)
        out = np.zeros((n, 3), dtype=np.float32)
        self.L.oracle_samples(_p(self.ubo), _p(params), gx, gy, first, n, _p(out))
        return out

    def cost_map(self, params, n, y0, y1):
        """(y1-y0, W, n, 3) uint32: SDF evaluations, rays and shaded bounces of every Scene() call (needs count=True)."""
        assert self.count
        self._bind()
        params = np.ascontiguousarray(params)
        W = int(np.ravel(params['resolution'])[0])
        out = np.zeros((y1 - y0, W, n, 3), dtype=np.uint32)
        self.L.oracle_cost_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        self.L.oracle_cost_map(_p(self.ubo), _p(params), n, _p(out), y0, y1)
        return out

    def dispatch_sum(self, params, first, n, image):
        self._bind()
        params = np.ascontiguousarray(params)
        self.L.oracle_dispatch_sum(_p(self.ubo), _p(params), first, n, _p(image))

    def sdf_eval(self, xyz, set1=1):
        """SDF() / SDFMATERIAL() at the points; set1 is the first mask word, or a sequence of up to four words."""
        self._bind()
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = xyz.shape[0]
        d = np.zeros(n, dtype=np.float32)
        m = np.zeros(n, dtype=np.float32)
        words = np.zeros(4, dtype=np.uint32)
        w = np.atleast_1d(np.asarray(set1, dtype=np.uint64))
        words[:len(w)] = w
        self.L.oracle_sdf_eval4(_p(self.ubo), _p(xyz), n, _p(words), _p(d), _p(m))
        return d, m

    def intersect(self, origin, direction):
        self._bind()
        o = np.asarray(origin, dtype=np.float32)
        d = np.asarray(direction, dtype=np.float32)
        out = np.zeros(5, dtype=np.float32)
        t = self.L.oracle_intersect(_p(self.ubo), _p(o), _p(d), _p(out))
        return float(t), out


def math_eval(fn, x, y=None):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), dtype=np.float32)
    out = np.zeros_like(x)
    L.oracle_math_eval(fn, _p(x), _p(y), _p(out), x.size)
    return out


def from_scene_file(path, count=False, threads=0):
    scene = pack.load_scene(path)
    return Oracle(pack.pack_ubo(scene), pack.sdf_sources(scene), count=count, threads=threads, surface_ext=pack.surface_ext(scene)), scene
