"""ctypes front-end of pathtracer_b200/lib/libpt_cuda.so (C ABI: include/pt_abi.h).

Class and method names mirror the reference's host functions for this path (src/pathtracer.cpp):
  Scene.load            ReadJSON + UpdateFromJSON            host:893-908, 2576-2722
  Scene.pack_ubo        UpdateUniformBuffer                  host:3642-3811
  Scene.pack_params     UpdatePushConstant                   host:3813-3834
  Renderer.set_scene    UpdateUniformBuffer + RecompileComputeShaders (InsertSDF)   host:2004-2054, 3836-3841
  Renderer.resize       CreateTexelBuffer                    host:2250-2269
  Renderer.dispatch     one vkCmdDispatch of DrawFrame       host:3586-3608, 3843-3880
  Renderer.render       offscreen MainLoop                   host:4005-4086
  Renderer.read_xyz     mapped texel memory                  host:3491-3518
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_STRICT, MODE_FAST = 0, 1
PIPE_MEGAKERNEL, PIPE_WAVEFRONT = 0, 1
UBO_FLOATS = 4097

PARAMS_DTYPE = np.dtype([
    ('resolution', '<i4', (2,)), ('frame', '<i4'), ('currentSamples', '<i4'), ('samplesPerFrame', '<i4'),
    ('FPS', '<f4'), ('persistence', '<f4'), ('pathLength', '<i4'), ('cameraAngle', '<f4', (2,)),
    ('cameraPosX', '<f4'), ('cameraPosY', '<f4'), ('cameraPosZ', '<f4'), ('ISO', '<i4'), ('cameraSize', '<f4'),
    ('apertureSize', '<f4'), ('apertureDist', '<f4'), ('lensRadius', '<f4'), ('lensFocalLength', '<f4'),
    ('lensThickness', '<f4'), ('lensDistance', '<f4'), ('tonemap', '<i4')])
assert PARAMS_DTYPE.itemsize == 88
# pt_surface_ext (pt_abi.h): the surface extensions of SURVEY 8f-4 -- not reference behaviour, off by default
SURFACE_EXT_DTYPE = np.dtype([('bsdf', '<i4'), ('roughness', '<f4'), ('ior', '<f4'), ('pad', '<f4')])
BSDF_REFERENCE, BSDF_MIRROR, BSDF_GLOSSY, BSDF_DIELECTRIC = 0, 1, 2, 3
MAX_SURFACE_EXT = 64


class PtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libpt_cuda error %d: %s' % (code, message))
        self.code = code


class LibraryNotBuilt(RuntimeError):
    pass


def library_path():
    return os.path.join(HERE, 'lib', 'libpt_cuda.so')


def build_library(verbose=False):
    """make -C pathtracer_b200/csrc (nvcc -gencode arch=compute_100a,code=sm_100a; cross-compiles without a GPU)."""
    r = subprocess.run(['make', '-C', os.path.join(HERE, 'csrc'), '-j8'], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError('building libpt_cuda failed:\n%s\n%s' % (r.stdout or '', r.stderr or ''))


_lib = None


def lib():
    """The loaded library. Raises LibraryNotBuilt (never falls back to anything) when the .so is missing."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise LibraryNotBuilt('%s not found: run `python -c "import __graft_entry__ as g; g.build()"` or '
                                  '`make -C pathtracer_b200/csrc`. There is no CPU fallback.' % path)
        L = C.CDLL(path)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.pt_version.restype = C.c_char_p
        L.pt_last_error.restype = C.c_char_p
        L.pt_last_error.argtypes = [vp]
        L.pt_create.argtypes = [ci, C.POINTER(vp)]
        L.pt_destroy.argtypes = [vp]
        L.pt_destroy.restype = None
        L.pt_set_mode.argtypes = [vp, ci]
        L.pt_set_jit.argtypes = [vp, ci]
        L.pt_set_pipeline.argtypes = [vp, ci]
        L.pt_multi_create.argtypes = [C.POINTER(ci), ci, ci, C.POINTER(vp)]
        L.pt_multi_destroy.argtypes = [vp]
        L.pt_multi_destroy.restype = None
        L.pt_multi_last_error.argtypes = [vp]
        L.pt_multi_last_error.restype = C.c_char_p
        L.pt_multi_num_devices.argtypes = [vp]
        L.pt_multi_ctx.argtypes = [vp, ci]
        L.pt_multi_ctx.restype = vp
        L.pt_multi_set_scene.argtypes = [vp, vp, C.POINTER(C.c_char_p), ci]
        L.pt_multi_resize.argtypes = [vp, ci, ci]
        L.pt_multi_render.argtypes = [vp, vp, ci, ci, ci, C.POINTER(C.c_double)]
        L.pt_multi_reduce_seconds.argtypes = [vp]
        L.pt_multi_reduce_seconds.restype = C.c_double
        L.pt_multi_read_xyz.argtypes = [vp, vp, C.c_size_t]
        L.pt_set_bvh.argtypes = [vp, ci]
        L.pt_bvh_active.argtypes = [vp]
        L.pt_set_scene.argtypes = [vp, vp, C.POINTER(C.c_char_p), ci]
        L.pt_set_surface_ext.argtypes = [vp, vp, ci]
        L.pt_multi_set_surface_ext.argtypes = [vp, vp, ci]
        L.pt_scene_surface_ext.argtypes = [vp, vp, ci]
        L.pt_resize.argtypes = [vp, ci, ci]
        L.pt_bind_image.argtypes = [vp, vp, ci, ci]
        L.pt_clear.argtypes = [vp]
        L.pt_dispatch.argtypes = [vp, vp]
        L.pt_dispatch_sum.argtypes = [vp, vp, ci, ci]
        L.pt_finalize.argtypes = [vp, vp, ci]
        L.pt_render.argtypes = [vp, vp, ci, ci]
        L.pt_read_xyz.argtypes = [vp, vp, C.c_size_t]
        L.pt_write_xyz.argtypes = [vp, vp, C.c_size_t]
        L.pt_read_xyz_async.argtypes = [vp, vp, C.c_size_t]
        L.pt_read_wait.argtypes = [vp]
        L.pt_render_resume.argtypes = [vp, vp, ci, ci, ci]
        L.pt_read_pfm.argtypes = [C.c_char_p, vp, ci, ci]
        L.pt_sync.argtypes = [vp]
        L.pt_image_ptr.argtypes = [vp]
        L.pt_image_ptr.restype = vp
        L.pt_stream_handle.argtypes = [vp]
        L.pt_stream_handle.restype = vp
        L.pt_kernel_time.argtypes = [vp, C.POINTER(cf), C.POINTER(C.c_longlong)]
        L.pt_scene_load_json.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.pt_scene_parse_json.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.pt_scene_free.argtypes = [vp]
        L.pt_scene_free.restype = None
        L.pt_scene_num_shots.argtypes = [vp]
        L.pt_scene_num_sdf.argtypes = [vp]
        L.pt_scene_sdf_glsl.argtypes = [vp, ci]
        L.pt_scene_sdf_glsl.restype = C.c_char_p
        L.pt_scene_pack_ubo.argtypes = [vp, vp]
        L.pt_scene_to_json.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.pt_scene_to_json.restype = C.c_long
        L.pt_scene_save_json.argtypes = [vp, C.c_char_p]
        L.pt_scene_pack_params.argtypes = [vp, ci, ci, ci, ci, ci, vp]
        L.pt_write_ppm.argtypes = [C.c_char_p, vp, ci, ci, ci]
        L.pt_write_pfm.argtypes = [C.c_char_p, vp, ci, ci, ci]
        L.pt_write_exr.argtypes = [C.c_char_p, vp, ci, ci, ci]
        L.pt_cie1931_table.restype = C.POINTER(cf)
        L.pt_sdf_translate.argtypes = [C.POINTER(C.c_char_p), ci, vp, C.c_char_p, C.c_size_t]
        L.pt_sdf_translate.restype = C.c_long
        L.pt_sdf_compile_check.argtypes = [C.POINTER(C.c_char_p), ci, vp, ci]
        L.pt_math_eval.argtypes = [vp, ci, vp, vp, vp, C.c_size_t]
        L.pt_sdf_eval.argtypes = [vp, vp, C.c_size_t, C.c_uint32, vp, vp]
        L.pt_sdf_eval4.argtypes = [vp, vp, C.c_size_t, vp, vp, vp]
        L.pt_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
        L.pt_get_option.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong)]
        L.pt_kernel_compile_check.argtypes = [vp, C.POINTER(C.c_char_p), ci, ci, ci]
        L.pt_kernel_compile_check_opts.argtypes = [vp, C.POINTER(C.c_char_p), ci, ci, ci, C.c_char_p]
        L.pt_fp32_peak.argtypes = [vp, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pt_debug_stats.argtypes = [vp, vp, ci]
        _lib = L
    return _lib


def _check(rc, ctx=None):
    if rc != 0:
        raise PtError(rc, (lib().pt_last_error(ctx) or b'').decode(errors='replace'))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _c_strings(strings):
    arr = (C.c_char_p * max(len(strings), 1))()
    for i, s in enumerate(strings):
        arr[i] = s if isinstance(s, bytes) else s.encode()
    return arr


def cie1931_table():
    return np.ctypeslib.as_array(lib().pt_cie1931_table(), shape=(1323,)).copy()


def sdf_translate(sources, sdfs_raw=None):
    """The CUDA/C++ translation unit pt_set_scene hands to NVRTC for these snippets (pt_sdf_front.cpp)."""
    L = lib()
    raw = np.zeros(6 * max(len(sources), 1), dtype=np.float32) if sdfs_raw is None else np.ascontiguousarray(sdfs_raw, dtype=np.float32)
    arr = _c_strings(sources)
    n = L.pt_sdf_translate(arr, len(sources), _ptr(raw), None, 0)
    if n < 0:
        _check(int(n))
    buf = C.create_string_buffer(int(n))
    L.pt_sdf_translate(arr, len(sources), _ptr(raw), buf, int(n))
    return buf.value.decode()


def sdf_compile_check(sources, sdfs_raw=None, mode=MODE_STRICT):
    """NVRTC compile-only check of the whole kernel with these snippets; needs no GPU."""
    raw = np.zeros(6 * max(len(sources), 1), dtype=np.float32) if sdfs_raw is None else np.ascontiguousarray(sdfs_raw, dtype=np.float32)
    _check(lib().pt_sdf_compile_check(_c_strings(sources), len(sources), _ptr(raw), mode))


def kernel_compile_check(ubo, sources, mode=MODE_STRICT, bake_counts=True, options=None, wavefront=False, bvh=False,
                         surface_ext=False):
    """NVRTC compile-only check of the kernel pt_set_scene would build (needs no GPU).  `options` is a dict of
    pt_set_option keys.  Returns the ptxas report (registers, spills, shared memory)."""
    ubo = np.ascontiguousarray(ubo, dtype=np.float32)
    opts = ','.join('%s=%d' % (k, int(v)) for k, v in (options or {}).items()).encode()
    m = int(mode) | (2 if wavefront else 0) | (4 if bvh else 0) | (8 if surface_ext else 0)
    _check(lib().pt_kernel_compile_check_opts(_ptr(ubo), _c_strings(list(sources)), len(sources), m, int(bool(bake_counts)), opts))
    return (lib().pt_last_error(None) or b'').decode(errors='replace')


class Scene:
    """A parsed scene file (same schema as the reference's scenes/*.json)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        _check(lib().pt_scene_load_json(os.fspath(path).encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def parse(cls, text):
        h = C.c_void_p()
        _check(lib().pt_scene_parse_json(text.encode() if isinstance(text, str) else text, C.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.pt_scene_free(self._h)
            self._h = None

    @property
    def num_shots(self):
        return lib().pt_scene_num_shots(self._h)

    @property
    def sdf_sources(self):
        L = lib()
        return [L.pt_scene_sdf_glsl(self._h, i) for i in range(L.pt_scene_num_sdf(self._h))]

    def to_json(self):
        """UpdateToJSON (host:2724-2858): the scene as JSON text, reals rounded to 1e-5."""
        n = lib().pt_scene_to_json(self._h, None, 0)
        if n < 0:
            _check(int(n))
        buf = C.create_string_buffer(int(n))
        lib().pt_scene_to_json(self._h, buf, int(n))
        return buf.value.decode()

    def save(self, path):
        _check(lib().pt_scene_save_json(self._h, os.fspath(path).encode()))

    def pack_ubo(self):
        ubo = np.zeros(UBO_FLOATS, dtype=np.float32)
        _check(lib().pt_scene_pack_ubo(self._h, _ptr(ubo)))
        return ubo

    def surface_ext(self):
        """The materials' optional "bsdf" / "roughness" / "ior" keys as a pt_surface_ext table (empty: none set)."""
        t = np.zeros(MAX_SURFACE_EXT, dtype=SURFACE_EXT_DTYPE)
        n = lib().pt_scene_surface_ext(self._h, _ptr(t), MAX_SURFACE_EXT)
        _check(min(n, 0))
        return t[:n].copy()

    def pack_params(self, shot=1, width=1280, height=720, spf=1, path_length=5):
        p = np.zeros((), dtype=PARAMS_DTYPE)
        _check(lib().pt_scene_pack_params(self._h, shot, width, height, spf, path_length, _ptr(p)))
        return p


class Renderer:
    """One device context (pt_ctx). Fails loudly without a GPU: there is no CPU fallback."""

    def __init__(self, device=0, mode=MODE_STRICT, jit=None, pipeline=PIPE_MEGAKERNEL, options=None):
        self._ctx = C.c_void_p()
        self.width = self.height = 0
        _check(lib().pt_create(device, C.byref(self._ctx)))
        _check(lib().pt_set_mode(self._ctx, mode), self._ctx)
        if jit is not None:
            _check(lib().pt_set_jit(self._ctx, jit), self._ctx)
        if pipeline != PIPE_MEGAKERNEL:
            _check(lib().pt_set_pipeline(self._ctx, pipeline), self._ctx)
        for k, v in (options or {}).items():
            self.set_option(k, v)
        self._keep = None

    def set_option(self, key, value):
        """pt_set_option: a tuning option of the run-time compiled kernels ("sched", "steal_s", ...); next set_scene."""
        _check(lib().pt_set_option(self._ctx, key.encode(), int(value)), self._ctx)

    def get_option(self, key):
        v = C.c_longlong()
        _check(lib().pt_get_option(self._ctx, key.encode(), C.byref(v)), self._ctx)
        return int(v.value)

    def fp32_peak(self, repeats=5):
        """Measured FP32 FMA peak of the device in TFLOP/s (pt_fp32_peak)."""
        t, ms = C.c_double(), C.c_double()
        _check(lib().pt_fp32_peak(self._ctx, repeats, C.byref(t), C.byref(ms)), self._ctx)
        return float(t.value), float(ms.value)

    def debug_stats(self, reset=True):
        out = (C.c_ulonglong * 16)()
        _check(lib().pt_debug_stats(self._ctx, out, 1 if reset else 0), self._ctx)
        return [int(v) for v in out]

    def close(self):
        if getattr(self, '_ctx', None) and _lib is not None:
            _lib.pt_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def set_mode(self, mode):
        _check(lib().pt_set_mode(self._ctx, mode), self._ctx)

    def set_jit(self, policy):
        _check(lib().pt_set_jit(self._ctx, policy), self._ctx)

    def set_pipeline(self, pipeline):
        """PIPE_MEGAKERNEL (default) or PIPE_WAVEFRONT; takes effect at the next set_scene."""
        _check(lib().pt_set_pipeline(self._ctx, pipeline), self._ctx)

    def set_bvh(self, min_prims):
        """Bounded-primitive count from which the closest-hit search walks the BVH (<= 0: never); next set_scene."""
        _check(lib().pt_set_bvh(self._ctx, min_prims), self._ctx)

    @property
    def bvh_active(self):
        return bool(lib().pt_bvh_active(self._ctx))

    def set_surface_ext(self, table=None):
        """pt_set_surface_ext: table = SURFACE_EXT_DTYPE array (entry i extends material i), None / empty = reference
        shading.  Takes effect at the next set_scene."""
        t = np.ascontiguousarray(table if table is not None else [], dtype=SURFACE_EXT_DTYPE)
        _check(lib().pt_set_surface_ext(self._ctx, _ptr(t) if t.size else None, int(t.size)), self._ctx)

    def set_scene(self, ubo, sdf_sources=()):
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        assert ubo.size == UBO_FLOATS
        _check(lib().pt_set_scene(self._ctx, _ptr(ubo), _c_strings(list(sdf_sources)), len(sdf_sources)), self._ctx)

    def resize(self, width, height):
        _check(lib().pt_resize(self._ctx, width, height), self._ctx)
        self.width, self.height = width, height
        self._keep = None

    def bind_image(self, tensor):
        """Render into caller-owned device memory: a torch CUDA tensor of shape (H, W, 4), float32, contiguous."""
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.dim() == 3 and tensor.shape[2] == 4
        h, w = int(tensor.shape[0]), int(tensor.shape[1])
        _check(lib().pt_bind_image(self._ctx, C.c_void_p(tensor.data_ptr()), w, h), self._ctx)
        self.width, self.height = w, h
        self._keep = tensor

    def clear(self):
        _check(lib().pt_clear(self._ctx), self._ctx)

    def dispatch(self, params):
        _check(lib().pt_dispatch(self._ctx, _ptr(np.ascontiguousarray(params))), self._ctx)

    def dispatch_sum(self, params, first_sample, n_samples):
        _check(lib().pt_dispatch_sum(self._ctx, _ptr(np.ascontiguousarray(params)), first_sample, n_samples), self._ctx)

    def finalize(self, params, total_samples):
        _check(lib().pt_finalize(self._ctx, _ptr(np.ascontiguousarray(params)), total_samples), self._ctx)

    def render(self, params, total_samples, spf):
        _check(lib().pt_render(self._ctx, _ptr(np.ascontiguousarray(params)), total_samples, spf), self._ctx)

    def sync(self):
        _check(lib().pt_sync(self._ctx), self._ctx)

    def read_xyz(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        _check(lib().pt_read_xyz(self._ctx, _ptr(out), out.size), self._ctx)
        return out

    def read_xyz_async(self, out):
        """Snapshot the image and copy it to `out` (a numpy view of pinned memory) while later dispatches run."""
        _check(lib().pt_read_xyz_async(self._ctx, _ptr(out), out.size), self._ctx)

    def read_wait(self):
        _check(lib().pt_read_wait(self._ctx), self._ctx)

    def write_xyz(self, image):
        """Upload a saved accumulation image (checkpoint resume)."""
        image = np.ascontiguousarray(image, dtype=np.float32)
        _check(lib().pt_write_xyz(self._ctx, _ptr(image), image.size), self._ctx)

    def render_resume(self, params, done_samples, total_samples, spf):
        _check(lib().pt_render_resume(self._ctx, _ptr(np.ascontiguousarray(params)), done_samples, total_samples, spf), self._ctx)

    def kernel_time(self):
        """(milliseconds of device time, number of kernel launches) since the previous call; CUDA events."""
        ms, n = C.c_float(), C.c_longlong()
        _check(lib().pt_kernel_time(self._ctx, C.byref(ms), C.byref(n)), self._ctx)
        return ms.value, n.value

    @property
    def stream(self):
        return lib().pt_stream_handle(self._ctx)

    def math_eval(self, fn, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float32)
        out = np.empty_like(x)
        _check(lib().pt_math_eval(self._ctx, fn, _ptr(x), None if y is None else _ptr(y), _ptr(out), x.size), self._ctx)
        return out

    def sdf_eval(self, xyz, set1=1):
        """SDF() / SDFMATERIAL() of the current scene at the points; set1 = the first mask word or up to four words."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = xyz.shape[0]
        d, m = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
        words = np.zeros(4, dtype=np.uint32)
        w = np.atleast_1d(np.asarray(set1, dtype=np.uint64))
        words[:len(w)] = w
        _check(lib().pt_sdf_eval4(self._ctx, _ptr(xyz), n, _ptr(words), _ptr(d), _ptr(m)), self._ctx)
        return d, m


class MultiRenderer:
    """pt_multi: the GPUs of one box from a single host thread -- sample-split, one NCCL reduce, finalize on the first
    device (include/pt_abi.h).  The one-process twin of bench.py's one-rank-per-GPU path."""

    def __init__(self, devices, mode=MODE_STRICT, jit=None):
        self._m = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = lib().pt_multi_create(arr, len(devices), mode, C.byref(self._m))
        if rc != 0:
            raise PtError(rc, (lib().pt_multi_last_error(None) or b'').decode())
        self.width = self.height = 0
        if jit is not None:
            for i in range(len(devices)):
                _check(lib().pt_set_jit(lib().pt_multi_ctx(self._m, i), jit))

    def _ok(self, rc):
        if rc != 0:
            raise PtError(rc, (lib().pt_multi_last_error(self._m) or b'').decode())

    def close(self):
        if getattr(self, '_m', None) and _lib is not None:
            _lib.pt_multi_destroy(self._m)
            self._m = None

    __del__ = close

    def set_surface_ext(self, table=None):
        t = np.ascontiguousarray(table if table is not None else [], dtype=SURFACE_EXT_DTYPE)
        self._ok(lib().pt_multi_set_surface_ext(self._m, _ptr(t) if t.size else None, int(t.size)))

    def set_scene(self, ubo, sdf_sources=()):
        ubo = np.ascontiguousarray(ubo, dtype=np.float32)
        assert ubo.size == UBO_FLOATS
        self._ok(lib().pt_multi_set_scene(self._m, _ptr(ubo), _c_strings(list(sdf_sources)), len(sdf_sources)))

    def resize(self, width, height):
        self._ok(lib().pt_multi_resize(self._m, width, height))
        self.width, self.height = width, height

    def render(self, params, first_sample, total_samples, spf):
        """Returns (seconds of dispatches + reduce + finalize, seconds of the reduce alone)."""
        params = np.ascontiguousarray(params)
        secs = C.c_double(0.0)
        self._ok(lib().pt_multi_render(self._m, _ptr(params), first_sample, total_samples, spf, C.byref(secs)))
        return secs.value, lib().pt_multi_reduce_seconds(self._m)

    def read_xyz(self):
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._ok(lib().pt_multi_read_xyz(self._m, _ptr(out), out.size))
        return out
