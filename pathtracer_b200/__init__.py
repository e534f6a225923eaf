"""pathtracer_b200 -- Python bindings (ctypes) of libpt_cuda, the B200-native drop-in for PathTracer's compute kernel.

Nothing in here computes: the hot path is hand-written sm_100a CUDA behind the C ABI of include/pt_abi.h.
There is no CPU fallback; importing works anywhere, creating a Renderer needs a CUDA device and the built library.
"""
from .api import (kernel_compile_check, LibraryNotBuilt, MultiRenderer, PtError, Renderer, Scene, build_library, lib, library_path, MODE_FAST, MODE_STRICT,
                  PARAMS_DTYPE, PIPE_MEGAKERNEL, PIPE_WAVEFRONT, UBO_FLOATS, SURFACE_EXT_DTYPE, BSDF_REFERENCE, BSDF_MIRROR,
                  BSDF_GLOSSY, BSDF_DIELECTRIC)

__all__ = ['LibraryNotBuilt', 'MultiRenderer', 'PtError', 'Renderer', 'Scene', 'build_library', 'lib', 'library_path', 'MODE_FAST',
           'MODE_STRICT', 'PARAMS_DTYPE', 'PIPE_MEGAKERNEL', 'PIPE_WAVEFRONT', 'UBO_FLOATS', 'SURFACE_EXT_DTYPE', 'BSDF_REFERENCE',
           'BSDF_MIRROR', 'BSDF_GLOSSY', 'BSDF_DIELECTRIC', 'kernel_compile_check']
