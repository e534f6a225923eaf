/* pt_internal.h -- internal C++ interfaces of libpt_cuda (not part of the ABI) */
#ifndef PT_INTERNAL_H
#define PT_INTERNAL_H

#include <string>
#include <vector>

#include "pt_abi.h"
#include "pt_math.h"
#include "pt_dev_scene.h"

/* pt_prepare.cpp */
int pt_prepare_scene(const pt_ubo* ubo, PtDevScene* sc, std::string* err);
bool pt_prepare_surface_ext(const pt_surface_ext* table, int n, PtDevScene* sc);
int pt_prepare_params(const pt_params* p, int accum_mode, int first_sample, int n_samples, PtDevParams* d,
                      std::string* err);

/* pt_bvh.cpp: BVH over the bounded primitives (layout: pt_bvh.h); blob = nodes + copy of the record pool */
int pt_bvh_bounded_prims(const PtDevScene* sc);
int pt_bvh_build(const PtDevScene* sc, std::vector<float>* blob, std::string* err);

/* pt_sdf_front.cpp: GLSL snippets -> CUDA translation unit text (prelude + snippets + dispatchers) */
int pt_sdf_generate(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, std::string* out, std::string* err);

/* pt_jit.cpp: NVRTC.  Returns a cubin for sm_100a in *cubin. */
/* Tuning options of the run-time compiled kernels (pt_set_option, pt_abi.h).  -1 = "auto": pt_jit_compile picks the
 * measured default for the scene.  They are part of the translation unit, hence of every cache key. */
struct PtKnobs {
    int sched = -1;       /* driver: 0 v1, 5 v2s, 7 v3s, 8 v2m */
    int sdf_reps = 8;     /* SDF() evaluations per execution of the SDF phase (8 vs 16: +1..2 % on cfg3/4a/5, profiles/r02_gpu3-4) */
    int feed_t = 8;       /* v2s / v2m: the SDF phase waits until every feeder phase has fewer lanes than this */
    int regen_t = 16;     /* v3s: waiting lanes that trigger a regeneration */
    int steal_s = -1;     /* samples per pixel per round of the pool (0 = the whole dispatch, fast mode only) */
    int min_blocks = -1;  /* __launch_bounds__ minimum CTAs per SM */
    int no_unroll = -1;   /* 1: keep the primitive loops rolled although the counts are baked */
    int pool_cap = 32;    /* v2m: slots of the per-warp pool of parked paths */
    int pool_min = 24;    /* v2m: marching rays from which the SDF phase runs ahead of the feeders */
    int stats = 0;        /* 1: build with the scheduling counters (pt_debug_stats) */
    int wf_refill = 8;    /* wavefront march kernel: evaluations between refills */
    int bvh_while_while = 0;
    int heavy_min = -1;   /* v2s: lanes that must wait before the box / lens / cyclide tests run as a phase of their own (0: inline) */
    int sin_poly_every = 0; /* fast mode: every k-th sin( of the SDF snippets is evaluated on the FMA pipe (0: none) */
    int pregen = -1;      /* v2s / v3s: camera rays from a generation kernel's records instead of inline (-1: auto) */
    int pathcolor_unroll = -1; /* SDF builds: PathColor's four CIE look-ups unrolled (-1: auto = with pregen) */
    int resolve = -1;     /* with pregen: radiance -> XYZ and the per-pixel sum in a resolve kernel (-1: auto = with pregen) */
};
int pt_knob_set(PtKnobs* k, const char* key, long long value);       /* 0, or -1 for an unknown key / bad value */
int pt_knob_get(const PtKnobs* k, const char* key, long long* value);
int pt_knobs_parse(PtKnobs* k, const char* text, std::string* err);  /* "key=value,key=value" */
struct PtJitOptions {
    int mode;            /* pt_mode */
    bool bake_counts;    /* compile primitive counts in as constants */
    bool wavefront;      /* also build the wavefront pipeline's kernels (pt_wavefront.cuh) */
    bool bvh;            /* closest hit through the BVH of pt_bvh.h instead of the brute-force scan */
    int counts[6];       /* spheres, planes, boxes, lenses, cyclides, sdfs */
    bool surface_ext = false; /* the scene carries surface extensions (pt_set_surface_ext): build the kernel with them */
    PtKnobs knobs;
};
/* the complete translation unit pt_jit_compile would hand to NVRTC (also the key of every kernel cache) */
std::string pt_jit_source(const std::string& sdf_unit, const PtJitOptions& opt);
bool pt_jit_uses_pregen(const std::string& source); /* whether that translation unit reads its camera rays from records */
bool pt_jit_uses_resolve(const std::string& source); /* ... and leaves XYZ projection and summation to the resolve kernel */
int pt_jit_compile(const std::string& sdf_unit, const PtJitOptions& opt, std::vector<char>* cubin, std::string* log);

/* pt_kernels_{strict,fast}.cu: statically compiled generic kernels (no SDF) and helpers */
extern "C" {
void pt_launch_strict(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image, void* stream);
void pt_launch_fast(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image, void* stream);
void pt_launch_finalize(void* image, int n_texels, float invTotal, float exposure, void* stream);
void pt_launch_fma_peak(float* out, int blocks, int threads, int iters, void* stream);
void pt_launch_math_eval(int fn, const float* x, const float* y, float* out, size_t n, void* stream);
const void* pt_static_kernel_strict(void);
const void* pt_static_kernel_fast(void);
}

/* pt_embedded.cpp (generated): the headers NVRTC needs, as strings */
struct PtEmbeddedHeader { const char* name; const char* text; };
extern const PtEmbeddedHeader pt_embedded_headers[];
extern const int pt_embedded_header_count;

#endif
