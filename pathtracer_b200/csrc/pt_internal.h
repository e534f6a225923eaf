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
int pt_prepare_params(const pt_params* p, int accum_mode, int first_sample, int n_samples, PtDevParams* d,
                      std::string* err);

/* pt_bvh.cpp: BVH over the bounded primitives (layout: pt_bvh.h); blob = nodes + copy of the record pool */
int pt_bvh_bounded_prims(const PtDevScene* sc);
int pt_bvh_build(const PtDevScene* sc, std::vector<float>* blob, std::string* err);

/* pt_sdf_front.cpp: GLSL snippets -> CUDA translation unit text (prelude + snippets + dispatchers) */
int pt_sdf_generate(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, std::string* out, std::string* err);

/* pt_jit.cpp: NVRTC.  Returns a cubin for sm_100a in *cubin. */
struct PtJitOptions {
    int mode;            /* pt_mode */
    bool bake_counts;    /* compile primitive counts in as constants */
    bool wavefront;      /* also build the wavefront pipeline's kernels (pt_wavefront.cuh) */
    bool bvh;            /* closest hit through the BVH of pt_bvh.h instead of the brute-force scan */
    int counts[6];       /* spheres, planes, boxes, lenses, cyclides, sdfs */
};
int pt_jit_compile(const std::string& sdf_unit, const PtJitOptions& opt, std::vector<char>* cubin, std::string* log);

/* pt_kernels_{strict,fast}.cu: statically compiled generic kernels (no SDF) and helpers */
extern "C" {
void pt_launch_strict(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image, void* stream);
void pt_launch_fast(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image, void* stream);
void pt_launch_finalize(void* image, int n_texels, float invTotal, float exposure, void* stream);
void pt_launch_math_eval(int fn, const float* x, const float* y, float* out, size_t n, void* stream);
const void* pt_static_kernel_strict(void);
const void* pt_static_kernel_fast(void);
}

/* pt_embedded.cpp (generated): the headers NVRTC needs, as strings */
struct PtEmbeddedHeader { const char* name; const char* text; };
extern const PtEmbeddedHeader pt_embedded_headers[];
extern const int pt_embedded_header_count;

#endif
