/* pt_lib.cpp -- the C ABI of libpt_cuda (include/pt_abi.h): device context, scene upload, dispatch, read-back.
 *
 * Replaces the Vulkan side of the reference's dispatch path: CreateUniformBuffer / CreateTexelBuffer /
 * UpdateUniformBuffer / UpdatePushConstant / RecordComputeCommandBuffer / vkQueueSubmit
 * (host:2230-2269, 3586-3608, 3642-3880).  One CUDA stream per context; every entry point returns a pt_status
 * and leaves the message in pt_last_error().  There is no CPU fallback: without a CUDA device pt_create fails.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "pt_internal.h"
#include "pt_bvh.h"

const std::string& pt_scene_last_error();

namespace {

thread_local std::string g_last_error;

struct JitKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    cudaKernel_t sdf_eval = nullptr; /* only for scenes with SDF snippets */
    cudaKernel_t gen = nullptr;      /* option "pregen": the camera-ray generation kernel (pt_gen_body) */
    cudaKernel_t resolve = nullptr;  /* option "resolve": radiance -> XYZ and the per-pixel sums (pt_resolve_body) */
    /* wavefront pipeline (pt_wavefront.cuh); wf_march only with SDF snippets */
    cudaKernel_t wf_sort = nullptr; /* surface extensions: SHADE's path rays grouped by lobe (option wf_sort) */
    cudaKernel_t wf_gen = nullptr, wf_isect = nullptr, wf_march = nullptr, wf_shade = nullptr, wf_final = nullptr,
                 wf_ctl = nullptr;
};

}  // namespace

struct pt_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int mode = PT_MODE_STRICT;
    int jit_policy = 1;      /* 0: never (static kernels only), 1: when the scene has SDFs, 2: always (baked counts) */
    int pipeline = PT_PIPE_MEGAKERNEL;
    int bvh_min = PT_BVH_DEFAULT_MIN_PRIMS; /* bounded primitives from which the BVH replaces the scan; <= 0: never */
    PtKnobs knobs;           /* pt_set_option: tuning options of the run-time compiled kernels */
    std::vector<pt_surface_ext> surface_ext; /* pt_set_surface_ext: applied by the next pt_set_scene */
    long long pregen_max_mb = 4096;  /* option "pregen": record buffer of one band; a dispatch that needs more runs as bands of
                                        CTA rows over two such buffers */
    float4* d_gen = nullptr;
    size_t gen_bytes = 0;
    bool gen_clamped = false; /* the buffer is smaller than asked for because memory was short */
    cudaStream_t gen_stream = nullptr; /* second stream of a banded pregen dispatch */
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    long long wf_sort = 0;      /* wavefront pipeline with surface extensions: sort SHADE's path rays by lobe (measured: slower) */
    long long wf_max_paths = 0; /* wavefront pipeline: paths in flight per chunk; 0 = auto (4 Mi without SDFs, 32 Mi with:
                                   profiles/r02_wf_l2) */
    bool bvh_active = false;
    PtWf wf;                 /* wavefront buffers (lazily allocated) */
    void* wf_block = nullptr;
    size_t wf_paths = 0, wf_pixels = 0;
    bool scene_set = false;
    pt_ubo ubo;
    PtDevScene dev_scene;
    float* d_ubo = nullptr;   /* flat copy of the uniform block (+ the BVH blob at PT_BVH_UBO_OFF) */
    /* pinned staging ring for the scene upload: pt_set_scene never synchronises the stream, so a host loop of
     * set_scene / dispatch / read_xyz_async keeps the GPU busy back to back */
    float* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    unsigned stage_next = 0;
    float* d_image = nullptr; /* accumulation image (RGBA32F) */
    bool own_image = false;
    int width = 0, height = 0;
    JitKernel* active_jit = nullptr; /* null: statically compiled kernel */
    std::map<std::string, JitKernel> jit_cache;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    /* asynchronous read-back (pt_read_xyz_async): snapshot buffer, copy stream, ordering events */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    float* d_snap = nullptr;
    size_t snap_bytes = 0;
    bool copy_pending = false;
    bool timing_open = false;
    long long launches = 0;
    std::string error;
};

namespace {

int fail(pt_ctx* ctx, int code, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->error = msg;
    return code;
}
int cuda_fail(pt_ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, PT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define PT_CUDA(ctx, call)                                  \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); \
    } while (0)

int launch_wavefront(pt_ctx* ctx, const PtDevParams& dp);

int launch(pt_ctx* ctx, const PtDevParams& dp) {
    if (ctx->pipeline == PT_PIPE_WAVEFRONT) return launch_wavefront(ctx, dp);
    if (!ctx->timing_open) {
        PT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        ctx->timing_open = true;
    }
    if (ctx->active_jit && ctx->active_jit->gen) {
        /* option "pregen": the generation kernel writes one 32-byte record per sample, the render kernel reads them; both
         * run on the same grid.  A dispatch whose records outgrow pregen_max_mb runs as bands of CTA rows, alternating
         * between two record buffers and two streams: band k + 1 is generated while band k renders and its render
         * kernel fills the SMs that band k's last wave leaves idle (bands write disjoint texels).  Band k + 2 reuses
         * band k's buffer in band k's stream, so stream order is all the synchronisation the buffers need; the second
         * stream is forked from and joined to the context's stream with events around the dispatch. */
        const unsigned gx = (unsigned)((dp.width + 15) / 16), gy = (unsigned)((dp.height + 7) / 8);
        const size_t per_row = (size_t)gx * 4u * 32u * (size_t)dp.samplesPerFrame; /* records of one CTA row */
        const size_t rec_bytes = ctx->active_jit->resolve ? 48u : 32u;             /* + the radiance bundle of option "resolve" */
        size_t rows = ((size_t)ctx->pregen_max_mb << 20) / (per_row * rec_bytes);
        if (rows < 1) rows = 1;
        if (rows > gy) rows = gy;
        const size_t nbuf_want = rows < gy ? 2 : 1;
        const size_t have_rows = ctx->gen_bytes / (per_row * rec_bytes * nbuf_want);
        if (have_rows < rows && !(ctx->gen_clamped && have_rows >= 1)) { /* grow (rarely): only now ask how much memory is free */
            const size_t want_rows = rows;
            size_t free_b = 0, total_b = 0;
            if (ctx->d_gen) {
                PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (ctx->gen_stream) PT_CUDA(ctx, cudaStreamSynchronize(ctx->gen_stream));
                cudaFree(ctx->d_gen);
                ctx->d_gen = nullptr;
                ctx->gen_bytes = 0;
            }
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && nbuf_want * rows * per_row * rec_bytes > free_b / 2)
                rows = (free_b / 2) / (per_row * rec_bytes * 2u);
            if (rows < 1) rows = 1;
            const size_t nbuf = rows < gy ? 2 : 1;
            if (cudaMalloc((void**)&ctx->d_gen, nbuf * rows * per_row * rec_bytes) != cudaSuccess) {
                (void)cudaGetLastError();
                ctx->d_gen = nullptr;
                return fail(ctx, PT_ERR_CUDA, "no memory for the camera-ray records of one row of tiles (" +
                                                  std::to_string((nbuf * rows * per_row * rec_bytes) >> 20) +
                                                  " MiB): dispatch fewer samples at a time or pt_set_option(ctx, \"pregen\", 0)");
            }
            ctx->gen_bytes = nbuf * rows * per_row * rec_bytes;
            ctx->gen_clamped = rows < want_rows; /* do not try again at every dispatch */
        }
        {   /* what the buffer we have allows */
            const size_t nbuf = rows < gy ? 2 : 1;
            if (rows * nbuf > ctx->gen_bytes / (per_row * rec_bytes)) rows = ctx->gen_bytes / (per_row * rec_bytes * 2u);
            if (rows < 1) rows = 1; /* cannot happen: the buffer holds at least one row per stream */
        }
        const bool two = rows < gy;
        if (two && !ctx->gen_stream) {
            PT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->gen_stream, cudaStreamNonBlocking));
            PT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            PT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        if (two) { /* fork: everything queued on the context's stream so far (scene upload, earlier dispatches, clears) comes first */
            PT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
            PT_CUDA(ctx, cudaStreamWaitEvent(ctx->gen_stream, ctx->ev_fork, 0));
        }
        const float* ubo = ctx->d_ubo;
        float* image = ctx->d_image;
        const dim3 block(128, 1, 1);
        unsigned k = 0;
        for (unsigned y0 = 0; y0 < gy; y0 += (unsigned)rows, k++) {
            const unsigned ny = (gy - y0 < (unsigned)rows) ? gy - y0 : (unsigned)rows;
            cudaStream_t st = (k & 1u) ? ctx->gen_stream : ctx->stream;
            PtDevParams band = dp;
            float4* buf = ctx->d_gen + (size_t)(k & 1u) * rows * per_row * (rec_bytes / 16u); /* the odd bands' buffer follows the even ones' */
            band.gen = buf;                         /* two planes of genCount float4 */
            band.rad = buf + 2u * (size_t)ny * per_row; /* one more for option "resolve" */
            band.genCount = (unsigned long long)ny * per_row;
            band.blockY0 = (int)y0;
            const dim3 grid(gx, ny, 1);
            float4* gen = buf;
            void* gargs[2] = {(void*)&band, (void*)&gen};
            PT_CUDA(ctx, cudaLaunchKernel((const void*)ctx->active_jit->gen, grid, block, gargs, 0, st));
            void* args[4] = {(void*)&ctx->dev_scene, (void*)&band, (void*)&ubo, (void*)&image};
            PT_CUDA(ctx, cudaLaunchKernel((const void*)ctx->active_jit->kernel, grid, block, args, 0, st));
            ctx->launches += 2;
            if (ctx->active_jit->resolve) {
                void* rargs[3] = {(void*)&band, (void*)&ubo, (void*)&image};
                PT_CUDA(ctx, cudaLaunchKernel((const void*)ctx->active_jit->resolve, grid, block, rargs, 0, st));
                ctx->launches++;
            }
        }
        if (two) { /* join */
            PT_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->gen_stream));
            PT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        }
        ctx->launches--; /* the render kernel of the last band is counted below */
    } else if (ctx->active_jit) {
        const dim3 grid((unsigned)((dp.width + 15) / 16), (unsigned)((dp.height + 7) / 8), 1), block(128, 1, 1);
        const float* ubo = ctx->d_ubo;
        float* image = ctx->d_image;
        void* args[4] = {(void*)&ctx->dev_scene, (void*)&dp, (void*)&ubo, (void*)&image};
        PT_CUDA(ctx, cudaLaunchKernel((const void*)ctx->active_jit->kernel, grid, block, args, 0, ctx->stream));
    } else if (ctx->mode == PT_MODE_FAST) {
        pt_launch_fast(&ctx->dev_scene, &dp, ctx->d_ubo, ctx->d_image, ctx->stream);
    } else {
        pt_launch_strict(&ctx->dev_scene, &dp, ctx->d_ubo, ctx->d_image, ctx->stream);
    }
    PT_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return PT_OK;
}

/* (Re)allocate the SoA path state for `paths` paths over `pixels` pixels: one block, carved into arrays */
int wf_alloc(pt_ctx* ctx, size_t paths, size_t pixels) {
    if (ctx->wf_block && ctx->wf_paths >= paths && ctx->wf_pixels >= pixels) return PT_OK;
    if (ctx->wf_block) { cudaFree(ctx->wf_block); ctx->wf_block = nullptr; }
    const size_t f4 = paths * 16, total = 10 * f4 + pixels * 16 + paths * 8 + 4 * paths * 4 + 256;
    PT_CUDA(ctx, cudaMalloc(&ctx->wf_block, total));
    char* b = (char*)ctx->wf_block;
    PtWf& w = ctx->wf;
    void** f4s[10] = {&w.rayO, &w.rayD, &w.wl, &w.rad, &w.thr, &w.shD, &w.shC, &w.hit0, &w.hit1, &w.col};
    for (int i = 0; i < 10; i++) { *f4s[i] = b; b += f4; }
    w.acc = b; b += pixels * 16;
    w.misc = b; b += paths * 8;
    void** qs[4] = {&w.qA, &w.qB, &w.qS, &w.qM};
    for (int i = 0; i < 4; i++) { *qs[i] = b; b += paths * 4; }
    w.cnt = b;
    ctx->wf_paths = paths;
    ctx->wf_pixels = pixels;
    return PT_OK;
}

/* One dispatch through the wavefront pipeline: chunks of samples x (GEN, per depth {ISECT, MARCH, SHADE} for the path
 * rays and again for the shadow rays, FINAL).  Queue sizes stay on the device; every kernel is launched with a fixed
 * persistent grid and reads its count there, so the host never synchronises inside a dispatch. */
int launch_wavefront(pt_ctx* ctx, const PtDevParams& dp0) {
    JitKernel* k = ctx->active_jit;
    if (!k || !k->wf_gen) return fail(ctx, PT_ERR_ARG, "wavefront pipeline: kernels not built (call pt_set_scene after pt_set_pipeline)");
    const size_t pixels = (size_t)dp0.width * (size_t)dp0.height;
    /* A chunk = a band of consecutive pixels x some samples, at most wf_max_paths paths: the frame is tiled so that the
     * ~190 B of state per path of one chunk can stay in the 126 MB L2 between the kernels of a bounce. */
    const size_t max_paths = ctx->wf_max_paths > 0 ? (size_t)ctx->wf_max_paths
                                                   : (ctx->dev_scene.nSdfs > 0 ? (size_t)32 << 20 : (size_t)4 << 20);
    const size_t band = pixels < max_paths ? pixels : max_paths; /* pixels per chunk */
    int chunk = (int)(max_paths / band);                          /* samples per chunk */
    if (chunk < 1) chunk = 1;
    if (chunk > dp0.samplesPerFrame) chunk = dp0.samplesPerFrame;
    int rc = wf_alloc(ctx, band * (size_t)chunk, band);
    if (rc != PT_OK) return rc;
    if (!ctx->timing_open) {
        PT_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        ctx->timing_open = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const dim3 block(128, 1, 1), grid((unsigned)(sms * 8), 1, 1), one(1, 1, 1);
    const float* ubo = ctx->d_ubo;
    float* image = ctx->d_image;
    PtDevParams dp = dp0;
    PtWf w = ctx->wf;
    const bool has_sdf = ctx->dev_scene.nSdfs > 0;
    auto run = [&](cudaKernel_t kern, const dim3& g, int which, int identity, int nargs) -> cudaError_t {
        void* args[7] = {(void*)&ctx->dev_scene, (void*)&dp, (void*)&ubo, (void*)&w, (void*)&which, (void*)&identity, nullptr};
        (void)nargs;
        ctx->launches++;
        return cudaLaunchKernel((const void*)kern, g, block, args, 0, ctx->stream);
    };
    auto ctl = [&](int op) -> cudaError_t {
        void* args[2] = {(void*)&w, (void*)&op};
        return cudaLaunchKernel((const void*)k->wf_ctl, one, dim3(32, 1, 1), args, 0, ctx->stream);
    };
    for (size_t pix0 = 0; pix0 < pixels; pix0 += band)
    for (int base = 0; base < dp0.samplesPerFrame; base += chunk) {
        const int ns = (dp0.samplesPerFrame - base < chunk) ? dp0.samplesPerFrame - base : chunk;
        w.pixBase = (unsigned)pix0;
        w.nPix = (unsigned)((pixels - pix0 < band) ? pixels - pix0 : band);
        w.P = (unsigned)((size_t)w.nPix * (size_t)ns);
        w.chunkBase = base;
        w.chunkSamples = ns;
        w.lastChunk = (base + ns >= dp0.samplesPerFrame) ? 1 : 0;
        w.qA = ctx->wf.qA;
        w.qB = ctx->wf.qB;
        PT_CUDA(ctx, ctl(0));
        PT_CUDA(ctx, run(k->wf_gen, grid, 0, 0, 4));
        for (int depth = 0; depth < dp0.pathLength; depth++) {
            const int identity = (depth == 0) ? 1 : 0;
            PT_CUDA(ctx, run(k->wf_isect, grid, 0, identity, 6));
            if (has_sdf) PT_CUDA(ctx, run(k->wf_march, grid, 0, 0, 4));
            if (ctx->wf_sort && ctx->dev_scene.nSurfaceExt > 0) { /* material-sorted shading: path rays grouped by lobe */
                PT_CUDA(ctx, run(k->wf_sort, grid, identity, 0, 6));
                PT_CUDA(ctx, ctl(3));
                PT_CUDA(ctx, run(k->wf_sort, grid, identity, 1, 6));
                PT_CUDA(ctx, ctl(4));
                PT_CUDA(ctx, run(k->wf_shade, grid, 2, identity, 6));
            } else {
                PT_CUDA(ctx, run(k->wf_shade, grid, 0, identity, 6));
            }
            if (ctx->dev_scene.numLights > 0.0f) { /* shadow rays of this depth */
                if (has_sdf) PT_CUDA(ctx, ctl(1));
                PT_CUDA(ctx, run(k->wf_isect, grid, 1, 0, 6));
                if (has_sdf) PT_CUDA(ctx, run(k->wf_march, grid, 0, 0, 4));
                PT_CUDA(ctx, run(k->wf_shade, grid, 1, 0, 6));
            }
            PT_CUDA(ctx, ctl(2));
            void* t = w.qA; w.qA = w.qB; w.qB = t;
        }
        {
            void* args[5] = {(void*)&ctx->dev_scene, (void*)&dp, (void*)&ubo, (void*)&w, (void*)&image};
            ctx->launches++;
            PT_CUDA(ctx, cudaLaunchKernel((const void*)k->wf_final, grid, block, args, 0, ctx->stream));
        }
    }
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int check_ready(pt_ctx* ctx, const pt_params* p) {
    if (!ctx || !p) return fail(ctx, PT_ERR_ARG, "null argument");
    if (!ctx->scene_set) return fail(ctx, PT_ERR_ARG, "pt_set_scene has not been called");
    if (!ctx->d_image) return fail(ctx, PT_ERR_ARG, "no image: call pt_resize or pt_bind_image first");
    if (p->resolution[0] != ctx->width || p->resolution[1] != ctx->height)
        return fail(ctx, PT_ERR_ARG, "params.resolution does not match the image size");
    return PT_OK;
}

}  // namespace

extern "C" {

const char* pt_version(void) { return "libpt_cuda 0.1 (sm_100a)"; }

const char* pt_last_error(const pt_ctx* ctx) {
    if (ctx) return ctx->error.c_str();
    if (g_last_error.empty()) return pt_scene_last_error().c_str();
    return g_last_error.c_str();
}

int pt_create(int device, pt_ctx** out) {
    if (!out) return fail(nullptr, PT_ERR_ARG, "pt_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(nullptr, PT_ERR_NOGPU, std::string("no CUDA device available (libpt_cuda has no CPU fallback): ") +
                                               cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, PT_ERR_ARG, "pt_create: device index out of range");
    pt_ctx* ctx = new pt_ctx();
    ctx->device = device;
    if (cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->num_sms <= 0) {
        (void)cudaGetLastError();
        ctx->num_sms = 148;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaMalloc((void**)&ctx->d_ubo, sizeof(float) * (PT_BVH_UBO_OFF + PT_BVH_MAX_FLOATS))) != cudaSuccess ||
        (e = cudaMallocHost((void**)&ctx->h_stage[0], sizeof(float) * (PT_BVH_UBO_OFF + PT_BVH_MAX_FLOATS))) != cudaSuccess ||
        (e = cudaMallocHost((void**)&ctx->h_stage[1], sizeof(float) * (PT_BVH_UBO_OFF + PT_BVH_MAX_FLOATS))) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming)) != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "pt_create");
        pt_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return PT_OK;
}

void pt_destroy(pt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->jit_cache)
        if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
    if (ctx->own_image && ctx->d_image) cudaFree(ctx->d_image);
    if (ctx->wf_block) cudaFree(ctx->wf_block);
    if (ctx->gen_stream) { cudaStreamSynchronize(ctx->gen_stream); cudaStreamDestroy(ctx->gen_stream); }
    if (ctx->d_gen) cudaFree(ctx->d_gen);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_snap) cudaEventDestroy(ctx->ev_snap);
    if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
    if (ctx->d_snap) cudaFree(ctx->d_snap);
    if (ctx->d_ubo) cudaFree(ctx->d_ubo);
    for (int i = 0; i < 2; i++) {
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->ev_stage[i]) cudaEventDestroy(ctx->ev_stage[i]);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int pt_set_mode(pt_ctx* ctx, int mode) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    if (mode != PT_MODE_STRICT && mode != PT_MODE_FAST) return fail(ctx, PT_ERR_ARG, "unknown mode");
    ctx->mode = mode;
    ctx->scene_set = false; /* kernels are selected in pt_set_scene */
    return PT_OK;
}

int pt_set_jit(pt_ctx* ctx, int policy) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    if (policy < 0 || policy > 2) return fail(ctx, PT_ERR_ARG, "jit policy must be 0, 1 or 2");
    ctx->jit_policy = policy;
    ctx->scene_set = false;
    return PT_OK;
}

int pt_set_bvh(pt_ctx* ctx, int min_prims) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    ctx->bvh_min = min_prims;
    ctx->scene_set = false;
    return PT_OK;
}

/* Tuning options (pt_abi.h): they select and parametrise the kernel pt_set_scene builds -- never the result */
int pt_set_option(pt_ctx* ctx, const char* key, long long value) {
    if (!ctx || !key) return fail(ctx, PT_ERR_ARG, "pt_set_option: null argument");
    if (std::string(key) == "pregen_max_mb") {
        if (value < 1) return fail(ctx, PT_ERR_ARG, "pt_set_option: pregen_max_mb must be positive");
        ctx->pregen_max_mb = value;
        return PT_OK;
    }
    if (std::string(key) == "wf_sort") {
        if (value < 0 || value > 1) return fail(ctx, PT_ERR_ARG, "pt_set_option: wf_sort is 0 or 1");
        ctx->wf_sort = value;
        return PT_OK;
    }
    if (std::string(key) == "wf_max_paths") {
        if (value < 0) return fail(ctx, PT_ERR_ARG, "pt_set_option: wf_max_paths must be positive (or 0 = auto)");
        ctx->wf_max_paths = value;
        return PT_OK;
    }
    if (pt_knob_set(&ctx->knobs, key, value) != 0)
        return fail(ctx, PT_ERR_ARG, std::string("pt_set_option: unknown option or value out of range: ") + key);
    ctx->scene_set = false; /* kernels are selected in pt_set_scene */
    return PT_OK;
}
int pt_get_option(const pt_ctx* ctx, const char* key, long long* value) {
    if (!ctx || !key || !value) return PT_ERR_ARG;
    if (std::string(key) == "wf_max_paths") { *value = ctx->wf_max_paths; return PT_OK; }
    if (std::string(key) == "wf_sort") { *value = ctx->wf_sort; return PT_OK; }
    if (std::string(key) == "pregen_max_mb") { *value = ctx->pregen_max_mb; return PT_OK; }
    return pt_knob_get(&ctx->knobs, key, value) == 0 ? PT_OK : PT_ERR_ARG;
}

int pt_bvh_active(const pt_ctx* ctx) { return (ctx && ctx->scene_set && ctx->bvh_active) ? 1 : 0; }

int pt_set_pipeline(pt_ctx* ctx, int pipeline) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    if (pipeline != PT_PIPE_MEGAKERNEL && pipeline != PT_PIPE_WAVEFRONT) return fail(ctx, PT_ERR_ARG, "unknown pipeline");
    ctx->pipeline = pipeline;
    ctx->scene_set = false;
    return PT_OK;
}

static_assert(PT_MAX_SURFACE_EXT == PT_DEV_MAX_SURFACE_EXT, "pt_abi.h and pt_dev_scene.h disagree");
int pt_set_surface_ext(pt_ctx* ctx, const pt_surface_ext* table, int n) {
    if (!ctx || n < 0 || (n > 0 && !table)) return fail(ctx, PT_ERR_ARG, "pt_set_surface_ext: bad argument");
    if (n > PT_MAX_SURFACE_EXT)
        return fail(ctx, PT_ERR_ARG, "pt_set_surface_ext: at most " + std::to_string(PT_MAX_SURFACE_EXT) + " materials can carry an extension");
    for (int i = 0; i < n; i++) {
        const pt_surface_ext& e = table[i];
        if (e.bsdf < PT_BSDF_REFERENCE || e.bsdf > PT_BSDF_DIELECTRIC)
            return fail(ctx, PT_ERR_ARG, "pt_set_surface_ext: entry " + std::to_string(i) + ": unknown bsdf " + std::to_string(e.bsdf));
        if (e.bsdf == PT_BSDF_GLOSSY && !(e.roughness >= 0.0f && e.roughness <= 1.0f))
            return fail(ctx, PT_ERR_ARG, "pt_set_surface_ext: entry " + std::to_string(i) + ": roughness must be in [0, 1]");
        if (e.bsdf == PT_BSDF_DIELECTRIC && !(e.ior == 0.0f || (e.ior >= 1.0f && e.ior <= 4.0f)))
            return fail(ctx, PT_ERR_ARG, "pt_set_surface_ext: entry " + std::to_string(i) + ": ior must be 0 (BK7) or in [1, 4]");
    }
    ctx->surface_ext.assign(table, table + n);
    return PT_OK;
}

int pt_set_scene(pt_ctx* ctx, const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf) {
    if (!ctx || !ubo) return fail(ctx, PT_ERR_ARG, "pt_set_scene: null argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    std::string err;
    PtDevScene sc;
    int rc = pt_prepare_scene(ubo, &sc, &err);
    if (rc != PT_OK) return fail(ctx, rc, err);
    if (n_sdf != sc.nSdfs)
        return fail(ctx, PT_ERR_ARG, "pt_set_scene: n_sdf (" + std::to_string(n_sdf) + ") != ubo.numObjects[5] (" +
                                         std::to_string(sc.nSdfs) + ")");
    /* surface extensions (not reference behaviour; empty table = the reference's shading, kernels built without them) */
    const bool surface_ext = pt_prepare_surface_ext(ctx->surface_ext.data(), (int)ctx->surface_ext.size(), &sc);
    if (surface_ext && ctx->jit_policy == 0)
        return fail(ctx, PT_ERR_COMPILE, "surface extensions need the run-time compiled kernel (pt_set_jit(ctx, 0) set)");
    JitKernel* jit = nullptr;
    const bool wavefront = ctx->pipeline == PT_PIPE_WAVEFRONT;
    /* enough bounded primitives for the tree to beat the scan (run-time compiled kernels only) */
    const int n_bounded = pt_bvh_bounded_prims(&sc);
    const bool bvh = ctx->jit_policy != 0 && ctx->bvh_min > 0 && n_bounded >= ctx->bvh_min && n_bounded >= 2 &&
                     n_bounded <= PT_BVH_MAX_PRIMS;
    std::vector<float> bvh_blob;
    if (bvh) {
        rc = pt_bvh_build(&sc, &bvh_blob, &err);
        if (rc != PT_OK) return fail(ctx, rc, err);
        if (bvh_blob.size() > (size_t)PT_BVH_MAX_FLOATS) return fail(ctx, PT_ERR_ARG, "BVH blob larger than its device buffer");
    }
    const bool want_jit = (n_sdf > 0) || ctx->jit_policy == 2 || wavefront || bvh || surface_ext;
    if ((n_sdf > 0 || wavefront) && ctx->jit_policy == 0)
        return fail(ctx, PT_ERR_COMPILE, "scenes with SDF snippets and the wavefront pipeline need run-time compilation (PT_JIT=0 set)");
    if (want_jit) {
        std::string unit;
        if (n_sdf > 0) {
            rc = pt_sdf_generate(sdf_glsl, n_sdf, ubo->sdfs, &unit, &err);
            if (rc != PT_OK) return fail(ctx, rc, err);
        }
        PtJitOptions opt;
        opt.mode = ctx->mode;
        opt.bake_counts = (ctx->jit_policy == 2);
        opt.wavefront = wavefront;
        opt.bvh = bvh;
        opt.surface_ext = surface_ext;
        const int counts[6] = {sc.nSpheres, sc.nPlanes, sc.nBoxes, sc.nLenses, sc.nCyclides, sc.nSdfs};
        memcpy(opt.counts, counts, sizeof counts);
        opt.knobs = ctx->knobs;
        /* the cache key is the complete translation unit: mode, baked counts, every resolved option, the SDF text */
        const std::string key = pt_jit_source(unit, opt);
        auto it = ctx->jit_cache.find(key);
        if (it == ctx->jit_cache.end()) {
            std::vector<char> cubin;
            std::string log;
            rc = pt_jit_compile(unit, opt, &cubin, &log);
            if (rc != PT_OK) return fail(ctx, rc, log);
            JitKernel jk;
            PT_CUDA(ctx, cudaLibraryLoadData(&jk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
            cudaError_t e = cudaLibraryGetKernel(&jk.kernel, jk.lib, "pt_render_jit");
            if (e != cudaSuccess) {
                cudaLibraryUnload(jk.lib);
                return cuda_fail(ctx, e, "cudaLibraryGetKernel(pt_render_jit)");
            }
            if (n_sdf > 0 && (e = cudaLibraryGetKernel(&jk.sdf_eval, jk.lib, "pt_sdf_eval_jit")) != cudaSuccess) {
                cudaLibraryUnload(jk.lib);
                return cuda_fail(ctx, e, "cudaLibraryGetKernel(pt_sdf_eval_jit)");
            }
            if (pt_jit_uses_pregen(key) && (e = cudaLibraryGetKernel(&jk.gen, jk.lib, "pt_gen_jit")) != cudaSuccess) {
                cudaLibraryUnload(jk.lib);
                return cuda_fail(ctx, e, "cudaLibraryGetKernel(pt_gen_jit)");
            }
            if (pt_jit_uses_resolve(key) && (e = cudaLibraryGetKernel(&jk.resolve, jk.lib, "pt_resolve_jit")) != cudaSuccess) {
                cudaLibraryUnload(jk.lib);
                return cuda_fail(ctx, e, "cudaLibraryGetKernel(pt_resolve_jit)");
            }
            if (wavefront) {
                struct { const char* name; cudaKernel_t* k; bool need; } wk[] = {
                    {"pt_wf_gen", &jk.wf_gen, true}, {"pt_wf_isect", &jk.wf_isect, true}, {"pt_wf_march", &jk.wf_march, n_sdf > 0},
                    {"pt_wf_shade", &jk.wf_shade, true}, {"pt_wf_final", &jk.wf_final, true}, {"pt_wf_ctl", &jk.wf_ctl, true},
                    {"pt_wf_sort", &jk.wf_sort, true}};
                for (auto& x : wk) {
                    if (!x.need) continue;
                    if ((e = cudaLibraryGetKernel(x.k, jk.lib, x.name)) != cudaSuccess) {
                        cudaLibraryUnload(jk.lib);
                        return cuda_fail(ctx, e, x.name);
                    }
                }
            }
            it = ctx->jit_cache.emplace(key, jk).first;
        }
        jit = &it->second;
    }
    ctx->ubo = *ubo;
    ctx->dev_scene = sc;
    ctx->active_jit = jit;
    ctx->bvh_active = bvh;
    {   /* one asynchronous upload from a pinned slot; the slot is reused two calls later, after its event */
        const unsigned slot = ctx->stage_next++ & 1u;
        PT_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage[slot]));
        float* h = ctx->h_stage[slot];
        memcpy(h, &ctx->ubo, sizeof(pt_ubo));
        size_t floats = PT_UBO_FLOATS;
        if (bvh) {
            memset(h + PT_UBO_FLOATS, 0, sizeof(float) * (PT_BVH_UBO_OFF - PT_UBO_FLOATS));
            memcpy(h + PT_BVH_UBO_OFF, bvh_blob.data(), bvh_blob.size() * sizeof(float));
            floats = PT_BVH_UBO_OFF + bvh_blob.size();
        }
        PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_ubo, h, floats * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        PT_CUDA(ctx, cudaEventRecord(ctx->ev_stage[slot], ctx->stream));
    }
    ctx->scene_set = true;
    ctx->error.clear();
    return PT_OK;
}

int pt_resize(pt_ctx* ctx, int width, int height) {
    if (!ctx || width <= 0 || height <= 0) return fail(ctx, PT_ERR_ARG, "pt_resize: bad size");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_image && ctx->d_image) cudaFree(ctx->d_image);
    ctx->d_image = nullptr;
    ctx->own_image = false;
    const size_t bytes = (size_t)width * (size_t)height * 16;
    PT_CUDA(ctx, cudaMalloc((void**)&ctx->d_image, bytes));
    ctx->own_image = true;
    ctx->width = width;
    ctx->height = height;
    /* the reference never clears a fresh texel buffer (n = 1 multiplies the old content by 0: SURVEY App. C-18);
     * garbage NaNs would survive that, so the image starts at zero here */
    PT_CUDA(ctx, cudaMemsetAsync(ctx->d_image, 0, bytes, ctx->stream));
    return PT_OK;
}

int pt_bind_image(pt_ctx* ctx, void* device_rgba32f, int width, int height) {
    if (!ctx || !device_rgba32f || width <= 0 || height <= 0) return fail(ctx, PT_ERR_ARG, "pt_bind_image: bad argument");
    if (((uintptr_t)device_rgba32f & 15) != 0) return fail(ctx, PT_ERR_ARG, "pt_bind_image: pointer must be 16-byte aligned");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_image && ctx->d_image) cudaFree(ctx->d_image);
    ctx->d_image = (float*)device_rgba32f;
    ctx->own_image = false;
    ctx->width = width;
    ctx->height = height;
    return PT_OK;
}

int pt_clear(pt_ctx* ctx) {
    if (!ctx || !ctx->d_image) return fail(ctx, PT_ERR_ARG, "pt_clear: no image");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaMemsetAsync(ctx->d_image, 0, (size_t)ctx->width * (size_t)ctx->height * 16, ctx->stream));
    return PT_OK;
}

int pt_dispatch(pt_ctx* ctx, const pt_params* params) {
    int rc = check_ready(ctx, params);
    if (rc != PT_OK) return rc;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PtDevParams dp;
    std::string err;
    rc = pt_prepare_params(params, 0, 0, 0, &dp, &err);
    if (rc != PT_OK) return fail(ctx, rc, err);
    return launch(ctx, dp);
}

int pt_dispatch_sum(pt_ctx* ctx, const pt_params* params, int first_sample, int n_samples) {
    int rc = check_ready(ctx, params);
    if (rc != PT_OK) return rc;
    if (n_samples <= 0) return fail(ctx, PT_ERR_ARG, "pt_dispatch_sum: n_samples must be positive");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PtDevParams dp;
    std::string err;
    rc = pt_prepare_params(params, 2, first_sample, n_samples, &dp, &err);
    if (rc != PT_OK) return fail(ctx, rc, err);
    return launch(ctx, dp);
}

int pt_finalize(pt_ctx* ctx, const pt_params* params, int total_samples) {
    int rc = check_ready(ctx, params);
    if (rc != PT_OK) return rc;
    if (total_samples <= 0) return fail(ctx, PT_ERR_ARG, "pt_finalize: total_samples must be positive");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    const float exposure = params->apertureSize * params->apertureSize * (float)params->ISO;
    pt_launch_finalize(ctx->d_image, ctx->width * ctx->height, 1.0f / (float)total_samples, exposure, ctx->stream);
    PT_CUDA(ctx, cudaGetLastError());
    return PT_OK;
}

int pt_render(pt_ctx* ctx, const pt_params* base, int total_samples, int samples_per_frame) {
    return pt_render_resume(ctx, base, 0, total_samples, samples_per_frame);
}

/* pt_render continuing a run that already holds done_samples samples per pixel in the image (a multiple of
 * samples_per_frame).  The reference's offscreen loop keeps dispatching while currentSamples < numSamples
 * (host:4042-4074), i.e. ceil(total / spf) dispatches of spf samples each -- a total that is not a multiple of
 * samples_per_frame renders the next multiple, as the reference does. */
int pt_render_resume(pt_ctx* ctx, const pt_params* base, int done_samples, int total_samples, int samples_per_frame) {
    if (!ctx || !base) return fail(ctx, PT_ERR_ARG, "pt_render: null argument");
    if (samples_per_frame <= 0 || total_samples <= 0 || done_samples < 0 || done_samples % samples_per_frame != 0)
        return fail(ctx, PT_ERR_ARG, "pt_render: need total_samples > 0, samples_per_frame > 0 and done_samples a multiple of it");
    pt_params p = *base;
    p.samplesPerFrame = samples_per_frame;
    const long long spf = samples_per_frame, total = total_samples;
    if (((total + spf - 1) / spf) * spf > 2147483647ll) return fail(ctx, PT_ERR_ARG, "pt_render: sample index would overflow the 32-bit frame counter");
    /* offscreen MainLoop bookkeeping (host:4042-4048): before dispatch j, frame = currentSamples = j * spf */
    for (long long j = done_samples / spf + 1; (j - 1) * spf < total; j++) {
        p.frame = (int)(j * spf);
        p.currentSamples = (int)(j * spf);
        int rc = pt_dispatch(ctx, &p);
        if (rc != PT_OK) return rc;
    }
    return pt_sync(ctx);
}

int pt_sync(pt_ctx* ctx) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PT_OK;
}

int pt_read_xyz(pt_ctx* ctx, float* rgba, size_t n_floats) {
    if (!ctx || !rgba || !ctx->d_image) return fail(ctx, PT_ERR_ARG, "pt_read_xyz: bad argument");
    const size_t need = (size_t)ctx->width * (size_t)ctx->height * 4;
    if (n_floats < need) return fail(ctx, PT_ERR_ARG, "pt_read_xyz: buffer too small");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->d_image, need * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PT_OK;
}

/* Read-back that overlaps the next dispatch: the image is snapshotted device-to-device on the compute stream (so later
 * dispatches may go on modifying it) and the snapshot travels to the host on a second stream.  `rgba` should be pinned
 * memory for the copy to be truly asynchronous; pt_read_wait() blocks until the last such copy has landed. */
int pt_read_xyz_async(pt_ctx* ctx, float* rgba, size_t n_floats) {
    if (!ctx || !rgba || !ctx->d_image) return fail(ctx, PT_ERR_ARG, "pt_read_xyz_async: bad argument");
    const size_t need = (size_t)ctx->width * (size_t)ctx->height * 4, bytes = need * sizeof(float);
    if (n_floats < need) return fail(ctx, PT_ERR_ARG, "pt_read_xyz_async: buffer too small");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) {
        PT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        PT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_snap, cudaEventDisableTiming));
        PT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
    }
    if (ctx->snap_bytes < bytes) {
        PT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->d_snap) cudaFree(ctx->d_snap);
        ctx->d_snap = nullptr;
        PT_CUDA(ctx, cudaMalloc((void**)&ctx->d_snap, bytes));
        ctx->snap_bytes = bytes;
    }
    if (ctx->copy_pending) PT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0)); /* snapshot buffer is free again */
    PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_snap, ctx->d_image, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    PT_CUDA(ctx, cudaEventRecord(ctx->ev_snap, ctx->stream));
    PT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_snap, 0));
    PT_CUDA(ctx, cudaMemcpyAsync(rgba, ctx->d_snap, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    PT_CUDA(ctx, cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
    ctx->copy_pending = true;
    return PT_OK;
}

int pt_read_wait(pt_ctx* ctx) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    if (!ctx->copy_pending) return PT_OK;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaEventSynchronize(ctx->ev_copied));
    ctx->copy_pending = false;
    return PT_OK;
}

/* Checkpoint / resume: the accumulation image IS the render state; uploading a saved one and
 * continuing with frame/currentSamples where the saved run stopped resumes it exactly. */
int pt_write_xyz(pt_ctx* ctx, const float* rgba, size_t n_floats) {
    if (!ctx || !rgba || !ctx->d_image) return fail(ctx, PT_ERR_ARG, "pt_write_xyz: bad argument");
    const size_t need = (size_t)ctx->width * (size_t)ctx->height * 4;
    if (n_floats < need) return fail(ctx, PT_ERR_ARG, "pt_write_xyz: buffer too small");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaMemcpyAsync(ctx->d_image, rgba, need * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PT_OK;
}

void* pt_image_ptr(pt_ctx* ctx) { return ctx ? (void*)ctx->d_image : nullptr; }
void* pt_stream_handle(pt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int pt_kernel_time(pt_ctx* ctx, float* ms, long long* launches) {
    if (!ctx) return fail(nullptr, PT_ERR_ARG, "null context");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    float t = 0.0f;
    if (ctx->timing_open) {
        PT_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        PT_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        PT_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
        ctx->timing_open = false;
    }
    if (ms) *ms = t;
    if (launches) *launches = ctx->launches;
    ctx->launches = 0;
    return PT_OK;
}

long pt_sdf_translate(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, char* out, size_t cap) {
    std::string unit, err;
    int rc = pt_sdf_generate(sdf_glsl, n_sdf, sdfs_raw, &unit, &err);
    if (rc != PT_OK) { fail(nullptr, rc, err); return rc; }
    if (out && cap > 0) {
        size_t n = unit.size() < cap - 1 ? unit.size() : cap - 1;
        memcpy(out, unit.data(), n);
        out[n] = '\0';
    }
    return (long)unit.size() + 1;
}

int pt_sdf_compile_check(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, int mode) {
    std::string unit, err, log;
    int rc = pt_sdf_generate(sdf_glsl, n_sdf, sdfs_raw, &unit, &err);
    if (rc != PT_OK) return fail(nullptr, rc, err);
    PtJitOptions opt;
    opt.mode = mode;
    opt.bake_counts = false;
    opt.wavefront = false;
    opt.bvh = false;
    memset(opt.counts, 0, sizeof opt.counts);
    opt.counts[5] = n_sdf;
    std::vector<char> cubin;
    rc = pt_jit_compile(unit, opt, &cubin, &log);
    if (rc != PT_OK) return fail(nullptr, rc, log);
    g_last_error = log;
    return PT_OK;
}

/* Compile (NVRTC, no GPU needed) exactly the kernel pt_set_scene would build for this scene, mode and jit policy.
 * The ptxas report (registers, spills) is left in pt_last_error(NULL). */
int pt_kernel_compile_check(const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf, int mode, int bake_counts) {
    return pt_kernel_compile_check_opts(ubo, sdf_glsl, n_sdf, mode, bake_counts, nullptr);
}
int pt_kernel_compile_check_opts(const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf, int mode, int bake_counts,
                                 const char* options) {
    if (!ubo) return fail(nullptr, PT_ERR_ARG, "null ubo");
    std::string unit, err, log;
    PtKnobs knobs;
    if (pt_knobs_parse(&knobs, options, &err) != 0) return fail(nullptr, PT_ERR_ARG, err);
    PtDevScene sc;
    int rc = pt_prepare_scene(ubo, &sc, &err);
    if (rc != PT_OK) return fail(nullptr, rc, err);
    if (n_sdf > 0) {
        rc = pt_sdf_generate(sdf_glsl, n_sdf, ubo->sdfs, &unit, &err);
        if (rc != PT_OK) return fail(nullptr, rc, err);
    }
    PtJitOptions opt;
    opt.mode = mode & 1;
    opt.bake_counts = bake_counts != 0;
    opt.wavefront = (mode & 2) != 0; /* mode bit 1: also build the wavefront kernels */
    opt.bvh = (mode & 4) != 0;       /* mode bit 2: closest hit through the BVH */
    opt.surface_ext = (mode & 8) != 0; /* mode bit 3: with the surface extensions (pt_set_surface_ext) */
    const int counts[6] = {sc.nSpheres, sc.nPlanes, sc.nBoxes, sc.nLenses, sc.nCyclides, sc.nSdfs};
    memcpy(opt.counts, counts, sizeof counts);
    opt.knobs = knobs;
    std::vector<char> cubin;
    rc = pt_jit_compile(unit, opt, &cubin, &log);
    if (rc != PT_OK) return fail(nullptr, rc, log);
    g_last_error = log;
    return PT_OK;
}

/* Measured FP32 peak of this device: a grid of independent FFMA chains (pt_kernels_fast.cu), timed with CUDA events on
 * the context's stream, best of `repeats`.  The roofline denominator of bench.py. */
int pt_fp32_peak(pt_ctx* ctx, int repeats, double* tflops, double* ms_best) {
    if (!ctx || !tflops) return fail(ctx, PT_ERR_ARG, "pt_fp32_peak: null argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d = nullptr;
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 4096;
    PT_CUDA(ctx, cudaMalloc((void**)&d, sizeof(float) * (size_t)blocks * threads));
    cudaEvent_t a, b;
    PT_CUDA(ctx, cudaEventCreate(&a));
    PT_CUDA(ctx, cudaEventCreate(&b));
    double best = 1e30;
    if (repeats < 1) repeats = 1;
    for (int r = 0; r < repeats + 1; r++) { /* first run warms up */
        cudaEventRecord(a, ctx->stream);
        pt_launch_fma_peak(d, blocks, threads, iters, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { cudaFree(d); return cuda_fail(ctx, e, "pt_fp32_peak"); }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads; /* 16 FFMA per iteration per thread */
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_best) *ms_best = best;
    return PT_OK;
}

int pt_sdf_eval(pt_ctx* ctx, const float* xyz, size_t n, unsigned set1, float* dist, float* material) {
    const unsigned sets[4] = {set1, 0u, 0u, 0u};
    return pt_sdf_eval4(ctx, xyz, n, sets, dist, material);
}
int pt_sdf_eval4(pt_ctx* ctx, const float* xyz, size_t n, const unsigned sets[4], float* dist, float* material) {
    if (!ctx || !xyz || !sets || n == 0) return fail(ctx, PT_ERR_ARG, "pt_sdf_eval: bad argument");
    if (!ctx->scene_set || !ctx->active_jit || !ctx->active_jit->sdf_eval)
        return fail(ctx, PT_ERR_ARG, "pt_sdf_eval: the current scene has no SDF snippets");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    float *dx = nullptr, *dd = nullptr, *dm = nullptr;
    PT_CUDA(ctx, cudaMalloc((void**)&dx, n * 12));
    PT_CUDA(ctx, cudaMalloc((void**)&dd, n * 4));
    PT_CUDA(ctx, cudaMalloc((void**)&dm, n * 4));
    cudaMemcpyAsync(dx, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream);
    unsigned long long nn = n;
    const float* a0 = dx;
    unsigned s0 = sets[0], s1 = sets[1], s2 = sets[2], s3 = sets[3];
    void* args[8] = {(void*)&a0, (void*)&nn, (void*)&s0, (void*)&s1, (void*)&s2, (void*)&s3, (void*)&dd, (void*)&dm};
    cudaError_t e = cudaLaunchKernel((const void*)ctx->active_jit->sdf_eval, dim3((unsigned)((n + 127) / 128)), dim3(128), args,
                                     0, ctx->stream);
    if (e == cudaSuccess && dist) e = cudaMemcpyAsync(dist, dd, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && material) e = cudaMemcpyAsync(material, dm, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dx); cudaFree(dd); cudaFree(dm);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "pt_sdf_eval");
    return PT_OK;
}

/* debug: scheduling statistics of a kernel built with option "stats" = 1 (16 counters; see pt_kernel.cuh) */
int pt_debug_stats(pt_ctx* ctx, unsigned long long* out16, int reset) {
    if (!ctx || !out16 || !ctx->active_jit) return fail(ctx, PT_ERR_ARG, "pt_debug_stats: needs a JIT kernel");
    void* dptr = nullptr;
    size_t bytes = 0;
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    PT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaError_t e = cudaLibraryGetGlobal(&dptr, &bytes, ctx->active_jit->lib, "pt_stats");
    if (e != cudaSuccess) return cuda_fail(ctx, e, "pt_stats symbol (pt_set_option(ctx, \"stats\", 1) before pt_set_scene)");
    PT_CUDA(ctx, cudaMemcpy(out16, dptr, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) PT_CUDA(ctx, cudaMemset(dptr, 0, 16 * sizeof(unsigned long long)));
    return PT_OK;
}

int pt_math_eval(pt_ctx* ctx, int fn, const float* x, const float* y, float* out, size_t n) {
    if (!ctx || !x || !out) return fail(ctx, PT_ERR_ARG, "pt_math_eval: null argument");
    PT_CUDA(ctx, cudaSetDevice(ctx->device));
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    PT_CUDA(ctx, cudaMalloc((void**)&dx, n * 4));
    PT_CUDA(ctx, cudaMalloc((void**)&dy, n * 4));
    PT_CUDA(ctx, cudaMalloc((void**)&dout, n * 4));
    cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (y) cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    else cudaMemsetAsync(dy, 0, n * 4, ctx->stream);
    pt_launch_math_eval(fn, dx, dy, dout, n, ctx->stream);
    cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "pt_math_eval");
    return PT_OK;
}

} /* extern "C" */
