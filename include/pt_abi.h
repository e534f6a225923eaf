/* pt_abi.h -- C ABI of libpt_cuda: the drop-in boundary for phyastro/PathTracer's compute dispatch.
 *
 * Every entry point names the reference interface it stands in for.  Citations: `shader.comp:N` is
 * src/shader.comp, `host:N` is src/pathtracer.cpp of the reference.
 *
 * The reference drives its kernel through a Vulkan compute pipeline whose whole interface is
 *   set 0 binding 0 : std430 uniform block of 4097 floats            (shader.comp:19-27 == host:187-195)
 *   set 0 binding 1 : rgba32f texel buffer of W*H texels, RMW         (shader.comp:29,1529-1532; host:2250-2269)
 *   push constants  : 88 bytes                                        (shader.comp:31-52 == host:197-218)
 *   source hook     : per-scene GLSL `float sdf(in vec3 p)` / `float sdfmaterial(in vec3 p)` pairs spliced into
 *                     the shader before compilation                   (shader.comp:704-719; host:2004-2054)
 * pt_ubo and pt_params below are bit-identical to those two blocks, so a maintainer of the reference can hand
 * `&ubo` and `&pushConstant` to this library unchanged.
 *
 * All functions return PT_OK (0) or a negative pt_status; pt_last_error() gives the message (NVRTC log included).
 * A pt_ctx is single-threaded; plain pointers and sizes only -- no C++ or torch types cross this boundary.
 */
#ifndef PT_ABI_H
#define PT_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_MAX_OBJECTS_SIZE 1024   /* host:39  / shader.comp:11 */
#define PT_MAX_SDFS_SIZE 768       /* host:40  / shader.comp:12 */
#define PT_MAX_MATERIALS_SIZE 783  /* host:41  / shader.comp:13 */
#define PT_MAX_LIGHTS_SIZE 128     /* host:42  / shader.comp:14 */
#define PT_MAX_LIGHTIDS_SIZE 64    /* host:43  / shader.comp:15 */
#define PT_CIE_SIZE 1323           /* 441 rows x 3, 360..800 nm (host:400-842) */
#define PT_MAX_SDF_SNIPPETS 128    /* sdfs[768] / 6 floats; four 32-bit masks set1..set4 (shader.comp:12, 706).  The reference
                                      only fills set1, i.e. 32 usable SDFs (734-738); libpt_cuda fills all four */

/* == UniformBufferObject (host:187-195) == `ubo` block (shader.comp:19-27); counts and ids are stored as floats */
typedef struct pt_ubo {
    float numObjects[7]; /* spheres, planes, boxes, lenses, cyclides, sdfs, sampled lights (host:3733-3739) */
    float objects[PT_MAX_OBJECTS_SIZE];
    float sdfs[PT_MAX_SDFS_SIZE];
    float materials[PT_MAX_MATERIALS_SIZE];
    float lights[PT_MAX_LIGHTS_SIZE];
    float lightIDs[PT_MAX_LIGHTIDS_SIZE];
    float CIEXYZ1931[PT_CIE_SIZE];
} pt_ubo; /* 4097 floats = 16388 bytes */

/* == PushConstantValues (host:197-218) == PushConstants (shader.comp:31-52) */
typedef struct pt_params {
    int32_t resolution[2];   /*  0 */
    int32_t frame;           /*  8 */
    int32_t currentSamples;  /* 12 */
    int32_t samplesPerFrame; /* 16 */
    float FPS;               /* 20 */
    float persistence;       /* 24 */
    int32_t pathLength;      /* 28 */
    float cameraAngle[2];    /* 32  = (-pitch, yaw) in degrees (host:3821) */
    float cameraPosX;        /* 40 */
    float cameraPosY;        /* 44 */
    float cameraPosZ;        /* 48 */
    int32_t ISO;             /* 52 */
    float cameraSize;        /* 56 */
    float apertureSize;      /* 60 */
    float apertureDist;      /* 64 */
    float lensRadius;        /* 68 */
    float lensFocalLength;   /* 72 */
    float lensThickness;     /* 76 */
    float lensDistance;      /* 80 */
    int32_t tonemap;         /* 84 */
} pt_params; /* 88 bytes */

typedef enum pt_status {
    PT_OK = 0,
    PT_ERR_ARG = -1,     /* bad argument / call order */
    PT_ERR_COMPILE = -2, /* SDF translation or NVRTC failure (the reference only prints these: host:1835-1860) */
    PT_ERR_CUDA = -3,    /* CUDA runtime/driver error (the reference throws std::runtime_error: host:4174-4179) */
    PT_ERR_IO = -4,      /* file / JSON error */
    PT_ERR_NOGPU = -5    /* no CUDA device: there is no CPU fallback */
} pt_status;

typedef enum pt_mode {
    PT_MODE_STRICT = 0, /* canonical IEEE semantics (pt_math.h, no contraction): bit-exact against oracle/ */
    PT_MODE_FAST = 1    /* fma contraction, MUFU intrinsics, approximate division: the throughput build */
} pt_mode;

typedef enum pt_pipeline {
    PT_PIPE_MEGAKERNEL = 0, /* one kernel per dispatch, path state in registers (pt_kernel.cuh) */
    PT_PIPE_WAVEFRONT = 1   /* generate / intersect / march / shade / accumulate kernels over SoA path state in HBM,
                               queues compacted by warp ballot (pt_wavefront.cuh); same results */
} pt_pipeline;

/* Surface extensions.  NOT reference behaviour: every surface of shader.comp is the Lambertian of lines
 * 1075-1091, and the reference's TODO.md:2 lists "Specular, Glossy Materials, Etc" as future work.  Off unless the
 * host asks: with no table set (the default) results are the reference's, bit for bit in strict mode. */
typedef enum pt_bsdf {
    PT_BSDF_REFERENCE = 0,  /* the reference's Lambertian with the material's spectrum */
    PT_BSDF_MIRROR = 1,     /* perfect reflection, tinted by the material's spectrum */
    PT_BSDF_GLOSSY = 2,     /* GGX conductor: alpha = roughness^2, Schlick Fresnel with F0 = the material's spectrum */
    PT_BSDF_DIELECTRIC = 3  /* smooth glass: exact Fresnel, refraction at `ior` (0: BK7 at the hero wavelength) */
} pt_bsdf;
typedef struct pt_surface_ext {
    int32_t bsdf;     /* pt_bsdf */
    float roughness;  /* PT_BSDF_GLOSSY */
    float ior;        /* PT_BSDF_DIELECTRIC */
    float pad;
} pt_surface_ext;
#define PT_MAX_SURFACE_EXT 64

typedef struct pt_ctx pt_ctx;

/* ---- device context (replaces InitVulkan + CreateComputePipeline, host:2424-2455, 2056-2092) ---------------- */
int pt_create(int device, pt_ctx** out);
void pt_destroy(pt_ctx* ctx);
const char* pt_last_error(const pt_ctx* ctx); /* ctx may be NULL: last error of a failed pt_create / loader call */
int pt_set_mode(pt_ctx* ctx, int mode);       /* default PT_MODE_STRICT; takes effect at the next pt_set_scene */
/* Run-time compilation policy (default 1): 0 = statically compiled kernels only (scenes
 * with SDF snippets are refused), 1 = NVRTC only for scenes with SDF snippets, 2 = always NVRTC, with the scene's
 * primitive counts baked in so the intersection loops unroll. Takes effect at the next pt_set_scene. */
int pt_set_jit(pt_ctx* ctx, int policy);
/* Megakernel (default) or wavefront pipeline; takes effect at the next pt_set_scene.  Both produce the same image
 * (bit-identical in strict mode); profiles/ compares them per scene. */
int pt_set_pipeline(pt_ctx* ctx, int pipeline);
/* Closest-hit search of Intersection / LightSourceVisibilityCheck (shader.comp:862-934, 1121-1216): the reference scans
 * every primitive; from min_prims spheres + boxes + lenses (default 12) libpt_cuda walks a host-built
 * BVH over those instead (run-time compiled kernels only, i.e. not with jit policy 0); planes and cyclides are still
 * scanned in order.  Same winner -- smallest t, ties to the lowest object index; bit-identical in strict mode.
 * min_prims <= 0 disables the tree.  Takes effect at the next pt_set_scene; pt_bvh_active tells whether the current scene uses it. */
int pt_set_bvh(pt_ctx* ctx, int min_prims);
int pt_bvh_active(const pt_ctx* ctx);
/* Tuning options of the run-time compiled kernels.  They choose HOW the kernel schedules the work, never WHAT it
 * computes: every setting renders the same image (bit-identical in strict mode).  -1 = auto (the measured default for
 * the scene).  Take effect at the next pt_set_scene; an unknown key or a value out of range is PT_ERR_ARG.
 *   "sched"       driver: 0 v1 nested loops, 7 v3s flat loop + sample pool, 5 v2s phase machine + sample pool,
 *                 8 v2m phase machine + pool of parked marching paths (SDF scenes)                        [-1]
 *   "sdf_reps"    SDF() evaluations per execution of the SDF phase                                          [8]
 *   "feed_t"      v2s / v2m: the SDF phase waits until every other phase has fewer lanes than this          [8]
 *   "regen_t"     v3s: finished lanes that trigger a regeneration                                           [16]
 *   "steal_s"     samples per pixel per round of the warp's sample pool; 0 = the whole dispatch (fast mode) [-1]
 *   "pool_cap" / "pool_min"   v2m: slots per warp / marching rays from which the SDF phase runs            [32 / 24]
 *   "min_blocks"  __launch_bounds__ minimum CTAs per SM                                                     [-1]
 *   "no_unroll"   1 keeps the primitive loops rolled although the counts are baked (jit policy 2)           [-1]
 *   "heavy_min"   v2s: > 0 runs the box / lens / cyclide tests as a phase of their own once so many lanes wait   [-1]
 *   "sin_poly_every"  fast mode: every k-th sin( of the SDF snippets runs on the FMA pipe instead of MUFU     [0]
 *   "pregen"      v2s / v3s: 1 = a generation kernel computes every sample's camera ray (seed, jitter, aperture, hero
 *                 wavelength, camera lens) with all lanes busy and the render kernel reads 32-byte records      [-1]
 *   "pathcolor_unroll"  SDF builds: the four CIE look-ups of a finished sample unrolled (auto: with pregen)      [-1]
 *   "resolve"     with pregen: 1 = the render kernel stores each sample's radiance bundle and a resolve kernel projects
 *                 to XYZ and sums per pixel in sample order (schedule-independent sums in fast mode; speed-neutral) [-1 = 0]
 *   "pregen_max_mb"  record buffer of one band in MiB; a dispatch that needs more runs as bands of 8-pixel rows
 *                 over two such buffers and two streams (4K x 64 spp: 17 GiB of records; bands of 4 / 2 / 1 GiB cost
 *                 0.2 / 0.7 / 4.5 %)                                                                          [4096]
 *   "stats"       1 builds the scheduling counters in (pt_debug_stats)                                      [0]
 *   "wf_sort"     wavefront pipeline, scenes with surface extensions: 1 = SHADE's path rays grouped by lobe by a
 *                 two-pass counting sort ("material-sorted shading"); same image; measured slower, hence           [0]
 *   "wf_refill", "wf_max_paths"   wavefront pipeline: evaluations between refills; paths in flight per chunk (a chunk is
 *                                 a band of consecutive pixels x some samples; 0 = auto: 4 Mi without SDFs, 32 Mi with)
 *   "bvh_while_while"             BVH traversal loop shape */
int pt_set_option(pt_ctx* ctx, const char* key, long long value);
int pt_get_option(const pt_ctx* ctx, const char* key, long long* value);

/* UpdateUniformBuffer + RecompileComputeShaders (host:3642-3811, 3836-3841, InsertSDF host:2004-2054).
 * sdf_glsl[i] is scene["sdf"][i]["glsl"] unchanged; n_sdf must equal ubo->numObjects[5].  With n_sdf > 0 the
 * kernel is rebuilt by NVRTC with the snippets translated against include/pt_glsl.h. */
int pt_set_scene(pt_ctx* ctx, const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf);
/* table[i] extends material i (0-based, the order of the scene file's "material" array) for the NEXT pt_set_scene;
 * n = 0 (the default) restores the reference's shading.  Needs the run-time compiled kernel (pt_set_jit != 0). */
int pt_set_surface_ext(pt_ctx* ctx, const pt_surface_ext* table, int n);

/* CreateTexelBuffer (host:2250-2269): library-owned W*H RGBA32F accumulation image, zero-filled */
int pt_resize(pt_ctx* ctx, int width, int height);
/* Same, but the caller owns the device memory (e.g. a torch tensor): W*H*4 floats, 16-byte aligned.  Not cleared. */
int pt_bind_image(pt_ctx* ctx, void* device_rgba32f, int width, int height);
int pt_clear(pt_ctx* ctx);

/* One vkCmdDispatch (host:3586-3608, DrawFrame host:3843-3880): renders samplesPerFrame samples per pixel with
 * sample indices frame-samplesPerFrame .. frame-1 and applies Accumulate() (shader.comp:1492-1507) to the image.
 * Asynchronous on the context's stream. */
int pt_dispatch(pt_ctx* ctx, const pt_params* params);

/* Sample-split building block: adds the raw XYZ of sample indices
 * first_sample .. first_sample+n_samples-1 to the image (no exposure, no mean).  pt_finalize then turns the sum
 * over total_samples into the reference's units: sum / total * apertureSize^2 * ISO, w = 1. */
int pt_dispatch_sum(pt_ctx* ctx, const pt_params* params, int first_sample, int n_samples);
int pt_finalize(pt_ctx* ctx, const pt_params* params, int total_samples);

/* Mapped texel memory (host:3491-3518 reads it through a persistent mapping): copies W*H*4 floats to the host.
 * Synchronises the stream. */
int pt_read_xyz(pt_ctx* ctx, float* rgba, size_t n_floats);
/* The same without blocking: the image is snapshotted on the device and copied to `rgba` (pinned host memory) on a
 * second stream while later dispatches run; pt_read_wait blocks until the most recent such copy has landed. */
int pt_read_xyz_async(pt_ctx* ctx, float* rgba, size_t n_floats);
int pt_read_wait(pt_ctx* ctx);
/* Checkpoint / resume (the accumulation image is the whole render state; the reference loses
 * it on exit): upload a saved image, then continue with pt_render_resume / pt_dispatch from the sample count reached */
int pt_write_xyz(pt_ctx* ctx, const float* rgba, size_t n_floats);
int pt_sync(pt_ctx* ctx);
void* pt_image_ptr(pt_ctx* ctx);    /* device pointer of the accumulation image */
void* pt_stream_handle(pt_ctx* ctx); /* cudaStream_t the kernels are launched on */
/* Device time of the kernels launched since the previous call (CUDA events on the context's stream), and how many */
int pt_kernel_time(pt_ctx* ctx, float* ms, long long* launches);

/* ---- convenience layer: the host code either side of the dispatch ------------------------------------------- */
typedef struct pt_scene pt_scene; /* parsed scene file: the std::vectors of host:1118-1127 */

/* ReadJSON + UpdateFromJSON (host:893-908, 2576-2722).  Same schema as scenes/ *.json; missing arrays are legal. */
int pt_scene_load_json(const char* path, pt_scene** out);
int pt_scene_parse_json(const char* text, pt_scene** out);
void pt_scene_free(pt_scene* scene);
int pt_scene_num_shots(const pt_scene* scene);
int pt_scene_num_sdf(const pt_scene* scene);
const char* pt_scene_sdf_glsl(const pt_scene* scene, int i);
/* UpdateToJSON + SaveScene (host:2724-2858, 3465-3472): the scene back as JSON text, reals rounded to 1e-5 like
 * RoundDecimal (host:935-943).  pt_scene_to_json returns the size needed (incl. NUL) and writes at most cap bytes. */
long pt_scene_to_json(const pt_scene* scene, char* out, size_t cap);
int pt_scene_save_json(const pt_scene* scene, const char* path);
/* UpdateUniformBuffer (host:3642-3811) incl. the CIE table copy of CreateUniformBuffer (host:2230-2248) */
int pt_scene_pack_ubo(const pt_scene* scene, pt_ubo* ubo);
/* The optional material keys "bsdf" ("mirror" | "glossy" | "dielectric"), "roughness", "ior" of the scene file (an
 * extension of the reference's schema; absent in every shipped scene).  Writes up to `max` entries and returns how many
 * to hand to pt_set_surface_ext: 0 when no material carries a "bsdf" key. */
int pt_scene_surface_ext(const pt_scene* scene, pt_surface_ext* out, int max);
/* UpdatePushConstant (host:3813-3834) for camera shot `shot` (1-based, host:1170); the per-dispatch fields are
 * set to the first offscreen dispatch: frame = currentSamples = samples_per_frame (host:4042-4048). */
int pt_scene_pack_params(const pt_scene* scene, int shot, int width, int height, int samples_per_frame,
                         int path_length, pt_params* params);

/* Offscreen MainLoop (host:4005-4086): total_samples/samples_per_frame dispatches with the reference's
 * frame/currentSamples bookkeeping. Blocks until done. */
int pt_render(pt_ctx* ctx, const pt_params* base, int total_samples, int samples_per_frame);
int pt_render_resume(pt_ctx* ctx, const pt_params* base, int done_samples, int total_samples, int samples_per_frame);

/* ---- the GPUs of one box from a single host thread (the reference drives one GPU) --------
 * One pt_ctx per device.  pt_multi_render gives device g the g-th contiguous slice of the sample-index range
 * (pt_dispatch_sum, issued round-robin: launches are asynchronous), then does the path's one exchange step -- a single
 * ncclReduce of the fp32 sum images to the first device over NVLink -- and pt_finalize there.  The union of the samples
 * equals a one-GPU run of the same range; images agree up to fp32 summation order.  libnccl.so.2 is dlopen'ed (env
 * PT_NCCL_LIB) and only needed for n_devices > 1.  bench.py does the same with one process per GPU. */
typedef struct pt_multi pt_multi;
int pt_multi_create(const int* devices, int n_devices, int mode, pt_multi** out);
void pt_multi_destroy(pt_multi* m);
const char* pt_multi_last_error(const pt_multi* m); /* m may be NULL: why the last pt_multi_create failed */
int pt_multi_num_devices(const pt_multi* m);
pt_ctx* pt_multi_ctx(pt_multi* m, int i);           /* for per-context settings (pt_set_jit, pt_set_bvh ...) */
int pt_multi_set_surface_ext(pt_multi* m, const pt_surface_ext* table, int n); /* before pt_multi_set_scene */
int pt_multi_set_scene(pt_multi* m, const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf);
int pt_multi_resize(pt_multi* m, int width, int height);
/* samples first_sample .. first_sample + total_samples - 1 of every pixel; blocks; seconds (optional) = host wall time */
int pt_multi_render(pt_multi* m, const pt_params* base, int first_sample, int total_samples, int samples_per_dispatch,
                    double* seconds);
double pt_multi_reduce_seconds(const pt_multi* m);  /* the ncclReduce of the last pt_multi_render alone */
int pt_multi_read_xyz(pt_multi* m, float* rgba, size_t n_floats); /* the final image, from the first device */

/* SaveRender + SavePPM (host:3491-3518, 918-933): display transform of shader.frag:31-93, 8-bit P6.
 * pt_write_pfm writes the raw XYZ (or linear sRGB when to_rgb != 0) as a bottom-up PF file. */
int pt_write_ppm(const char* path, const float* rgba, int width, int height, int tonemap);
int pt_write_pfm(const char* path, const float* rgba, int width, int height, int to_rgb);
int pt_read_pfm(const char* path, float* rgba, int width, int height); /* reads a to_rgb = 0 file back (w = 1) */
/* OpenEXR (single-part scanline, uncompressed, 3 x FLOAT): channels R,G,B (linear sRGB) or X,Y,Z (to_rgb == 0) */
int pt_write_exr(const char* path, const float* rgba, int width, int height, int to_rgb);

/* ---- introspection used by the tests --------------------------------------------------------------------- */
const float* pt_cie1931_table(void); /* 1323 floats */
/* GLSL snippet -> CUDA/C++ text exactly as pt_set_scene feeds NVRTC (prelude + snippets + dispatchers).
 * sdfs_raw = ubo.sdfs (6 floats per SDF: position, bounding size).  The text also compiles with g++, which is how
 * the CPU-side tests check the front end.  Returns the required size (including NUL); writes at most cap bytes. */
long pt_sdf_translate(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, char* out, size_t cap);
/* Compile-only check of the NVRTC path (needs no GPU): 0 on success, log via pt_last_error(NULL) */
int pt_sdf_compile_check(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, int mode);
/* Same for the whole kernel pt_set_scene would build for (ubo, snippets, mode); bake_counts as jit policy 2 */
int pt_kernel_compile_check(const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf, int mode, int bake_counts);
/* The same with tuning options given as "key=value,key=value" (see pt_set_option) */
int pt_kernel_compile_check_opts(const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf, int mode, int bake_counts,
                                 const char* options);
/* FP32 peak of the device, measured: independent FFMA chains on every SM, CUDA-event time, best of `repeats`.
 * bench.py's roofline denominator (MEASURED_PEAKS.json holds no FP32 figure). */
int pt_fp32_peak(pt_ctx* ctx, int repeats, double* tflops, double* ms_best);
/* Scheduling statistics of a kernel built with option "stats" = 1: per phase p of the v2 driver, out16[2p] = times the
 * phase ran, out16[2p+1] = lanes it served (debug / profiles only) */
int pt_debug_stats(pt_ctx* ctx, unsigned long long* out16, int reset);
/* Evaluate a pt_math.h function on the device: fn 0 sin,1 cos,2 acos,3 exp2,4 log2,5 exp,6 log,7 pow(x,y),8 PCG32 */
int pt_math_eval(pt_ctx* ctx, int fn, const float* x, const float* y, float* out, size_t n);
/* Evaluate SDF()/SDFMATERIAL() of the current scene at n points (xyz triples); set1 mask as given */
int pt_sdf_eval(pt_ctx* ctx, const float* xyz, size_t n, unsigned set1, float* dist, float* material);
int pt_sdf_eval4(pt_ctx* ctx, const float* xyz, size_t n, const unsigned sets[4], float* dist, float* material);
const char* pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PT_ABI_H */
