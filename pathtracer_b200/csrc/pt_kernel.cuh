/* pt_kernel.cuh -- the spectral path-tracing megakernel for sm_100a (hand-written CUDA; no GLSL translation).
 *
 * One source, compiled several ways:
 *   * nvcc, PT_MODE_STRICT: -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, transcendentals from pt_math.h.
 *     Bit-exact against the CPU oracle (oracle/oracle.cpp), which is the parity gate.
 *   * nvcc, PT_FAST: fma contraction, approximate division/sqrt, MUFU sin/cos/ex2/lg2.  The throughput build.
 *   * NVRTC (pt_jit.cpp), either mode, with the scene's SDF snippets spliced in (PT_HAS_SDF) and optionally the
 *     primitive counts baked in as compile-time constants (PT_N_*), so the intersection loops unroll and every
 *     scene constant becomes an immediate constant-bank operand.
 *
 * What the kernel computes is the reference's shader.comp main() (shader.comp:1525-1533): per pixel,
 * samplesPerFrame calls of Scene() (camera ray through the BK7 lens, hero-wavelength bundle, path tracing with
 * MIS light sampling and Russian roulette, CIE XYZ projection), the exposure scale and Accumulate().
 * Function names follow the shader; each carries the shader.comp lines it implements.
 *
 * How it differs from a transliteration (all of it result-preserving; see DESIGN.md "Kernel"):
 *   - per-object constants (rotation matrices, lens cap geometry, bounding radii, camera basis, light sampling
 *     cascade) arrive precomputed in constant memory (pt_dev_scene.h) instead of being rebuilt per ray;
 *   - the shadow ray (LightSourceVisibilityCheck) shares the intersection code but skips normals and materials,
 *     which it never uses; sphere tracing for it skips the 6 normal probes and the material evaluation;
 *   - EvaluateBRDF's spectral term is evaluated once per bounce and reused for the light sample;
 *   - the whole uniform block (CIE table, material/light tables: divergent indices) is staged in shared memory;
 *   - a warp covers an 8x4 pixel tile so primary rays stay coherent.
 *
 * Layout of this file: value types and mode-dependent primitives; the shader's functions (RNG, spectral helpers,
 * intersections, SDF search); then the DRIVERS that run them --
 *   v1  (PT_SCHED 0)  nested sample / bounce loops per thread            default for scenes without SDFs
 *   v2  (PT_SCHED 1)  four phases over explicit PathState, one phase per warp iteration chosen by ballot,
 *                     lanes refill themselves with their pixel's next sample   default for scenes with SDFs
 *   v2p (PT_SCHED 2)  v2 + a CTA-shared march pool (measured slower: profiles/r01_pool)
 *   v3  (PT_SCHED 3)  v1's loop bodies in one flat loop with gated path regeneration (PT_REGEN_T)
 *   v2d (PT_SCHED 4)  v2 with two pixels per lane, the idle one parked in shared memory: a lane only waits for the SDF
 *                     phase when both its paths do
 *   v2sp (PT_SCHED 6) v2s with persistent warps that stream over tiles claimed from a global counter, several tiles in
 *                     flight per warp (fast mode; strict builds fall back to v2s)
 *   v3s (PT_SCHED 7)  v3's flat loop with v2s' sample pool (verified on the host emulator, not yet measured)
 *   v2s (PT_SCHED 5)  v2 with in-warp sample stealing: the warp's 32 x S samples are a pool of work items, a lane that
 *                     finishes a path takes the next one whichever pixel it belongs to; per-sample XYZ in shared memory,
 *                     summed per pixel in sample order at the end of a round
 * -- and the kernel entry macros.  pt_wavefront.cuh runs the same phases as separate kernels over state in HBM.
 * The kernel is instruction-cache bound (16-byte SASS, 28-45 KB per scene): single call sites and rolled loops
 * are deliberate (profiles/README.md).
 */
#ifndef PT_KERNEL_CUH
#define PT_KERNEL_CUH

#include "pt_math.h"
#include "pt_dev_scene.h"
#ifndef PT_BVH
#define PT_BVH 0 /* 1: the JIT found enough bounded primitives for the tree of pt_bvh.h to pay (pt_lib.cpp) */
#endif
#if PT_BVH
#include "pt_bvh.h"
#endif

#ifndef PT_BLOCK_THREADS
#define PT_BLOCK_THREADS 128
#endif
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 4
#endif
#ifndef PT_HAS_SDF
#define PT_HAS_SDF 0
#endif

#define PT_DEV __device__ __forceinline__
#define PT_DEV_NOINLINE __device__ __noinline__

#define PT_CIE_FLOATS 1323

/* primitive counts: compile-time constants when the JIT bakes them in, else fields of the scene */
#ifdef PT_N_SPHERES_CONST
#define PT_N_SPHERES(c) (PT_N_SPHERES_CONST)
#define PT_N_PLANES(c) (PT_N_PLANES_CONST)
#define PT_N_BOXES(c) (PT_N_BOXES_CONST)
#define PT_N_LENSES(c) (PT_N_LENSES_CONST)
#define PT_N_CYCLIDES(c) (PT_N_CYCLIDES_CONST)
#define PT_N_SDF(c) (PT_N_SDF_CONST)
#if defined(PT_NO_UNROLL) && PT_NO_UNROLL
#define PT_UNROLL_PRIMS _Pragma("unroll 1")
#else
#define PT_UNROLL_PRIMS _Pragma("unroll")
#endif
#define PT_OFF_PLANES(sc) (8 * PT_N_SPHERES_CONST)
#define PT_OFF_BOXES(sc) (PT_OFF_PLANES(sc) + 4 * PT_N_PLANES_CONST)
#define PT_OFF_LENSES(sc) (PT_OFF_BOXES(sc) + 20 * PT_N_BOXES_CONST)
#define PT_OFF_CYCLIDES(sc) (PT_OFF_LENSES(sc) + 20 * PT_N_LENSES_CONST)
#define PT_OFF_SDFS(sc) (PT_OFF_CYCLIDES(sc) + 24 * PT_N_CYCLIDES_CONST)
#else
#define PT_OFF_PLANES(sc) ((sc).offPlanes)
#define PT_OFF_BOXES(sc) ((sc).offBoxes)
#define PT_OFF_LENSES(sc) ((sc).offLenses)
#define PT_OFF_CYCLIDES(sc) ((sc).offCyclides)
#define PT_OFF_SDFS(sc) ((sc).offSdfs)
#define PT_N_SPHERES(c) ((c).sc->nSpheres)
#define PT_N_PLANES(c) ((c).sc->nPlanes)
#define PT_N_BOXES(c) ((c).sc->nBoxes)
#define PT_N_LENSES(c) ((c).sc->nLenses)
#define PT_N_CYCLIDES(c) ((c).sc->nCyclides)
#define PT_N_SDF(c) ((c).sc->nSdfs)
#define PT_UNROLL_PRIMS
#endif

/* ---- mode-dependent primitives ------------------------------------------------------------------------------ */
#ifdef PT_FAST
#define PTK_SIN(x) __sinf(x)
#define PTK_COS(x) __cosf(x)
#define PTK_ACOS(x) acosf(x)
#define PTK_EXP(x) __expf(x)
#define PTK_POW(x, y) __powf(x, y)
#define PTK_SQRT(x) sqrtf(x)
#define PTK_DIV(a, b) __fdividef(a, b)
#define PTK_MIN(x, y) fminf(x, y)
#define PTK_MAX(x, y) fmaxf(x, y)
static __device__ __forceinline__ float ptk_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
static __device__ __forceinline__ float ptk_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
#define PTK_SIN(x) pt_sin(x)
#define PTK_COS(x) pt_cos(x)
#define PTK_ACOS(x) pt_acos(x)
#define PTK_EXP(x) pt_exp(x)
#define PTK_POW(x, y) pt_pow(x, y)
#define PTK_SQRT(x) sqrtf(x)          /* IEEE with -prec-sqrt=true */
#define PTK_DIV(a, b) ((a) / (b))     /* IEEE with -prec-div=true */
#define PTK_MIN(x, y) (((y) < (x)) ? (y) : (x)) /* GLSL min/max, NaN behaviour included */
#define PTK_MAX(x, y) (((x) < (y)) ? (y) : (x))
#endif

#if PT_HAS_SDF
/* provided by the generated translation unit (pt_sdf_front.cpp): the dispatchers InsertSDF builds (host:2004-2054) */
__device__ float pt_sdf_dispatch(float px, float py, float pz, unsigned set1);
__device__ float pt_sdfmaterial_dispatch(float px, float py, float pz, unsigned set1);
#endif

namespace PT_KERNEL_NS {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

PT_DEV V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_DEV V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
PT_DEV V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_DEV V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_DEV V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_DEV V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PT_DEV V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PT_DEV V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_DEV V3 div3(V3 a, V3 b) { return mk3(PTK_DIV(a.x, b.x), PTK_DIV(a.y, b.y), PTK_DIV(a.z, b.z)); }
#ifdef PT_FAST /* one MUFU.RCP shared by the components */
PT_DEV V3 div3(V3 a, float s) { const float r = ptk_rcp(s); return mk3(a.x * r, a.y * r, a.z * r); }
#else
PT_DEV V3 div3(V3 a, float s) { return mk3(PTK_DIV(a.x, s), PTK_DIV(a.y, s), PTK_DIV(a.z, s)); }
#endif
PT_DEV V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PT_DEV V4 operator*(V4 a, V4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PT_DEV V4 operator*(V4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
PT_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_DEV float length(V3 a) { return PTK_SQRT(dot(a, a)); }
#ifdef PT_FAST
PT_DEV V3 normalize(V3 a) { const float r = ptk_rsqrt(dot(a, a)); return mk3(a.x * r, a.y * r, a.z * r); }
#else
PT_DEV V3 normalize(V3 a) { return div3(a, length(a)); } /* GLSL 4.50 8.5: v / length(v) */
#endif
/* (b * num) / den per component, the shape of `EvaluateBRDF(...) * costheta / pdf` (shader.comp:1332,1374) */
#ifdef PT_FAST
PT_DEV V4 mulDiv4(V4 b, float num, float den) { const float k = num * ptk_rcp(den); return mk4(b.x * k, b.y * k, b.z * k, b.w * k); }
#else
PT_DEV V4 mulDiv4(V4 b, float num, float den) {
    return mk4(PTK_DIV(b.x * num, den), PTK_DIV(b.y * num, den), PTK_DIV(b.z * num, den), PTK_DIV(b.w * num, den));
}
#endif
PT_DEV V3 fma3(V3 a, float t, V3 c) { return mk3(fmaf(a.x, t, c.x), fmaf(a.y, t, c.y), fmaf(a.z, t, c.z)); }
PT_DEV float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
PT_DEV float gstep(float e, float x) { return (x < e) ? 0.0f : 1.0f; }
/* v * M (row vector times column-major matrix): component j = dot(v, column j) */
PT_DEV V3 mulVM(V3 v, const float* m) {
    return mk3(v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[3] + v.y * m[4] + v.z * m[5],
               v.x * m[6] + v.y * m[7] + v.z * m[8]);
}
/* M * v = col0*v.x + col1*v.y + col2*v.z */
PT_DEV V3 mulMV(const float* m, V3 v) {
    return mk3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
               m[2] * v.x + m[5] * v.y + m[8] * v.z);
}

struct Ray { V3 origin, dir; };

/* Everything a thread needs besides its registers */
struct Ctx {
    const PtDevScene* sc;   /* constant bank (kernel parameter) */
    const PtDevParams* pr;  /* constant bank (kernel parameter) */
    const float* ubo;       /* global: flat copy of the 4097-float uniform block */
    const float* s_tab;     /* shared: the whole 4097-float uniform block */
};

#define PT_SH_BASE 0 /* the whole block is staged: 16 388 B of shared memory per CTA, one LDS per table read */
#define PT_SH_FLOATS (PT_UBO_FLOATS - PT_SH_BASE)

/* clamped flat read of the uniform block: same rule as oracle Shader::at() */
PT_DEV float uboAt(const Ctx& c, int flat) {
    return c.s_tab[min(max(flat, 0), PT_UBO_FLOATS - 1)];
}

/* ---- RNG (shader.comp:937-958): pure uint32 arithmetic, bit-exact ------------------------------------------------ */
PT_DEV void PCG32(unsigned& seed) {
    unsigned state = seed * 747796405u + 2891336453u;
    unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    seed = (word >> 22u) ^ word;
}
PT_DEV float RandomFloatPCG32(unsigned& seed) {
    PCG32(seed);
    return __uint2float_rn(seed) * 2.3283064365386963e-10f; /* float(seed) / 2^32: exact scaling */
}

/* ---- spectral helpers ------------------------------------------------------------------------------------------ */
/* shader.comp:129-140.  index3 <= 1320; the +3..+5 reads of wave == 800 run past the table and clamp to the last
 * float of the block like the oracle's (unreachable in practice: the bundle lives on [390, 720)). */
PT_DEV V3 WaveToXYZ(const Ctx& c, float wave) {
    V3 XYZ = mk3(0.0f, 0.0f, 0.0f);
    if ((wave >= 360.0f) && (wave <= 800.0f)) {
        const float fl = floorf(wave);
        const int index3 = 3 * __float2int_rz(fl - 360.0f);
        const float a = wave - fl;
        const float oma = 1.0f - a;
        const float* t = c.s_tab + PT_OFF_CIE;
        const int i3 = min(index3 + 3, PT_CIE_FLOATS - 1), i4 = min(index3 + 4, PT_CIE_FLOATS - 1), i5 = min(index3 + 5, PT_CIE_FLOATS - 1);
        XYZ = mk3(t[index3] * oma + t[i3] * a, t[index3 + 1] * oma + t[i4] * a, t[index3 + 2] * oma + t[i5] * a);
    }
    return XYZ;
}

/* shader.comp:971-974 */
PT_DEV float gmod330(float x) { return x - 330.0f * floorf(PTK_DIV(x, 330.0f)); }
PT_DEV V4 SampleWavelengths(float l_h) {
    const float b = l_h - 390.0f;
    return mk4(390.0f + gmod330(b + 82.5f), 390.0f + gmod330(b + 165.0f), 390.0f + gmod330(b + 247.5f),
               390.0f + gmod330(b + 330.0f));
}

/* shader.comp:1030-1038 with EvaluateBRDF's "/ PI" (1075-1080): f = SPD / PI for the four wavelengths */
PT_DEV V4 EvaluateBRDF(V4 l, float peak, float sigma, float invertf) {
    const float den = 2.0f * sigma * sigma;
    const float a = (float)__float2int_rz(invertf); /* int(mat.reflection.z) -> float for mix() */
    const float oma = 1.0f - a;
    float r[4];
    const float lv[4] = {l.x, l.y, l.z, l.w};
#ifdef PT_FAST
    const float rden = ptk_rcp(den);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float x = (lv[i] - peak) * rden;
        const float e = PTK_EXP(-x * x);
        r[i] = (e * oma + (1.0f - e) * a) * 0.318309886f;
    }
#else
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float x = PTK_DIV(lv[i] - peak, den);
        const float e = PTK_EXP(-x * x);
        r[i] = PTK_DIV(e * oma + (1.0f - e) * a, PT_PI_F);
    }
#endif
    return mk4(r[0], r[1], r[2], r[3]);
}

/* shader.comp:1040-1055.  temperature/luminosity already clamped by max(., 0). */
PT_DEV V4 Emit(V4 l, float temperature, float luminosity) {
    float r[4];
    const float lv[4] = {l.x, l.y, l.z, l.w};
#ifdef PT_FAST
    /* same formula with the powers as multiplications and shared reciprocals: lum/peak * c1 * lm^-5 / (e^(c2/(lm T)) - 1) */
    const float t2 = temperature * temperature;
    const float k = luminosity * ptk_rcp(4.0956746759e-6f * (t2 * t2 * temperature));
    const float rT = ptk_rcp(temperature);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float rl = ptk_rcp(lv[i] * 1e-9f);
        const float rl2 = rl * rl;
        const float num = 1.1910429724e-16f * (rl2 * rl2 * rl);
        const float den = PTK_EXP(0.014387768775f * rl * rT) - 1.0f;
        r[i] = num * ptk_rcp(den) * k;
    }
#else
    const float peak = 4.0956746759e-6f * PTK_POW(temperature, 5.0f);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float lm = lv[i] * 1e-9f;
        const float num = 1.1910429724e-16f * PTK_POW(lm, -5.0f);
        const float den = PTK_EXP(PTK_DIV(0.014387768775f, lm * temperature)) - 1.0f;
        r[i] = PTK_DIV(PTK_DIV(num, den), peak) * luminosity;
    }
#endif
    return mk4(r[0], r[1], r[2], r[3]);
}

/* shader.comp:1064-1073 */
PT_DEV float RefractiveIndexBK7Glass(float l) {
    l *= 1e-3f;
    const float l2 = l * l;
    float n2 = 1.0f;
    n2 += PTK_DIV(1.03961212f * l2, l2 - 6.00069867e-3f);
    n2 += PTK_DIV(0.231792344f * l2, l2 - 2.00179144e-2f);
    n2 += PTK_DIV(1.01046945f * l2, l2 - 1.03560653e2f);
    return PTK_SQRT(n2);
}

/* shader.comp:216-235: material / light lookup through the flat uniform block */
PT_DEV void GetMaterialMix(const Ctx& c, float materialID, float& peak, float& sigma, float& invertf) {
    const float fl = floorf(materialID);
    const int i1 = 3 * __float2int_rz(fl);
    const int i2 = 3 * __float2int_rz(ceilf(materialID));
    const float x = materialID - fl;
    const float omx = 1.0f - x;
    peak = uboAt(c, PT_OFF_MAT + i1) * omx + uboAt(c, PT_OFF_MAT + i2) * x;
    sigma = uboAt(c, PT_OFF_MAT + i1 + 1) * omx + uboAt(c, PT_OFF_MAT + i2 + 1) * x;
    invertf = uboAt(c, PT_OFF_MAT + i1 + 2) * omx + uboAt(c, PT_OFF_MAT + i2 + 2) * x;
}
PT_DEV void GetLightMix(const Ctx& c, float lightID, float& temperature, float& luminosity) {
    const int index = __float2int_rz(lightID);
    if (index == -1) {
        temperature = 5500.0f;
        luminosity = 0.0f;
        return;
    }
    temperature = uboAt(c, PT_OFF_LGT + 2 * index);
    luminosity = uboAt(c, PT_OFF_LGT + 2 * index + 1);
}

/* ---- hit record -------------------------------------------------------------------------------------------- */
struct Hit {
    float t;          /* hitdist */
    V3 normal;
    float materialID;
    float lightID;
    int objectID;     /* global object index (LightSourceVisibilityCheck, shader.comp:1127) */
};

/* The reference scans primitives in index order with `if (t < hit.t)`: smallest t wins, ties go to the lowest index.
 * The BVH visits them in ray order, so there the tie rule is spelled out (kTie); the in-order scan does not need it. */
PT_DEV bool CloserHit(float t, int objectID, const Hit& h, const bool kTie) {
    return (t < h.t) || (kTie && (t == h.t) && (objectID < h.objectID));
}

/* shader.comp:263-276 */
PT_DEV bool BoundingSphere(const Ray& ray, float px, float py, float pz, float radius2) {
    const V3 lo = mk3(ray.origin.x - px, ray.origin.y - py, ray.origin.z - pz);
    const float b = dot(ray.dir, lo);
    const float cc = dot(lo, lo) - radius2;
    if ((b * b) < cc) return false;
    if ((b >= 0.0f) && (cc >= 0.0f)) return false;
    return true;
}

/* shader.comp:289-317 */
PT_DEV void SphereIntersection(const Ray& ray, const PtDevSphere& o, int objectID, Hit& h, const bool kShadow, const bool kTie = false) {
    const V3 lo = mk3(ray.origin.x - o.px, ray.origin.y - o.py, ray.origin.z - o.pz);
    const float b = 2.0f * dot(ray.dir, lo);
    const float cc = dot(lo, lo) - o.r2;
    const float discriminant = b * b - 4.0f * cc;
    if (discriminant < 0.0f) return;
    const float sqrtD = PTK_SQRT(discriminant);
    const float t1 = (-b - sqrtD) * 0.5f;
    const float t2 = (-b + sqrtD) * 0.5f;
    const float t = (t1 > 0.0f) ? t1 : t2;
    if (t < 1e-4f) return;
    if (CloserHit(t, objectID, h, kTie)) {
        h.t = t;
        h.objectID = objectID;
        if (!kShadow) {
            const float isOutside = (t1 > 0.0f) ? 1.0f : -1.0f;
            h.normal = normalize(fma3(ray.dir, t, lo) * isOutside);
            h.materialID = o.materialID;
            h.lightID = o.lightID;
        }
    }
}

/* shader.comp:319-335 */
PT_DEV void PlaneIntersection(const Ray& ray, const PtDevPlane& o, int objectID, Hit& h, const bool kShadow, const bool kTie = false) {
    const float loy = ray.origin.y - o.py;
    const float t = PTK_DIV(-loy, ray.dir.y);
    if (t < 1e-4f) return;
    if (CloserHit(t, objectID, h, kTie)) {
        h.t = t;
        h.objectID = objectID;
        if (!kShadow) {
            /* faceforward((0,1,0), dir, (0,1,0)): dot(Nref, I) = 0*dx + 1*dy + 0*dz */
            const float d = 0.0f * ray.dir.x + 1.0f * ray.dir.y + 0.0f * ray.dir.z;
            h.normal = (d < 0.0f) ? mk3(0.0f, 1.0f, 0.0f) : mk3(-0.0f, -1.0f, -0.0f);
            h.materialID = o.materialID;
            h.lightID = o.lightID;
        }
    }
}

/* shader.comp:337-364 */
PT_DEV void BoxIntersection(const Ray& ray, const PtDevBox& o, int objectID, Hit& h, const bool kShadow, const bool kTie = false) {
    const V3 lo = mulVM(mk3(ray.origin.x - o.px, ray.origin.y - o.py, ray.origin.z - o.pz), o.m);
    const V3 dir = mulVM(ray.dir, o.m);
    const V3 invdir = mk3(PTK_DIV(1.0f, dir.x), PTK_DIV(1.0f, dir.y), PTK_DIV(1.0f, dir.z));
    const V3 a = mk3(fmaf(o.sx, -0.5f, -lo.x) * invdir.x, fmaf(o.sy, -0.5f, -lo.y) * invdir.y,
                     fmaf(o.sz, -0.5f, -lo.z) * invdir.z);
    const V3 b = mk3(fmaf(o.sx, 0.5f, -lo.x) * invdir.x, fmaf(o.sy, 0.5f, -lo.y) * invdir.y,
                     fmaf(o.sz, 0.5f, -lo.z) * invdir.z);
    const V3 tMin = mk3(PTK_MIN(a.x, b.x), PTK_MIN(a.y, b.y), PTK_MIN(a.z, b.z));
    const V3 tMax = mk3(PTK_MAX(a.x, b.x), PTK_MAX(a.y, b.y), PTK_MAX(a.z, b.z));
    const float t1 = PTK_MAX(PTK_MAX(tMin.x, tMin.y), tMin.z);
    const float t2 = PTK_MIN(PTK_MIN(tMax.x, tMax.y), tMax.z);
    const float t = (t1 < 0.0f) ? t2 : t1;
    if ((t1 > t2) || (t < 1e-4f)) return;
    if (CloserHit(t, objectID, h, kTie)) {
        h.t = t;
        h.objectID = objectID;
        if (!kShadow) {
            const V3 p = mk3(fabsf(PTK_DIV(lo.x + dir.x * t, o.sx)), fabsf(PTK_DIV(lo.y + dir.y * t, o.sy)),
                             fabsf(PTK_DIV(lo.z + dir.z * t, o.sz)));
            const float pm = PTK_MAX(PTK_MAX(p.x, p.y), p.z);
            const V3 n = mk3(gstep(pm, p.x) * -gsign(dir.x), gstep(pm, p.y) * -gsign(dir.y),
                             gstep(pm, p.z) * -gsign(dir.z));
            h.normal = mulMV(o.m, n);
            h.materialID = o.materialID;
            h.lightID = o.lightID;
        }
    }
}

/* shader.comp:366-448: LensIntersection = the two spherical caps of SphereSliceIntersection, first slice then
 * second.  The caps differ only by a sign s = +1 / -1 (origin shift `+= shift` / `-= shift`, cut test
 * `x > -sliceOffset` / `x < sliceOffset`, i.e. `s*x > -sliceOffset`; multiplying by +-1 is exact), so one copy of
 * the code runs twice: the kernel is instruction-cache bound, and this body is instantiated for the scene's
 * lenses and for the camera lens. */
PT_DEV void LensIntersection(const Ray& ray, const PtDevLens& o, int objectID, Hit& h, int& isOutside, const bool kShadow, const bool kTie = false) {
    const V3 lo0 = mulVM(mk3(ray.origin.x - o.px, ray.origin.y - o.py, ray.origin.z - o.pz), o.m);
    const V3 ldir = mulVM(ray.dir, o.m);
#pragma unroll 1
    for (int slice = 0; slice < 2; slice++) {
        const float s = (slice == 0) ? 1.0f : -1.0f;
        V3 lo = lo0;
        lo.x += s * o.shift;
        const float b = 2.0f * dot(ldir, lo);
        const float cc = dot(lo, lo) - o.sradius2;
        const float discriminant = b * b - 4.0f * cc;
        if (discriminant < 0.0f) continue;
        const float sqrtD = PTK_SQRT(discriminant);
        float t1 = (-b - sqrtD) * 0.5f;
        float t2 = (-b + sqrtD) * 0.5f;
        t1 = ((s * fmaf(ldir.x, t1, lo.x)) > -o.sliceOffset) ? 1e6f : t1;
        t2 = ((s * fmaf(ldir.x, t2, lo.x)) > -o.sliceOffset) ? 1e6f : t2;
        float t = (t1 > 0.0f) ? t1 : 1e6f;
        int isOut = 1;
        if (t2 < t) {
            t = t2;
            isOut = -1;
        }
        if (t < 1e-4f) continue;
        if (CloserHit(t, objectID, h, kTie)) {
            h.t = t;
            h.objectID = objectID;
            if (!kShadow) {
                h.normal = mulMV(o.m, normalize(fma3(ldir, t, lo) * (float)isOut));
                isOutside = (o.invertSide == 0.0f) ? isOut : -isOut;
                h.materialID = o.materialID;
                h.lightID = o.lightID;
            }
        }
    }
}

/* ---- Dupin cyclide (shader.comp:450-541, 633-679) ------------------------------------------------------------ */
PT_DEV float EvalCubic1(float b, float c, float d, float x) { return x * (x * (x * 1.0f + b) + c) + d; }
PT_DEV float EvalQuadratic3(float b2, float c, float x) { return x * (x * 3.0f + b2) + c; }

/* shader.comp:474-503: only roots.x is consumed by SolveQuartic, so only it is returned (the other two roots of
 * the three-real-root branch never influence anything). */
PT_DEV float SolveCubicFirstRoot(float b, float c, float d) {
    const float ONEBYTHREE = 0.3333333f;
    const float bdiv3 = b * ONEBYTHREE;
    const float Q = c * ONEBYTHREE - bdiv3 * bdiv3;
    const float R = 0.5f * bdiv3 * c - bdiv3 * bdiv3 * bdiv3 - 0.5f * d;
    const float D = Q * Q * Q + R * R;
    float root;
    if (D > 0.0f) {
        const float sD = PTK_SQRT(D);
        const float u = R + sD;
        const float v = R - sD;
        const float S = gsign(u) * PTK_POW(fabsf(u), ONEBYTHREE);
        const float T = gsign(v) * PTK_POW(fabsf(v), ONEBYTHREE);
        root = S + T - bdiv3;
    } else {
        const float sqrtnegQ = PTK_SQRT(-Q);
        const float thetadiv3 = PTK_ACOS(PTK_DIV(R, sqrtnegQ * sqrtnegQ * sqrtnegQ)) * ONEBYTHREE;
        root = 2.0f * sqrtnegQ * PTK_COS(thetadiv3) - bdiv3;
    }
    const float b2 = 2.0f * b;
#pragma unroll
    for (int i = 0; i < 2; i++) root -= PTK_DIV(EvalCubic1(b, c, d, root), EvalQuadratic3(b2, c, root));
    return root;
}

PT_DEV float QuarticNewton(float a, float b, float c, float d, float e, float a4, float b3, float c2, float x) {
    const float q = x * (x * (x * (x * a + b) + c) + d) + e;
    const float dq = x * (x * (x * a4 + b3) + c2) + d;
    return x - PTK_DIV(q, dq);
}

/* shader.comp:505-541 + the root selection of 648-655: smallest positive real root, or 1e6 */
PT_DEV float SolveQuarticNearest(float a, float b, float c, float d, float e) {
    const float inva = PTK_DIV(1.0f, a);
    const float inva2 = inva * 0.5f;
    const float inva2a2 = inva2 * inva2;
    const float bb = b * b;
    const float p = -1.5f * bb * inva2a2 + c * inva;
    const float q = bb * b * inva2a2 * inva2 - b * c * inva * inva2 + d * inva;
    const float r = -0.1875f * bb * bb * inva2a2 * inva2a2 + 0.5f * c * bb * inva2a2 * inva2 - b * d * inva2a2 + e * inva;
    const float sx = SolveCubicFirstRoot(0.5f * -p, -r, 0.5f * p * r - 0.125f * q * q);
    const float s2subp = 2.0f * sx - p;
    float t = 1e6f;
    if (s2subp < 0.0f) return t;
    const float invs2subp = -2.0f * sx - p;
    const float sqrts2subp = PTK_SQRT(s2subp);
    const float q2divsqrt = PTK_DIV(2.0f * q, sqrts2subp);
    const float invaddq2div = invs2subp + q2divsqrt;
    const float invsubq2div = invs2subp - q2divsqrt;
    const float bdiv4a = 0.25f * inva * b;
    const float a4 = 4.0f * a, b3 = 3.0f * b, c2 = 2.0f * c;
    if (invaddq2div >= 0.0f) {
        const float sq = PTK_SQRT(invaddq2div);
        const float r0 = QuarticNewton(a, b, c, d, e, a4, b3, c2, 0.5f * (-sqrts2subp + 1.0f * sq) - bdiv4a);
        const float r1 = QuarticNewton(a, b, c, d, e, a4, b3, c2, 0.5f * (-sqrts2subp + -1.0f * sq) - bdiv4a);
        if ((r0 < t) && (r0 > 0.0f)) t = r0;
        if ((r1 < t) && (r1 > 0.0f)) t = r1;
    }
    if (invsubq2div >= 0.0f) {
        const float sq = PTK_SQRT(invsubq2div);
        const float r2 = QuarticNewton(a, b, c, d, e, a4, b3, c2, 0.5f * (sqrts2subp + 1.0f * sq) - bdiv4a);
        const float r3 = QuarticNewton(a, b, c, d, e, a4, b3, c2, 0.5f * (sqrts2subp + -1.0f * sq) - bdiv4a);
        if ((r2 < t) && (r2 > 0.0f)) t = r2;
        if ((r3 < t) && (r3 > 0.0f)) t = r3;
    }
    return t;
}

/* shader.comp:633-679.  The one out-of-line function of the kernel (the quartic solver is large and rarely reached):
 * everything goes in and out BY VALUE, in registers.  Passing the caller's Ray / Hit by reference made their address
 * escape, which parked the whole PathState of the in-warp drivers in local memory (176-byte stack frame, 4 % of all
 * executed instructions local loads/stores on the scenes with a cyclide).  Returns (t, normal) of an accepted hit --
 * `t < hT`, or `t == hT` when the caller says a tie goes to this object -- or t = -1 (roots are positive). */
PT_DEV_NOINLINE float4 DupinCyclideCore(float rox, float roy, float roz, float rdx, float rdy, float rdz, const PtDevCyclide& ob,
                                        float hT, int tieWins, int wantNormal) {
    const V3 lo = mulVM(mk3(rox - ob.px, roy - ob.py, roz - ob.pz), ob.m);
    const V3 ld = mulVM(mk3(rdx, rdy, rdz), ob.m);
    /* .xzy swizzle after the divide by scale */
    const V3 o = mk3(PTK_DIV(lo.x, ob.sx), PTK_DIV(lo.z, ob.sz), PTK_DIV(lo.y, ob.sy));
    const V3 d = mk3(PTK_DIV(ld.x, ob.sx), PTK_DIV(ld.z, ob.sz), PTK_DIV(ld.y, ob.sy));
    const float A = ob.a, B = ob.b, C = ob.c, D = ob.d;
    const V3 dd = d * d, oo = o * o, od = o * d;
    const V3 dyzx = mk3(d.y, d.z, d.x), dzxy = mk3(d.z, d.x, d.y);
    const V3 oyzx = mk3(o.y, o.z, o.x), ozxy = mk3(o.z, o.x, o.y);
    const V3 dyzx2 = dyzx * dyzx, dzxy2 = dzxy * dzxy;
    const float BBmDD = B * B - D * D;
    const float a4 = dot(dd, dd) + 2.0f * dot(dd, dyzx2);
    const float a3 = 4.0f * (dot(o, dd * d) + dot(od, dyzx2) + dot(od, dzxy2));
    const float a2 = 6.0f * dot(oo, dd) + 8.0f * dot(od, oyzx * dyzx) + 2.0f * (dot(oo, dyzx2) + dot(oo, dzxy2)) +
                     2.0f * BBmDD * dot(d, d) - 4.0f * (A * A * d.x * d.x + B * B * d.y * d.y);
    const float a1 = 4.0f * (dot(oo * o, d) + dot(oo, oyzx * dyzx) + dot(oo, ozxy * dzxy) + 2.0f * A * C * D * d.x +
                             BBmDD * dot(o, d) - 2.0f * (A * A * o.x * d.x + B * B * o.y * d.y));
    const float a0 = dot(oo, oo) + 2.0f * dot(oo, oyzx * oyzx) + B * B * B * B + D * D * D * D - 2.0f * B * B * D * D -
                     4.0f * C * C * D * D + 8.0f * A * C * D * o.x + 2.0f * BBmDD * dot(o, o) -
                     4.0f * (A * A * o.x * o.x + B * B * o.y * o.y);
    const float t = SolveQuarticNearest(a4, a3, a2, a1, a0);
    float4 r = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
    if ((t < hT) || (tieWins && (t == hT))) { /* CloserHit */
        r.x = t;
        if (wantNormal) {
            const float x = o.x + d.x * t;
            const float y = o.y + d.y * t;
            const float z = o.z + d.z * t;
            const float term1 = x * x + y * y + z * z + B * B - D * D;
            V3 n;
            n.x = 4.0f * (x * term1 - 2.0f * A * (A * x - C * D));
            n.y = 4.0f * z * term1;
            n.z = 4.0f * y * (term1 - 2.0f * B * B);
            n = normalize(n);
            r.y = n.x; r.z = n.y; r.w = n.z;
        }
    }
    return r;
}
PT_DEV void DupinCyclide(const Ray& ray, const PtDevCyclide& ob, int objectID, Hit& h, const bool kShadow, const bool kTie = false) {
    const int tieWins = (kTie && (objectID < h.objectID)) ? 1 : 0;
    const float4 r = DupinCyclideCore(ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x, ray.dir.y, ray.dir.z, ob, h.t,
                                      tieWins, kShadow ? 0 : 1);
    if (r.x >= 0.0f) {
        h.t = r.x;
        h.objectID = objectID;
        if (!kShadow) {
            h.normal = mk3(r.y, r.z, r.w);
            h.materialID = ob.materialID;
            h.lightID = ob.lightID;
        }
    }
}

/* ---- SDF sphere tracing (shader.comp:704-860) ------------------------------------------------------------------ */
#if PT_HAS_SDF
PT_DEV float SDF(V3 p, unsigned set1) { return ::pt_sdf_dispatch(p.x, p.y, p.z, set1); }

/* shader.comp:278-287 with box = SDF bounding box */
PT_DEV void RayIntersectAABB(V3 origin, V3 invdir, const PtDevSdf& s, float& t1, float& t2) {
    const V3 lo = mk3(origin.x - s.px, origin.y - s.py, origin.z - s.pz);
    const V3 a = mk3(fmaf(s.sx, -0.5f, -lo.x) * invdir.x, fmaf(s.sy, -0.5f, -lo.y) * invdir.y,
                     fmaf(s.sz, -0.5f, -lo.z) * invdir.z);
    const V3 b = mk3(fmaf(s.sx, 0.5f, -lo.x) * invdir.x, fmaf(s.sy, 0.5f, -lo.y) * invdir.y,
                     fmaf(s.sz, 0.5f, -lo.z) * invdir.z);
    t1 = PTK_MAX(PTK_MAX(PTK_MIN(a.x, b.x), PTK_MIN(a.y, b.y)), PTK_MIN(a.z, b.z));
    t2 = PTK_MIN(PTK_MIN(PTK_MAX(a.x, b.x), PTK_MAX(a.y, b.y)), PTK_MAX(a.z, b.z));
}

/* shader.comp:732-777 */
PT_DEV bool SearchSDF(const Ctx& c, V3 p, V3 invdir, float& tMin, float& tMax, unsigned& set1) {
    bool isFoundSDF = false;
    set1 = 0u;
    const int n = PT_N_SDF(c);
    for (int i = 0; i < n; i++) {
        float bx, by;
        RayIntersectAABB(p, invdir, reinterpret_cast<const PtDevSdf*>(c.sc->pool + PT_OFF_SDFS(*c.sc))[i], bx, by);
        if ((bx > by) || (by < 0.0f)) continue;
        const unsigned bit = 1u << (unsigned)i;
        if (bx < tMin) {
            isFoundSDF = true;
            if (by < tMin) {
                tMin = bx; tMax = by;
                set1 = bit;
            } else {
                if (by < tMax) {
                    tMin = bx;
                    set1 += bit;
                } else {
                    tMin = bx; tMax = by;
                    set1 += bit;
                }
            }
        } else {
            if (bx < tMax) {
                isFoundSDF = true;
                if (by > tMax) tMax = by;
                set1 += bit;
            }
        }
    }
    return isFoundSDF;
}

/* shader.comp:779-860.  For shadow rays the normal probes and the material are skipped (never read). */
PT_DEV_NOINLINE void SphereTracing(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    const float MAXDIST = 1e5f;
    float t = 1e-3f;
    float insT = 0.0f;
    const float omegaMax = 1.70f;
    const float omegaSpeed = 0.20f;
    float omega = omegaMax;
    float previousRadius = 0.0f;
    V3 p = ray.origin;
    const V3 invdir = mk3(PTK_DIV(1.0f, ray.dir.x), PTK_DIV(1.0f, ray.dir.y), PTK_DIV(1.0f, ray.dir.z));
    int points = 0;
    float tMin = MAXDIST, tMax = MAXDIST;
    unsigned set1 = 0u;
    if (SearchSDF(c, p, invdir, tMin, tMax, set1)) {
        t = PTK_MAX(tMin, t);
        p = fma3(ray.dir, t, ray.origin);
    } else {
        return;
    }
    const float k = gsign(SDF(ray.origin, set1));

    for (int i = 0; i < 512; i++) {
        const float radius = SDF(p, set1);
        if (insT > (fabsf(previousRadius) + fabsf(radius))) {
            t -= insT;
            omega = 1.0f;
            insT = previousRadius * omega * k;
            t += insT;
            p = fma3(ray.dir, t, ray.origin);
            continue;
        }
        if (fabsf(radius) < 1e-4f) break;
        if (t > tMax) points += 1; else points = 0;
        if (points >= 2) {
            t = tMax + 1e-3f;
            tMin = MAXDIST; tMax = MAXDIST;
            if (SearchSDF(c, fma3(ray.dir, t, ray.origin), invdir, tMin, tMax, set1)) {
                tMin += t; tMax += t;
                t = PTK_MAX(tMin, t);
                p = fma3(ray.dir, t, ray.origin);
                continue;
            } else {
                return;
            }
        }
        insT = radius * omega * k;
        t += insT;
        p = fma3(ray.dir, t, ray.origin);
        const float omegaSpeedFactor = PTK_MIN(PTK_DIV(radius, previousRadius), 0.99f);
        omega += omegaSpeed * (PTK_MIN(PTK_DIV(1.0f, 1.0f - omegaSpeedFactor), omegaMax) - omega);
        previousRadius = radius;
    }

    if (t < h.t) {
        h.t = t - 1e-3f;
        h.objectID = -1;
        if (!kShadow) {
            p = fma3(ray.dir, t, ray.origin);
            const float e = 1e-4f; /* shader.comp:721-730 */
            const float nx = SDF(mk3(p.x + e, p.y + 0.0f, p.z + 0.0f), set1) - SDF(mk3(p.x - e, p.y - 0.0f, p.z - 0.0f), set1);
            const float ny = SDF(mk3(p.x + 0.0f, p.y + e, p.z + 0.0f), set1) - SDF(mk3(p.x - 0.0f, p.y - e, p.z - 0.0f), set1);
            const float nz = SDF(mk3(p.x + 0.0f, p.y + 0.0f, p.z + e), set1) - SDF(mk3(p.x - 0.0f, p.y - 0.0f, p.z - e), set1);
            h.normal = normalize(mk3(nx, ny, nz));
            h.materialID = ::pt_sdfmaterial_dispatch(p.x, p.y, p.z, set1);
            h.lightID = -1.0f;
        }
    }
}
#endif /* PT_HAS_SDF */

#if PT_BVH
/* a primitive record of the pool's global-memory copy, as 16-byte loads through the read-only path */
template <class T>
PT_DEV T LoadRecord(const float* p) {
    alignas(16) T r;
    float4* d = reinterpret_cast<float4*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = __ldg(reinterpret_cast<const float4*>(p) + k);
    return r;
}
#endif

/* shader.comp:862-934 (kShadow = false) and 1121-1216 (kShadow = true): closest hit over the analytic primitives --
 * the reference's brute-force scan in type order, or (PT_BVH) the same search through the tree of pt_bvh.h */
PT_DEV void IntersectionAnalytic(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    const PtDevScene& sc = *c.sc;
    h.t = 1e5f;
    h.objectID = -1;
    if (!kShadow) {
        h.normal = mk3(0.0f, 0.0f, 0.0f);
        h.materialID = 0.0f;
        h.lightID = -1.0f;
    }
#if PT_BVH
    /* planes (unbounded) in order, then the tree over spheres, boxes and lenses -- the tie rule makes the visiting
     * order irrelevant -- then the cyclides in order like the reference: they have the highest indices, so the plain
     * `t < hit.t` is still the right rule, and they cannot live in the tree because the quartic solver reports
     * spurious roots anywhere along a ray that merely passes the bounding sphere (pt_bvh.cpp). */
    const int nSb = PT_N_SPHERES(c), nPb = PT_N_PLANES(c), nBb = PT_N_BOXES(c), nLb = PT_N_LENSES(c);
    for (int i = 0; i < nPb; i++) PlaneIntersection(ray, reinterpret_cast<const PtDevPlane*>(sc.pool + PT_OFF_PLANES(sc))[i], nSb + i, h, kShadow);
    const float* bvh = c.ubo + PT_BVH_UBO_OFF;
    const float* recs = bvh + PT_BVH_HEADER_FLOATS + PT_BVH_NODE_FLOATS * (nSb + nBb + nLb - 1); /* the pool's copy in global memory */
    const int offB = PT_OFF_BOXES(sc), offL = PT_OFF_LENSES(sc);
    pt_bvh_traverse(bvh, ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x, ray.dir.y, ray.dir.z, h.t, [&](int ref) {
        const int type = ref >> 16, i = ref & 0xffff;
        if (type == PT_BVH_SPHERE) {
            const PtDevSphere o = LoadRecord<PtDevSphere>(recs + 8 * i);
            SphereIntersection(ray, o, i, h, kShadow, true);
        } else if (type == PT_BVH_BOX) {
            if (nBb > 0) {
                const PtDevBox o = LoadRecord<PtDevBox>(recs + offB + 20 * i);
                if (BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) BoxIntersection(ray, o, nSb + nPb + i, h, kShadow, true);
            }
        } else {
            if (nLb > 0) {
                const PtDevLens o = LoadRecord<PtDevLens>(recs + offL + 20 * i);
                if (BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) {
                    int isOutside = 1;
                    LensIntersection(ray, o, nSb + nPb + nBb + i, h, isOutside, kShadow, true);
                }
            }
        }
    });
    int base = nSb + nPb + nBb + nLb;
#else
    int base = 0;
    const int nS = PT_N_SPHERES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nS; i++) SphereIntersection(ray, reinterpret_cast<const PtDevSphere*>(sc.pool)[i], base + i, h, kShadow);
    base += nS;
    const int nP = PT_N_PLANES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nP; i++) PlaneIntersection(ray, reinterpret_cast<const PtDevPlane*>(sc.pool + PT_OFF_PLANES(sc))[i], base + i, h, kShadow);
    base += nP;
    const int nB = PT_N_BOXES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nB; i++) {
        const PtDevBox& o = reinterpret_cast<const PtDevBox*>(sc.pool + PT_OFF_BOXES(sc))[i];
        if (!BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) continue;
        BoxIntersection(ray, o, base + i, h, kShadow);
    }
    base += nB;
    const int nL = PT_N_LENSES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nL; i++) {
        const PtDevLens& o = reinterpret_cast<const PtDevLens*>(sc.pool + PT_OFF_LENSES(sc))[i];
        if (!BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) continue;
        int isOutside = 1;
        LensIntersection(ray, o, base + i, h, isOutside, kShadow);
    }
    base += nL;
#endif
    const int nC = PT_N_CYCLIDES(c);
    for (int i = 0; i < nC; i++) {
        const PtDevCyclide& o = reinterpret_cast<const PtDevCyclide*>(sc.pool + PT_OFF_CYCLIDES(sc))[i];
        if (!BoundingSphere(ray, o.px, o.py, o.pz, o.brad)) continue;
        DupinCyclide(ray, o, base + i, h, kShadow);
    }
}

/* shader.comp:862-934 / 1121-1216 in one piece (used by the v1 driver) */
PT_DEV void Intersection(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    IntersectionAnalytic(c, ray, h, kShadow);
#if PT_HAS_SDF
    SphereTracing(c, ray, h, kShadow);
#endif
}

/* ---- sampling (shader.comp:976-1028, 1093-1119) ------------------------------------------------------------------ */
PT_DEV V3 SampleCosineDirectionHemisphere(V3 normal, unsigned& seed) {
    const float rx = RandomFloatPCG32(seed);
    const float ry = RandomFloatPCG32(seed);
    const float phi = 2.0f * PT_PI_F * ry;
    const float sinTheta = 2.0f * rx - 1.0f;
    const float cosTheta = PTK_SQRT(fmaf(-sinTheta, sinTheta, 1.0f));
    const V3 u = mk3(PTK_COS(phi) * cosTheta, PTK_SIN(phi) * cosTheta, sinTheta);
    return normalize(normal + u);
}
PT_DEV V3 SampleCosineUnitCone(unsigned& seed, float cosThetaMax) {
    const float rx = RandomFloatPCG32(seed);
    const float ry = RandomFloatPCG32(seed);
    const float cosAlphaMax = 2.0f * cosThetaMax * cosThetaMax - 1.0f;
    const float phi = 2.0f * PT_PI_F * ry;
    const float cosTheta = (1.0f - cosAlphaMax) * rx + cosAlphaMax;
    const float sinTheta = PTK_SQRT(fmaf(-cosTheta, cosTheta, 1.0f));
    return normalize(mk3(PTK_COS(phi) * sinTheta, PTK_SIN(phi) * sinTheta, cosTheta + 1.0f));
}
PT_DEV V3 ToWorld(V3 v, V3 n) {
    V3 b1 = mk3(0.0f, -1.0f, 0.0f);
    V3 b2 = mk3(-1.0f, 0.0f, 0.0f);
    if (n.z >= -0.9999999f) {
        const float a = PTK_DIV(1.0f, 1.0f + n.z);
        const float b = -n.x * n.y * a;
        b1 = mk3(1.0f - (n.x * n.x * a), b, -n.x);
        b2 = mk3(b, 1.0f - (n.y * n.y * a), -n.y);
    }
    return b1 * v.x + b2 * v.y + n * v.z;
}

/* ---- one path (shader.comp:1298-1407) --------------------------------------------------------------------------- */
PT_DEV V4 TracePath(const Ctx& c, V4 l, Ray ray, unsigned& seed) {
    const PtDevScene& sc = *c.sc;
    V4 radiance = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    V4 rayradiance = mk4(1.0f, 1.0f, 1.0f, 1.0f);
    float MISBRDFWeight = 1.0f;
    const int pathLength = c.pr->pathLength;
    for (int bounce = 0; bounce < pathLength; bounce++) {
        /* TraceRay, shader.comp:1345-1391 */
        Hit h;
        Intersection(c, ray, h, false);
        if (!(h.t < 1e5f)) break; /* miss: black environment */
        float temperature, luminosity;
        GetLightMix(c, h.lightID, temperature, luminosity);
        if (luminosity > 0.0f) { /* emitter hit terminates the path */
            const V4 e = Emit(l, PTK_MAX(temperature, 0.0f), PTK_MAX(luminosity, 0.0f));
            radiance = radiance + (e * rayradiance) * MISBRDFWeight;
            break;
        }
        float peak, sigma, invertf;
        GetMaterialMix(c, h.materialID, peak, sigma, invertf);
        const V4 brdf = EvaluateBRDF(l, peak, sigma, invertf);

        Ray outRay;
        outRay.origin = fma3(ray.dir, h.t, ray.origin);
        outRay.dir = SampleCosineDirectionHemisphere(h.normal, seed);
        const float BRDFpdf = PTK_DIV(dot(outRay.dir, h.normal), PT_PI_F);

        /* SampleLightSource, shader.comp:1298-1343 */
        if (sc.numLights > 0.0f) {
            const int randomLight = __float2int_rz(floorf(RandomFloatPCG32(seed) * sc.numLights));
            const PtDevLightSlot& ls = sc.lightSlots[randomLight < sc.nLightSlots ? randomLight : sc.nLightSlots - 1];
            const V3 toLight = mk3(ls.px - outRay.origin.x, ls.py - outRay.origin.y, ls.pz - outRay.origin.z);
            const float invLightDistance = PTK_DIV(1.0f, length(toLight));
            const V3 lightDir = toLight * invLightDistance;
            const float sinthetaMax = PTK_MIN(ls.boundingRadius * invLightDistance, 1.0f);
            const float costhetaMax = PTK_SQRT(1.0f - sinthetaMax * sinthetaMax);
            Ray shadowRay;
            shadowRay.origin = outRay.origin;
            shadowRay.dir = ToWorld(SampleCosineUnitCone(seed, costhetaMax), lightDir);
            float lightpdf = sc.invNumLights;
            lightpdf *= PTK_DIV(dot(shadowRay.dir, lightDir), PT_PI_F * (1.0f - costhetaMax * costhetaMax));
            MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + lightpdf * lightpdf);
            const float costheta = dot(shadowRay.dir, h.normal);
            const float deathProbability = 1.25f * PTK_MAX(MISBRDFWeight - 0.2f, 0.0f);
            if (costheta >= 0.0f) {
                if (RandomFloatPCG32(seed) > deathProbability) {
                    Hit sh;
                    Intersection(c, shadowRay, sh, true);
                    if (sh.objectID == ls.objectID) {
                        float lt, ll;
                        GetLightMix(c, ls.lightID, lt, ll);
                        const V4 rr = rayradiance * mulDiv4(brdf, costheta, lightpdf);
                        const V4 e = Emit(l, PTK_MAX(lt, 0.0f), PTK_MAX(ll, 0.0f));
                        radiance = radiance + (e * rr) * (1.0f - MISBRDFWeight);
                    }
                } else {
                    MISBRDFWeight = 1.0f;
                }
            }
        } else {
            MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + 0.0f * 0.0f);
        }

        const float costheta = dot(outRay.dir, h.normal);
        rayradiance = rayradiance * mulDiv4(brdf, costheta, BRDFpdf);
        const float mx = PTK_MAX(rayradiance.x, PTK_MAX(rayradiance.y, PTK_MAX(rayradiance.z, rayradiance.w)));
        const float rayProbability = PTK_MIN(PTK_MAX(mx, 0.0f), 0.99f);
        if (RandomFloatPCG32(seed) > rayProbability) break;
        rayradiance = rayradiance * PTK_DIV(1.0f, rayProbability);
        ray = outRay;
    }
    return radiance;
}

/* shader.comp:1409-1444: two refractions through the camera's BK7 lens */
PT_DEV void TracePathLens(const Ctx& c, float l, Ray& ray) {
    const PtDevLens& lens = c.pr->camLens;
#pragma unroll 1
    for (int i = 0; i < 2; i++) {
        Hit h;
        h.t = 1e6f;
        h.normal = mk3(0.0f, 0.0f, 0.0f);
        h.materialID = 0.0f; h.lightID = -1.0f; h.objectID = -1;
        int isOutside = 1;
        LensIntersection(ray, lens, 0, h, isOutside, false);
        float n1 = 1.0f, n2 = 1.0f;
        if (isOutside == 1) n2 = RefractiveIndexBK7Glass(l); else n1 = RefractiveIndexBK7Glass(l);
        const float n12 = PTK_DIV(n1, n2);
        l = l * n12;
        ray.origin = fma3(ray.dir, h.t, ray.origin);
        /* refract(I, N, eta), GLSL 4.50 8.5 */
        const float ndi = dot(h.normal, ray.dir);
        const float k = 1.0f - n12 * n12 * (1.0f - ndi * ndi);
        if (k < 0.0f) {
            ray.dir = mk3(0.0f, 0.0f, 0.0f);
        } else {
            const float f = n12 * ndi + PTK_SQRT(k);
            ray.dir = mk3(n12 * ray.dir.x - f * h.normal.x, n12 * ray.dir.y - f * h.normal.y, n12 * ray.dir.z - f * h.normal.z);
        }
    }
}

/* shader.comp:1446-1490: one spectral path sample -> CIE XYZ */
PT_DEV V3 Scene(const Ctx& c, unsigned xyx, unsigned xyy, float uvx, float uvy, int k) {
    const PtDevParams& pr = *c.pr;
    unsigned seed = (unsigned)(pr.firstSample + k); /* GenerateSeed, shader.comp:948-958 */
    PCG32(seed);
    seed += xyx + (unsigned)pr.width * xyy;

    const float j1 = RandomFloatPCG32(seed);
    const float j2 = RandomFloatPCG32(seed);
    uvx = uvx + PTK_DIV(2.0f * j1 - 0.5f, pr.resX);
    uvy = uvy + PTK_DIV(2.0f * j2 - 0.5f, pr.resY);
    uvx *= pr.sensorScale;
    uvy *= pr.sensorScale;
    const V3 camPos = mk3(pr.camPosX, pr.camPosY, pr.camPosZ);
    Ray ray;
    ray.origin = camPos + mulVM(mk3(uvx, uvy, 0.0f), pr.camM);
    /* SampleUniformUnitDisk, shader.comp:976-982 */
    const float rx = RandomFloatPCG32(seed);
    const float ry = RandomFloatPCG32(seed);
    const float phi = 2.0f * PT_PI_F * ry;
    const float dd = PTK_SQRT(rx);
    const float diskx = pr.halfAperture * (dd * PTK_COS(phi));
    const float disky = pr.halfAperture * (dd * PTK_SIN(phi));
    const V3 pointOnAperture = camPos + mulVM(mk3(diskx, disky, pr.apertureDist), pr.camM);
    ray.dir = normalize(pointOnAperture - ray.origin);

    const float r5 = RandomFloatPCG32(seed);
    const float l_h = 360.0f * (1.0f - r5) + 800.0f * r5; /* mix(360, 800, r) */
    TracePathLens(c, l_h, ray);
    const V4 l = SampleWavelengths(l_h);
    const V4 radiance = TracePath(c, l, ray, seed);
    const V3 w0 = WaveToXYZ(c, l.x), w1 = WaveToXYZ(c, l.y), w2 = WaveToXYZ(c, l.z), w3 = WaveToXYZ(c, l.w);
    V3 color;
    color.x = 0.0f + (radiance.x * w0.x + radiance.y * w1.x + radiance.z * w2.x + radiance.w * w3.x) * 330.0f * 0.25f;
    color.y = 0.0f + (radiance.x * w0.y + radiance.y * w1.y + radiance.z * w2.y + radiance.w * w3.y) * 330.0f * 0.25f;
    color.z = 0.0f + (radiance.x * w0.z + radiance.y * w1.z + radiance.z * w2.z + radiance.w * w3.z) * 330.0f * 0.25f;
    if ((color.x != color.x) || (color.y != color.y) || (color.z != color.z)) return mk3(0.0f, 0.0f, 0.0f);
    return color;
}

/* Rendering()'s tail + Accumulate() + imageStore, shader.comp:1492-1533 */
PT_DEV void StoreTexel(const PtDevParams& pr, float4* __restrict__ image, int gx, int gy, V3 outColor) {
    float4* texel = image + ((size_t)gx + (size_t)pr.width * (size_t)gy);
    if (pr.accumMode == 2) { /* raw sum for the sample-split path (pt_dispatch_sum) */
        float4 v = *texel;
        v.x += outColor.x; v.y += outColor.y; v.z += outColor.z;
        *texel = v;
        return;
    }
    outColor = div3(outColor, pr.spfFloat);
    outColor = outColor * pr.exposure;
    const float4 in = *texel;
    if (pr.accumMode == 1) {
        const float w = pr.accumWeight;
        outColor = mk3(((1.0f - w) * outColor.x) + (w * in.x), ((1.0f - w) * outColor.y) + (w * in.y),
                       ((1.0f - w) * outColor.z) + (w * in.z));
    } else {
        const float n = pr.accumN, nm1 = pr.accumNm1;
        outColor = mk3(PTK_DIV(nm1 * in.x + outColor.x, n), PTK_DIV(nm1 * in.y + outColor.y, n),
                       PTK_DIV(nm1 * in.z + outColor.z, n));
    }
    *texel = make_float4(outColor.x, outColor.y, outColor.z, 1.0f);
}

/* ---- driver v1: one thread = one pixel, nested sample / bounce loops (kept for A/B measurements, PT_SCHED=0) ----
 * Grid: 2-D tiles of 16x8 pixels per 128-thread block; each warp owns an 8x4 sub-tile. */
__device__ __forceinline__ void pt_render_body_v1(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                  float4* __restrict__ image, float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (gx >= pr.width || gy >= pr.height) return;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const unsigned xyy = (unsigned)pr.height - (unsigned)gy; /* shader.comp:1510 */
    const float uvx = PTK_DIV(2.0f * __uint2float_rn(xyx) - pr.resX, pr.resY);
    const float uvy = PTK_DIV(2.0f * __uint2float_rn(xyy) - pr.resY, pr.resY);

    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    const int spf = pr.samplesPerFrame;
#pragma unroll 1
    for (int k = 0; k < spf; k++) outColor = outColor + Scene(c, xyx, xyy, uvx, uvy, k);
    StoreTexel(pr, image, gx, gy, outColor);
}

/* ---- the path as four phases over explicit per-path state --------------------------------------------------------
 * v1 leaves most lanes idle: a warp waits for its longest path at every sample, the shadow ray runs under a
 * divergent branch, and with SDFs only the lanes whose ray entered a bounding box march, for 1..512 steps each
 * (ncu on v1: 17.3 of 32 lanes active on scene1, 7.5 of 32 on the mandelbulb scene).
 * Below, the SAME per-path arithmetic in the same order (so strict mode stays bit-exact) is cut into four phases
 * that read and write an explicit PathState:
 *   NEW    project the finished path to XYZ (Scene()'s tail, shader.comp:1477-1489), then camera ray + lens +
 *          wavelengths for the next sample index                                    (shader.comp:1446-1472)
 *   ISECT  brute-force primitives for the current ray, path OR shadow ray alike, then SearchSDF
 *                                                                                  (shader.comp:862-924, 1121-1205)
 *   SDF    ONE evaluation of the injected SDF(): sign probe, march step or one of the six normal probes -- every
 *          consumer of the distance function funnels through this single site       (shader.comp:779-860, 721-730)
 *   SHADE  emitter / BSDF sample / light sample with MIS / Russian roulette, or the verdict of a pending shadow
 *          ray (traced before the next path ray; LightSourceVisibilityCheck draws no random numbers)
 *                                                                                  (shader.comp:1298-1407)
 * Two drivers run these phases: v2 keeps the state in registers and lets each warp execute, per iteration, the
 * phase its lanes vote for; the wavefront pipeline keeps it in HBM (SoA) and runs one kernel per phase. */
enum { PT_ST_NEW = 0, PT_ST_ISECT = 1, PT_ST_SDF = 2, PT_ST_SHADE = 3, PT_ST_DONE = 4 };
enum { PT_SUB_SIGN = 0, PT_SUB_STEP = 1, PT_SUB_N0 = 2 }; /* N0..N5 = 2..7: +x -x +y -y +z -z */

struct PathState {
    Ray ray;            /* current path ray; while a shadow ray is pending ray.dir already holds the NEXT path direction */
    V4 l;               /* the wavelength bundle */
    V4 radiance, rayradiance;
    float MISBRDFWeight;
    unsigned seed;
    int bounce;
    bool isShadow;      /* the ray being traced is the shadow ray (origin = ray.origin, direction = shDir) */
    bool pathAlive;     /* whether the path continues after the pending shadow ray */
    bool pendingFinish; /* a finished path whose radiance is still to be projected to XYZ */
    V3 shDir;
    V4 shContrib;       /* added to radiance iff the shadow ray sees object shObj */
    int shObj;
    Hit h;
};

struct MarchState {     /* SphereTracing's locals, shader.comp:780-794 */
    float mt, insT, omega, previousRadius, tMax, ksign;
    float probe, nrm0, nrm1, nrm2;
    int points, iter, sub;
    unsigned set1;
};

PT_DEV void PathStateInit(PathState& ps) {
    const V3 z3 = mk3(0.0f, 0.0f, 0.0f);
    const V4 z4 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    ps.ray.origin = z3; ps.ray.dir = z3; ps.l = z4; ps.radiance = z4; ps.rayradiance = z4;
    ps.MISBRDFWeight = 1.0f; ps.seed = 0u; ps.bounce = 0;
    ps.isShadow = false; ps.pathAlive = false; ps.pendingFinish = false;
    ps.shDir = z3; ps.shContrib = z4; ps.shObj = 0;
    ps.h.t = 1e5f; ps.h.normal = z3; ps.h.materialID = 0.0f; ps.h.lightID = -1.0f; ps.h.objectID = -1;
}
PT_DEV void MarchStateInit(MarchState& ms) {
    ms.mt = 0.0f; ms.insT = 0.0f; ms.omega = 1.7f; ms.previousRadius = 0.0f; ms.tMax = 1e5f; ms.ksign = 0.0f;
    ms.probe = 0.0f; ms.nrm0 = 0.0f; ms.nrm1 = 0.0f; ms.nrm2 = 0.0f;
    ms.points = 0; ms.iter = 0; ms.sub = PT_SUB_SIGN; ms.set1 = 0u;
}

/* Scene()'s tail, shader.comp:1477-1489: radiance -> XYZ, NaNs dropped.  The four table look-ups run through one
 * rolled loop body (instruction-cache footprint); the sum keeps the shader's order
 * ((rad.x*W(l.x) + rad.y*W(l.y)) + rad.z*W(l.z)) + rad.w*W(l.w), the first product initialising it. */
PT_DEV V3 PathColor(const Ctx& c, const PathState& ps) {
    V4 l = ps.l, r = ps.radiance;
    V3 sum = mk3(0.0f, 0.0f, 0.0f);
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const V3 w = WaveToXYZ(c, l.x);
        sum = (i == 0) ? mk3(r.x * w.x, r.x * w.y, r.x * w.z) : mk3(sum.x + r.x * w.x, sum.y + r.x * w.y, sum.z + r.x * w.z);
        l = mk4(l.y, l.z, l.w, l.x);
        r = mk4(r.y, r.z, r.w, r.x);
    }
    V3 color;
    color.x = 0.0f + sum.x * 330.0f * 0.25f;
    color.y = 0.0f + sum.y * 330.0f * 0.25f;
    color.z = 0.0f + sum.z * 330.0f * 0.25f;
    if ((color.x != color.x) || (color.y != color.y) || (color.z != color.z)) color = mk3(0.0f, 0.0f, 0.0f);
    return color;
}

/* NEW: Scene() up to TracePath, shader.comp:1446-1472, for sample index pr.firstSample + k of pixel (xyx, xyy).
 * Returns the next phase (ISECT, or NEW again with pendingFinish when pathLength <= 0). */
PT_DEV int PhaseNew(const Ctx& c, PathState& ps, unsigned xyx, unsigned xyy, int k) {
    const PtDevParams& pr = *c.pr;
    const V3 camPos = mk3(pr.camPosX, pr.camPosY, pr.camPosZ);
    const float uvx0 = PTK_DIV(2.0f * __uint2float_rn(xyx) - pr.resX, pr.resY); /* shader.comp:1511 */
    const float uvy0 = PTK_DIV(2.0f * __uint2float_rn(xyy) - pr.resY, pr.resY);
    unsigned seed = (unsigned)(pr.firstSample + k); /* GenerateSeed, shader.comp:948-958 */
    PCG32(seed);
    seed += xyx + (unsigned)pr.width * xyy;
    const float j1 = RandomFloatPCG32(seed);
    const float j2 = RandomFloatPCG32(seed);
    float uvx = uvx0 + PTK_DIV(2.0f * j1 - 0.5f, pr.resX);
    float uvy = uvy0 + PTK_DIV(2.0f * j2 - 0.5f, pr.resY);
    uvx *= pr.sensorScale;
    uvy *= pr.sensorScale;
    Ray ray;
    ray.origin = camPos + mulVM(mk3(uvx, uvy, 0.0f), pr.camM);
    const float rx = RandomFloatPCG32(seed); /* SampleUniformUnitDisk, shader.comp:976-982 */
    const float ry = RandomFloatPCG32(seed);
    const float phi = 2.0f * PT_PI_F * ry;
    const float dd = PTK_SQRT(rx);
    const float diskx = pr.halfAperture * (dd * PTK_COS(phi));
    const float disky = pr.halfAperture * (dd * PTK_SIN(phi));
    const V3 pointOnAperture = camPos + mulVM(mk3(diskx, disky, pr.apertureDist), pr.camM);
    ray.dir = normalize(pointOnAperture - ray.origin);
    const float r5 = RandomFloatPCG32(seed);
    const float l_h = 360.0f * (1.0f - r5) + 800.0f * r5; /* mix(360, 800, r) */
    TracePathLens(c, l_h, ray);
    ps.ray = ray;
    ps.l = SampleWavelengths(l_h);
    ps.seed = seed;
    ps.radiance = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    ps.rayradiance = mk4(1.0f, 1.0f, 1.0f, 1.0f);
    ps.MISBRDFWeight = 1.0f;
    ps.bounce = 0;
    ps.isShadow = false;
    if (pr.pathLength > 0) return PT_ST_ISECT;
    ps.pendingFinish = true;
    return PT_ST_NEW;
}

/* ISECT: Intersection / LightSourceVisibilityCheck up to (and with) SphereTracing's prologue */
PT_DEV int PhaseIsect(const Ctx& c, PathState& ps, MarchState& ms) {
    Ray r;
    r.origin = ps.ray.origin;
    r.dir = ps.isShadow ? ps.shDir : ps.ray.dir;
    IntersectionAnalytic(c, r, ps.h, ps.isShadow);
#if PT_HAS_SDF
    const V3 invdir = mk3(PTK_DIV(1.0f, r.dir.x), PTK_DIV(1.0f, r.dir.y), PTK_DIV(1.0f, r.dir.z)); /* shader.comp:780-801 */
    float tMin = 1e5f;
    ms.tMax = 1e5f;
    if (SearchSDF(c, r.origin, invdir, tMin, ms.tMax, ms.set1)) {
        ms.mt = PTK_MAX(tMin, 1e-3f);
        ms.insT = 0.0f; ms.omega = 1.70f; ms.previousRadius = 0.0f; ms.points = 0; ms.iter = 0;
        ms.sub = PT_SUB_SIGN;
        return PT_ST_SDF;
    }
#endif
    return PT_ST_SHADE;
}

#if PT_HAS_SDF
/* The SDF phase in three pieces, so a driver can decide WHO evaluates the distance function at WHICH point:
 * SdfMarchPoint (where the march wants its next evaluation), SdfProbePoint (the j-th of the six normal probes around
 * p: +x -x +y -y +z -z, CalculateNumericalSDFNormals shader.comp:721-730) and SdfMarchConsume (SphereTracing's
 * bookkeeping for one evaluated distance, shader.comp:801-858). */
PT_DEV V3 SdfMarchPoint(const PathState& ps, const MarchState& ms) {
    if (ms.sub == PT_SUB_SIGN) return ps.ray.origin; /* k = sign(SDF(ray.origin)), shader.comp:801 */
    return fma3(ps.isShadow ? ps.shDir : ps.ray.dir, ms.mt, ps.ray.origin);
}
PT_DEV V3 SdfProbePoint(V3 p, int j) {
    const int axis = j >> 1;
    const float ex = (axis == 0) ? 1e-4f : 0.0f, ey = (axis == 1) ? 1e-4f : 0.0f, ez = (axis == 2) ? 1e-4f : 0.0f;
    return (j & 1) ? mk3(p.x - ex, p.y - ey, p.z - ez) : mk3(p.x + ex, p.y + ey, p.z + ez);
}
/* One evaluated distance d for a march in sub-state SIGN or STEP.  Returns SDF (more to do; ms.sub == PT_SUB_N0
 * means "converged on a path ray: the six normal probes are next") or SHADE. */
PT_DEV int SdfMarchConsume(const Ctx& c, PathState& ps, MarchState& ms, float d) {
    if (ms.sub == PT_SUB_SIGN) {
        ms.ksign = gsign(d);
        ms.sub = PT_SUB_STEP;
        return PT_ST_SDF;
    }
    /* one iteration of the loop at shader.comp:803-849 */
    const V3 dir = ps.isShadow ? ps.shDir : ps.ray.dir;
    const float radius = d;
    bool finished = false, nohit = false;
    if (ms.insT > (fabsf(ms.previousRadius) + fabsf(radius))) {
        ms.mt -= ms.insT;
        ms.omega = 1.0f;
        ms.insT = ms.previousRadius * ms.omega * ms.ksign;
        ms.mt += ms.insT;
    } else if (fabsf(radius) < 1e-4f) {
        finished = true;
    } else {
        if (ms.mt > ms.tMax) ms.points += 1; else ms.points = 0;
        if (ms.points >= 2) {
            ms.mt = ms.tMax + 1e-3f;
            float tMin = 1e5f;
            ms.tMax = 1e5f;
            const V3 invdir = mk3(PTK_DIV(1.0f, dir.x), PTK_DIV(1.0f, dir.y), PTK_DIV(1.0f, dir.z));
            if (SearchSDF(c, fma3(dir, ms.mt, ps.ray.origin), invdir, tMin, ms.tMax, ms.set1)) {
                tMin += ms.mt; ms.tMax += ms.mt;
                ms.mt = PTK_MAX(tMin, ms.mt);
            } else {
                nohit = true;
            }
        } else {
            ms.insT = radius * ms.omega * ms.ksign;
            ms.mt += ms.insT;
            const float omegaSpeedFactor = PTK_MIN(PTK_DIV(radius, ms.previousRadius), 0.99f);
            ms.omega += 0.20f * (PTK_MIN(PTK_DIV(1.0f, 1.0f - omegaSpeedFactor), 1.70f) - ms.omega);
            ms.previousRadius = radius;
        }
    }
    if (!finished && !nohit) {
        ms.iter++;
        if (ms.iter >= 512) finished = true; /* falls through to "hit" unconverged: SURVEY App. C-9 */
    }
    if (nohit) return PT_ST_SHADE;
    if (finished) { /* shader.comp:851-858 */
        if (ms.mt < ps.h.t) {
            ps.h.t = ms.mt - 1e-3f;
            ps.h.objectID = -1;
            if (ps.isShadow) return PT_ST_SHADE;
            ms.sub = PT_SUB_N0;
            return PT_ST_SDF;
        }
        return PT_ST_SHADE;
    }
    return PT_ST_SDF;
}
/* The hit of a converged path ray from its six probe values (shader.comp:853-857) */
PT_DEV void SdfFinishHit(PathState& ps, const MarchState& ms, float e0, float e1, float e2, float e3, float e4, float e5) {
    const V3 p = fma3(ps.ray.dir, ms.mt, ps.ray.origin);
    ps.h.normal = normalize(mk3(e0 - e1, e2 - e3, e4 - e5));
    ps.h.materialID = ::pt_sdfmaterial_dispatch(p.x, p.y, p.z, ms.set1);
    ps.h.lightID = -1.0f;
}

/* SDF, one lane on its own: one SDF() evaluation and the bookkeeping around it.  Returns SDF (more to do) or SHADE. */
PT_DEV int PhaseSdfEval(const Ctx& c, PathState& ps, MarchState& ms) {
    const V3 pm = SdfMarchPoint(ps, ms);
    const V3 pe = (ms.sub >= PT_SUB_N0) ? SdfProbePoint(pm, ms.sub - PT_SUB_N0) : pm;
    const float d = SDF(pe, ms.set1); /* the one SDF() site of the kernel */
    if (ms.sub < PT_SUB_N0) return SdfMarchConsume(c, ps, ms, d);
    const int j = ms.sub - PT_SUB_N0;
    if ((j & 1) == 0) {
        ms.probe = d;
    } else {
        const float g = ms.probe - d;
        if (j == 1) ms.nrm0 = g; else if (j == 3) ms.nrm1 = g; else ms.nrm2 = g;
    }
    ms.sub++;
    if (j == 5) {
        const V3 p = fma3(ps.ray.dir, ms.mt, ps.ray.origin);
        ps.h.normal = normalize(mk3(ms.nrm0, ms.nrm1, ms.nrm2));
        ps.h.materialID = ::pt_sdfmaterial_dispatch(p.x, p.y, p.z, ms.set1);
        ps.h.lightID = -1.0f;
        return PT_ST_SHADE;
    }
    return PT_ST_SDF;
}

/* SDF, the whole warp at once (v2 driver): per round every lane evaluates SDF() at most once -- marching lanes at
 * their own next point, and, when some lanes have converged on a path ray, lanes 0..6m-1 at the six normal probes
 * of the first m <= 5 of them (a marching lane drafted as a prober just advances one round later).  A converged
 * ray's normal thus costs one round with six lanes busy instead of six rounds with one, and runs concurrently with
 * the other lanes' marching.  Same evaluations, same arithmetic: bit-exact. */
PT_DEV int PhaseSdfWarp(const Ctx& c, PathState& ps, MarchState& ms, int st, int rounds) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll 1
    for (int rep = 0; rep < rounds; rep++) {
        const bool inSdf = (st == PT_ST_SDF);
        const bool needN = inSdf && (ms.sub == PT_SUB_N0);
        const unsigned needMask = __ballot_sync(0xffffffffu, needN);
        if (__ballot_sync(0xffffffffu, inSdf) == 0u) break;
        const int m = min(__popc(needMask), 5);
        const bool prober = (int)lane < 6 * m;
        V3 pe = ps.ray.origin;
        unsigned set = ms.set1;
        bool doEval = false;
        if (m > 0) { /* warp-uniform */
            const V3 ph = fma3(ps.ray.dir, ms.mt, ps.ray.origin); /* meaningful on the converged lanes */
            const int slot = (int)lane / 6, probe = (int)lane - 6 * slot;
            const int src = prober ? (int)__fns(needMask, 0u, slot + 1) : 0;
            const float qx = __shfl_sync(0xffffffffu, ph.x, src), qy = __shfl_sync(0xffffffffu, ph.y, src),
                        qz = __shfl_sync(0xffffffffu, ph.z, src);
            const unsigned qs = __shfl_sync(0xffffffffu, ms.set1, src);
            if (prober) {
                pe = SdfProbePoint(mk3(qx, qy, qz), probe);
                set = qs;
                doEval = true;
            }
        }
        const bool marching = inSdf && (ms.sub < PT_SUB_N0) && !prober;
        if (marching) {
            pe = SdfMarchPoint(ps, ms);
            set = ms.set1;
            doEval = true;
        }
        float d = 0.0f;
        if (doEval) d = SDF(pe, set); /* the one SDF() site of the kernel */
        if (marching) st = SdfMarchConsume(c, ps, ms, d);
        if (m > 0) {
            const int nr = __popc(needMask & ((1u << lane) - 1u));
            const bool served = needN && (nr < m);
            const int b0 = served ? 6 * nr : 0;
            const float e0 = __shfl_sync(0xffffffffu, d, b0), e1 = __shfl_sync(0xffffffffu, d, b0 + 1),
                        e2 = __shfl_sync(0xffffffffu, d, b0 + 2), e3 = __shfl_sync(0xffffffffu, d, b0 + 3),
                        e4 = __shfl_sync(0xffffffffu, d, b0 + 4), e5 = __shfl_sync(0xffffffffu, d, b0 + 5);
            if (served) {
                SdfFinishHit(ps, ms, e0, e1, e2, e3, e4, e5);
                st = PT_ST_SHADE;
            }
        }
    }
    return st;
}
#endif

/* SHADE: TraceRay after Intersection (shader.comp:1352-1390) with SampleLightSource (1298-1343), or the verdict of
 * the pending shadow ray (1218-1222, 1328-1334).  Returns ISECT (another ray to trace) or NEW (path finished). */
PT_DEV int PhaseShade(const Ctx& c, PathState& ps) {
    const PtDevScene& sc = *c.sc;
    bool done = false;
    int next = PT_ST_ISECT;
    if (ps.isShadow) {
        if (ps.h.objectID == ps.shObj) ps.radiance = ps.radiance + ps.shContrib;
        ps.isShadow = false;
        if (!ps.pathAlive) done = true;
    } else if (!(ps.h.t < 1e5f)) { /* miss: black environment */
        done = true;
    } else {
        float emitT, emitL; /* the light Emit() is evaluated for: the one hit, or the one sampled */
        GetLightMix(c, ps.h.lightID, emitT, emitL);
        const bool emitterHit = emitL > 0.0f; /* terminates the path, shader.comp:1359-1364 */
        bool needShadow = false, alive = false;
        V4 rr = ps.rayradiance;
        float emitScale = ps.MISBRDFWeight;
        V3 outOrigin = ps.ray.origin, outDir = ps.ray.dir;
        if (!emitterHit) {
            float peak, sigma, invertf;
            GetMaterialMix(c, ps.h.materialID, peak, sigma, invertf);
            const V4 brdf = EvaluateBRDF(ps.l, peak, sigma, invertf);
            const V3 n = ps.h.normal;
            outOrigin = fma3(ps.ray.dir, ps.h.t, ps.ray.origin);
            outDir = SampleCosineDirectionHemisphere(n, ps.seed);
            const float BRDFpdf = PTK_DIV(dot(outDir, n), PT_PI_F);
            if (sc.numLights > 0.0f) { /* SampleLightSource, shader.comp:1298-1343 */
                const int randomLight = __float2int_rz(floorf(RandomFloatPCG32(ps.seed) * sc.numLights));
                const PtDevLightSlot& ls = sc.lightSlots[randomLight < sc.nLightSlots ? randomLight : sc.nLightSlots - 1];
                const V3 toLight = mk3(ls.px - outOrigin.x, ls.py - outOrigin.y, ls.pz - outOrigin.z);
                const float invLightDistance = PTK_DIV(1.0f, length(toLight));
                const V3 lightDir = toLight * invLightDistance;
                const float sinthetaMax = PTK_MIN(ls.boundingRadius * invLightDistance, 1.0f);
                const float costhetaMax = PTK_SQRT(1.0f - sinthetaMax * sinthetaMax);
                const V3 sdir = ToWorld(SampleCosineUnitCone(ps.seed, costhetaMax), lightDir);
                float lightpdf = sc.invNumLights;
                lightpdf *= PTK_DIV(dot(sdir, lightDir), PT_PI_F * (1.0f - costhetaMax * costhetaMax));
                ps.MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + lightpdf * lightpdf);
                const float costheta = dot(sdir, n);
                const float deathProbability = 1.25f * PTK_MAX(ps.MISBRDFWeight - 0.2f, 0.0f);
                if (costheta >= 0.0f) {
                    if (RandomFloatPCG32(ps.seed) > deathProbability) {
                        GetLightMix(c, ls.lightID, emitT, emitL);
                        rr = ps.rayradiance * mulDiv4(brdf, costheta, lightpdf);
                        emitScale = 1.0f - ps.MISBRDFWeight;
                        ps.shDir = sdir;
                        ps.shObj = ls.objectID;
                        needShadow = true;
                    } else {
                        ps.MISBRDFWeight = 1.0f;
                    }
                }
            } else {
                ps.MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + 0.0f * 0.0f);
            }
            const float costheta = dot(outDir, n);
            ps.rayradiance = ps.rayradiance * mulDiv4(brdf, costheta, BRDFpdf);
            const V4 t4 = ps.rayradiance;
            const float mx = PTK_MAX(t4.x, PTK_MAX(t4.y, PTK_MAX(t4.z, t4.w)));
            const float rayProbability = PTK_MIN(PTK_MAX(mx, 0.0f), 0.99f);
            alive = !(RandomFloatPCG32(ps.seed) > rayProbability);
            if (alive) ps.rayradiance = ps.rayradiance * PTK_DIV(1.0f, rayProbability);
            ps.bounce++;
            if (ps.bounce >= c.pr->pathLength) alive = false; /* TracePath's loop bound, shader.comp:1400 */
        }
        if (emitterHit || needShadow) { /* the one Emit() site: (Emit * rr) * scale in both uses */
            const V4 e = Emit(ps.l, PTK_MAX(emitT, 0.0f), PTK_MAX(emitL, 0.0f));
            const V4 contrib = (e * rr) * emitScale;
            if (emitterHit) ps.radiance = ps.radiance + contrib; else ps.shContrib = contrib;
        }
        ps.ray.origin = outOrigin;
        ps.ray.dir = outDir; /* next path direction; a pending shadow ray travels along shDir */
        if (emitterHit) {
            done = true;
        } else if (needShadow) {
            ps.isShadow = true;
            ps.pathAlive = alive;
        } else if (!alive) {
            done = true;
        }
    }
    if (done) {
        ps.pendingFinish = true;
        next = PT_ST_NEW;
    }
    return next;
}

/* The two outcomes of a traced ray that need no shading work -- the verdict of a shadow ray and a path ray that left
 * the scene -- exactly as PhaseShade handles them (shader.comp:1218-1222,1328-1334 and 1387-1389).  The in-warp
 * drivers apply this right after the intersection / march phases, so the SHADE phase only ever runs real shading.
 * Returns the next phase, or SHADE when the hit needs shading. */
PT_DEV int PhaseTrivial(PathState& ps) {
    if (ps.isShadow) {
        if (ps.h.objectID == ps.shObj) ps.radiance = ps.radiance + ps.shContrib;
        ps.isShadow = false;
        if (ps.pathAlive) return PT_ST_ISECT;
        ps.pendingFinish = true;
        return PT_ST_NEW;
    }
    if (!(ps.h.t < 1e5f)) {
        ps.pendingFinish = true;
        return PT_ST_NEW;
    }
    return PT_ST_SHADE;
}

/* ---- driver v2: in-warp scheduled state machine, all state in registers ------------------------------------------
 * Every lane runs the phases above for its own pixel; per iteration the warp executes ONE phase, chosen by ballot.
 * A lane that finishes a path refills itself with its pixel's next sample index instead of idling, so the
 * expensive phases run with more lanes populated than in v1.  No queues, no HBM traffic. */
#ifdef PT_STATS
/* scheduling statistics (debug builds only, env PT_STATS=1): per phase, [2p] = executions, [2p+1] = lanes served */
} /* namespace */
extern "C" __device__ unsigned long long pt_stats[16];
namespace PT_KERNEL_NS {
#define PT_STAT(p, mask) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&pt_stats[2 * (p)], 1ull); atomicAdd(&pt_stats[2 * (p) + 1], (unsigned long long)__popc(mask)); } } while (0)
#else
#define PT_STAT(p, mask) do { } while (0)
#endif
#ifndef PT_SDF_REPS
#define PT_SDF_REPS 16
#endif
#ifndef PT_FEED_T
#define PT_FEED_T 8
#endif
#ifndef PT_COOP_NORMALS
#define PT_COOP_NORMALS 0 /* measured slower than one lane per ray on every SDF workload: profiles/r01_coop */
#endif
#ifndef PT_SDF_MIN
#define PT_SDF_MIN 0  /* lanes that must be waiting in the SDF phase before it runs while a feeder still has work */
#endif
#ifndef PT_SDF_EXIT
#define PT_SDF_EXIT 0 /* leave the SDF phase early (checked every 4 evaluations) once fewer lanes than this still march */
#endif

__device__ __forceinline__ void pt_render_body_v2(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                  float4* __restrict__ image, float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const unsigned xyy = (unsigned)pr.height - (unsigned)gy; /* shader.comp:1510 */
    const int spf = pr.samplesPerFrame;

    int st = (inRange && spf > 0) ? PT_ST_NEW : PT_ST_DONE;
    int k = 0;
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps;
    PathStateInit(ps);
    MarchState ms;
    MarchStateInit(ms);

    for (;;) {
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
#if PT_HAS_SDF
        const unsigned bSdf = __ballot_sync(0xffffffffu, st == PT_ST_SDF);
#else
        const unsigned bSdf = 0u;
#endif
        if ((bNew | bIs | bSdf | bSh) == 0u) break;
        /* Phase selection.  Executing a phase costs the same whatever its population.  The fullest of the three
         * "feeder" phases wins (ties go to the later pipeline stage); lanes waiting in the SDF phase take their
         * PT_SDF_REPS evaluations when every feeder has fewer than PT_FEED_T lanes waiting.  Without SDFs this
         * is "the phase most lanes are waiting for". */
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
#if PT_HAS_SDF
        /* an SDF execution costs an order of magnitude more than a feeder phase whatever its population: with
         * PT_SDF_MIN > 0 a thin feeder runs first (it may send more lanes marching) unless enough lanes already wait */
        if (bSdf != 0u && (best == 0 || (best < PT_FEED_T && __popc(bSdf) >= PT_SDF_MIN))) phase = PT_ST_SDF;
#endif
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? bSdf : bSh)));
#ifdef PT_STATS
        if (phase == PT_ST_SDF && (threadIdx.x & 31) == 0) atomicAdd(&pt_stats[8 + ((__popc(bSdf) - 1) >> 2)], 1ull); /* population histogram, buckets of 4 */
#endif
        if (phase == PT_ST_NEW) {
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) {
                    outColor = outColor + PathColor(c, ps);
                    ps.pendingFinish = false;
                }
                if (k < spf) {
                    st = PhaseNew(c, ps, xyx, xyy, k);
                    k++;
                } else {
                    st = PT_ST_DONE;
                }
            }
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
            }
        }
#if PT_HAS_SDF
        else if (phase == PT_ST_SDF) {
#if PT_COOP_NORMALS
            st = PhaseSdfWarp(c, ps, ms, st, PT_SDF_REPS);
#else
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (st == PT_ST_SDF) st = PhaseSdfEval(c, ps, ms);
#if PT_SDF_EXIT > 0
                if ((rep & 3) == 3 && __popc(__ballot_sync(0xffffffffu, st == PT_ST_SDF)) < PT_SDF_EXIT) break;
#endif
            }
#endif
            if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
        }
#endif
        else {
            if (st == PT_ST_SHADE) st = PhaseShade(c, ps);
        }
    }
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}


/* ---- driver v2s (PT_SCHED=5): v2 with in-warp sample stealing ----------------------------------------------------
 * In v2 a lane owns one pixel and runs that pixel's samples one after the other, so a warp lives as long as its most
 * expensive pixel: on the SDF scenes a few lanes of a tile march the fractal through all their samples while the
 * others finished long ago (oracle cost maps, profiles/r01_steal: the mean lane carries 0.35 (menger) / 0.64
 * (mandelbulb) / 0.72 (terrain) of the work of the busiest lane of its warp at 16 samples per pixel).
 * Here the warp owns the 32 x S samples of its 8x4 tile as a pool of work items (item = 32 * sample + pixel, so the
 * first 32 items are v2's initial state and primary rays start coherent); a lane that finishes a path claims the
 * next unclaimed item, whichever pixel it belongs to (claims happen in the NEW phase, which the warp executes
 * converged: rank by ballot, no atomics).  A finished sample's XYZ goes to a per-warp table in shared memory
 * (S x 32 x 3 floats); when the pool is empty and all lanes idle, lane p adds pixel p's S entries IN SAMPLE ORDER --
 * the same additions in the same order as v1 / v2, so strict mode stays bit-exact -- and the next round of S
 * samples starts.  Scene() depends only on (pixel, sample index), never on the lane that runs it. */
#ifndef PT_STEAL_S
#define PT_STEAL_S 16
#endif
/* PT_STEAL_S == 0 (fast mode only): ONE round with all samplesPerFrame samples of the dispatch in the pool, and a
 * finished sample is added straight to its pixel's running sum in shared memory (red.shared.add.f32) -- no table,
 * hence no bound on the pool, at the price of a summation order that follows the schedule instead of the sample
 * index (the schedule of a warp is a pure function of its inputs, so renders still repeat bit for bit in practice). */
#if PT_STEAL_S == 0
#define PT_STEAL_WORDS (3 * 32) /* per warp: the running XYZ sums of the tile's pixels */
#else
#define PT_STEAL_WORDS (3 * PT_STEAL_S * 32) /* per warp */
#endif
static_assert(PT_STEAL_S >= 0 && PT_STEAL_S <= 20, "PT_STEAL_S: the per-sample table must fit the 48 KB of static shared memory next to the uniform block");
enum { PT_ST_IDLE = 5 };

/* PT_MPARK (v2s, SDF scenes): rays that march do not wait in a lane.  The march lengths are heavy-tailed (1 .. 512
 * evaluations), so whatever batch of lanes enters the SDF phase together, its later executions serve the few long
 * rays only (61-73 % of v2's SDF executions find 1-4 lanes, profiles/r01_sdfsched).  With PT_MPARK a lane whose ray
 * must march -- after ISECT found a bounding box, and again after every SDF execution that did not finish it -- writes
 * the whole path (49 words: PathState, MarchState, the item it belongs to) to a per-warp stack in shared memory and is
 * free for other work; in the NEW phase free lanes take parked paths back, a batch at a time: as soon as
 * PT_MPARK_MIN paths wait and as many lanes are free (or the pool is empty).  Every SDF execution thus starts with a
 * batch, and the feeder phases carry no lanes that only wait for it.  A path is the same arithmetic whichever lane
 * holds it (strict mode stays bit-exact); with a full stack the ray simply stays in its lane as before. */
#ifndef PT_MPARK
#define PT_MPARK 0
#endif
#ifndef PT_MPARK_CAP
#define PT_MPARK_CAP 20
#endif
#ifndef PT_MPARK_MIN
#define PT_MPARK_MIN 12
#endif
#define PT_MPARK_FIELDS 49
#if PT_MPARK && PT_HAS_SDF
#define PT_MPARK_WORDS (PT_MPARK_FIELDS * PT_MPARK_CAP) /* per warp, [field][position]: a batch of lanes hits distinct banks */
#define PT_MPARK_XFER(X)                                                                                       \
    X(0, ps.ray.origin.x) X(1, ps.ray.origin.y) X(2, ps.ray.origin.z) X(3, ps.ray.dir.x) X(4, ps.ray.dir.y)    \
    X(5, ps.ray.dir.z) X(6, ps.l.x) X(7, ps.l.y) X(8, ps.l.z) X(9, ps.l.w) X(10, ps.radiance.x)                \
    X(11, ps.radiance.y) X(12, ps.radiance.z) X(13, ps.radiance.w) X(14, ps.rayradiance.x)                    \
    X(15, ps.rayradiance.y) X(16, ps.rayradiance.z) X(17, ps.rayradiance.w) X(18, ps.MISBRDFWeight)           \
    X(19, ps.shDir.x) X(20, ps.shDir.y) X(21, ps.shDir.z) X(22, ps.shContrib.x) X(23, ps.shContrib.y)          \
    X(24, ps.shContrib.z) X(25, ps.shContrib.w) X(26, ps.h.t) X(27, ps.h.normal.x) X(28, ps.h.normal.y)        \
    X(29, ps.h.normal.z) X(30, ps.h.materialID) X(31, ps.h.lightID) X(32, ms.mt) X(33, ms.tMax)               \
    X(40, ms.insT) X(41, ms.omega) X(42, ms.previousRadius) X(43, ms.ksign) X(44, ms.probe) X(45, ms.nrm0)     \
    X(46, ms.nrm1) X(47, ms.nrm2)
#else
#define PT_MPARK_WORDS 0
#endif

__device__ __forceinline__ void pt_render_body_v2s(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_colAll) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = blockIdx.y * 8 + (warp >> 1) * 4;
    const int gx = tileX + (lane & 7), gy = tileY + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    float* s_col = s_colAll + warp * (PT_STEAL_WORDS + PT_MPARK_WORDS);
#if PT_MPARK && PT_HAS_SDF
    float* s_park = s_col + PT_STEAL_WORDS;
    int parked = 0;                                      /* paths on the warp's stack (warp-uniform) */
#endif

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const int spf = pr.samplesPerFrame;
    const bool warpLive = (tileX < pr.width) && (tileY < pr.height) && (spf > 0); /* warp-uniform */
    int roundBase = 0;                                   /* sample index of the round's first sample */
#if PT_STEAL_S == 0
    int roundN = spf;
    s_col[lane] = 0.0f; s_col[32 + lane] = 0.0f; s_col[64 + lane] = 0.0f;
    __syncwarp();
#else
    int roundN = spf < PT_STEAL_S ? spf : PT_STEAL_S;    /* samples per pixel in this round */
#endif
    int next = 0;                                        /* first unclaimed item of the round (warp-uniform) */
    int item = 0;                                        /* the item this lane is working on */

    int st = warpLive ? PT_ST_NEW : PT_ST_DONE;
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps;
    PathStateInit(ps);
    MarchState ms;
    MarchStateInit(ms);

    for (;;) {
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
#if PT_HAS_SDF
        const unsigned bSdf = __ballot_sync(0xffffffffu, st == PT_ST_SDF);
#else
        const unsigned bSdf = 0u;
#endif
        if ((bNew | bIs | bSdf | bSh) == 0u) {
            if (!warpLive) break;
            /* end of a round: every item is done.  Lane p sums pixel p's samples in index order. */
            __syncwarp();
#if PT_STEAL_S == 0
            outColor = mk3(s_col[lane], s_col[32 + lane], s_col[64 + lane]);
            break;
#else
            if (inRange) {
#pragma unroll 1
                for (int kk = 0; kk < roundN; kk++) {
                    const float* e = s_col + (3 * kk) * 32 + lane;
                    outColor = outColor + mk3(e[0], e[32], e[64]);
                }
            }
            __syncwarp();
            roundBase += roundN;
            if (roundBase >= spf) break;
            roundN = (spf - roundBase) < PT_STEAL_S ? (spf - roundBase) : PT_STEAL_S;
            next = 0;
            st = PT_ST_NEW;
            continue;
#endif
        }
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
#if PT_HAS_SDF
        if (bSdf != 0u && (best == 0 || (best < PT_FEED_T && __popc(bSdf) >= PT_SDF_MIN))) phase = PT_ST_SDF;
#endif
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? bSdf : bSh)));
        bool entered = false; /* this lane's ray found an SDF bounding box in this iteration's ISECT phase */
        if (phase == PT_ST_NEW) {
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) { /* the sample this lane just finished: item -> (pixel, sample of the round) */
                    const V3 col = PathColor(c, ps);
#if PT_STEAL_S == 0
                    float* e = s_col + (item & 31);
                    atomicAdd(e, col.x); atomicAdd(e + 32, col.y); atomicAdd(e + 64, col.z);
#else
                    float* e = s_col + (3 * (item >> 5)) * 32 + (item & 31);
                    e[0] = col.x; e[32] = col.y; e[64] = col.z;
#endif
                    ps.pendingFinish = false;
                }
            }
            int rank = __popc(bNew & ((1u << lane) - 1u)), nfree = __popc(bNew);
#if PT_MPARK && PT_HAS_SDF
            /* free lanes take parked paths back, the whole batch at once */
            const int npop = parked < nfree ? parked : nfree; /* a batch needs parked paths AND free lanes */
            if (npop > 0 && (npop >= PT_MPARK_MIN || next >= 32 * roundN)) {
                if (st == PT_ST_NEW && rank < npop) {
                    const float* e = s_park + (parked - 1 - rank);
#define PT_MPARK_LD(i, f) f = e[(i) * PT_MPARK_CAP];
                    PT_MPARK_XFER(PT_MPARK_LD)
#undef PT_MPARK_LD
                    ps.seed = __float_as_uint(e[34 * PT_MPARK_CAP]);
                    const unsigned pk = __float_as_uint(e[35 * PT_MPARK_CAP]); /* bounce | isShadow << 30 | pathAlive << 31 */
                    ps.bounce = (int)(pk & 0x3fffffffu); ps.isShadow = ((pk >> 30) & 1u) != 0u; ps.pathAlive = (pk >> 31) != 0u;
                    ps.shObj = __float_as_int(e[36 * PT_MPARK_CAP]);
                    ps.h.objectID = __float_as_int(e[37 * PT_MPARK_CAP]);
                    ms.set1 = __float_as_uint(e[38 * PT_MPARK_CAP]);
                    item = __float_as_int(e[39 * PT_MPARK_CAP]);
                    const unsigned mk = __float_as_uint(e[48 * PT_MPARK_CAP]); /* points | iter << 12 | sub << 24 */
                    ms.points = (int)(mk & 0xfffu); ms.iter = (int)((mk >> 12) & 0xfffu); ms.sub = (int)(mk >> 24);
                    st = PT_ST_SDF;
                }
                parked -= npop;
                rank -= npop; /* the lanes served above are no longer NEW */
                nfree -= npop;
                __syncwarp();
            }
#endif
            if (st == PT_ST_NEW) {
                item = next + rank;
                if (item < 32 * roundN) {
                    const int q = item & 31;
                    const int qx = tileX + (q & 7), qy = tileY + (q >> 3);
                    if ((qx < pr.width) && (qy < pr.height)) /* else: a pixel beyond the image edge; claim again */
                        st = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, roundBase + (item >> 5));
                } else {
                    st = PT_ST_IDLE;
                }
            }
            next += nfree;
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
                entered = (st == PT_ST_SDF);
            }
#if !(PT_MPARK && PT_HAS_SDF)
            (void)entered;
#endif
        }
#if PT_HAS_SDF
        else if (phase == PT_ST_SDF) {
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (st == PT_ST_SDF) st = PhaseSdfEval(c, ps, ms);
#if PT_SDF_EXIT > 0
                if ((rep & 3) == 3 && __popc(__ballot_sync(0xffffffffu, st == PT_ST_SDF)) < PT_SDF_EXIT) break;
#endif
            }
            if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
        }
#endif
        else {
            if (st == PT_ST_SHADE) st = PhaseShade(c, ps);
        }
#if PT_MPARK && PT_HAS_SDF
        /* the one parking site: rays that entered a bounding box in this ISECT phase, or that this SDF execution did not
         * finish, go to the stack (as many as fit); their lanes are free again */
        if (phase == PT_ST_ISECT || phase == PT_ST_SDF) {
            const bool doPark = (phase == PT_ST_ISECT) ? entered : (st == PT_ST_SDF);
            const unsigned bEnt = __ballot_sync(0xffffffffu, doPark);
            if (bEnt != 0u) {
                const int pos = parked + __popc(bEnt & ((1u << lane) - 1u));
                if (doPark && pos < PT_MPARK_CAP) {
                    float* e = s_park + pos;
#define PT_MPARK_ST(i, f) e[(i) * PT_MPARK_CAP] = f;
                    PT_MPARK_XFER(PT_MPARK_ST)
#undef PT_MPARK_ST
                    e[34 * PT_MPARK_CAP] = __uint_as_float(ps.seed);
                    e[35 * PT_MPARK_CAP] = __uint_as_float((unsigned)ps.bounce | ((unsigned)ps.isShadow << 30) | ((unsigned)ps.pathAlive << 31));
                    e[36 * PT_MPARK_CAP] = __int_as_float(ps.shObj);
                    e[37 * PT_MPARK_CAP] = __int_as_float(ps.h.objectID);
                    e[38 * PT_MPARK_CAP] = __uint_as_float(ms.set1);
                    e[39 * PT_MPARK_CAP] = __int_as_float(item);
                    e[48 * PT_MPARK_CAP] = __uint_as_float((unsigned)ms.points | ((unsigned)ms.iter << 12) | ((unsigned)ms.sub << 24));
                    st = PT_ST_NEW; /* nothing pending: the NEW phase hands this lane a parked path or a new item */
                }
                const int room = PT_MPARK_CAP - parked, n = __popc(bEnt);
                parked += n < room ? n : room;
                __syncwarp();
            }
        }
#endif
    }
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}


/* ---- driver v2sp (PT_SCHED=6, fast mode): v2s with persistent warps streaming over tiles ---------------------------
 * v2s still drains its pool at the end of every 8x4 tile: the last paths of a tile run alone (an expensive path on the
 * menger scene is worth a tenth of a whole tile at 64 samples per pixel), and a small samplesPerFrame leaves nothing
 * to steal.  Here a warp does not belong to a tile.  The grid is persistent (SMs x resident CTAs); a warp claims tiles
 * from a global counter and keeps up to PT_TILE_SLOTS of them in flight: when the items of the newest tile are all
 * claimed the lanes that come free open the next tile while the stragglers of the older ones finish.  Per slot the
 * warp keeps, in shared memory, the tile id, the number of finished items and the running XYZ sums of the tile's 32
 * pixels (red.shared.add.f32, schedule order: fast mode only -- strict builds use v2s with its per-sample table).  A
 * tile whose items are all finished is written out (StoreTexel: the reference's Accumulate) by the whole warp and its
 * slot is recycled.  All bookkeeping happens in the NEW phase, which the warp executes converged (ballots, no locks).
 * The tile counter lives in the module (pt_tile_ctr[0]); the last warp to leave resets it (pt_tile_ctr[1] counts
 * leavers), so back-to-back launches on a stream need no host-side reset. */
#ifndef PT_TILE_SLOTS
#define PT_TILE_SLOTS 4
#endif
#define PT_V2SP_WORDS (PT_TILE_SLOTS * (3 * 32 + 2)) /* per warp: sums, then tile id and finished count per slot */

__device__ __forceinline__ void pt_render_body_v2sp(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                    float4* __restrict__ image, float* s_tab, float* s_all, unsigned* tileCtr) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_acc = s_all + warp * PT_V2SP_WORDS;                    /* [slot][channel][pixel] */
    int* s_tile = reinterpret_cast<int*>(s_acc + PT_TILE_SLOTS * 96); /* [slot]: the tile's pixel origin x0 | y0 << 16, or -1 */
    int* s_done = s_tile + PT_TILE_SLOTS;                           /* [slot]: finished (or skipped) items */
    for (int i = lane; i < PT_TILE_SLOTS * 96; i += 32) s_acc[i] = 0.0f;
    if (lane < PT_TILE_SLOTS) { s_tile[lane] = -1; s_done[lane] = 0; }
    __syncwarp();

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const int spf = pr.samplesPerFrame;
    const int tilesX = (pr.width + 7) >> 3, tilesY = (pr.height + 3) >> 2;
    const int numTiles = spf > 0 ? tilesX * tilesY : 0;
    const int total = 32 * spf;      /* items per tile: item = 32 * sample + pixel */
    int curSlot = -1, next = 0;      /* the slot items are being claimed from (warp-uniform) */
    bool exhausted = false;          /* the global counter ran past the last tile (warp-uniform) */
    int item = 0, slot = 0;          /* this lane's current item and the slot of its tile */

    int st = PT_ST_NEW;
    PathState ps;
    PathStateInit(ps);
    MarchState ms;
    MarchStateInit(ms);

    for (;;) {
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
#if PT_HAS_SDF
        const unsigned bSdf = __ballot_sync(0xffffffffu, st == PT_ST_SDF);
#else
        const unsigned bSdf = 0u;
#endif
        if ((bNew | bIs | bSdf | bSh) == 0u) break; /* every lane idle: no tile left and none in flight */
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
#if PT_HAS_SDF
        if (bSdf != 0u && (best == 0 || (best < PT_FEED_T && __popc(bSdf) >= PT_SDF_MIN))) phase = PT_ST_SDF;
#endif
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? bSdf : bSh)));
        if (phase == PT_ST_NEW) {
            /* 1. finished samples -> their pixel's sum; per slot, count them */
            const bool fin = (st == PT_ST_NEW) && ps.pendingFinish;
            if (fin) {
                const V3 col = PathColor(c, ps);
                float* e = s_acc + slot * 96 + (item & 31);
                atomicAdd(e, col.x); atomicAdd(e + 32, col.y); atomicAdd(e + 64, col.z);
                atomicAdd(s_done + slot, 1);
                ps.pendingFinish = false;
            }
            __syncwarp();
            /* 2. tiles whose items are all finished: write them out, recycle the slot, wake the lanes waiting for one */
            bool wake = false;
#pragma unroll 1
            for (int s = 0; s < PT_TILE_SLOTS; s++) {
                const int org = s_tile[s];
                if (org < 0 || s_done[s] != total) continue; /* warp-uniform */
                const int gx = (org & 0xffff) + (lane & 7), gy = (org >> 16) + (lane >> 3);
                float* e = s_acc + s * 96 + lane;
                if (gx < pr.width && gy < pr.height) StoreTexel(pr, image, gx, gy, mk3(e[0], e[32], e[64]));
                e[0] = 0.0f; e[32] = 0.0f; e[64] = 0.0f;
                __syncwarp();
                if (lane == 0) { s_tile[s] = -1; s_done[s] = 0; }
                __syncwarp();
                wake = true;
            }
            if (wake && st == PT_ST_IDLE) st = PT_ST_NEW;
            /* 3. claims: rank the lanes that want an item; serve them from the current tile, opening new ones as needed */
            const bool want = (st == PT_ST_NEW);
            const unsigned wmask = __ballot_sync(0xffffffffu, want);
            const int rank = __popc(wmask & ((1u << lane) - 1u));
            int remaining = __popc(wmask), base = 0;
            bool got = false;
#pragma unroll 1
            while (remaining > 0) {
                const int avail = curSlot >= 0 ? total - next : 0;
                const int take = avail < remaining ? avail : remaining;
                if (want && !got && rank >= base && rank < base + take) {
                    item = next + (rank - base);
                    slot = curSlot;
                    got = true;
                }
                next += take; base += take; remaining -= take;
                if (remaining == 0 || exhausted) break;
                int freeSlot = -1;
#pragma unroll 1
                for (int s = PT_TILE_SLOTS - 1; s >= 0; s--) if (s_tile[s] < 0) freeSlot = s;
                if (freeSlot < 0) break; /* every slot still has stragglers: the unserved lanes wait */
                unsigned id = 0u;
                if (lane == 0) id = atomicAdd(tileCtr, 1u);
                id = __shfl_sync(0xffffffffu, id, 0);
                if (id >= (unsigned)numTiles) { exhausted = true; break; }
                if (lane == 0) { /* the tile's pixel origin, packed: x0 | y0 << 16 */
                    const int ty = (int)id / tilesX, tx = (int)id - ty * tilesX;
                    s_tile[freeSlot] = (tx * 8) | ((ty * 4) << 16);
                    s_done[freeSlot] = 0;
                }
                __syncwarp();
                curSlot = freeSlot;
                next = 0;
            }
            if (want) {
                if (got) {
                    const int org = s_tile[slot];
                    const int q = item & 31;
                    const int qx = (org & 0xffff) + (q & 7), qy = (org >> 16) + (q >> 3);
                    if ((qx < pr.width) && (qy < pr.height)) st = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, item >> 5);
                    else atomicAdd(s_done + slot, 1); /* a pixel beyond the image edge: the item counts as finished, the lane claims again */
                } else {
                    st = PT_ST_IDLE;
                }
            }
            __syncwarp();
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
            }
        }
#if PT_HAS_SDF
        else if (phase == PT_ST_SDF) {
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (st == PT_ST_SDF) st = PhaseSdfEval(c, ps, ms);
#if PT_SDF_EXIT > 0
                if ((rep & 3) == 3 && __popc(__ballot_sync(0xffffffffu, st == PT_ST_SDF)) < PT_SDF_EXIT) break;
#endif
            }
            if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
        }
#endif
        else {
            if (st == PT_ST_SHADE) st = PhaseShade(c, ps);
        }
    }
    /* the last warp of the grid to leave rearms the tile counter for the next launch */
    if (lane == 0) {
        __threadfence();
        const unsigned leavers = atomicAdd(tileCtr + 1, 1u) + 1u;
        if (leavers == gridDim.x * gridDim.y * (PT_BLOCK_THREADS / 32)) {
            tileCtr[0] = 0u;
            tileCtr[1] = 0u;
            __threadfence();
        }
    }
}

/* One iteration of TracePath's loop (shader.comp:1393-1407): TraceRay (1345-1391) with SampleLightSource (1298-1343)
 * inline, verbatim from TracePath above, on the path held in `ps`.  Returns whether the path goes on.  Shared by the
 * flat-loop drivers v3 and v3s. */
PT_DEV bool TraceRayFlat(const Ctx& c, PathState& ps, const int pathLength) {
    const PtDevScene& sc = *c.sc;
    bool goOn = false;
    Hit h;
    Intersection(c, ps.ray, h, false);
    if (h.t < 1e5f) {
        float temperature, luminosity;
        GetLightMix(c, h.lightID, temperature, luminosity);
        if (luminosity > 0.0f) { /* emitter hit terminates the path */
            const V4 e = Emit(ps.l, PTK_MAX(temperature, 0.0f), PTK_MAX(luminosity, 0.0f));
            ps.radiance = ps.radiance + (e * ps.rayradiance) * ps.MISBRDFWeight;
        } else {
            float peak, sigma, invertf;
            GetMaterialMix(c, h.materialID, peak, sigma, invertf);
            const V4 brdf = EvaluateBRDF(ps.l, peak, sigma, invertf);
            Ray outRay;
            outRay.origin = fma3(ps.ray.dir, h.t, ps.ray.origin);
            outRay.dir = SampleCosineDirectionHemisphere(h.normal, ps.seed);
            const float BRDFpdf = PTK_DIV(dot(outRay.dir, h.normal), PT_PI_F);
            if (sc.numLights > 0.0f) { /* SampleLightSource, shader.comp:1298-1343 */
                const int randomLight = __float2int_rz(floorf(RandomFloatPCG32(ps.seed) * sc.numLights));
                const PtDevLightSlot& ls = sc.lightSlots[randomLight < sc.nLightSlots ? randomLight : sc.nLightSlots - 1];
                const V3 toLight = mk3(ls.px - outRay.origin.x, ls.py - outRay.origin.y, ls.pz - outRay.origin.z);
                const float invLightDistance = PTK_DIV(1.0f, length(toLight));
                const V3 lightDir = toLight * invLightDistance;
                const float sinthetaMax = PTK_MIN(ls.boundingRadius * invLightDistance, 1.0f);
                const float costhetaMax = PTK_SQRT(1.0f - sinthetaMax * sinthetaMax);
                Ray shadowRay;
                shadowRay.origin = outRay.origin;
                shadowRay.dir = ToWorld(SampleCosineUnitCone(ps.seed, costhetaMax), lightDir);
                float lightpdf = sc.invNumLights;
                lightpdf *= PTK_DIV(dot(shadowRay.dir, lightDir), PT_PI_F * (1.0f - costhetaMax * costhetaMax));
                ps.MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + lightpdf * lightpdf);
                const float costheta = dot(shadowRay.dir, h.normal);
                const float deathProbability = 1.25f * PTK_MAX(ps.MISBRDFWeight - 0.2f, 0.0f);
                if (costheta >= 0.0f) {
                    if (RandomFloatPCG32(ps.seed) > deathProbability) {
                        Hit sh;
                        Intersection(c, shadowRay, sh, true);
                        if (sh.objectID == ls.objectID) {
                            float lt, ll;
                            GetLightMix(c, ls.lightID, lt, ll);
                            const V4 rr = ps.rayradiance * mulDiv4(brdf, costheta, lightpdf);
                            const V4 e = Emit(ps.l, PTK_MAX(lt, 0.0f), PTK_MAX(ll, 0.0f));
                            ps.radiance = ps.radiance + (e * rr) * (1.0f - ps.MISBRDFWeight);
                        }
                    } else {
                        ps.MISBRDFWeight = 1.0f;
                    }
                }
            } else {
                ps.MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + 0.0f * 0.0f);
            }
            const float costheta = dot(outRay.dir, h.normal);
            ps.rayradiance = ps.rayradiance * mulDiv4(brdf, costheta, BRDFpdf);
            const float mx = PTK_MAX(ps.rayradiance.x, PTK_MAX(ps.rayradiance.y, PTK_MAX(ps.rayradiance.z, ps.rayradiance.w)));
            const float rayProbability = PTK_MIN(PTK_MAX(mx, 0.0f), 0.99f);
            if (!(RandomFloatPCG32(ps.seed) > rayProbability)) {
                ps.rayradiance = ps.rayradiance * PTK_DIV(1.0f, rayProbability);
                ps.ray = outRay;
                ps.bounce++;
                goOn = ps.bounce < pathLength;
            }
        }
    }
    return goOn;
}

/* ---- driver v3: v1's loop bodies, flattened, with gated path regeneration ----------------------------------------
 * What v1 loses on the analytic scenes is not the shape of a path but its tail: on scene1 nine lanes in ten are done
 * after two rays, yet the warp runs bounce 2's shading, shadow ray and the third and fourth intersection for the
 * one to four lanes still alive -- about half of all executions of the intersection code are that sparse (ncu: 17
 * of 32 lanes per instruction).  Here the sample loop and the bounce loop are ONE loop: an iteration is one bounce
 * (TraceRay, shader.comp:1345-1391, verbatim from TracePath above) for every lane with a live path, and the lanes
 * whose path has ended start their pixel's next sample (Scene() up to TracePath, 1446-1472) as soon as at least
 * PT_REGEN_T of them are waiting (or nobody is alive), so the stragglers' late bounces ride along with the next
 * sample's first ones.  Per-lane arithmetic and sample order are v1's: bit-exact in strict mode.  Unlike v2 there is
 * no phase vote and no explicit hit/shadow state: only v1's live variables, one extra flag and the sample counter. */
#ifndef PT_REGEN_T
#define PT_REGEN_T 16
#endif
__device__ __forceinline__ void pt_render_body_v3(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                  float4* __restrict__ image, float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const unsigned xyy = (unsigned)pr.height - (unsigned)gy; /* shader.comp:1510 */
    const int spf = inRange ? pr.samplesPerFrame : 0;
    const int pathLength = pr.pathLength;

    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps; /* only ray, l, radiance, rayradiance, MISBRDFWeight, seed, bounce and pendingFinish are used */
    PathStateInit(ps);
    bool alive = false;
    int k = 0;

    for (;;) {
        const bool wantNew = !alive && ((k < spf) || ps.pendingFinish);
        const unsigned bNew = __ballot_sync(0xffffffffu, wantNew);
        const unsigned bAlive = __ballot_sync(0xffffffffu, alive);
        if ((bNew | bAlive) == 0u) break;
        if ((__popc(bNew) >= PT_REGEN_T) || (bAlive == 0u)) { /* warp-uniform */
            if (wantNew) {
                if (ps.pendingFinish) { /* Scene()'s tail for the path that ended, shader.comp:1477-1489 */
                    outColor = outColor + PathColor(c, ps);
                    ps.pendingFinish = false;
                }
                if (k < spf) {
                    const int next = PhaseNew(c, ps, xyx, xyy, k);
                    k++;
                    alive = (next == PT_ST_ISECT); /* pathLength <= 0: PhaseNew left pendingFinish set */
                }
            }
        }
        if (alive) { /* one iteration of TracePath's loop, shader.comp:1393-1407 */
            const bool goOn = TraceRayFlat(c, ps, pathLength);
            if (!goOn) {
                alive = false;
                ps.pendingFinish = true;
            }
        }
    }
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}

/* ---- driver v3s (PT_SCHED=7): v3 with in-warp sample stealing ------------------------------------------------------
 * v3 keeps v1's compact loop bodies (no phase vote, no hit / shadow state) but ties a lane to its pixel's sample
 * sequence; v2s pools the tile's samples but pays for its phase machine on the scenes without SDFs (cfg2: 9.85 against
 * v1's 10.96 Gsamples/s).  This is v3's flat loop with v2s' pool: at a regeneration (PT_REGEN_T lanes waiting, or nobody
 * alive) the waiting lanes deposit their finished sample and claim the next items of the tile's 32 x S pool by ballot
 * rank, whichever pixel they belong to.  Strict mode: per-sample XYZ table, summed per pixel in sample order at the end
 * of a round (bit-exact); fast mode (PT_STEAL_S = 0): the whole dispatch is one pool, sums in shared memory.
 * Verified against the oracle on the host SIMT emulator (tests/test_simt_emulation.py); NOT yet measured on a GPU --
 * a round-2 candidate for the scenes without SDFs. */
__device__ __forceinline__ void pt_render_body_v3s(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_colAll) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = blockIdx.y * 8 + (warp >> 1) * 4;
    const int gx = tileX + (lane & 7), gy = tileY + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    float* s_col = s_colAll + warp * PT_STEAL_WORDS;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const int spf = pr.samplesPerFrame;
    const int pathLength = pr.pathLength;
    const bool warpLive = (tileX < pr.width) && (tileY < pr.height) && (spf > 0); /* warp-uniform */
    int roundBase = 0;
#if PT_STEAL_S == 0
    int roundN = warpLive ? spf : 0;
    s_col[lane] = 0.0f; s_col[32 + lane] = 0.0f; s_col[64 + lane] = 0.0f;
    __syncwarp();
#else
    int roundN = warpLive ? (spf < PT_STEAL_S ? spf : PT_STEAL_S) : 0;
#endif
    int next = 0, item = 0;

    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps; /* only ray, l, radiance, rayradiance, MISBRDFWeight, seed, bounce and pendingFinish are used */
    PathStateInit(ps);
    bool alive = false;

    for (;;) {
        const bool wantNew = !alive && ((next < 32 * roundN) || ps.pendingFinish);
        const unsigned bNew = __ballot_sync(0xffffffffu, wantNew);
        const unsigned bAlive = __ballot_sync(0xffffffffu, alive);
        if ((bNew | bAlive) == 0u) {
            if (!warpLive) break;
            /* end of a round: lane p sums pixel p's samples in index order */
            __syncwarp();
#if PT_STEAL_S == 0
            outColor = mk3(s_col[lane], s_col[32 + lane], s_col[64 + lane]);
            break;
#else
            if (inRange) {
#pragma unroll 1
                for (int kk = 0; kk < roundN; kk++) {
                    const float* e = s_col + (3 * kk) * 32 + lane;
                    outColor = outColor + mk3(e[0], e[32], e[64]);
                }
            }
            __syncwarp();
            roundBase += roundN;
            if (roundBase >= spf) break;
            roundN = (spf - roundBase) < PT_STEAL_S ? (spf - roundBase) : PT_STEAL_S;
            next = 0;
            continue;
#endif
        }
        if ((__popc(bNew) >= PT_REGEN_T) || (bAlive == 0u)) { /* warp-uniform */
            if (wantNew) {
                if (ps.pendingFinish) { /* Scene()'s tail for the path that ended, shader.comp:1477-1489 */
                    const V3 col = PathColor(c, ps);
#if PT_STEAL_S == 0
                    float* e = s_col + (item & 31);
                    atomicAdd(e, col.x); atomicAdd(e + 32, col.y); atomicAdd(e + 64, col.z);
#else
                    float* e = s_col + (3 * (item >> 5)) * 32 + (item & 31);
                    e[0] = col.x; e[32] = col.y; e[64] = col.z;
#endif
                    ps.pendingFinish = false;
                }
                item = next + __popc(bNew & ((1u << lane) - 1u));
                if (item < 32 * roundN) {
                    const int q = item & 31;
                    const int qx = tileX + (q & 7), qy = tileY + (q >> 3);
                    if ((qx < pr.width) && (qy < pr.height)) { /* else: a pixel beyond the image edge; claim again */
                        const int nextState = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, roundBase + (item >> 5));
                        alive = (nextState == PT_ST_ISECT); /* pathLength <= 0: PhaseNew left pendingFinish set */
                    }
                }
            }
            next += __popc(bNew);
        }
        if (alive) { /* one iteration of TracePath's loop, shader.comp:1393-1407 */
            if (!TraceRayFlat(c, ps, pathLength)) {
                alive = false;
                ps.pendingFinish = true;
            }
        }
    }
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}

/* ---- driver v2d: v2 with TWO pixels per lane, the idle one parked in shared memory ---------------------------------
 * v2 cannot fill its phases because a lane is tied to one pixel's sample sequence: a lane whose ray waits in front of
 * the SDF phase is a lane the feeder phases do not have, so the (expensive) SDF phase runs when the feeders dry up,
 * with whoever happens to wait -- 1-4 of 32 lanes in 61-73 % of its executions on the fractal scenes
 * (profiles/r01_sdfsched).  Here every lane owns two pixels (rows gy and gy + 8 of a 16x16 tile).  One path lives in
 * registers, the other is parked in the lane's own column of shared memory (56 words, no atomics, no queues).  A lane
 * whose active path has to wait for the SDF phase swaps to its other pixel and keeps feeding; it only counts as waiting
 * when BOTH its paths wait (the older one is marched first).  So the SDF phase starts with most of the warp in it, and
 * the feeders lose a lane only when it has nothing else to do.  Per pixel the samples still run in order through the
 * same phases: bit-exact in strict mode. */
#define PT_PARK_WORDS 56
#ifndef PT_SWAP_MIN
#define PT_SWAP_MIN 8
#endif
PT_DEV void ParkSwap(float* col, PathState& ps, MarchState& ms, int& st, int& k, V3& outColor) {
    int w = 0;
#define PT_SWF(x) { const float t_ = col[w * PT_BLOCK_THREADS]; col[w * PT_BLOCK_THREADS] = (x); (x) = t_; w++; }
#define PT_SWI(x) { const int t_ = __float_as_int(col[w * PT_BLOCK_THREADS]); col[w * PT_BLOCK_THREADS] = __int_as_float((int)(x)); (x) = t_; w++; }
    PT_SWF(ps.ray.origin.x) PT_SWF(ps.ray.origin.y) PT_SWF(ps.ray.origin.z)
    PT_SWF(ps.ray.dir.x) PT_SWF(ps.ray.dir.y) PT_SWF(ps.ray.dir.z)
    PT_SWF(ps.l.x) PT_SWF(ps.l.y) PT_SWF(ps.l.z) PT_SWF(ps.l.w)
    PT_SWF(ps.radiance.x) PT_SWF(ps.radiance.y) PT_SWF(ps.radiance.z) PT_SWF(ps.radiance.w)
    PT_SWF(ps.rayradiance.x) PT_SWF(ps.rayradiance.y) PT_SWF(ps.rayradiance.z) PT_SWF(ps.rayradiance.w)
    PT_SWF(ps.MISBRDFWeight)
    PT_SWF(ps.shDir.x) PT_SWF(ps.shDir.y) PT_SWF(ps.shDir.z)
    PT_SWF(ps.shContrib.x) PT_SWF(ps.shContrib.y) PT_SWF(ps.shContrib.z) PT_SWF(ps.shContrib.w)
    PT_SWF(ps.h.t) PT_SWF(ps.h.normal.x) PT_SWF(ps.h.normal.y) PT_SWF(ps.h.normal.z) PT_SWF(ps.h.materialID) PT_SWF(ps.h.lightID)
    PT_SWF(ms.mt) PT_SWF(ms.insT) PT_SWF(ms.omega) PT_SWF(ms.previousRadius) PT_SWF(ms.tMax) PT_SWF(ms.ksign)
    PT_SWF(ms.probe) PT_SWF(ms.nrm0) PT_SWF(ms.nrm1) PT_SWF(ms.nrm2)
    PT_SWF(outColor.x) PT_SWF(outColor.y) PT_SWF(outColor.z)
    { /* seed and set1 are full 32-bit words */
        unsigned s_ = ps.seed; int si_ = (int)s_; PT_SWI(si_) ps.seed = (unsigned)si_;
        unsigned m_ = ms.set1; int mi_ = (int)m_; PT_SWI(mi_) ms.set1 = (unsigned)mi_;
    }
    int flags = ps.bounce | (ps.isShadow ? (1 << 28) : 0) | (ps.pathAlive ? (1 << 29) : 0) | (ps.pendingFinish ? (1 << 30) : 0);
    PT_SWI(flags)
    ps.bounce = flags & ((1 << 28) - 1);
    ps.isShadow = (flags & (1 << 28)) != 0; ps.pathAlive = (flags & (1 << 29)) != 0; ps.pendingFinish = (flags & (1 << 30)) != 0;
    PT_SWI(ps.shObj) PT_SWI(ps.h.objectID)
    PT_SWI(ms.points) PT_SWI(ms.iter) PT_SWI(ms.sub)
    PT_SWI(st) PT_SWI(k)
#undef PT_SWF
#undef PT_SWI
}

__device__ __forceinline__ void pt_render_body_v2d(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_park) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy0 = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3); /* the lane's pixels: rows gy0 and gy0 + 8 */
    float* col = s_park + threadIdx.x;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const int spf = pr.samplesPerFrame;
    const bool inA = (gx < pr.width) && (gy0 < pr.height), inB = (gx < pr.width) && (gy0 + 8 < pr.height);

    PathState ps;
    MarchState ms;
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    int k = 0;
    /* pixel B starts parked: a path in phase NEW (or DONE) carries no other state (k = 0, colour 0, nothing pending) */
    const int stOther0 = (inB && spf > 0) ? PT_ST_NEW : PT_ST_DONE;
#pragma unroll 1
    for (int w = 0; w < PT_PARK_WORDS; w++) col[w * PT_BLOCK_THREADS] = 0.0f;
    col[(PT_PARK_WORDS - 2) * PT_BLOCK_THREADS] = __int_as_float(stOther0); /* ParkSwap's layout: ..., st, k */
    int stOther = stOther0; /* register mirror of the parked path's phase */
    PathStateInit(ps);
    MarchStateInit(ms);
    int st = (inA && spf > 0) ? PT_ST_NEW : PT_ST_DONE;
    int cur = 0; /* which pixel is in registers: 0 = row gy0, 1 = row gy0 + 8 */
    bool settled = false; /* the active path is the older of two that wait for the SDF phase */
    V3 colorA = mk3(0.0f, 0.0f, 0.0f), colorB = mk3(0.0f, 0.0f, 0.0f);

    for (;;) {
        { /* lane-local: keep a runnable path in registers; when both wait for the SDF phase, march the older one */
            const bool blocked = (st == PT_ST_SDF) || (st == PT_ST_DONE);
            const bool otherRunnable = (stOther != PT_ST_SDF) && (stOther != PT_ST_DONE);
            if (st != PT_ST_SDF) settled = false;
            /* both wait for the SDF phase: bring the one that has waited longer (the parked one) in, once */
            const bool toOlder = (stOther == PT_ST_SDF) && ((st == PT_ST_DONE) || ((st == PT_ST_SDF) && !settled));
            const bool wantSwap = (blocked && otherRunnable) || toOlder;
            /* the exchange costs ~250 instructions for the whole warp however many lanes take part: batch it -- wait for
             * PT_SWAP_MIN candidates unless no lane has anything else to run */
            const unsigned bWant = __ballot_sync(0xffffffffu, wantSwap);
            const unsigned bRun = __ballot_sync(0xffffffffu, (st == PT_ST_NEW) || (st == PT_ST_ISECT) || (st == PT_ST_SHADE));
            if (bWant != 0u && (__popc(bWant) >= PT_SWAP_MIN || bRun == 0u)) {
                if (wantSwap) {
                    if (st == PT_ST_DONE) { if (cur == 0) colorA = outColor; else colorB = outColor; }
                    const int mine = st;
                    ParkSwap(col, ps, ms, st, k, outColor);
                    stOther = mine;
                    cur ^= 1;
                    settled = toOlder;
                }
            }
        }
        const unsigned xyy = (unsigned)pr.height - (unsigned)(gy0 + 8 * cur); /* shader.comp:1510 */
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
#if PT_HAS_SDF
        const unsigned bSdf = __ballot_sync(0xffffffffu, st == PT_ST_SDF);
#else
        const unsigned bSdf = 0u;
#endif
        if ((bNew | bIs | bSdf | bSh) == 0u) break; /* every lane: active DONE, and then the parked one is DONE too */
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
#if PT_HAS_SDF
        if (bSdf != 0u && (best < PT_FEED_T || best == 0)) phase = PT_ST_SDF;
#endif
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? bSdf : bSh)));
#ifdef PT_STATS
        if (phase == PT_ST_SDF && (threadIdx.x & 31) == 0) atomicAdd(&pt_stats[8 + ((__popc(bSdf) - 1) >> 2)], 1ull);
#endif
        if (phase == PT_ST_NEW) {
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) {
                    outColor = outColor + PathColor(c, ps);
                    ps.pendingFinish = false;
                }
                if (k < spf) {
                    st = PhaseNew(c, ps, xyx, xyy, k);
                    k++;
                } else {
                    st = PT_ST_DONE;
                }
            }
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
            }
        }
#if PT_HAS_SDF
        else if (phase == PT_ST_SDF) {
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (st == PT_ST_SDF) st = PhaseSdfEval(c, ps, ms);
            }
            if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
        }
#endif
        else {
            if (st == PT_ST_SHADE) st = PhaseShade(c, ps);
        }
    }
    if (cur == 0) colorA = outColor; else colorB = outColor;
    if (inA) StoreTexel(pr, image, gx, gy0, colorA);
    if (inB) StoreTexel(pr, image, gx, gy0 + 8, colorB);
}

#if PT_HAS_SDF
/* ---- driver v2p: v2 + a march pool shared by the warps of a CTA --------------------------------------------------
 * In v2 only the lanes of ONE warp that happen to be marching populate the SDF phase (ncu: 7 of 32 on the
 * mandelbulb scene).  Here a lane whose ray entered an SDF box publishes the march as a job in shared memory
 * (origin, direction, SphereTracing's locals: 24 words) and waits; whenever a warp runs the SDF phase, ALL its
 * lanes -- whatever their own state -- claim pending jobs of ANY thread of the CTA (128-bit pending mask, claims by
 * atomicAnd), advance them PT_SDF_REPS evaluations with the same PhaseSdfEval, and put them back or mark them done.
 * The per-job arithmetic does not depend on who executes it, so strict mode stays bit-exact. */
enum { PT_ST_WAIT = 5 };
enum { PJ_OX = 0, PJ_OY, PJ_OZ, PJ_DX, PJ_DY, PJ_DZ, PJ_SHADOW, PJ_HT, PJ_MT, PJ_INST, PJ_OMEGA, PJ_PREV, PJ_TMAX, PJ_KSIGN,
       PJ_PROBE, PJ_N0, PJ_N1, PJ_N2, PJ_POINTS, PJ_ITER, PJ_SUB, PJ_SET1, PJ_OBJ, PJ_MAT, PJ_DONE, PJ_FIELDS };
#ifndef PT_POOL_MIN
#define PT_POOL_MIN 24
#endif

__device__ __forceinline__ void pt_render_body_v2p(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_job, unsigned* s_mask) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    if (threadIdx.x < 4) s_mask[threadIdx.x] = 0u;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    const int tid = threadIdx.x;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const unsigned xyy = (unsigned)pr.height - (unsigned)gy; /* shader.comp:1510 */
    const int spf = pr.samplesPerFrame;

    int st = (inRange && spf > 0) ? PT_ST_NEW : PT_ST_DONE;
    int k = 0;
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps;
    PathStateInit(ps);
    volatile float* job = s_job;
    volatile unsigned* vmask = s_mask;
#define PJ(f, t) job[(f) * PT_BLOCK_THREADS + (t)]

    for (;;) {
        /* a finished job: take the hit back */
        if (st == PT_ST_WAIT && PJ(PJ_DONE, tid) != 0.0f) {
            ps.h.t = PJ(PJ_HT, tid);
            ps.h.objectID = __float_as_int(PJ(PJ_OBJ, tid));
            if (__float_as_int(PJ(PJ_SUB, tid)) == PT_SUB_N0 + 6) { /* a path ray hit the SDF: normal and material came with it */
                ps.h.normal = mk3(PJ(PJ_N0, tid), PJ(PJ_N1, tid), PJ(PJ_N2, tid));
                ps.h.materialID = PJ(PJ_MAT, tid);
                ps.h.lightID = -1.0f;
            }
            st = PhaseTrivial(ps);
        }
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
        const unsigned bWait = __ballot_sync(0xffffffffu, st == PT_ST_WAIT);
        if ((bNew | bIs | bSh | bWait) == 0u) break;
        /* one snapshot of the pending mask for the whole warp (lane 0's), so ranks and the vote are consistent */
        const unsigned m0 = __shfl_sync(0xffffffffu, vmask[0], 0), m1 = __shfl_sync(0xffffffffu, vmask[1], 0),
                       m2 = __shfl_sync(0xffffffffu, vmask[2], 0), m3 = __shfl_sync(0xffffffffu, vmask[3], 0);
        const int pending = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
        /* march when a (nearly) full warp's worth of jobs is waiting in the CTA, or when this warp's feeders are thin */
        if (pending > 0 && (pending >= PT_POOL_MIN || best < PT_FEED_T)) phase = PT_ST_SDF;
        if (best == 0 && phase != PT_ST_SDF) { /* everyone here waits on jobs other warps are running */
            __nanosleep(200);
            continue;
        }
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? 0u : bSh)));
        if (phase == PT_ST_NEW) {
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) {
                    outColor = outColor + PathColor(c, ps);
                    ps.pendingFinish = false;
                }
                if (k < spf) {
                    st = PhaseNew(c, ps, xyx, xyy, k);
                    k++;
                } else {
                    st = PT_ST_DONE;
                }
            }
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
                MarchState ms;
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SDF) { /* publish the march as a job of the CTA's pool */
                    const V3 d = ps.isShadow ? ps.shDir : ps.ray.dir;
                    PJ(PJ_OX, tid) = ps.ray.origin.x; PJ(PJ_OY, tid) = ps.ray.origin.y; PJ(PJ_OZ, tid) = ps.ray.origin.z;
                    PJ(PJ_DX, tid) = d.x; PJ(PJ_DY, tid) = d.y; PJ(PJ_DZ, tid) = d.z;
                    PJ(PJ_SHADOW, tid) = ps.isShadow ? 1.0f : 0.0f;
                    PJ(PJ_HT, tid) = ps.h.t;
                    PJ(PJ_OBJ, tid) = __int_as_float(ps.h.objectID);
                    PJ(PJ_MT, tid) = ms.mt; PJ(PJ_INST, tid) = ms.insT; PJ(PJ_OMEGA, tid) = ms.omega;
                    PJ(PJ_PREV, tid) = ms.previousRadius; PJ(PJ_TMAX, tid) = ms.tMax; PJ(PJ_KSIGN, tid) = 0.0f;
                    PJ(PJ_PROBE, tid) = 0.0f; PJ(PJ_N0, tid) = 0.0f; PJ(PJ_N1, tid) = 0.0f; PJ(PJ_N2, tid) = 0.0f;
                    PJ(PJ_POINTS, tid) = __int_as_float(ms.points); PJ(PJ_ITER, tid) = __int_as_float(ms.iter);
                    PJ(PJ_SUB, tid) = __int_as_float(ms.sub); PJ(PJ_SET1, tid) = __uint_as_float(ms.set1);
                    PJ(PJ_DONE, tid) = 0.0f;
                    __threadfence_block();
                    atomicOr(&s_mask[warp], 1u << lane);
                    st = PT_ST_WAIT;
                } else {
                    st = PhaseTrivial(ps);
                }
            }
        } else if (phase == PT_ST_SDF) {
            /* lane L takes the L-th pending job of the CTA (if it wins the atomicAnd against the other warps) */
            int j = -1;
            {
                int r = lane;
                const unsigned mm[4] = {m0, m1, m2, m3};
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const int n = __popc(mm[w]);
                    if (j < 0 && r < n) j = w * 32 + (int)__fns(mm[w], 0u, r + 1);
                    r -= n;
                }
            }
            if (j >= 0) {
                const unsigned bit = 1u << (j & 31);
                if ((atomicAnd(&s_mask[j >> 5], ~bit) & bit) == 0u) j = -1; /* another warp got it */
            }
            if (j >= 0) {
                __threadfence_block();
                PathState js;
                MarchState ms;
                js.ray.origin = mk3(PJ(PJ_OX, j), PJ(PJ_OY, j), PJ(PJ_OZ, j));
                js.ray.dir = mk3(PJ(PJ_DX, j), PJ(PJ_DY, j), PJ(PJ_DZ, j));
                js.shDir = js.ray.dir;
                js.isShadow = PJ(PJ_SHADOW, j) != 0.0f;
                js.h.t = PJ(PJ_HT, j);
                js.h.objectID = __float_as_int(PJ(PJ_OBJ, j));
                js.h.normal = mk3(0.0f, 0.0f, 0.0f); js.h.materialID = 0.0f; js.h.lightID = -1.0f;
                ms.mt = PJ(PJ_MT, j); ms.insT = PJ(PJ_INST, j); ms.omega = PJ(PJ_OMEGA, j); ms.previousRadius = PJ(PJ_PREV, j);
                ms.tMax = PJ(PJ_TMAX, j); ms.ksign = PJ(PJ_KSIGN, j); ms.probe = PJ(PJ_PROBE, j);
                ms.nrm0 = PJ(PJ_N0, j); ms.nrm1 = PJ(PJ_N1, j); ms.nrm2 = PJ(PJ_N2, j);
                ms.points = __float_as_int(PJ(PJ_POINTS, j)); ms.iter = __float_as_int(PJ(PJ_ITER, j));
                ms.sub = __float_as_int(PJ(PJ_SUB, j)); ms.set1 = __float_as_uint(PJ(PJ_SET1, j));
                int jst = PT_ST_SDF;
#pragma unroll 1
                for (int rep = 0; rep < PT_SDF_REPS && jst == PT_ST_SDF; rep++) jst = PhaseSdfEval(c, js, ms);
                PJ(PJ_HT, j) = js.h.t;
                PJ(PJ_OBJ, j) = __int_as_float(js.h.objectID);
                PJ(PJ_SUB, j) = __int_as_float(ms.sub);
                if (jst == PT_ST_SDF) {
                    PJ(PJ_MT, j) = ms.mt; PJ(PJ_INST, j) = ms.insT; PJ(PJ_OMEGA, j) = ms.omega; PJ(PJ_PREV, j) = ms.previousRadius;
                    PJ(PJ_TMAX, j) = ms.tMax; PJ(PJ_KSIGN, j) = ms.ksign; PJ(PJ_PROBE, j) = ms.probe;
                    PJ(PJ_N0, j) = ms.nrm0; PJ(PJ_N1, j) = ms.nrm1; PJ(PJ_N2, j) = ms.nrm2;
                    PJ(PJ_POINTS, j) = __int_as_float(ms.points); PJ(PJ_ITER, j) = __int_as_float(ms.iter);
                    PJ(PJ_SET1, j) = __uint_as_float(ms.set1);
                    __threadfence_block();
                    atomicOr(&s_mask[j >> 5], 1u << (j & 31));
                } else {
                    if (ms.sub == PT_SUB_N0 + 6) {
                        PJ(PJ_N0, j) = js.h.normal.x; PJ(PJ_N1, j) = js.h.normal.y; PJ(PJ_N2, j) = js.h.normal.z;
                        PJ(PJ_MAT, j) = js.h.materialID;
                    }
                    __threadfence_block();
                    PJ(PJ_DONE, j) = 1.0f;
                }
            }
        } else {
            if (st == PT_ST_SHADE) st = PhaseShade(c, ps);
        }
    }
#undef PJ
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}
#endif /* PT_HAS_SDF */

#ifndef PT_SCHED
#define PT_SCHED 1
#endif
__device__ __forceinline__ void pt_render_body(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                               float4* __restrict__ image, float* s_tab) {
#if PT_SCHED == 3
    pt_render_body_v3(sc, pr, ubo, image, s_tab);
#elif PT_SCHED
    pt_render_body_v2(sc, pr, ubo, image, s_tab);
#else
    pt_render_body_v1(sc, pr, ubo, image, s_tab);
#endif
}

} /* namespace PT_KERNEL_NS */

/* the kernel entry point; the name distinguishes the strict / fast / JIT instances */
#if PT_SCHED == 4
#define PT_ROWS_PER_BLOCK 16
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __device__ int pt_rows_per_block = 16;                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_park[PT_PARK_WORDS * PT_BLOCK_THREADS];                                           \
        PT_KERNEL_NS::pt_render_body_v2d(sc, pr, ubo, image, s_tab, s_park);                                 \
    }
#elif PT_SCHED == 6
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __device__ int pt_persistent_ctas_per_sm = PT_MIN_BLOCKS;                                     \
    extern "C" __device__ unsigned pt_tile_ctr[2] = {0u, 0u};                                                \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_v2sp[PT_V2SP_WORDS * (PT_BLOCK_THREADS / 32)];                                    \
        PT_KERNEL_NS::pt_render_body_v2sp(sc, pr, ubo, image, s_tab, s_v2sp, pt_tile_ctr);                   \
    }
#elif PT_SCHED == 7
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_col[PT_STEAL_WORDS * (PT_BLOCK_THREADS / 32)];                                    \
        PT_KERNEL_NS::pt_render_body_v3s(sc, pr, ubo, image, s_tab, s_col);                                  \
    }
#elif PT_SCHED == 5
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_col[(PT_STEAL_WORDS + PT_MPARK_WORDS) * (PT_BLOCK_THREADS / 32)];                 \
        PT_KERNEL_NS::pt_render_body_v2s(sc, pr, ubo, image, s_tab, s_col);                                  \
    }
#elif PT_HAS_SDF && PT_SCHED == 2
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_job[PT_KERNEL_NS::PJ_FIELDS * PT_BLOCK_THREADS];                                  \
        __shared__ unsigned s_mask[4];                                                                       \
        PT_KERNEL_NS::pt_render_body_v2p(sc, pr, ubo, image, s_tab, s_job, s_mask);                          \
    }
#else
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        PT_KERNEL_NS::pt_render_body(sc, pr, ubo, image, s_tab);                                             \
    }
#endif


#if PT_HAS_SDF
/* SDF()/SDFMATERIAL() at arbitrary points: used by the tests to compare the NVRTC build of the snippets with the
 * CPU oracle's g++ build bit for bit (pt_sdf_eval) */
#define PT_DEFINE_SDF_EVAL_KERNEL(name)                                                                      \
    extern "C" __global__ void name(const float* __restrict__ xyz, unsigned long long n, unsigned set1,      \
                                    float* __restrict__ dist, float* __restrict__ material) {                \
        const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;              \
        if (i >= n) return;                                                                                  \
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];                                  \
        if (dist) dist[i] = ::pt_sdf_dispatch(x, y, z, set1);                                                  \
        if (material) material[i] = ::pt_sdfmaterial_dispatch(x, y, z, set1);                                  \
    }
#endif

#endif /* PT_KERNEL_CUH */
