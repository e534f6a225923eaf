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
 * Layout of this file: value types and mode-dependent primitives; the shader's leaf functions (RNG, spectral helpers,
 * intersections, SDF search); then ONE statement of a path as four phases over an explicit PathState --
 *   PhaseNew (camera + lens + wavelengths), PhaseIsect (closest hit over the primitives + SearchSDF), PhaseSdfEval (one
 *   SDF() evaluation of the march or of the numerical normal), PhaseShade (emitter / BSDF / light sample / roulette),
 *   PathColor (XYZ projection) --
 * and the DRIVERS that schedule those phases over the lanes of a warp:
 *   v1  (PT_SCHED 0)  nested sample / bounce loops per thread (TraceRayFlat = ISECT + SHADE + shadow ISECT inline)
 *   v3s (PT_SCHED 7)  v1's loop bodies in one flat loop; finished lanes regenerate from the tile's sample pool
 *   v2s (PT_SCHED 5)  one phase per warp iteration chosen by ballot; finished lanes take the next sample of the tile's
 *                     pool whichever pixel it belongs to (in-warp sample stealing)
 *   v2m (PT_SCHED 8)  (pt_driver_v2m.cuh) v2s + a per-warp pool of PARKED paths in shared memory: a ray that has to march is parked with its
 *                     whole path state and its lane takes other work; the SDF phase marches parked rays with all 32
 *                     lanes whatever paths those lanes hold in registers
 * -- two small kernels that take the coherent ends of Scene() out of the pooled drivers' hot loop (options pregen /
 * resolve): pt_gen_body computes PhaseNew's camera half for every sample of a dispatch into 32-byte records with all
 * lanes busy, pt_resolve_body projects stored radiance bundles to XYZ and sums them per pixel in sample order --
 * and the kernel entry macros.  pt_wavefront.cuh runs the same phases as separate kernels over state in HBM.
 * Surface extensions (PT_EXT_BSDF: mirror / glossy / dielectric lobes; NOT reference behaviour) are compiled in only for
 * scenes that ask for them (pt_set_surface_ext).
 * Drivers that were measured and dropped (v2, v2p CTA job board, v2d two pixels per lane, v3, v2sp persistent tiles,
 * march parking inside v2s) live in the history of this file and in profiles/r01_*; DESIGN.md section 7 has their numbers.
 * The kernel is instruction-cache bound (16-byte SASS): single call sites and rolled loops are deliberate.
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
#ifndef PT_PATHCOLOR_UNROLL
#define PT_PATHCOLOR_UNROLL 0 /* experiment: PathColor's four table look-ups unrolled also in SDF builds */
#endif
#ifndef PT_EXT_BSDF
#define PT_EXT_BSDF 0 /* 1: the scene carries surface extensions (pt_set_surface_ext); 0 = the reference's shading, untouched */
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
__device__ float pt_sdf_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4);
__device__ float pt_sdfmaterial_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4);
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
    /* the four candidates 0.5 * (-+sqrts2subp +- sq) - bdiv4a in the shader's order (r0..r3), through ONE copy of the
     * Newton step (instruction-cache footprint): the signs enter as factors of +-1, which is exact */
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const float inv = (k < 2) ? invaddq2div : invsubq2div;
        if (inv >= 0.0f) {
            const float sq = PTK_SQRT(inv);
            const float s1 = (k < 2) ? -1.0f : 1.0f, s2 = (k & 1) ? -1.0f : 1.0f;
            const float rk = QuarticNewton(a, b, c, d, e, a4, b3, c2, 0.5f * (s1 * sqrts2subp + s2 * sq) - bdiv4a);
            if ((rk < t) && (rk > 0.0f)) t = rk;
        }
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

/* Which SDFs the bounding-box search found: bit i % 32 of word i / 32 -- the shader's four masks set1..set4
 * (shader.comp:706-719).  The reference fills set1 only (732-738: "Program To Use set2, set3, set4 Not Yet Written"),
 * which caps it at 32 SDFs; the other words are filled here the way InsertSDF's dispatcher lines already read them
 * (host:2012: set<i/32 + 1> & 2^(i % 32)) -- SURVEY 8f-3.  PT_SDF_WORDS = ceil(n_sdf / 32) is baked in by the JIT: with
 * at most 32 SDFs (every shipped scene) the set is one register and the arithmetic is the reference's. */
#ifndef PT_SDF_WORDS
#define PT_SDF_WORDS 1
#endif
struct SdfSet { unsigned w[PT_SDF_WORDS]; };
PT_DEV void SdfSetClear(SdfSet& s) {
#pragma unroll
    for (int k = 0; k < PT_SDF_WORDS; k++) s.w[k] = 0u;
}
PT_DEV void SdfSetAdd(SdfSet& s, int i) { /* set += 1 << i */
    const unsigned bit = 1u << (unsigned)(i & 31);
#pragma unroll
    for (int k = 0; k < PT_SDF_WORDS; k++)
        if ((i >> 5) == k) s.w[k] += bit;
}
#define PT_SDF_SET_ARGS(s) (s).w[0], (PT_SDF_WORDS > 1 ? (s).w[PT_SDF_WORDS > 1 ? 1 : 0] : 0u), \
                           (PT_SDF_WORDS > 2 ? (s).w[PT_SDF_WORDS > 2 ? 2 : 0] : 0u), (PT_SDF_WORDS > 3 ? (s).w[PT_SDF_WORDS > 3 ? 3 : 0] : 0u)
/* ---- SDF sphere tracing (shader.comp:704-860) ------------------------------------------------------------------ */
#if PT_HAS_SDF
PT_DEV float SDF(V3 p, const SdfSet& set1) { return ::pt_sdf_dispatch(p.x, p.y, p.z, PT_SDF_SET_ARGS(set1)); }

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
PT_DEV bool SearchSDF(const Ctx& c, V3 p, V3 invdir, float& tMin, float& tMax, SdfSet& set1) {
    bool isFoundSDF = false;
    SdfSetClear(set1);
    const int n = PT_N_SDF(c);
    for (int i = 0; i < n; i++) {
        float bx, by;
        RayIntersectAABB(p, invdir, reinterpret_cast<const PtDevSdf*>(c.sc->pool + PT_OFF_SDFS(*c.sc))[i], bx, by);
        if ((bx > by) || (by < 0.0f)) continue;
        if (bx < tMin) {
            isFoundSDF = true;
            if (by < tMin) {
                tMin = bx; tMax = by;
                SdfSetClear(set1); SdfSetAdd(set1, i);
            } else {
                if (by < tMax) {
                    tMin = bx;
                    SdfSetAdd(set1, i);
                } else {
                    tMin = bx; tMax = by;
                    SdfSetAdd(set1, i);
                }
            }
        } else {
            if (bx < tMax) {
                isFoundSDF = true;
                if (by > tMax) tMax = by;
                SdfSetAdd(set1, i);
            }
        }
    }
    return isFoundSDF;
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


#if !PT_BVH
/* The scan in two halves, for the drivers that run the heavy tests as a phase of their own (v2s, option heavy_min):
 * IntersectLight = spheres and planes, and for boxes / lenses / cyclides only the bounding-sphere cull -- whoever passes
 * is a bit in the returned candidate mask (bit = position among the boxes, lenses, cyclides in scan order; at most 32,
 * pt_jit.cpp keeps the option off beyond that);  IntersectHeavy = the full tests of the candidates, in the same order.
 * Light then Heavy is IntersectionAnalytic's scan operation for operation: spheres and planes come first in the
 * reference's order anyway (shader.comp:866-924), so `t < hit.t` sees the same sequence of hits. */
PT_DEV unsigned IntersectLight(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    const PtDevScene& sc = *c.sc;
    h.t = 1e5f;
    h.objectID = -1;
    if (!kShadow) {
        h.normal = mk3(0.0f, 0.0f, 0.0f);
        h.materialID = 0.0f;
        h.lightID = -1.0f;
    }
    int base = 0;
    const int nS = PT_N_SPHERES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nS; i++) SphereIntersection(ray, reinterpret_cast<const PtDevSphere*>(sc.pool)[i], base + i, h, kShadow);
    base += nS;
    const int nP = PT_N_PLANES(c);
    PT_UNROLL_PRIMS
    for (int i = 0; i < nP; i++) PlaneIntersection(ray, reinterpret_cast<const PtDevPlane*>(sc.pool + PT_OFF_PLANES(sc))[i], base + i, h, kShadow);
    unsigned cand = 0u;
    int bit = 0;
    const int nB = PT_N_BOXES(c), nL = PT_N_LENSES(c), nC = PT_N_CYCLIDES(c);
    for (int i = 0; i < nB; i++, bit++) {
        const PtDevBox& o = reinterpret_cast<const PtDevBox*>(sc.pool + PT_OFF_BOXES(sc))[i];
        if (BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) cand |= 1u << bit;
    }
    for (int i = 0; i < nL; i++, bit++) {
        const PtDevLens& o = reinterpret_cast<const PtDevLens*>(sc.pool + PT_OFF_LENSES(sc))[i];
        if (BoundingSphere(ray, o.px, o.py, o.pz, o.bound2)) cand |= 1u << bit;
    }
    for (int i = 0; i < nC; i++, bit++) {
        const PtDevCyclide& o = reinterpret_cast<const PtDevCyclide*>(sc.pool + PT_OFF_CYCLIDES(sc))[i];
        if (BoundingSphere(ray, o.px, o.py, o.pz, o.brad)) cand |= 1u << bit;
    }
    return cand;
}
PT_DEV void IntersectHeavy(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow, unsigned cand) {
    const PtDevScene& sc = *c.sc;
    const int nB = PT_N_BOXES(c), nL = PT_N_LENSES(c), nC = PT_N_CYCLIDES(c);
    int base = PT_N_SPHERES(c) + PT_N_PLANES(c), bit = 0;
    for (int i = 0; i < nB; i++, bit++) {
        if (((cand >> bit) & 1u) == 0u) continue;
        BoxIntersection(ray, reinterpret_cast<const PtDevBox*>(sc.pool + PT_OFF_BOXES(sc))[i], base + i, h, kShadow);
    }
    base += nB;
    for (int i = 0; i < nL; i++, bit++) {
        if (((cand >> bit) & 1u) == 0u) continue;
        int isOutside = 1;
        LensIntersection(ray, reinterpret_cast<const PtDevLens*>(sc.pool + PT_OFF_LENSES(sc))[i], base + i, h, isOutside, kShadow);
    }
    base += nL;
    for (int i = 0; i < nC; i++, bit++) {
        if (((cand >> bit) & 1u) == 0u) continue;
        DupinCyclide(ray, reinterpret_cast<const PtDevCyclide*>(sc.pool + PT_OFF_CYCLIDES(sc))[i], base + i, h, kShadow);
    }
}
#endif


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
    bool inside;        /* PT_EXT_BSDF: the path is inside a dielectric (toggled by every refraction); else always false */
    V3 shDir;
    V3 traceDir;        /* direction of the ray being traced: shDir while a shadow ray is pending, else ray.dir.  The phases that
                           serve both kinds of ray read this instead of selecting (three FSEL per intersection and per SDF
                           evaluation, on the half-rate ALU pipe); drivers that never read it pay nothing (dead stores) */
    V4 shContrib;       /* added to radiance iff the shadow ray sees object shObj (kDeferEmit: the factor Emit() is scaled by) */
    int shObj;
    float shScale, shT, shL; /* kDeferEmit only: (Emit(l, shT, shL) * shContrib) * shScale is the contribution */
    Hit h;
};

struct MarchState {     /* SphereTracing's locals, shader.comp:780-794 */
    float mt, insT, omega, previousRadius, tMax, ksign;
    float probe, nrm0, nrm1, nrm2;
    int points, iter, sub;
    SdfSet set1;
};

PT_DEV void PathStateInit(PathState& ps) {
    const V3 z3 = mk3(0.0f, 0.0f, 0.0f);
    const V4 z4 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    ps.ray.origin = z3; ps.ray.dir = z3; ps.l = z4; ps.radiance = z4; ps.rayradiance = z4;
    ps.MISBRDFWeight = 1.0f; ps.seed = 0u; ps.bounce = 0;
    ps.isShadow = false; ps.pathAlive = false; ps.pendingFinish = false; ps.inside = false;
    ps.shDir = z3; ps.traceDir = z3; ps.shContrib = z4; ps.shObj = 0; ps.shScale = 0.0f; ps.shT = 0.0f; ps.shL = 0.0f;
    ps.h.t = 1e5f; ps.h.normal = z3; ps.h.materialID = 0.0f; ps.h.lightID = -1.0f; ps.h.objectID = -1;
}
PT_DEV void MarchStateInit(MarchState& ms) {
    ms.mt = 0.0f; ms.insT = 0.0f; ms.omega = 1.7f; ms.previousRadius = 0.0f; ms.tMax = 1e5f; ms.ksign = 0.0f;
    ms.probe = 0.0f; ms.nrm0 = 0.0f; ms.nrm1 = 0.0f; ms.nrm2 = 0.0f;
    ms.points = 0; ms.iter = 0; ms.sub = PT_SUB_SIGN; SdfSetClear(ms.set1);
}

/* Scene()'s tail, shader.comp:1477-1489: radiance -> XYZ, NaNs dropped.  The four table look-ups run through one
 * rolled loop body (instruction-cache footprint); the sum keeps the shader's order
 * ((rad.x*W(l.x) + rad.y*W(l.y)) + rad.z*W(l.z)) + rad.w*W(l.w), the first product initialising it. */
PT_DEV V3 PathColorOf(const Ctx& c, V4 l, V4 r) {
    V3 sum = mk3(0.0f, 0.0f, 0.0f);
#if PT_HAS_SDF && !PT_PATHCOLOR_UNROLL /* rolled where the kernel outgrows the instruction cache; unrolled (no rotations) where it does not */
#pragma unroll 1
#else
#pragma unroll
#endif
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
PT_DEV V3 PathColor(const Ctx& c, const PathState& ps) { return PathColorOf(c, ps.l, ps.radiance); }

/* NEW: Scene() up to TracePath, shader.comp:1446-1472, for sample index pr.firstSample + k of pixel (xyx, xyy), in two
 * halves.  PhaseNewCamera is everything that depends on (pixel, sample index) alone -- seed, sensor jitter, aperture
 * sample, hero wavelength, the two refractions through the camera lens: 350 instructions every lane executes identically.
 * The pooled drivers can take its result from a record a generation kernel wrote beforehand (PT_PREGEN, pt_gen_body):
 * there all 32 lanes are busy, here only the lanes that happen to start a sample.  PhaseNewFinish resets the path.
 * Returns the next phase (ISECT, or NEW again with pendingFinish when pathLength <= 0). */
PT_DEV void PhaseNewCamera(const Ctx& c, unsigned xyx, unsigned xyy, int k, Ray& ray, float& l_h, unsigned& seedOut) {
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
    l_h = 360.0f * (1.0f - r5) + 800.0f * r5; /* mix(360, 800, r) */
    TracePathLens(c, l_h, ray);
    seedOut = seed;
}
PT_DEV int PhaseNewFinish(const Ctx& c, PathState& ps, const Ray& ray, float l_h, unsigned seed) {
    ps.ray = ray;
    ps.traceDir = ray.dir;
    ps.l = SampleWavelengths(l_h);
    ps.seed = seed;
    ps.radiance = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    ps.rayradiance = mk4(1.0f, 1.0f, 1.0f, 1.0f);
    ps.MISBRDFWeight = 1.0f;
    ps.bounce = 0;
    ps.isShadow = false;
    ps.inside = false;
    if (c.pr->pathLength > 0) return PT_ST_ISECT;
    ps.pendingFinish = true;
    return PT_ST_NEW;
}
PT_DEV int PhaseNew(const Ctx& c, PathState& ps, unsigned xyx, unsigned xyy, int k) {
    Ray ray;
    float l_h;
    unsigned seed;
    PhaseNewCamera(c, xyx, xyy, k, ray, l_h, seed);
    return PhaseNewFinish(c, ps, ray, l_h, seed);
}

#ifndef PT_PREGEN
#define PT_PREGEN 0 /* 1: the pooled drivers (v2s, v3s) read PhaseNewCamera's result from pr.gen instead of computing it */
#endif
/* Generation kernel of PT_PREGEN: the launch geometry of the render kernel (one warp = one 8x4 tile), lane p computes
 * the camera rays of pixel p of its tile for every sample of the dispatch -- no divergence, no table, 30 registers.
 * Record layout: PtDevParams::gen.  32 B per sample written once and read once: at 12 Gsamples/s 0.8 TB/s of the 7.7. */
__device__ __forceinline__ void pt_gen_body(const PtDevParams& pr, float4* __restrict__ gen) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4;
    const int qx = tileX + (lane & 7), qy = tileY + (lane >> 3);
    if (qx >= pr.width || qy >= pr.height) return;
    Ctx c;
    c.sc = nullptr; c.pr = &pr; c.ubo = nullptr; c.s_tab = nullptr;
    const unsigned long long tile = ((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * (PT_BLOCK_THREADS / 32) + (unsigned)warp;
    float4* a = gen + tile * 32ull * (unsigned long long)pr.samplesPerFrame + (unsigned)lane;
    float4* b = a + pr.genCount;
#pragma unroll 1
    for (int k = 0; k < pr.samplesPerFrame; k++) {
        Ray ray;
        float l_h;
        unsigned seed;
        PhaseNewCamera(c, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, k, ray, l_h, seed);
        a[32 * k] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x);
        b[32 * k] = make_float4(ray.dir.y, ray.dir.z, l_h, __uint_as_float(seed));
    }
}
#ifndef PT_RESOLVE
#define PT_RESOLVE 0 /* 1 (with PT_PREGEN): finished samples leave their radiance bundle in pr.rad; pt_resolve_body projects and sums */
#endif
/* Resolve kernel of PT_RESOLVE: Scene()'s tail (radiance -> XYZ, shader.comp:1477-1489) and Rendering()'s sum over the
 * samples (1492-1533) for pixel p of the warp's tile, in SAMPLE ORDER -- so the image does not depend on the schedule in
 * any mode -- with all lanes busy, where the render kernel ran the four table look-ups for whichever lanes had just
 * finished a path (17 of 32 on cfg5, 8 % of its instructions).  The wavelength bundle is rebuilt from the hero
 * wavelength of the generation record. */
__device__ __forceinline__ void pt_resolve_body(const PtDevParams& pr, const float* __restrict__ ubo, float4* __restrict__ image,
                                                float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (gx >= pr.width || gy >= pr.height) return;
    Ctx c;
    c.sc = nullptr; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;
    const unsigned long long tile = ((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * (PT_BLOCK_THREADS / 32) + (unsigned)warp;
    const unsigned long long r0 = tile * 32ull * (unsigned long long)pr.samplesPerFrame + (unsigned)lane;
    const float4* rad = pr.rad + r0;
    const float4* rec = pr.gen + pr.genCount + r0; /* (dir.y, dir.z, hero wavelength, seed) */
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
#pragma unroll 1
    for (int k = 0; k < pr.samplesPerFrame; k++) {
        const float4 R = __ldg(rad + 32 * k);
        const float l_h = __ldg(rec + 32 * k).z;
        outColor = outColor + PathColorOf(c, SampleWavelengths(l_h), mk4(R.x, R.y, R.z, R.w));
    }
    StoreTexel(pr, image, gx, gy, outColor);
}
#if PT_RESOLVE
PT_DEV void ResolveDeposit(const PtDevParams& pr, unsigned long long tileRecord0, int k, int q, V4 radiance) {
    pr.rad[tileRecord0 + 32ull * (unsigned)k + (unsigned)q] = make_float4(radiance.x, radiance.y, radiance.z, radiance.w);
}
#endif
#if PT_PREGEN
/* NEW for item (sample k of the dispatch, pixel q of the warp's tile) from its record */
PT_DEV int PhaseNewFromRecord(const Ctx& c, PathState& ps, unsigned long long tileRecord0, int k, int q) {
    const float4* a = c.pr->gen + tileRecord0 + 32ull * (unsigned)k + (unsigned)q;
    const float4 A = __ldg(a), B = __ldg(a + c.pr->genCount);
    Ray ray;
    ray.origin = mk3(A.x, A.y, A.z);
    ray.dir = mk3(A.w, B.x, B.y);
    return PhaseNewFinish(c, ps, ray, B.z, __float_as_uint(B.w));
}
#endif


#if PT_HAS_SDF
/* SphereTracing's prologue (shader.comp:780-801): the bounding boxes the ray meets.  True = there is a march to run. */
PT_DEV bool MarchBegin(const Ctx& c, V3 origin, V3 dir, MarchState& ms) {
    const V3 invdir = mk3(PTK_DIV(1.0f, dir.x), PTK_DIV(1.0f, dir.y), PTK_DIV(1.0f, dir.z));
    float tMin = 1e5f;
    ms.tMax = 1e5f;
    if (!SearchSDF(c, origin, invdir, tMin, ms.tMax, ms.set1)) return false;
    ms.mt = PTK_MAX(tMin, 1e-3f);
    ms.insT = 0.0f; ms.omega = 1.70f; ms.previousRadius = 0.0f; ms.points = 0; ms.iter = 0;
    ms.sub = PT_SUB_SIGN;
    return true;
}
#endif

/* ISECT: Intersection / LightSourceVisibilityCheck up to (and with) SphereTracing's prologue */
PT_DEV int PhaseIsect(const Ctx& c, PathState& ps, MarchState& ms) {
    Ray r;
    r.origin = ps.ray.origin;
    r.dir = ps.traceDir;
    IntersectionAnalytic(c, r, ps.h, ps.isShadow);
#if PT_HAS_SDF
    if (MarchBegin(c, r.origin, r.dir, ms)) return PT_ST_SDF;
#endif
    return PT_ST_SHADE;
}

#if !PT_BVH
/* ISECT in two phases (option heavy_min): the light half returns the phase that follows the intersection (SDF / SHADE)
 * and the candidates of the heavy half; ms.points / ms.iter carry them while the lane waits for the HEAVY phase */
PT_DEV int PhaseIsectLight(const Ctx& c, PathState& ps, MarchState& ms, unsigned& cand) {
    Ray r;
    r.origin = ps.ray.origin;
    r.dir = ps.traceDir;
    cand = IntersectLight(c, r, ps.h, ps.isShadow);
#if PT_HAS_SDF
    if (MarchBegin(c, r.origin, r.dir, ms)) return PT_ST_SDF;
#endif
    return PT_ST_SHADE;
}
PT_DEV void PhaseHeavy(const Ctx& c, PathState& ps, unsigned cand) {
    Ray r;
    r.origin = ps.ray.origin;
    r.dir = ps.traceDir;
    IntersectHeavy(c, r, ps.h, ps.isShadow, cand);
}
#endif

#if PT_HAS_SDF
/* The SDF phase in pieces, so a driver can decide WHO evaluates the distance function at WHICH point:
 * SdfMarchPoint (where the march wants its next evaluation), SdfProbePoint (the j-th of the six normal probes around
 * p: +x -x +y -y +z -z, CalculateNumericalSDFNormals shader.comp:721-730) and SdfMarchConsume (SphereTracing's
 * bookkeeping for one evaluated distance, shader.comp:801-858). */
PT_DEV V3 SdfMarchPoint(const PathState& ps, const MarchState& ms) {
    if (ms.sub == PT_SUB_SIGN) return ps.ray.origin; /* k = sign(SDF(ray.origin)), shader.comp:801 */
    return fma3(ps.traceDir, ms.mt, ps.ray.origin);
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
    const V3 dir = ps.traceDir;
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

/* SDF: one SDF() evaluation and the bookkeeping around it.  Returns SDF (more to do) or SHADE.
 * This is the ONE SDF() site of a megakernel: sign probe, march steps and the six normal probes funnel through it. */
PT_DEV int PhaseSdfEval(const Ctx& c, PathState& ps, MarchState& ms) {
    const V3 pm = SdfMarchPoint(ps, ms);
    const V3 pe = (ms.sub >= PT_SUB_N0) ? SdfProbePoint(pm, ms.sub - PT_SUB_N0) : pm;
    const float d = SDF(pe, ms.set1);
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
        ps.h.materialID = ::pt_sdfmaterial_dispatch(p.x, p.y, p.z, PT_SDF_SET_ARGS(ms.set1));
        ps.h.lightID = -1.0f;
        return PT_ST_SHADE;
    }
    return PT_ST_SDF;
}
#endif

/* The two outcomes of a traced ray that need no shading work -- the verdict of a shadow ray and a path ray that left
 * the scene (shader.comp:1218-1222,1328-1334 and 1387-1389).  Returns the next phase, or SHADE when the hit needs
 * shading. */
PT_DEV int PhaseTrivial(PathState& ps) {
    if (ps.isShadow) {
        if (ps.h.objectID == ps.shObj) ps.radiance = ps.radiance + ps.shContrib;
        ps.isShadow = false;
        ps.traceDir = ps.ray.dir; /* the path ray that was waiting behind the shadow ray */
        if (ps.pathAlive) return PT_ST_ISECT;
        ps.pendingFinish = true;
        return PT_ST_NEW;
    }
    if (!(ps.h.t < 1e5f)) { /* miss: black environment */
        ps.pendingFinish = true;
        return PT_ST_NEW;
    }
    return PT_ST_SHADE;
}

#if PT_EXT_BSDF
/* ---- surface extensions --------------------------------------------------------------------------------
 * NOT reference behaviour: shader.comp shades every surface as a Lambertian (1075-1091) and its TODO.md:2 lists
 * "Specular, Glossy Materials, Etc" as future work, so nothing here can be parity-checked against the reference.  The
 * checker is oracle/oracle.cpp's own statement of the same definitions plus the physical
 * invariants of tests/test_surface_ext.py.  A scene without extensions compiles none of this.
 *   mirror      o = reflect(d, n); throughput *= R(lambda); no light sampling (delta lobe)
 *   glossy      GGX conductor, alpha = roughness^2, Schlick Fresnel with F0 = R(lambda), Smith G1*G1; the half vector is
 *               drawn from D(h)(n.h); light sampling + the reference's MIS weighting with this lobe's pdf
 *   dielectric  exact unpolarised Fresnel at eta = n1/n2 (n = ior, or BK7 at the hero wavelength l.w like the camera
 *               lens: the bundle follows the hero's direction); reflect with probability F, else refract and tint by
 *               R(lambda); `inside` toggles at every refraction
 * R(lambda) = EvaluateBRDF * PI = the material's spectrum; n is turned to face the incoming ray first. */
PT_DEV PtDevSurfaceExt SurfaceExtOf(const Ctx& c, float materialID) {
    PtDevSurfaceExt e;
    e.bsdf = 0; e.roughness = 0.0f; e.ior = 0.0f; e.pad = 0.0f;
    const int m = __float2int_rz(floorf(materialID));
    if (m >= 0 && m < c.sc->nSurfaceExt) e = c.sc->surfaceExt[m];
    return e;
}
PT_DEV V3 Reflect3(V3 I, V3 N) { /* GLSL 4.50 8.5: I - 2 dot(N, I) N */
    const float d = 2.0f * dot(N, I);
    return mk3(I.x - d * N.x, I.y - d * N.y, I.z - d * N.z);
}
PT_DEV float GgxD(float nh, float a2) {
    const float q = nh * nh * (a2 - 1.0f) + 1.0f;
    return PTK_DIV(a2, PT_PI_F * (q * q));
}
PT_DEV float GgxG1(float nv, float a2) { return PTK_DIV(2.0f * nv, nv + PTK_SQRT(a2 + (1.0f - a2) * (nv * nv))); }
PT_DEV V4 SchlickF(V4 f0, float ih) {
    const float m = PTK_MIN(PTK_MAX(1.0f - ih, 0.0f), 1.0f);
    const float m2 = m * m;
    const float m5 = m2 * m2 * m;
    return mk4(f0.x + (1.0f - f0.x) * m5, f0.y + (1.0f - f0.y) * m5, f0.z + (1.0f - f0.z) * m5, f0.w + (1.0f - f0.w) * m5);
}
PT_DEV float GgxAlpha2(float roughness) {
    const float a = roughness * roughness;
    return PTK_MAX(a * a, 1e-8f);
}
/* the GGX lobe f(i, o) without the cosine; i = -ray direction, all three on the side of n */
PT_DEV V4 GgxEval(V3 i, V3 o, V3 n, float a2, V4 f0) {
    const float ni = dot(n, i), no = dot(n, o);
    if (!(ni > 0.0f) || !(no > 0.0f)) return mk4(0.0f, 0.0f, 0.0f, 0.0f);
    const V3 h = normalize(i + o);
    const float nh = dot(n, h), ih = dot(i, h);
    const float k = PTK_DIV(GgxD(nh, a2) * (GgxG1(ni, a2) * GgxG1(no, a2)), 4.0f * ni * no);
    return SchlickF(f0, ih) * k;
}
/* One scattering event of an extended surface.  In: ray direction d, facing normal n (dot(d, n) <= 0), spectrum R.
 * Out: the next direction, the throughput factor f cos / pdf, the lobe's pdf for the MIS weight (0 for the delta lobes),
 * whether the path dies here (sampled below the surface).  Draws: mirror none, glossy 2, dielectric 1. */
PT_DEV void SurfaceExtSample(const PtDevSurfaceExt& ext, V3 d, V3 n, V4 R, V4 l, unsigned& seed, bool& inside, V3& outDir, V4& weight,
                             float& pdf, bool& dead) {
    pdf = 0.0f;
    dead = false;
    weight = R;
    if (ext.bsdf == 1) {
        outDir = Reflect3(d, n);
    } else if (ext.bsdf == 2) {
        const float a2 = GgxAlpha2(ext.roughness);
        const float u1 = RandomFloatPCG32(seed);
        const float u2 = RandomFloatPCG32(seed);
        const float cos2 = PTK_DIV(1.0f - u1, 1.0f + (a2 - 1.0f) * u1);
        const float cosT = PTK_SQRT(cos2);
        const float sinT = PTK_SQRT(PTK_MAX(1.0f - cos2, 0.0f));
        const float phi = 2.0f * PT_PI_F * u2;
        const V3 h = ToWorld(mk3(PTK_COS(phi) * sinT, PTK_SIN(phi) * sinT, cosT), n);
        outDir = Reflect3(d, h);
        const V3 i = -d;
        const float ni = dot(n, i), no = dot(n, outDir), nh = dot(n, h), ih = dot(i, h);
        if (!(no > 0.0f) || !(ih > 0.0f) || !(ni > 0.0f)) {
            dead = true;
            weight = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        } else {
            pdf = PTK_DIV(GgxD(nh, a2) * nh, 4.0f * ih);
            weight = SchlickF(R, ih) * PTK_DIV((GgxG1(ni, a2) * GgxG1(no, a2)) * ih, ni * nh);
        }
    } else {
        const float ng = (ext.ior > 0.0f) ? ext.ior : RefractiveIndexBK7Glass(l.w);
        const float n1 = inside ? ng : 1.0f, n2 = inside ? 1.0f : ng;
        const float eta = PTK_DIV(n1, n2);
        const float cosi = -dot(d, n);
        const float sin2t = eta * eta * (1.0f - cosi * cosi);
        float F = 1.0f, cost = 0.0f;
        if (sin2t < 1.0f) {
            cost = PTK_SQRT(1.0f - sin2t);
            const float rs = PTK_DIV(n1 * cosi - n2 * cost, n1 * cosi + n2 * cost);
            const float rp = PTK_DIV(n2 * cosi - n1 * cost, n2 * cosi + n1 * cost);
            F = 0.5f * (rs * rs + rp * rp);
        }
        const float u = RandomFloatPCG32(seed);
        if (u < F) {
            outDir = Reflect3(d, n);
            weight = mk4(1.0f, 1.0f, 1.0f, 1.0f);
        } else {
            const float k = eta * cosi - cost;
            outDir = normalize(mk3(eta * d.x + k * n.x, eta * d.y + k * n.y, eta * d.z + k * n.z));
            inside = !inside;
        }
    }
}
#endif /* PT_EXT_BSDF */

/* SHADE: TraceRay after Intersection (shader.comp:1352-1390) with SampleLightSource (1298-1343) for a path ray that hit
 * something -- THE statement of the bounce arithmetic, shared by every driver and by the wavefront pipeline.
 * The light sample's visibility test is not traced here: the sampled direction, the object to look for and the
 * contribution it would add are left in shDir / shObj / shContrib with isShadow set (LightSourceVisibilityCheck draws no
 * random numbers and its result only gates that addition, so running it after the roulette below changes nothing).
 * Returns ISECT (another ray to trace: the shadow ray if isShadow, else the next path ray) or NEW (path finished). */
template <bool kDeferEmit>
PT_DEV int PhaseShadeHitT(const Ctx& c, PathState& ps) {
    const PtDevScene& sc = *c.sc;
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
        V3 n = ps.h.normal;
        outOrigin = fma3(ps.ray.dir, ps.h.t, ps.ray.origin);
        float BRDFpdf;
        bool sampleLights = sc.numLights > 0.0f;
#if PT_EXT_BSDF
        const PtDevSurfaceExt ext = SurfaceExtOf(c, ps.h.materialID);
        const V4 spectrum = brdf * PT_PI_F; /* R(lambda) */
        V4 extWeight = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        bool extDead = false;
        if (ext.bsdf != 0) {
            if (dot(ps.ray.dir, n) > 0.0f) n = -n; /* cyclide and SDF normals are not turned towards the ray */
            SurfaceExtSample(ext, ps.ray.dir, n, spectrum, ps.l, ps.seed, ps.inside, outDir, extWeight, BRDFpdf, extDead);
            /* the next ray starts off the surface, on the side it leaves to (the reference's primitives reject hits nearer
             * than 1e-4 and never see a ray that starts on a surface and goes INTO it; SDF hits sit 1e-3 in front) */
            const float side = (dot(outDir, n) > 0.0f) ? 1.0f : -1.0f;
            const float eps = side * ((ps.h.objectID < 0) ? 2e-3f : 2e-4f);
            outOrigin = mk3(outOrigin.x + n.x * eps, outOrigin.y + n.y * eps, outOrigin.z + n.z * eps);
            sampleLights = sampleLights && (ext.bsdf == 2) && !extDead; /* delta lobes: nothing to connect */
        } else
#endif
        {
            outDir = SampleCosineDirectionHemisphere(n, ps.seed);
            BRDFpdf = PTK_DIV(dot(outDir, n), PT_PI_F);
        }
        if (sampleLights) { /* SampleLightSource, shader.comp:1298-1343 */
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
#if PT_EXT_BSDF
                    const V4 fLight = (ext.bsdf == 2) ? GgxEval(-ps.ray.dir, sdir, n, GgxAlpha2(ext.roughness), spectrum) : brdf;
#else
                    const V4 fLight = brdf;
#endif
                    rr = ps.rayradiance * mulDiv4(fLight, costheta, lightpdf);
                    emitScale = 1.0f - ps.MISBRDFWeight;
                    ps.shDir = sdir;
                    ps.shObj = ls.objectID;
                    needShadow = true;
                } else {
                    ps.MISBRDFWeight = 1.0f;
                }
            }
        } else {
#if PT_EXT_BSDF
            if (ext.bsdf != 0) ps.MISBRDFWeight = 1.0f; /* the next emitter hit counts in full */
            else
#endif
            ps.MISBRDFWeight = PTK_DIV(BRDFpdf * BRDFpdf, BRDFpdf * BRDFpdf + 0.0f * 0.0f);
        }
#if PT_EXT_BSDF
        if (ext.bsdf != 0) {
            ps.rayradiance = ps.rayradiance * extWeight;
        } else
#endif
        {
            const float costheta = dot(outDir, n);
            ps.rayradiance = ps.rayradiance * mulDiv4(brdf, costheta, BRDFpdf);
        }
        const V4 t4 = ps.rayradiance;
        const float mx = PTK_MAX(t4.x, PTK_MAX(t4.y, PTK_MAX(t4.z, t4.w)));
        const float rayProbability = PTK_MIN(PTK_MAX(mx, 0.0f), 0.99f);
        alive = !(RandomFloatPCG32(ps.seed) > rayProbability);
        if (alive) ps.rayradiance = ps.rayradiance * PTK_DIV(1.0f, rayProbability);
#if PT_EXT_BSDF
        if (extDead) alive = false;
#endif
        ps.bounce++;
        if (ps.bounce >= c.pr->pathLength) alive = false; /* TracePath's loop bound, shader.comp:1400 */
    }
    if (kDeferEmit && needShadow) {
        /* the drivers that trace the shadow ray right away (TraceRayFlat) evaluate Emit() only if the light is seen:
         * most light samples are occluded or face away.  Same products in the same order: (Emit * rr) * scale. */
        ps.shContrib = rr; ps.shScale = emitScale; ps.shT = emitT; ps.shL = emitL;
    } else if (emitterHit || needShadow) { /* the one Emit() site: (Emit * rr) * scale in both uses */
        const V4 e = Emit(ps.l, PTK_MAX(emitT, 0.0f), PTK_MAX(emitL, 0.0f));
        const V4 contrib = (e * rr) * emitScale;
        if (emitterHit) ps.radiance = ps.radiance + contrib; else ps.shContrib = contrib;
    }
    ps.ray.origin = outOrigin;
    ps.ray.dir = outDir; /* next path direction; a pending shadow ray travels along shDir */
    ps.traceDir = needShadow ? ps.shDir : outDir;
    if (needShadow) {
        ps.isShadow = true;
        ps.pathAlive = alive;
        return PT_ST_ISECT;
    }
    if (emitterHit || !alive) {
        ps.pendingFinish = true;
        return PT_ST_NEW;
    }
    return PT_ST_ISECT;
}
PT_DEV int PhaseShadeHit(const Ctx& c, PathState& ps) { return PhaseShadeHitT<false>(c, ps); }
/* SHADE for callers that have not applied PhaseTrivial themselves (the wavefront pipeline's SHADE kernel) */
PT_DEV int PhaseShade(const Ctx& c, PathState& ps) {
    const int t = PhaseTrivial(ps);
    return (t == PT_ST_SHADE) ? PhaseShadeHit(c, ps) : t;
}

/* ---- the phases run back to back by one lane: the loop bodies of the v1 and v3s drivers -----------------------------
 * shader.comp:862-934 / 1121-1216: Intersection / LightSourceVisibilityCheck in one piece (kShadow is a compile-time
 * constant at both call sites, so the shadow flavour drops normals, materials, probes). */
#if PT_HAS_SDF
PT_DEV_NOINLINE void SphereTracing(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    PathState js;
    js.ray = ray; js.shDir = ray.dir; js.traceDir = ray.dir; js.isShadow = kShadow; js.h = h;
    MarchState ms;
    if (!MarchBegin(c, ray.origin, ray.dir, ms)) return;
    int st = PT_ST_SDF;
#pragma unroll 1
    while (st == PT_ST_SDF) st = PhaseSdfEval(c, js, ms);
    h = js.h;
}
#endif
PT_DEV void Intersection(const Ctx& c, const Ray& ray, Hit& h, const bool kShadow) {
    IntersectionAnalytic(c, ray, h, kShadow);
#if PT_HAS_SDF
    SphereTracing(c, ray, h, kShadow);
#endif
}
/* One iteration of TracePath's loop (shader.comp:1393-1407) on the path held in `ps`: TraceRay (1345-1391) = closest hit,
 * PhaseShadeHit, and the light sample's visibility test when one is pending.  Returns whether the path goes on. */
PT_DEV bool TraceRayFlat(const Ctx& c, PathState& ps) {
    Intersection(c, ps.ray, ps.h, false);
    int next = PhaseTrivial(ps);
    if (next == PT_ST_SHADE) next = PhaseShadeHitT<true>(c, ps);
    if (ps.isShadow) { /* the light sample's visibility test, shader.comp:1121-1223, 1328-1334 */
        Ray sr;
        sr.origin = ps.ray.origin;
        sr.dir = ps.shDir;
        Intersection(c, sr, ps.h, true);
        if (ps.h.objectID == ps.shObj) {
            const V4 e = Emit(ps.l, PTK_MAX(ps.shT, 0.0f), PTK_MAX(ps.shL, 0.0f));
            ps.radiance = ps.radiance + (e * ps.shContrib) * ps.shScale;
        }
        ps.isShadow = false;
        if (ps.pathAlive) {
            next = PT_ST_ISECT;
        } else {
            ps.pendingFinish = true;
            next = PT_ST_NEW;
        }
    }
    return next == PT_ST_ISECT;
}

/* ---- scheduling statistics and knobs of the in-warp drivers ---------------------------------------------------------- */
#ifdef PT_STATS
/* debug builds only (option "stats"): per phase, [2p] = executions, [2p+1] = lanes served; [8..15] = SDF executions by
 * participants (1-4, 5-8, ... 29-32) */
} /* namespace */
extern "C" __device__ unsigned long long pt_stats[16];
namespace PT_KERNEL_NS {
#define PT_STAT(p, mask) do { if ((threadIdx.x & 31) == 0) { atomicAdd(&pt_stats[2 * (p)], 1ull); atomicAdd(&pt_stats[2 * (p) + 1], (unsigned long long)__popc(mask)); } } while (0)
#define PT_STAT_SDF(n) do { if ((threadIdx.x & 31) == 0 && (n) > 0) atomicAdd(&pt_stats[8 + (((n) - 1) >> 2)], 1ull); } while (0)
#else
#define PT_STAT(p, mask) do { } while (0)
#define PT_STAT_SDF(n) do { } while (0)
#endif
#ifndef PT_SDF_REPS
#define PT_SDF_REPS 16 /* SDF() evaluations per execution of the SDF phase */
#endif
#ifndef PT_FEED_T
#define PT_FEED_T 8    /* the SDF phase runs once every feeder phase (NEW / ISECT / SHADE) has fewer lanes than this */
#endif
#ifndef PT_REGEN_T
#define PT_REGEN_T 16  /* v3s: lanes that must wait for a new sample before the warp regenerates */
#endif
#ifndef PT_STEAL_S
#define PT_STEAL_S 16
#endif
/* PT_STEAL_S > 0: the warp's pool holds S samples per pixel at a time and a finished sample's XYZ goes to a per-warp
 * table in shared memory (S x 32 x 3 floats); when the pool is empty lane p adds pixel p's S entries IN SAMPLE ORDER --
 * the additions of the reference's loop in its order, so strict mode stays bit-exact -- and the next round starts.
 * PT_STEAL_S == 0 (fast mode only): ONE round with all samplesPerFrame samples of the dispatch in the pool, and a
 * finished sample is added straight to its pixel's running sum in shared memory (red.shared.add.f32) -- no table,
 * hence no bound on the pool, at the price of a summation order that follows the schedule instead of the sample
 * index (the schedule of a warp is a pure function of its inputs, so renders still repeat bit for bit in practice). */
#if PT_STEAL_S == 0
#define PT_STEAL_WORDS (3 * 32) /* per warp: the running XYZ sums of the tile's pixels */
#else
#define PT_STEAL_WORDS (3 * PT_STEAL_S * 32) /* per warp */
#endif
static_assert(PT_STEAL_S >= 0 && PT_STEAL_S <= 20, "PT_STEAL_S: the per-sample table must fit the 48 KB of static shared memory next to the uniform block");
enum { PT_ST_IDLE = 5, PT_ST_HEAVY = 6 };
#ifndef PT_HEAVY_MIN
#define PT_HEAVY_MIN 0 /* v2s: > 0 runs the box / lens / cyclide tests as a phase of their own, once this many lanes wait for it
                          (or the feeders run dry): fewer, fuller executions of the kernel's coldest 450 instructions */
#endif
#if PT_BVH
#undef PT_HEAVY_MIN
#define PT_HEAVY_MIN 0
#endif

/* the tile's sample pool, shared by v3s / v2s / v2m: deposit a finished sample, close a round */
PT_DEV void PoolDeposit(float* s_col, int item, V3 col) {
#if PT_STEAL_S == 0
    float* e = s_col + (item & 31);
    atomicAdd(e, col.x); atomicAdd(e + 32, col.y); atomicAdd(e + 64, col.z);
#else
    float* e = s_col + (3 * (item >> 5)) * 32 + (item & 31);
    e[0] = col.x; e[32] = col.y; e[64] = col.z;
#endif
}
/* end of a round (every item finished): lane p takes pixel p's sum.  Returns false when the dispatch is complete. */
PT_DEV bool PoolCloseRound(float* s_col, int lane, bool inRange, int spf, int& roundBase, int& roundN, V3& outColor) {
    __syncwarp();
#if PT_STEAL_S == 0
    (void)inRange; (void)spf; (void)roundBase; (void)roundN;
    outColor = mk3(s_col[lane], s_col[32 + lane], s_col[64 + lane]);
    return false;
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
    if (roundBase >= spf) return false;
    roundN = (spf - roundBase) < PT_STEAL_S ? (spf - roundBase) : PT_STEAL_S;
    return true;
#endif
}

/* ---- driver v1 (PT_SCHED 0): one thread = one pixel, nested sample / bounce loops -----------------------------------
 * Grid: 2-D tiles of 16x8 pixels per 128-thread block; each warp owns an 8x4 sub-tile. */
__device__ __forceinline__ void pt_render_body_v1(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                  float4* __restrict__ image, float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int gy = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (gx >= pr.width || gy >= pr.height) return;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const unsigned xyx = (unsigned)gx;
    const unsigned xyy = (unsigned)pr.height - (unsigned)gy; /* shader.comp:1510 */

    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    const int spf = pr.samplesPerFrame;
    PathState ps;
    PathStateInit(ps);
#pragma unroll 1
    for (int k = 0; k < spf; k++) { /* Scene(), shader.comp:1446-1490 */
        if (PhaseNew(c, ps, xyx, xyy, k) == PT_ST_ISECT) {
#pragma unroll 1
            while (TraceRayFlat(c, ps)) { }
        }
        ps.pendingFinish = false;
        outColor = outColor + PathColor(c, ps);
    }
    StoreTexel(pr, image, gx, gy, outColor);
}

/* ---- driver v3s (PT_SCHED 7): v1's loop bodies in ONE flat loop, with the tile's sample pool ------------------------
 * What v1 loses on the analytic scenes is a path's tail: on scene1 nine lanes in ten are done after two rays, yet the
 * warp runs the later bounces for the one to four lanes still alive (ncu on v1: 17 of 32 lanes per instruction).  Here
 * the sample loop and the bounce loop are one loop: an iteration is one bounce (TraceRayFlat) for every lane with a
 * live path, and as soon as PT_REGEN_T lanes wait (or nobody is alive) the waiting lanes deposit their finished sample
 * and claim the next items of the tile's 32 x S pool by ballot rank, whichever pixel they belong to, so the
 * stragglers' late bounces ride along with the next samples' first ones.  Scene() depends only on (pixel, sample
 * index), never on the lane that runs it: bit-exact in strict mode (per-sample table, summed in sample order). */
__device__ __forceinline__ void pt_render_body_v3s(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_colAll) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4;
#if PT_PREGEN
    const unsigned long long tileRecord0 = (((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * (PT_BLOCK_THREADS / 32) + (unsigned)warp) *
                                           32ull * (unsigned long long)pr.samplesPerFrame;
#endif
    const int gx = tileX + (lane & 7), gy = tileY + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    float* s_col = s_colAll + warp * PT_STEAL_WORDS;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const int spf = pr.samplesPerFrame;
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
    PathState ps;
    PathStateInit(ps);
    bool alive = false;

    for (;;) {
        const bool wantNew = !alive && ((next < 32 * roundN) || ps.pendingFinish);
        const unsigned bNew = __ballot_sync(0xffffffffu, wantNew);
        const unsigned bAlive = __ballot_sync(0xffffffffu, alive);
        if ((bNew | bAlive) == 0u) {
            if (!warpLive) break;
            if (!PoolCloseRound(s_col, lane, inRange, spf, roundBase, roundN, outColor)) break;
            next = 0;
            continue;
        }
        if ((__popc(bNew) >= PT_REGEN_T) || (bAlive == 0u)) { /* warp-uniform */
            if (wantNew) {
                if (ps.pendingFinish) { /* Scene()'s tail for the path that ended, shader.comp:1477-1489 */
#if PT_RESOLVE
                    ResolveDeposit(pr, tileRecord0, roundBase + (item >> 5), item & 31, ps.radiance);
#else
                    PoolDeposit(s_col, item, PathColor(c, ps));
#endif
                    ps.pendingFinish = false;
                }
                item = next + __popc(bNew & ((1u << lane) - 1u));
                if (item < 32 * roundN) {
                    const int q = item & 31;
                    const int qx = tileX + (q & 7), qy = tileY + (q >> 3);
                    if ((qx < pr.width) && (qy < pr.height)) { /* else: a pixel beyond the image edge; claim again */
#if PT_PREGEN
                        const int nextState = PhaseNewFromRecord(c, ps, tileRecord0, roundBase + (item >> 5), q);
#else
                        const int nextState = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, roundBase + (item >> 5));
#endif
                        alive = (nextState == PT_ST_ISECT); /* pathLength <= 0: PhaseNew left pendingFinish set */
                    }
                }
            }
            next += __popc(bNew);
        }
        if (alive) { /* one iteration of TracePath's loop, shader.comp:1393-1407 */
            if (!TraceRayFlat(c, ps)) alive = false; /* the phases left pendingFinish set */
        }
    }
#if PT_RESOLVE
    (void)inRange; (void)outColor; (void)image; (void)gx; (void)gy; /* pt_resolve_body writes the texels */
#else
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
#endif
}

/* ---- driver v2s (PT_SCHED 5): one phase per warp iteration, with the tile's sample pool ------------------------------
 * Every lane is a state machine over the four phases; per iteration the warp executes ONE phase chosen by ballot (the
 * fullest feeder first; the SDF phase, PT_SDF_REPS evaluations at a time, once every feeder has fewer than PT_FEED_T
 * lanes), so path and shadow rays share one intersection site and sign probe, march steps and normal probes one SDF()
 * site.  The warp owns the 32 x S samples of its 8x4 tile as a pool of work items (item = 32 * sample + pixel, so the
 * first 32 items are one sample of every pixel and primary rays start coherent); a lane that finishes a path claims the
 * next unclaimed item, whichever pixel it belongs to (claims happen in the NEW phase, which the warp executes
 * converged: rank by ballot, no atomics).  With a lane tied to its pixel's sample sequence a warp lives as long as its
 * most expensive pixel (the mean lane carries 0.35 / 0.64 / 0.72 of the busiest lane's work on menger / mandelbulb /
 * terrain, profiles/r01_steal); the pool balances that. */
__device__ __forceinline__ void pt_render_body_v2s(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_colAll) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4;
#if PT_PREGEN
    const unsigned long long tileRecord0 = (((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * (PT_BLOCK_THREADS / 32) + (unsigned)warp) *
                                           32ull * (unsigned long long)pr.samplesPerFrame;
#endif
    const int gx = tileX + (lane & 7), gy = tileY + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    float* s_col = s_colAll + warp * PT_STEAL_WORDS;

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
#if PT_HEAVY_MIN > 0
        const unsigned bHv = __ballot_sync(0xffffffffu, st == PT_ST_HEAVY);
#else
        const unsigned bHv = 0u;
#endif
        if ((bNew | bIs | bSdf | bSh | bHv) == 0u) {
            if (!warpLive) break;
            if (!PoolCloseRound(s_col, lane, inRange, spf, roundBase, roundN, outColor)) break;
            next = 0;
            st = PT_ST_NEW;
            continue;
        }
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
#if PT_HAS_SDF
        if (bSdf != 0u && best < PT_FEED_T) phase = PT_ST_SDF;
#endif
#if PT_HEAVY_MIN > 0
        if (bHv != 0u && (__popc(bHv) >= PT_HEAVY_MIN || best < PT_FEED_T)) phase = PT_ST_HEAVY;
#endif
        PT_STAT(phase == PT_ST_HEAVY ? PT_ST_ISECT : phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? bSdf : bSh)));
        const bool ran = (st == phase) && (phase == PT_ST_ISECT || phase == PT_ST_SDF);
        if (phase == PT_ST_NEW) {
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) { /* the sample this lane just finished: item -> (pixel, sample of the round) */
#if PT_RESOLVE
                    ResolveDeposit(pr, tileRecord0, roundBase + (item >> 5), item & 31, ps.radiance);
#else
                    PoolDeposit(s_col, item, PathColor(c, ps));
#endif
                    ps.pendingFinish = false;
                }
                item = next + __popc(bNew & ((1u << lane) - 1u));
                if (item < 32 * roundN) {
                    const int q = item & 31;
                    const int qx = tileX + (q & 7), qy = tileY + (q >> 3);
                    if ((qx < pr.width) && (qy < pr.height)) { /* else: a pixel beyond the image edge; claim again */
#if PT_PREGEN
                        st = PhaseNewFromRecord(c, ps, tileRecord0, roundBase + (item >> 5), q);
#else
                        st = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, roundBase + (item >> 5));
#endif
                    }
                } else {
                    st = PT_ST_IDLE;
                }
            }
            next += __popc(bNew);
        } else if (phase == PT_ST_ISECT) {
            if (st == PT_ST_ISECT) {
#if PT_HEAVY_MIN > 0
                unsigned cand;
                st = PhaseIsectLight(c, ps, ms, cand);
                if (cand != 0u) { /* wait for the HEAVY phase; what follows it and who is to be tested ride in ms */
                    ms.iter = (st == PT_ST_SDF) ? 0 : -1;
                    ms.points = (int)cand;
                    st = PT_ST_HEAVY;
                } else if (st == PT_ST_SHADE) {
                    st = PhaseTrivial(ps);
                }
#else
                st = PhaseIsect(c, ps, ms);
#endif
            }
        }
#if PT_HEAVY_MIN > 0
        else if (phase == PT_ST_HEAVY) {
            if (st == PT_ST_HEAVY) {
                PhaseHeavy(c, ps, (unsigned)ms.points);
                st = (ms.iter == 0) ? PT_ST_SDF : PT_ST_SHADE;
                ms.points = 0;
                ms.iter = 0;
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
            }
        }
#endif
#if PT_HAS_SDF
        else if (phase == PT_ST_SDF) {
            PT_STAT_SDF(__popc(bSdf));
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (st == PT_ST_SDF) st = PhaseSdfEval(c, ps, ms);
            }
        }
#endif
        else {
            if (st == PT_ST_SHADE) st = PhaseShadeHit(c, ps);
        }
        /* A traced ray's verdict, at ONE site for the ISECT and SDF phases (instruction-cache footprint): only lanes
         * that took part in this phase can have arrived at SHADE just now. */
        if (ran && st == PT_ST_SHADE) st = PhaseTrivial(ps);
    }
#if PT_RESOLVE
    (void)inRange; (void)outColor; (void)image; (void)gx; (void)gy; /* pt_resolve_body writes the texels */
#else
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
#endif
}


#ifndef PT_SCHED
#define PT_SCHED 0
#endif
#if PT_SCHED == 8 && PT_HAS_SDF && PT_SDF_WORDS > 1
#error "v2m parks one mask word per path: more than 32 SDFs need the v2s driver (pt_jit.cpp selects it)"
#endif
#if PT_SCHED == 8 && !PT_HAS_SDF
#undef PT_SCHED
#define PT_SCHED 5 /* nothing marches: v2m is v2s */
#endif
#if PT_SCHED == 8
#include "pt_driver_v2m.cuh" /* the pool of parked marching paths: measured slower, kept out of the default drivers' way */
#endif

} /* namespace PT_KERNEL_NS */

/* the kernel entry point; the name distinguishes the strict / fast / JIT instances */
#if PT_SCHED == 8
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        __shared__ float s_col[(PT_STEAL_WORDS + PT_POOL_WORDS) * (PT_BLOCK_THREADS / 32)];                  \
        PT_KERNEL_NS::pt_render_body_v2m(sc, pr, ubo, image, s_tab, s_col);                                  \
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
        __shared__ float s_col[PT_STEAL_WORDS * (PT_BLOCK_THREADS / 32)];                                    \
        PT_KERNEL_NS::pt_render_body_v2s(sc, pr, ubo, image, s_tab, s_col);                                  \
    }
#elif PT_SCHED == 0
#define PT_DEFINE_RENDER_KERNEL(name)                                                                        \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                            \
    name(const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr,                      \
         const float* __restrict__ ubo, float4* __restrict__ image) {                                        \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        PT_KERNEL_NS::pt_render_body_v1(sc, pr, ubo, image, s_tab);                                          \
    }
#else
#error "PT_SCHED: 0 (v1), 5 (v2s), 7 (v3s) or 8 (v2m)"
#endif

/* the generation kernel of option "pregen" (same grid and block as the render kernel) */
#define PT_DEFINE_GEN_KERNEL(name)                                                                           \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS)                                           \
    name(const __grid_constant__ PtDevParams pr, float4* __restrict__ gen) { PT_KERNEL_NS::pt_gen_body(pr, gen); }

/* the resolve kernel of option "resolve" (same grid and block again) */
#define PT_DEFINE_RESOLVE_KERNEL(name)                                                                       \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS)                                           \
    name(const __grid_constant__ PtDevParams pr, const float* __restrict__ ubo, float4* __restrict__ image) { \
        __shared__ float s_tab[PT_SH_FLOATS];                                                                \
        PT_KERNEL_NS::pt_resolve_body(pr, ubo, image, s_tab);                                                \
    }

#if PT_HAS_SDF
/* SDF()/SDFMATERIAL() at arbitrary points: used by the tests to compare the NVRTC build of the snippets with the
 * CPU oracle's g++ build bit for bit (pt_sdf_eval) */
#define PT_DEFINE_SDF_EVAL_KERNEL(name)                                                                      \
    extern "C" __global__ void name(const float* __restrict__ xyz, unsigned long long n, unsigned set1,      \
                                    unsigned set2, unsigned set3, unsigned set4,                             \
                                    float* __restrict__ dist, float* __restrict__ material) {                \
        const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;              \
        if (i >= n) return;                                                                                  \
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];                                  \
        if (dist) dist[i] = ::pt_sdf_dispatch(x, y, z, set1, set2, set3, set4);                                            \
        if (material) material[i] = ::pt_sdfmaterial_dispatch(x, y, z, set1, set2, set3, set4);                            \
    }
#endif

#endif /* PT_KERNEL_CUH */
