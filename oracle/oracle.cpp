/* oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A literal, scalar CPU restatement of the reference's compute shader src/shader.comp (1533 lines of GLSL), used
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker the
 * CUDA path is compared against.  Nothing under pathtracer_b200/ links or calls this file.
 *
 * PARITY PIN STATUS: "parity unpinned by reference tests".  The reference ships no tests, golden images or
 * known-answer vectors and cannot be compiled or run in this image (no Vulkan loader/ICD,
 * no glslang: SURVEY.md section 0-3), so this restatement is pinned by (a) line-by-line correspondence with
 * shader.comp -- every function below carries the shader.comp line range it follows -- and (b) the first-
 * principles known-answer tests of SURVEY.md App. E (tests/test_oracle_kat.py).
 *
 * Float semantics are the canonical ones of SURVEY.md App. F: each GLSL operator is the correctly rounded
 * binary32 operation in source order, no contraction except where the source says fma(); vector operations are
 * componentwise; dot() accumulates left to right; length = sqrt(dot); normalize = v / length(v);
 * min(x,y) = y<x ? y : x; max(x,y) = x<y ? y : x; transcendentals come from include/pt_math.h.
 * Build with -ffp-contract=off -fno-fast-math (oracle/Makefile) or the above does not hold.
 *
 * Citations `shader.comp:N` are line numbers of /root/reference/src/shader.comp; `host:N` of src/pathtracer.cpp.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pt_abi.h"
#include "pt_math.h"

#ifdef PT_ORACLE_BVH
/* liboracle_bvh.so only: a CHECKER OF THE PRODUCT'S BVH (pathtracer_b200/csrc/pt_bvh.h, pt_bvh.cpp, which the
 * Makefile compiles in).  The closest-hit search runs through that tree with this file's own primitive routines at
 * the leaves; tests/test_bvh.py requires the rendered images to equal liboracle.so's (the literal in-order scan of
 * shader.comp:862-934, 1121-1216) bit for bit.  liboracle.so itself contains none of this. */
#include <cmath>
#include <string>
#include <vector>
#include "pt_internal.h"
#include "pt_bvh.h"
static std::vector<float> g_bvh_blob;
#endif

#include <vector>
/* Surface extensions (include/pt_abi.h pt_surface_ext): NOT reference behaviour -- shader.comp has one
 * surface model, the Lambertian of lines 1075-1091, and the reference's TODO.md:2 lists specular / glossy materials as
 * future work.  PARITY UNPINNED for everything reached through this table: the reference holds nothing to check it
 * against.  What follows is this oracle's own statement of the definitions in DESIGN.md section 4 (the CUDA code states
 * them a second time; the two are compared bit for bit), and tests/test_surface_ext.py holds it to the physics:
 * reflection and Snell's law, Fresnel limits, energy bounds.  With an empty table (the default) none of it is reached. */
static std::vector<pt_surface_ext> g_surface_ext;

namespace {

/* ------------------------------------------------------------------------------------------------------------
 * operation counters; compiled in with -DPT_COUNT
 * ---------------------------------------------------------------------------------------------------------- */
enum {
    C_SAMPLES, C_RAYS_PATH, C_RAYS_SHADOW, C_SPHERE, C_SPHERE_HIT, C_PLANE, C_PLANE_HIT, C_BSPHERE, C_BOX, C_BOX_HIT,
    C_LENS, C_SLICE_HIT, C_CYCLIDE, C_CYCLIDE_3ROOT, C_CYCLIDE_HIT, C_SEARCHSDF, C_ST_CALLS, C_ST_ENTER, C_ST_ITER,
    C_ST_BACKSTEP, C_ST_HIT, C_SDF_EVAL, C_SDFMAT_EVAL, C_BOUNCE, C_EMIT_HIT, C_LIGHT_SAMPLE, C_LIGHT_VISIBLE,
    C_RNG, C_MISS, C_ST_ITER_BEYOND, C_ST_ENTER_BEYOND, C_N
};
struct Counters { unsigned long long v[C_N]; };
#ifdef PT_COUNT
thread_local Counters* g_cnt = nullptr;
#define CNT(name, n) do { if (g_cnt) g_cnt->v[name] += (n); } while (0)
#else
#define CNT(name, n) do { } while (0)
#endif

/* ------------------------------------------------------------------------------------------------------------
 * GLSL value types and builtins, App. F semantics
 * ---------------------------------------------------------------------------------------------------------- */
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M3 { V3 c[3]; }; /* column-major like GLSL: c[col] */

inline V2 v2(float x, float y) { V2 r = {x, y}; return r; }
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 v3(float s) { V3 r = {s, s, s}; return r; }
inline V4 v4(float x, float y, float z, float w) { V4 r = {x, y, z, w}; return r; }
inline V4 v4(float s) { V4 r = {s, s, s, s}; return r; }

inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float gstep(float e, float x) { return (x < e) ? 0.0f : 1.0f; }
inline float gmod(float x, float y) { return x - y * pt_floor(x / y); }
inline float gabs(float x) { return pt_abs(x); }
inline float gfma(float a, float b, float c) { return pt_fma(a, b, c); }
inline float gsqrt(float x) { return pt_sqrt(x); }
inline float gsin(float x) { return pt_sin(x); }
inline float gcos(float x) { return pt_cos(x); }
inline float gacos(float x) { return pt_acos(x); }
inline float gexp(float x) { return pt_exp(x); }
inline float gpow(float x, float y) { return pt_pow(x, y); }

inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 operator/(float s, V3 a) { return v3(s / a.x, s / a.y, s / a.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator+(V3 a, float s) { return v3(a.x + s, a.y + s, a.z + s); }
inline V3 operator-(V3 a, float s) { return v3(a.x - s, a.y - s, a.z - s); }
inline V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
inline V2 operator*(V2 a, V2 b) { return v2(a.x * b.x, a.y * b.y); }
inline V2 operator/(V2 a, V2 b) { return v2(a.x / b.x, a.y / b.y); }
inline V2 operator*(V2 a, float s) { return v2(a.x * s, a.y * s); }
inline V2 operator*(float s, V2 a) { return v2(s * a.x, s * a.y); }
inline V2 operator/(V2 a, float s) { return v2(a.x / s, a.y / s); }
inline V2 operator+(V2 a, float s) { return v2(a.x + s, a.y + s); }
inline V2 operator-(V2 a, float s) { return v2(a.x - s, a.y - s); }
inline V4 operator+(V4 a, V4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline V4 operator-(V4 a, V4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline V4 operator*(V4 a, V4 b) { return v4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline V4 operator/(V4 a, V4 b) { return v4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline V4 operator*(V4 a, float s) { return v4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline V4 operator*(float s, V4 a) { return v4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline V4 operator/(V4 a, float s) { return v4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline V4 operator/(float s, V4 a) { return v4(s / a.x, s / a.y, s / a.z, s / a.w); }
inline V4 operator+(V4 a, float s) { return v4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline V4 operator+(float s, V4 a) { return v4(s + a.x, s + a.y, s + a.z, s + a.w); }
inline V4 operator-(V4 a, float s) { return v4(a.x - s, a.y - s, a.z - s, a.w - s); }
inline V4 operator-(float s, V4 a) { return v4(s - a.x, s - a.y, s - a.z, s - a.w); }
inline V4 operator-(V4 a) { return v4(-a.x, -a.y, -a.z, -a.w); }

inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(V3 a) { return gsqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 vfma(V3 a, V3 b, V3 c) { return v3(gfma(a.x, b.x, c.x), gfma(a.y, b.y, c.y), gfma(a.z, b.z, c.z)); }
inline V3 vmin(V3 a, V3 b) { return v3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
inline V3 vmax(V3 a, V3 b) { return v3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
inline V3 vabs(V3 a) { return v3(gabs(a.x), gabs(a.y), gabs(a.z)); }
inline V3 vsign(V3 a) { return v3(gsign(a.x), gsign(a.y), gsign(a.z)); }
inline V3 vsin(V3 a) { return v3(gsin(a.x), gsin(a.y), gsin(a.z)); }
inline V3 vcos(V3 a) { return v3(gcos(a.x), gcos(a.y), gcos(a.z)); }
inline V3 vstep(float e, V3 x) { return v3(gstep(e, x.x), gstep(e, x.y), gstep(e, x.z)); }
inline V3 vmix(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
inline V3 faceforward(V3 N, V3 I, V3 Nref) { return (dot(Nref, I) < 0.0f) ? N : -N; }
inline V3 refract(V3 I, V3 N, float eta) {
    float k = 1.0f - eta * eta * (1.0f - dot(N, I) * dot(N, I));
    if (k < 0.0f) return v3(0.0f);
    return eta * I - (eta * dot(N, I) + gsqrt(k)) * N;
}
/* swizzles used by the shader */
inline V3 xzy(V3 a) { return v3(a.x, a.z, a.y); }
inline V3 yzx(V3 a) { return v3(a.y, a.z, a.x); }
inline V3 zxy(V3 a) { return v3(a.z, a.x, a.y); }

/* mat3(a..i) is column-major: columns (a,b,c), (d,e,f), (g,h,i) */
inline M3 m3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    M3 m; m.c[0] = v3(a, b, c); m.c[1] = v3(d, e, f); m.c[2] = v3(g, h, i); return m;
}
/* A*B: column j of the product is A.c0*B[j][0] + A.c1*B[j][1] + A.c2*B[j][2] */
inline M3 operator*(const M3& A, const M3& B) {
    M3 r;
    for (int j = 0; j < 3; j++) r.c[j] = A.c[0] * B.c[j].x + A.c[1] * B.c[j].y + A.c[2] * B.c[j].z;
    return r;
}
/* v*M (row vector): component j = dot(v, column j); M*v = c0*v.x + c1*v.y + c2*v.z */
inline V3 operator*(V3 v, const M3& M) { return v3(dot(v, M.c[0]), dot(v, M.c[1]), dot(v, M.c[2])); }
inline V3 operator*(const M3& M, V3 v) { return M.c[0] * v.x + M.c[1] * v.y + M.c[2] * v.z; }

/* ------------------------------------------------------------------------------------------------------------
 * shader state: the uniform block, the push constants, the injected SDF dispatchers
 * ---------------------------------------------------------------------------------------------------------- */
/* which SDFs a ray's bounding-box search found: bit i % 32 of word i / 32 -- the four masks set1..set4 the shader declares
 * (shader.comp:706-719, 732-738).  The reference only ever fills set1 ("Program To Use set2, set3, set4 Not Yet
 * Written", its `1 << uint(i)` is undefined from i = 32 on); filling the other three the way InsertSDF's generated
 * dispatcher lines already read them (host:2012, 2029: set<i/32 + 1> & 2^(i % 32)) is the extension of SURVEY 8f-3.
 * With at most 32 SDFs only w[0] is ever non-zero and everything is the reference's arithmetic. */
struct SdfSet { unsigned w[4]; };
inline SdfSet sdfset(unsigned a = 0u, unsigned b = 0u, unsigned c = 0u, unsigned d = 0u) { SdfSet s = {{a, b, c, d}}; return s; }
inline SdfSet sdfset_bit(int i) { SdfSet s = sdfset(); if (i >= 0 && i < 128) s.w[i >> 5] = 1u << (unsigned)(i & 31); return s; }
inline void sdfset_add(SdfSet& s, int i) { if (i >= 0 && i < 128) s.w[i >> 5] += 1u << (unsigned)(i & 31); }
typedef float (*sdf_dispatch_fn)(const float* sdfs, float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4);

const float MAXDIST = 1e5f;                                   /* shader.comp:8 */
const float PI = 3.141592653589792623810034526344f;           /* shader.comp:9  (rounds to 0x40490FDB) */
const float ONEBYTHREE = 0.3333333f;                          /* shader.comp:10 */

struct Ray { V3 origin, dir; };                                                    /* shader.comp:54-57 */
struct Sphere { V3 pos; float radius; int materialID, lightID; };                 /* shader.comp:59-64 */
struct Plane { V3 pos; int materialID, lightID; };                                /* shader.comp:66-70 */
struct Box { V3 pos, rotation, size; int materialID, lightID; };                  /* shader.comp:72-78 */
struct SphereSlice { V3 pos; float radius, sliceSize; bool is1stSlice; V3 rotation; int materialID, lightID; };
struct Lens { V3 pos, rotation; float radius, focalLength, thickness; bool isConverging; int materialID, lightID; };
struct Cyclide { V3 pos, rotation, scale; float a, b, c, d, brad; int materialID, lightID; };
struct Sdf { V3 pos, size; };
struct Material { V3 reflection; };
struct Light { V2 emission; };

/* flat offsets of the arrays inside the 4097-float block (shader.comp:19-27) */
enum { OFF_NUM = 0, OFF_OBJ = 7, OFF_SDF = 7 + 1024, OFF_MAT = OFF_SDF + 768, OFF_LGT = OFF_MAT + 783,
       OFF_LID = OFF_LGT + 128, OFF_CIE = OFF_LID + 64, UBO_FLOATS = OFF_CIE + 1323 };

struct Shader {
    const float* ubo;  /* 4097 floats */
    pt_params pc;
    sdf_dispatch_fn sdf_fn, sdfmat_fn;
    V3 cameraPos;      /* shader.comp:127 */

    /* An array read at a computed index.  GLSL leaves out-of-range indexing undefined; the block is contiguous
     * memory, so the canonical choice here (and in the CUDA kernels) is flat indexing clamped to the block. */
    float at(int base, int i) const {
        long k = (long)base + (long)i;
        if (k < 0) k = 0;
        if (k > UBO_FLOATS - 1) k = UBO_FLOATS - 1;
        return ubo[k];
    }
    float numObjects(int i) const { return ubo[OFF_NUM + i]; }
    float objects(int i) const { return at(OFF_OBJ, i); }
    float sdfs(int i) const { return at(OFF_SDF, i); }
    float materials(int i) const { return at(OFF_MAT, i); }
    float lights(int i) const { return at(OFF_LGT, i); }
    float lightIDs(int i) const { return at(OFF_LID, i); }
    float CIE(int i) const { return at(OFF_CIE, i); }

    /* shader.comp:129-140 */
    V3 WaveToXYZ(float wave) const {
        V3 XYZ = v3(0.0f);
        if ((wave >= 360.0f) && (wave <= 800.0f)) {
            int index3 = 3 * pt_f2i(pt_floor(wave) - 360.0f);
            V3 t1 = v3(CIE(index3), CIE(index3 + 1), CIE(index3 + 2));
            V3 t2 = v3(CIE(index3 + 3), CIE(index3 + 4), CIE(index3 + 5));
            XYZ = vmix(t1, t2, wave - pt_floor(wave));
        }
        return XYZ;
    }

    /* shader.comp:143-152 */
    static M3 RotationMatrix(V3 angle) {
        angle = angle * 0.0174532925199f;
        V3 s = vsin(angle);
        V3 c = vcos(angle);
        M3 mX = m3(1.0f, 0.0f, 0.0f, 0.0f, c.x, -s.x, 0.0f, s.x, c.x);
        M3 mY = m3(c.y, 0.0f, s.y, 0.0f, 1.0f, 0.0f, -s.y, 0.0f, c.y);
        M3 mZ = m3(c.z, -s.z, 0.0f, s.z, c.z, 0.0f, 0.0f, 0.0f, 1.0f);
        return mX * mY * mZ;
    }

    /* shader.comp:154-235 */
    void UnpackSphere(Sphere& o, int index) const {
        index *= 6;
        o.pos = v3(objects(index), objects(index + 1), objects(index + 2));
        o.radius = objects(index + 3);
        o.materialID = pt_f2i(objects(index + 4)) - 1;
        o.lightID = pt_f2i(objects(index + 5)) - 1;
    }
    void UnpackPlane(Plane& o, int index, int offset) const {
        index *= 5;
        o.pos = v3(objects(index + offset), objects(index + 1 + offset), objects(index + 2 + offset));
        o.materialID = pt_f2i(objects(index + 3 + offset)) - 1;
        o.lightID = pt_f2i(objects(index + 4 + offset)) - 1;
    }
    void UnpackBox(Box& o, int index, int offset) const {
        index *= 11;
        o.pos = v3(objects(index + offset), objects(index + 1 + offset), objects(index + 2 + offset));
        o.rotation = v3(objects(index + 3 + offset), objects(index + 4 + offset), objects(index + 5 + offset));
        o.size = v3(objects(index + 6 + offset), objects(index + 7 + offset), objects(index + 8 + offset));
        o.materialID = pt_f2i(objects(index + 9 + offset)) - 1;
        o.lightID = pt_f2i(objects(index + 10 + offset)) - 1;
    }
    void UnpackLens(Lens& o, int index, int offset) const {
        index *= 12;
        o.pos = v3(objects(index + offset), objects(index + 1 + offset), objects(index + 2 + offset));
        o.rotation = v3(objects(index + 3 + offset), objects(index + 4 + offset), objects(index + 5 + offset));
        o.radius = objects(index + 6 + offset);
        o.focalLength = objects(index + 7 + offset);
        o.thickness = objects(index + 8 + offset);
        o.isConverging = (objects(index + 9 + offset) != 0.0f);
        o.materialID = pt_f2i(objects(index + 10 + offset)) - 1;
        o.lightID = pt_f2i(objects(index + 11 + offset)) - 1;
    }
    void UnpackCyclide(Cyclide& o, int index, int offset) const {
        index *= 16;
        o.pos = v3(objects(index + offset), objects(index + offset + 1), objects(index + offset + 2));
        o.rotation = v3(objects(index + 3 + offset), objects(index + 4 + offset), objects(index + 5 + offset));
        o.scale = v3(objects(index + 6 + offset), objects(index + 7 + offset), objects(index + 8 + offset));
        o.a = objects(index + 9 + offset);
        o.b = objects(index + 10 + offset);
        o.c = objects(index + 11 + offset);
        o.d = objects(index + 12 + offset);
        o.brad = objects(index + 13 + offset);
        o.materialID = pt_f2i(objects(index + 14 + offset)) - 1;
        o.lightID = pt_f2i(objects(index + 15 + offset)) - 1;
    }
    void UnpackSDF(Sdf& o, int index) const {
        index *= 6;
        o.pos = v3(sdfs(index), sdfs(index + 1), sdfs(index + 2));
        o.size = v3(sdfs(index + 3), sdfs(index + 4), sdfs(index + 5));
    }
    void UnpackMaterial(Material& m, int index) const {
        index *= 3;
        m.reflection.x = materials(index);
        m.reflection.y = materials(index + 1);
        m.reflection.z = materials(index + 2);
    }
    void UnpackLight(Light& lt, int index) const {
        if (index == -1) {
            lt.emission.x = 5500.0f;
            lt.emission.y = 0.0f;
            return;
        }
        index *= 2;
        lt.emission.x = lights(index);
        lt.emission.y = lights(index + 1);
    }

    /* shader.comp:237-249 */
    void GetMaterialMix(Material& mat, float materialID) const {
        Material material1, material2;
        UnpackMaterial(material1, pt_f2i(pt_floor(materialID)));
        UnpackMaterial(material2, pt_f2i(pt_ceil(materialID)));
        float x = materialID - pt_floor(materialID);
        mat.reflection.x = gmix(material1.reflection.x, material2.reflection.x, x);
        mat.reflection.y = gmix(material1.reflection.y, material2.reflection.y, x);
        mat.reflection.z = gmix(material1.reflection.z, material2.reflection.z, x);
    }
    /* shader.comp:251-254 */
    void GetLightMix(Light& lt, float lightID) const { UnpackLight(lt, pt_f2i(lightID)); }

    /* shader.comp:256-261 */
    static void SortMinMax(V3& t1, V3& t2) {
        V3 a = t1, b = t2;
        t1 = vmin(a, b);
        t2 = vmax(a, b);
    }

    /* shader.comp:263-276 */
    static bool BoundingSphere(const Ray& ray, V3 pos, float radius2) {
        CNT(C_BSPHERE, 1);
        V3 localorigin = ray.origin - pos;
        float b = dot(ray.dir, localorigin);
        float c = dot(localorigin, localorigin) - radius2;
        if ((b * b) < c) return false;
        if ((b >= 0.0f) && (c >= 0.0f)) return false;
        return true;
    }

    /* shader.comp:278-287 */
    static V2 RayIntersectAABB(V3 origin, V3 invdir, const Box& object) {
        V3 localorigin = origin - object.pos;
        V3 tMin = vfma(object.size, v3(-0.5f), -localorigin) * invdir;
        V3 tMax = vfma(object.size, v3(0.5f), -localorigin) * invdir;
        SortMinMax(tMin, tMax);
        float t1 = gmax(gmax(tMin.x, tMin.y), tMin.z);
        float t2 = gmin(gmin(tMax.x, tMax.y), tMax.z);
        return v2(t1, t2);
    }

    /* shader.comp:289-317 */
    static bool SphereIntersection(const Ray& ray, const Sphere& object, float& hitdist, V3& normal,
                                   float& materialID, float& lightID) {
        CNT(C_SPHERE, 1);
        V3 localorigin = ray.origin - object.pos;
        float b = 2.0f * dot(ray.dir, localorigin);
        float c = dot(localorigin, localorigin) - (object.radius * object.radius);
        float discriminant = b * b - 4.0f * c;
        float t = 1e6f;
        int isOutside = 1;
        if (discriminant < 0.0f) return false;
        float sqrtD = gsqrt(discriminant);
        float t1 = (-b - sqrtD) * 0.5f;
        float t2 = (-b + sqrtD) * 0.5f;
        t = (t1 > 0.0f) ? t1 : t2;
        isOutside = (t1 > 0.0f) ? 1 : -1;
        if (t < 1e-4f) return false;
        if (t < hitdist) {
            CNT(C_SPHERE_HIT, 1);
            hitdist = t;
            normal = normalize(vfma(ray.dir, v3(t), localorigin) * (float)isOutside);
            materialID = (float)object.materialID;
            lightID = (float)object.lightID;
            return true;
        }
        return false;
    }

    /* shader.comp:319-335 */
    static bool PlaneIntersection(const Ray& ray, const Plane& object, float& hitdist, V3& normal,
                                  float& materialID, float& lightID) {
        CNT(C_PLANE, 1);
        V3 localorigin = ray.origin - object.pos;
        float t = -localorigin.y / ray.dir.y;
        if (t < 1e-4f) return false;
        if (t < hitdist) {
            CNT(C_PLANE_HIT, 1);
            hitdist = t;
            normal = faceforward(v3(0.0f, 1.0f, 0.0f), ray.dir, v3(0.0f, 1.0f, 0.0f));
            materialID = (float)object.materialID;
            lightID = (float)object.lightID;
            return true;
        }
        return false;
    }

    /* shader.comp:337-364 */
    static bool BoxIntersection(Ray ray, const Box& object, float& hitdist, V3& normal, float& materialID,
                                float& lightID) {
        CNT(C_BOX, 1);
        M3 matrix = RotationMatrix(object.rotation);
        V3 localorigin = (ray.origin - object.pos) * matrix;
        ray.dir = ray.dir * matrix;
        V3 invdir = 1.0f / ray.dir;
        V3 tMin = vfma(object.size, v3(-0.5f), -localorigin) * invdir;
        V3 tMax = vfma(object.size, v3(0.5f), -localorigin) * invdir;
        SortMinMax(tMin, tMax);
        float t1 = gmax(gmax(tMin.x, tMin.y), tMin.z);
        float t2 = gmin(gmin(tMax.x, tMax.y), tMax.z);
        float t = (t1 < 0.0f) ? t2 : t1;
        if ((t1 > t2) || (t < 1e-4f)) return false;
        if (t < hitdist) {
            CNT(C_BOX_HIT, 1);
            hitdist = t;
            V3 p = vabs((localorigin + ray.dir * t) / object.size);
            normal = matrix * (vstep(gmax(gmax(p.x, p.y), p.z), p) * -vsign(ray.dir));
            materialID = (float)object.materialID;
            lightID = (float)object.lightID;
            return true;
        }
        return false;
    }

    /* shader.comp:366-417 */
    static bool SphereSliceIntersection(const Ray& ray, const SphereSlice& object, float localSlicePos,
                                        bool isSideInvert, float& hitdist, V3& normal, int& isOutside,
                                        float& materialID, float& lightID) {
        M3 matrix = RotationMatrix(object.rotation);
        Ray localRay;
        float sliceOffset = object.radius - object.sliceSize;
        localRay.origin = ray.origin - object.pos;
        localRay.origin = localRay.origin * matrix;
        if (object.is1stSlice) {
            localRay.origin.x += localSlicePos - object.sliceSize - sliceOffset;
        } else {
            localRay.origin.x -= localSlicePos - object.sliceSize - sliceOffset;
        }
        localRay.dir = ray.dir * matrix;
        float b = 2.0f * dot(localRay.dir, localRay.origin);
        float c = dot(localRay.origin, localRay.origin) - (object.radius * object.radius);
        float discriminant = b * b - 4.0f * c;
        float t = 1e6f;
        int isOut = 1;
        if (discriminant < 0.0f) return false;
        float sqrtD = gsqrt(discriminant);
        float t1 = (-b - sqrtD) * 0.5f;
        float t2 = (-b + sqrtD) * 0.5f;
        if (object.is1stSlice) {
            t1 = (gfma(localRay.dir.x, t1, localRay.origin.x) > -sliceOffset) ? 1e6f : t1;
            t2 = (gfma(localRay.dir.x, t2, localRay.origin.x) > -sliceOffset) ? 1e6f : t2;
        } else {
            t1 = (gfma(localRay.dir.x, t1, localRay.origin.x) < sliceOffset) ? 1e6f : t1;
            t2 = (gfma(localRay.dir.x, t2, localRay.origin.x) < sliceOffset) ? 1e6f : t2;
        }
        t = (t1 > 0.0f) ? t1 : t;
        if (t2 < t) {
            t = t2;
            isOut = -1;
        }
        if (t < 1e-4f) return false;
        if (t < hitdist) {
            CNT(C_SLICE_HIT, 1);
            hitdist = t;
            normal = matrix * normalize(vfma(localRay.dir, v3(t), localRay.origin) * (float)isOut);
            isOutside = (!isSideInvert) ? isOut : -isOut;
            materialID = (float)object.materialID;
            lightID = (float)object.lightID;
            return true;
        }
        return false;
    }

    /* shader.comp:419-448 */
    static bool LensIntersection(const Ray& ray, const Lens& object, float& hitdist, V3& normal, int& isOutside,
                                 float& materialID, float& lightID) {
        CNT(C_LENS, 1);
        float lensThicknessHalf = 2.0f * object.focalLength -
                                  gsqrt(4.0f * object.focalLength * object.focalLength - object.radius * object.radius);
        const bool lensSlicePart[2] = {true, false};
        float lensSlicePos = 0.5f * (object.isConverging ? object.thickness : -object.thickness);
        if (object.isConverging) lensSlicePos += lensThicknessHalf;
        bool isIntersect = false;
        for (int i = 0; i < 2; i++) {
            SphereSlice slice;
            slice.pos = object.pos;
            slice.radius = 2.0f * object.focalLength;
            slice.sliceSize = lensThicknessHalf;
            slice.is1stSlice = lensSlicePart[i];
            slice.rotation = object.rotation;
            slice.materialID = object.materialID;
            slice.lightID = object.lightID;
            if (SphereSliceIntersection(ray, slice, lensSlicePos, !object.isConverging, hitdist, normal, isOutside,
                                        materialID, lightID)) {
                isIntersect = true;
            }
        }
        return isIntersect;
    }

    /* shader.comp:450-472 */
    static float EvalQuadratic(float a, float b, float c, float x) { return x * (x * a + b) + c; }
    static V3 EvalQuadratic(float a, float b, float c, V3 x) { return x * (x * a + b) + c; }
    static float EvalCubic(float a, float b, float c, float d, float x) { return x * (x * (x * a + b) + c) + d; }
    static V2 EvalCubic(float a, float b, float c, float d, V2 x) { return x * (x * (x * a + b) + c) + d; }
    static V3 EvalCubic(float a, float b, float c, float d, V3 x) { return x * (x * (x * a + b) + c) + d; }
    static V2 EvalQuartic(float a, float b, float c, float d, float e, V2 x) {
        return x * (x * (x * (x * a + b) + c) + d) + e;
    }

    /* shader.comp:474-503 (the return value is never used by the callers) */
    static void SolveCubic(float b, float c, float d, V3& roots) {
        float bdiv3 = b * ONEBYTHREE;
        float Q = c * ONEBYTHREE - bdiv3 * bdiv3;
        float R = 0.5f * bdiv3 * c - bdiv3 * bdiv3 * bdiv3 - 0.5f * d;
        float D = Q * Q * Q + R * R;
        if (D > 0.0f) {
            float u = R + gsqrt(D);
            float v = R - gsqrt(D);
            float S = gsign(u) * gpow(gabs(u), ONEBYTHREE);
            float T = gsign(v) * gpow(gabs(v), ONEBYTHREE);
            roots.x = S + T - bdiv3;
            for (int i = 0; i < 2; i++) {
                roots.x -= EvalCubic(1.0f, b, c, d, roots.x) / EvalQuadratic(3.0f, 2.0f * b, c, roots.x);
            }
            return;
        }
        CNT(C_CYCLIDE_3ROOT, 1);
        float sqrtnegQ = gsqrt(-Q);
        float thetadiv3 = gacos(R / (sqrtnegQ * sqrtnegQ * sqrtnegQ)) * ONEBYTHREE;
        float TWOPIBYTHREE = gfma(2.0f, PI, ONEBYTHREE); /* sic: 2*pi + 1/3 */
        roots = 2.0f * sqrtnegQ *
                    v3(gcos(thetadiv3), gcos(thetadiv3 + TWOPIBYTHREE), gcos(gfma(2.0f, TWOPIBYTHREE, thetadiv3))) -
                bdiv3;
        for (int i = 0; i < 2; i++) {
            roots = roots - EvalCubic(1.0f, b, c, d, roots) / EvalQuadratic(3.0f, 2.0f * b, c, roots);
        }
    }

    /* shader.comp:505-541; isReal[4] out */
    static void SolveQuartic(float a, float b, float c, float d, float e, float roots[4], bool isReal[4]) {
        isReal[0] = isReal[1] = isReal[2] = isReal[3] = false;
        float inva = 1.0f / a;
        float inva2 = inva * 0.5f;
        float inva2a2 = inva2 * inva2;
        float bb = b * b;
        float p = -1.5f * bb * inva2a2 + c * inva;
        float q = bb * b * inva2a2 * inva2 - b * c * inva * inva2 + d * inva;
        float r = -0.1875f * bb * bb * inva2a2 * inva2a2 + 0.5f * c * bb * inva2a2 * inva2 - b * d * inva2a2 + e * inva;
        V3 s = v3(0.0f);
        SolveCubic(0.5f * -p, -r, 0.5f * p * r - 0.125f * q * q, s);
        float s2subp = 2.0f * s.x - p;
        if (s2subp < 0.0f) return;
        float invs2subp = -2.0f * s.x - p;
        float sqrts2subp = gsqrt(s2subp);
        float q2divsqrt = 2.0f * q / sqrts2subp;
        float invaddq2div = invs2subp + q2divsqrt;
        float invsubq2div = invs2subp - q2divsqrt;
        float bdiv4a = 0.25f * inva * b;
        if (invaddq2div >= 0.0f) {
            V2 rt = 0.5f * (v2(1.0f, -1.0f) * gsqrt(invaddq2div) + (-sqrts2subp)) - bdiv4a;
            rt = rt - EvalQuartic(a, b, c, d, e, rt) / EvalCubic(4.0f * a, 3.0f * b, 2.0f * c, d, rt);
            roots[0] = rt.x; roots[1] = rt.y;
            isReal[0] = isReal[1] = true;
        }
        if (invsubq2div >= 0.0f) {
            V2 rt = 0.5f * (v2(1.0f, -1.0f) * gsqrt(invsubq2div) + sqrts2subp) - bdiv4a;
            rt = rt - EvalQuartic(a, b, c, d, e, rt) / EvalCubic(4.0f * a, 3.0f * b, 2.0f * c, d, rt);
            roots[2] = rt.x; roots[3] = rt.y;
            isReal[2] = isReal[3] = true;
        }
    }

    /* shader.comp:633-679 */
    static bool DupinCyclide(const Ray& ray, const Cyclide& object, float& hitdist, V3& normal, float& materialID,
                             float& lightID) {
        CNT(C_CYCLIDE, 1);
        M3 matrix = RotationMatrix(object.rotation);
        V3 o = xzy(((ray.origin - object.pos) * matrix) / object.scale);
        V3 d = xzy((ray.dir * matrix) / object.scale);
        const float A = object.a, B = object.b, C = object.c, D = object.d;
        float a4 = dot(d * d, d * d) + 2.0f * dot(d * d, yzx(d) * yzx(d));
        float a3 = 4.0f * (dot(o, d * d * d) + dot(o * d, yzx(d) * yzx(d)) + dot(o * d, zxy(d) * zxy(d)));
        float a2 = 6.0f * dot(o * o, d * d) + 8.0f * dot(o * d, yzx(o) * yzx(d)) +
                   2.0f * (dot(o * o, yzx(d) * yzx(d)) + dot(o * o, zxy(d) * zxy(d))) +
                   2.0f * (B * B - D * D) * dot(d, d) - 4.0f * (A * A * d.x * d.x + B * B * d.y * d.y);
        float a1 = 4.0f * (dot(o * o * o, d) + dot(o * o, yzx(o) * yzx(d)) + dot(o * o, zxy(o) * zxy(d)) +
                           2.0f * A * C * D * d.x + (B * B - D * D) * dot(o, d) -
                           2.0f * (A * A * o.x * d.x + B * B * o.y * d.y));
        float a0 = dot(o * o, o * o) + 2.0f * dot(o * o, yzx(o) * yzx(o)) + B * B * B * B + D * D * D * D -
                   2.0f * B * B * D * D - 4.0f * C * C * D * D + 8.0f * A * C * D * o.x +
                   2.0f * (B * B - D * D) * dot(o, o) - 4.0f * (A * A * o.x * o.x + B * B * o.y * o.y);

        float roots[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        bool isReal[4];
        SolveQuartic(a4, a3, a2, a1, a0, roots, isReal);

        float t = 1e6f;
        for (int i = 0; i < 4; i++) {
            if (isReal[i]) {
                if ((roots[i] < t) && (roots[i] > 0.0f)) t = roots[i];
            }
        }
        if (t < hitdist) {
            CNT(C_CYCLIDE_HIT, 1);
            hitdist = t;
            float x = o.x + d.x * t;
            float y = o.y + d.y * t;
            float z = o.z + d.z * t;
            float term1 = x * x + y * y + z * z + B * B - D * D;
            normal.x = 4.0f * (x * term1 - 2.0f * A * (A * x - C * D));
            normal.y = 4.0f * z * term1;
            normal.z = 4.0f * y * (term1 - 2.0f * B * B);
            normal = normalize(normal);
            materialID = (float)object.materialID;
            lightID = (float)object.lightID;
            return true;
        }
        return false;
    }

    /* shader.comp:706-719: the dispatchers InsertSDF (host:2004-2054) generates; built by oracle/sdf_build.py */
    float SDF(V3 p, const SdfSet& set1) const {
        CNT(C_SDF_EVAL, 1);
        if (!sdf_fn) return MAXDIST;
        return sdf_fn(ubo + OFF_SDF, p.x, p.y, p.z, set1.w[0], set1.w[1], set1.w[2], set1.w[3]);
    }
    float SDFMATERIAL(V3 p, const SdfSet& set1) const {
        CNT(C_SDFMAT_EVAL, 1);
        if (!sdfmat_fn) return 0.0f;
        return sdfmat_fn(ubo + OFF_SDF, p.x, p.y, p.z, set1.w[0], set1.w[1], set1.w[2], set1.w[3]);
    }

    /* shader.comp:721-730 */
    V3 CalculateNumericalSDFNormals(V3 p, const SdfSet& set1) const {
        float epsilon = 1e-4f;
        V3 hx = v3(epsilon, 0.0f, 0.0f), hy = v3(0.0f, epsilon, 0.0f), hz = v3(0.0f, 0.0f, epsilon);
        float nx = SDF(p + hx, set1) - SDF(p - hx, set1);
        float ny = SDF(p + hy, set1) - SDF(p - hy, set1);
        float nz = SDF(p + hz, set1) - SDF(p - hz, set1);
        return normalize(v3(nx, ny, nz));
    }

    /* shader.comp:732-777 (set1 here = the four masks; see SdfSet) */
    bool SearchSDF(V3 p, V3 invdir, V2& tMinMax, SdfSet& set1) const {
        CNT(C_SEARCHSDF, 1);
        bool isFoundSDF = false;
        set1 = sdfset();
        for (int i = 0; (float)i < numObjects(5); i++) {
            Box boundingBox;
            Sdf sdf;
            UnpackSDF(sdf, i);
            boundingBox.pos = sdf.pos;
            boundingBox.size = sdf.size;
            V2 boxMinMax = RayIntersectAABB(p, invdir, boundingBox);
            if ((boxMinMax.x > boxMinMax.y) || (boxMinMax.y < 0.0f)) continue;
            if (boxMinMax.x < tMinMax.x) {
                isFoundSDF = true;
                if (boxMinMax.y < tMinMax.x) {
                    tMinMax = boxMinMax;
                    set1 = sdfset_bit(i);
                } else {
                    if (boxMinMax.y < tMinMax.y) {
                        tMinMax.x = boxMinMax.x;
                        sdfset_add(set1, i);
                    } else {
                        tMinMax = boxMinMax;
                        sdfset_add(set1, i);
                    }
                }
            } else {
                if (boxMinMax.x < tMinMax.y) {
                    isFoundSDF = true;
                    if (boxMinMax.y > tMinMax.y) {
                        tMinMax.y = boxMinMax.y;
                        sdfset_add(set1, i);
                    } else {
                        sdfset_add(set1, i);
                    }
                }
            }
        }
        return isFoundSDF;
    }

    /* shader.comp:779-860 */
    bool SphereTracing(const Ray& ray, float& hitdist, V3& normal, float& materialID, float& lightID) const {
        CNT(C_ST_CALLS, 1);
        float t = 1e-3f;
        float insT = 0.0f;
        float omegaMax = 1.70f;
        float omegaSpeed = 0.20f;
        float omega = omegaMax;
        float omegaSpeedFactor = 0.0f;
        float previousRadius = 0.0f;
        V3 p = ray.origin;
        V3 invdir = 1.0f / ray.dir;
        int points = 0;
        V2 tMinMax = v2(MAXDIST, MAXDIST);
        SdfSet set1 = sdfset();
        if (SearchSDF(p, invdir, tMinMax, set1)) {
            t = gmax(tMinMax.x, t);
            p = vfma(ray.dir, v3(t), ray.origin);
        } else {
            return false;
        }
        CNT(C_ST_ENTER, 1);
        if (t > hitdist) CNT(C_ST_ENTER_BEYOND, 1); /* the bounding box starts behind the nearest analytic hit */
        float k = gsign(SDF(ray.origin, set1));

        for (int i = 0; i < 512; i++) {
            CNT(C_ST_ITER, 1);
            if (t > hitdist) CNT(C_ST_ITER_BEYOND, 1); /* marching behind the nearest analytic hit (cannot win if t never returns) */
            float radius = SDF(p, set1);
            if (insT > (gabs(previousRadius) + gabs(radius))) {
                CNT(C_ST_BACKSTEP, 1);
                t -= insT;
                omega = 1.0f;
                insT = previousRadius * omega * k;
                t += insT;
                p = vfma(ray.dir, v3(t), ray.origin);
                continue;
            }
            if (gabs(radius) < 1e-4f) break;
            if (t > tMinMax.y) {
                points += 1;
            } else {
                points = 0;
            }
            if (points >= 2) {
                t = tMinMax.y + 1e-3f;
                tMinMax = v2(MAXDIST, MAXDIST);
                if (SearchSDF(vfma(ray.dir, v3(t), ray.origin), invdir, tMinMax, set1)) {
                    tMinMax = tMinMax + v2(t, t);
                    t = gmax(tMinMax.x, t);
                    p = vfma(ray.dir, v3(t), ray.origin);
                    continue;
                } else {
                    return false;
                }
            }
            insT = radius * omega * k;
            t += insT;
            p = vfma(ray.dir, v3(t), ray.origin);
            omegaSpeedFactor = gmin(radius / previousRadius, 0.99f);
            omega += omegaSpeed * (gmin(1.0f / (1.0f - omegaSpeedFactor), omegaMax) - omega);
            previousRadius = radius;
        }

        if (t < hitdist) {
            CNT(C_ST_HIT, 1);
            hitdist = t - 1e-3f;
            p = vfma(ray.dir, v3(t), ray.origin);
            normal = CalculateNumericalSDFNormals(p, set1);
            materialID = SDFMATERIAL(p, set1);
            lightID = -1.0f;
            return true;
        }
        return false;
    }

    static float LensBoundingRadius2(const Lens& object) { /* shader.comp:899-905 == 1172-1178 */
        float boundingRadius = 0.0f;
        if (object.isConverging) {
            boundingRadius = (object.radius * object.radius) + (0.25f * object.thickness * object.thickness);
        } else {
            boundingRadius = 0.5f * object.thickness + 2.0f * object.focalLength -
                             gsqrt(4.0f * object.focalLength * object.focalLength - object.radius * object.radius);
            boundingRadius = boundingRadius * boundingRadius + object.radius * object.radius;
        }
        return boundingRadius;
    }

#ifdef PT_ORACLE_BVH
    /* Closest hit over the analytic primitives through the product's tree: planes in order, then every sphere, box
     * and lens the traversal reaches, then the cyclides in order (not in the tree: pt_bvh.h).  Winner = smallest t,
     * ties to the lowest object index -- the outcome of the in-order `t < hitdist` scan.  The primitive routines
     * update on t < bound, so a lower-index candidate is offered the next float above hitdist as its bound (it then
     * wins on t <= hitdist). */
    void ClosestAnalyticBvh(const Ray& ray, float& hitdist, V3& normal, float& materialID, float& lightID, int& objectID) const {
        const int nS = pt_f2i(numObjects(0)), nP = pt_f2i(numObjects(1)), nB = pt_f2i(numObjects(2)), nL = pt_f2i(numObjects(3));
        const int offP = 6 * nS, offB = offP + 5 * nP, offL = offB + 11 * nB, offC = offL + 12 * nL;
        for (int i = 0; i < nP; i++) {
            Plane object;
            UnpackPlane(object, i, offP);
            if (PlaneIntersection(ray, object, hitdist, normal, materialID, lightID)) objectID = nS + i;
        }
        pt_bvh_traverse(g_bvh_blob.data(), ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x, ray.dir.y, ray.dir.z, hitdist,
                        [&](int ref) {
            const int type = ref >> 16, i = ref & 0xffff;
            const int base = (type == PT_BVH_SPHERE) ? 0 : (type == PT_BVH_BOX) ? nS + nP : nS + nP + nB;
            const int id = base + i;
            float bound = (id < objectID) ? std::nextafterf(hitdist, 3.0e38f) : hitdist;
            bool hit = false;
            if (type == PT_BVH_SPHERE) {
                Sphere object;
                UnpackSphere(object, i);
                hit = SphereIntersection(ray, object, bound, normal, materialID, lightID);
            } else if (type == PT_BVH_BOX) {
                Box object;
                UnpackBox(object, i, offB);
                if (BoundingSphere(ray, object.pos, 0.25f * dot(object.size, object.size)))
                    hit = BoxIntersection(ray, object, bound, normal, materialID, lightID);
            } else {
                Lens object;
                UnpackLens(object, i, offL);
                int isOutside = 1;
                if (BoundingSphere(ray, object.pos, LensBoundingRadius2(object)))
                    hit = LensIntersection(ray, object, bound, normal, isOutside, materialID, lightID);
            }
            if (hit) { hitdist = bound; objectID = id; }
        });
        for (int i = 0; (float)i < numObjects(4); i++) {
            Cyclide object;
            UnpackCyclide(object, i, offC);
            if (!BoundingSphere(ray, object.pos, object.brad)) continue;
            if (DupinCyclide(ray, object, hitdist, normal, materialID, lightID)) objectID = nS + nP + nB + nL + i;
        }
    }
#endif

    /* shader.comp:862-934 */
    float Intersection(const Ray& ray, V3& normal, float& materialID, float& lightID, bool* sdfHit = nullptr) const {
        CNT(C_RAYS_PATH, 1);
#ifdef PT_ORACLE_BVH
        if (!g_bvh_blob.empty()) {
            float hd = MAXDIST;
            int objectID = -1;
            ClosestAnalyticBvh(ray, hd, normal, materialID, lightID, objectID);
            const bool viaSdf = SphereTracing(ray, hd, normal, materialID, lightID);
            if (sdfHit) *sdfHit = viaSdf;
            return hd;
        }
#endif
        float hitdist = MAXDIST;
        int offset = 0;
        for (int i = 0; (float)i < numObjects(0); i++) {
            Sphere object;
            UnpackSphere(object, i);
            SphereIntersection(ray, object, hitdist, normal, materialID, lightID);
        }
        offset += 6 * pt_f2i(numObjects(0));
        for (int i = 0; (float)i < numObjects(1); i++) {
            Plane object;
            UnpackPlane(object, i, offset);
            PlaneIntersection(ray, object, hitdist, normal, materialID, lightID);
        }
        offset += 5 * pt_f2i(numObjects(1));
        for (int i = 0; (float)i < numObjects(2); i++) {
            Box object;
            UnpackBox(object, i, offset);
            if (!BoundingSphere(ray, object.pos, 0.25f * dot(object.size, object.size))) continue;
            BoxIntersection(ray, object, hitdist, normal, materialID, lightID);
        }
        offset += 11 * pt_f2i(numObjects(2));
        for (int i = 0; (float)i < numObjects(3); i++) {
            Lens object;
            UnpackLens(object, i, offset);
            int isOutside = 1;
            if (!BoundingSphere(ray, object.pos, LensBoundingRadius2(object))) continue;
            LensIntersection(ray, object, hitdist, normal, isOutside, materialID, lightID);
        }
        offset += 12 * pt_f2i(numObjects(3));
        for (int i = 0; (float)i < numObjects(4); i++) {
            Cyclide object;
            UnpackCyclide(object, i, offset);
            if (!BoundingSphere(ray, object.pos, object.brad)) continue;
            DupinCyclide(ray, object, hitdist, normal, materialID, lightID);
        }
        offset += 16 * pt_f2i(numObjects(4));
        const bool viaSdf = SphereTracing(ray, hitdist, normal, materialID, lightID);
        if (sdfHit) *sdfHit = viaSdf; /* surface extensions only: the hit point then lies 1e-3 in front of the surface */
        return hitdist;
    }

    /* shader.comp:937-946 */
    static void PCG32(uint32_t& seed) {
        uint32_t state = seed * 747796405u + 2891336453u;
        uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
        seed = (word >> 22u) ^ word;
    }
    static float RandomFloatPCG32(uint32_t& seed) {
        CNT(C_RNG, 1);
        PCG32(seed);
        return (float)seed / (float)0xFFFFFFFFu; /* float(0xFFFFFFFFu) == 2^32 */
    }
    /* shader.comp:948-958 */
    uint32_t GenerateSeed(uint32_t xyx, uint32_t xyy, int k) const {
        uint32_t seed = (uint32_t)(pc.frame - pc.samplesPerFrame + k);
        PCG32(seed);
        seed += xyx + (uint32_t)pc.resolution[0] * xyy;
        return seed;
    }

    /* shader.comp:960-974 */
    static float SampleHeroWavelength(float l_min, float l_max, uint32_t& seed) {
        return gmix(l_min, l_max, RandomFloatPCG32(seed));
    }
    static float InverseSampleWavelengthPDF(float l_min, float l_max) { return l_max - l_min; }
    static V4 SampleWavelengths(float l_h) {
        V4 x = (l_h - 390.0f) + (0.25f * v4(1.0f, 2.0f, 3.0f, 4.0f)) * 330.0f;
        V4 m = v4(gmod(x.x, 330.0f), gmod(x.y, 330.0f), gmod(x.z, 330.0f), gmod(x.w, 330.0f));
        return 390.0f + m;
    }

    /* shader.comp:976-982 */
    static V2 SampleUniformUnitDisk(uint32_t& seed) {
        float rx = RandomFloatPCG32(seed);
        float ry = RandomFloatPCG32(seed);
        float phi = 2.0f * PI * ry;
        float d = gsqrt(rx);
        return d * v2(gcos(phi), gsin(phi));
    }
    /* shader.comp:984-997 */
    static V3 SampleUniformUnitSphere(uint32_t& seed) {
        float rx = RandomFloatPCG32(seed);
        float ry = RandomFloatPCG32(seed);
        float phi = 2.0f * PI * ry;
        float sinTheta = 2.0f * rx - 1.0f;
        float cosTheta = gsqrt(gfma(-sinTheta, sinTheta, 1.0f));
        float x = gcos(phi) * cosTheta;
        float y = gsin(phi) * cosTheta;
        float z = sinTheta;
        return v3(x, y, z);
    }
    /* shader.comp:999-1010 */
    static V3 SampleCosineDirectionHemisphere(V3 normal, uint32_t& seed) {
        V3 sumvector = normal + SampleUniformUnitSphere(seed);
        return normalize(sumvector);
    }
    static float CosineDirectionPDF(float cosTheta) { return cosTheta / PI; }
    /* shader.comp:1012-1028 */
    static V3 SampleCosineUnitCone(uint32_t& seed, float cosThetaMax) {
        float rx = RandomFloatPCG32(seed);
        float ry = RandomFloatPCG32(seed);
        float cosAlphaMax = 2.0f * cosThetaMax * cosThetaMax - 1.0f;
        float phi = 2.0f * PI * ry;
        float cosTheta = (1.0f - cosAlphaMax) * rx + cosAlphaMax;
        float sinTheta = gsqrt(gfma(-cosTheta, cosTheta, 1.0f));
        float x = gcos(phi) * sinTheta;
        float y = gsin(phi) * sinTheta;
        float z = cosTheta + 1.0f;
        return normalize(v3(x, y, z));
    }
    static float CosineUnitConePDF(float cosTheta, float cosThetaMax) {
        return cosTheta / (PI * (1.0f - cosThetaMax * cosThetaMax));
    }

    /* shader.comp:1030-1038 */
    static V4 SpectralPowerDistribution(V4 l, float l_peak, float d, int invert) {
        V4 x = (l - l_peak) / (2.0f * d * d);
        V4 e = -x * x;
        V4 radiance = v4(gexp(e.x), gexp(e.y), gexp(e.z), gexp(e.w));
        float a = (float)invert;
        radiance = radiance * (1.0f - a) + (1.0f - radiance) * a;
        return radiance;
    }
    /* shader.comp:1040-1055 */
    static V4 BlackBodyRadiation(V4 l, float T) {
        V4 num = 1.1910429724e-16f * v4(gpow(l.x, -5.0f), gpow(l.y, -5.0f), gpow(l.z, -5.0f), gpow(l.w, -5.0f));
        V4 q = 0.014387768775f / (l * T);
        V4 den = v4(gexp(q.x), gexp(q.y), gexp(q.z), gexp(q.w)) - 1.0f;
        return num / den;
    }
    static float BlackBodyRadiationPeak(float T) { return 4.0956746759e-6f * gpow(T, 5.0f); }
    static V4 Emit(V4 l, const Light& lt) {
        float temperature = gmax(lt.emission.x, 0.0f);
        return (BlackBodyRadiation(l * 1e-9f, temperature) / BlackBodyRadiationPeak(temperature)) *
               gmax(lt.emission.y, 0.0f);
    }
    /* shader.comp:1064-1073 */
    static float RefractiveIndexBK7Glass(float l) {
        l *= 1e-3f;
        float l2 = l * l;
        float n2 = 1.0f;
        n2 += (1.03961212f * l2) / (l2 - 6.00069867e-3f);
        n2 += (0.231792344f * l2) / (l2 - 2.00179144e-2f);
        n2 += (1.01046945f * l2) / (l2 - 1.03560653e2f);
        return gsqrt(n2);
    }
    /* shader.comp:1075-1091 */
    static V4 EvaluateBRDF(V4 l, const Material& mat) {
        return SpectralPowerDistribution(l, mat.reflection.x, mat.reflection.y, pt_f2i(mat.reflection.z)) / PI;
    }
    static V3 SampleBRDF(V3 normal, uint32_t& seed) { return SampleCosineDirectionHemisphere(normal, seed); }
    static float BRDFPDF(V3 outDir, V3 normal) { return CosineDirectionPDF(dot(outDir, normal)); }

    /* shader.comp:1093-1119 */
    static void OrthonormalBasis(V3& b1, V3& b2, V3 n) {
        b1 = v3(0.0f, -1.0f, 0.0f);
        b2 = v3(-1.0f, 0.0f, 0.0f);
        if (n.z >= -0.9999999f) {
            float a = 1.0f / (1.0f + n.z);
            float b = -n.x * n.y * a;
            b1 = v3(1.0f - (n.x * n.x * a), b, -n.x);
            b2 = v3(b, 1.0f - (n.y * n.y * a), -n.y);
        }
    }
    static V3 ToWorld(V3 v, V3 n) {
        V3 s = v3(0.0f), t = v3(0.0f);
        OrthonormalBasis(s, t, n);
        return s * v.x + t * v.y + n * v.z;
    }

    /* shader.comp:1121-1223 */
    bool LightSourceVisibilityCheck(const Ray& ray, int lightObjectID) const {
        CNT(C_RAYS_SHADOW, 1);
        float hitdist = MAXDIST;
        V3 normal = v3(0.0f);
        float materialID = 0.0f;
        float lightID = -1.0f;
        int objectID = -1;
#ifdef PT_ORACLE_BVH
        if (!g_bvh_blob.empty()) {
            ClosestAnalyticBvh(ray, hitdist, normal, materialID, lightID, objectID);
            if (SphereTracing(ray, hitdist, normal, materialID, lightID)) objectID = -1;
            return objectID == lightObjectID;
        }
#endif
        int offset = 0;
        int objectOffset = 0;
        for (int i = 0; (float)i < numObjects(0); i++) {
            Sphere object;
            UnpackSphere(object, i);
            if (SphereIntersection(ray, object, hitdist, normal, materialID, lightID)) objectID = i;
        }
        offset += 6 * pt_f2i(numObjects(0));
        objectOffset += pt_f2i(numObjects(0));
        for (int i = 0; (float)i < numObjects(1); i++) {
            Plane object;
            UnpackPlane(object, i, offset);
            if (PlaneIntersection(ray, object, hitdist, normal, materialID, lightID)) objectID = i + objectOffset;
        }
        offset += 5 * pt_f2i(numObjects(1));
        objectOffset += pt_f2i(numObjects(1));
        for (int i = 0; (float)i < numObjects(2); i++) {
            Box object;
            UnpackBox(object, i, offset);
            if (!BoundingSphere(ray, object.pos, 0.25f * dot(object.size, object.size))) continue;
            if (BoxIntersection(ray, object, hitdist, normal, materialID, lightID)) objectID = i + objectOffset;
        }
        offset += 11 * pt_f2i(numObjects(2));
        objectOffset += pt_f2i(numObjects(2));
        for (int i = 0; (float)i < numObjects(3); i++) {
            Lens object;
            UnpackLens(object, i, offset);
            int isOutside = 1;
            if (!BoundingSphere(ray, object.pos, LensBoundingRadius2(object))) continue;
            if (LensIntersection(ray, object, hitdist, normal, isOutside, materialID, lightID))
                objectID = i + objectOffset;
        }
        offset += 12 * pt_f2i(numObjects(3));
        objectOffset += pt_f2i(numObjects(3));
        for (int i = 0; (float)i < numObjects(4); i++) {
            Cyclide object;
            UnpackCyclide(object, i, offset);
            if (!BoundingSphere(ray, object.pos, object.brad)) continue;
            if (DupinCyclide(ray, object, hitdist, normal, materialID, lightID)) objectID = i + objectOffset;
        }
        offset += 16 * pt_f2i(numObjects(4));
        objectOffset += pt_f2i(numObjects(4));
        if (SphereTracing(ray, hitdist, normal, materialID, lightID)) objectID = -1;
        return objectID == lightObjectID;
    }

    /* shader.comp:1225-1285, including the wrong counts of the lens and cyclide branches */
    int SampleRandomLightSource(uint32_t& seed, float& boundingRadius, V3& pos, float& lightID) const {
        int randomLight = pt_f2i(pt_floor(RandomFloatPCG32(seed) * numObjects(6)));
        int randomLightID = pt_f2i(lightIDs(randomLight));
        int offset = 0;
        if (randomLightID < pt_f2i(numObjects(0))) {
            Sphere object;
            UnpackSphere(object, randomLightID);
            boundingRadius = object.radius;
            pos = object.pos;
            lightID = (float)object.lightID;
            return pt_f2i(lightIDs(randomLight));
        }
        randomLightID -= pt_f2i(numObjects(0));
        offset += 6 * pt_f2i(numObjects(0));
        if (randomLightID < pt_f2i(numObjects(1))) {
            Plane object;
            UnpackPlane(object, randomLightID, offset);
            boundingRadius = 1e5f;
            pos = object.pos;
            lightID = (float)object.lightID;
            return pt_f2i(lightIDs(randomLight));
        }
        randomLightID -= pt_f2i(numObjects(1));
        offset += 5 * pt_f2i(numObjects(1));
        if (randomLightID < pt_f2i(numObjects(2))) {
            Box object;
            UnpackBox(object, randomLightID, offset);
            boundingRadius = 0.5f * length(object.size);
            pos = object.pos;
            lightID = (float)object.lightID;
            return pt_f2i(lightIDs(randomLight));
        }
        randomLightID -= pt_f2i(numObjects(2));
        offset += 11 * pt_f2i(numObjects(2));
        if (randomLightID < pt_f2i(numObjects(2))) { /* sic */
            Lens object;
            UnpackLens(object, randomLightID, offset);
            boundingRadius = gsqrt((object.radius * object.radius) + (0.25f * object.thickness * object.thickness));
            pos = object.pos;
            lightID = (float)object.lightID;
            return pt_f2i(lightIDs(randomLight));
        }
        randomLightID -= pt_f2i(numObjects(3));
        offset += 12 * pt_f2i(numObjects(3));
        if (randomLightID < pt_f2i(numObjects(3))) { /* sic */
            Cyclide object;
            UnpackCyclide(object, randomLightID, offset);
            boundingRadius = gsqrt(object.brad);
            pos = object.pos;
            lightID = (float)object.lightID;
            return pt_f2i(lightIDs(randomLight));
        }
        return 0;
    }
    float SampleRandomLightSourcePDF() const { return 1.0f / numObjects(6); }          /* shader.comp:1287-1290 */
    static float MISPowerHeuristicsBeta2(float pdf1, float pdf2) {                       /* shader.comp:1293-1296 */
        return pdf1 * pdf1 / (pdf1 * pdf1 + pdf2 * pdf2);
    }

    /* ---- surface extensions (not in shader.comp; see the note at g_surface_ext) ---------------------------------- */
    struct GlossyLobe { V3 i; float a2; V4 R; };
    static const pt_surface_ext* SurfaceExtOf(float materialID) {
        const int m = pt_f2i(pt_floor(materialID));
        if (m < 0 || m >= (int)g_surface_ext.size() || g_surface_ext[(size_t)m].bsdf == PT_BSDF_REFERENCE) return nullptr;
        return &g_surface_ext[(size_t)m];
    }
    static V3 Reflect(V3 I, V3 N) { float d = 2.0f * dot(N, I); return v3(I.x - d * N.x, I.y - d * N.y, I.z - d * N.z); }
    static float GgxD(float nh, float a2) {
        float q = nh * nh * (a2 - 1.0f) + 1.0f;
        return a2 / (PI * (q * q));
    }
    static float GgxG1(float nv, float a2) { return (2.0f * nv) / (nv + gsqrt(a2 + (1.0f - a2) * (nv * nv))); }
    static V4 SchlickF(V4 f0, float ih) {
        float m = gclamp(1.0f - ih, 0.0f, 1.0f);
        float m2 = m * m;
        float m5 = m2 * m2 * m;
        return f0 + (1.0f - f0) * m5;
    }
    static float GgxAlpha2(float roughness) {
        float a = roughness * roughness;
        return gmax(a * a, 1e-8f);
    }
    static V4 GgxEval(V3 i, V3 o, V3 n, float a2, V4 f0) {
        float ni = dot(n, i), no = dot(n, o);
        if (!(ni > 0.0f) || !(no > 0.0f)) return v4(0.0f);
        V3 h = normalize(i + o);
        float nh = dot(n, h), ih = dot(i, h);
        float k = (GgxD(nh, a2) * (GgxG1(ni, a2) * GgxG1(no, a2))) / (4.0f * ni * no);
        return SchlickF(f0, ih) * k;
    }
    /* one scattering event: next direction, throughput factor f cos / pdf, lobe pdf (0: delta), path dies */
    static void SurfaceExtSample(const pt_surface_ext& ext, V3 d, V3 n, V4 R, V4 l, uint32_t& seed, bool& inside, V3& outDir,
                                 V4& weight, float& pdf, bool& dead) {
        pdf = 0.0f;
        dead = false;
        weight = R;
        if (ext.bsdf == PT_BSDF_MIRROR) {
            outDir = Reflect(d, n);
        } else if (ext.bsdf == PT_BSDF_GLOSSY) {
            float a2 = GgxAlpha2(ext.roughness);
            float u1 = RandomFloatPCG32(seed);
            float u2 = RandomFloatPCG32(seed);
            float cos2 = (1.0f - u1) / (1.0f + (a2 - 1.0f) * u1);
            float cosT = gsqrt(cos2);
            float sinT = gsqrt(gmax(1.0f - cos2, 0.0f));
            float phi = 2.0f * PI * u2;
            V3 h = ToWorld(v3(gcos(phi) * sinT, gsin(phi) * sinT, cosT), n);
            outDir = Reflect(d, h);
            V3 i = -d;
            float ni = dot(n, i), no = dot(n, outDir), nh = dot(n, h), ih = dot(i, h);
            if (!(no > 0.0f) || !(ih > 0.0f) || !(ni > 0.0f)) {
                dead = true;
                weight = v4(0.0f);
            } else {
                pdf = (GgxD(nh, a2) * nh) / (4.0f * ih);
                weight = SchlickF(R, ih) * (((GgxG1(ni, a2) * GgxG1(no, a2)) * ih) / (ni * nh));
            }
        } else { /* PT_BSDF_DIELECTRIC */
            float ng = (ext.ior > 0.0f) ? ext.ior : RefractiveIndexBK7Glass(l.w);
            float n1 = inside ? ng : 1.0f, n2 = inside ? 1.0f : ng;
            float eta = n1 / n2;
            float cosi = -dot(d, n);
            float sin2t = eta * eta * (1.0f - cosi * cosi);
            float F = 1.0f, cost = 0.0f;
            if (sin2t < 1.0f) {
                cost = gsqrt(1.0f - sin2t);
                float rs = (n1 * cosi - n2 * cost) / (n1 * cosi + n2 * cost);
                float rp = (n2 * cosi - n1 * cost) / (n2 * cosi + n1 * cost);
                F = 0.5f * (rs * rs + rp * rp);
            }
            float u = RandomFloatPCG32(seed);
            if (u < F) {
                outDir = Reflect(d, n);
                weight = v4(1.0f);
            } else {
                float k = eta * cosi - cost;
                outDir = normalize(v3(eta * d.x + k * n.x, eta * d.y + k * n.y, eta * d.z + k * n.z));
                inside = !inside;
            }
        }
    }

    /* shader.comp:1298-1343 (glossy: not in the reference -- the light sample is weighted with the extension's lobe) */
    V4 SampleLightSource(V4 l, V4 rayradiance, Ray outRay, V3 normal, const Material& mat, uint32_t& seed,
                         float BRDFpdf, float& MISBRDFWeight, const GlossyLobe* glossy = nullptr) const {
        float boundingRadius = 0.0f;
        V3 lightPos = v3(0.0f);
        float lightIDOut = -1.0f;
        int lightObjectID = 0;
        float lightpdf = 0.0f;
        if (numObjects(6) > 0.0f) {
            CNT(C_LIGHT_SAMPLE, 1);
            lightObjectID = SampleRandomLightSource(seed, boundingRadius, lightPos, lightIDOut);
            float invLightDistance = 1.0f / length(lightPos - outRay.origin);
            V3 lightDir = (lightPos - outRay.origin) * invLightDistance;
            float sinthetaMax = gmin(boundingRadius * invLightDistance, 1.0f);
            float costhetaMax = gsqrt(1.0f - sinthetaMax * sinthetaMax);
            outRay.dir = ToWorld(SampleCosineUnitCone(seed, costhetaMax), lightDir);
            lightpdf = SampleRandomLightSourcePDF();
            lightpdf *= CosineUnitConePDF(dot(outRay.dir, lightDir), costhetaMax);
            MISBRDFWeight = MISPowerHeuristicsBeta2(BRDFpdf, lightpdf);
            float costheta = dot(outRay.dir, normal);
            float deathProbability = 1.25f * gmax(MISBRDFWeight - 0.2f, 0.0f);
            if (costheta >= 0.0f) {
                if (RandomFloatPCG32(seed) > deathProbability) {
                    bool isVisible = LightSourceVisibilityCheck(outRay, lightObjectID);
                    if (isVisible) {
                        CNT(C_LIGHT_VISIBLE, 1);
                        Light lt;
                        GetLightMix(lt, lightIDOut);
                        V4 f = glossy ? GgxEval(glossy->i, outRay.dir, normal, glossy->a2, glossy->R) : EvaluateBRDF(l, mat);
                        rayradiance = rayradiance * (f * costheta / lightpdf);
                        return Emit(l, lt) * rayradiance * (1.0f - MISBRDFWeight);
                    }
                } else {
                    MISBRDFWeight = 1.0f;
                }
            }
            return v4(0.0f);
        }
        MISBRDFWeight = MISPowerHeuristicsBeta2(BRDFpdf, lightpdf);
        return v4(0.0f);
    }

    /* shader.comp:1345-1391 */
    V4 TraceRay(V4 l, V4& rayradiance, Ray& inRay, uint32_t& seed, float& MISBRDFWeight, bool& isTerminate, bool& inside) const {
        V4 radiance = v4(0.0f);
        V3 normal = v3(0.0f);
        float materialID = 0.0f;
        float lightID = -1.0f;
        bool sdfHit = false;
        float hitdist = Intersection(inRay, normal, materialID, lightID, &sdfHit);
        Material mat;
        Light lt;
        GetMaterialMix(mat, materialID);
        GetLightMix(lt, lightID);
        Ray outRay = inRay;
        if (hitdist < MAXDIST) {
            if (lt.emission.y > 0.0f) {
                CNT(C_EMIT_HIT, 1);
                radiance = Emit(l, lt) * rayradiance * MISBRDFWeight;
                isTerminate = true;
                return radiance;
            }
            CNT(C_BOUNCE, 1);
            outRay.origin = vfma(inRay.dir, v3(hitdist), inRay.origin);
            if (const pt_surface_ext* ext = SurfaceExtOf(materialID)) { /* not reference behaviour: surface extensions */
                V3 nf = (dot(inRay.dir, normal) > 0.0f) ? -normal : normal;
                V4 R = EvaluateBRDF(l, mat) * PI;
                V4 weight = v4(0.0f);
                float pdf = 0.0f;
                bool dead = false;
                SurfaceExtSample(*ext, inRay.dir, nf, R, l, seed, inside, outRay.dir, weight, pdf, dead);
                /* The next ray starts off the surface, on the side it leaves to: the reference's primitives reject hits
                 * nearer than 1e-4 and its Lambertian never sends a ray INTO a surface; a refracted ray that starts on a
                 * sphere sees its near root at +-1 ulp and, when it comes out positive, misses the far side as well.
                 * SDF hits already sit 1e-3 in front of their surface (shader.comp:853), so they step further. */
                float side = (dot(outRay.dir, nf) > 0.0f) ? 1.0f : -1.0f;
                float eps = side * (sdfHit ? 2e-3f : 2e-4f);
                outRay.origin = v3(outRay.origin.x + nf.x * eps, outRay.origin.y + nf.y * eps, outRay.origin.z + nf.z * eps);
                if (ext->bsdf == PT_BSDF_GLOSSY && !dead && numObjects(6) > 0.0f) {
                    GlossyLobe lobe = {-inRay.dir, GgxAlpha2(ext->roughness), R};
                    radiance = SampleLightSource(l, rayradiance, outRay, nf, mat, seed, pdf, MISBRDFWeight, &lobe);
                } else {
                    MISBRDFWeight = 1.0f;
                }
                rayradiance = rayradiance * weight;
                float p = gclamp(gmax(rayradiance.x, gmax(rayradiance.y, gmax(rayradiance.z, rayradiance.w))), 0.0f, 0.99f);
                if ((RandomFloatPCG32(seed) > p) || dead) {
                    isTerminate = true;
                    return radiance;
                }
                rayradiance = rayradiance * (1.0f / p);
                inRay = outRay;
                return radiance;
            }
            outRay.dir = SampleBRDF(normal, seed);
            float BRDFpdf = BRDFPDF(outRay.dir, normal);
            radiance = SampleLightSource(l, rayradiance, outRay, normal, mat, seed, BRDFpdf, MISBRDFWeight);
            float costheta = dot(outRay.dir, normal);
            rayradiance = rayradiance * (EvaluateBRDF(l, mat) * costheta / BRDFpdf);
            float rayProbability =
                gclamp(gmax(rayradiance.x, gmax(rayradiance.y, gmax(rayradiance.z, rayradiance.w))), 0.0f, 0.99f);
            if (RandomFloatPCG32(seed) > rayProbability) {
                isTerminate = true;
                return radiance;
            }
            rayradiance = rayradiance * (1.0f / rayProbability);
            inRay = outRay;
        } else {
            CNT(C_MISS, 1);
            isTerminate = true;
        }
        return radiance;
    }

    /* shader.comp:1393-1407 */
    V4 TracePath(V4 l, Ray ray, uint32_t& seed) const {
        V4 radiance = v4(0.0f);
        V4 rayradiance = v4(1.0f);
        float MISBRDFWeight = 1.0f;
        bool isTerminate = false;
        bool inside = false; /* surface extensions only: inside a dielectric */
        for (int i = 0; i < pc.pathLength; i++) {
            radiance = radiance + TraceRay(l, rayradiance, ray, seed, MISBRDFWeight, isTerminate, inside);
            if (isTerminate) break;
        }
        return radiance;
    }

    /* shader.comp:1409-1444 */
    void TracePathLens(float l, Ray& ray, V3 forwardDir) const {
        Lens object;
        object.radius = pc.lensRadius;
        object.focalLength = pc.lensFocalLength;
        object.thickness = pc.lensThickness;
        object.isConverging = true;
        object.pos = cameraPos + forwardDir * pc.lensDistance;
        object.rotation = v3(0.0f, 90.0f - pc.cameraAngle[1], pc.cameraAngle[0]);
        object.materialID = 0;
        object.lightID = 0;
#ifdef PT_COUNT
        Counters* saved_cnt = g_cnt; /* the camera lens is part of Scene()'s fixed cost */
        g_cnt = nullptr;
#endif
        for (int i = 0; i < 2; i++) {
            float hitdist = 1e6f;
            V3 normal = v3(0.0f);
            int isOutside = 1;
            float materialID = 0.0f;
            float lightID = -1.0f;
            LensIntersection(ray, object, hitdist, normal, isOutside, materialID, lightID);
            float n1 = 1.0f, n2 = 1.0f;
            if (isOutside == 1) {
                n1 = 1.0f;
                n2 = RefractiveIndexBK7Glass(l);
            } else {
                n1 = RefractiveIndexBK7Glass(l);
                n2 = 1.0f;
            }
            float n12 = n1 / n2;
            l = l * n12;
            ray.origin = vfma(ray.dir, v3(hitdist), ray.origin);
            ray.dir = refract(ray.dir, normal, n12);
        }
#ifdef PT_COUNT
        g_cnt = saved_cnt;
#endif
    }

    /* shader.comp:1446-1490 */
    V3 Scene(uint32_t xyx, uint32_t xyy, V2 uv, int k) const {
        CNT(C_SAMPLES, 1);
        uint32_t seed = GenerateSeed(xyx, xyy, k);
        float j1 = RandomFloatPCG32(seed);
        float j2 = RandomFloatPCG32(seed);
        uv = uv + v2(2.0f * j1 - 0.5f, 2.0f * j2 - 0.5f) / v2((float)pc.resolution[0], (float)pc.resolution[1]);

        Ray ray;
        M3 matrix = RotationMatrix(v3(pc.cameraAngle[0], pc.cameraAngle[1], 0.0f));
        uv = uv * (-pc.cameraSize * 0.5f);
        ray.origin = cameraPos + (v3(uv.x, uv.y, 0.0f) * matrix);
        V2 disk = 0.5f * pc.apertureSize * SampleUniformUnitDisk(seed);
        V3 pointOnAperture = cameraPos + (v3(disk.x, disk.y, pc.apertureDist) * matrix);
        ray.dir = normalize(pointOnAperture - ray.origin);
        V3 forwardDir = v3(matrix.c[0].z, matrix.c[1].z, matrix.c[2].z);

        V3 color = v3(0.0f);
        float l_h = SampleHeroWavelength(360.0f, 800.0f, seed);
        TracePathLens(l_h, ray, forwardDir);
        V4 l = SampleWavelengths(l_h);
        float invNuml = 0.25f;
        V4 radiance = TracePath(l, ray, seed);
        color = color + (radiance.x * WaveToXYZ(l.x) + radiance.y * WaveToXYZ(l.y) + radiance.z * WaveToXYZ(l.z) +
                         radiance.w * WaveToXYZ(l.w)) *
                            InverseSampleWavelengthPDF(390.0f, 720.0f) * invNuml;
        if (color.x != color.x) return v3(0.0f);
        if (color.y != color.y) return v3(0.0f);
        if (color.z != color.z) return v3(0.0f);
        return color;
    }

    /* shader.comp:1492-1507 */
    void Accumulate(V3 inColor, V3& outColor) const {
        if ((pc.currentSamples == pc.samplesPerFrame) && (pc.frame > pc.samplesPerFrame)) {
            float weight = gpow(2.0f, -8.0f / (pc.FPS * pc.persistence));
            outColor = ((1.0f - weight) * outColor) + (weight * inColor);
        } else {
            int unitSamples = pc.currentSamples / pc.samplesPerFrame;
            outColor = ((float)(unitSamples - 1) * inColor + outColor) / (float)unitSamples;
        }
    }

    /* shader.comp:1509-1523 */
    V3 Rendering(uint32_t gidx, uint32_t gidy, V3 inColor) const {
        uint32_t xyx = gidx;
        uint32_t xyy = (uint32_t)pc.resolution[1] - gidy;
        V2 uv = (2.0f * v2((float)xyx, (float)xyy) - v2((float)pc.resolution[0], (float)pc.resolution[1])) /
                (float)pc.resolution[1];
        V3 outColor = v3(0.0f);
        for (int i = 0; i < pc.samplesPerFrame; i++) outColor = outColor + Scene(xyx, xyy, uv, i);
        outColor = outColor / (float)pc.samplesPerFrame;
        outColor = outColor * (pc.apertureSize * pc.apertureSize * (float)pc.ISO);
        Accumulate(inColor, outColor);
        return outColor;
    }
};

void* g_sdf_handle = nullptr;
sdf_dispatch_fn g_sdf_fn = nullptr, g_sdfmat_fn = nullptr;
int g_threads = 0;

Shader make_shader(const pt_ubo* ubo, const pt_params* pc) {
    Shader s;
    s.ubo = reinterpret_cast<const float*>(ubo);
    s.pc = *pc;
    s.sdf_fn = g_sdf_fn;
    s.sdfmat_fn = g_sdfmat_fn;
    s.cameraPos = v3(pc->cameraPosX, pc->cameraPosY, pc->cameraPosZ);
    return s;
}

}  // namespace

/* ------------------------------------------------------------------------------------------------------------
 * C entry points for the tests (ctypes)
 * ---------------------------------------------------------------------------------------------------------- */
extern "C" {

/* Load the dispatchers built by oracle/sdf_build.py for the current scene (NULL/"" = scene has no SDF). */
int oracle_load_sdf(const char* so_path) {
    if (g_sdf_handle) { dlclose(g_sdf_handle); g_sdf_handle = nullptr; }
    g_sdf_fn = g_sdfmat_fn = nullptr;
    if (!so_path || !so_path[0]) return 0;
    g_sdf_handle = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!g_sdf_handle) { fprintf(stderr, "oracle_load_sdf: %s\n", dlerror()); return -1; }
    g_sdf_fn = (sdf_dispatch_fn)dlsym(g_sdf_handle, "oracle_SDF4");
    g_sdfmat_fn = (sdf_dispatch_fn)dlsym(g_sdf_handle, "oracle_SDFMATERIAL4");
    return (g_sdf_fn && g_sdfmat_fn) ? 0 : -2;
}

#ifdef PT_ORACLE_BVH
/* Build the product's tree for this block (exactly what pt_set_scene does) and route the closest-hit search through
 * it; ubo == NULL goes back to the in-order scan.  Returns the number of floats of the blob, < 0 on error. */
int oracle_bvh_build(const pt_ubo* ubo) {
    g_bvh_blob.clear();
    if (!ubo) return 0;
    static PtDevScene sc;
    std::string err;
    if (pt_prepare_scene(ubo, &sc, &err) != PT_OK) { fprintf(stderr, "oracle_bvh_build: %s\n", err.c_str()); return -1; }
    if (pt_bvh_build(&sc, &g_bvh_blob, &err) != PT_OK) { fprintf(stderr, "oracle_bvh_build: %s\n", err.c_str()); g_bvh_blob.clear(); return -1; }
    return (int)g_bvh_blob.size();
}
#endif

/* surface extensions (pt_set_surface_ext's twin): n = 0 restores the reference's shading.  Process-global. */
int oracle_set_surface_ext(const pt_surface_ext* table, int n) {
    if (n < 0 || n > PT_MAX_SURFACE_EXT || (n > 0 && !table)) return -1;
    g_surface_ext.assign(table, table + n);
    return 0;
}

void oracle_set_threads(int n) { g_threads = n; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* main() of shader.comp:1525-1533 for every texel: one vkCmdDispatch.  image = W*H RGBA32F, read-modify-write.
 * Only in-range texels are processed (the reference's `>` bounds test lets row H / column W run and store out
 * of range: SURVEY App. C-2; those stores are not reproduced).  counters may be NULL (needs -DPT_COUNT). */
int oracle_dispatch_rows(const pt_ubo* ubo, const pt_params* pc, float* image, unsigned long long* counters,
                         int row_start, int row_step);
int oracle_dispatch(const pt_ubo* ubo, const pt_params* pc, float* image, unsigned long long* counters) {
    return oracle_dispatch_rows(ubo, pc, image, counters, 0, 1);
}

/* Same, restricted to rows row_start, row_start + row_step, ... of the full-resolution frame: a bounded but
 * representative sample of the workload for the CPU baseline timings of bench.py. */
int oracle_dispatch_rows(const pt_ubo* ubo, const pt_params* pc, float* image, unsigned long long* counters,
                         int row_start, int row_step) {
    const Shader sh = make_shader(ubo, pc);
    if (row_step < 1 || row_start < 0) return -1;
    const int W = pc->resolution[0], H = pc->resolution[1];
    if (W <= 0 || H <= 0 || pc->samplesPerFrame <= 0) return -1;
#ifdef _OPENMP
    int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    Counters total;
    memset(&total, 0, sizeof total);
    (void)nt;
#pragma omp parallel num_threads(nt)
    {
#ifdef PT_COUNT
        Counters local;
        memset(&local, 0, sizeof local);
        g_cnt = &local;
#endif
#pragma omp for schedule(dynamic, 1)
        for (int gy = row_start; gy < H; gy += row_step) {
            for (int gx = 0; gx < W; gx++) {
                float* texel = image + 4 * ((size_t)gx + (size_t)W * (size_t)gy);
                V3 in = v3(texel[0], texel[1], texel[2]);
                V3 out = sh.Rendering((uint32_t)gx, (uint32_t)gy, in);
                texel[0] = out.x; texel[1] = out.y; texel[2] = out.z; texel[3] = 1.0f;
            }
        }
#ifdef PT_COUNT
#pragma omp critical
        for (int i = 0; i < C_N; i++) total.v[i] += local.v[i];
        g_cnt = nullptr;
#endif
    }
    if (counters) memcpy(counters, total.v, sizeof total.v);
    return 0;
}

/* Raw per-sample XYZ of Scene() for one pixel and sample indices first..first+n-1 (no exposure, no mean):
 * out = n x 3 floats.  Used for per-sample parity and for the sum-mode (sample-split) reference. */
int oracle_samples(const pt_ubo* ubo, const pt_params* pc, int gx, int gy, int first, int n, float* out) {
    pt_params p = *pc;
    p.samplesPerFrame = 1;
    for (int k = 0; k < n; k++) {
        p.frame = first + k + 1; /* uint(frame - spf + 0) == first + k */
        const Shader sh = make_shader(ubo, &p);
        uint32_t xyx = (uint32_t)gx, xyy = (uint32_t)p.resolution[1] - (uint32_t)gy;
        V2 uv = (2.0f * v2((float)xyx, (float)xyy) - v2((float)p.resolution[0], (float)p.resolution[1])) /
                (float)p.resolution[1];
        V3 c = sh.Scene(xyx, xyy, uv, 0);
        out[3 * k] = c.x; out[3 * k + 1] = c.y; out[3 * k + 2] = c.z;
    }
    return 0;
}

#ifdef PT_COUNT
/* Per-sample cost map (analysis of lane load balance, tests/lane_balance.py): for rows y0..y1-1 and sample indices
 * 0..n-1 of every pixel, out[((gy - y0) * W + gx) * n + k] = {SDF evaluations, rays traced (path + shadow), shaded
 * bounces} of that one Scene() call. */
int oracle_cost_map(const pt_ubo* ubo, const pt_params* pc, int n, unsigned* out, int y0, int y1) {
    const int W = pc->resolution[0];
#pragma omp parallel for schedule(dynamic, 1)
    for (int gy = y0; gy < y1; gy++) {
        Counters cn;
        g_cnt = &cn;
        for (int gx = 0; gx < W; gx++) {
            pt_params p = *pc;
            p.samplesPerFrame = 1;
            for (int k = 0; k < n; k++) {
                memset(&cn, 0, sizeof cn);
                p.frame = k + 1; /* uint(frame - spf + 0) == k */
                const Shader sh = make_shader(ubo, &p);
                const uint32_t xyx = (uint32_t)gx, xyy = (uint32_t)p.resolution[1] - (uint32_t)gy;
                const V2 uv = (2.0f * v2((float)xyx, (float)xyy) - v2((float)p.resolution[0], (float)p.resolution[1])) /
                              (float)p.resolution[1];
                (void)sh.Scene(xyx, xyy, uv, 0);
                unsigned* o = out + 3 * (((size_t)(gy - y0) * W + gx) * n + k);
                o[0] = (unsigned)cn.v[C_SDF_EVAL];
                o[1] = (unsigned)(cn.v[C_RAYS_PATH] + cn.v[C_RAYS_SHADOW]);
                o[2] = (unsigned)cn.v[C_BOUNCE];
            }
        }
        g_cnt = nullptr;
    }
    return 0;
}
#endif

/* Sum mode: image.xyz += sum over sample indices [first, first+n) of Scene(), in index order; w untouched. */
int oracle_dispatch_sum(const pt_ubo* ubo, const pt_params* pc, int first, int n, float* image) {
    const int W = pc->resolution[0], H = pc->resolution[1];
#ifdef _OPENMP
    int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int gy = 0; gy < H; gy++) {
        for (int gx = 0; gx < W; gx++) {
            float* texel = image + 4 * ((size_t)gx + (size_t)W * (size_t)gy);
            V3 acc = v3(0.0f);
            pt_params p = *pc;
            p.samplesPerFrame = n;
            p.frame = first + n;
            const Shader sh = make_shader(ubo, &p);
            uint32_t xyx = (uint32_t)gx, xyy = (uint32_t)p.resolution[1] - (uint32_t)gy;
            V2 uv = (2.0f * v2((float)xyx, (float)xyy) - v2((float)p.resolution[0], (float)p.resolution[1])) /
                    (float)p.resolution[1];
            for (int k = 0; k < n; k++) acc = acc + sh.Scene(xyx, xyy, uv, k);
            texel[0] += acc.x; texel[1] += acc.y; texel[2] += acc.z;
        }
    }
    return 0;
}

/* ---- unit-level entry points for the known-answer tests -------------------------------------- */
unsigned oracle_pcg32(unsigned s) { uint32_t v = s; Shader::PCG32(v); return v; }
unsigned oracle_generate_seed(const pt_params* pc, int gx, int gy, int k) {
    pt_ubo dummy; (void)dummy;
    Shader s; s.pc = *pc;
    return s.GenerateSeed((uint32_t)gx, (uint32_t)pc->resolution[1] - (uint32_t)gy, k);
}
float oracle_random_float(unsigned* s) { uint32_t v = *s; float f = Shader::RandomFloatPCG32(v); *s = v; return f; }
void oracle_wave_to_xyz(const pt_ubo* ubo, float wave, float* xyz) {
    pt_params pc; memset(&pc, 0, sizeof pc);
    Shader s = make_shader(ubo, &pc);
    V3 v = s.WaveToXYZ(wave); xyz[0] = v.x; xyz[1] = v.y; xyz[2] = v.z;
}
void oracle_sample_wavelengths(float l_h, float* out4) {
    V4 v = Shader::SampleWavelengths(l_h); out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
float oracle_bk7(float l) { return Shader::RefractiveIndexBK7Glass(l); }
void oracle_emit(const float* l4, float temperature, float luminosity, float* out4) {
    Light lt; lt.emission = v2(temperature, luminosity);
    V4 v = Shader::Emit(v4(l4[0], l4[1], l4[2], l4[3]), lt); out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
void oracle_spd(const float* l4, float peak, float sigma, int invert, float* out4) {
    V4 v = Shader::SpectralPowerDistribution(v4(l4[0], l4[1], l4[2], l4[3]), peak, sigma, invert);
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
/* surface extensions, one scattering event (tests/test_surface_ext.py): d, n unit vectors with dot(d, n) <= 0.
 * Returns 1 when the path dies; seed and inside are updated in place; out = dir xyz, weight xyzw, pdf */
int oracle_surface_ext_sample(const pt_surface_ext* ext, const float* d3, const float* n3, const float* R4, const float* l4,
                              uint32_t* seed, int* inside, float* out8) {
    V3 o = v3(0.0f); V4 w = v4(0.0f); float pdf = 0.0f; bool dead = false, in = *inside != 0;
    Shader::SurfaceExtSample(*ext, v3(d3[0], d3[1], d3[2]), v3(n3[0], n3[1], n3[2]), v4(R4[0], R4[1], R4[2], R4[3]),
                             v4(l4[0], l4[1], l4[2], l4[3]), *seed, in, o, w, pdf, dead);
    *inside = in ? 1 : 0;
    out8[0] = o.x; out8[1] = o.y; out8[2] = o.z; out8[3] = w.x; out8[4] = w.y; out8[5] = w.z; out8[6] = w.w; out8[7] = pdf;
    return dead ? 1 : 0;
}
void oracle_ggx_eval(const float* i3, const float* o3, const float* n3, float roughness, const float* f04, float* out4) {
    V4 f = Shader::GgxEval(v3(i3[0], i3[1], i3[2]), v3(o3[0], o3[1], o3[2]), v3(n3[0], n3[1], n3[2]), Shader::GgxAlpha2(roughness),
                           v4(f04[0], f04[1], f04[2], f04[3]));
    out4[0] = f.x; out4[1] = f.y; out4[2] = f.z; out4[3] = f.w;
}
void oracle_rotation_matrix(const float* deg3, float* m9) { /* column-major like GLSL */
    M3 m = Shader::RotationMatrix(v3(deg3[0], deg3[1], deg3[2]));
    for (int c = 0; c < 3; c++) { m9[3 * c] = m.c[c].x; m9[3 * c + 1] = m.c[c].y; m9[3 * c + 2] = m.c[c].z; }
}
/* closest hit of one ray against the scene: returns hitdist; out = normal xyz, materialID, lightID */
float oracle_intersect(const pt_ubo* ubo, const float* origin, const float* dir, float* out5) {
    pt_params pc; memset(&pc, 0, sizeof pc);
    Shader s = make_shader(ubo, &pc);
    Ray r; r.origin = v3(origin[0], origin[1], origin[2]); r.dir = v3(dir[0], dir[1], dir[2]);
    V3 n = v3(0.0f); float m = 0.0f, l = -1.0f;
    float t = s.Intersection(r, n, m, l);
    out5[0] = n.x; out5[1] = n.y; out5[2] = n.z; out5[3] = m; out5[4] = l;
    return t;
}
int oracle_visible(const pt_ubo* ubo, const float* origin, const float* dir, int lightObjectID) {
    pt_params pc; memset(&pc, 0, sizeof pc);
    Shader s = make_shader(ubo, &pc);
    Ray r; r.origin = v3(origin[0], origin[1], origin[2]); r.dir = v3(dir[0], dir[1], dir[2]);
    return s.LightSourceVisibilityCheck(r, lightObjectID) ? 1 : 0;
}
void oracle_solve_quartic(const float* coef5, float* roots4, int* real4) {
    bool isReal[4]; float r[4] = {0, 0, 0, 0};
    Shader::SolveQuartic(coef5[0], coef5[1], coef5[2], coef5[3], coef5[4], r, isReal);
    for (int i = 0; i < 4; i++) { roots4[i] = r[i]; real4[i] = isReal[i]; }
}
void oracle_sdf_eval4(const pt_ubo* ubo, const float* xyz, size_t n, const unsigned* sets4, float* dist, float* material);
void oracle_sdf_eval(const pt_ubo* ubo, const float* xyz, size_t n, unsigned set1_word, float* dist, float* material) {
    const unsigned sets[4] = {set1_word, 0u, 0u, 0u};
    oracle_sdf_eval4(ubo, xyz, n, sets, dist, material);
}
void oracle_sdf_eval4(const pt_ubo* ubo, const float* xyz, size_t n, const unsigned* sets4, float* dist, float* material) {
    pt_params pc; memset(&pc, 0, sizeof pc);
    Shader s = make_shader(ubo, &pc);
    const SdfSet set1 = sdfset(sets4[0], sets4[1], sets4[2], sets4[3]);
    for (size_t i = 0; i < n; i++) {
        V3 p = v3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        if (dist) dist[i] = s.SDF(p, set1);
        if (material) material[i] = s.SDFMATERIAL(p, set1);
    }
}
/* fn: 0 sin, 1 cos, 2 acos, 3 exp2, 4 log2, 5 exp, 6 log, 7 pow(x,y) -- same numbering as pt_math_eval */
void oracle_math_eval(int fn, const float* x, const float* y, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        switch (fn) {
            case 0: out[i] = pt_sin(x[i]); break;
            case 1: out[i] = pt_cos(x[i]); break;
            case 2: out[i] = pt_acos(x[i]); break;
            case 3: out[i] = pt_exp2(x[i]); break;
            case 4: out[i] = pt_log2(x[i]); break;
            case 5: out[i] = pt_exp(x[i]); break;
            case 6: out[i] = pt_log(x[i]); break;
            case 7: out[i] = pt_pow(x[i], y[i]); break;
            default: out[i] = 0.0f;
        }
    }
}
/* The samplers on their own (statistical known-answer tests): n draws from one PCG stream.
 * kind 0 SampleUniformUnitDisk -> (x, y, 0); 1 SampleUniformUnitSphere; 2 SampleCosineDirectionHemisphere(normal n3);
 * 3 SampleCosineUnitCone(cosThetaMax = param) in the cone's frame; 4 ToWorld(SampleCosineUnitCone(param), n3). */
void oracle_sample(int kind, unsigned seed, size_t n, float param, const float* n3, float* out) {
    uint32_t s = seed;
    const V3 nn = n3 ? v3(n3[0], n3[1], n3[2]) : v3(0.0f, 0.0f, 1.0f);
    for (size_t i = 0; i < n; i++) {
        V3 r = v3(0.0f);
        if (kind == 0) { V2 d = Shader::SampleUniformUnitDisk(s); r = v3(d.x, d.y, 0.0f); }
        else if (kind == 1) r = Shader::SampleUniformUnitSphere(s);
        else if (kind == 2) r = Shader::SampleCosineDirectionHemisphere(nn, s);
        else if (kind == 3) r = Shader::SampleCosineUnitCone(s, param);
        else r = Shader::ToWorld(Shader::SampleCosineUnitCone(s, param), nn);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
/* TracePath (shader.comp:1393-1407) on its own: the mean radiance of n paths started with the same ray and wavelength
 * bundle, one PCG stream.  For the closed-form direct-lighting check of tests/test_oracle_kat.py. */
void oracle_trace_path(const pt_ubo* ubo, const pt_params* pc, const float* o3, const float* d3, const float* l4,
                       unsigned seed, size_t n, double* mean4) {
    const Shader sh = make_shader(ubo, pc);
    uint32_t s = seed;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t i = 0; i < n; i++) {
        Ray ray;
        ray.origin = v3(o3[0], o3[1], o3[2]);
        ray.dir = v3(d3[0], d3[1], d3[2]);
        const V4 r = sh.TracePath(v4(l4[0], l4[1], l4[2], l4[3]), ray, s);
        acc[0] += r.x; acc[1] += r.y; acc[2] += r.z; acc[3] += r.w;
    }
    for (int k = 0; k < 4; k++) mean4[k] = acc[k] / (double)(n ? n : 1);
}
/* TracePathLens (shader.comp:1409-1444) for a given ray and hero wavelength; forwardDir as Scene() derives it
 * (1456-1466).  out9 = origin(3), direction(3) after the two refractions, then forwardDir(3). */
void oracle_lens_ray(const pt_ubo* ubo, const pt_params* pc, const float* o3, const float* d3, float l_h, float* out6) {
    const Shader sh = make_shader(ubo, pc);
    const M3 matrix = Shader::RotationMatrix(v3(pc->cameraAngle[0], pc->cameraAngle[1], 0.0f));
    const V3 forwardDir = v3(matrix.c[0].z, matrix.c[1].z, matrix.c[2].z);
    Ray ray;
    ray.origin = v3(o3[0], o3[1], o3[2]);
    ray.dir = v3(d3[0], d3[1], d3[2]);
    sh.TracePathLens(l_h, ray, forwardDir);
    out6[0] = ray.origin.x; out6[1] = ray.origin.y; out6[2] = ray.origin.z;
    out6[3] = ray.dir.x; out6[4] = ray.dir.y; out6[5] = ray.dir.z;
    out6[6] = forwardDir.x; out6[7] = forwardDir.y; out6[8] = forwardDir.z;
}
float oracle_cone_pdf(float cosTheta, float cosThetaMax) { return Shader::CosineUnitConePDF(cosTheta, cosThetaMax); }
float oracle_mis_weight(float pdf1, float pdf2) { return Shader::MISPowerHeuristicsBeta2(pdf1, pdf2); }
void oracle_orthonormal_basis(const float* n3, float* b6) {
    V3 b1 = v3(0.0f), b2 = v3(0.0f);
    Shader::OrthonormalBasis(b1, b2, v3(n3[0], n3[1], n3[2]));
    b6[0] = b1.x; b6[1] = b1.y; b6[2] = b1.z; b6[3] = b2.x; b6[4] = b2.y; b6[5] = b2.z;
}

int oracle_num_counters(void) { return C_N; }
const char* oracle_counter_names(void) {
    return "samples,rays_path,rays_shadow,sphere,sphere_hit,plane,plane_hit,bsphere,box,box_hit,lens,slice_hit,"
           "cyclide,cyclide_3root,cyclide_hit,searchsdf,st_calls,st_enter,st_iter,st_backstep,st_hit,sdf_eval,"
           "sdfmat_eval,bounce,emit_hit,light_sample,light_visible,rng,miss,st_iter_beyond,st_enter_beyond";
}
int oracle_has_counters(void) {
#ifdef PT_COUNT
    return 1;
#else
    return 0;
#endif
}

} /* extern "C" */
