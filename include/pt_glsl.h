/* pt_glsl.h -- the small GLSL-compatible vector header the scene's SDF snippets are compiled against.
 *
 * The reference splices scene["sdf"][i]["glsl"] into shader.comp and lets glslang compile it (host:2004-2054).
 * Here the same text (after the token-level rewrite described in pathtracer_b200/csrc/pt_sdf_front.cpp: float
 * literal suffixes, parameter qualifiers, swizzles -> swizzle calls) is compiled by NVRTC for the GPU and by g++
 * for the CPU oracle, inside `namespace ptglsl`, against the types and builtins below.
 *
 * Semantics are the canonical ones of SURVEY.md App. F (GLSL 4.50 section 8 definitions, componentwise, source
 * order): min(x,y) = y<x ? y : x, max(x,y) = x<y ? y : x, mix = x*(1-a)+y*a, mod = x - y*floor(x/y),
 * dot accumulates left to right, length = sqrt(dot), normalize = v/length(v).
 * Transcendentals go through PT_FN_* so that the strict build uses pt_math.h (bit-exact CPU<->GPU) and the fast
 * build the MUFU intrinsics.  Builtins provided: sin cos tan asin acos atan (one and two arguments) sinh cosh tanh exp
 * exp2 log log2 pow sqrt inversesqrt radians degrees abs sign floor ceil round trunc fract mod min max clamp mix step
 * smoothstep fma length distance dot cross normalize reflect refract faceforward transpose determinant inverse
 * matrixCompMult outerProduct.  Types: float int uint bool, vec2/3/4 with every swizzle (reads through sw<..>(), writes
 * through an lvalue proxy: `p.xz = ..`, `p.xz *= mat2(..)`), mat2/3/4 (column-major, `m[c][r]`, the GLSL constructors
 * and products).  The front end (pt_sdf_front.cpp) rewrites swizzles into the sw / lsw calls.
 */
#ifndef PT_GLSL_H
#define PT_GLSL_H

#include "pt_math.h"

#if defined(PT_FAST) && defined(__CUDA_ARCH__)
#define PT_FN_SIN(x) __sinf(x)
#define PT_FN_COS(x) __cosf(x)
#define PT_FN_ACOS(x) acosf(x)
#define PT_FN_ASIN(x) asinf(x)
#define PT_FN_TAN(x) __tanf(x)
#define PT_FN_ATAN(x) atanf(x)
#define PT_FN_ATAN2(y, x) atan2f(y, x)
#define PT_FN_EXP(x) __expf(x)
#define PT_FN_EXP2(x) exp2f(x)
#define PT_FN_LOG(x) __logf(x)
#define PT_FN_LOG2(x) __log2f(x)
#define PT_FN_POW(x, y) __powf(x, y)
#define PT_FN_SQRT(x) sqrtf(x)
#define PT_FN_RSQRT(x) rsqrtf(x)
#define PT_FN_FMA(a, b, c) fmaf(a, b, c)
/* one FMNMX instead of FSETP + FSEL: min / max are 8 % of the menger scene's executed instructions, all on the half-rate
 * ALU pipe.  Differs from GLSL's (y < x) ? y : x only when an operand is NaN (IEEE minNum / maxNum return the other one) */
#define PT_FN_MIN(x, y) fminf(x, y)
#define PT_FN_MAX(x, y) fmaxf(x, y)
/* x - y * floor(x / y) as one FFMA after the floor (with y = 2 the compiler otherwise emits t + t and two FADDs) */
#define PT_FN_MOD(x, y) fmaf(-(y), floorf((x) / (y)), (x))
#else
#define PT_FN_MOD(x, y) ((x) - (y) * pt_floor((x) / (y)))
#define PT_FN_MIN(x, y) (((y) < (x)) ? (y) : (x))
#define PT_FN_MAX(x, y) (((x) < (y)) ? (y) : (x))
#define PT_FN_SIN(x) pt_sin(x)
#define PT_FN_COS(x) pt_cos(x)
#define PT_FN_ACOS(x) pt_acos(x)
#define PT_FN_ASIN(x) pt_asin(x)
#define PT_FN_TAN(x) pt_tan(x)
#define PT_FN_ATAN(x) pt_atan(x)
#define PT_FN_ATAN2(y, x) pt_atan2(y, x)
#define PT_FN_EXP(x) pt_exp(x)
#define PT_FN_EXP2(x) pt_exp2(x)
#define PT_FN_LOG(x) pt_log(x)
#define PT_FN_LOG2(x) pt_log2(x)
#define PT_FN_POW(x, y) pt_pow(x, y)
#define PT_FN_SQRT(x) pt_sqrt(x)
#define PT_FN_RSQRT(x) pt_rsqrt(x)
#define PT_FN_FMA(a, b, c) pt_fma(a, b, c)
#endif

namespace ptglsl {

typedef unsigned int uint;

struct vec2;
struct vec3;
struct vec4;

struct mat2;
struct mat3;
struct mat4;

/* lvalue swizzles: `v.xz = e`, `v.xz += e`, `v.zyx *= m` ... bind references to the named components */
struct ref2;
struct ref3;
struct ref4;

#define PT_SW2(T2, a, b) PT_HD T2 a##b() const;
#define PT_SW3(T3, a, b, c) PT_HD T3 a##b##c() const;
/* generic swizzle reads / writes, what the front end emits: .sw2<0,2>() etc. */
#define PT_SWIZZLE_MEMBERS                                                                                   \
    template <int A, int B> PT_HD vec2 sw2() const;                                                          \
    template <int A, int B, int C> PT_HD vec3 sw3() const;                                                   \
    template <int A, int B, int C, int D> PT_HD vec4 sw4() const;                                            \
    template <int A, int B> PT_HD ref2 lsw2();                                                               \
    template <int A, int B, int C> PT_HD ref3 lsw3();                                                        \
    template <int A, int B, int C, int D> PT_HD ref4 lsw4();

struct vec2 {
    float x, y;
    PT_HD vec2() : x(0.0f), y(0.0f) {}
    PT_HD explicit vec2(float s) : x(s), y(s) {}
    PT_HD vec2(float x_, float y_) : x(x_), y(y_) {}
    PT_HD explicit vec2(const vec3& v);
    PT_HD float& operator[](int i) { return i == 0 ? x : y; }
    PT_HD float operator[](int i) const { return i == 0 ? x : y; }
    PT_SW2(vec2, x, x) PT_SW2(vec2, x, y) PT_SW2(vec2, y, x) PT_SW2(vec2, y, y)
    PT_SWIZZLE_MEMBERS
};

struct vec3 {
    float x, y, z;
    PT_HD vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    PT_HD explicit vec3(float s) : x(s), y(s), z(s) {}
    PT_HD vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    PT_HD vec3(const vec2& v, float z_) : x(v.x), y(v.y), z(z_) {}
    PT_HD vec3(float x_, const vec2& v) : x(x_), y(v.x), z(v.y) {}
    PT_HD explicit vec3(const vec4& v);
    PT_HD float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    PT_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
#define PT_SW3_ROW(a, b) PT_SW3(vec3, a, b, x) PT_SW3(vec3, a, b, y) PT_SW3(vec3, a, b, z)
#define PT_SW3_BLK(a) PT_SW2(vec2, a, x) PT_SW2(vec2, a, y) PT_SW2(vec2, a, z) PT_SW3_ROW(a, x) PT_SW3_ROW(a, y) PT_SW3_ROW(a, z)
    PT_SW3_BLK(x) PT_SW3_BLK(y) PT_SW3_BLK(z)
    PT_SWIZZLE_MEMBERS
};

struct vec4 {
    float x, y, z, w;
    PT_HD vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    PT_HD explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    PT_HD vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    PT_HD vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    PT_HD vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    PT_HD vec4(const vec2& a, float z_, float w_) : x(a.x), y(a.y), z(z_), w(w_) {}
    PT_HD vec4(float x_, const vec3& v) : x(x_), y(v.x), z(v.y), w(v.z) {}
    PT_HD float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    PT_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    PT_HD vec3 xyz() const { return vec3(x, y, z); }
    PT_HD vec2 xy() const { return vec2(x, y); }
    PT_HD vec2 zw() const { return vec2(z, w); }
    PT_HD vec4 xyzw() const { return *this; }
    PT_SWIZZLE_MEMBERS
};

PT_HD vec2::vec2(const vec3& v) : x(v.x), y(v.y) {}
PT_HD vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

/* the proxies: `v.xz = e` stores e's components; `v.xz op= e` is `v.xz = v.xz op e` with e a scalar, a vector or (for *=) a matrix */
#define PT_REF_COMPOUND(R, V)                                                                          \
    template <class T> PT_HD const R& operator+=(const T& t) const { return *this = V(*this) + t; }    \
    template <class T> PT_HD const R& operator-=(const T& t) const { return *this = V(*this) - t; }    \
    template <class T> PT_HD const R& operator*=(const T& t) const { return *this = V(*this) * t; }    \
    template <class T> PT_HD const R& operator/=(const T& t) const { return *this = V(*this) / t; }
struct ref2 {
    float &a, &b;
    PT_HD ref2(float& a_, float& b_) : a(a_), b(b_) {}
    PT_HD operator vec2() const { return vec2(a, b); }
    PT_HD const ref2& operator=(const vec2& v) const { a = v.x; b = v.y; return *this; }
    PT_REF_COMPOUND(ref2, vec2)
};
struct ref3 {
    float &a, &b, &c;
    PT_HD ref3(float& a_, float& b_, float& c_) : a(a_), b(b_), c(c_) {}
    PT_HD operator vec3() const { return vec3(a, b, c); }
    PT_HD const ref3& operator=(const vec3& v) const { a = v.x; b = v.y; c = v.z; return *this; }
    PT_REF_COMPOUND(ref3, vec3)
};
struct ref4 {
    float &a, &b, &c, &d;
    PT_HD ref4(float& a_, float& b_, float& c_, float& d_) : a(a_), b(b_), c(c_), d(d_) {}
    PT_HD operator vec4() const { return vec4(a, b, c, d); }
    PT_HD const ref4& operator=(const vec4& v) const { a = v.x; b = v.y; c = v.z; d = v.w; return *this; }
    PT_REF_COMPOUND(ref4, vec4)
};
#undef PT_REF_COMPOUND
#define PT_SWIZZLE_DEFS(S)                                                                                                     \
    template <int A, int B> PT_HD vec2 S::sw2() const { return vec2((*this)[A], (*this)[B]); }                                  \
    template <int A, int B, int C> PT_HD vec3 S::sw3() const { return vec3((*this)[A], (*this)[B], (*this)[C]); }               \
    template <int A, int B, int C, int D> PT_HD vec4 S::sw4() const { return vec4((*this)[A], (*this)[B], (*this)[C], (*this)[D]); } \
    template <int A, int B> PT_HD ref2 S::lsw2() { return ref2((*this)[A], (*this)[B]); }                                       \
    template <int A, int B, int C> PT_HD ref3 S::lsw3() { return ref3((*this)[A], (*this)[B], (*this)[C]); }                    \
    template <int A, int B, int C, int D> PT_HD ref4 S::lsw4() { return ref4((*this)[A], (*this)[B], (*this)[C], (*this)[D]); }
PT_SWIZZLE_DEFS(vec2) PT_SWIZZLE_DEFS(vec3) PT_SWIZZLE_DEFS(vec4)
#undef PT_SWIZZLE_DEFS
#undef PT_SWIZZLE_MEMBERS

#undef PT_SW2
#undef PT_SW3
#define PT_SW2(S, a, b) PT_HD vec2 S::a##b() const { return vec2(a, b); }
#define PT_SW3(S, a, b, c) PT_HD vec3 S::a##b##c() const { return vec3(a, b, c); }
PT_SW2(vec2, x, x) PT_SW2(vec2, x, y) PT_SW2(vec2, y, x) PT_SW2(vec2, y, y)
#undef PT_SW3_ROW
#undef PT_SW3_BLK
#define PT_SW3_ROW(a, b) PT_SW3(vec3, a, b, x) PT_SW3(vec3, a, b, y) PT_SW3(vec3, a, b, z)
#define PT_SW3_BLK(a) PT_SW2(vec3, a, x) PT_SW2(vec3, a, y) PT_SW2(vec3, a, z) PT_SW3_ROW(a, x) PT_SW3_ROW(a, y) PT_SW3_ROW(a, z)
PT_SW3_BLK(x) PT_SW3_BLK(y) PT_SW3_BLK(z)
#undef PT_SW2
#undef PT_SW3
#undef PT_SW3_ROW
#undef PT_SW3_BLK

/* ---- scalar builtins ---------------------------------------------------------------------------------------- */
PT_HD float sin(float x) { return PT_FN_SIN(x); }
PT_HD float cos(float x) { return PT_FN_COS(x); }
PT_HD float acos(float x) { return PT_FN_ACOS(x); }
PT_HD float asin(float x) { return PT_FN_ASIN(x); }
PT_HD float tan(float x) { return PT_FN_TAN(x); }
PT_HD float atan(float x) { return PT_FN_ATAN(x); }
PT_HD float atan(float y, float x) { return PT_FN_ATAN2(y, x); }
PT_HD float exp(float x) { return PT_FN_EXP(x); }
PT_HD float exp2(float x) { return PT_FN_EXP2(x); }
PT_HD float log(float x) { return PT_FN_LOG(x); }
PT_HD float log2(float x) { return PT_FN_LOG2(x); }
PT_HD float pow(float x, float y) { return PT_FN_POW(x, y); }
PT_HD float sqrt(float x) { return PT_FN_SQRT(x); }
PT_HD float inversesqrt(float x) { return PT_FN_RSQRT(x); }
PT_HD float fma(float a, float b, float c) { return PT_FN_FMA(a, b, c); }
PT_HD float abs(float x) { return pt_abs(x); }
PT_HD int abs(int x) { return x < 0 ? -x : x; }
PT_HD float floor(float x) { return pt_floor(x); }
PT_HD float ceil(float x) { return pt_ceil(x); }
PT_HD float fract(float x) { return x - pt_floor(x); }
PT_HD float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
PT_HD float min(float x, float y) { return PT_FN_MIN(x, y); } /* GLSL 4.50 8.3: (y < x) ? y : x */
PT_HD float max(float x, float y) { return PT_FN_MAX(x, y); } /*               (x < y) ? y : x */
PT_HD int min(int x, int y) { return (y < x) ? y : x; }
PT_HD int max(int x, int y) { return (x < y) ? y : x; }
PT_HD float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
PT_HD float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
PT_HD float step(float e, float x) { return (x < e) ? 0.0f : 1.0f; }
PT_HD float mod(float x, float y) { return PT_FN_MOD(x, y); } /* GLSL 4.50 8.3: x - y * floor(x / y) */
PT_HD float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
PT_HD float radians(float d) { return d * 0.0174532925199432958f; }
PT_HD float degrees(float r) { return r * 57.295779513082320877f; }
PT_HD float trunc(float x) { return x < 0.0f ? pt_ceil(x) : pt_floor(x); }
PT_HD float round(float x) { return x < 0.0f ? -pt_floor(0.5f - x) : pt_floor(x + 0.5f); } /* halves away from zero */
PT_HD float sinh(float x) { return 0.5f * (exp(x) - exp(-x)); }
PT_HD float cosh(float x) { return 0.5f * (exp(x) + exp(-x)); }
PT_HD float tanh(float x) { float a = exp(x), b = exp(-x); return (a - b) / (a + b); }
PT_HD int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
PT_HD float mix(float x, float y, bool a) { return a ? y : x; }

/* sin on the FMA pipe (fast device builds): MUFU.SIN issues at 1/8 of the FP32 rate, so a snippet made of sines -- the
 * terrain's 22 per evaluation -- is bound by the XU pipe while the FMA pipe idles.  PT_SIN_SITE(n, x) is what the front
 * end emits for the n-th `sin(` of a snippet; every PT_SIN_POLY_EVERY-th site (option "sin_poly_every", 0 = none) takes
 * this polynomial instead: x in turns, nearest integer subtracted with the 1.5 * 2^23 trick, then t P(t^2) on
 * [-1/2, 1/2] (six terms, max abs error 7e-7 -- __sinf's own is 5e-7 near 0 and grows with |x|).  Strict builds and the
 * CPU never see it: there every site is pt_sin. */
#if defined(PT_FAST) && defined(__CUDA_ARCH__) && defined(PT_SIN_POLY_EVERY) && PT_SIN_POLY_EVERY > 0
PT_HD float sin_fma(float x) {
    float t = x * 0.15915494309189535f;
    const float r = (t + 12582912.0f) - 12582912.0f;
    t = t - r;
    const float z = t * t;
    float p = fmaf(z, -12.27126693725586f, 41.20539855957031f);
    p = fmaf(z, p, -76.5801010131836f);
    p = fmaf(z, p, 81.59618377685547f);
    p = fmaf(z, p, -41.34142303466797f);
    p = fmaf(z, p, 6.283182621002197f);
    return p * t;
}
#define PT_SIN_SITE(n, x) ((((n) % PT_SIN_POLY_EVERY) == 0) ? sin_fma(x) : sin(x))
#else
#define PT_SIN_SITE(n, x) sin(x)
#endif

/* ---- componentwise lifting ---------------------------------------------------------------------------------- */
#define PT_V_UN(f)                                                          \
    PT_HD vec2 f(const vec2& a) { return vec2(f(a.x), f(a.y)); }            \
    PT_HD vec3 f(const vec3& a) { return vec3(f(a.x), f(a.y), f(a.z)); }    \
    PT_HD vec4 f(const vec4& a) { return vec4(f(a.x), f(a.y), f(a.z), f(a.w)); }
PT_V_UN(sin) PT_V_UN(cos) PT_V_UN(acos) PT_V_UN(exp) PT_V_UN(exp2) PT_V_UN(log) PT_V_UN(log2) PT_V_UN(sqrt)
PT_V_UN(inversesqrt) PT_V_UN(abs) PT_V_UN(floor) PT_V_UN(ceil) PT_V_UN(fract) PT_V_UN(sign)
#if defined(PT_FAST) && defined(__CUDA_ARCH__) && defined(PT_SIN_POLY_EVERY) && PT_SIN_POLY_EVERY > 0
PT_V_UN(sin_fma)
#endif
PT_V_UN(tan) PT_V_UN(asin) PT_V_UN(atan) PT_V_UN(radians) PT_V_UN(degrees) PT_V_UN(trunc) PT_V_UN(round) PT_V_UN(sinh) PT_V_UN(cosh) PT_V_UN(tanh)
#undef PT_V_UN

/* f(vec, vec) and f(vec, float) */
#define PT_V_BIN(f)                                                                               \
    PT_HD vec2 f(const vec2& a, const vec2& b) { return vec2(f(a.x, b.x), f(a.y, b.y)); }         \
    PT_HD vec3 f(const vec3& a, const vec3& b) { return vec3(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z)); } \
    PT_HD vec4 f(const vec4& a, const vec4& b) { return vec4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w)); } \
    PT_HD vec2 f(const vec2& a, float b) { return vec2(f(a.x, b), f(a.y, b)); }                   \
    PT_HD vec3 f(const vec3& a, float b) { return vec3(f(a.x, b), f(a.y, b), f(a.z, b)); }        \
    PT_HD vec4 f(const vec4& a, float b) { return vec4(f(a.x, b), f(a.y, b), f(a.z, b), f(a.w, b)); }
PT_V_BIN(min) PT_V_BIN(max) PT_V_BIN(mod) PT_V_BIN(pow) PT_V_BIN(atan)
#undef PT_V_BIN
PT_HD vec2 step(const vec2& e, const vec2& x) { return vec2(step(e.x, x.x), step(e.y, x.y)); }
PT_HD vec3 step(const vec3& e, const vec3& x) { return vec3(step(e.x, x.x), step(e.y, x.y), step(e.z, x.z)); }
PT_HD vec2 step(float e, const vec2& x) { return vec2(step(e, x.x), step(e, x.y)); }
PT_HD vec3 step(float e, const vec3& x) { return vec3(step(e, x.x), step(e, x.y), step(e, x.z)); }

/* ---- operators ---------------------------------------------------------------------------------------------- */
#define PT_V_OP(op)                                                                                        \
    PT_HD vec2 operator op(const vec2& a, const vec2& b) { return vec2(a.x op b.x, a.y op b.y); }          \
    PT_HD vec2 operator op(const vec2& a, float b) { return vec2(a.x op b, a.y op b); }                    \
    PT_HD vec2 operator op(float a, const vec2& b) { return vec2(a op b.x, a op b.y); }                    \
    PT_HD vec3 operator op(const vec3& a, const vec3& b) { return vec3(a.x op b.x, a.y op b.y, a.z op b.z); } \
    PT_HD vec3 operator op(const vec3& a, float b) { return vec3(a.x op b, a.y op b, a.z op b); }          \
    PT_HD vec3 operator op(float a, const vec3& b) { return vec3(a op b.x, a op b.y, a op b.z); }          \
    PT_HD vec4 operator op(const vec4& a, const vec4& b) { return vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    PT_HD vec4 operator op(const vec4& a, float b) { return vec4(a.x op b, a.y op b, a.z op b, a.w op b); } \
    PT_HD vec4 operator op(float a, const vec4& b) { return vec4(a op b.x, a op b.y, a op b.z, a op b.w); } \
    PT_HD vec2& operator op##=(vec2& a, const vec2& b) { a = a op b; return a; }                           \
    PT_HD vec2& operator op##=(vec2& a, float b) { a = a op b; return a; }                                 \
    PT_HD vec3& operator op##=(vec3& a, const vec3& b) { a = a op b; return a; }                           \
    PT_HD vec3& operator op##=(vec3& a, float b) { a = a op b; return a; }                                 \
    PT_HD vec4& operator op##=(vec4& a, const vec4& b) { a = a op b; return a; }                           \
    PT_HD vec4& operator op##=(vec4& a, float b) { a = a op b; return a; }
PT_V_OP(+) PT_V_OP(-) PT_V_OP(*) PT_V_OP(/)
#undef PT_V_OP
PT_HD vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
PT_HD vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
PT_HD vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

/* ---- geometric ---------------------------------------------------------------------------------------------- */
PT_HD float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
PT_HD float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
PT_HD float length(float a) { return pt_abs(a); }
PT_HD float length(const vec2& a) { return sqrt(dot(a, a)); }
PT_HD float length(const vec3& a) { return sqrt(dot(a, a)); }
PT_HD float length(const vec4& a) { return sqrt(dot(a, a)); }
PT_HD float distance(const vec2& a, const vec2& b) { return length(a - b); }
PT_HD float distance(const vec3& a, const vec3& b) { return length(a - b); }
PT_HD vec2 normalize(const vec2& a) { return a / length(a); }
PT_HD vec3 normalize(const vec3& a) { return a / length(a); }
PT_HD vec4 normalize(const vec4& a) { return a / length(a); }
PT_HD vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
PT_HD vec3 reflect(const vec3& i, const vec3& n) { return i - 2.0f * dot(n, i) * n; }
PT_HD vec2 reflect(const vec2& i, const vec2& n) { return i - 2.0f * dot(n, i) * n; }
PT_HD vec3 faceforward(const vec3& n, const vec3& i, const vec3& nref) { return dot(nref, i) < 0.0f ? n : -n; }
PT_HD vec3 refract(const vec3& i, const vec3& n, float eta) {
    const float ndi = dot(n, i), k = 1.0f - eta * eta * (1.0f - ndi * ndi);
    return k < 0.0f ? vec3(0.0f) : eta * i - (eta * ndi + sqrt(k)) * n;
}
PT_HD float distance(float a, float b) { return pt_abs(a - b); }
PT_HD vec2 smoothstep(float e0, float e1, const vec2& x) { return vec2(smoothstep(e0, e1, x.x), smoothstep(e0, e1, x.y)); }
PT_HD vec3 smoothstep(float e0, float e1, const vec3& x) { return vec3(smoothstep(e0, e1, x.x), smoothstep(e0, e1, x.y), smoothstep(e0, e1, x.z)); }
PT_HD vec2 mix(const vec2& x, const vec2& y, const vec2& a) { return x * (vec2(1.0f) - a) + y * a; }
PT_HD vec4 mix(const vec4& x, const vec4& y, const vec4& a) { return x * (vec4(1.0f) - a) + y * a; }
PT_HD vec4 fma(const vec4& a, const vec4& b, const vec4& c) { return vec4(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y), fma(a.z, b.z, c.z), fma(a.w, b.w, c.w)); }
PT_HD vec4 clamp(const vec4& x, const vec4& lo, const vec4& hi) { return min(max(x, lo), hi); }

PT_HD vec2 clamp(const vec2& x, float lo, float hi) { return min(max(x, lo), hi); }
PT_HD vec3 clamp(const vec3& x, float lo, float hi) { return min(max(x, lo), hi); }
PT_HD vec4 clamp(const vec4& x, float lo, float hi) { return min(max(x, lo), hi); }
PT_HD vec2 clamp(const vec2& x, const vec2& lo, const vec2& hi) { return min(max(x, lo), hi); }
PT_HD vec3 clamp(const vec3& x, const vec3& lo, const vec3& hi) { return min(max(x, lo), hi); }
PT_HD vec2 mix(const vec2& x, const vec2& y, float a) { return x * (1.0f - a) + y * a; }
PT_HD vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
PT_HD vec4 mix(const vec4& x, const vec4& y, float a) { return x * (1.0f - a) + y * a; }
PT_HD vec3 mix(const vec3& x, const vec3& y, const vec3& a) { return x * (vec3(1.0f) - a) + y * a; }
PT_HD vec2 fma(const vec2& a, const vec2& b, const vec2& c) { return vec2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
PT_HD vec3 fma(const vec3& a, const vec3& b, const vec3& c) {
    return vec3(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y), fma(a.z, b.z, c.z));
}


/* ---- matrices: column-major like GLSL, m[c] is column c, m[c][r] one element (GLSL 4.50 5.4.2, 5.10) ------------- */
struct mat2 {
    vec2 c[2];
    PT_HD mat2() {}
    PT_HD explicit mat2(float s) { c[0] = vec2(s, 0.0f); c[1] = vec2(0.0f, s); }
    PT_HD mat2(float a, float b, float d, float e) { c[0] = vec2(a, b); c[1] = vec2(d, e); }
    PT_HD mat2(const vec2& c0, const vec2& c1) { c[0] = c0; c[1] = c1; }
    PT_HD explicit mat2(const mat3& m);
    PT_HD vec2& operator[](int i) { return c[i]; }
    PT_HD const vec2& operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    PT_HD mat3() {}
    PT_HD explicit mat3(float s) { c[0] = vec3(s, 0.0f, 0.0f); c[1] = vec3(0.0f, s, 0.0f); c[2] = vec3(0.0f, 0.0f, s); }
    PT_HD mat3(float a, float b, float d, float e, float f, float g, float h, float i, float j) { c[0] = vec3(a, b, d); c[1] = vec3(e, f, g); c[2] = vec3(h, i, j); }
    PT_HD mat3(const vec3& c0, const vec3& c1, const vec3& c2) { c[0] = c0; c[1] = c1; c[2] = c2; }
    PT_HD explicit mat3(const mat4& m);
    PT_HD explicit mat3(const mat2& m) { c[0] = vec3(m.c[0], 0.0f); c[1] = vec3(m.c[1], 0.0f); c[2] = vec3(0.0f, 0.0f, 1.0f); }
    PT_HD vec3& operator[](int i) { return c[i]; }
    PT_HD const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    PT_HD mat4() {}
    PT_HD explicit mat4(float s) { c[0] = vec4(s, 0.0f, 0.0f, 0.0f); c[1] = vec4(0.0f, s, 0.0f, 0.0f); c[2] = vec4(0.0f, 0.0f, s, 0.0f); c[3] = vec4(0.0f, 0.0f, 0.0f, s); }
    PT_HD mat4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3, float c0, float c1, float c2, float c3,
               float d0, float d1, float d2, float d3) { c[0] = vec4(a0, a1, a2, a3); c[1] = vec4(b0, b1, b2, b3); c[2] = vec4(c0, c1, c2, c3); c[3] = vec4(d0, d1, d2, d3); }
    PT_HD mat4(const vec4& c0, const vec4& c1, const vec4& c2, const vec4& c3) { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }
    PT_HD explicit mat4(const mat3& m) { c[0] = vec4(m.c[0], 0.0f); c[1] = vec4(m.c[1], 0.0f); c[2] = vec4(m.c[2], 0.0f); c[3] = vec4(0.0f, 0.0f, 0.0f, 1.0f); }
    PT_HD vec4& operator[](int i) { return c[i]; }
    PT_HD const vec4& operator[](int i) const { return c[i]; }
};
PT_HD mat2::mat2(const mat3& m) { c[0] = vec2(m.c[0]); c[1] = vec2(m.c[1]); }
PT_HD mat3::mat3(const mat4& m) { c[0] = vec3(m.c[0]); c[1] = vec3(m.c[1]); c[2] = vec3(m.c[2]); }
/* M * v: linear combination of the columns, accumulated left to right;  v * M: one dot product per column */
PT_HD vec2 operator*(const mat2& m, const vec2& v) { return m.c[0] * v.x + m.c[1] * v.y; }
PT_HD vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
PT_HD vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
PT_HD vec2 operator*(const vec2& v, const mat2& m) { return vec2(dot(v, m.c[0]), dot(v, m.c[1])); }
PT_HD vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
PT_HD vec4 operator*(const vec4& v, const mat4& m) { return vec4(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]), dot(v, m.c[3])); }
PT_HD mat2 operator*(const mat2& a, const mat2& b) { return mat2(a * b.c[0], a * b.c[1]); }
PT_HD mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
PT_HD mat4 operator*(const mat4& a, const mat4& b) { return mat4(a * b.c[0], a * b.c[1], a * b.c[2], a * b.c[3]); }
#define PT_M_SCALAR(M, N, ...)                                                                            \
    PT_HD M operator*(const M& a, float s) { M r; for (int i = 0; i < N; i++) r.c[i] = a.c[i] * s; return r; } \
    PT_HD M operator*(float s, const M& a) { M r; for (int i = 0; i < N; i++) r.c[i] = s * a.c[i]; return r; } \
    PT_HD M operator/(const M& a, float s) { M r; for (int i = 0; i < N; i++) r.c[i] = a.c[i] / s; return r; } \
    PT_HD M operator+(const M& a, const M& b) { M r; for (int i = 0; i < N; i++) r.c[i] = a.c[i] + b.c[i]; return r; } \
    PT_HD M operator-(const M& a, const M& b) { M r; for (int i = 0; i < N; i++) r.c[i] = a.c[i] - b.c[i]; return r; } \
    PT_HD M operator-(const M& a) { M r; for (int i = 0; i < N; i++) r.c[i] = -a.c[i]; return r; }         \
    PT_HD M matrixCompMult(const M& a, const M& b) { M r; for (int i = 0; i < N; i++) r.c[i] = a.c[i] * b.c[i]; return r; } \
    PT_HD M& operator*=(M& a, const M& b) { a = a * b; return a; }                                         \
    PT_HD M& operator*=(M& a, float s) { a = a * s; return a; }
PT_M_SCALAR(mat2, 2) PT_M_SCALAR(mat3, 3) PT_M_SCALAR(mat4, 4)
#undef PT_M_SCALAR
PT_HD vec2& operator*=(vec2& v, const mat2& m) { v = v * m; return v; }
PT_HD vec3& operator*=(vec3& v, const mat3& m) { v = v * m; return v; }
PT_HD vec4& operator*=(vec4& v, const mat4& m) { v = v * m; return v; }
PT_HD mat2 transpose(const mat2& m) { return mat2(m.c[0].x, m.c[1].x, m.c[0].y, m.c[1].y); }
PT_HD mat3 transpose(const mat3& m) { return mat3(m.c[0].x, m.c[1].x, m.c[2].x, m.c[0].y, m.c[1].y, m.c[2].y, m.c[0].z, m.c[1].z, m.c[2].z); }
PT_HD mat4 transpose(const mat4& m) {
    return mat4(m.c[0].x, m.c[1].x, m.c[2].x, m.c[3].x, m.c[0].y, m.c[1].y, m.c[2].y, m.c[3].y, m.c[0].z, m.c[1].z, m.c[2].z, m.c[3].z,
                m.c[0].w, m.c[1].w, m.c[2].w, m.c[3].w);
}
PT_HD float determinant(const mat2& m) { return m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y; }
PT_HD float determinant(const mat3& m) {
    return m.c[0].x * (m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z) - m.c[1].x * (m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z) +
           m.c[2].x * (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z);
}
PT_HD mat2 inverse(const mat2& m) { const float d = 1.0f / determinant(m); return mat2(m.c[1].y * d, -m.c[0].y * d, -m.c[1].x * d, m.c[0].x * d); }
PT_HD mat3 inverse(const mat3& m) {
    const float d = 1.0f / determinant(m);
    return mat3((m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z) * d, -(m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z) * d, (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z) * d,
                -(m.c[1].x * m.c[2].z - m.c[2].x * m.c[1].z) * d, (m.c[0].x * m.c[2].z - m.c[2].x * m.c[0].z) * d, -(m.c[0].x * m.c[1].z - m.c[1].x * m.c[0].z) * d,
                (m.c[1].x * m.c[2].y - m.c[2].x * m.c[1].y) * d, -(m.c[0].x * m.c[2].y - m.c[2].x * m.c[0].y) * d, (m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y) * d);
}
PT_HD mat2 outerProduct(const vec2& c, const vec2& r) { return mat2(c * r.x, c * r.y); }
PT_HD mat3 outerProduct(const vec3& c, const vec3& r) { return mat3(c * r.x, c * r.y, c * r.z); }

/* ---- helpers of shader.comp that snippets may call (shader.comp:7-10) ---------------------------------------- */
#define MINDIST 1e-5f
#define MAXDIST 1e5f
#define PI 3.141592653589792623810034526344f
#define ONEBYTHREE 0.3333333f

} /* namespace ptglsl */
#endif /* PT_GLSL_H */
