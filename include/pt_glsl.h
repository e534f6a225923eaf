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
 * build the MUFU intrinsics.  Builtins provided: sin cos acos exp exp2 log log2 pow sqrt inversesqrt abs sign floor
 * ceil fract mod min max clamp mix step smoothstep fma length distance dot cross normalize reflect.
 * (tan/asin/atan are not used by shader.comp or any shipped snippet and are not provided.)
 */
#ifndef PT_GLSL_H
#define PT_GLSL_H

#include "pt_math.h"

#if defined(PT_FAST) && defined(__CUDA_ARCH__)
#define PT_FN_SIN(x) __sinf(x)
#define PT_FN_COS(x) __cosf(x)
#define PT_FN_ACOS(x) acosf(x)
#define PT_FN_EXP(x) __expf(x)
#define PT_FN_EXP2(x) exp2f(x)
#define PT_FN_LOG(x) __logf(x)
#define PT_FN_LOG2(x) __log2f(x)
#define PT_FN_POW(x, y) __powf(x, y)
#define PT_FN_SQRT(x) sqrtf(x)
#define PT_FN_RSQRT(x) rsqrtf(x)
#define PT_FN_FMA(a, b, c) fmaf(a, b, c)
#else
#define PT_FN_SIN(x) pt_sin(x)
#define PT_FN_COS(x) pt_cos(x)
#define PT_FN_ACOS(x) pt_acos(x)
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

#define PT_SW2(T2, a, b) PT_HD T2 a##b() const;
#define PT_SW3(T3, a, b, c) PT_HD T3 a##b##c() const;

struct vec2 {
    float x, y;
    PT_HD vec2() : x(0.0f), y(0.0f) {}
    PT_HD explicit vec2(float s) : x(s), y(s) {}
    PT_HD vec2(float x_, float y_) : x(x_), y(y_) {}
    PT_HD explicit vec2(const vec3& v);
    PT_HD float& operator[](int i) { return i == 0 ? x : y; }
    PT_HD float operator[](int i) const { return i == 0 ? x : y; }
    PT_SW2(vec2, x, x) PT_SW2(vec2, x, y) PT_SW2(vec2, y, x) PT_SW2(vec2, y, y)
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
};

struct vec4 {
    float x, y, z, w;
    PT_HD vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    PT_HD explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    PT_HD vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    PT_HD vec4(const vec3& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    PT_HD vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    PT_HD float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    PT_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    PT_HD vec3 xyz() const { return vec3(x, y, z); }
    PT_HD vec2 xy() const { return vec2(x, y); }
    PT_HD vec2 zw() const { return vec2(z, w); }
    PT_HD vec4 xyzw() const { return *this; }
};

PT_HD vec2::vec2(const vec3& v) : x(v.x), y(v.y) {}
PT_HD vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

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
PT_HD float min(float x, float y) { return (y < x) ? y : x; }
PT_HD float max(float x, float y) { return (x < y) ? y : x; }
PT_HD int min(int x, int y) { return (y < x) ? y : x; }
PT_HD int max(int x, int y) { return (x < y) ? y : x; }
PT_HD float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
PT_HD float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
PT_HD float step(float e, float x) { return (x < e) ? 0.0f : 1.0f; }
PT_HD float mod(float x, float y) { return x - y * pt_floor(x / y); }
PT_HD float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

/* ---- componentwise lifting ---------------------------------------------------------------------------------- */
#define PT_V_UN(f)                                                          \
    PT_HD vec2 f(const vec2& a) { return vec2(f(a.x), f(a.y)); }            \
    PT_HD vec3 f(const vec3& a) { return vec3(f(a.x), f(a.y), f(a.z)); }    \
    PT_HD vec4 f(const vec4& a) { return vec4(f(a.x), f(a.y), f(a.z), f(a.w)); }
PT_V_UN(sin) PT_V_UN(cos) PT_V_UN(acos) PT_V_UN(exp) PT_V_UN(exp2) PT_V_UN(log) PT_V_UN(log2) PT_V_UN(sqrt)
PT_V_UN(inversesqrt) PT_V_UN(abs) PT_V_UN(floor) PT_V_UN(ceil) PT_V_UN(fract) PT_V_UN(sign)
#undef PT_V_UN

/* f(vec, vec) and f(vec, float) */
#define PT_V_BIN(f)                                                                               \
    PT_HD vec2 f(const vec2& a, const vec2& b) { return vec2(f(a.x, b.x), f(a.y, b.y)); }         \
    PT_HD vec3 f(const vec3& a, const vec3& b) { return vec3(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z)); } \
    PT_HD vec4 f(const vec4& a, const vec4& b) { return vec4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w)); } \
    PT_HD vec2 f(const vec2& a, float b) { return vec2(f(a.x, b), f(a.y, b)); }                   \
    PT_HD vec3 f(const vec3& a, float b) { return vec3(f(a.x, b), f(a.y, b), f(a.z, b)); }        \
    PT_HD vec4 f(const vec4& a, float b) { return vec4(f(a.x, b), f(a.y, b), f(a.z, b), f(a.w, b)); }
PT_V_BIN(min) PT_V_BIN(max) PT_V_BIN(mod) PT_V_BIN(pow)
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

/* ---- helpers of shader.comp that snippets may call (shader.comp:7-10) ---------------------------------------- */
#define MINDIST 1e-5f
#define MAXDIST 1e5f
#define PI 3.141592653589792623810034526344f
#define ONEBYTHREE 0.3333333f

} /* namespace ptglsl */
#endif /* PT_GLSL_H */
