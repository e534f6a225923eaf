/* pt_math.h -- pinned IEEE-754 binary32 math for libpt_cuda and its CPU oracle.
 *
 * Why this exists: GLSL leaves sin/cos/acos/exp/log/pow accuracy to the driver, so
 * "the" float output of the reference's shader.comp is only defined up to a tolerance.  This header freezes ONE
 * interpretation: every function below is built only from + - * / sqrt fma and integer bit
 * operations, each of which is correctly rounded on x86-64 (SSE/FMA, no contraction) and on sm_100a (the _rn
 * intrinsics are never contracted or approximated, whatever -fmad / -use_fast_math say).  The same inputs
 * therefore give the same bits on the CPU oracle and in the strict CUDA kernels.
 *
 * Polynomial coefficients come from tools/gen_math_coeffs.py (least squares on Chebyshev nodes, rounded to fp32).
 * Accuracy (vs. float64 libm, see tests/test_pt_math.py): sin/cos <= 2 ulp for |x| < 1e4, acos <= 3 ulp,
 * exp2 <= 2 ulp, log2 <= 3 ulp away from 1.
 *
 * Freestanding: compiles under g++ (host), nvcc (host+device) and NVRTC (device only, no libc headers).
 */
#ifndef PT_MATH_H
#define PT_MATH_H

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD inline
#endif

#if !defined(__CUDACC_RTC__)
#include <math.h>
#include <string.h>
#endif

/* ---- correctly rounded primitives ------------------------------------------------------------------------- */
#if defined(__CUDA_ARCH__)
PT_HD float pt_add(float a, float b) { return __fadd_rn(a, b); }
PT_HD float pt_sub(float a, float b) { return __fsub_rn(a, b); }
PT_HD float pt_mul(float a, float b) { return __fmul_rn(a, b); }
PT_HD float pt_div(float a, float b) { return __fdiv_rn(a, b); }
PT_HD float pt_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PT_HD float pt_sqrt(float a) { return __fsqrt_rn(a); }
PT_HD float pt_floor(float a) { return floorf(a); }
PT_HD float pt_ceil(float a) { return ceilf(a); }
PT_HD float pt_abs(float a) { return fabsf(a); }
PT_HD unsigned pt_f2u(float a) { return __float_as_uint(a); }
PT_HD float pt_u2f(unsigned a) { return __uint_as_float(a); }
/* float -> int, truncating; NaN -> 0, out of range saturates (this is what cvt.rzi.s32.f32 does) */
PT_HD int pt_f2i(float a) { return __float2int_rz(a); }
#else
PT_HD float pt_add(float a, float b) { return a + b; }
PT_HD float pt_sub(float a, float b) { return a - b; }
PT_HD float pt_mul(float a, float b) { return a * b; }
PT_HD float pt_div(float a, float b) { return a / b; }
PT_HD float pt_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
PT_HD float pt_sqrt(float a) { return __builtin_sqrtf(a); }
PT_HD float pt_floor(float a) { return __builtin_floorf(a); }
PT_HD float pt_ceil(float a) { return __builtin_ceilf(a); }
PT_HD float pt_abs(float a) { return __builtin_fabsf(a); }
PT_HD unsigned pt_f2u(float a) { unsigned u; memcpy(&u, &a, 4); return u; }
PT_HD float pt_u2f(unsigned a) { float f; memcpy(&f, &a, 4); return f; }
PT_HD int pt_f2i(float a) {
    if (!(a == a)) return 0;
    if (a >= 2147483648.0f) return 2147483647;
    if (a <= -2147483648.0f) return (-2147483647 - 1);
    return (int)a;
}
#endif

#define PT_PI_F 3.141592741e+00f      /* fp32 nearest to pi = 0x40490FDB, same as the shader's PI literal */
#define PT_PIO2_F 1.570796371e+00f

/* ---- sin / cos -------------------------------------------------------------------------------------------- */
/* Reduce x to r in [-pi/4, pi/4] and a quadrant: x = q*(pi/2) + r.  3-term Cody-Waite with fma. */
PT_HD float pt__reduce_pio2(float x, int* q) {
    if (!(pt_abs(x) < 4194304.0f)) {
        /* huge or non-finite: fold by 2*pi first (inaccurate but deterministic); inf/NaN fall through to NaN */
        x = pt_fma(-6.283185482e+00f, pt_floor(pt_div(x, 6.283185482e+00f)), x);
        if (!(pt_abs(x) < 4194304.0f)) { *q = 0; return pt_u2f(0x7fc00000u); }
    }
    float kf = pt_add(pt_mul(x, 6.366197467e-01f), 12582912.0f); /* round to nearest integer */
    kf = pt_sub(kf, 12582912.0f);
    float r = pt_fma(kf, -1.570796371e+00f, x);
    r = pt_fma(kf, 4.371138829e-08f, r);
    r = pt_fma(kf, 1.715124510e-15f, r);
    *q = (int)kf;
    return r;
}
PT_HD float pt__sin_poly(float r) {
    float z = pt_mul(r, r);
    float p = pt_fma(z, 2.725813147e-06f, -1.984017144e-04f);
    p = pt_fma(z, p, 8.333331905e-03f);
    p = pt_fma(z, p, -1.666666716e-01f);
    return pt_fma(pt_mul(r, z), p, r);
}
PT_HD float pt__cos_poly(float r) {
    float z = pt_mul(r, r);
    float p = pt_fma(z, -2.730780011e-07f, 2.480067087e-05f);
    p = pt_fma(z, p, -1.388888806e-03f);
    p = pt_fma(z, p, 4.166666791e-02f);
    return pt_fma(pt_mul(z, z), p, pt_fma(z, -0.5f, 1.0f));
}
PT_HD float pt_sin(float x) {
    int q;
    float r = pt__reduce_pio2(x, &q);
    float v = (q & 1) ? pt__cos_poly(r) : pt__sin_poly(r);
    return (q & 2) ? -v : v;
}
PT_HD float pt_cos(float x) {
    int q;
    float r = pt__reduce_pio2(x, &q);
    float v = (q & 1) ? pt__sin_poly(r) : pt__cos_poly(r);
    return ((q + 1) & 2) ? -v : v;
}

/* ---- acos ------------------------------------------------------------------------------------------------- */
PT_HD float pt__asin_poly(float x) { /* |x| <= 0.5 */
    float z = pt_mul(x, x);
    float p = pt_fma(z, 3.341218084e-02f, 1.733727753e-02f);
    p = pt_fma(z, p, 3.105503134e-02f);
    p = pt_fma(z, p, 4.460414127e-02f);
    p = pt_fma(z, p, 7.500075549e-02f);
    p = pt_fma(z, p, 1.666666716e-01f);
    return pt_fma(pt_mul(x, z), p, x);
}
PT_HD float pt_acos(float x) {
    float a = pt_abs(x);
    if (!(a <= 1.0f)) return pt_u2f(0x7fc00000u); /* NaN outside [-1,1] (and for NaN) */
    if (a <= 0.5f) return pt_sub(PT_PIO2_F, pt__asin_poly(x));
    float s = pt_sqrt(pt_mul(pt_sub(1.0f, a), 0.5f)); /* acos(a) = 2 asin(sqrt((1-a)/2)) */
    float t = pt_mul(2.0f, pt__asin_poly(s));
    return (x < 0.0f) ? pt_sub(PT_PI_F, t) : t;
}

/* ---- exp2 / log2 ------------------------------------------------------------------------------------------ */
PT_HD float pt_exp2(float x) {
    if (!(x == x)) return x;
    if (x >= 128.0f) return pt_u2f(0x7f800000u);
    if (x < -150.0f) return 0.0f;
    float kf = pt_sub(pt_add(x, 12582912.0f), 12582912.0f); /* nearest integer */
    float f = pt_sub(x, kf);                                   /* exact, in [-0.5, 0.5] */
    float p = pt_fma(f, 1.530370173e-05f, 1.546144777e-04f);
    p = pt_fma(f, p, 1.333347871e-03f);
    p = pt_fma(f, p, 9.618056938e-03f);
    p = pt_fma(f, p, 5.550410971e-02f);
    p = pt_fma(f, p, 2.402265072e-01f);
    p = pt_fma(f, p, 6.931471825e-01f);
    p = pt_fma(f, p, 1.0f);
    int k = (int)kf;
    /* scale by 2^k in two steps so that subnormal results round once at the end */
    int k1 = k / 2, k2 = k - k1;
    float s1 = pt_u2f((unsigned)(k1 + 127) << 23);
    float s2 = pt_u2f((unsigned)(k2 + 127) << 23);
    return pt_mul(pt_mul(p, s1), s2);
}
PT_HD float pt_log2(float x) {
    if (!(x == x)) return x;
    if (x < 0.0f) return pt_u2f(0x7fc00000u);
    if (x == 0.0f) return pt_u2f(0xff800000u);
    unsigned u = pt_f2u(x);
    if (u == 0x7f800000u) return x;
    int e = 0;
    if (u < 0x00800000u) { /* subnormal: scale up by 2^24 */
        x = pt_mul(x, 16777216.0f);
        u = pt_f2u(x);
        e = -24;
    }
    /* mantissa m in [sqrt(1/2), sqrt(2)) */
    unsigned v = u - 0x3f3504f3u;    /* 0x3f3504f3 = sqrt(1/2); two's-complement wrap is intended */
    e += ((int)v) >> 23;             /* arithmetic shift = floor division by 2^23 */
    float m = pt_u2f((v & 0x007fffffu) + 0x3f3504f3u);
    float s = pt_div(pt_sub(m, 1.0f), pt_add(m, 1.0f));
    float z = pt_mul(s, s);
    float p = pt_fma(z, 3.403142393e-01f, 4.116994441e-01f);
    p = pt_fma(z, p, 5.770830512e-01f);
    p = pt_fma(z, p, 9.617967010e-01f);
    p = pt_fma(z, p, 2.885390043e+00f);
    return pt_fma(s, p, (float)e);
}

/* ---- asin / atan / atan2 / tan (not used by shader.comp or the shipped snippets; third-party SDF code uses them) ------ */
PT_HD float pt_asin(float x) {
    float a = pt_abs(x);
    if (!(a <= 1.0f)) return pt_u2f(0x7fc00000u);
    if (a <= 0.5f) return pt__asin_poly(x);
    float s = pt_sqrt(pt_mul(pt_sub(1.0f, a), 0.5f)); /* asin(a) = pi/2 - 2 asin(sqrt((1-a)/2)) */
    float t = pt_sub(PT_PIO2_F, pt_mul(2.0f, pt__asin_poly(s)));
    return (x < 0.0f) ? -t : t;
}
/* atan: Cephes' single-precision scheme -- reduce to |r| <= tan(pi/8) (r = -1/a above tan(3pi/8), (a-1)/(a+1) above
 * tan(pi/8)), then r + r^3 P(r^2) with its four published coefficients; <= 2 ulp */
PT_HD float pt_atan(float x) {
    if (!(x == x)) return x;
    float a = pt_abs(x), base = 0.0f, r = a;
    if (a > 2.414213562373095f) { base = PT_PIO2_F; r = pt_div(-1.0f, a); }
    else if (a > 0.4142135623730950f) { base = 7.853981853e-01f; r = pt_div(pt_sub(a, 1.0f), pt_add(a, 1.0f)); }
    float z = pt_mul(r, r);
    float p = pt_fma(z, 8.05374449538e-2f, -1.38776856032e-1f);
    p = pt_fma(z, p, 1.99777106478e-1f);
    p = pt_fma(z, p, -3.33329491539e-1f);
    float y = pt_add(base, pt_fma(pt_mul(p, z), r, r));
    return (x < 0.0f) ? -y : y;
}
/* GLSL atan(y, x): the angle of (x, y) in (-pi, pi]; atan(0, 0) is undefined in GLSL, 0 here */
PT_HD float pt_atan2(float y, float x) {
    if (!(x == x) || !(y == y)) return pt_u2f(0x7fc00000u);
    if (x == 0.0f) return (y > 0.0f) ? PT_PIO2_F : ((y < 0.0f) ? -PT_PIO2_F : 0.0f);
    float t = pt_atan(pt_div(y, x));
    if (x > 0.0f) return t;
    return (y < 0.0f) ? pt_sub(t, PT_PI_F) : pt_add(t, PT_PI_F);
}
PT_HD float pt_sin(float x);
PT_HD float pt_cos(float x);
PT_HD float pt_tan(float x) { return pt_div(pt_sin(x), pt_cos(x)); }

/* ---- the GLSL transcendental set, per App. F --------------------------------------------------------------- */
PT_HD float pt_exp(float x) { return pt_exp2(pt_mul(x, 1.442695022e+00f)); }
PT_HD float pt_log(float x) { return pt_mul(pt_log2(x), 6.931471825e-01f); }
PT_HD float pt_pow(float x, float y) { return pt_exp2(pt_mul(y, pt_log2(x))); }
PT_HD float pt_rsqrt(float x) { return pt_div(1.0f, pt_sqrt(x)); }

#endif /* PT_MATH_H */
