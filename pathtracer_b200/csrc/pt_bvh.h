/* pt_bvh.h -- bounding-volume hierarchy over the scene's spheres, boxes and lenses.
 *
 * The reference finds the closest hit by brute force over every primitive (Intersection shader.comp:862-934 and
 * LightSourceVisibilityCheck 1121-1216), which is right for its shipped scenes (<= 8 primitives) and quadratic
 * pain at the uniform block's capacity (170 spheres).  B200 has no RT cores, so this is a plain binary BVH walked
 * with a per-thread stack.  It is an ACCELERATION of the same search, not a different one:
 *   - every leaf is ONE primitive and runs the very same intersection routine as the brute-force loop;
 *   - the winner is "smallest t, ties to the lowest object index", which is what the reference's in-order
 *     `if (t < hit.t)` scan yields, so the visiting order cannot change the result;
 *   - node boxes are padded (pt_bvh.cpp) by more than the fp32 noise of the primitives' own quadratics, so culling
 *     a node never removes a primitive the brute-force scan would have reported.
 * Planes are infinite and stay outside the tree (scanned first, in order).  So do the Dupin cyclides (scanned last, in
 * order): the reference's quartic solver (shader.comp:505-541) returns spurious roots anywhere along a ray that merely
 * passes a cyclide's bounding sphere -- "hits" in mid-air, many units from the surface, which the in-order scan
 * reports and which no box around the surface can anticipate (found by tests/test_bvh.py's oracle-through-the-tree
 * check).  They are the rarest and by far the costliest primitive, behind a bounding-sphere cull of their own.
 *
 * Layout (built on the host by pt_bvh.cpp, appended to the device copy of the uniform block at float offset
 * PT_BVH_UBO_OFF): a 16-byte header, n-1 inner nodes of 64 bytes holding BOTH children's boxes, then a copy of the
 * PtDevScene record pool (per-lane primitive indices diverge, and divergent constant-bank reads serialise; from
 * global memory the records are __ldg'd as float4 and stay in L1).
 *      float4 header = (C.x, C.y, C.z, R - S)   centre and radius of everything bounded; S = the scale the static
 *                                               padding of the boxes covers (pt_bvh.cpp)
 *      float4 a = (c0.min.x, c0.min.y, c0.min.z, c0.max.x)
 *      float4 b = (c0.max.y, c0.max.z, c1.min.x, c1.min.y)
 *      float4 c = (c1.min.z, c1.max.x, c1.max.y, c1.max.z)
 *      int4   d = (ref0, ref1, 0, 0)      ref >= 0: inner node index;  ref < 0: ~ref = (type << 16) | index in type
 * This header is shared by g++ (builder, host-side checker in tests/), nvcc and NVRTC.
 */
#ifndef PT_BVH_H
#define PT_BVH_H

#include "pt_dev_scene.h"

#define PT_BVH_UBO_OFF 4100      /* float offset inside the device ubo buffer: 16 400 B, 16-byte aligned */
#define PT_BVH_HEADER_FLOATS 4
#define PT_BVH_NODE_FLOATS 16
/* fp32 noise of SphereIntersection's `b*b - 4*c`, relative to D^2 (D = distance of the ray origin): <= 6e-7 worst case
 * (three roundings in each dot product).  tests/bvh_check.cpp sees the first tree/scan mismatches at 2^-21 = 4.8e-7
 * (1 in 10^7 rays, D = 68 000) and none at 2^-20; 2^-18 = 3.8e-6 keeps a factor 6 over the bound. */
#define PT_BVH_KAPPA 3.8146973e-6f
#define PT_BVH_SQRT_KAPPA 1.953125e-3f
#define PT_BVH_STACK 32          /* the builder bounds the depth at PT_BVH_MAX_DEPTH */
#define PT_BVH_MAX_DEPTH 28
#define PT_BVH_MAX_PRIMS 256     /* > the 170 spheres the uniform block can describe */
#define PT_BVH_MAX_FLOATS (PT_BVH_HEADER_FLOATS + PT_BVH_NODE_FLOATS * PT_BVH_MAX_PRIMS + PT_DEV_POOL_FLOATS + 4)
#ifndef PT_BVH_WHILE_WHILE
#define PT_BVH_WHILE_WHILE 0     /* loop shape of pt_bvh_traverse; A/B on B200: profiles/r01_bvh */
#endif
#define PT_BVH_DEFAULT_MIN_PRIMS 12 /* spheres + boxes + lenses from which the tree replaces the brute-force scan */

enum { PT_BVH_SPHERE = 0, PT_BVH_BOX = 1, PT_BVH_LENS = 2 };

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define PT_BVH_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PT_BVH_HD inline
#endif

struct PtBvhF4 { float x, y, z, w; };

PT_BVH_HD PtBvhF4 pt_bvh_load4(const float* base, int index4) {
#ifdef __CUDA_ARCH__
    const float4 v = __ldg(reinterpret_cast<const float4*>(base) + index4);
    PtBvhF4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    PtBvhF4 r; r.x = base[4 * index4]; r.y = base[4 * index4 + 1]; r.z = base[4 * index4 + 2]; r.w = base[4 * index4 + 3];
    return r;
#endif
}

PT_BVH_HD int pt_bvh_float_as_int(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    union { float f; int i; } u; u.f = f; return u.i;
#endif
}

/* ray / box slabs, one fused multiply-add per plane: (lo - o) / d = lo * (1/d) - o * (1/d), the second product hoisted
 * per ray.  1/d is clamped to +-1e25 (pt_bvh_traverse) so that a direction component of 0 yields huge finite products
 * of the right sign instead of inf - inf.  The rounding differs from the primitives' own arithmetic by ulps; the
 * padding of the boxes is orders of magnitude larger. */
PT_BVH_HD void pt_bvh_slab(float lox, float loy, float loz, float hix, float hiy, float hiz, float nxl, float nyl, float nzl,
                           float nxh, float nyh, float nzh, float ix, float iy, float iz, float& tn, float& tf) {
    const float ax = fmaf(lox, ix, nxl), bx = fmaf(hix, ix, nxh);
    const float ay = fmaf(loy, iy, nyl), by = fmaf(hiy, iy, nyh);
    const float az = fmaf(loz, iz, nzl), bz = fmaf(hiz, iz, nzh);
    tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}

/* Walks the tree (bvh = header, nodes) for the ray (o, d); calls leaf(ref) with ref = (type << 16) | index for every
 * primitive whose padded box the ray enters no later than tBest.  leaf() is expected to shrink tBest (it aliases the
 * hit record's t).
 * The static padding of the boxes covers ray origins up to S from the primitives.  A ray that starts farther away --
 * off an infinite plane, toward the horizon -- sees primitives whose quadratic is mostly rounding noise (a sphere
 * "is" wherever b*b - 4*c happens to come out non-negative, units off at D = 5000); for it every box is inflated by
 * e = sqrt(kappa) * (D - S), which bounds the growth of that noise band (d/dD of sqrt(r^2 + kappa D^2) <= sqrt(kappa)).
 * The inflation is folded into the ray origin per axis side, so the slab test costs the same. */
template <class Leaf>
PT_BVH_HD void pt_bvh_traverse(const float* bvh, float ox, float oy, float oz, float dx, float dy, float dz,
                               const float& tBest, Leaf leaf) {
    const PtBvhF4 hdr = pt_bvh_load4(bvh, 0);
    const float* nodes = bvh + PT_BVH_HEADER_FLOATS;
    const float cx = ox - hdr.x, cy = oy - hdr.y, cz = oz - hdr.z;
    const float infl = PT_BVH_SQRT_KAPPA * fmaxf(sqrtf(cx * cx + cy * cy + cz * cz) + hdr.w, 0.0f);
    const float ix = fminf(fmaxf(1.0f / dx, -1e25f), 1e25f), iy = fminf(fmaxf(1.0f / dy, -1e25f), 1e25f),
                iz = fminf(fmaxf(1.0f / dz, -1e25f), 1e25f);
    /* -(o + e) / d for the lower planes, -(o - e) / d for the upper ones: lo - (o + e) = (lo - e) - o */
    const float oxl = -(ox + infl) * ix, oyl = -(oy + infl) * iy, ozl = -(oz + infl) * iz;
    const float oxh = -(ox - infl) * ix, oyh = -(oy - infl) * iy, ozh = -(oz - infl) * iz;
    int sref[PT_BVH_STACK];
    float stn[PT_BVH_STACK];
    int sp = 0;
    int cur = 0;
#if PT_BVH_WHILE_WHILE
    /* "while-while": every lane first descends to its next leaf (the node code runs with all traversing lanes), then
     * the leaves are tested together (the primitive code runs with all lanes that found one) */
    bool done = false;
    while (!done) {
        while (cur >= 0) {
            const PtBvhF4 a = pt_bvh_load4(nodes, 4 * cur), b = pt_bvh_load4(nodes, 4 * cur + 1),
                          c = pt_bvh_load4(nodes, 4 * cur + 2), d = pt_bvh_load4(nodes, 4 * cur + 3);
            float n0, f0, n1, f1;
            pt_bvh_slab(a.x, a.y, a.z, a.w, b.x, b.y, oxl, oyl, ozl, oxh, oyh, ozh, ix, iy, iz, n0, f0);
            pt_bvh_slab(b.z, b.w, c.x, c.y, c.z, c.w, oxl, oyl, ozl, oxh, oyh, ozh, ix, iy, iz, n1, f1);
            const bool h0 = (n0 <= f0) && (f0 >= 0.0f) && (n0 <= tBest);
            const bool h1 = (n1 <= f1) && (f1 >= 0.0f) && (n1 <= tBest);
            const int r0 = pt_bvh_float_as_int(d.x), r1 = pt_bvh_float_as_int(d.y);
            if (h0 && h1) {
                const bool firstIs0 = n0 <= n1;
                if (sp < PT_BVH_STACK) {
                    sref[sp] = firstIs0 ? r1 : r0;
                    stn[sp] = firstIs0 ? n1 : n0;
                    sp++;
                }
                cur = firstIs0 ? r0 : r1;
            } else if (h0) {
                cur = r0;
            } else if (h1) {
                cur = r1;
            } else {
                float tn;
                do {
                    if (sp == 0) { done = true; cur = -1; break; }
                    --sp;
                    cur = sref[sp];
                    tn = stn[sp];
                } while (tn > tBest);
            }
        }
        if (!done) {
            leaf(~cur);
            float tn;
            do {
                if (sp == 0) { done = true; break; }
                --sp;
                cur = sref[sp];
                tn = stn[sp];
            } while (tn > tBest);
        }
    }
#else
    for (;;) {
        if (cur >= 0) {
            const PtBvhF4 a = pt_bvh_load4(nodes, 4 * cur), b = pt_bvh_load4(nodes, 4 * cur + 1),
                          c = pt_bvh_load4(nodes, 4 * cur + 2), d = pt_bvh_load4(nodes, 4 * cur + 3);
            float n0, f0, n1, f1;
            pt_bvh_slab(a.x, a.y, a.z, a.w, b.x, b.y, oxl, oyl, ozl, oxh, oyh, ozh, ix, iy, iz, n0, f0);
            pt_bvh_slab(b.z, b.w, c.x, c.y, c.z, c.w, oxl, oyl, ozl, oxh, oyh, ozh, ix, iy, iz, n1, f1);
            const bool h0 = (n0 <= f0) && (f0 >= 0.0f) && (n0 <= tBest);
            const bool h1 = (n1 <= f1) && (f1 >= 0.0f) && (n1 <= tBest);
            const int r0 = pt_bvh_float_as_int(d.x), r1 = pt_bvh_float_as_int(d.y);
            if (h0 && h1) {
                const bool firstIs0 = n0 <= n1;
                if (sp < PT_BVH_STACK) {
                    sref[sp] = firstIs0 ? r1 : r0;
                    stn[sp] = firstIs0 ? n1 : n0;
                    sp++;
                }
                cur = firstIs0 ? r0 : r1;
                continue;
            }
            if (h0) { cur = r0; continue; }
            if (h1) { cur = r1; continue; }
        } else {
            leaf(~cur);
        }
        float tn;
        do {
            if (sp == 0) return;
            --sp;
            cur = sref[sp];
            tn = stn[sp];
        } while (tn > tBest);
    }
#endif
}

#endif /* PT_BVH_H */
