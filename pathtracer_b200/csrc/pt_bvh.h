/* pt_bvh.h -- bounding-volume hierarchy over the scene's bounded primitives (spheres, boxes, lenses, cyclides).
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
 * Planes are infinite and stay outside the tree (scanned first, in order).
 *
 * Layout (built on the host by pt_bvh.cpp, appended to the device copy of the uniform block at float offset
 * PT_BVH_UBO_OFF): n-1 inner nodes of 64 bytes holding BOTH children's boxes, then a copy of the PtDevScene record
 * pool (per-lane primitive indices diverge, and divergent constant-bank reads serialise; from global memory the
 * records are __ldg'd as float4 and stay in L1).
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
#define PT_BVH_NODE_FLOATS 16
#define PT_BVH_STACK 32          /* the builder bounds the depth at PT_BVH_MAX_DEPTH */
#define PT_BVH_MAX_DEPTH 28
#define PT_BVH_MAX_PRIMS 256     /* > the 170 spheres the uniform block can describe */
#define PT_BVH_MAX_FLOATS (PT_BVH_NODE_FLOATS * PT_BVH_MAX_PRIMS + PT_DEV_POOL_FLOATS + 4)
#define PT_BVH_DEFAULT_MIN_PRIMS 12 /* bounded primitives from which the tree replaces the brute-force scan */

enum { PT_BVH_SPHERE = 0, PT_BVH_BOX = 1, PT_BVH_LENS = 2, PT_BVH_CYCLIDE = 3 };

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

/* ray / box slabs; fminf/fmaxf drop the NaN of 0 * inf (origin on a slab plane, direction parallel to it) */
PT_BVH_HD void pt_bvh_slab(float lox, float loy, float loz, float hix, float hiy, float hiz, float ox, float oy, float oz,
                           float ix, float iy, float iz, float& tn, float& tf) {
    const float ax = (lox - ox) * ix, bx = (hix - ox) * ix;
    const float ay = (loy - oy) * iy, by = (hiy - oy) * iy;
    const float az = (loz - oz) * iz, bz = (hiz - oz) * iz;
    tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}

/* Walks the tree for the ray (o, d); calls leaf(ref) with ref = (type << 16) | index for every primitive whose padded
 * box the ray enters no later than tBest.  leaf() is expected to shrink tBest (it aliases the hit record's t). */
template <class Leaf>
PT_BVH_HD void pt_bvh_traverse(const float* nodes, float ox, float oy, float oz, float dx, float dy, float dz,
                               const float& tBest, Leaf leaf) {
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    int sref[PT_BVH_STACK];
    float stn[PT_BVH_STACK];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const PtBvhF4 a = pt_bvh_load4(nodes, 4 * cur), b = pt_bvh_load4(nodes, 4 * cur + 1),
                          c = pt_bvh_load4(nodes, 4 * cur + 2), d = pt_bvh_load4(nodes, 4 * cur + 3);
            float n0, f0, n1, f1;
            pt_bvh_slab(a.x, a.y, a.z, a.w, b.x, b.y, ox, oy, oz, ix, iy, iz, n0, f0);
            pt_bvh_slab(b.z, b.w, c.x, c.y, c.z, c.w, ox, oy, oz, ix, iy, iz, n1, f1);
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
}

#endif /* PT_BVH_H */
