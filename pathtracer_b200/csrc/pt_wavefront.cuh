/* pt_wavefront.cuh -- the wavefront pipeline: the same four phases as the megakernel's v2 driver (pt_kernel.cuh:
 * PhaseNew / PhaseIsect / PhaseSdfEval / PhaseShade), one kernel per phase, path state in HBM.
 *
 * Why: inside a warp a path tracer diverges -- paths end at different bounces, only some rays enter an SDF bounding
 * box, marches take 1..512 steps.  Here every kernel's threads each take ONE item of a queue, so all lanes of a
 * warp do the same phase; the march kernel refills a lane from the queue the moment its ray converges.  The price
 * is HBM traffic for the state (SoA, one float4 per field group so a warp reads 512 contiguous bytes per field)
 * and more launches; pt_kernel.cuh's megakernel pays neither.  profiles/ holds the measured comparison.
 *
 * Layout (P paths in flight = pixels x samples of the chunk; path p belongs to pixel pixBase + p % nPix of the frame and
 * to sample chunkBase + p / nPix; a chunk is a band of nPix consecutive pixels x some samples, sized by wf_max_paths):
 *   rayO [P] float4  ray origin xyz, MISBRDFWeight         rayD  [P] float4  ray dir xyz (next path direction)
 *   wl   [P] float4  the wavelength bundle                 rad   [P] float4  radiance
 *   thr  [P] float4  rayradiance (throughput)              shD   [P] float4  shadow dir xyz, shObj (int bits)
 *   shC  [P] float4  pending light contribution            hit0  [P] float4  t, normal xyz
 *   hit1 [P] float4  materialID, lightID, objectID bits    misc  [P] uint2   seed, bounce | inside<<29 | isShadow<<30 | pathAlive<<31
 *   col  [P] float4  XYZ of the finished path              acc   [nPix] float4 per-pixel sum, added in sample order
 * Queues hold path indices; pushes are warp-aggregated (ballot + popc prefix, one atomicAdd per warp).
 * Per bounce depth:  ISECT(qA) -> MARCH(qM) -> SHADE(qA: -> qS shadow rays, qB next depth, or finished)
 *                    ISECT(qS) -> MARCH(qM) -> SHADE(qS: verdict -> qB or finished)      then qA <-> qB.
 * The arithmetic per path is exactly the megakernel's, and pixels are summed in sample-index order by the FINAL
 * kernel, so the strict build stays bit-exact against the oracle.
 */
#ifndef PT_WAVEFRONT_CUH
#define PT_WAVEFRONT_CUH

#include "pt_kernel.cuh"

#ifndef PT_WF_REFILL
#define PT_WF_REFILL 8 /* the march kernel refills once this many lanes of a warp are idle */
#endif

enum { PT_WF_NA = 0, PT_WF_NB = 1, PT_WF_NS = 2, PT_WF_NM = 3, PT_WF_HEAD = 4 };

namespace PT_KERNEL_NS {

PT_DEV unsigned LaneId() { return threadIdx.x & 31u; }

/* all 32 lanes of the warp must call this together */
PT_DEV void QueuePush(unsigned* q, unsigned* counter, bool pred, unsigned value) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    unsigned base = 0u;
    if ((int)LaneId() == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) q[base + (unsigned)__popc(m & ((1u << LaneId()) - 1u))] = value;
}

PT_DEV void PixelOf(const PtDevParams& pr, unsigned pix, unsigned& xyx, unsigned& xyy) {
    const unsigned gx = pix % (unsigned)pr.width, gy = pix / (unsigned)pr.width;
    xyx = gx;
    xyy = (unsigned)pr.height - gy; /* shader.comp:1510 */
}

PT_DEV void StageTable(const float* __restrict__ ubo, float* s_tab) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += blockDim.x) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();
}

PT_DEV unsigned PackMisc(const PathState& ps) {
    return ((unsigned)ps.bounce & 0x1fffffffu) | (ps.inside ? 0x20000000u : 0u) | (ps.isShadow ? 0x40000000u : 0u) |
           (ps.pathAlive ? 0x80000000u : 0u);
}
PT_DEV void UnpackMisc(unsigned v, PathState& ps) {
    ps.bounce = (int)(v & 0x1fffffffu);
    ps.inside = PT_EXT_BSDF ? ((v & 0x20000000u) != 0u) : false; /* only the surface extensions ever set it */
    ps.isShadow = (v & 0x40000000u) != 0u;
    ps.pathAlive = (v & 0x80000000u) != 0u;
}

/* the ray of path p as the intersection phases need it */
PT_DEV void LoadRay(const PtWf& w, unsigned p, PathState& ps) {
    const uint2 m = w.misc[p];
    UnpackMisc(m.y, ps);
    const float4 o = w.rayO[p];
    ps.ray.origin = mk3(o.x, o.y, o.z);
    if (ps.isShadow) {
        const float4 d = w.shD[p];
        ps.shDir = mk3(d.x, d.y, d.z);
        ps.traceDir = ps.shDir;
    } else {
        const float4 d = w.rayD[p];
        ps.ray.dir = mk3(d.x, d.y, d.z);
        ps.traceDir = ps.ray.dir;
    }
}
PT_DEV void StoreHit(const PtWf& w, unsigned p, const Hit& h) {
    w.hit0[p] = make_float4(h.t, h.normal.x, h.normal.y, h.normal.z);
    w.hit1[p] = make_float4(h.materialID, h.lightID, __int_as_float(h.objectID), 0.0f);
}
PT_DEV void LoadHit(const PtWf& w, unsigned p, Hit& h) {
    const float4 a = w.hit0[p], b = w.hit1[p];
    h.t = a.x; h.normal = mk3(a.y, a.z, a.w);
    h.materialID = b.x; h.lightID = b.y; h.objectID = __float_as_int(b.z);
}

/* ---- GEN: PhaseNew for every path of the chunk ------------------------------------------------------------------ */
PT_DEV void WfGen(const Ctx& c, const PtWf& w) {
    const PtDevParams& pr = *c.pr;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < w.P; p += gridDim.x * blockDim.x) {
        const unsigned pix = w.pixBase + p % w.nPix, s = p / w.nPix;
        unsigned xyx, xyy;
        PixelOf(pr, pix, xyx, xyy);
        PathState ps;
        PathStateInit(ps);
        const int st = PhaseNew(c, ps, xyx, xyy, w.chunkBase + (int)s);
        w.rayO[p] = make_float4(ps.ray.origin.x, ps.ray.origin.y, ps.ray.origin.z, ps.MISBRDFWeight);
        w.rayD[p] = make_float4(ps.ray.dir.x, ps.ray.dir.y, ps.ray.dir.z, 0.0f);
        w.wl[p] = make_float4(ps.l.x, ps.l.y, ps.l.z, ps.l.w);
        w.rad[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        w.thr[p] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        w.misc[p] = make_uint2(ps.seed, PackMisc(ps));
        if (st != PT_ST_ISECT) { /* pathLength <= 0: the path is finished before it starts */
            const V3 col = PathColor(c, ps);
            w.col[p] = make_float4(col.x, col.y, col.z, 0.0f);
        }
    }
}

/* ---- ISECT: analytic primitives for every ray of a queue; rays that enter an SDF box go to the march queue -------
 * identity != 0: the queue is 0..count-1 (depth 0, straight after GEN) */
PT_DEV void WfIsect(const Ctx& c, const PtWf& w, const unsigned* __restrict__ q, unsigned count, int identity) {
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (count + stride - 1u) / stride;
    for (unsigned r = 0; r < rounds; r++) { /* uniform trip count: QueuePush needs whole warps */
        const unsigned i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < count;
        bool toMarch = false;
        unsigned p = 0u;
        if (valid) {
            p = identity ? i : q[i];
            PathState ps;
            MarchState ms;
            LoadRay(w, p, ps);
            toMarch = (PhaseIsect(c, ps, ms) == PT_ST_SDF);
            StoreHit(w, p, ps.h);
        }
#if PT_HAS_SDF
        QueuePush(w.qM, w.cnt + PT_WF_NM, toMarch, p);
#else
        (void)toMarch;
#endif
    }
}

/* ---- MARCH: sphere tracing with per-lane refill -----------------------------------------------------------------
 * Persistent warps; a lane whose ray has converged (or left the boxes) writes its hit and takes the next ray of the
 * queue, so the single SDF() site below runs with (nearly) all lanes busy until the queue is drained. */
#if PT_HAS_SDF
PT_DEV void WfMarch(const Ctx& c, const PtWf& w) {
    const unsigned count = w.cnt[PT_WF_NM];
    PathState ps;
    MarchState ms;
    PathStateInit(ps);
    MarchStateInit(ms);
    bool busy = false;
    unsigned p = 0u;
    bool drained = false;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (!drained && (idle == 0xffffffffu || __popc(idle) >= PT_WF_REFILL)) {
            const int leader = __ffs(idle) - 1;
            unsigned base = 0u;
            if ((int)LaneId() == leader) base = atomicAdd(w.cnt + PT_WF_HEAD, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!busy) {
                const unsigned i = base + (unsigned)__popc(idle & ((1u << LaneId()) - 1u));
                if (i < count) {
                    p = w.qM[i];
                    LoadRay(w, p, ps);
                    LoadHit(w, p, ps.h);
                    /* SphereTracing's prologue again (PhaseIsect found a box for this ray, so it does here too) */
                    const V3 dir = ps.isShadow ? ps.shDir : ps.ray.dir;
                    const V3 invdir = mk3(PTK_DIV(1.0f, dir.x), PTK_DIV(1.0f, dir.y), PTK_DIV(1.0f, dir.z));
                    float tMin = 1e5f;
                    ms.tMax = 1e5f;
                    busy = SearchSDF(c, ps.ray.origin, invdir, tMin, ms.tMax, ms.set1);
                    ms.mt = PTK_MAX(tMin, 1e-3f);
                    ms.insT = 0.0f; ms.omega = 1.70f; ms.previousRadius = 0.0f; ms.points = 0; ms.iter = 0;
                    ms.sub = PT_SUB_SIGN;
                }
            }
            if (base + (unsigned)__popc(idle) >= count) drained = true; /* uniform: same base for the whole warp */
        }
        if (__ballot_sync(0xffffffffu, busy) == 0u) {
            if (drained) break;
            continue;
        }
        if (busy) {
            if (PhaseSdfEval(c, ps, ms) != PT_ST_SDF) {
                StoreHit(w, p, ps.h);
                busy = false;
            }
        }
    }
}
#endif

/* ---- SHADE: PhaseShade for every item of a queue; routes the path to the shadow queue, the next depth or the end -- */
PT_DEV void WfShade(const Ctx& c, const PtWf& w, const unsigned* __restrict__ q, unsigned count, int identity) {
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (count + stride - 1u) / stride;
    for (unsigned r = 0; r < rounds; r++) {
        const unsigned i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < count;
        bool toShadow = false, toNext = false;
        unsigned p = 0u;
        if (valid) {
            p = identity ? i : q[i];
            PathState ps;
            const uint2 m = w.misc[p];
            ps.seed = m.x;
            UnpackMisc(m.y, ps);
            const float4 o = w.rayO[p], d = w.rayD[p], l = w.wl[p], ra = w.rad[p], th = w.thr[p];
            ps.ray.origin = mk3(o.x, o.y, o.z);
            ps.MISBRDFWeight = o.w;
            ps.ray.dir = mk3(d.x, d.y, d.z);
            ps.l = mk4(l.x, l.y, l.z, l.w);
            ps.radiance = mk4(ra.x, ra.y, ra.z, ra.w);
            ps.rayradiance = mk4(th.x, th.y, th.z, th.w);
            ps.pendingFinish = false;
            ps.shDir = mk3(0.0f, 0.0f, 0.0f);
            ps.shContrib = mk4(0.0f, 0.0f, 0.0f, 0.0f);
            ps.shObj = 0;
            if (ps.isShadow) {
                const float4 sd = w.shD[p], sc4 = w.shC[p];
                ps.shObj = __float_as_int(sd.w);
                ps.shContrib = mk4(sc4.x, sc4.y, sc4.z, sc4.w);
            }
            LoadHit(w, p, ps.h);
            const bool wasShadow = ps.isShadow;
            const int next = PhaseShade(c, ps);
            w.rad[p] = make_float4(ps.radiance.x, ps.radiance.y, ps.radiance.z, ps.radiance.w);
            if (next == PT_ST_NEW) {
                const V3 col = PathColor(c, ps);
                w.col[p] = make_float4(col.x, col.y, col.z, 0.0f);
            } else {
                w.misc[p] = make_uint2(ps.seed, PackMisc(ps));
                if (!wasShadow) { /* a verdict changes nothing but radiance and the flags */
                    w.rayO[p] = make_float4(ps.ray.origin.x, ps.ray.origin.y, ps.ray.origin.z, ps.MISBRDFWeight);
                    w.rayD[p] = make_float4(ps.ray.dir.x, ps.ray.dir.y, ps.ray.dir.z, 0.0f);
                    w.thr[p] = make_float4(ps.rayradiance.x, ps.rayradiance.y, ps.rayradiance.z, ps.rayradiance.w);
                    if (ps.isShadow) {
                        w.shD[p] = make_float4(ps.shDir.x, ps.shDir.y, ps.shDir.z, __int_as_float(ps.shObj));
                        w.shC[p] = make_float4(ps.shContrib.x, ps.shContrib.y, ps.shContrib.z, ps.shContrib.w);
                    }
                }
                toShadow = ps.isShadow;
                toNext = !ps.isShadow;
            }
        }
        QueuePush(w.qS, w.cnt + PT_WF_NS, toShadow, p);
        QueuePush(w.qB, w.cnt + PT_WF_NB, toNext, p);
    }
}

/* ---- SORT (surface extensions only, option wf_sort): the path rays about to be shaded, grouped by the lobe of the surface
 * they hit -- 0 = the reference's Lambertian, misses and emitters, 1 mirror, 2 glossy, 3 dielectric -- so a warp of the SHADE
 * kernel runs one lobe's code instead of up to four (the "material-sorted shading" of the wavefront design; with the
 * reference's single surface model there is nothing to sort).  A counting sort in two passes over the queue: pass 0 counts
 * the classes (one atomic per warp and class), pt_wf_ctl(3) turns the counts into bases, pass 1 scatters into the march
 * queue's buffer, idle at this point.  The order inside a class follows the atomics; nothing depends on it (FINAL sums by
 * sample index). */
enum { PT_WF_BIN0 = 8, PT_WF_BASE0 = 12 };
PT_DEV int WfLobeClass(const Ctx& c, const PtWf& w, unsigned p) {
#if PT_EXT_BSDF
    const float4 a = w.hit0[p], b = w.hit1[p];
    if (!(a.x < 1e5f)) return 0;
    float emitT, emitL;
    GetLightMix(c, b.y, emitT, emitL);
    if (emitL > 0.0f) return 0;
    return SurfaceExtOf(c, b.x).bsdf & 3;
#else
    (void)c; (void)w; (void)p;
    return 0;
#endif
}
PT_DEV void WfSort(const Ctx& c, const PtWf& w, const unsigned* __restrict__ q, unsigned count, int identity, int pass) {
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (count + stride - 1u) / stride;
    for (unsigned r = 0; r < rounds; r++) {
        const unsigned i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < count;
        const unsigned p = valid ? (identity ? i : q[i]) : 0u;
        const int cls = valid ? WfLobeClass(c, w, p) : -1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (pass == 0) {
                const unsigned m = __ballot_sync(0xffffffffu, cls == k);
                if (m != 0u && LaneId() == (unsigned)(__ffs(m) - 1)) atomicAdd(w.cnt + PT_WF_BIN0 + k, (unsigned)__popc(m));
            } else {
                QueuePush(w.qM + w.cnt[PT_WF_BASE0 + k], w.cnt + PT_WF_BIN0 + k, cls == k, p);
            }
        }
    }
}

/* ---- FINAL: per pixel, add the chunk's finished paths in sample-index order; after the last chunk, Rendering()'s
 * tail + Accumulate() + imageStore (StoreTexel) ------------------------------------------------------------------- */
PT_DEV void WfFinal(const PtDevParams& pr, const PtWf& w, float4* __restrict__ image) {
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < w.nPix; pix += gridDim.x * blockDim.x) {
        V3 outColor = mk3(0.0f, 0.0f, 0.0f);
        if (w.chunkBase != 0) {
            const float4 a = w.acc[pix];
            outColor = mk3(a.x, a.y, a.z);
        }
        for (int s = 0; s < w.chunkSamples; s++) {
            const float4 cc = w.col[(size_t)s * w.nPix + pix];
            outColor = outColor + mk3(cc.x, cc.y, cc.z);
        }
        if (w.lastChunk) {
            const unsigned g = w.pixBase + pix; /* pix counts inside the chunk's band of the frame */
            StoreTexel(pr, image, (int)(g % (unsigned)pr.width), (int)(g / (unsigned)pr.width), outColor);
        } else {
            w.acc[pix] = make_float4(outColor.x, outColor.y, outColor.z, 0.0f);
        }
    }
}

} /* namespace PT_KERNEL_NS */

#define PT_WF_CTX()                                                  \
    __shared__ float s_tab[PT_SH_FLOATS];                            \
    PT_KERNEL_NS::StageTable(ubo, s_tab);                            \
    PT_KERNEL_NS::Ctx c;                                             \
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

#define PT_WF_SIG const __grid_constant__ PtDevScene sc, const __grid_constant__ PtDevParams pr, const float* __restrict__ ubo, const __grid_constant__ PtWf w

#if PT_HAS_SDF
#define PT_DEFINE_WF_MARCH                                                                                     \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS) pt_wf_march(PT_WF_SIG) {     \
        PT_WF_CTX();                                                                                           \
        PT_KERNEL_NS::WfMarch(c, w);                                                                           \
    }
#else
#define PT_DEFINE_WF_MARCH
#endif

/* which: 0 = queue A (identity when identity != 0), 1 = queue S */
#define PT_DEFINE_WAVEFRONT_KERNELS                                                                            \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS) pt_wf_gen(PT_WF_SIG) {       \
        PT_WF_CTX();                                                                                           \
        PT_KERNEL_NS::WfGen(c, w);                                                                             \
    }                                                                                                          \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                              \
    pt_wf_isect(PT_WF_SIG, int which, int identity) {                                                          \
        PT_WF_CTX();                                                                                           \
        const unsigned n = identity ? w.P : w.cnt[which ? PT_WF_NS : PT_WF_NA];                                \
        PT_KERNEL_NS::WfIsect(c, w, which ? w.qS : w.qA, n, identity);                                         \
    }                                                                                                          \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                              \
    pt_wf_shade(PT_WF_SIG, int which, int identity) {                                                          \
        PT_WF_CTX();                                                                                           \
        const unsigned n = identity ? w.P : w.cnt[which == 1 ? PT_WF_NS : PT_WF_NA];                           \
        PT_KERNEL_NS::WfShade(c, w, which == 2 ? w.qM : (which ? w.qS : w.qA), n, which == 2 ? 0 : identity);  \
    }                                                                                                          \
    extern "C" __global__ void __launch_bounds__(PT_BLOCK_THREADS, PT_MIN_BLOCKS)                              \
    pt_wf_sort(PT_WF_SIG, int identity, int pass) {                                                            \
        PT_WF_CTX();                                                                                           \
        const unsigned n = identity ? w.P : w.cnt[PT_WF_NA];                                                   \
        PT_KERNEL_NS::WfSort(c, w, w.qA, n, identity, pass);                                                   \
    }                                                                                                          \
    extern "C" __global__ void pt_wf_ctl(const __grid_constant__ PtWf w, int op) {                             \
        if (threadIdx.x != 0 || blockIdx.x != 0) return;                                                       \
        if (op == 0) { for (int i = 0; i < 16; i++) w.cnt[i] = 0u; }             /* start of a chunk */       \
        else if (op == 3) { unsigned b = 0u;                                      /* SORT: counts -> bases */  \
            for (int k = 0; k < 4; k++) { w.cnt[12 + k] = b; b += w.cnt[8 + k]; w.cnt[8 + k] = 0u; } }        \
        else if (op == 4) { for (int k = 0; k < 4; k++) w.cnt[8 + k] = 0u; }     /* after SORT's scatter */   \
        else if (op == 1) { w.cnt[PT_WF_NM] = 0u; w.cnt[PT_WF_HEAD] = 0u; }      /* before the next march */   \
        else { w.cnt[PT_WF_NA] = w.cnt[PT_WF_NB]; w.cnt[PT_WF_NB] = 0u; w.cnt[PT_WF_NS] = 0u;                  \
               w.cnt[PT_WF_NM] = 0u; w.cnt[PT_WF_HEAD] = 0u; }                   /* next depth (host swaps qA/qB) */ \
    }                                                                                                          \
    extern "C" __global__ void pt_wf_final(PT_WF_SIG, float4* __restrict__ image) {                            \
        (void)sc; (void)ubo;                                                                                   \
        PT_KERNEL_NS::WfFinal(pr, w, image);                                                                   \
    }                                                                                                          \
    PT_DEFINE_WF_MARCH

#endif /* PT_WAVEFRONT_CUH */
