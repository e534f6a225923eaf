/* pt_driver_v2m.cuh -- driver v2m (PT_SCHED 8) of the megakernel: included by pt_kernel.cuh, inside its namespace, only
 * when that driver is asked for (pt_set_option "sched" = 8) and the scene has SDFs.  Measured slower than v2s in both
 * rounds (2.73 vs 3.87 Gsamples/s on cfg5 -- its pool code costs the instruction cache more than its
 * fuller phases save); kept as the documented, bit-exact experiment, out of the way of the drivers that ship as defaults. */
#ifndef PT_DRIVER_V2M_CUH
#define PT_DRIVER_V2M_CUH

/* ---- driver v2m (PT_SCHED 8): v2s + a per-warp pool of parked paths ---------------------------------------------------
 * In v2s a lane whose ray must march waits in the SDF state until the feeders run dry, and the SDF phase then serves
 * whoever waits: 11-15 of 32 lanes on the fractal scenes, for 35-42 % of all executed instructions.  While a lane is
 * tied to the path it holds, utilisation is conserved: filling the SDF phase by holding it back empties the feeders by
 * the same amount (profiles/r01_sdfsched).  The way out is more paths than lanes:
 *   * a ray that enters an SDF bounding box is PARKED: the lane writes the whole path (PT_POOL_FIELDS words) to a free
 *     slot of the warp's pool in shared memory and is free again -- it takes the next item or a finished parked path;
 *   * the SDF phase marches parked rays with ALL 32 lanes, whatever paths those lanes hold in registers (their own
 *     PathState is not touched): a lane loads the march job of a slot (origin, direction, SphereTracing's locals),
 *     advances it PT_SDF_REPS evaluations and writes it back, or -- finished -- writes the hit and marks the slot READY;
 *   * in the NEW phase free lanes pick READY paths up (before new items) and carry them on: PhaseTrivial, then SHADE.
 * The phase runs when PT_POOL_MIN rays wait, or when the feeders have nothing to do.  With no free slot the ray stays in
 * its lane and marches from there in the same phase (the pool cannot deadlock).  All bookkeeping is warp-uniform masks
 * and ballots; no atomics, no fences beyond __syncwarp.  A path is the same arithmetic whichever lane holds it: strict
 * mode stays bit-exact. */
#ifndef PT_POOL_CAP
#define PT_POOL_CAP 32 /* slots per warp (<= 32: one mask word) */
#endif
#ifndef PT_POOL_MIN
#define PT_POOL_MIN 24 /* marching rays (parked + in-lane) from which the SDF phase runs ahead of the feeders */
#endif
static_assert(PT_POOL_CAP >= 1 && PT_POOL_CAP <= 32, "PT_POOL_CAP: one 32-bit mask");
#define PT_POOL_FIELDS 50
#define PT_POOL_WORDS (PT_POOL_FIELDS * PT_POOL_CAP) /* per warp, [field][slot]: lanes of one access hit distinct banks */
/* field numbers: the march job first (what the SDF phase reads / writes), then the rest of the path */
enum { PF_OX = 0, PF_OY, PF_OZ, PF_DX, PF_DY, PF_DZ, PF_SX, PF_SY, PF_SZ, PF_FLAGS, PF_HT, PF_HOBJ, PF_MT, PF_INST, PF_OMEGA,
       PF_PREV, PF_TMAX, PF_KSIGN, PF_PROBE, PF_N0, PF_N1, PF_N2, PF_MPACK, PF_SET1, PF_HNX, PF_HNY, PF_HNZ, PF_HMAT,
       PF_HLIGHT, PF_LX, PF_LY, PF_LZ, PF_LW, PF_RX, PF_RY, PF_RZ, PF_RW, PF_TX, PF_TY, PF_TZ, PF_TW, PF_MIS, PF_SEED,
       PF_CX, PF_CY, PF_CZ, PF_CW, PF_SHOBJ, PF_ITEM, PF_SPARE };
static_assert(PF_SPARE + 1 == PT_POOL_FIELDS, "pool layout");
#define PT_PF(f) e[(f) * PT_POOL_CAP]

PT_DEV unsigned PoolPackMarch(const MarchState& ms) { return (unsigned)ms.points | ((unsigned)ms.iter << 12) | ((unsigned)ms.sub << 24); }
PT_DEV void PoolUnpackMarch(unsigned mk, MarchState& ms) { ms.points = (int)(mk & 0xfffu); ms.iter = (int)((mk >> 12) & 0xfffu); ms.sub = (int)(mk >> 24); }
PT_DEV void PoolStoreMarch(float* e, const MarchState& ms) {
    PT_PF(PF_MT) = ms.mt; PT_PF(PF_INST) = ms.insT; PT_PF(PF_OMEGA) = ms.omega; PT_PF(PF_PREV) = ms.previousRadius;
    PT_PF(PF_TMAX) = ms.tMax; PT_PF(PF_KSIGN) = ms.ksign; PT_PF(PF_PROBE) = ms.probe; PT_PF(PF_N0) = ms.nrm0;
    PT_PF(PF_N1) = ms.nrm1; PT_PF(PF_N2) = ms.nrm2; PT_PF(PF_MPACK) = __uint_as_float(PoolPackMarch(ms));
    PT_PF(PF_SET1) = __uint_as_float(ms.set1.w[0]);
}
PT_DEV void PoolLoadMarch(const float* e, MarchState& ms) {
    ms.mt = PT_PF(PF_MT); ms.insT = PT_PF(PF_INST); ms.omega = PT_PF(PF_OMEGA); ms.previousRadius = PT_PF(PF_PREV);
    ms.tMax = PT_PF(PF_TMAX); ms.ksign = PT_PF(PF_KSIGN); ms.probe = PT_PF(PF_PROBE); ms.nrm0 = PT_PF(PF_N0);
    ms.nrm1 = PT_PF(PF_N1); ms.nrm2 = PT_PF(PF_N2); PoolUnpackMarch(__float_as_uint(PT_PF(PF_MPACK)), ms);
    ms.set1.w[0] = __float_as_uint(PT_PF(PF_SET1));
}
/* the whole path into a slot (its ray is about to march: ps.h holds the analytic hit, ms SphereTracing's prologue) */
PT_DEV void PoolPark(float* e, const PathState& ps, const MarchState& ms, int item) {
    PT_PF(PF_OX) = ps.ray.origin.x; PT_PF(PF_OY) = ps.ray.origin.y; PT_PF(PF_OZ) = ps.ray.origin.z;
    PT_PF(PF_DX) = ps.ray.dir.x; PT_PF(PF_DY) = ps.ray.dir.y; PT_PF(PF_DZ) = ps.ray.dir.z;
    PT_PF(PF_SX) = ps.traceDir.x; PT_PF(PF_SY) = ps.traceDir.y; PT_PF(PF_SZ) = ps.traceDir.z; /* the marching ray's direction */
    PT_PF(PF_FLAGS) = __uint_as_float((unsigned)ps.bounce | ((unsigned)ps.inside << 29) | ((unsigned)ps.isShadow << 30) | ((unsigned)ps.pathAlive << 31));
    PT_PF(PF_HT) = ps.h.t; PT_PF(PF_HOBJ) = __int_as_float(ps.h.objectID);
    PoolStoreMarch(e, ms);
    PT_PF(PF_HNX) = ps.h.normal.x; PT_PF(PF_HNY) = ps.h.normal.y; PT_PF(PF_HNZ) = ps.h.normal.z;
    PT_PF(PF_HMAT) = ps.h.materialID; PT_PF(PF_HLIGHT) = ps.h.lightID;
    PT_PF(PF_LX) = ps.l.x; PT_PF(PF_LY) = ps.l.y; PT_PF(PF_LZ) = ps.l.z; PT_PF(PF_LW) = ps.l.w;
    PT_PF(PF_RX) = ps.radiance.x; PT_PF(PF_RY) = ps.radiance.y; PT_PF(PF_RZ) = ps.radiance.z; PT_PF(PF_RW) = ps.radiance.w;
    PT_PF(PF_TX) = ps.rayradiance.x; PT_PF(PF_TY) = ps.rayradiance.y; PT_PF(PF_TZ) = ps.rayradiance.z; PT_PF(PF_TW) = ps.rayradiance.w;
    PT_PF(PF_MIS) = ps.MISBRDFWeight; PT_PF(PF_SEED) = __uint_as_float(ps.seed);
    PT_PF(PF_CX) = ps.shContrib.x; PT_PF(PF_CY) = ps.shContrib.y; PT_PF(PF_CZ) = ps.shContrib.z; PT_PF(PF_CW) = ps.shContrib.w;
    PT_PF(PF_SHOBJ) = __int_as_float(ps.shObj); PT_PF(PF_ITEM) = __int_as_float(item);
}
/* a READY path back into a lane: everything but the march state (the march is over) */
PT_DEV void PoolPickup(const float* e, PathState& ps, int& item) {
    ps.ray.origin = mk3(PT_PF(PF_OX), PT_PF(PF_OY), PT_PF(PF_OZ));
    ps.ray.dir = mk3(PT_PF(PF_DX), PT_PF(PF_DY), PT_PF(PF_DZ));
    ps.shDir = mk3(PT_PF(PF_SX), PT_PF(PF_SY), PT_PF(PF_SZ));
    ps.traceDir = ps.shDir; /* (a READY path goes through PhaseTrivial / SHADE, which set it again, before anything reads it) */
    const unsigned pk = __float_as_uint(PT_PF(PF_FLAGS));
    ps.bounce = (int)(pk & 0x1fffffffu); ps.inside = PT_EXT_BSDF ? (((pk >> 29) & 1u) != 0u) : false; ps.isShadow = ((pk >> 30) & 1u) != 0u; ps.pathAlive = (pk >> 31) != 0u;
    ps.pendingFinish = false;
    ps.h.t = PT_PF(PF_HT); ps.h.objectID = __float_as_int(PT_PF(PF_HOBJ));
    ps.h.normal = mk3(PT_PF(PF_HNX), PT_PF(PF_HNY), PT_PF(PF_HNZ));
    ps.h.materialID = PT_PF(PF_HMAT); ps.h.lightID = PT_PF(PF_HLIGHT);
    ps.l = mk4(PT_PF(PF_LX), PT_PF(PF_LY), PT_PF(PF_LZ), PT_PF(PF_LW));
    ps.radiance = mk4(PT_PF(PF_RX), PT_PF(PF_RY), PT_PF(PF_RZ), PT_PF(PF_RW));
    ps.rayradiance = mk4(PT_PF(PF_TX), PT_PF(PF_TY), PT_PF(PF_TZ), PT_PF(PF_TW));
    ps.MISBRDFWeight = PT_PF(PF_MIS); ps.seed = __float_as_uint(PT_PF(PF_SEED));
    ps.shContrib = mk4(PT_PF(PF_CX), PT_PF(PF_CY), PT_PF(PF_CZ), PT_PF(PF_CW));
    ps.shObj = __float_as_int(PT_PF(PF_SHOBJ)); item = __float_as_int(PT_PF(PF_ITEM));
}

__device__ __forceinline__ void pt_render_body_v2m(const PtDevScene& sc, const PtDevParams& pr, const float* __restrict__ ubo,
                                                   float4* __restrict__ image, float* s_tab, float* s_colAll) {
    for (int i = threadIdx.x; i < PT_SH_FLOATS; i += PT_BLOCK_THREADS) s_tab[i] = __ldg(ubo + PT_SH_BASE + i);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tileX = blockIdx.x * 16 + (warp & 1) * 8, tileY = ((int)blockIdx.y + pr.blockY0) * 8 + (warp >> 1) * 4;
    const int gx = tileX + (lane & 7), gy = tileY + (lane >> 3);
    const bool inRange = (gx < pr.width) && (gy < pr.height);
    float* s_col = s_colAll + warp * (PT_STEAL_WORDS + PT_POOL_WORDS);
    float* s_pool = s_col + PT_STEAL_WORDS;
    const unsigned capMask = (PT_POOL_CAP == 32) ? 0xffffffffu : ((1u << (PT_POOL_CAP & 31)) - 1u);
    const unsigned below = (1u << lane) - 1u;

    Ctx c;
    c.sc = &sc; c.pr = &pr; c.ubo = ubo; c.s_tab = s_tab;

    const int spf = pr.samplesPerFrame;
    const bool warpLive = (tileX < pr.width) && (tileY < pr.height) && (spf > 0); /* warp-uniform */
    int roundBase = 0;
#if PT_STEAL_S == 0
    int roundN = spf;
    s_col[lane] = 0.0f; s_col[32 + lane] = 0.0f; s_col[64 + lane] = 0.0f;
    __syncwarp();
#else
    int roundN = spf < PT_STEAL_S ? spf : PT_STEAL_S;
#endif
    int next = 0, item = 0;
    unsigned mMarch = 0u, mReady = 0u; /* slots holding an unfinished march / a finished one waiting for a lane (warp-uniform) */

    int st = warpLive ? PT_ST_NEW : PT_ST_DONE;
    V3 outColor = mk3(0.0f, 0.0f, 0.0f);
    PathState ps;
    PathStateInit(ps);
    MarchState ms;
    MarchStateInit(ms);

    for (;;) {
        if (st == PT_ST_IDLE && mReady != 0u) st = PT_ST_NEW; /* a parked path finished: somebody has to carry it on */
        const unsigned bNew = __ballot_sync(0xffffffffu, st == PT_ST_NEW);
        const unsigned bIs = __ballot_sync(0xffffffffu, st == PT_ST_ISECT);
        const unsigned bSh = __ballot_sync(0xffffffffu, st == PT_ST_SHADE);
        const unsigned bSdf = __ballot_sync(0xffffffffu, st == PT_ST_SDF); /* rays marching in their lane (pool was full) */
        if ((bNew | bIs | bSdf | bSh | mMarch) == 0u) { /* (mReady != 0 implies bNew != 0) */
            if (!warpLive) break;
            if (!PoolCloseRound(s_col, lane, inRange, spf, roundBase, roundN, outColor)) break;
            next = 0;
            st = PT_ST_NEW;
            continue;
        }
        int phase = PT_ST_NEW, best = __popc(bNew);
        if (__popc(bIs) >= best) { best = __popc(bIs); phase = PT_ST_ISECT; }
        if (__popc(bSh) >= best) { best = __popc(bSh); phase = PT_ST_SHADE; }
        const int marching = __popc(mMarch) + __popc(bSdf);
        if (marching > 0 && (marching >= PT_POOL_MIN || best == 0 || (bSdf != 0u && best < PT_FEED_T))) phase = PT_ST_SDF;
        PT_STAT(phase, phase == PT_ST_NEW ? bNew : (phase == PT_ST_ISECT ? bIs : (phase == PT_ST_SDF ? 0u : bSh)));
        if (phase == PT_ST_NEW) {
            const int rank = __popc(bNew & below), nReady = __popc(mReady);
            const int take = (__popc(bNew) < nReady) ? __popc(bNew) : nReady; /* READY paths picked up in this execution */
            unsigned taken = 0u;
            if (st == PT_ST_NEW) {
                if (ps.pendingFinish) {
                    PoolDeposit(s_col, item, PathColor(c, ps));
                    ps.pendingFinish = false;
                }
                if (rank < take) { /* carry a finished parked path on */
                    const int slot = (int)__fns(mReady, 0u, rank + 1);
                    PoolPickup(s_pool + slot, ps, item);
                    taken = 1u << slot;
                    st = PhaseTrivial(ps);
                } else {
                    item = next + (rank - take);
                    if (item < 32 * roundN) {
                        const int q = item & 31;
                        const int qx = tileX + (q & 7), qy = tileY + (q >> 3);
                        if ((qx < pr.width) && (qy < pr.height)) /* else: a pixel beyond the image edge; claim again */
                            st = PhaseNew(c, ps, (unsigned)qx, (unsigned)pr.height - (unsigned)qy, roundBase + (item >> 5));
                    } else {
                        st = PT_ST_IDLE;
                    }
                }
            }
            next += __popc(bNew) - take;
            if (take > 0) {
                mReady &= ~__reduce_or_sync(0xffffffffu, taken);
                __syncwarp();
            }
        } else if (phase == PT_ST_ISECT) {
            bool entered = false;
            if (st == PT_ST_ISECT) {
                st = PhaseIsect(c, ps, ms);
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
                entered = (st == PT_ST_SDF);
            }
            const unsigned bEnt = __ballot_sync(0xffffffffu, entered);
            if (bEnt != 0u) { /* park the rays that have to march, as many as there are free slots */
                const unsigned mFree = ~(mMarch | mReady) & capMask;
                const int r = __popc(bEnt & below), nFree = __popc(mFree);
                unsigned parkedBit = 0u;
                if (entered && r < nFree) {
                    const int slot = (int)__fns(mFree, 0u, r + 1);
                    PoolPark(s_pool + slot, ps, ms, item);
                    parkedBit = 1u << slot;
                    st = PT_ST_NEW; /* nothing pending: the NEW phase hands this lane a READY path or a new item */
                }
                mMarch |= __reduce_or_sync(0xffffffffu, parkedBit);
                __syncwarp();
            }
        } else if (phase == PT_ST_SDF) {
            /* who marches what: lanes with a ray of their own march that; the others take the parked rays in slot order */
            const bool own = (st == PT_ST_SDF);
            const int r = __popc(~bSdf & below);
            const bool job = !own && (r < __popc(mMarch));
            PT_STAT_SDF(__popc(bSdf) + min(32 - __popc(bSdf), __popc(mMarch)));
#ifdef PT_STATS
            if ((threadIdx.x & 31) == 0) atomicAdd(&pt_stats[2 * PT_ST_SDF + 1], (unsigned long long)(__popc(bSdf) + min(32 - __popc(bSdf), __popc(mMarch))));
#endif
            float* e = s_pool + (job ? (int)__fns(mMarch, 0u, r + 1) : 0);
            PathState js; /* the marching ray: only ray.origin, the direction, isShadow and h are read / written */
            MarchState jm = ms;
            js.ray.origin = ps.ray.origin;
            js.ray.dir = ps.traceDir;
            js.isShadow = ps.isShadow;
            js.h = ps.h;
            if (job) {
                js.ray.origin = mk3(PT_PF(PF_OX), PT_PF(PF_OY), PT_PF(PF_OZ));
                js.isShadow = ((__float_as_uint(PT_PF(PF_FLAGS)) >> 30) & 1u) != 0u;
                js.ray.dir = mk3(PT_PF(PF_SX), PT_PF(PF_SY), PT_PF(PF_SZ)); /* the traced direction, shadow or path ray alike */
                js.h.t = PT_PF(PF_HT);
                js.h.objectID = __float_as_int(PT_PF(PF_HOBJ));
                PoolLoadMarch(e, jm);
            }
            js.shDir = js.ray.dir;
            js.traceDir = js.ray.dir;
            int jst = (own || job) ? PT_ST_SDF : PT_ST_DONE;
#pragma unroll 1
            for (int rep = 0; rep < PT_SDF_REPS; rep++) {
                if (jst == PT_ST_SDF) jst = PhaseSdfEval(c, js, jm);
            }
            unsigned doneBit = 0u;
            if (own) {
                ps.h = js.h;
                ms = jm;
                st = jst;
                if (st == PT_ST_SHADE) st = PhaseTrivial(ps);
            } else if (job) {
                /* (a converged path ray has its hit distance before its normal: h.t travels with an unfinished job too) */
                PT_PF(PF_HT) = js.h.t; PT_PF(PF_HOBJ) = __int_as_float(js.h.objectID);
                if (jst == PT_ST_SDF) {
                    PoolStoreMarch(e, jm);
                } else { /* finished: the hit goes to the slot, the path waits there for a free lane */
                    if (jm.sub == PT_SUB_N0 + 6) { /* a path ray that hit the SDF: normal and material came with it */
                        PT_PF(PF_HNX) = js.h.normal.x; PT_PF(PF_HNY) = js.h.normal.y; PT_PF(PF_HNZ) = js.h.normal.z;
                        PT_PF(PF_HMAT) = js.h.materialID; PT_PF(PF_HLIGHT) = js.h.lightID;
                    }
                    doneBit = 1u << (unsigned)(e - s_pool);
                }
            }
            if (mMarch != 0u) {
                doneBit = __reduce_or_sync(0xffffffffu, doneBit);
                mMarch &= ~doneBit;
                mReady |= doneBit;
                __syncwarp();
            }
        } else {
            if (st == PT_ST_SHADE) st = PhaseShadeHit(c, ps);
        }
    }
    if (inRange) StoreTexel(pr, image, gx, gy, outColor);
}
#undef PT_PF

#endif /* PT_DRIVER_V2M_CUH */
