/* pt_prepare.cpp -- host-side evaluation of the per-scene and per-dispatch constants the kernels consume.
 *
 * Everything here is arithmetic the reference's shader performs per ray on values that only depend on the
 * uniform block / push constants (see pt_dev_scene.h).  It is evaluated with pt_math.h and in the shader's fp32
 * operation order, and this file MUST be compiled with -ffp-contract=off (csrc/Makefile does), so the results are
 * bit-identical to the per-ray evaluation the strict kernels would otherwise do.
 */
#include "pt_internal.h"

#include <string.h>

namespace {

inline int f2i(float v) { return pt_f2i(v); }

struct Flat {
    const float* u;
    float at(int base, int i) const { /* clamped flat indexing: same rule as the kernels and the oracle */
        long k = (long)base + (long)i;
        if (k < 0) k = 0;
        if (k > PT_UBO_FLOATS - 1) k = PT_UBO_FLOATS - 1;
        return u[k];
    }
    float obj(int i) const { return at(PT_OFF_OBJ, i); }
};

/* RotationMatrix (shader.comp:143-152): mX * mY * mZ with GLSL's column-major constructors; out[3*col+row] */
void rotation_matrix(float ax, float ay, float az, float* out) {
    const float k = 0.0174532925199f;
    ax = ax * k; ay = ay * k; az = az * k;
    const float sx = pt_sin(ax), sy = pt_sin(ay), sz = pt_sin(az);
    const float cx = pt_cos(ax), cy = pt_cos(ay), cz = pt_cos(az);
    const float mX[9] = {1.0f, 0.0f, 0.0f, 0.0f, cx, -sx, 0.0f, sx, cx};
    const float mY[9] = {cy, 0.0f, sy, 0.0f, 1.0f, 0.0f, -sy, 0.0f, cy};
    const float mZ[9] = {cz, -sz, 0.0f, sz, cz, 0.0f, 0.0f, 0.0f, 1.0f};
    /* (A*B) column j = A.col0*B[j][0] + A.col1*B[j][1] + A.col2*B[j][2], summed left to right */
    float t[9];
    for (int j = 0; j < 3; j++)
        for (int r = 0; r < 3; r++)
            t[3 * j + r] = mX[r] * mY[3 * j] + mX[3 + r] * mY[3 * j + 1] + mX[6 + r] * mY[3 * j + 2];
    for (int j = 0; j < 3; j++)
        for (int r = 0; r < 3; r++)
            out[3 * j + r] = t[r] * mZ[3 * j] + t[3 + r] * mZ[3 * j + 1] + t[6 + r] * mZ[3 * j + 2];
}

/* the constants of LensIntersection / SphereSliceIntersection (shader.comp:366-448) for one lens */
void prepare_lens(float px, float py, float pz, float rx, float ry, float rz, float radius, float focalLength,
                  float thickness, bool isConverging, PtDevLens* L) {
    L->px = px; L->py = py; L->pz = pz;
    rotation_matrix(rx, ry, rz, L->m);
    const float lensThicknessHalf =
        2.0f * focalLength - pt_sqrt(4.0f * focalLength * focalLength - radius * radius);
    float lensSlicePos = 0.5f * (isConverging ? thickness : -thickness);
    if (isConverging) lensSlicePos += lensThicknessHalf;
    L->sradius = 2.0f * focalLength;
    L->sradius2 = L->sradius * L->sradius;
    L->sliceOffset = L->sradius - lensThicknessHalf;
    L->shift = lensSlicePos - lensThicknessHalf - L->sliceOffset;
    L->invertSide = isConverging ? 0.0f : 1.0f;
    /* shader.comp:899-905 */
    float boundingRadius;
    if (isConverging) {
        boundingRadius = (radius * radius) + (0.25f * thickness * thickness);
    } else {
        boundingRadius = 0.5f * thickness + 2.0f * focalLength -
                         pt_sqrt(4.0f * focalLength * focalLength - radius * radius);
        boundingRadius = boundingRadius * boundingRadius + radius * radius;
    }
    L->bound2 = boundingRadius;
}

}  // namespace

/* The surface extension table of pt_set_surface_ext into the scene the kernels read.  Returns whether any entry leaves
 * the reference's shading (only then is the kernel built with the extension code). */
bool pt_prepare_surface_ext(const pt_surface_ext* table, int n, PtDevScene* sc) {
    bool any = false;
    sc->nSurfaceExt = 0;
    memset(sc->surfaceExt, 0, sizeof sc->surfaceExt);
    for (int i = 0; i < n && i < PT_DEV_MAX_SURFACE_EXT; i++) {
        sc->surfaceExt[i].bsdf = table[i].bsdf;
        sc->surfaceExt[i].roughness = table[i].roughness;
        sc->surfaceExt[i].ior = table[i].ior;
        if (table[i].bsdf != PT_BSDF_REFERENCE) { any = true; sc->nSurfaceExt = i + 1; }
    }
    return any;
}

int pt_prepare_scene(const pt_ubo* ubo, PtDevScene* sc, std::string* err) {
    memset(sc, 0, sizeof *sc);
    Flat F = {reinterpret_cast<const float*>(ubo)};
    const float* num = ubo->numObjects;
    /* loop bounds are `float(i) < numObjects[k]` in the shader; offsets use int(numObjects[k]) */
    auto count = [](float n) { int c = 0; while ((float)c < n && c < 100000) c++; return c; };
    const int nS = count(num[0]), nP = count(num[1]), nB = count(num[2]), nL = count(num[3]), nC = count(num[4]);
    const int nSdf = count(num[5]);
    /* anything the reference's 1024-float objects[] can describe fits the pool; counts that overrun the arrays
     * (a hand-made ubo) are refused rather than read out of bounds */
    if (6L * nS + 5L * nP + 11L * nB + 12L * nL + 16L * nC > PT_MAX_OBJECTS_SIZE || nSdf > PT_DEV_MAX_SDFS) {
        if (err) *err = "numObjects[] describes more objects than the uniform block holds (or more than 128 SDFs)";
        return PT_ERR_ARG;
    }
    sc->nSpheres = nS; sc->nPlanes = nP; sc->nBoxes = nB; sc->nLenses = nL; sc->nCyclides = nC; sc->nSdfs = nSdf;
    sc->offPlanes = 8 * nS;
    sc->offBoxes = sc->offPlanes + 4 * nP;
    sc->offLenses = sc->offBoxes + 20 * nB;
    sc->offCyclides = sc->offLenses + 20 * nL;
    sc->offSdfs = sc->offCyclides + 24 * nC;
    if (sc->offSdfs + 8 * nSdf > PT_DEV_POOL_FLOATS) {
        if (err) *err = "device scene pool overflow";
        return PT_ERR_ARG;
    }
    PtDevSphere* spheres = reinterpret_cast<PtDevSphere*>(sc->pool);
    PtDevPlane* planes = reinterpret_cast<PtDevPlane*>(sc->pool + sc->offPlanes);
    PtDevBox* boxes = reinterpret_cast<PtDevBox*>(sc->pool + sc->offBoxes);
    PtDevLens* lenses = reinterpret_cast<PtDevLens*>(sc->pool + sc->offLenses);
    PtDevCyclide* cyclides = reinterpret_cast<PtDevCyclide*>(sc->pool + sc->offCyclides);
    PtDevSdf* sdfs = reinterpret_cast<PtDevSdf*>(sc->pool + sc->offSdfs);
    const int iS = f2i(num[0]), iP = f2i(num[1]), iB = f2i(num[2]), iL = f2i(num[3]);

    int offset = 0;
    for (int i = 0; i < nS; i++) { /* UnpackSphere shader.comp:154-161 */
        PtDevSphere& o = spheres[i];
        const int k = 6 * i;
        o.px = F.obj(k); o.py = F.obj(k + 1); o.pz = F.obj(k + 2); o.radius = F.obj(k + 3);
        o.r2 = o.radius * o.radius;
        o.materialID = (float)(f2i(F.obj(k + 4)) - 1);
        o.lightID = (float)(f2i(F.obj(k + 5)) - 1);
    }
    offset += 6 * iS;
    for (int i = 0; i < nP; i++) { /* UnpackPlane shader.comp:163-169 */
        PtDevPlane& o = planes[i];
        const int k = 5 * i + offset;
        o.py = F.obj(k + 1);
        o.materialID = (float)(f2i(F.obj(k + 3)) - 1);
        o.lightID = (float)(f2i(F.obj(k + 4)) - 1);
    }
    offset += 5 * iP;
    for (int i = 0; i < nB; i++) { /* UnpackBox shader.comp:171-179 */
        PtDevBox& o = boxes[i];
        const int k = 11 * i + offset;
        o.px = F.obj(k); o.py = F.obj(k + 1); o.pz = F.obj(k + 2);
        rotation_matrix(F.obj(k + 3), F.obj(k + 4), F.obj(k + 5), o.m);
        o.sx = F.obj(k + 6); o.sy = F.obj(k + 7); o.sz = F.obj(k + 8);
        o.bound2 = 0.25f * (o.sx * o.sx + o.sy * o.sy + o.sz * o.sz);
        o.materialID = (float)(f2i(F.obj(k + 9)) - 1);
        o.lightID = (float)(f2i(F.obj(k + 10)) - 1);
    }
    offset += 11 * iB;
    for (int i = 0; i < nL; i++) { /* UnpackLens shader.comp:181-192 */
        PtDevLens& o = lenses[i];
        const int k = 12 * i + offset;
        prepare_lens(F.obj(k), F.obj(k + 1), F.obj(k + 2), F.obj(k + 3), F.obj(k + 4), F.obj(k + 5), F.obj(k + 6),
                     F.obj(k + 7), F.obj(k + 8), F.obj(k + 9) != 0.0f, &o);
        o.materialID = (float)(f2i(F.obj(k + 10)) - 1);
        o.lightID = (float)(f2i(F.obj(k + 11)) - 1);
    }
    offset += 12 * iL;
    for (int i = 0; i < nC; i++) { /* UnpackCyclide shader.comp:194-207 */
        PtDevCyclide& o = cyclides[i];
        const int k = 16 * i + offset;
        o.px = F.obj(k); o.py = F.obj(k + 1); o.pz = F.obj(k + 2);
        rotation_matrix(F.obj(k + 3), F.obj(k + 4), F.obj(k + 5), o.m);
        o.sx = F.obj(k + 6); o.sy = F.obj(k + 7); o.sz = F.obj(k + 8);
        o.a = F.obj(k + 9); o.b = F.obj(k + 10); o.c = F.obj(k + 11); o.d = F.obj(k + 12);
        o.brad = F.obj(k + 13);
        o.materialID = (float)(f2i(F.obj(k + 14)) - 1);
        o.lightID = (float)(f2i(F.obj(k + 15)) - 1);
    }
    for (int i = 0; i < nSdf; i++) { /* UnpackSDF shader.comp:209-214 */
        PtDevSdf& o = sdfs[i];
        o.px = F.at(PT_OFF_SDF, 6 * i); o.py = F.at(PT_OFF_SDF, 6 * i + 1); o.pz = F.at(PT_OFF_SDF, 6 * i + 2);
        o.sx = F.at(PT_OFF_SDF, 6 * i + 3); o.sy = F.at(PT_OFF_SDF, 6 * i + 4); o.sz = F.at(PT_OFF_SDF, 6 * i + 5);
    }

    /* SampleRandomLightSource (shader.comp:1225-1285) for every value randomLight can take: 0..numLights */
    sc->numLights = num[6];
    sc->invNumLights = 1.0f / num[6];
    const int nLights = count(num[6]);
    if (nLights + 1 > PT_DEV_MAX_LIGHT_SLOTS) {
        if (err) *err = "too many sampled lights";
        return PT_ERR_ARG;
    }
    sc->nLightSlots = (num[6] > 0.0f) ? nLights + 1 : 0;
    for (int j = 0; j < sc->nLightSlots; j++) {
        PtDevLightSlot& s = sc->lightSlots[j];
        const int ret = f2i(F.at(PT_OFF_LID, j));
        int id = ret;
        int off = 0;
        /* defaults of SampleLightSource's locals (shader.comp:1301-1304) for the fall-through `return 0` */
        s.px = s.py = s.pz = 0.0f; s.boundingRadius = 0.0f; s.lightID = -1.0f; s.objectID = 0;
        if (id < iS) {
            const int k = 6 * id;
            s.boundingRadius = F.obj(k + 3);
            s.px = F.obj(k); s.py = F.obj(k + 1); s.pz = F.obj(k + 2);
            s.lightID = (float)(f2i(F.obj(k + 5)) - 1);
            s.objectID = ret;
            continue;
        }
        id -= iS; off += 6 * iS;
        if (id < iP) {
            const int k = 5 * id + off;
            s.boundingRadius = 1e5f;
            s.px = F.obj(k); s.py = F.obj(k + 1); s.pz = F.obj(k + 2);
            s.lightID = (float)(f2i(F.obj(k + 4)) - 1);
            s.objectID = ret;
            continue;
        }
        id -= iP; off += 5 * iP;
        if (id < iB) {
            const int k = 11 * id + off;
            const float sx = F.obj(k + 6), sy = F.obj(k + 7), sz = F.obj(k + 8);
            s.boundingRadius = 0.5f * pt_sqrt(sx * sx + sy * sy + sz * sz);
            s.px = F.obj(k); s.py = F.obj(k + 1); s.pz = F.obj(k + 2);
            s.lightID = (float)(f2i(F.obj(k + 10)) - 1);
            s.objectID = ret;
            continue;
        }
        id -= iB; off += 11 * iB;
        if (id < iB) { /* sic: shader.comp:1264 compares with the box count */
            const int k = 12 * id + off;
            const float radius = F.obj(k + 6), thickness = F.obj(k + 8);
            s.boundingRadius = pt_sqrt((radius * radius) + (0.25f * thickness * thickness));
            s.px = F.obj(k); s.py = F.obj(k + 1); s.pz = F.obj(k + 2);
            s.lightID = (float)(f2i(F.obj(k + 11)) - 1);
            s.objectID = ret;
            continue;
        }
        id -= iL; off += 12 * iL;
        if (id < iL) { /* sic: shader.comp:1275 compares with the lens count */
            const int k = 16 * id + off;
            s.boundingRadius = pt_sqrt(F.obj(k + 13));
            s.px = F.obj(k); s.py = F.obj(k + 1); s.pz = F.obj(k + 2);
            s.lightID = (float)(f2i(F.obj(k + 15)) - 1);
            s.objectID = ret;
            continue;
        }
    }
    return PT_OK;
}

int pt_prepare_params(const pt_params* p, int accum_mode, int first_sample, int n_samples, PtDevParams* d,
                      std::string* err) {
    memset(d, 0, sizeof *d);
    if (p->resolution[0] <= 0 || p->resolution[1] <= 0) {
        if (err) *err = "resolution must be positive";
        return PT_ERR_ARG;
    }
    d->width = p->resolution[0];
    d->height = p->resolution[1];
    d->pathLength = p->pathLength;
    if (accum_mode == 2) { /* raw sum of an explicit sample range */
        d->firstSample = first_sample;
        d->samplesPerFrame = n_samples;
        d->accumMode = 2;
    } else {
        if (p->samplesPerFrame <= 0) {
            if (err) *err = "samplesPerFrame must be positive";
            return PT_ERR_ARG;
        }
        d->firstSample = p->frame - p->samplesPerFrame; /* shader.comp:954 */
        d->samplesPerFrame = p->samplesPerFrame;
        if ((p->currentSamples == p->samplesPerFrame) && (p->frame > p->samplesPerFrame)) { /* shader.comp:1500 */
            d->accumMode = 1;
            d->accumWeight = pt_pow(2.0f, -8.0f / (p->FPS * p->persistence));
        } else {
            d->accumMode = 0;
            const int unitSamples = p->currentSamples / p->samplesPerFrame; /* shader.comp:1504 */
            if (unitSamples <= 0) {
                if (err) *err = "currentSamples / samplesPerFrame must be >= 1";
                return PT_ERR_ARG;
            }
            d->accumN = (float)unitSamples;
            d->accumNm1 = (float)(unitSamples - 1);
        }
    }
    d->spfFloat = (float)d->samplesPerFrame;
    d->exposure = p->apertureSize * p->apertureSize * (float)p->ISO; /* shader.comp:1519 */
    d->resX = (float)p->resolution[0];
    d->resY = (float)p->resolution[1];
    d->camPosX = p->cameraPosX; d->camPosY = p->cameraPosY; d->camPosZ = p->cameraPosZ;
    d->sensorScale = -p->cameraSize * 0.5f;   /* shader.comp:1457 */
    d->halfAperture = 0.5f * p->apertureSize; /* shader.comp:1461 */
    d->apertureDist = p->apertureDist;
    rotation_matrix(p->cameraAngle[0], p->cameraAngle[1], 0.0f, d->camM); /* shader.comp:1456 */
    /* TracePathLens' lens (shader.comp:1411-1418): pos = cameraPos + forwardDir * lensDistance */
    const float fx = d->camM[2], fy = d->camM[5], fz = d->camM[8]; /* (M[0][2], M[1][2], M[2][2]) */
    prepare_lens(p->cameraPosX + fx * p->lensDistance, p->cameraPosY + fy * p->lensDistance,
                 p->cameraPosZ + fz * p->lensDistance, 0.0f, 90.0f - p->cameraAngle[1], p->cameraAngle[0],
                 p->lensRadius, p->lensFocalLength, p->lensThickness, true, &d->camLens);
    d->camLens.materialID = 0.0f;
    d->camLens.lightID = 0.0f;
    return PT_OK;
}
