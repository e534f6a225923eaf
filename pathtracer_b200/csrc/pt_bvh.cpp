/* pt_bvh.cpp -- host-side BVH build over the bounded primitives of a prepared scene (layout: pt_bvh.h).
 *
 * Full-sweep SAH, one primitive per leaf (170 primitives at most: the build is microseconds).  The boxes are
 * PADDED so that culling stays conservative with respect to the primitives' own fp32 arithmetic: the reference's
 * quadratic `b*b - 4*c` carries an absolute error of up to 2.4e-6 * D^2 for a ray origin at distance D, so a sphere of
 * radius r "exists" for the shader out to sqrt(r^2 + 6e-7 D^2) at worst.  Every primitive's bounding radius is therefore
 * grown to sqrt(r^2 + 2^-18 S^2) + 2^-16 S, S = the scene's scale (2 x the diagonal of everything bounded, origin
 * included), which covers every ray that starts within S of the primitives.  Rays from farther away exist -- a path
 * that hit the infinite ground plane near the horizon and bounces back -- and for them the shader's spheres are
 * mostly rounding noise (tests/bvh_check.cpp: hits units away from the geometry at D = 5000); pt_bvh_traverse
 * inflates every box per ray by 2^-9 (D - S) for those (header: centre and radius of the bounded set).
 * With both, the tree returns bit for bit what the in-order scan returns (tests/test_bvh.py on the host,
 * tests/test_gpu_parity.py::test_bvh_* on the GPU against the oracle's scan).
 */
#include <algorithm>
#include <string.h>

#include "pt_internal.h"
#include "pt_bvh.h"

namespace {

struct Prim {
    float lo[3], hi[3], c[3];
    int ref;
};

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; } }
    void add(const Prim& p) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p.lo[k]); hi[k] = std::max(hi[k], p.hi[k]); } }
    float area() const {
        const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return 2.0f * (x * y + y * z + z * x);
    }
};

struct Builder {
    std::vector<Prim> prims;
    std::vector<float> nodes;
    int nNodes = 0;

    Box bounds(const std::vector<int>& ids) const {
        Box b; b.reset();
        for (int i : ids) b.add(prims[i]);
        return b;
    }

    /* returns the child reference for this subset: ~leaf or the index of a freshly written inner node */
    int build(std::vector<int>& ids, int depth) {
        if (ids.size() == 1) return ~prims[ids[0]].ref;
        const int me = nNodes++;
        size_t bestSplit = ids.size() / 2;
        int bestAxis = 0;
        if (depth < 16) {
            float bestCost = 3.0e38f;
            for (int axis = 0; axis < 3; axis++) {
                std::vector<int> s = ids;
                std::stable_sort(s.begin(), s.end(), [&](int a, int b) { return prims[a].c[axis] < prims[b].c[axis]; });
                std::vector<float> right(s.size() + 1, 0.0f);
                Box b; b.reset();
                for (size_t i = s.size(); i-- > 1;) { b.add(prims[s[i]]); right[i] = b.area(); }
                b.reset();
                for (size_t i = 1; i < s.size(); i++) {
                    b.add(prims[s[i - 1]]);
                    const float cost = b.area() * (float)i + right[i] * (float)(s.size() - i);
                    if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = i; }
                }
            }
        } else { /* depth guard: median on the longest axis halves the set every level */
            const Box b = bounds(ids);
            for (int k = 1; k < 3; k++)
                if (b.hi[k] - b.lo[k] > b.hi[bestAxis] - b.lo[bestAxis]) bestAxis = k;
        }
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return prims[a].c[bestAxis] < prims[b].c[bestAxis]; });
        std::vector<int> left(ids.begin(), ids.begin() + bestSplit), right(ids.begin() + bestSplit, ids.end());
        const Box b0 = bounds(left), b1 = bounds(right);
        const int r0 = build(left, depth + 1);
        const int r1 = build(right, depth + 1);
        float* n = nodes.data() + (size_t)PT_BVH_NODE_FLOATS * me;
        n[0] = b0.lo[0]; n[1] = b0.lo[1]; n[2] = b0.lo[2]; n[3] = b0.hi[0];
        n[4] = b0.hi[1]; n[5] = b0.hi[2]; n[6] = b1.lo[0]; n[7] = b1.lo[1];
        n[8] = b1.lo[2]; n[9] = b1.hi[0]; n[10] = b1.hi[1]; n[11] = b1.hi[2];
        memcpy(n + 12, &r0, 4);
        memcpy(n + 13, &r1, 4);
        n[14] = 0.0f; n[15] = 0.0f;
        return me;
    }
};

}  // namespace

/* what goes into the tree: spheres, boxes, lenses (planes are unbounded; cyclides: see pt_bvh.h) */
int pt_bvh_bounded_prims(const PtDevScene* sc) { return sc->nSpheres + sc->nBoxes + sc->nLenses; }

/* blob = header, (n - 1) nodes, then sc->pool[0 .. offCyclides) rounded up to a multiple of 4 floats */
int pt_bvh_build(const PtDevScene* sc, std::vector<float>* blob, std::string* err) {
    const int n = pt_bvh_bounded_prims(sc);
    if (n < 2 || n > PT_BVH_MAX_PRIMS) {
        if (err) *err = "pt_bvh_build: needs 2.." + std::to_string(PT_BVH_MAX_PRIMS) + " bounded primitives";
        return PT_ERR_ARG;
    }
    const PtDevSphere* spheres = reinterpret_cast<const PtDevSphere*>(sc->pool);
    const PtDevPlane* planes = reinterpret_cast<const PtDevPlane*>(sc->pool + sc->offPlanes);
    const PtDevBox* boxes = reinterpret_cast<const PtDevBox*>(sc->pool + sc->offBoxes);
    const PtDevLens* lenses = reinterpret_cast<const PtDevLens*>(sc->pool + sc->offLenses);
    const PtDevSdf* sdfs = reinterpret_cast<const PtDevSdf*>(sc->pool + sc->offSdfs);

    struct Ball { float x, y, z, r; int ref; };
    std::vector<Ball> balls;
    for (int i = 0; i < sc->nSpheres; i++)
        balls.push_back({spheres[i].px, spheres[i].py, spheres[i].pz, fabsf(spheres[i].radius), (PT_BVH_SPHERE << 16) | i});
    for (int i = 0; i < sc->nBoxes; i++) /* the OBB lies inside the sphere BoundingSphere() culls with (shader.comp:887) */
        balls.push_back({boxes[i].px, boxes[i].py, boxes[i].pz, sqrtf(fmaxf(boxes[i].bound2, 0.0f)), (PT_BVH_BOX << 16) | i});
    for (int i = 0; i < sc->nLenses; i++) {
        /* the culling sphere of shader.comp:899-905 does not contain the apex of a thick lens; cover the caps too.
         * In the lens frame a cap spans x in [-(shift + sradius), -(shift + sradius) + h] (mirrored for the second
         * one), h = sradius - sliceOffset its height, and sqrt(sradius^2 - sliceOffset^2) across (shader.comp:366-431) */
        const PtDevLens& o = lenses[i];
        const float axial = fabsf(o.shift + o.sradius) + fabsf(o.sradius - o.sliceOffset);
        const float across2 = fmaxf(o.sradius2 - o.sliceOffset * o.sliceOffset, 0.0f);
        const float r = fmaxf(sqrtf(fmaxf(o.bound2, 0.0f)), sqrtf(axial * axial + across2));
        balls.push_back({o.px, o.py, o.pz, r, (PT_BVH_LENS << 16) | i});
    }
    const size_t firstLens = balls.size() - (size_t)sc->nLenses;
    /* scene scale: everything bounded, the SDF boxes, the plane heights and the origin */
    float lo[3] = {0.0f, 0.0f, 0.0f}, hi[3] = {0.0f, 0.0f, 0.0f};
    auto grow = [&](float x, float y, float z, float r) {
        const float p[3] = {x, y, z};
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k] - r); hi[k] = std::max(hi[k], p[k] + r); }
    };
    for (const Ball& b : balls) grow(b.x, b.y, b.z, b.r);
    for (int i = 0; i < sc->nSdfs; i++)
        grow(sdfs[i].px, sdfs[i].py, sdfs[i].pz, 0.5f * fmaxf(fmaxf(fabsf(sdfs[i].sx), fabsf(sdfs[i].sy)), fabsf(sdfs[i].sz)));
    for (int i = 0; i < sc->nPlanes; i++) grow(0.0f, planes[i].py, 0.0f, 0.0f);
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    const float S = 2.0f * sqrtf(dx * dx + dy * dy + dz * dz);
    if (!(S < 1e18f)) {
        if (err) *err = "pt_bvh_build: scene extent is not finite";
        return PT_ERR_ARG;
    }

    Builder B;
    for (size_t k = 0; k < balls.size(); k++) {
        const Ball& b = balls[k];
        /* spheres (and the culling spheres of boxes) carry the noise of a quadratic of radius r; a lens cap is cut
         * from a sphere of radius 2f >> r, whose grazing hits wander by sqrt(kappa) S along the ray: linear padding */
        const float r = (k >= firstLens) ? b.r + PT_BVH_SQRT_KAPPA * S + 1.5258789e-5f * S
                                         : sqrtf(b.r * b.r + PT_BVH_KAPPA * S * S) + 1.5258789e-5f * S; /* + 2^-16 S */
        Prim p;
        const float c[3] = {b.x, b.y, b.z};
        for (int k = 0; k < 3; k++) { p.lo[k] = c[k] - r; p.hi[k] = c[k] + r; p.c[k] = c[k]; }
        p.ref = b.ref;
        B.prims.push_back(p);
    }
    B.nodes.assign((size_t)PT_BVH_NODE_FLOATS * (n - 1), 0.0f);
    std::vector<int> ids(n);
    for (int i = 0; i < n; i++) ids[i] = i;
    const int root = B.build(ids, 0);
    if (root != 0 || B.nNodes != n - 1) {
        if (err) *err = "pt_bvh_build: internal error";
        return PT_ERR_ARG;
    }
    const int poolFloats = (sc->offCyclides + 3) & ~3; /* spheres, planes, boxes, lenses */
    /* header: centre and radius of everything bounded (for the far-origin inflation of pt_bvh_traverse) */
    float blo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (const Ball& b : balls) {
        const float c[3] = {b.x, b.y, b.z};
        for (int k = 0; k < 3; k++) { blo[k] = std::min(blo[k], c[k] - b.r); bhi[k] = std::max(bhi[k], c[k] + b.r); }
    }
    const float C[3] = {0.5f * (blo[0] + bhi[0]), 0.5f * (blo[1] + bhi[1]), 0.5f * (blo[2] + bhi[2])};
    const float ex = bhi[0] - blo[0], ey = bhi[1] - blo[1], ez = bhi[2] - blo[2];
    const float R = 0.5f * sqrtf(ex * ex + ey * ey + ez * ez);
    blob->assign({C[0], C[1], C[2], R - S});
    blob->insert(blob->end(), B.nodes.begin(), B.nodes.end());
    blob->insert(blob->end(), sc->pool, sc->pool + poolFloats);
    return PT_OK;
}
