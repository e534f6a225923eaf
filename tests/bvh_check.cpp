/* bvh_check.cpp -- host-side checker of the BVH builder and of the traversal template (pt_bvh.h, pt_bvh.cpp).
 *
 * Test infrastructure (compiled and run by tests/test_bvh.py with g++, no GPU): reads a packed uniform block,
 * prepares the scene and builds the tree exactly as pt_set_scene does, then shoots rays and compares the closest
 * hit found through pt_bvh_traverse with the in-order brute-force scan -- (t, object index) must be identical bit
 * for bit, ties included.  Spheres use SphereIntersection's arithmetic (shader.comp:289-317; g++ -ffp-contract=off
 * gives the strict kernels' rounding); boxes and lenses are stood in for by the sphere BoundingSphere() culls them
 * with, which is what decides whether their intersection routine runs at all (cyclides are not in the tree; the real
 * box / lens / cyclide routines go through the tree in tests/test_bvh.py's oracle-through-the-tree check).
 *
 * usage: bvh_check <ubo.bin> <n_rays> <seed> <camx> <camy> <camz>     prints "rays N mismatches M visits V visits_near V2 prims K"
 * (leaf tests per ray: all rays / rays that start within the scene's scale)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <string>
#include <vector>

#include "pt_internal.h"
#include "pt_bvh.h"

namespace {

struct Ball { float x, y, z, r2; int id; };

unsigned g_state;
float rnd() { /* PCG32 of shader.comp:937-946: any deterministic stream will do */
    unsigned state = g_state * 747796405u + 2891336453u;
    unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    g_state = (word >> 22u) ^ word;
    return (float)g_state * 2.3283064365386963e-10f;
}

inline void sphere_hit(const Ball& o, const float* ro, const float* rd, float& bestT, int& bestId, bool tie) {
    const float lox = ro[0] - o.x, loy = ro[1] - o.y, loz = ro[2] - o.z;
    const float b = 2.0f * (rd[0] * lox + rd[1] * loy + rd[2] * loz);
    const float cc = (lox * lox + loy * loy + loz * loz) - o.r2;
    const float disc = b * b - 4.0f * cc;
    if (disc < 0.0f) return;
    const float s = sqrtf(disc);
    const float t1 = (-b - s) * 0.5f, t2 = (-b + s) * 0.5f;
    const float t = (t1 > 0.0f) ? t1 : t2;
    if (t < 1e-4f) return;
    if ((t < bestT) || (tie && t == bestT && o.id < bestId)) { bestT = t; bestId = o.id; }
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: bvh_check ubo.bin n_rays seed camx camy camz\n"); return 2; }
    static pt_ubo ubo;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(&ubo, sizeof ubo, 1, f) != 1) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    const long nRays = atol(argv[2]);
    g_state = (unsigned)atol(argv[3]);
    const float cam[3] = {(float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6])};

    static PtDevScene sc;
    std::string err;
    if (pt_prepare_scene(&ubo, &sc, &err) != PT_OK) { fprintf(stderr, "prepare: %s\n", err.c_str()); return 2; }
    std::vector<float> blob;
    if (pt_bvh_build(&sc, &blob, &err) != PT_OK) { fprintf(stderr, "build: %s\n", err.c_str()); return 2; }
    const int nPrims = pt_bvh_bounded_prims(&sc);
    const float* nodes = blob.data() + PT_BVH_HEADER_FLOATS;
    const float* recs = nodes + PT_BVH_NODE_FLOATS * (nPrims - 1);
    if (memcmp(recs, sc.pool, sizeof(float) * sc.offCyclides) != 0) { fprintf(stderr, "record copy differs from the pool\n"); return 1; }

    /* every bounded primitive as (centre, culling radius^2, global object index), in the reference's order */
    std::vector<Ball> balls;
    std::vector<int> typeBase = {0, sc.nSpheres + sc.nPlanes, sc.nSpheres + sc.nPlanes + sc.nBoxes,
                                 sc.nSpheres + sc.nPlanes + sc.nBoxes + sc.nLenses};
    const PtDevSphere* S = reinterpret_cast<const PtDevSphere*>(sc.pool);
    const PtDevBox* B = reinterpret_cast<const PtDevBox*>(sc.pool + sc.offBoxes);
    const PtDevLens* L = reinterpret_cast<const PtDevLens*>(sc.pool + sc.offLenses);
    std::vector<std::vector<Ball>> byType(3);
    for (int i = 0; i < sc.nSpheres; i++) byType[0].push_back({S[i].px, S[i].py, S[i].pz, S[i].r2, typeBase[0] + i});
    for (int i = 0; i < sc.nBoxes; i++) byType[1].push_back({B[i].px, B[i].py, B[i].pz, B[i].bound2, typeBase[1] + i});
    for (int i = 0; i < sc.nLenses; i++) byType[2].push_back({L[i].px, L[i].py, L[i].pz, L[i].bound2, typeBase[2] + i});
    for (auto& v : byType) balls.insert(balls.end(), v.begin(), v.end());

    /* every leaf must appear exactly once */
    std::vector<int> seen(nPrims, 0);
    for (int n = 0; n < nPrims - 1; n++)
        for (int k = 0; k < 2; k++) {
            int ref;
            memcpy(&ref, nodes + PT_BVH_NODE_FLOATS * n + 12 + k, 4);
            if (ref < 0) {
                const int type = (~ref) >> 16, idx = (~ref) & 0xffff;
                if (type < 0 || type > 2 || idx >= (int)byType[type].size()) { fprintf(stderr, "bad leaf ref\n"); return 1; }
                int flat = idx;
                for (int t = 0; t < type; t++) flat += (int)byType[t].size();
                seen[flat]++;
            } else if (ref <= n || ref >= nPrims - 1) { fprintf(stderr, "bad inner ref %d at node %d\n", ref, n); return 1; }
        }
    for (int i = 0; i < nPrims; i++)
        if (seen[i] != 1) { fprintf(stderr, "primitive %d referenced %d times\n", i, seen[i]); return 1; }

    long mismatches = 0, visits = 0, visitsNear = 0, nNear = 0;
    float ro[3], rd[3];
    for (long r = 0; r < nRays; r++) {
        const Ball& target = balls[(size_t)(rnd() * (float)balls.size()) % balls.size()];
        const int kind = (int)(r % 5);
        if (kind == 0) { /* primary-like: from the camera toward a point near a primitive */
            memcpy(ro, cam, sizeof ro);
        } else if (kind == 1) { /* secondary: from a point on some primitive's surface */
            const Ball& from = balls[(size_t)(rnd() * (float)balls.size()) % balls.size()];
            float u[3] = {rnd() - 0.5f, rnd() - 0.5f, rnd() - 0.5f};
            const float len = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) + 1e-20f, rad = sqrtf(from.r2);
            ro[0] = from.x + u[0] / len * rad; ro[1] = from.y + u[1] / len * rad; ro[2] = from.z + u[2] / len * rad;
        } else if (kind == 2) { /* from a point on the ground plane, well outside the cluster */
            ro[0] = (rnd() - 0.5f) * 60.0f; ro[1] = 0.0f; ro[2] = (rnd() - 0.5f) * 60.0f;
        } else if (kind == 4) { /* from a far point of the ground plane (the horizon of a rendered frame): 1e2 .. 1e5 away */
            const float dist = 100.0f * powf(10.0f, 3.0f * rnd()), phi = 6.2831853f * rnd();
            ro[0] = dist * cosf(phi); ro[1] = 0.0f; ro[2] = dist * sinf(phi);
        } else { /* from inside a primitive */
            ro[0] = target.x; ro[1] = target.y; ro[2] = target.z;
        }
        /* aim at a point within 1.2 culling radii of the target's centre: plenty of grazing rays */
        const float rad = 1.2f * sqrtf(target.r2);
        float aim[3] = {target.x + (2.0f * rnd() - 1.0f) * rad, target.y + (2.0f * rnd() - 1.0f) * rad, target.z + (2.0f * rnd() - 1.0f) * rad};
        if (kind == 3) { aim[0] = ro[0] + rnd() - 0.5f; aim[1] = ro[1] + rnd() - 0.5f; aim[2] = ro[2] + rnd() - 0.5f; }
        float d[3] = {aim[0] - ro[0], aim[1] - ro[1], aim[2] - ro[2]};
        const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (!(len > 0.0f)) continue;
        rd[0] = d[0] / len; rd[1] = d[1] / len; rd[2] = d[2] / len;
        if (r % 97 == 0) rd[(r / 97) % 3] = 0.0f; /* axis-parallel components: 1/0 = inf in the slab test */

        float tA = 1e5f; int idA = -1;
        for (const Ball& b : balls) sphere_hit(b, ro, rd, tA, idA, false);
        float tB = 1e5f; int idB = -1;
        pt_bvh_traverse(blob.data(), ro[0], ro[1], ro[2], rd[0], rd[1], rd[2], tB, [&](int ref) {
            const int type = ref >> 16, idx = ref & 0xffff;
            visits++;
            if (kind != 4) visitsNear++;
            sphere_hit(byType[type][idx], ro, rd, tB, idB, true);
        });
        if (kind != 4) nNear++;
        if (memcmp(&tA, &tB, 4) != 0 || idA != idB) {
            if (mismatches < 5)
                fprintf(stderr, "ray %ld kind %d: scan (%.9g, %d) tree (%.9g, %d) o=(%g %g %g) d=(%g %g %g)\n", r, kind, tA, idA, tB,
                        idB, ro[0], ro[1], ro[2], rd[0], rd[1], rd[2]);
            mismatches++;
        }
    }
    printf("rays %ld mismatches %ld visits %.3f visits_near %.3f prims %d\n", nRays, mismatches, (double)visits / (double)nRays,
           (double)visitsNear / (double)(nNear > 0 ? nNear : 1), nPrims);
    return mismatches ? 1 : 0;
}
