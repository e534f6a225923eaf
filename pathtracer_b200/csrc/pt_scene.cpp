/* pt_scene.cpp -- scene files -> the kernel's data contract.
 *
 * Host-side mirror of the reference's scene model (host:110-185) and of the three functions that sit between a
 * scene file and the dispatch:
 *   UpdateFromJSON      host:2576-2722   JSON -> object vectors (same schema; missing arrays are legal)
 *   UpdateUniformBuffer host:3642-3811   object vectors -> the 4097-float uniform block
 *   UpdatePushConstant  host:3813-3834   camera + render settings -> the 88-byte push block
 * The packer reproduces the observable quirks: light registration of lenses and cyclides is keyed on
 * planes[i].lightID (host:3704,3728), the cyclide bounding radius is packed as brad^2 * max(scale)^2
 * (host:3723-3725), ids stay 1-based floats (the shader subtracts 1), cameraAngle = (-angle.y, angle.x).
 */
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "pt_abi.h"
#include "pt_json.h"

namespace {

const float kCIE[PT_CIE_SIZE] = {
#include "pt_cie1931.inc"
};

struct Sphere { float pos[3]; float radius; int materialID, lightID; };
struct Plane { float pos[3]; int materialID, lightID; };
struct Box { float pos[3], rotation[3], size[3]; int materialID, lightID; };
struct Lens { float pos[3], rotation[3]; float radius, focalLength, thickness; bool isConverging; int materialID, lightID; };
struct Cyclide { float pos[3], rotation[3], scale[3]; float a, b, c, d, brad; int materialID, lightID; };
struct Sdf { float pos[3], size[3]; std::string glsl; };
struct Material { float reflection[3]; int bsdf = 0; float roughness = 0.0f, ior = 0.0f; /* surface extension (pt_abi.h), absent in the reference's schema */ };
struct Light { float emission[2]; };
struct CameraShot { float pos[3]; float angle[2]; };
struct Camera { int ISO; float size, apertureSize, apertureDist, lensRadius, lensFocalLength, lensThickness, lensDistance; };

thread_local std::string g_error;

float f(const PtJson& v) { return (float)v.number(); }     /* double -> float like nlohmann's get<float>() */
int i32(const PtJson& v) { return (int)v.number(); }       /* double/int -> int (truncation) */
void vec3(const PtJson& v, float* o) { for (int k = 0; k < 3; k++) o[k] = f(v.at(k)); }

}  // namespace

struct pt_scene {
    std::vector<CameraShot> shots;
    Camera camera;
    std::vector<Sphere> spheres;
    std::vector<Plane> planes;
    std::vector<Box> boxes;
    std::vector<Lens> lenses;
    std::vector<Cyclide> cyclides;
    std::vector<Sdf> sdfs;
    std::vector<Material> materials;
    std::vector<Light> lights;
};

const std::string& pt_scene_last_error() { return g_error; }

extern "C" {

const float* pt_cie1931_table(void) { return kCIE; }

int pt_scene_parse_json(const char* text, pt_scene** out) {
    if (!text || !out) { g_error = "pt_scene_parse_json: null argument"; return PT_ERR_ARG; }
    PtJson j;
    std::string err;
    if (!pt_json_parse(text, strlen(text), &j, &err)) { g_error = "scene JSON: " + err; return PT_ERR_IO; }
    if (j.type != PtJson::Object || !j.find("camera")) { g_error = "scene JSON: no \"camera\" object"; return PT_ERR_IO; }
    pt_scene* s = new pt_scene();
    const PtJson& cam = j["camera"];
    const int numShots = i32(cam["numShots"]);
    if (numShots < 1 || (size_t)numShots > cam["position"].size() || (size_t)numShots > cam["angle"].size()) {
        delete s;
        g_error = "scene JSON: camera.numShots does not match camera.position / camera.angle";
        return PT_ERR_IO;
    }
    s->shots.resize((size_t)numShots);
    for (int k = 0; k < numShots; k++) { /* host:2577-2585 */
        vec3(cam["position"].at(k), s->shots[k].pos);
        s->shots[k].angle[0] = f(cam["angle"].at(k).at(0));
        s->shots[k].angle[1] = f(cam["angle"].at(k).at(1));
    }
    s->camera.ISO = i32(cam["ISO"]); /* host:2594-2608 */
    s->camera.size = f(cam["size"]);
    s->camera.apertureSize = f(cam["apertureSize"]);
    s->camera.apertureDist = f(cam["apertureDistance"]);
    s->camera.lensRadius = f(cam["lensRadius"]);
    s->camera.lensFocalLength = f(cam["lensFocalLength"]);
    s->camera.lensThickness = f(cam["lensThickness"]);
    s->camera.lensDistance = f(cam["lensDistance"]);

    const PtJson& sp = j["sphere"]; /* host:2610-2620 */
    s->spheres.resize(sp.size());
    for (size_t k = 0; k < sp.size(); k++) {
        vec3(sp.at(k)["position"], s->spheres[k].pos);
        s->spheres[k].radius = f(sp.at(k)["radius"]);
        s->spheres[k].materialID = i32(sp.at(k)["materialID"]);
        s->spheres[k].lightID = i32(sp.at(k)["lightID"]);
    }
    const PtJson& pl = j["plane"]; /* host:2622-2630 */
    s->planes.resize(pl.size());
    for (size_t k = 0; k < pl.size(); k++) {
        vec3(pl.at(k)["position"], s->planes[k].pos);
        s->planes[k].materialID = i32(pl.at(k)["materialID"]);
        s->planes[k].lightID = i32(pl.at(k)["lightID"]);
    }
    const PtJson& bx = j["box"]; /* host:2632-2648 */
    s->boxes.resize(bx.size());
    for (size_t k = 0; k < bx.size(); k++) {
        vec3(bx.at(k)["position"], s->boxes[k].pos);
        vec3(bx.at(k)["rotation"], s->boxes[k].rotation);
        vec3(bx.at(k)["size"], s->boxes[k].size);
        s->boxes[k].materialID = i32(bx.at(k)["materialID"]);
        s->boxes[k].lightID = i32(bx.at(k)["lightID"]);
    }
    const PtJson& ln = j["lens"]; /* host:2650-2670 */
    s->lenses.resize(ln.size());
    for (size_t k = 0; k < ln.size(); k++) {
        vec3(ln.at(k)["position"], s->lenses[k].pos);
        vec3(ln.at(k)["rotation"], s->lenses[k].rotation);
        s->lenses[k].radius = f(ln.at(k)["radius"]);
        s->lenses[k].focalLength = f(ln.at(k)["focalLength"]);
        s->lenses[k].thickness = f(ln.at(k)["thickness"]);
        s->lenses[k].isConverging = ln.at(k)["isConverging"].truthy();
        s->lenses[k].materialID = i32(ln.at(k)["materialID"]);
        s->lenses[k].lightID = i32(ln.at(k)["lightID"]);
    }
    const PtJson& cy = j["cyclide"]; /* host:2672-2695 */
    s->cyclides.resize(cy.size());
    for (size_t k = 0; k < cy.size(); k++) {
        vec3(cy.at(k)["position"], s->cyclides[k].pos);
        vec3(cy.at(k)["rotation"], s->cyclides[k].rotation);
        vec3(cy.at(k)["scale"], s->cyclides[k].scale);
        s->cyclides[k].a = f(cy.at(k)["a"]);
        s->cyclides[k].b = f(cy.at(k)["b"]);
        s->cyclides[k].c = f(cy.at(k)["c"]);
        s->cyclides[k].d = f(cy.at(k)["d"]);
        s->cyclides[k].brad = f(cy.at(k)["boundingRadius"]);
        s->cyclides[k].materialID = i32(cy.at(k)["materialID"]);
        s->cyclides[k].lightID = i32(cy.at(k)["lightID"]);
    }
    const PtJson& sd = j["sdf"]; /* host:2697-2708 */
    s->sdfs.resize(sd.size());
    for (size_t k = 0; k < sd.size(); k++) {
        vec3(sd.at(k)["position"], s->sdfs[k].pos);
        vec3(sd.at(k)["boundingSize"], s->sdfs[k].size);
        s->sdfs[k].glsl = sd.at(k)["glsl"].str;
    }
    const PtJson& mt = j["material"]; /* host:2710-2715 */
    s->materials.resize(mt.size());
    for (size_t k = 0; k < mt.size(); k++) {
        const PtJson& r = mt.at(k)["reflection"];
        s->materials[k].reflection[0] = f(r["peakWavelength"]);
        s->materials[k].reflection[1] = f(r["sigma"]);
        s->materials[k].reflection[2] = r["isInvert"].truthy() ? 1.0f : 0.0f;
        if (const PtJson* b = mt.at(k).find("bsdf")) { /* extension of the schema; the reference ignores unknown keys */
            const std::string& n = b->str;
            int id = -1;
            if (n == "reference" || n == "diffuse") id = PT_BSDF_REFERENCE;
            else if (n == "mirror") id = PT_BSDF_MIRROR;
            else if (n == "glossy") id = PT_BSDF_GLOSSY;
            else if (n == "dielectric") id = PT_BSDF_DIELECTRIC;
            if (b->type != PtJson::String || id < 0) {
                g_error = "scene JSON: material " + std::to_string(k) + ": \"bsdf\" must be \"reference\", \"mirror\", \"glossy\" or \"dielectric\"";
                delete s;
                return PT_ERR_IO;
            }
            s->materials[k].bsdf = id;
            if (const PtJson* v = mt.at(k).find("roughness")) s->materials[k].roughness = f(*v);
            if (const PtJson* v = mt.at(k).find("ior")) s->materials[k].ior = f(*v);
        }
    }
    const PtJson& lt = j["light"]; /* host:2717-2721 */
    s->lights.resize(lt.size());
    for (size_t k = 0; k < lt.size(); k++) {
        const PtJson& e = lt.at(k)["emission"];
        s->lights[k].emission[0] = f(e["temperature"]);
        s->lights[k].emission[1] = f(e["luminosity"]);
    }
    *out = s;
    return PT_OK;
}

int pt_scene_load_json(const char* path, pt_scene** out) { /* ReadFile + parse: host:893-908 */
    if (!path || !out) { g_error = "pt_scene_load_json: null argument"; return PT_ERR_ARG; }
    FILE* fp = fopen(path, "rb");
    if (!fp) { g_error = std::string("cannot open scene file ") + path; return PT_ERR_IO; }
    std::string text;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
    fclose(fp);
    return pt_scene_parse_json(text.c_str(), out);
}

void pt_scene_free(pt_scene* scene) { delete scene; }
int pt_scene_num_shots(const pt_scene* scene) { return scene ? (int)scene->shots.size() : 0; }
int pt_scene_num_sdf(const pt_scene* scene) { return scene ? (int)scene->sdfs.size() : 0; }
const char* pt_scene_sdf_glsl(const pt_scene* scene, int i) {
    if (!scene || i < 0 || (size_t)i >= scene->sdfs.size()) return nullptr;
    return scene->sdfs[(size_t)i].glsl.c_str();
}

/* UpdateToJSON + SaveScene (host:2724-2858, 3465-3472): same key order, reals rounded to 1e-5 by RoundDecimal
 * (host:935-943: static_cast<int>(x * 1e5 +- 0.5) / 1e5), 4-space indentation.  Returns the size needed
 * (including NUL); writes at most cap bytes. */
long pt_scene_to_json(const pt_scene* s, char* out, size_t cap) {
    if (!s) { g_error = "pt_scene_to_json: null scene"; return PT_ERR_ARG; }
    auto num = [](double v) {
        PtJson j;
        j.type = PtJson::Number;
        v = (v >= 0.0) ? (double)(int)(v * 1e5 + 0.5) : (double)(int)(v * 1e5 - 0.5);
        j.num = v / 1e5;
        return j;
    };
    auto integer = [](long long v) { PtJson j; j.type = PtJson::Number; j.num = (double)v; j.is_int = true; return j; };
    auto boolean = [](bool v) { PtJson j; j.type = PtJson::Bool; j.b = v; return j; };
    auto arr3 = [&](const float* v, int n) { PtJson j; j.type = PtJson::Array; for (int i = 0; i < n; i++) j.arr.push_back(num(v[i])); return j; };
    auto object = []() { PtJson j; j.type = PtJson::Object; return j; };
    auto array = []() { PtJson j; j.type = PtJson::Array; return j; };
    PtJson root = object(), cam = object(), pos = array(), ang = array();
    cam.obj.emplace_back("numShots", integer((long long)s->shots.size()));
    for (const CameraShot& c : s->shots) { pos.arr.push_back(arr3(c.pos, 3)); ang.arr.push_back(arr3(c.angle, 2)); }
    cam.obj.emplace_back("position", pos);
    cam.obj.emplace_back("angle", ang);
    cam.obj.emplace_back("ISO", integer(s->camera.ISO));
    cam.obj.emplace_back("size", num(s->camera.size));
    cam.obj.emplace_back("apertureSize", num(s->camera.apertureSize));
    cam.obj.emplace_back("apertureDistance", num(s->camera.apertureDist));
    cam.obj.emplace_back("lensRadius", num(s->camera.lensRadius));
    cam.obj.emplace_back("lensFocalLength", num(s->camera.lensFocalLength));
    cam.obj.emplace_back("lensThickness", num(s->camera.lensThickness));
    cam.obj.emplace_back("lensDistance", num(s->camera.lensDistance));
    root.obj.emplace_back("camera", cam);
    auto ids = [&](PtJson& o, int materialID, int lightID) {
        o.obj.emplace_back("materialID", integer(materialID));
        o.obj.emplace_back("lightID", integer(lightID));
    };
    /* like nlohmann's operator[] on an index, an array key only exists when it has at least one element */
    if (!s->spheres.empty()) {
        PtJson a = array();
        for (const Sphere& v : s->spheres) {
            PtJson o = object();
            o.obj.emplace_back("position", arr3(v.pos, 3));
            o.obj.emplace_back("radius", num(v.radius));
            ids(o, v.materialID, v.lightID);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("sphere", a);
    }
    if (!s->planes.empty()) {
        PtJson a = array();
        for (const Plane& v : s->planes) {
            PtJson o = object();
            o.obj.emplace_back("position", arr3(v.pos, 3));
            ids(o, v.materialID, v.lightID);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("plane", a);
    }
    if (!s->boxes.empty()) {
        PtJson a = array();
        for (const Box& v : s->boxes) {
            PtJson o = object();
            o.obj.emplace_back("position", arr3(v.pos, 3));
            o.obj.emplace_back("rotation", arr3(v.rotation, 3));
            o.obj.emplace_back("size", arr3(v.size, 3));
            ids(o, v.materialID, v.lightID);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("box", a);
    }
    if (!s->lenses.empty()) {
        PtJson a = array();
        for (const Lens& v : s->lenses) {
            PtJson o = object();
            o.obj.emplace_back("position", arr3(v.pos, 3));
            o.obj.emplace_back("rotation", arr3(v.rotation, 3));
            o.obj.emplace_back("radius", num(v.radius));
            o.obj.emplace_back("focalLength", num(v.focalLength));
            o.obj.emplace_back("thickness", num(v.thickness));
            o.obj.emplace_back("isConverging", boolean(v.isConverging));
            ids(o, v.materialID, v.lightID);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("lens", a);
    }
    if (!s->cyclides.empty()) {
        PtJson a = array();
        for (const Cyclide& v : s->cyclides) {
            PtJson o = object();
            o.obj.emplace_back("position", arr3(v.pos, 3));
            o.obj.emplace_back("rotation", arr3(v.rotation, 3));
            o.obj.emplace_back("scale", arr3(v.scale, 3));
            o.obj.emplace_back("a", num(v.a));
            o.obj.emplace_back("b", num(v.b));
            o.obj.emplace_back("c", num(v.c));
            o.obj.emplace_back("d", num(v.d));
            o.obj.emplace_back("boundingRadius", num(v.brad));
            ids(o, v.materialID, v.lightID);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("cyclide", a);
    }
    if (!s->sdfs.empty()) {
        PtJson a = array();
        for (const Sdf& v : s->sdfs) {
            PtJson o = object(), g;
            o.obj.emplace_back("position", arr3(v.pos, 3));
            o.obj.emplace_back("boundingSize", arr3(v.size, 3));
            g.type = PtJson::String;
            g.str = v.glsl;
            o.obj.emplace_back("glsl", g);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("sdf", a);
    }
    if (!s->materials.empty()) {
        PtJson a = array();
        for (const Material& v : s->materials) {
            PtJson o = object(), r = object();
            r.obj.emplace_back("peakWavelength", num(v.reflection[0]));
            r.obj.emplace_back("sigma", num(v.reflection[1]));
            r.obj.emplace_back("isInvert", boolean(v.reflection[2] != 0.0f));
            o.obj.emplace_back("reflection", r);
            if (v.bsdf != PT_BSDF_REFERENCE) {
                static const char* const names[4] = {"reference", "mirror", "glossy", "dielectric"};
                PtJson b;
                b.type = PtJson::String;
                b.str = names[v.bsdf];
                o.obj.emplace_back("bsdf", b);
                if (v.bsdf == PT_BSDF_GLOSSY) o.obj.emplace_back("roughness", num(v.roughness));
                if (v.bsdf == PT_BSDF_DIELECTRIC) o.obj.emplace_back("ior", num(v.ior));
            }
            a.arr.push_back(o);
        }
        root.obj.emplace_back("material", a);
    }
    if (!s->lights.empty()) {
        PtJson a = array();
        for (const Light& v : s->lights) {
            PtJson o = object(), e = object();
            e.obj.emplace_back("temperature", num(v.emission[0]));
            e.obj.emplace_back("luminosity", num(v.emission[1]));
            o.obj.emplace_back("emission", e);
            a.arr.push_back(o);
        }
        root.obj.emplace_back("light", a);
    }
    const std::string text = pt_json_dump(root, 4);
    if (out && cap > 0) {
        const size_t n = text.size() < cap - 1 ? text.size() : cap - 1;
        memcpy(out, text.data(), n);
        out[n] = '\0';
    }
    return (long)text.size() + 1;
}

int pt_scene_save_json(const pt_scene* s, const char* path) {
    if (!s || !path) { g_error = "pt_scene_save_json: null argument"; return PT_ERR_ARG; }
    const long n = pt_scene_to_json(s, nullptr, 0);
    if (n < 0) return (int)n;
    std::string buf((size_t)n, '\0');
    pt_scene_to_json(s, &buf[0], (size_t)n);
    FILE* fp = fopen(path, "wb");
    if (!fp) { g_error = std::string("cannot write ") + path; return PT_ERR_IO; }
    const bool ok = fwrite(buf.data(), 1, (size_t)n - 1, fp) == (size_t)n - 1;
    return (fclose(fp) == 0 && ok) ? PT_OK : PT_ERR_IO;
}

int pt_scene_surface_ext(const pt_scene* s, pt_surface_ext* out, int max) {
    if (!s || (max > 0 && !out)) { g_error = "pt_scene_surface_ext: null argument"; return PT_ERR_ARG; }
    int n = 0;
    for (size_t k = 0; k < s->materials.size(); k++)
        if (s->materials[k].bsdf != PT_BSDF_REFERENCE) n = (int)k + 1;
    if (n > PT_MAX_SURFACE_EXT) { g_error = "scene: only the first " + std::to_string(PT_MAX_SURFACE_EXT) + " materials can carry a \"bsdf\""; return PT_ERR_ARG; }
    for (int k = 0; k < n && k < max; k++) {
        out[k].bsdf = s->materials[k].bsdf;
        out[k].roughness = s->materials[k].roughness;
        out[k].ior = s->materials[k].ior;
        out[k].pad = 0.0f;
    }
    return n;
}

int pt_scene_pack_ubo(const pt_scene* s, pt_ubo* ubo) { /* host:3642-3811 */
    if (!s || !ubo) { g_error = "pt_scene_pack_ubo: null argument"; return PT_ERR_ARG; }
    std::vector<float> objects, sdfs, materials, lights, lightIDs;
    const size_t nS = s->spheres.size(), nP = s->planes.size(), nB = s->boxes.size(), nL = s->lenses.size();
    for (size_t i = 0; i < nS; i++) {
        const Sphere& o = s->spheres[i];
        objects.insert(objects.end(), {o.pos[0], o.pos[1], o.pos[2], o.radius, (float)o.materialID, (float)o.lightID});
        if (o.lightID > 0) lightIDs.push_back((float)i);
    }
    for (size_t i = 0; i < nP; i++) {
        const Plane& o = s->planes[i];
        objects.insert(objects.end(), {o.pos[0], o.pos[1], o.pos[2], (float)o.materialID, (float)o.lightID});
        if (o.lightID > 0) lightIDs.push_back((float)(nS + i));
    }
    for (size_t i = 0; i < nB; i++) {
        const Box& o = s->boxes[i];
        objects.insert(objects.end(), {o.pos[0], o.pos[1], o.pos[2], o.rotation[0], o.rotation[1], o.rotation[2],
                                       o.size[0], o.size[1], o.size[2], (float)o.materialID, (float)o.lightID});
        if (o.lightID > 0) lightIDs.push_back((float)(nS + nP + i));
    }
    for (size_t i = 0; i < nL; i++) {
        const Lens& o = s->lenses[i];
        objects.insert(objects.end(), {o.pos[0], o.pos[1], o.pos[2], o.rotation[0], o.rotation[1], o.rotation[2],
                                       o.radius, o.focalLength, o.thickness, o.isConverging ? 1.0f : 0.0f,
                                       (float)o.materialID, (float)o.lightID});
        /* host:3704 reads planes[i].lightID (sic).  Past the end of `planes` that is an out-of-bounds vector
         * read in the reference; here it is defined as "not registered". */
        if (i < nP && s->planes[i].lightID > 0) lightIDs.push_back((float)(nS + nP + nB + i));
    }
    for (size_t i = 0; i < s->cyclides.size(); i++) {
        const Cyclide& o = s->cyclides[i];
        const float a = o.scale[0] < o.scale[1] ? o.scale[1] : o.scale[0]; /* glm::max(x, y) = (x < y) ? y : x */
        const float m = a < o.scale[2] ? o.scale[2] : a;
        objects.insert(objects.end(), {o.pos[0], o.pos[1], o.pos[2], o.rotation[0], o.rotation[1], o.rotation[2],
                                       o.scale[0], o.scale[1], o.scale[2], o.a, o.b, o.c, o.d,
                                       o.brad * o.brad * (m * m), (float)o.materialID, (float)o.lightID});
        if (i < nP && s->planes[i].lightID > 0) lightIDs.push_back((float)(nS + nP + nB + nL + i)); /* host:3728 (sic) */
    }
    for (const Sdf& o : s->sdfs) sdfs.insert(sdfs.end(), {o.pos[0], o.pos[1], o.pos[2], o.size[0], o.size[1], o.size[2]});
    for (const Material& o : s->materials) materials.insert(materials.end(), {o.reflection[0], o.reflection[1], o.reflection[2]});
    for (const Light& o : s->lights) lights.insert(lights.end(), {o.emission[0], o.emission[1]});

    memset(ubo, 0, sizeof *ubo);
    ubo->numObjects[0] = (float)nS;
    ubo->numObjects[1] = (float)nP;
    ubo->numObjects[2] = (float)nB;
    ubo->numObjects[3] = (float)nL;
    ubo->numObjects[4] = (float)s->cyclides.size();
    ubo->numObjects[5] = (float)s->sdfs.size();
    ubo->numObjects[6] = (float)lightIDs.size();
    auto fill = [](float* dst, size_t cap, const std::vector<float>& src) { /* zero-padded, silently truncated */
        for (size_t i = 0; i < cap; i++) dst[i] = (src.size() > i) ? src[i] : 0.0f;
    };
    fill(ubo->objects, PT_MAX_OBJECTS_SIZE, objects);
    fill(ubo->sdfs, PT_MAX_SDFS_SIZE, sdfs);
    fill(ubo->materials, PT_MAX_MATERIALS_SIZE, materials);
    fill(ubo->lights, PT_MAX_LIGHTS_SIZE, lights);
    fill(ubo->lightIDs, PT_MAX_LIGHTIDS_SIZE, lightIDs);
    memcpy(ubo->CIEXYZ1931, kCIE, sizeof kCIE); /* CreateUniformBuffer host:2230-2248 */
    return PT_OK;
}

int pt_scene_pack_params(const pt_scene* s, int shot, int width, int height, int samples_per_frame, int path_length,
                         pt_params* p) { /* host:3813-3834 */
    if (!s || !p) { g_error = "pt_scene_pack_params: null argument"; return PT_ERR_ARG; }
    if (shot < 1 || (size_t)shot > s->shots.size()) { g_error = "camera shot index out of range"; return PT_ERR_ARG; }
    if (width <= 0 || height <= 0 || samples_per_frame <= 0 || path_length < 0) {
        g_error = "pt_scene_pack_params: bad render settings";
        return PT_ERR_ARG;
    }
    const CameraShot& c = s->shots[(size_t)shot - 1];
    memset(p, 0, sizeof *p);
    p->resolution[0] = width;
    p->resolution[1] = height;
    p->frame = samples_per_frame;          /* first offscreen dispatch: host:4042-4048 */
    p->currentSamples = samples_per_frame;
    p->samplesPerFrame = samples_per_frame;
    p->FPS = 60.0f;
    p->persistence = 0.0625f;              /* host:1174 */
    p->pathLength = path_length;
    p->cameraAngle[0] = -c.angle[1];
    p->cameraAngle[1] = c.angle[0];
    p->cameraPosX = c.pos[0];
    p->cameraPosY = c.pos[1];
    p->cameraPosZ = c.pos[2];
    p->ISO = s->camera.ISO;
    p->cameraSize = s->camera.size;
    p->apertureSize = s->camera.apertureSize;
    p->apertureDist = s->camera.apertureDist;
    p->lensRadius = s->camera.lensRadius;
    p->lensFocalLength = s->camera.lensFocalLength;
    p->lensThickness = s->camera.lensThickness;
    p->lensDistance = s->camera.lensDistance;
    p->tonemap = 3;                         /* host:35 */
    return PT_OK;
}

} /* extern "C" */
