"""oracle/_ref: the reference's OWN sources, compiled for the CPU -- TEST INFRASTRUCTURE, never on the product path.

The reference cannot run as an application here (no Vulkan loader / ICD, no glslang, no GLFW), but
its two source files largely can be compiled against the header-only libraries it vendors itself:

  (a) libref_host.so   <-  src/pathtracer.cpp, the functions of the data contract taken VERBATIM by name
        structs + MAX_* defines (host:39-43, 110-218), CIE table (host:397-842), ReadFile .. sRGBCompanding
        (host:866-1071), App::InsertSDF (host:2004-2054), App::UpdateFromJSON (host:2576-2722), App::UpdateToJSON
        (host:2724-2858), App::UpdateUniformBuffer (host:3642-3811), App::UpdatePushConstant (host:3813-3834) and the
        per-pixel body of App::SaveRender (host:3501-3512), wrapped in a generated `struct App` that declares only the
        members those functions touch, over includes/json/json.hpp and includes/glm.
  (b) ref_shader_<tag>.so  <-  src/shader.comp, mechanically rewritten into one C++ struct over glm (GLM_FORCE_SWIZZLE):
        the text first goes through the reference's own InsertSDF (in libref_host.so, on a CRLF copy -- its offsets
        assume CRLF) with the scene's SDF snippets, then through `translate_shader` below:
        layout blocks -> members, `in/inout/out` -> value / reference parameters, float literals get `f`, swizzles
        become glm swizzle calls, main -> shader_main.  No arithmetic is restated: every expression of the shader is
        compiled as written, evaluated by glm's implementation of the GLSL built-ins and libm.

Nothing is copied into the repository: the extracted / rewritten sources and the shared objects live only in
oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the prebuilt objects travel to the GPU box where /root/reference does
not exist).  `python -m oracle.ref_build` builds everything for the shipped scenes; tests/test_ref_pin.py uses it to pin
oracle/oracle.cpp, oracle/pack.py and the product's packer to the reference itself.

Differences from a Vulkan run, all stated: (1) invocations with gid.x == W or gid.y == H (the shader's `>` bounds test,
shader.comp:1526) are not launched -- their stores fall outside the texel buffer; (2) a fresh texel buffer is
zero-filled; (3) float semantics are glm's + x86-64 SSE2 + glibc libm, compiled -ffp-contract=off: one legal realisation
of GLSL's tolerance-only arithmetic (glm's scalar fma() is std::fma, one rounding -- the vector one is supplied componentwise by the harness; normalize is v*inversesqrt(dot)), not
"the" reference output -- which does not exist.
"""
import hashlib
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, '_ref')
REF = os.environ.get('PT_REFERENCE_DIR', '/root/reference')
CXX = '/usr/bin/g++' if os.access('/usr/bin/g++', os.X_OK) else 'g++'
# -ffp-contract=off: expressions are evaluated as written (no silent fusing); plain x86-64 so the objects run on the
# GPU box's host CPU whatever it is.  -O2: -O3 buys < 5 % here and doubles the compile time.
CXXFLAGS = ['-O2', '-std=c++17', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-w']


def have_reference():
    return os.path.exists(os.path.join(REF, 'src', 'shader.comp')) and \
        os.path.exists(os.path.join(REF, 'includes', 'glm', 'glm.hpp'))


# ------------------------------------------------------------------------------------------------------------------
# (a) host side: verbatim extraction by name
# ------------------------------------------------------------------------------------------------------------------

def _block(lines, start, what):
    """Lines from index `start` up to and including the one where the brace depth returns to zero."""
    depth, seen = 0, False
    for j in range(start, len(lines)):
        for ch in lines[j]:
            if ch == '{':
                depth += 1
                seen = True
            elif ch == '}':
                depth -= 1
        if seen and depth == 0:
            return lines[start:j + 1], j
    raise RuntimeError('ref_build: unterminated block for ' + what)


def _find(lines, pattern, what, begin=0):
    rx = re.compile(pattern)
    for j in range(begin, len(lines)):
        if rx.search(lines[j]):
            return j
    raise RuntimeError('ref_build: %s not found in the reference (pattern %r)' % (what, pattern))


def extract_host(src):
    """Returns (file_scope_text, member_text, save_loop_text, line_map) cut out of src/pathtracer.cpp."""
    lines = src.replace('\r\n', '\n').split('\n')
    where = {}
    file_scope = []
    for name in ('MAX_OBJECTS_SIZE', 'MAX_SDFS_SIZE', 'MAX_MATERIALS_SIZE', 'MAX_LIGHTS_SIZE', 'MAX_LIGHTIDS_SIZE'):
        j = _find(lines, r'^#define %s\b' % name, name)
        file_scope.append(lines[j])
        where[name] = j + 1
    for name in ('MAX_FRAMES_IN_FLIGHT', 'TONEMAP'):
        j = _find(lines, r'^const \w+ %s\b' % name, name)
        file_scope.append(lines[j])
    for name in ('sphere', 'plane', 'box', 'lens', 'cyclide', 'sdf', 'material', 'light', 'Camera', 'CameraShot',
                 'UniformBufferObject', 'PushConstantValues'):
        j = _find(lines, r'^struct %s \{' % name, 'struct ' + name)
        blk, e = _block(lines, j, name)
        file_scope += blk
        where['struct ' + name] = (j + 1, e + 1)
    j = _find(lines, r'^const float CIEXYZ1931\[1323\] = \{', 'CIE table')
    blk, e = _block(lines, j, 'CIE table')
    file_scope += blk
    where['CIEXYZ1931'] = (j + 1, e + 1)
    file_scope.append('std::string computeShaderCode{};')
    a = _find(lines, r'^std::string ReadFile\(', 'ReadFile')
    b = _find(lines, r'^class App \{', 'class App')
    file_scope += lines[a:b]
    where['ReadFile..sRGBCompanding'] = (a + 1, b)
    members = []
    for name in ('InsertSDF', 'UpdateFromJSON', 'UpdateToJSON', 'UpdateUniformBuffer', 'UpdatePushConstant'):
        j = _find(lines, r'^\tvoid %s\(\) \{' % name, 'App::' + name, b)
        blk, e = _block(lines, j, name)
        members += blk + ['']
        where['App::' + name] = (j + 1, e + 1)
    s = _find(lines, r'^\tvoid SaveRender\(\) \{', 'App::SaveRender', b)
    j = _find(lines, r'for \(int i = 0; i < \(W \* H\); i\+\+\) \{', 'SaveRender pixel loop', s)
    blk, e = _block(lines, j, 'SaveRender pixel loop')
    where['App::SaveRender pixel loop'] = (j + 1, e + 1)
    return '\n'.join(file_scope), '\n'.join(members), '\n'.join(blk), where


HOST_HARNESS = r'''// GENERATED by oracle/ref_build.py from %(ref)s/src/pathtracer.cpp -- do not commit (oracle/_ref is git-ignored).
// Everything between the BEGIN/END REFERENCE markers is the reference's text, verbatim; the rest is harness.
#include <glm/glm.hpp>
#include <json/json.hpp>
#include <iostream>
#include <stdexcept>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <array>
#include <algorithm>
#include <fstream>
#include <cmath>
#include <cstdint>
// ---- BEGIN REFERENCE (file scope) ----
%(file_scope)s
// ---- END REFERENCE ----
static_assert(sizeof(UniformBufferObject) == 4097 * 4, "ubo block");
static_assert(sizeof(PushConstantValues) == 88, "push constants");
struct App {
    // the members the extracted functions touch, with the reference's names and defaults (host:1088-1089, 1129, 1146,
    // 1157-1191)
    int W = 1280, H = 720;
    UniformBufferObject ubo;
    std::vector<void*> uniformBuffersMapped;
    PushConstantValues pushConstant;
    bool isUpdateUBO = true;
    int samplesPerFrame = 1, frame = 0, currentSamples = 0;
    float frameTime = 0.0166f;
    int cameraShotIndex = 1;
    float persistence = 0.0625f;
    int pathLength = 5;
    int tonemap = TONEMAP;
    nlohmann::ordered_json scene;
    Camera camera{};
    std::vector<CameraShot> cameraShots;
    std::vector<sphere> spheres;
    std::vector<plane> planes;
    std::vector<box> boxes;
    std::vector<lens> lenses;
    std::vector<cyclide> cyclides;
    std::vector<sdf> sdfs;
    std::vector<material> materials;
    std::vector<light> lights;
    UniformBufferObject mapped[MAX_FRAMES_IN_FLIGHT];
    App() { for (int k = 0; k < MAX_FRAMES_IN_FLIGHT; k++) uniformBuffersMapped.push_back(&mapped[k]); }
// ---- BEGIN REFERENCE (App members) ----
%(members)s
// ---- END REFERENCE ----
    void SaveRenderPixels(const float* pixels, char* pixelsRGB) {
// ---- BEGIN REFERENCE (App::SaveRender, per-pixel loop) ----
%(save_loop)s
// ---- END REFERENCE ----
    }
};

static thread_local std::string g_err;
#define GUARD(body) try { body } catch (const std::exception& e) { g_err = e.what(); return -1; }

extern "C" {
const char* ref_last_error() { return g_err.c_str(); }
void* ref_app_new() { return new App(); }
void ref_app_free(void* a) { delete (App*)a; }
// LoadScene (host:3442-3463) without the dialog: ReadJSON + UpdateFromJSON + UpdateUniformBuffer
int ref_load_scene(void* h, const char* path, int shot) {
    GUARD(
        App* a = (App*)h;
        a->scene = ReadJSON(path);
        a->cameraShotIndex = shot;
        a->UpdateFromJSON();
        a->isUpdateUBO = true;
        std::memcpy(a->ubo.CIEXYZ1931, CIEXYZ1931, sizeof(CIEXYZ1931));  // CreateUniformBuffer, host:2230-2248
        a->UpdateUniformBuffer();
        return 0;
    )
}
int ref_load_scene_text(void* h, const char* text, int shot) {
    GUARD(
        App* a = (App*)h;
        a->scene = nlohmann::ordered_json::parse(text);
        a->cameraShotIndex = shot;
        a->UpdateFromJSON();
        a->isUpdateUBO = true;
        std::memcpy(a->ubo.CIEXYZ1931, CIEXYZ1931, sizeof(CIEXYZ1931));
        a->UpdateUniformBuffer();
        return 0;
    )
}
int ref_get_ubo(void* h, float* out4097) { std::memcpy(out4097, ((App*)h)->uniformBuffersMapped[0], 4097 * 4); return 0; }
int ref_num_sdfs(void* h) { return (int)((App*)h)->sdfs.size(); }
const char* ref_sdf_glsl(void* h, int i) { return ((App*)h)->sdfs[i].glsl.c_str(); }
int ref_num_shots(void* h) { return (int)((App*)h)->cameraShots.size(); }
// UpdatePushConstant for the state the offscreen MainLoop (host:4042-4048) would hold
int ref_get_push(void* h, int W, int H, int frame, int currentSamples, int samplesPerFrame, float frameTime,
                 float persistence, int pathLength, int tonemap, void* out88) {
    App* a = (App*)h;
    a->W = W; a->H = H; a->frame = frame; a->currentSamples = currentSamples; a->samplesPerFrame = samplesPerFrame;
    a->frameTime = frameTime; a->persistence = persistence; a->pathLength = pathLength; a->tonemap = tonemap;
    a->UpdatePushConstant();
    std::memcpy(out88, &a->pushConstant, 88);
    return 0;
}
// InsertSDF on the given shader text; returns the length written (or needed)
long ref_insert_sdf(void* h, const char* shader, char* out, long cap) {
    GUARD(
        computeShaderCode = shader;
        ((App*)h)->InsertSDF();
        long n = (long)computeShaderCode.size();
        if (n < cap) std::memcpy(out, computeShaderCode.c_str(), n + 1);
        return n;
    )
}
// UpdateToJSON: the scene as the reference would save it (dump(4), host:3470)
long ref_to_json(void* h, char* out, long cap) {
    GUARD(
        App* a = (App*)h;
        a->UpdateToJSON();
        std::string s = a->scene.dump(4);
        long n = (long)s.size();
        if (n < cap) std::memcpy(out, s.c_str(), n + 1);
        return n;
    )
}
// SaveRender's per-pixel display transform: pixels = W*H RGBA32F, rgb = W*H*3 bytes
int ref_save_render_pixels(void* h, const float* pixels, int W, int H, int tonemap, char* rgb) {
    App* a = (App*)h;
    a->W = W; a->H = H; a->tonemap = tonemap;
    a->SaveRenderPixels(pixels, rgb);
    return 0;
}
int ref_save_ppm(const char* path, int W, int H, char* rgb) { GUARD( SavePPM(path, W, H, rgb); return 0; ) }
double ref_round_decimal(double x, double p) { return RoundDecimal(x, p); }
float ref_host_spd(float l, float peak, float d, float inv) { return SpectralPowerDistribution(l, peak, d, inv); }
float ref_host_blackbody(float l, float T) { return BlackBodyRadiation(l, T); }
float ref_host_blackbody_peak(float T) { return BlackBodyRadiationPeak(T); }
const float* ref_cie_table() { return CIEXYZ1931; }
}
'''


# ------------------------------------------------------------------------------------------------------------------
# (b) shader side: GLSL -> C++ over glm, mechanically
# ------------------------------------------------------------------------------------------------------------------

_FLOAT = re.compile(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])')
_SWZ_IDX = {'x': 0, 'y': 1, 'z': 2, 'w': 3}


def _strip_comments(s):
    s = re.sub(r'/\*.*?\*/', '', s, flags=re.S)
    return re.sub(r'//[^\n]*', '', s)


def _array_constructors(s):
    """GLSL `float[3](a, b, c)` / `vec2[](u, v)` -> C++ braced initialisers"""
    out, i = [], 0
    rx = re.compile(r'\b(?:float|int|uint|bool|[iub]?vec[234]|mat[234])\s*\[\s*\d*\s*\]\s*\(')
    while True:
        m = rx.search(s, i)
        if not m:
            out.append(s[i:])
            return ''.join(out)
        depth, j = 1, m.end()
        while depth and j < len(s):
            depth += {'(': 1, ')': -1}.get(s[j], 0)
            j += 1
        out.append(s[i:m.start()] + '{' + _array_constructors(s[m.end():j - 1]) + '}')
        i = j


def _swizzle_stores(s):
    """`v.xy = E;` / `v.zw -= E;` -> swz_store<..>(v, E): glm's function-style swizzles are rvalues only."""
    def repl(m):
        var, sw, op, rhs = m.group(1), m.group(2), m.group(3), m.group(4)
        idx = ', '.join(str(_SWZ_IDX[c]) for c in sw)
        if op != '=':
            rhs = '%s.%s %s (%s)' % (var, sw, op[0], rhs)
        return 'swz_store%d<%s>(%s, %s);' % (len(sw), idx, var, rhs)
    return re.sub(r'\b(\w+)\.([xyzw]{2,4})\s*(=|[-+*/]=)(?!=)\s*([^;]+);', repl, s)


def _xyzw_spelling(s):
    """rgba / stpq swizzles spelled as xyzw (glm's function-style swizzles exist for all three; one spelling keeps the
    store rewrite simple)"""
    def repl(m):
        w = m.group(1)
        for alphabet in ('rgba', 'stpq'):
            if all(c in alphabet for c in w):
                return '.' + ''.join('xyzw'[alphabet.index(c)] for c in w)
        return m.group(0)
    return re.sub(r'\.([rgbastpq]{1,4})\b(?!\s*\()', repl, s)


def _split_args(text):
    args, depth, cur = [], 0, ''
    for ch in text:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            args.append(cur)
            cur = ''
        else:
            cur += ch
    args.append(cur)
    return args


def _ordered_constructors(s):
    """GLSL evaluates call arguments left to right; C++ leaves the order open (g++: right to left).  Where two or
    more arguments of one call draw from the RNG (call a function with an inout / out parameter), the order is part of
    the reference's behaviour: constructor calls become braced initialisers, whose order C++ fixes;
    any other such call is refused rather than silently reordered."""
    mutators = set()
    for m in re.finditer(r'\b\w+\s+(\w+)\s*\(([^)]*)\)\s*\{', s):
        if re.search(r'\b(?:inout|out)\b', m.group(2)):
            mutators.add(m.group(1))
    if not mutators:
        return s
    mut = re.compile(r'\b(?:%s)\s*\(' % '|'.join(sorted(mutators)))
    out, i = [], 0
    call = re.compile(r'\b(\w+)\s*\(')
    while True:
        m = call.search(s, i)
        if not m:
            out.append(s[i:])
            break
        name, start = m.group(1), m.end()
        depth, j = 1, start
        while depth and j < len(s):
            depth += {'(': 1, ')': -1}.get(s[j], 0)
            j += 1
        inner = s[start:j - 1]
        if name in ('if', 'for', 'while', 'switch', 'return') or \
                sum(1 for a in _split_args(inner) if mut.search(a)) < 2:
            out.append(s[i:start])
            i = start
            continue
        if not re.fullmatch(r'[iub]?vec[234]|mat[234]', name):
            raise RuntimeError('ref_build: call of %s() has several RNG-drawing arguments; order cannot be kept' % name)
        out.append(s[i:m.start()] + name + '{' + inner + '}')
        i = j
    return ''.join(out)


def translate_shader(glsl):
    """src/shader.comp (after InsertSDF) -> the body of `struct RefShader`.  Purely syntactic; see the module docstring."""
    s = glsl.replace('\r\n', '\n').replace('\r', '\n')
    s = _strip_comments(s)
    s = re.sub(r'^\s*#version[^\n]*\n', '\n', s, flags=re.M)
    s = re.sub(r'^\s*#extension[^\n]*\n', '\n', s, flags=re.M)
    s = re.sub(r'^\s*precision[^\n]*\n', '\n', s, flags=re.M)
    s = re.sub(r'layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;', '', s)
    # uniform block -> pointers into the flat 4097-float block, offsets in declaration order (scalar block layout)
    m = re.search(r'layout\s*\([^)]*std430[^)]*\)\s*uniform\s+ubo\s*\{(.*?)\}\s*;', s, flags=re.S)
    if not m:
        raise RuntimeError('ref_build: uniform block not found in shader.comp')
    defines = dict(re.findall(r'^#define\s+(\w+)\s+(\d+)\s*$', s, flags=re.M))
    fields, off = [], 0
    for typ, name, n in re.findall(r'(\w+)\s+(\w+)\s*\[\s*(\w+)\s*\]\s*;', m.group(1)):
        if typ != 'float':
            raise RuntimeError('ref_build: unexpected uniform block member type ' + typ)
        n = int(defines.get(n, n))
        fields.append((name, off, n))
        off += n
    if off != 4097:
        raise RuntimeError('ref_build: uniform block has %d floats, expected 4097' % off)
    ubo_members = ''.join('    const float* %s;\n' % f[0] for f in fields)
    ubo_init = ''.join('        %s = block + %d;\n' % (f[0], f[1]) for f in fields)
    s = s[:m.start()] + '/*UBO*/' + s[m.end():]
    # texel buffer
    s, n = re.subn(r'layout\s*\([^)]*rgba32f[^)]*\)\s*uniform\s+imageBuffer\s+(\w+)\s*;', r'vec4* \1;', s)
    if n != 1:
        raise RuntimeError('ref_build: imageBuffer declaration not found')
    # push constants -> a base struct (members reachable by their bare names)
    m = re.search(r'layout\s*\(\s*push_constant\s*\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*;', s, flags=re.S)
    if not m:
        raise RuntimeError('ref_build: push-constant block not found')
    push_struct = 'struct RefPush {%s};' % m.group(2)
    s = s[:m.start()] + s[m.end():]
    s = _ordered_constructors(s)
    # parameter qualifiers
    s = re.sub(r'\b(?:inout|out)\s+(\w+)\s+(\w+)', r'\1& \2', s)
    s = re.sub(r'([(,]\s*)(?:const\s+)?in\s+(\w+)\s+(\w+)', r'\1\2 \3', s)
    s = _FLOAT.sub(lambda k: k.group(1) + 'f', s)
    # `v *= M` with M a matrix variable is `v = v * M` in GLSL; glm's vec::operator*= would take M for a scalar
    mats = set(re.findall(r'\bmat[234]\s+(\w+)\s*[=;,)]', s))
    if mats:
        s = re.sub(r'\b([\w.]+)\s*\*=\s*(%s)\s*;' % '|'.join(sorted(mats)), r'\1 = \1 * \2;', s)
    s = _array_constructors(s)
    # rgba / stpq spellings occur in SDF snippets only; the shader's own structs have members named a, b, c, d
    a, b = s.find('vec2 smin('), s.find('float SDF(')
    if 0 <= a < b:
        a = s.index('}', a) + 1
        s = s[:a] + _xyzw_spelling(s[a:b]) + s[b:]
    s = _swizzle_stores(s)
    s = re.sub(r'\.([xyzw]{2,4})\b(?!\s*\()', r'.\1()', s)
    s, n = re.subn(r'\bvoid\s+main\s*\(\s*\)', 'void shader_main()', s)
    if n != 1:
        raise RuntimeError('ref_build: main() not found')
    push_struct = _FLOAT.sub(lambda k: k.group(1) + 'f', push_struct)
    return s, push_struct, ubo_members, ubo_init


SHADER_HARNESS = r'''// GENERATED by oracle/ref_build.py from %(ref)s/src/shader.comp -- do not commit (oracle/_ref is git-ignored).
// The struct body is the reference shader (after the reference's InsertSDF), rewritten syntactically only.
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <cstring>
#include <cstdint>
#include <omp.h>
// glm 0.9.9.7 declares the vector fma() but only defines the scalar one (= std::fma, one rounding): componentwise here
namespace glm {
template <> inline vec2 fma<vec2>(vec2 const& a, vec2 const& b, vec2 const& c) { return vec2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
template <> inline vec3 fma<vec3>(vec3 const& a, vec3 const& b, vec3 const& c) { return vec3(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z)); }
template <> inline vec4 fma<vec4>(vec4 const& a, vec4 const& b, vec4 const& c) { return vec4(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z), std::fma(a.w, b.w, c.w)); }
}
using namespace glm;
// GLSL's implicit int -> float conversions that glm's templates do not perform
inline vec2 operator/(const vec2& a, const ivec2& b) { return a / vec2(b); }
inline vec2 operator-(const vec2& a, const ivec2& b) { return a - vec2(b); }
inline vec2 operator/(const vec2& a, int b) { return a / float(b); }
inline vec3 operator*(int a, const vec3& b) { return float(a) * b; }
inline vec3 operator*(const vec3& a, int b) { return a * float(b); }
inline vec3 operator/(const vec3& a, int b) { return a / float(b); }
inline float mix(float x, float y, int a) { return glm::mix(x, y, float(a)); }
inline vec4 mix(const vec4& x, const vec4& y, int a) { return glm::mix(x, y, float(a)); }
template <int A, int B, class V, class R> inline void swz_store2(V& v, const R& r) { R t = r; v[A] = t[0]; v[B] = t[1]; }
template <int A, int B, int C, class V, class R> inline void swz_store3(V& v, const R& r) { R t = r; v[A] = t[0]; v[B] = t[1]; v[C] = t[2]; }
inline vec4 imageLoad(vec4* b, int i) { return b[i]; }
inline void imageStore(vec4* b, int i, const vec4& v) { b[i] = v; }
%(push_struct)s
static_assert(sizeof(RefPush) == 88, "push constants");
struct RefShader : RefPush {
%(ubo_members)s
    uvec3 gl_GlobalInvocationID;
    // the push-constant base is filled BEFORE the shader's global initialisers (vec3 cameraPos = ...) run
    static RefPush load_push(const void* push) { RefPush p; std::memcpy(&p, push, 88); return p; }
    RefShader(const float* block, const void* push, float* texels, unsigned gx, unsigned gy) : RefPush(load_push(push)) {
%(ubo_init)s
        texelBuffer = reinterpret_cast<vec4*>(texels);
        gl_GlobalInvocationID = uvec3(gx, gy, 0u);
    }
// ---- BEGIN REFERENCE (src/shader.comp, rewritten) ----
%(body)s
// ---- END REFERENCE ----
};

extern "C" {
// one vkCmdDispatch over rows [row0, row1) stepping row_step (the `>` overhang invocations are not launched)
void ref_dispatch(const float* ubo, const void* push, float* texels, int row0, int row1, int row_step, int threads) {
    RefPush pc;
    std::memcpy(&pc, push, 88);
    const int W = pc.resolution.x;
    if (threads <= 0) threads = omp_get_num_procs();
    if (row_step < 1) row_step = 1;
    const int nrows = (row1 - row0 + row_step - 1) / row_step;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int r = 0; r < nrows; r++) {
        const int y = row0 + r * row_step;
        for (int x = 0; x < W; x++) {
            RefShader s(ubo, push, texels, (unsigned)x, (unsigned)y);
            s.shader_main();
        }
    }
}
// leaf entry points: the shader's own functions, called one at a time (KAT-level pins of the oracle)
uint32_t ref_pcg32(uint32_t seed) { float u[4097] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); glm::uint v = seed; s.PCG32(v); return v; }
float ref_random_float(uint32_t* seed) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); glm::uint v = *seed; float f = s.RandomFloatPCG32(v); *seed = v; return f; }
uint32_t ref_generate_seed(const void* push, unsigned x, unsigned y, int k) { float u[1] = {0}; RefShader s(u, push, nullptr, 0, 0); return s.GenerateSeed(uvec2(x, y), k); }
void ref_wave_to_xyz(const float* ubo, float wave, float* out3) { char p[88] = {0}; RefShader s(ubo, p, nullptr, 0, 0); vec3 v = s.WaveToXYZ(wave); out3[0] = v.x; out3[1] = v.y; out3[2] = v.z; }
void ref_sample_wavelengths(float lh, float* out4) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); vec4 v = s.SampleWavelengths(lh); std::memcpy(out4, &v, 16); }
float ref_bk7(float l) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); return s.RefractiveIndexBK7Glass(l); }
void ref_rotation_matrix(const float* angle3, float* out9) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); mat3 m = s.RotationMatrix(vec3(angle3[0], angle3[1], angle3[2])); std::memcpy(out9, &m, 36); }
void ref_emit(const float* l4, float temperature, float luminosity, float* out4) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    RefShader::light lt; lt.emission = vec2(temperature, luminosity);
    vec4 v = s.Emit(vec4(l4[0], l4[1], l4[2], l4[3]), lt); std::memcpy(out4, &v, 16);
}
void ref_spd(const float* l4, float peak, float sigma, int invert, float* out4) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    vec4 v = s.SpectralPowerDistribution(vec4(l4[0], l4[1], l4[2], l4[3]), peak, sigma, invert); std::memcpy(out4, &v, 16);
}
// Intersection(ray) for n rays: out = {t, nx, ny, nz, materialID, lightID} per ray
void ref_intersect(const float* ubo, const void* push, const float* od6, int n, float* out6) {
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        RefShader s(ubo, push, nullptr, 0, 0);
        RefShader::Ray ray; ray.origin = vec3(od6[6*i], od6[6*i+1], od6[6*i+2]); ray.dir = vec3(od6[6*i+3], od6[6*i+4], od6[6*i+5]);
        vec3 nrm(0.0f); float mat = 0.0f, lt = 0.0f;
        float t = s.Intersection(ray, nrm, mat, lt);
        out6[6*i] = t; out6[6*i+1] = nrm.x; out6[6*i+2] = nrm.y; out6[6*i+3] = nrm.z; out6[6*i+4] = mat; out6[6*i+5] = lt;
    }
}
// SDF(p, set1) and SDFMATERIAL(p, set1) of the injected dispatchers for n points
void ref_sdf_eval(const float* ubo, const float* p3, int n, uint32_t set1, float* dist, float* mat) {
    char p[88] = {0};
    for (int i = 0; i < n; i++) {
        RefShader s(ubo, p, nullptr, 0, 0);
        vec3 q(p3[3*i], p3[3*i+1], p3[3*i+2]);
        if (dist) dist[i] = s.SDF(q, set1, 0u, 0u, 0u);
        if (mat) mat[i] = s.SDFMATERIAL(q, set1, 0u, 0u, 0u);
    }
}
// the camera: TracePathLens(l, ray, forwardDir) on one ray
void ref_lens_ray(const float* ubo, const void* push, float l, float* od6, const float* fwd3) {
    RefShader s(ubo, push, nullptr, 0, 0);
    RefShader::Ray ray; ray.origin = vec3(od6[0], od6[1], od6[2]); ray.dir = vec3(od6[3], od6[4], od6[5]);
    s.TracePathLens(l, ray, vec3(fwd3[0], fwd3[1], fwd3[2]));
    od6[0] = ray.origin.x; od6[1] = ray.origin.y; od6[2] = ray.origin.z; od6[3] = ray.dir.x; od6[4] = ray.dir.y; od6[5] = ray.dir.z;
}
int ref_max_threads() { return omp_get_num_procs(); }
void ref_pcg32_n(const uint32_t* in, int n, uint32_t* out) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    for (int i = 0; i < n; i++) { glm::uint v = in[i]; s.PCG32(v); out[i] = v; }
}
// GenerateSeed for n (gid.x, gid.y, k) triples, with xy derived as Rendering() does (shader.comp:1510)
void ref_generate_seed_n(const void* push, const int* gxyk, int n, uint32_t* out) {
    float u[1] = {0}; RefShader s(u, push, nullptr, 0, 0);
    for (int i = 0; i < n; i++)
        out[i] = s.GenerateSeed(uvec2((glm::uint)gxyk[3*i], (glm::uint)(s.resolution.y - gxyk[3*i+1])), gxyk[3*i+2]);
}
// n draws from one PCG stream; same `kind` numbering as oracle_sample
void ref_sample(int kind, uint32_t seed, int n, float param, const float* n3, float* out) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    glm::uint sd = seed;
    const vec3 nn = n3 ? vec3(n3[0], n3[1], n3[2]) : vec3(0.0f, 0.0f, 1.0f);
    for (int i = 0; i < n; i++) {
        vec3 r(0.0f);
        if (kind == 0) { vec2 d = s.SampleUniformUnitDisk(sd); r = vec3(d.x, d.y, 0.0f); }
        else if (kind == 1) r = s.SampleUniformUnitSphere(sd);
        else if (kind == 2) r = s.SampleCosineDirectionHemisphere(nn, sd);
        else if (kind == 3) r = s.SampleCosineUnitCone(sd, param);
        else r = s.ToWorld(s.SampleCosineUnitCone(sd, param), nn);
        out[3*i] = r.x; out[3*i+1] = r.y; out[3*i+2] = r.z;
    }
}
int ref_visible(const float* ubo, const void* push, const float* od6, int lightObjectID) {
    RefShader s(ubo, push, nullptr, 0, 0);
    RefShader::Ray ray; ray.origin = vec3(od6[0], od6[1], od6[2]); ray.dir = vec3(od6[3], od6[4], od6[5]);
    return s.LightSourceVisibilityCheck(ray, lightObjectID) ? 1 : 0;
}
void ref_solve_quartic(const float* coef5, float* roots4, int* real4) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    vec4 r(0.0f);
    bvec4 b = s.SolveQuartic(coef5[0], coef5[1], coef5[2], coef5[3], coef5[4], r);
    for (int i = 0; i < 4; i++) { roots4[i] = r[i]; real4[i] = b[i]; }
}
float ref_cone_pdf(float c, float cmax) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); return s.CosineUnitConePDF(c, cmax); }
float ref_mis_weight(float a, float b) { float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0); return s.MISPowerHeuristicsBeta2(a, b); }
void ref_orthonormal_basis(const float* n3, float* b6) {
    float u[1] = {0}; char p[88] = {0}; RefShader s(u, p, nullptr, 0, 0);
    vec3 b1(0.0f), b2(0.0f);
    s.OrthonormalBasis(b1, b2, vec3(n3[0], n3[1], n3[2]));
    b6[0] = b1.x; b6[1] = b1.y; b6[2] = b1.z; b6[3] = b2.x; b6[4] = b2.y; b6[5] = b2.z;
}
// mean radiance of n TracePath() calls from one ray / wavelength bundle, one PCG stream (as oracle_trace_path)
void ref_trace_path(const float* ubo, const void* push, const float* od6, const float* l4, uint32_t seed, int n, double* mean4) {
    RefShader s(ubo, push, nullptr, 0, 0);
    glm::uint sd = seed;
    double acc[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        RefShader::Ray ray; ray.origin = vec3(od6[0], od6[1], od6[2]); ray.dir = vec3(od6[3], od6[4], od6[5]);
        vec4 r = s.TracePath(vec4(l4[0], l4[1], l4[2], l4[3]), ray, sd);
        for (int k = 0; k < 4; k++) acc[k] += r[k];
    }
    for (int k = 0; k < 4; k++) mean4[k] = acc[k] / (double)(n ? n : 1);
}
// Accumulate(inColor, outColor) for the state in `push`
void ref_accumulate(const void* push, const float* in3, float* out3) {
    float u[1] = {0}; RefShader s(u, push, nullptr, 0, 0);
    vec3 o(out3[0], out3[1], out3[2]);
    s.Accumulate(vec3(in3[0], in3[1], in3[2]), o);
    out3[0] = o.x; out3[1] = o.y; out3[2] = o.z;
}
}
'''


# ------------------------------------------------------------------------------------------------------------------
# build driver
# ------------------------------------------------------------------------------------------------------------------

def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('ref_build: %s failed:\n%s' % (' '.join(cmd[:3]), (r.stdout + r.stderr)[-6000:]))


def _write_if_changed(path, text):
    if os.path.exists(path):
        with open(path) as f:
            if f.read() == text:
                return False
    with open(path, 'w') as f:
        f.write(text)
    return True


def host_so():
    return os.path.join(OUT, 'libref_host.so')


def build_host(force=False):
    """oracle/_ref/libref_host.so from src/pathtracer.cpp (needs /root/reference).  Returns its path."""
    so = host_so()
    if not have_reference():
        if os.path.exists(so):
            return so
        raise RuntimeError('ref_build: %s is absent and %s was not prebuilt' % (REF, so))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(REF, 'src', 'pathtracer.cpp'), encoding='utf-8', errors='replace') as f:
        src = f.read()
    file_scope, members, save_loop, where = extract_host(src)
    text = HOST_HARNESS % dict(ref=REF, file_scope=file_scope, members=members, save_loop=save_loop)
    cpp = os.path.join(OUT, 'ref_host.cpp')
    changed = _write_if_changed(cpp, text)
    with open(os.path.join(OUT, 'ref_host.lines.json'), 'w') as f:
        json.dump(where, f, indent=1)
    if changed or force or not os.path.exists(so):
        tmp = so + '.tmp%d' % os.getpid()
        _run([CXX, *CXXFLAGS, '-I' + os.path.join(REF, 'includes'), '-o', tmp, cpp])
        os.replace(tmp, so)
    return so


def _host_lib():
    import ctypes as C
    L = C.CDLL(build_host())
    L.ref_app_new.restype = C.c_void_p
    L.ref_app_free.argtypes = [C.c_void_p]
    L.ref_load_scene_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.ref_insert_sdf.restype = C.c_long
    L.ref_insert_sdf.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_long]
    L.ref_last_error.restype = C.c_char_p
    return L


def sdf_tag(sdf_sources):
    """Names the shader build of a scene: the SDF snippets are the only per-scene text."""
    h = hashlib.sha1()
    for s in sdf_sources:
        h.update(s.encode() if isinstance(s, str) else s)
        h.update(b'\0')
    return h.hexdigest()[:16] if sdf_sources else 'nosdf'


def shader_so(sdf_sources):
    return os.path.join(OUT, 'ref_shader_%s.so' % sdf_tag(list(sdf_sources)))


def inserted_shader(scene_text):
    """shader.comp after the reference's own InsertSDF for this scene (CRLF in, LF out)."""
    import ctypes as C
    L = _host_lib()
    with open(os.path.join(REF, 'src', 'shader.comp'), 'rb') as f:
        glsl = f.read().decode('utf-8')
    if '\r\n' not in glsl:  # InsertSDF's +26/+34/+31 offsets only land between the lines of a CRLF file (App. C-1)
        glsl = glsl.replace('\n', '\r\n')
    app = L.ref_app_new()
    try:
        if L.ref_load_scene_text(app, scene_text.encode(), 1) != 0:
            raise RuntimeError('ref_build: reference loader rejected the scene: ' + L.ref_last_error().decode())
        cap = len(glsl) * 2 + (1 << 20)
        buf = C.create_string_buffer(cap)
        n = L.ref_insert_sdf(app, glsl.encode(), buf, cap)
        if n < 0 or n >= cap:
            raise RuntimeError('ref_build: InsertSDF failed: ' + L.ref_last_error().decode())
        return buf.value.decode().replace('\r\n', '\n')
    finally:
        L.ref_app_free(app)


def build_shader(scene_path, force=False):
    """oracle/_ref/ref_shader_<tag>.so for one scene file.  Returns its path (prebuilt objects are reused)."""
    with open(scene_path) as f:
        scene_text = f.read()
    sources = [s['glsl'] for s in json.loads(scene_text).get('sdf', [])]
    so = shader_so(sources)
    if not have_reference():
        if os.path.exists(so):
            return so
        raise RuntimeError('ref_build: %s is absent and %s was not prebuilt' % (REF, so))
    os.makedirs(OUT, exist_ok=True)
    body, push_struct, ubo_members, ubo_init = translate_shader(inserted_shader(scene_text))
    text = SHADER_HARNESS % dict(ref=REF, body=body, push_struct=push_struct, ubo_members=ubo_members,
                                 ubo_init=ubo_init)
    cpp = so[:-3] + '.cpp'
    changed = _write_if_changed(cpp, text)
    if changed or force or not os.path.exists(so):
        tmp = so + '.tmp%d' % os.getpid()
        _run([CXX, *CXXFLAGS, '-I' + os.path.join(REF, 'includes'), '-o', tmp, cpp])
        os.replace(tmp, so)
    return so


def build_all(scene_paths=None, jobs=4):
    """Host library + one shader object per distinct SDF set of the given scenes (default: scenes/*.json)."""
    from concurrent.futures import ThreadPoolExecutor
    if scene_paths is None:
        d = os.path.join(ROOT, 'scenes')
        scene_paths = sorted(os.path.join(d, n) for n in os.listdir(d) if n.endswith('.json'))
    build_host()
    seen, todo = set(), []
    for p in scene_paths:
        with open(p) as f:
            tag = sdf_tag([s['glsl'] for s in json.load(f).get('sdf', [])])
        if tag not in seen:
            seen.add(tag)
            todo.append(p)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        return list(ex.map(build_shader, todo))


if __name__ == '__main__':
    if not have_reference():
        print('ref_build: %s not present; nothing to build' % REF)
        sys.exit(0)
    for so in build_all(sys.argv[1:] or None):
        print(so)
