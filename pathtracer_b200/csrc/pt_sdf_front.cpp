/* pt_sdf_front.cpp -- the SDF plug-in front end: scene["sdf"][i]["glsl"] -> a CUDA translation unit.
 *
 * The reference splices each snippet into shader.comp as text and lets glslang compile the result
 * (InsertSDF, host:2004-2054).  The snippet contract is `float sdf(in vec3 p)` + `float sdfmaterial(in vec3 p)`
 * (plus free helper functions), evaluated in the SDF's translated frame.  This file reproduces what InsertSDF
 * does to the text --
 *   * the first occurrence of the substring "sdf" becomes SDF<i+1>, then the first "sdfmaterial" becomes
 *     SDF<i+1>MATERIAL                                                            (host:2015-2017)
 *   * SDF() / SDFMATERIAL() get one line per snippet, snippet 1 first, and inside SDFMATERIAL() the material
 *     line comes before the distance line                                         (host:2023-2051)
 * -- and adds the token-level rewrite that lets NVRTC (and g++, for the CPU-side tests) compile GLSL snippet
 * text against include/pt_glsl.h:
 *   * float literals get an `f` suffix (GLSL literals are 32-bit; unsuffixed C++ literals would be double)
 *   * parameter qualifiers: `in T x` -> `T x`, `out T x` / `inout T x` -> `T& x`; precision qualifiers vanish
 *   * multi-component swizzles `.xzy` -> `.xzy()` (rgba / stpq spellings are mapped to xyzw)
 *   * comments are dropped.
 * Snippet text is otherwise passed through unchanged: mandelbulb / menger / blob / terrain load as shipped.
 */
#include <stdio.h>
#include <string.h>

#include <string>

#include "pt_internal.h"

namespace {

bool is_ident_start(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
bool is_ident(char c) { return is_ident_start(c) || (c >= '0' && c <= '9'); }
bool is_digit(char c) { return c >= '0' && c <= '9'; }

bool is_type_name(const std::string& w) {
    static const char* k[] = {"float", "int", "uint", "bool", "vec2", "vec3", "vec4", "mat3", nullptr};
    for (int i = 0; k[i]; i++)
        if (w == k[i]) return true;
    return false;
}

/* next identifier starting at or after position i (skipping blanks); empty if the next token is not one */
std::string peek_ident(const std::string& s, size_t i) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n')) i++;
    size_t j = i;
    if (j < s.size() && is_ident_start(s[j])) {
        while (j < s.size() && is_ident(s[j])) j++;
    }
    return s.substr(i, j - i);
}

bool swizzle_letters(const std::string& w, std::string* mapped) {
    if (w.size() < 2 || w.size() > 4) return false;
    const char* sets[3] = {"xyzw", "rgba", "stpq"};
    for (int k = 0; k < 3; k++) {
        std::string m;
        bool ok = true;
        for (char ch : w) {
            const char* p = strchr(sets[k], ch);
            if (!p) { ok = false; break; }
            m.push_back("xyzw"[p - sets[k]]);
        }
        if (ok) { *mapped = m; return true; }
    }
    return false;
}

/* token-level GLSL -> C++/CUDA rewrite of one snippet */
bool rewrite(const std::string& in, std::string* out, std::string* err) {
    const std::string& s = in;
    std::string o;
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
        const char c = s[i];
        if (c == '/' && i + 1 < n && s[i + 1] == '/') { /* line comment */
            while (i < n && s[i] != '\n') i++;
            continue;
        }
        if (c == '/' && i + 1 < n && s[i + 1] == '*') { /* block comment */
            size_t e = s.find("*/", i + 2);
            if (e == std::string::npos) { *err = "unterminated comment in SDF snippet"; return false; }
            for (size_t k = i; k < e + 2; k++)
                if (s[k] == '\n') o.push_back('\n');
            o.push_back(' ');
            i = e + 2;
            continue;
        }
        if (c == '#') { *err = "preprocessor directives are not supported in SDF snippets"; return false; }
        if (is_digit(c) || (c == '.' && i + 1 < n && is_digit(s[i + 1]))) { /* numeric literal */
            size_t j = i;
            bool is_float = false, is_hex = false;
            if (c == '0' && j + 1 < n && (s[j + 1] == 'x' || s[j + 1] == 'X')) {
                is_hex = true;
                j += 2;
                while (j < n && isxdigit((unsigned char)s[j])) j++;
            } else {
                while (j < n && is_digit(s[j])) j++;
                if (j < n && s[j] == '.') { is_float = true; j++; while (j < n && is_digit(s[j])) j++; }
                if (j < n && (s[j] == 'e' || s[j] == 'E')) {
                    size_t k = j + 1;
                    if (k < n && (s[k] == '+' || s[k] == '-')) k++;
                    if (k < n && is_digit(s[k])) {
                        is_float = true;
                        j = k;
                        while (j < n && is_digit(s[j])) j++;
                    }
                }
            }
            o.append(s, i, j - i);
            if (is_float && !is_hex) {
                if (j + 1 < n && (s[j] == 'l' || s[j] == 'L') && (s[j + 1] == 'f' || s[j + 1] == 'F')) {
                    *err = "double-precision literals are not supported in SDF snippets";
                    return false;
                }
                if (j < n && (s[j] == 'f' || s[j] == 'F')) j++; /* already suffixed */
                o.push_back('f');
            } else if (j < n && (s[j] == 'u' || s[j] == 'U')) {
                o.push_back('u');
                j++;
            }
            i = j;
            continue;
        }
        if (is_ident_start(c)) {
            size_t j = i;
            while (j < n && is_ident(s[j])) j++;
            const std::string w = s.substr(i, j - i);
            if (w == "highp" || w == "mediump" || w == "lowp") { i = j; continue; }
            if (w == "in" || w == "out" || w == "inout") {
                const std::string next = peek_ident(s, j);
                if (is_type_name(next)) {
                    if (w == "in") { i = j; continue; } /* by-value, mutable copy: plain C++ parameter */
                    /* out / inout: emit "T&" and skip the type token */
                    size_t k = j;
                    while (k < n && (s[k] == ' ' || s[k] == '\t' || s[k] == '\n')) k++;
                    o += next;
                    o.push_back('&');
                    i = k + next.size();
                    continue;
                }
            }
            o += w;
            i = j;
            continue;
        }
        if (c == '.' && i + 1 < n && is_ident_start(s[i + 1])) { /* member access: maybe a swizzle */
            size_t j = i + 1;
            while (j < n && is_ident(s[j])) j++;
            const std::string w = s.substr(i + 1, j - i - 1);
            std::string mapped;
            size_t k = j;
            while (k < n && (s[k] == ' ' || s[k] == '\t')) k++;
            const bool is_call = (k < n && s[k] == '(');
            if (!is_call && swizzle_letters(w, &mapped)) {
                o.push_back('.');
                o += mapped;
                o += "()";
            } else if (!is_call && w.size() == 1 && strchr("rgbastpq", w[0])) {
                const char* sets[2] = {"rgba", "stpq"};
                char m = w[0];
                for (int q = 0; q < 2; q++) {
                    const char* p = strchr(sets[q], w[0]);
                    if (p) m = "xyzw"[p - sets[q]];
                }
                o.push_back('.');
                o.push_back(m);
            } else {
                o.push_back('.');
                o += w;
            }
            i = j;
            continue;
        }
        o.push_back(c);
        i++;
    }
    *out = o;
    return true;
}

void append_float_literal(float v, std::string* o) {
    char b[48];
    if (v != v) { *o += "(0.0f/0.0f)"; return; }
    if (v - v != 0.0f) { *o += (v > 0.0f) ? "(1.0f/0.0f)" : "(-1.0f/0.0f)"; return; }
    snprintf(b, sizeof b, "%.9g", (double)v); /* 9 significant digits round-trip binary32 exactly */
    *o += b;
    if (!strpbrk(b, ".eE")) *o += ".0";
    *o += "f";
}

/* helper functions of shader.comp that snippets may call: minMaterial (681-683), smin (686-702) */
const char* kHelpers = R"(
PT_SDF_FN float minMaterial(float x, float y, float material1, float material2) {
    return (x < y) ? material1 : material2;
}
PT_SDF_FN float smin(float x, float y) {
    float k = 0.02f;
    k *= 6.0f;
    float h = max(k - abs(x - y), 0.0f) / k;
    float m = h * h * h * 0.5f;
    float s = m * k * ONEBYTHREE;
    return min(x, y) - s;
}
PT_SDF_FN vec2 smin(vec2 x, vec2 y) {
    float k = 0.02f;
    k *= 6.0f;
    float h = max(k - abs(x.x - y.x), 0.0f) / k;
    float m = h * h * h * 0.5f;
    float s = m * k * ONEBYTHREE;
    return (x.x < y.x) ? vec2(x.x - s, x.y + (y.y - x.y) * m) : vec2(y.x - s, x.y + (y.y - x.y) * (1.0f - m));
}
)";

}  // namespace

int pt_sdf_generate(const char* const* sdf_glsl, int n_sdf, const float* sdfs_raw, std::string* out, std::string* err) {
    if (n_sdf < 0 || n_sdf > PT_MAX_SDF_SNIPPETS) {
        *err = "at most 32 SDFs are usable (only set1 of the four masks is ever filled: shader.comp:734-738)";
        return PT_ERR_ARG;
    }
    std::string o;
    o += "/* generated by libpt_cuda (pt_sdf_front.cpp) from the scene's SDF snippets */\n";
    o += "#include \"pt_glsl.h\"\n";
    o += "#if defined(__CUDACC__) || defined(__CUDACC_RTC__)\n";
    o += "#define PT_SDF_FN __device__\n#define PT_SDF_TABLE static __device__ __constant__\n#define PT_SDF_ENTRY __device__\n";
    o += "#else\n#define PT_SDF_FN inline\n#define PT_SDF_TABLE static const\n#define PT_SDF_ENTRY extern \"C\"\n#endif\n";
    o += "namespace ptglsl {\n";
    /* the sdfs[] array of the uniform block (host:3753-3768), baked in: dispatcher lines index it like the
     * reference's generated GLSL does, and snippets may read it */
    const int n_table = n_sdf > 0 ? 6 * n_sdf : 1;
    o += "PT_SDF_TABLE float sdfs[" + std::to_string(n_table) + "] = {";
    for (int i = 0; i < n_table; i++) {
        if (i) o += ", ";
        append_float_literal((sdfs_raw && i < 6 * n_sdf) ? sdfs_raw[i] : 0.0f, &o);
    }
    o += "};\n";
    o += kHelpers;
    for (int i = 0; i < n_sdf; i++) {
        if (!sdf_glsl || !sdf_glsl[i]) { *err = "null SDF snippet"; return PT_ERR_ARG; }
        std::string src(sdf_glsl[i]);
        /* CRLF -> LF (scene files store the snippets with \r\n) */
        std::string t;
        for (size_t k = 0; k < src.size(); k++) {
            if (src[k] == '\r') { if (k + 1 < src.size() && src[k + 1] == '\n') continue; t.push_back('\n'); }
            else t.push_back(src[k]);
        }
        const std::string name = "SDF" + std::to_string(i + 1);
        size_t p = t.find("sdf"); /* host:2015 */
        if (p == std::string::npos) { *err = "SDF snippet " + std::to_string(i + 1) + " defines no sdf()"; return PT_ERR_COMPILE; }
        t.replace(p, 3, name);
        p = t.find("sdfmaterial"); /* host:2017 */
        if (p == std::string::npos) { *err = "SDF snippet " + std::to_string(i + 1) + " defines no sdfmaterial()"; return PT_ERR_COMPILE; }
        t.replace(p, 11, name + "MATERIAL");
        std::string r;
        if (!rewrite(t, &r, err)) return PT_ERR_COMPILE;
        o += "/* ---- snippet " + std::to_string(i + 1) + " ---- */\n";
        o += r;
        o += "\n";
    }
    std::string sdf_lines, mat_lines;
    for (int i = 0; i < n_sdf; i++) {
        const std::string code = std::to_string(1u << (i % 32)) + "u";
        const std::string num = std::to_string(i + 1);
        const std::string pos = "(p - vec3(sdfs[" + std::to_string(6 * i) + "], sdfs[" + std::to_string(6 * i + 1) +
                                "], sdfs[" + std::to_string(6 * i + 2) + "]))";
        const std::string cond = "    if ((set1 & " + code + ") == " + code + ") ";
        const std::string sdf_line = cond + "sdf = min(sdf, SDF" + num + pos + ");\n";
        const std::string mat_line = cond + "sdfmaterial = minMaterial(sdf, SDF" + num + pos + ", sdfmaterial, SDF" + num +
                                     "MATERIAL" + pos + ");\n";
        sdf_lines += sdf_line;
        mat_lines += mat_line + sdf_line;
    }
    o += "/* shader.comp:706-711 */\nPT_SDF_FN float SDF(vec3 p, uint set1) {\n    float sdf = MAXDIST;\n" + sdf_lines +
         "    return sdf;\n}\n";
    o += "/* shader.comp:713-719 */\nPT_SDF_FN float SDFMATERIAL(vec3 p, uint set1) {\n    float sdf = MAXDIST;\n"
         "    float sdfmaterial = 0.0f;\n" + mat_lines + "    (void)sdf;\n    return sdfmaterial;\n}\n";
    o += "} /* namespace ptglsl */\n";
    o += "PT_SDF_ENTRY float pt_sdf_dispatch(float px, float py, float pz, unsigned set1) {\n"
         "    return ptglsl::SDF(ptglsl::vec3(px, py, pz), set1);\n}\n";
    o += "PT_SDF_ENTRY float pt_sdfmaterial_dispatch(float px, float py, float pz, unsigned set1) {\n"
         "    return ptglsl::SDFMATERIAL(ptglsl::vec3(px, py, pz), set1);\n}\n";
    *out = o;
    return PT_OK;
}
