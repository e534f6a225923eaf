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
 *   * multi-component swizzles: reads `.xzy` -> `.sw3<0,2,1>()`, writes `v.xz = e` / `v.xz *= m` -> `.lsw2<0,2>() = e`
 *     (an lvalue proxy of pt_glsl.h); rgba / stpq spellings are mapped to xyzw
 *   * array constructors `float[3](a, b, c)` / `vec2[](..)` -> braced initialisers
 *   * `#define` / `#undef` / `#if*` lines pass through (every macro a snippet defines is #undef'ed after the snippets, so
 *     it cannot reach the kernel text that follows); `#include`, `#pragma`, `#line`, `#error`, `#extension` are refused
 *   * comments are dropped.
 * Snippet text is otherwise passed through unchanged: mandelbulb / menger / blob / terrain load as shipped.
 *
 * Scene files are data, and in the reference their snippets were sandboxed GLSL.  Here the text ends up in a CUDA C++
 * translation unit inside the host process's context, so everything GLSL does not have is refused before NVRTC sees it:
 * pointers and references (unary `*` / `&`, `->`, `T*`), `::`, casts, `asm`, string / character literals, C++ keywords,
 * and identifiers that reach into CUDA, libc or this library (`__*`, `cuda*`, `atomic*`, `threadIdx`, `printf`, `pt_*` ..).
 */
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "pt_internal.h"

namespace {

bool is_ident_start(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
bool is_ident(char c) { return is_ident_start(c) || (c >= '0' && c <= '9'); }
bool is_digit(char c) { return c >= '0' && c <= '9'; }

bool is_type_name(const std::string& w) {
    static const char* k[] = {"float", "int", "uint", "bool", "vec2", "vec3", "vec4", "mat2", "mat3", "mat4", "void", nullptr};
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

/* words GLSL does not have (or reserves) and identifiers that would reach outside the snippet's sandbox */
bool forbidden_word(const std::string& w) {
    static const char* k[] = {"asm", "reinterpret_cast", "static_cast", "const_cast", "dynamic_cast", "new", "delete", "goto", "extern",
                              "volatile", "typedef", "using", "namespace", "template", "typename", "class", "union", "enum", "sizeof",
                              "alignof", "decltype", "auto", "operator", "this", "throw", "try", "catch", "friend", "virtual", "public",
                              "private", "protected", "mutable", "register", "static", "inline", "constexpr", "nullptr", "char", "short",
                              "long", "double", "unsigned", "signed", "printf", "malloc", "free", "assert", "memcpy", "memset", "threadIdx",
                              "blockIdx", "blockDim", "gridDim", "warpSize", "clock", "clock64", "sdfs_raw", "ptglsl", nullptr};
    for (int i = 0; k[i]; i++)
        if (w == k[i]) return true;
    if (w.size() >= 2 && w[0] == '_' && w[1] == '_') return true;
    static const char* pre[] = {"cuda", "atomic", "pt_", "PT_", "ptk", "nvrtc", nullptr};
    for (int i = 0; pre[i]; i++)
        if (w.compare(0, strlen(pre[i]), pre[i]) == 0) return true;
    return false;
}

/* token-level GLSL -> C++/CUDA rewrite of one snippet; macro names it #defines are appended to *macros */
bool rewrite(const std::string& in, std::string* out, std::string* err, std::vector<std::string>* macros, int* sin_sites) {
    const std::string& s = in;
    std::string o;
    size_t i = 0;
    const size_t n = s.size();
    /* last significant token: 'v' a value (identifier, literal, `)` or `]`), 't' a type name, 'o' anything after which an
     * operand is expected.  A `*` or `&` where an operand is expected, or right after a type name, is a pointer. */
    char last = 'o';
    std::vector<char> closers; /* what each open `(` closes with: `)` or, for an array constructor, `}` */
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
        if (c == '#') { /* preprocessor: #define / #undef / conditionals pass through; the body is rewritten like any text */
            size_t j = i + 1;
            while (j < n && (s[j] == ' ' || s[j] == '\t')) j++;
            size_t e = j;
            while (e < n && is_ident(s[e])) e++;
            const std::string d = s.substr(j, e - j);
            if (!(d == "define" || d == "undef" || d == "if" || d == "ifdef" || d == "ifndef" || d == "else" || d == "elif" || d == "endif")) {
                *err = "preprocessor directive #" + d + " is not supported in SDF snippets";
                return false;
            }
            if (d == "define") {
                const std::string name = peek_ident(s, e);
                if (name.empty() || forbidden_word(name) || is_type_name(name)) { *err = "bad macro name in #define " + name; return false; }
                macros->push_back(name);
            }
            o.push_back('#');
            o += d;
            i = e;
            last = 'o';
            continue;
        }
        if (c == '"' || c == '\'') { *err = "string and character literals are not GLSL"; return false; }
        /* what would let a snippet assemble a refused identifier behind this filter's back: line splicing (and universal
         * character names), the digraph spelling of # / ## (token pasting), and bytes outside GLSL's character set */
        if (c == '\\') { *err = "'\\' (line continuation) is not accepted in SDF snippets"; return false; }
        if (c == '%' && i + 1 < n && s[i + 1] == ':') { *err = "the digraph '%:' is not GLSL"; return false; }
        if (c == '$' || c == '@' || c == '`' || (unsigned char)c >= 0x80 || ((unsigned char)c < 0x20 && c != '\n' && c != '\r' && c != '\t')) {
            *err = "character outside GLSL's character set in SDF snippet";
            return false;
        }
        if (c == ':' && i + 1 < n && s[i + 1] == ':') { *err = "'::' is not GLSL"; return false; }
        if (c == '-' && i + 1 < n && s[i + 1] == '>') { *err = "'->' is not GLSL"; return false; }
        if ((c == '*' || c == '&') && !(i + 1 < n && s[i + 1] == '=') && !(c == '&' && i + 1 < n && s[i + 1] == '&') &&
            !(c == '&' && i > 0 && s[i - 1] == '&')) {
            if (last != 'v') { *err = std::string("unary '") + c + "' (pointers / references) is not GLSL"; return false; }
        }
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
            last = 'v';
            continue;
        }
        if (is_ident_start(c)) {
            size_t j = i;
            while (j < n && is_ident(s[j])) j++;
            const std::string w = s.substr(i, j - i);
            if (w == "highp" || w == "mediump" || w == "lowp") { i = j; continue; }
            if (forbidden_word(w)) { *err = "'" + w + "' is not available to SDF snippets"; return false; }
            if (is_type_name(w)) { /* array constructor: T[n]( .. ) or T[]( .. )  ->  { .. } */
                size_t k = j;
                while (k < n && (s[k] == ' ' || s[k] == '\t')) k++;
                if (k < n && s[k] == '[') {
                    size_t q = k + 1;
                    while (q < n && (is_digit(s[q]) || s[q] == ' ')) q++;
                    if (q < n && s[q] == ']') {
                        size_t r = q + 1;
                        while (r < n && (s[r] == ' ' || s[r] == '\t')) r++;
                        if (r < n && s[r] == '(') {
                            o.push_back('{');
                            closers.push_back('}');
                            i = r + 1;
                            last = 'o';
                            continue;
                        }
                    }
                }
            }
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
            if (w == "sin" && !(i > 0 && s[i - 1] == '.')) { /* the n-th sin( of the unit: PT_SIN_SITE(n, ..) (pt_glsl.h) */
                size_t k = j;
                while (k < n && (s[k] == ' ' || s[k] == '\t')) k++;
                if (k < n && s[k] == '(') {
                    o += "PT_SIN_SITE(" + std::to_string((*sin_sites)++) + ", ";
                    closers.push_back(')');
                    i = k + 1;
                    last = 'o';
                    continue;
                }
            }
            o += w;
            i = j;
            last = is_type_name(w) ? 't' : ((w == "return" || w == "else" || w == "const" || w == "in" || w == "out" || w == "inout") ? 'o' : 'v');
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
                /* `.xz = e`, `.xz += e` ... write through the lvalue proxy; every other use reads a value */
                bool lvalue = false;
                if (k < n) {
                    if (s[k] == '=' && !(k + 1 < n && s[k + 1] == '=')) lvalue = true;
                    if ((s[k] == '+' || s[k] == '-' || s[k] == '*' || s[k] == '/') && k + 1 < n && s[k + 1] == '=') lvalue = true;
                }
                o += lvalue ? ".lsw" : ".sw";
                o += std::to_string(mapped.size());
                o.push_back('<');
                for (size_t q = 0; q < mapped.size(); q++) {
                    if (q) o.push_back(',');
                    o.push_back((char)('0' + (strchr("xyzw", mapped[q]) - "xyzw")));
                }
                o += ">()";
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
            last = 'v';
            continue;
        }
        if (c == '(') { closers.push_back(')'); o.push_back('('); i++; last = 'o'; continue; }
        if (c == ')') {
            const char cl = closers.empty() ? ')' : closers.back();
            if (!closers.empty()) closers.pop_back();
            o.push_back(cl);
            i++;
            last = 'v';
            continue;
        }
        if (c == ']') last = 'v';
        else if (c == '+' && i + 1 < n && s[i + 1] == '+') { o += "++"; i += 2; continue; } /* x++ / ++x leave `last` as it is */
        else if (c == '-' && i + 1 < n && s[i + 1] == '-') { o += "--"; i += 2; continue; }
        else if (c != ' ' && c != '\t' && c != '\n' && c != '\r') last = 'o';
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
        *err = "at most 128 SDFs fit the uniform block's sdfs[768] and the shader's four 32-bit masks (shader.comp:12, 706)";
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
    std::vector<std::string> macros;
    int sin_sites = 0;
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
        if (!rewrite(t, &r, err, &macros, &sin_sites)) { *err = "SDF snippet " + std::to_string(i + 1) + ": " + *err; return PT_ERR_COMPILE; }
        o += "/* ---- snippet " + std::to_string(i + 1) + " ---- */\n";
        o += r;
        o += "\n";
    }
    for (const std::string& m : macros) o += "#undef " + m + "\n"; /* a snippet's macros end with the snippets */
    std::string sdf_lines, mat_lines;
    for (int i = 0; i < n_sdf; i++) {
        const std::string code = std::to_string(1u << (i % 32)) + "u";
        const std::string num = std::to_string(i + 1);
        const std::string pos = "(p - vec3(sdfs[" + std::to_string(6 * i) + "], sdfs[" + std::to_string(6 * i + 1) +
                                "], sdfs[" + std::to_string(6 * i + 2) + "]))";
        /* InsertSDF: "(set" + ((i - i % 32) / 32 + 1) + " & " + 2^(i % 32) + ") == ..." (host:2012, 2029-2033) */
        const std::string cond = "    if ((set" + std::to_string(i / 32 + 1) + " & " + code + ") == " + code + ") ";
        const std::string sdf_line = cond + "sdf = min(sdf, SDF" + num + pos + ");\n";
        const std::string mat_line = cond + "sdfmaterial = minMaterial(sdf, SDF" + num + pos + ", sdfmaterial, SDF" + num +
                                     "MATERIAL" + pos + ");\n";
        sdf_lines += sdf_line;
        mat_lines += mat_line + sdf_line;
    }
    o += "/* shader.comp:706-711 */\nPT_SDF_FN float SDF(vec3 p, uint set1, uint set2, uint set3, uint set4) {\n    float sdf = MAXDIST;\n" + sdf_lines +
         "    return sdf;\n}\n";
    o += "/* shader.comp:713-719 */\nPT_SDF_FN float SDFMATERIAL(vec3 p, uint set1, uint set2, uint set3, uint set4) {\n    float sdf = MAXDIST;\n"
         "    float sdfmaterial = 0.0f;\n" + mat_lines + "    (void)sdf; (void)set1; (void)set2; (void)set3; (void)set4;\n    return sdfmaterial;\n}\n";
    o += "} /* namespace ptglsl */\n";
    o += "PT_SDF_ENTRY float pt_sdf_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4) {\n"
         "    return ptglsl::SDF(ptglsl::vec3(px, py, pz), set1, set2, set3, set4);\n}\n";
    o += "PT_SDF_ENTRY float pt_sdfmaterial_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4) {\n"
         "    return ptglsl::SDFMATERIAL(ptglsl::vec3(px, py, pz), set1, set2, set3, set4);\n}\n";
    *out = o;
    return PT_OK;
}
