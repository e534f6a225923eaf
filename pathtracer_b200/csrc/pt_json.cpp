/* pt_json.cpp -- recursive-descent JSON (RFC 8259) reader and a pretty printer. */
#include "pt_json.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

const PtJson kNull;

struct Parser {
    const char* s;
    size_t n, i = 0;
    std::string err;

    bool fail(const char* msg) {
        if (err.empty()) {
            size_t line = 1, col = 1;
            for (size_t k = 0; k < i && k < n; k++) {
                if (s[k] == '\n') { line++; col = 1; } else col++;
            }
            char buf[160];
            snprintf(buf, sizeof buf, "%zu:%zu: %s", line, col, msg);
            err = buf;
        }
        return false;
    }
    void ws() { while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
    bool lit(const char* w) {
        size_t l = strlen(w);
        if (i + l <= n && memcmp(s + i, w, l) == 0) { i += l; return true; }
        return false;
    }
    static void utf8(unsigned cp, std::string* o) {
        if (cp < 0x80) o->push_back((char)cp);
        else if (cp < 0x800) { o->push_back((char)(0xC0 | (cp >> 6))); o->push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            o->push_back((char)(0xE0 | (cp >> 12))); o->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            o->push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            o->push_back((char)(0xF0 | (cp >> 18))); o->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            o->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); o->push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    bool hex4(unsigned* v) {
        if (i + 4 > n) return fail("truncated \\u escape");
        unsigned r = 0;
        for (int k = 0; k < 4; k++) {
            char c = s[i++];
            r <<= 4;
            if (c >= '0' && c <= '9') r |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') r |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') r |= (unsigned)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        *v = r;
        return true;
    }
    bool string(std::string* out) {
        if (i >= n || s[i] != '"') return fail("expected string");
        i++;
        out->clear();
        while (i < n) {
            unsigned char c = (unsigned char)s[i++];
            if (c == '"') return true;
            if (c < 0x20) return fail("control character in string");
            if (c != '\\') { out->push_back((char)c); continue; }
            if (i >= n) break;
            char e = s[i++];
            switch (e) {
                case '"': out->push_back('"'); break;
                case '\\': out->push_back('\\'); break;
                case '/': out->push_back('/'); break;
                case 'b': out->push_back('\b'); break;
                case 'f': out->push_back('\f'); break;
                case 'n': out->push_back('\n'); break;
                case 'r': out->push_back('\r'); break;
                case 't': out->push_back('\t'); break;
                case 'u': {
                    unsigned cp;
                    if (!hex4(&cp)) return false;
                    if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < n && s[i] == '\\' && s[i + 1] == 'u') {
                        i += 2;
                        unsigned lo;
                        if (!hex4(&lo)) return false;
                        if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(cp, out);
                    break;
                }
                default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool number(PtJson* out) {
        size_t st = i;
        bool is_int = true;
        if (i < n && s[i] == '-') i++;
        if (i >= n || !(s[i] >= '0' && s[i] <= '9')) return fail("bad number");
        while (i < n && s[i] >= '0' && s[i] <= '9') i++;
        if (i < n && s[i] == '.') {
            is_int = false;
            i++;
            if (i >= n || !(s[i] >= '0' && s[i] <= '9')) return fail("bad fraction");
            while (i < n && s[i] >= '0' && s[i] <= '9') i++;
        }
        if (i < n && (s[i] == 'e' || s[i] == 'E')) {
            is_int = false;
            i++;
            if (i < n && (s[i] == '+' || s[i] == '-')) i++;
            if (i >= n || !(s[i] >= '0' && s[i] <= '9')) return fail("bad exponent");
            while (i < n && s[i] >= '0' && s[i] <= '9') i++;
        }
        std::string tok(s + st, i - st);
        out->type = PtJson::Number;
        out->num = strtod(tok.c_str(), nullptr);
        out->is_int = is_int;
        return true;
    }
    bool value(PtJson* out, int depth) {
        if (depth > 256) return fail("nesting too deep");
        ws();
        if (i >= n) return fail("unexpected end of input");
        char c = s[i];
        if (c == '{') {
            i++;
            out->type = PtJson::Object;
            ws();
            if (i < n && s[i] == '}') { i++; return true; }
            for (;;) {
                ws();
                std::string key;
                if (!string(&key)) return false;
                ws();
                if (i >= n || s[i] != ':') return fail("expected ':'");
                i++;
                out->obj.emplace_back(key, PtJson());
                if (!value(&out->obj.back().second, depth + 1)) return false;
                ws();
                if (i < n && s[i] == ',') { i++; continue; }
                if (i < n && s[i] == '}') { i++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            i++;
            out->type = PtJson::Array;
            ws();
            if (i < n && s[i] == ']') { i++; return true; }
            for (;;) {
                out->arr.emplace_back();
                if (!value(&out->arr.back(), depth + 1)) return false;
                ws();
                if (i < n && s[i] == ',') { i++; continue; }
                if (i < n && s[i] == ']') { i++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { out->type = PtJson::String; return string(&out->str); }
        if (lit("true")) { out->type = PtJson::Bool; out->b = true; return true; }
        if (lit("false")) { out->type = PtJson::Bool; out->b = false; return true; }
        if (lit("null")) { out->type = PtJson::Null; return true; }
        return number(out);
    }
};

void dump_string(const std::string& s, std::string* o) {
    o->push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': *o += "\\\""; break;
            case '\\': *o += "\\\\"; break;
            case '\b': *o += "\\b"; break;
            case '\f': *o += "\\f"; break;
            case '\n': *o += "\\n"; break;
            case '\r': *o += "\\r"; break;
            case '\t': *o += "\\t"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); *o += b; }
                else o->push_back((char)c);
        }
    }
    o->push_back('"');
}

void dump(const PtJson& v, int indent, int level, std::string* o) {
    auto pad = [&](int l) { o->append((size_t)(indent * l), ' '); };
    switch (v.type) {
        case PtJson::Null: *o += "null"; break;
        case PtJson::Bool: *o += v.b ? "true" : "false"; break;
        case PtJson::Number: {
            char b[64];
            if (v.is_int && fabs(v.num) < 9e15) snprintf(b, sizeof b, "%lld", (long long)v.num);
            else {
                snprintf(b, sizeof b, "%.17g", v.num);
                for (int p = 1; p < 17; p++) { /* shortest representation that round-trips */
                    char t[64];
                    snprintf(t, sizeof t, "%.*g", p, v.num);
                    if (strtod(t, nullptr) == v.num) { strcpy(b, t); break; }
                }
                if (!strpbrk(b, ".eEn")) strcat(b, ".0");
            }
            *o += b;
            break;
        }
        case PtJson::String: dump_string(v.str, o); break;
        case PtJson::Array:
            if (v.arr.empty()) { *o += "[]"; break; }
            *o += "[\n";
            for (size_t k = 0; k < v.arr.size(); k++) {
                pad(level + 1);
                dump(v.arr[k], indent, level + 1, o);
                *o += (k + 1 < v.arr.size()) ? ",\n" : "\n";
            }
            pad(level);
            *o += "]";
            break;
        case PtJson::Object:
            if (v.obj.empty()) { *o += "{}"; break; }
            *o += "{\n";
            for (size_t k = 0; k < v.obj.size(); k++) {
                pad(level + 1);
                dump_string(v.obj[k].first, o);
                *o += ": ";
                dump(v.obj[k].second, indent, level + 1, o);
                *o += (k + 1 < v.obj.size()) ? ",\n" : "\n";
            }
            pad(level);
            *o += "}";
            break;
    }
}

}  // namespace

const PtJson* PtJson::find(const char* key) const {
    if (type != Object) return nullptr;
    for (const auto& kv : obj)
        if (kv.first == key) return &kv.second;
    return nullptr;
}
const PtJson& PtJson::at(size_t i) const { return (type == Array && i < arr.size()) ? arr[i] : kNull; }
const PtJson& PtJson::operator[](const char* key) const {
    const PtJson* p = find(key);
    return p ? *p : kNull;
}
double PtJson::number(double dflt) const {
    if (type == Number) return num;
    if (type == Bool) return b ? 1.0 : 0.0;
    return dflt;
}
bool PtJson::truthy() const {
    if (type == Bool) return b;
    if (type == Number) return num != 0.0;
    return false;
}

bool pt_json_parse(const char* text, size_t len, PtJson* out, std::string* err) {
    Parser p{text, len};
    *out = PtJson();
    if (len >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) p.i = 3;
    bool ok = p.value(out, 0);
    if (ok) {
        p.ws();
        if (p.i != p.n) ok = p.fail("trailing characters");
    }
    if (!ok && err) *err = p.err;
    return ok;
}

std::string pt_json_dump(const PtJson& v, int indent) {
    std::string o;
    dump(v, indent, 0, &o);
    return o;
}
