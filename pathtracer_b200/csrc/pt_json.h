/* pt_json.h -- a small JSON reader/writer, enough for the reference's scene files.
 * The reference parses scenes with nlohmann::ordered_json 3.11.3 (host:893-908); numbers become double and are
 * narrowed to float/int on assignment (host:2576-2722).  Object member order is preserved, like ordered_json. */
#ifndef PT_JSON_H
#define PT_JSON_H

#include <string>
#include <utility>
#include <vector>

struct PtJson {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    bool is_int = false; /* the literal had no '.', 'e' or 'E' */
    std::string str;
    std::vector<PtJson> arr;
    std::vector<std::pair<std::string, PtJson>> obj;

    const PtJson* find(const char* key) const;
    /* nlohmann's operator[] on a missing key of a const-less object yields null; size() of null is 0 */
    size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
    const PtJson& at(size_t i) const;
    const PtJson& operator[](const char* key) const;
    double number(double dflt = 0.0) const;
    bool truthy() const;
};

/* returns false and fills err ("line:col: message") on malformed input */
bool pt_json_parse(const char* text, size_t len, PtJson* out, std::string* err);
/* 4-space indented dump in member order (SaveScene writes scene.dump(4): host:3465-3472) */
std::string pt_json_dump(const PtJson& v, int indent = 4);

#endif
