// Native OBJ reader: the geometry half of the reference's parseObjFile (src/objloader.js:149-238), so that large
// meshes load in milliseconds instead of seconds of host-language string handling (the BVH
// build half is bvh_build.cpp).  Same semantics, line by line:
//   * lines are split on '\n'; blank lines and lines whose first non-blank character is '#' are skipped (:155-156,176-177);
//   * tokens are maximal runs of non-whitespace (`l.match(/\S+/g)`);
//   * `mtllib` / `usemtl` record their first argument; material indices follow first appearance (:160-161,185-190);
//   * `f`: every corner token is matched with /(\d+)(?:\/(\d*)(?:\/(\d+))?)?/ — the first run of digits anywhere in the
//     token, so a negative (relative) index loses its sign exactly as it does in the reference — values are 1-based,
//     a missing vt / vn is -1; polygons are fan-triangulated [0, i-1, i] (:144-147,193-196);
//   * every other argument goes through Number.parseFloat: the longest prefix that is a decimal literal or
//     "Infinity", NaN if there is none (:209-210);
//   * `v x y z [w]` (w defaults to 1), `vt u [v [w]]` (missing, NaN and 0 all become 0: `t[2] || 0`), `vn x y z` (w = 0);
//     `s`, `o`, `g`, `vp` are ignored; anything else is the reference's parse error (:213-234).
// Values are stored as f32, like the reference's Vec (src/math.js:160).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jsrt.h"

namespace {

inline bool isSpace(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v' || c == '\n'; }
inline bool isDigit(char c) { return c >= '0' && c <= '9'; }

// Number.parseFloat on one token
double parseFloatJs(const char* p, const char* e) {
    const char* s = p;
    if (s < e && (*s == '+' || *s == '-')) ++s;
    if (e - s >= 8 && !memcmp(s, "Infinity", 8)) return (*p == '-') ? -INFINITY : INFINITY;
    const char* q = s;
    while (q < e && isDigit(*q)) ++q;
    const bool int_digits = q > s;
    bool frac_digits = false;
    if (q < e && *q == '.') {
        const char* r = q + 1;
        while (r < e && isDigit(*r)) ++r;
        frac_digits = r > q + 1;
        if (int_digits || frac_digits) q = r;
    }
    if (!int_digits && !frac_digits) return NAN;
    if (q < e && (*q == 'e' || *q == 'E')) {
        const char* r = q + 1;
        if (r < e && (*r == '+' || *r == '-')) ++r;
        if (r < e && isDigit(*r)) { while (r < e && isDigit(*r)) ++r; q = r; }
    }
    const std::string lit(p, q);          // a plain decimal literal: strtod reads exactly this
    return strtod(lit.c_str(), nullptr);
}

struct Tok { const char* b; const char* e; };

}  // namespace

struct jsrt_obj {
    std::vector<float> positions, texcoords, normals;       // 4 / 3 / 4 floats per entry
    std::vector<int32_t> faces;                             // 9 per triangle: (v, vt, vn) x 3, 0-based, -1 = absent
    std::vector<int32_t> face_material;
    std::vector<std::string> material_names, mtllibs;
    std::string error;
};

extern "C" {

jsrt_obj* jsrt_obj_parse(const char* text, size_t len) {
    jsrt_obj* o = new jsrt_obj;
    if (!text) { o->error = "jsrt: null OBJ text"; return o; }
    const char* p = text; const char* const end = text + len;
    int cur = -1;
    std::vector<Tok> t;
    std::vector<int32_t> idx;
    while (p <= end) {
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!le) le = end;
        // tokenize
        t.clear();
        for (const char* q = p; q < le;) {
            while (q < le && isSpace((unsigned char)*q)) ++q;
            if (q >= le) break;
            const char* b = q;
            while (q < le && !isSpace((unsigned char)*q)) ++q;
            t.push_back(Tok{b, q});
        }
        const char* line_b = p; const char* line_e = le;
        p = le + 1;
        if (t.empty() || *t[0].b == '#') { if (le == end) break; continue; }
        const std::string key(t[0].b, t[0].e);
        auto arg = [&](size_t i) { return i < t.size() ? std::string(t[i].b, t[i].e) : std::string("undefined"); };
        auto num = [&](size_t i) { return i < t.size() ? parseFloatJs(t[i].b, t[i].e) : (double)NAN; };
        if (key == "mtllib") { o->mtllibs.push_back(arg(1)); }
        else if (key == "usemtl") {
            const std::string name = arg(1);
            cur = -1;
            for (size_t i = 0; i < o->material_names.size(); ++i) if (o->material_names[i] == name) { cur = (int)i; break; }
            if (cur < 0) { o->material_names.push_back(name); cur = (int)o->material_names.size() - 1; }
        }
        else if (key == "f") {
            idx.clear();
            for (size_t k = 1; k < t.size(); ++k) {
                const char* q = t[k].b; const char* e = t[k].e;
                while (q < e && !isDigit(*q)) ++q;
                if (q >= e) { o->error = "Error while attempting to parse obj file on line \"" + std::string(line_b, line_e) + "\""; return o; }
                int32_t v[3] = {-1, -1, -1};
                auto digits = [&](int32_t& out) { long long x = 0; bool any = false; while (q < e && isDigit(*q)) { if (x < (1LL << 40)) x = x * 10 + (*q - '0'); ++q; any = true; }
                                                  if (any) out = (int32_t)std::min<long long>(x - 1, 2147483647LL); return any; };
                digits(v[0]);
                if (q < e && *q == '/') {
                    ++q; digits(v[1]);                                  // (\d*): may be empty
                    if (q < e && *q == '/') { const char* save = q; ++q; if (!digits(v[2])) q = save; }      // (\d+): all or nothing
                }
                idx.push_back(v[0]); idx.push_back(v[1]); idx.push_back(v[2]);
            }
            const size_t n = idx.size() / 3;
            for (size_t i = 2; i < n; ++i) {
                const size_t c[3] = {0, i - 1, i};
                for (size_t k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) o->faces.push_back(idx[3 * c[k] + j]);
                o->face_material.push_back(cur);
            }
        }
        else if (key == "v") {
            o->positions.push_back((float)num(1)); o->positions.push_back((float)num(2)); o->positions.push_back((float)num(3));
            o->positions.push_back(t.size() < 5 ? 1.f : (float)num(4));
        }
        else if (key == "vt") {
            auto orZero = [&](size_t i) { const double x = num(i); return (x == x && x != 0) ? (float)x : 0.f; };
            o->texcoords.push_back((float)num(1)); o->texcoords.push_back(orZero(2)); o->texcoords.push_back(orZero(3));
        }
        else if (key == "vn") {
            o->normals.push_back((float)num(1)); o->normals.push_back((float)num(2)); o->normals.push_back((float)num(3)); o->normals.push_back(0.f);
        }
        else if (key == "s" || key == "o" || key == "g" || key == "vp") { /* ignored by the reference */ }
        else { o->error = "Error while attempting to parse obj file on line \"" + std::string(line_b, line_e) + "\""; return o; }
        if (le == end) break;
    }
    return o;
}

const char* jsrt_obj_error(const jsrt_obj* o) { return (o && !o->error.empty()) ? o->error.c_str() : nullptr; }

void jsrt_obj_counts(const jsrt_obj* o, int32_t counts[6]) {
    counts[0] = (int32_t)(o->positions.size() / 4); counts[1] = (int32_t)(o->texcoords.size() / 3); counts[2] = (int32_t)(o->normals.size() / 4);
    counts[3] = (int32_t)o->face_material.size(); counts[4] = (int32_t)o->material_names.size(); counts[5] = (int32_t)o->mtllibs.size();
}

void jsrt_obj_copy(const jsrt_obj* o, float* positions, float* texcoords, float* normals, int32_t* faces, int32_t* face_material) {
    if (positions && !o->positions.empty()) memcpy(positions, o->positions.data(), o->positions.size() * sizeof(float));
    if (texcoords && !o->texcoords.empty()) memcpy(texcoords, o->texcoords.data(), o->texcoords.size() * sizeof(float));
    if (normals && !o->normals.empty()) memcpy(normals, o->normals.data(), o->normals.size() * sizeof(float));
    if (faces && !o->faces.empty()) memcpy(faces, o->faces.data(), o->faces.size() * sizeof(int32_t));
    if (face_material && !o->face_material.empty()) memcpy(face_material, o->face_material.data(), o->face_material.size() * sizeof(int32_t));
}

const char* jsrt_obj_material_name(const jsrt_obj* o, int i) { return (i >= 0 && (size_t)i < o->material_names.size()) ? o->material_names[i].c_str() : nullptr; }
const char* jsrt_obj_mtllib(const jsrt_obj* o, int i) { return (i >= 0 && (size_t)i < o->mtllibs.size()) ? o->mtllibs[i].c_str() : nullptr; }
void jsrt_obj_free(jsrt_obj* o) { delete o; }

}  // extern "C"
