// See wire.h.  JSON reader accepts, besides RFC 8259, the tokens Infinity /
// -Infinity / NaN (Python's json module writes them; JSON.stringify writes
// null, handled by WireDoc::number's nil_value).
#include "wire.h"

#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace jsrt {

static const int kMaxDepth = 4096;

// the children of the container just parsed sit on top of the scratch stack, from `scratch_start`: move them to the arena
uint64_t WireDoc::commit(size_t scratch_start) {
    uint64_t first = arena_.size();
    arena_.insert(arena_.end(), scratch_.begin() + (std::ptrdiff_t)scratch_start, scratch_.end());
    scratch_.resize(scratch_start);
    return first;
}

static inline void skipWs(const char*& p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
}
static inline bool lit(const char*& p, const char* e, const char* s) {
    size_t n = strlen(s);
    if ((size_t)(e - p) >= n && !memcmp(p, s, n)) { p += n; return true; }
    return false;
}

Val WireDoc::parseJson(const char*& p, const char* e, int depth) {
    if (depth > kMaxDepth) fail("jsrt: JSON nesting too deep");
    skipWs(p, e);
    if (p >= e) fail("jsrt: unexpected end of JSON");
    Val v;
    char c = *p;
    if (c == '{') {
        ++p; v.type = Val::MAP;
        const size_t start = scratch_.size();
        skipWs(p, e);
        if (p < e && *p == '}') { ++p; v.first = arena_.size(); return v; }
        for (;;) {
            skipWs(p, e);
            if (p >= e || *p != '"') fail("jsrt: JSON object key expected");
            { const Val k = parseJson(p, e, depth + 1); scratch_.push_back(k); }
            skipWs(p, e);
            if (p >= e || *p != ':') fail("jsrt: JSON ':' expected");
            ++p;
            { const Val k = parseJson(p, e, depth + 1); scratch_.push_back(k); }
            skipWs(p, e);
            if (p < e && *p == ',') { ++p; continue; }
            if (p < e && *p == '}') { ++p; break; }
            fail("jsrt: JSON object malformed");
        }
        v.count = (uint32_t)((scratch_.size() - start) / 2);
        v.first = commit(start);
        return v;
    }
    if (c == '[') {
        ++p; v.type = Val::ARR;
        const size_t start = scratch_.size();
        skipWs(p, e);
        if (p < e && *p == ']') { ++p; v.first = arena_.size(); return v; }
        for (;;) {
            { const Val k = parseJson(p, e, depth + 1); scratch_.push_back(k); }
            skipWs(p, e);
            if (p < e && *p == ',') { ++p; continue; }
            if (p < e && *p == ']') { ++p; break; }
            fail("jsrt: JSON array malformed");
        }
        v.count = (uint32_t)(scratch_.size() - start);
        v.first = commit(start);
        return v;
    }
    if (c == '"') {
        ++p; v.type = Val::STR; v.str = p;
        while (p < e && *p != '"') { if (*p == '\\') ++p; ++p; }   // escapes are kept raw; type names and keys have none
        if (p >= e) fail("jsrt: unterminated JSON string");
        v.count = (uint32_t)(p - v.str);
        ++p;
        return v;
    }
    if (lit(p, e, "true")) { v.type = Val::BOOL; v.b = true; return v; }
    if (lit(p, e, "false")) { v.type = Val::BOOL; v.b = false; return v; }
    if (lit(p, e, "null")) { v.type = Val::NIL; return v; }
    v.type = Val::NUM;
    if (lit(p, e, "NaN")) { v.num = std::nan(""); return v; }
    if (lit(p, e, "Infinity")) { v.num = std::numeric_limits<double>::infinity(); return v; }
    if (lit(p, e, "-Infinity")) { v.num = -std::numeric_limits<double>::infinity(); return v; }
    // The blob is (pointer, length), not a C string (a Node Buffer, an mmap): the numeric token is copied, bounded by
    // `e`, into a terminated local buffer before strtod sees it.
    size_t n = 0;
    while (p + n < e && n < 63) {
        const char ch = p[n];
        if (!((ch >= '0' && ch <= '9') || ch == '-' || ch == '+' || ch == '.' || ch == 'e' || ch == 'E')) break;
        ++n;
    }
#if defined(__cpp_lib_to_chars) && __cpp_lib_to_chars >= 201611L
    {   // std::from_chars: bounded by construction, correctly rounded, several times faster than strtod (a 100 000-triangle
        // scene is 10 million numbers)
        const auto r = std::from_chars(p, p + n, v.num);
        if (r.ec == std::errc() && r.ptr != p) { p = r.ptr; return v; }
    }
#endif
    char tok[64];
    memcpy(tok, p, n);
    tok[n] = 0;
    char* end = nullptr;
    v.num = strtod(tok, &end);
    if (end == tok) fail("jsrt: JSON value expected");
    p += end - tok;
    return v;
}

// ---- msgpack ------------------------------------------------------------------
static inline uint64_t be(const uint8_t*& p, const uint8_t* e, int n) {
    if (e - p < n) fail("jsrt: truncated msgpack");
    uint64_t x = 0;
    for (int i = 0; i < n; ++i) x = (x << 8) | *p++;
    return x;
}

Val WireDoc::parseMsgpack(const uint8_t*& p, const uint8_t* e, int depth) {
    if (depth > kMaxDepth) fail("jsrt: msgpack nesting too deep");
    if (p >= e) fail("jsrt: truncated msgpack");
    Val v;
    uint8_t t = *p++;
    auto str = [&](uint64_t n) { if ((uint64_t)(e - p) < n) fail("jsrt: truncated msgpack string"); v.type = Val::STR; v.str = (const char*)p; v.count = (uint32_t)n; p += n; };
    // every element occupies at least one byte: a count beyond the remaining input is malformed (and must not be reserved)
    auto arr = [&](uint64_t n) {
        if (n > (uint64_t)(e - p)) fail("jsrt: truncated msgpack array");
        v.type = Val::ARR; const size_t start = scratch_.size();
        for (uint64_t i = 0; i < n; ++i) { const Val k = parseMsgpack(p, e, depth + 1); scratch_.push_back(k); }
        v.count = (uint32_t)n; v.first = commit(start);
    };
    auto map = [&](uint64_t n) {
        if (2 * n > (uint64_t)(e - p)) fail("jsrt: truncated msgpack map");
        v.type = Val::MAP; const size_t start = scratch_.size();
        for (uint64_t i = 0; i < 2 * n; ++i) { const Val k = parseMsgpack(p, e, depth + 1); scratch_.push_back(k); }
        v.count = (uint32_t)n; v.first = commit(start);
    };
    if (t <= 0x7f) { v.type = Val::NUM; v.num = t; }
    else if (t >= 0xe0) { v.type = Val::NUM; v.num = (int8_t)t; }
    else if (t >= 0xa0 && t <= 0xbf) str(t & 0x1f);
    else if (t >= 0x90 && t <= 0x9f) arr(t & 0x0f);
    else if (t >= 0x80 && t <= 0x8f) map(t & 0x0f);
    else switch (t) {
        case 0xc0: v.type = Val::NIL; break;
        case 0xc2: v.type = Val::BOOL; v.b = false; break;
        case 0xc3: v.type = Val::BOOL; v.b = true; break;
        case 0xca: { uint32_t x = (uint32_t)be(p, e, 4); float f; memcpy(&f, &x, 4); v.type = Val::NUM; v.num = f; break; }
        case 0xcb: { uint64_t x = be(p, e, 8); double d; memcpy(&d, &x, 8); v.type = Val::NUM; v.num = d; break; }
        case 0xcc: v.type = Val::NUM; v.num = (double)be(p, e, 1); break;
        case 0xcd: v.type = Val::NUM; v.num = (double)be(p, e, 2); break;
        case 0xce: v.type = Val::NUM; v.num = (double)be(p, e, 4); break;
        case 0xcf: v.type = Val::NUM; v.num = (double)be(p, e, 8); break;
        case 0xd0: v.type = Val::NUM; v.num = (double)(int8_t)be(p, e, 1); break;
        case 0xd1: v.type = Val::NUM; v.num = (double)(int16_t)be(p, e, 2); break;
        case 0xd2: v.type = Val::NUM; v.num = (double)(int32_t)be(p, e, 4); break;
        case 0xd3: v.type = Val::NUM; v.num = (double)(int64_t)be(p, e, 8); break;
        case 0xd9: str(be(p, e, 1)); break;
        case 0xda: str(be(p, e, 2)); break;
        case 0xdb: str(be(p, e, 4)); break;
        case 0xc4: str(be(p, e, 1)); break;   // bin8/16/32 treated as strings
        case 0xc5: str(be(p, e, 2)); break;
        case 0xc6: str(be(p, e, 4)); break;
        case 0xdc: arr(be(p, e, 2)); break;
        case 0xdd: arr(be(p, e, 4)); break;
        case 0xde: map(be(p, e, 2)); break;
        case 0xdf: map(be(p, e, 4)); break;
        default: fail("jsrt: unsupported msgpack type byte");
    }
    return v;
}

WireDoc::WireDoc(const uint8_t* blob, size_t len, int format) {
    if (!blob || !len) fail("jsrt: empty scene blob");
    // one value takes at least one byte of msgpack and two of JSON; half of that bound avoids all but one regrowth of the arena
    // (a 100 000-triangle scene is ~50 M values)
    arena_.reserve(len / (format == 0 ? 4 : 2) + 16);
    if (format == 0) {
        const char* p = (const char*)blob; const char* e = p + len;
        root_ = parseJson(p, e, 0);
        skipWs(p, e);
        if (p != e) fail("jsrt: trailing bytes after JSON document");
    } else if (format == 1) {
        const uint8_t* p = blob; const uint8_t* e = blob + len;
        root_ = parseMsgpack(p, e, 0);
        if (p != e) fail("jsrt: trailing bytes after msgpack document");
    } else fail("jsrt: unknown scene format (0 = JSON, 1 = msgpack)");
    scratch_ = std::vector<Val>();
    indexGraph();
}

bool WireDoc::keyEq(const Val* k, const char* s) {
    size_t n = strlen(s);
    return k->type == Val::STR && k->count == n && !memcmp(k->str, s, n);
}

const Val* WireDoc::mapGet(const Val* map, const char* key) const {
    if (!map || map->type != Val::MAP) return nullptr;
    for (uint32_t i = 0; i < map->count; ++i)
        if (keyEq(&arena_[map->first + 2 * i], key)) return &arena_[map->first + 2 * i + 1];
    return nullptr;
}

// One pass in document order: register type names (`_t` = [name, idx] on first
// appearance, src/serializer.js:38-41) and referenced objects (`_r` next to `_t`).
void WireDoc::indexGraph() {
    std::vector<Val*> stack{&root_};
    // document order matters only for type names; a DFS that pushes children in
    // reverse keeps it.
    while (!stack.empty()) {
        Val* v = stack.back(); stack.pop_back();
        if (v->type == Val::MAP) {
            // where the serializer's keys sit in this map (first occurrence, like mapGet)
            v->t_at = v->v_at = Val::kKeyAbsent;
            const Val* r = nullptr;
            for (uint32_t i = 0; i < v->count; ++i) {
                const Val* k = &arena_[v->first + 2 * i];
                if (k->type != Val::STR || k->count != 2 || k->str[0] != '_') continue;
                if (k->str[1] == 't') { if (v->t_at == Val::kKeyAbsent) v->t_at = i < Val::kKeyAbsent ? (uint8_t)i : Val::kKeyUnknown; }
                else if (k->str[1] == 'v') { if (v->v_at == Val::kKeyAbsent) v->v_at = i < Val::kKeyAbsent ? (uint8_t)i : Val::kKeyUnknown; }
                else if (k->str[1] == 'r' && !r) r = &arena_[v->first + 2 * i + 1];
            }
            const Val* t = tOf(v);
            if (t) {
                if (t->type == Val::ARR && t->count == 2) {
                    const Val* nm = child(t, 0); const Val* ix = child(t, 1);
                    // `_t` = [className, typeIndex]: the index is a small integer (one per class that occurs in the scene)
                    if (nm->type != Val::STR || ix->type != Val::NUM || !(ix->num >= 0 && ix->num < 65536) || ix->num != (double)(long long)ix->num)
                        fail("jsrt: malformed _t in scene blob (expected [className, typeIndex])");
                    size_t idx = (size_t)ix->num;
                    if (typenames_.size() <= idx) typenames_.resize(idx + 1);
                    typenames_[idx] = std::string(nm->str, nm->count);
                }
                if (r && r->type == Val::NUM && r->num >= -9e15 && r->num <= 9e15) {
                    const long long id = (long long)r->num;
                    // the serializer numbers its objects 1, 2, 3, ... (src/serializer.js:31): a table; anything else: the map
                    if (id >= 0 && id < (1LL << 27)) { if ((size_t)id >= ref_table_.size()) ref_table_.resize((size_t)id + 1 + ref_table_.size() / 2, nullptr); ref_table_[(size_t)id] = v; }
                    else refs_[id] = v;
                }
            }
            for (uint32_t i = v->count; i-- > 0;) stack.push_back(&arena_[v->first + 2 * i + 1]);
        } else if (v->type == Val::ARR) {
            for (uint32_t i = v->count; i-- > 0;) stack.push_back(&arena_[v->first + i]);
        }
    }
}

const Val* WireDoc::tOf(const Val* m) const {
    if (m->t_at < Val::kKeyAbsent) return &arena_[m->first + 2 * (uint64_t)m->t_at + 1];
    return m->t_at == Val::kKeyAbsent ? nullptr : mapGet(m, "_t");
}
const Val* WireDoc::vOf(const Val* m) const {
    if (m->v_at < Val::kKeyAbsent) return &arena_[m->first + 2 * (uint64_t)m->v_at + 1];
    return m->v_at == Val::kKeyAbsent ? nullptr : mapGet(m, "_v");
}

const Val* WireDoc::resolve(const Val* v) const {
    if (!v || v->type != Val::MAP) return v;
    if (tOf(v)) return v;
    // a back-reference is `{_r: id}` alone (src/serializer.js:22-26)
    const Val* r = (v->count == 1 && keyEq(&arena_[v->first], "_r")) ? &arena_[v->first + 1] : mapGet(v, "_r");
    if (!r) return v;
    if (r->type != Val::NUM || !(r->num >= -9e15 && r->num <= 9e15)) fail("jsrt: malformed reference _r in scene blob");
    const long long id = (long long)r->num;
    if (id >= 0 && (size_t)id < ref_table_.size()) {
        if (!ref_table_[(size_t)id]) fail("jsrt: dangling reference _r in scene blob");
        return ref_table_[(size_t)id];
    }
    auto it = refs_.find(id);
    if (it == refs_.end()) fail("jsrt: dangling reference _r in scene blob");
    return it->second;
}

bool WireDoc::isObject(const Val* v) const { return v && v->type == Val::MAP && tOf(v); }

const std::string& WireDoc::typeName(const Val* v) const {
    if (!isObject(v)) return empty_;
    const Val* t = tOf(v);
    const Val* ix = (t->type == Val::ARR && t->count == 2) ? child(t, 1) : t;
    if (ix->type != Val::NUM || !(ix->num >= 0 && ix->num < 65536)) fail("jsrt: malformed _t in scene blob");
    const size_t idx = (size_t)ix->num;
    if (idx >= typenames_.size()) fail("jsrt: type index without a name in scene blob");
    return typenames_[idx];
}

const Val* WireDoc::payload(const Val* v) const { return isObject(v) ? vOf(v) : v; }

const Val* WireDoc::field(const Val* obj, const char* key) const {
    obj = resolve(obj);
    const Val* pv = payload(obj);
    const Val* f = mapGet(pv, key);
    return f ? resolve(f) : nullptr;
}

uint32_t WireDoc::length(const Val* obj) const {
    const Val* pv = payload(resolve(obj));
    return (pv && pv->type == Val::ARR) ? pv->count : 0;
}

const Val* WireDoc::at(const Val* obj, uint32_t i) const {
    const Val* pv = payload(resolve(obj));
    if (!pv || pv->type != Val::ARR || i >= pv->count) fail("jsrt: array index out of range in scene blob");
    return resolve(child(pv, i));
}

double WireDoc::number(const Val* v, double nil_value) const {
    if (!v || v->type == Val::NIL) return nil_value;
    if (v->type == Val::NUM) return v->num;
    if (v->type == Val::BOOL) return v->b ? 1 : 0;
    fail("jsrt: number expected in scene blob");
}

bool WireDoc::truthy(const Val* v) const {
    if (!v) return false;
    switch (v->type) {
        case Val::NIL: return false;
        case Val::BOOL: return v->b;
        case Val::NUM: return v->num != 0 && v->num == v->num;
        case Val::STR: return v->count != 0;
        default: return true;
    }
}

int WireDoc::vec(const Val* v, double out[4], double nil_value) const {
    const Val* pv = payload(resolve(v));
    if (!pv || pv->type != Val::ARR) fail("jsrt: Vec expected in scene blob");
    int n = (int)pv->count;
    for (int i = 0; i < 4; ++i) out[i] = 0;
    for (int i = 0; i < n && i < 4; ++i) out[i] = number(child(pv, i), nil_value);
    return n;
}

void WireDoc::mat4(const Val* v, double out[16]) const {
    const Val* pv = payload(resolve(v));
    if (!pv || pv->type != Val::ARR || pv->count != 4) fail("jsrt: 4x4 Mat expected in scene blob");
    for (int r = 0; r < 4; ++r) {
        const Val* row = payload(resolve(child(pv, r)));
        if (!row || row->type != Val::ARR || row->count != 4) fail("jsrt: 4x4 Mat row expected in scene blob");
        for (int c = 0; c < 4; ++c) out[r * 4 + c] = number(child(row, c), std::numeric_limits<double>::infinity());
    }
}

}  // namespace jsrt
