// TEST INFRASTRUCTURE (see oracle_math.h).  Minimal JSON reader for the
// serializer wire format (reference src/serializer.js:12-60), independent of
// the product's reader in jsraytracer_b200/csrc/.  Accepts the non-standard
// tokens Infinity / -Infinity / NaN that Python's json module writes, because
// JSON.stringify would turn the scene's legitimate infinities into null.
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct JV {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
    double num = 0;
    bool b = false;
    std::string s;
    std::vector<JV*> arr;
    std::vector<std::pair<std::string, JV*>> obj;
    const JV* get(const char* k) const {
        for (auto& kv : obj) if (kv.first == k) return kv.second;
        return nullptr;
    }
    bool has(const char* k) const { return get(k) != nullptr; }
};

class JsonParser {
public:
    JsonParser(const char* p, size_t n) : p_(p), e_(p + n) {}
    ~JsonParser() { for (JV* v : pool_) delete v; }
    JV* parse() { ws(); JV* v = value(); ws(); if (p_ != e_) fail("trailing data"); return v; }

private:
    const char *p_, *e_;
    std::vector<JV*> pool_;
    JV* mk(JV::Type t) { JV* v = new JV; v->t = t; pool_.push_back(v); return v; }
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("oracle json: ") + m); }
    void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
    bool lit(const char* s) { size_t n = strlen(s); if ((size_t)(e_ - p_) >= n && !memcmp(p_, s, n)) { p_ += n; return true; } return false; }
    JV* value() {
        if (p_ >= e_) fail("eof");
        char c = *p_;
        if (c == '{') {
            ++p_; JV* v = mk(JV::OBJ); ws();
            if (*p_ == '}') { ++p_; return v; }
            for (;;) {
                ws(); if (*p_ != '"') fail("key");
                std::string k = str(); ws(); if (*p_ != ':') fail("colon"); ++p_; ws();
                v->obj.emplace_back(k, value()); ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; return v; }
                fail("object");
            }
        }
        if (c == '[') {
            ++p_; JV* v = mk(JV::ARR); ws();
            if (*p_ == ']') { ++p_; return v; }
            for (;;) {
                ws(); v->arr.push_back(value()); ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; return v; }
                fail("array");
            }
        }
        if (c == '"') { JV* v = mk(JV::STR); v->s = str(); return v; }
        if (lit("true")) { JV* v = mk(JV::BOOL); v->b = true; return v; }
        if (lit("false")) { JV* v = mk(JV::BOOL); v->b = false; return v; }
        if (lit("null")) return mk(JV::NUL);
        if (lit("NaN")) { JV* v = mk(JV::NUM); v->num = std::nan(""); return v; }
        if (lit("Infinity")) { JV* v = mk(JV::NUM); v->num = INF; return v; }
        if (lit("-Infinity")) { JV* v = mk(JV::NUM); v->num = -INF; return v; }
        char* end = nullptr;
        double d = strtod(p_, &end);
        if (end == p_) fail("value");
        p_ = end;
        JV* v = mk(JV::NUM); v->num = d; return v;
    }
    std::string str() {
        ++p_; std::string out;
        while (p_ < e_ && *p_ != '"') {
            if (*p_ == '\\') { ++p_; char c = *p_++; out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c); }
            else out.push_back(*p_++);
        }
        ++p_; return out;
    }
};

}  // namespace orc
