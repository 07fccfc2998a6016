// Reader for the drop-in boundary's wire format: the object graph written by
// the reference's Serializer (src/serializer.js:12-60) as JSON text or msgpack
// (tests/test_to_json.js:32-39).  Zero-copy where possible: strings point into
// the caller's blob, containers index into one arena.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace jsrt {

struct Val {
    enum Type : uint8_t { NIL, BOOL, NUM, STR, ARR, MAP };
    Type type = NIL;
    bool b = false;
    // MAP only, filled by WireDoc::indexGraph: which pair holds the serializer's `_t` / `_v` keys (the flattener asks for them
    // tens of millions of times on a 100 000-triangle scene).  kKeyUnknown: not indexed, scan; kKeyAbsent: the map has no such key.
    static constexpr uint8_t kKeyUnknown = 255, kKeyAbsent = 254;
    uint8_t t_at = kKeyUnknown, v_at = kKeyUnknown;
    uint32_t count = 0;      // ARR: elements; MAP: pairs; STR: byte length
    union {                  // by type: 16 bytes per value instead of 32 (a 100 000-triangle scene is ~50 M values)
        double num = 0;      // NUM
        uint64_t first;      // ARR/MAP: arena index of the first child (MAP: key,value,key,value...)
        const char* str;     // STR: points into the caller's blob
    };
};
static_assert(sizeof(Val) == 16, "Val is meant to stay two words");

class WireDoc {
public:
    // format: 0 = JSON, 1 = msgpack (include/jsrt.h JSRT_FORMAT_*)
    WireDoc(const uint8_t* blob, size_t len, int format);

    const Val* root() const { return &root_; }
    // ---- plain value access --------------------------------------------------
    const Val* child(const Val* arr, uint32_t i) const { return &arena_[arr->first + i]; }
    const Val* mapGet(const Val* map, const char* key) const;
    static bool keyEq(const Val* k, const char* s);

    // ---- object-graph access (src/serializer.js:76-130) ------------------------
    // Follows {_r:id} to the defining {_t,_v,_r}; returns v itself if it is not a reference.
    const Val* resolve(const Val* v) const;
    bool isObject(const Val* v) const;               // has _t
    const std::string& typeName(const Val* v) const;  // constructor name of a resolved object
    const Val* payload(const Val* v) const;           // the _v of a resolved object
    // field of an object's _v map, resolved; nullptr if absent
    const Val* field(const Val* obj, const char* key) const;
    // element of an object whose _v is an array (Array / Vec / Mat), resolved
    uint32_t length(const Val* obj) const;
    const Val* at(const Val* obj, uint32_t i) const;

    // numbers: JSON.stringify turns non-finite numbers into null; `nil_value` is
    // what a null stands for at this site (+Inf for IOR / SDF sizes / AABB halves).
    double number(const Val* v, double nil_value) const;
    bool truthy(const Val* v) const;
    // Vec -> up to 4 doubles; returns component count
    int vec(const Val* v, double out[4], double nil_value = 0) const;
    // Mat (4 rows of 4) -> row-major doubles
    void mat4(const Val* v, double out[16]) const;

private:
    std::vector<Val> arena_;
    std::vector<Val> scratch_;     // children of the containers being parsed (one shared stack instead of a vector per container)
    Val root_;
    std::vector<const Val*> ref_table_;                   // _r id -> defining object, for the dense ids the serializer writes
    std::unordered_map<long long, const Val*> refs_;     // ... and for any other id
    std::vector<std::string> typenames_;
    std::string empty_;

    Val parseJson(const char*& p, const char* e, int depth);
    Val parseMsgpack(const uint8_t*& p, const uint8_t* e, int depth);
    uint64_t commit(size_t scratch_start);
    void indexGraph();
    const Val* tOf(const Val* map) const;      // value of the `_t` / `_v` key of a MAP (nullptr if absent), through the index
    const Val* vOf(const Val* map) const;
};

[[noreturn]] inline void fail(const std::string& m) { throw std::runtime_error(m); }

// Recursion guard for walks over the object graph: `_r` references can make a malformed blob cyclic, and a cycle
// must end in an error, not in a stack overflow inside the host process.
struct NestGuard {
    int& depth;
    NestGuard(int& d, int limit) : depth(d) { if (depth >= limit) fail("jsrt: scene graph nested too deep (cyclic reference?)"); ++depth; }
    ~NestGuard() { --depth; }
    NestGuard(const NestGuard&) = delete;
    NestGuard& operator=(const NestGuard&) = delete;
};

}  // namespace jsrt
