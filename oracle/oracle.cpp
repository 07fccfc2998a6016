// TEST INFRASTRUCTURE — CPU restatement oracle (see oracle_math.h header).
// PARITY PINNED TO THE REFERENCE ITSELF (see oracle_math.h): bit-identical to runs of the reference's own sources
// (oracle/jsvm) on 34 of its 37 demo scenes (31 as fixtures) — tests/test_refjs_pin.py.
//
// Restates, function by function, the reference's CPU render path:
//   src/renderers.js  src/cameras.js  src/world.js  src/aggregates.js
//   src/geometry.js   src/sdf.js      src/materials.js  src/lights.js
//   src/pixelbuffer.js
// with the same recursion structure, the same traversal order and tie rules,
// and the f32-store / f64-op numeric model.  The scene arrives in the
// serializer wire format (src/serializer.js) as JSON text.
//
// Differences from the reference, all forced: Math.random() is replaced by the
// counter-based RNG of oracle_math.h (the GPU uses the same one); pixels are
// farmed to std::threads (the reference stripes columns over web workers,
// src/worker.js:30-32; results are per-pixel independent either way).
#include "oracle_math.h"
#include "oracle_json.h"

#include <atomic>
#include <functional>
#include <thread>
#include <unordered_map>

namespace orc {

static const double PI = 3.141592653589793;

// ---------------------------------------------------------------------------
// counters (define the roofline's algorithmic work)
enum RayClass { RC_PRIMARY = 0, RC_SECONDARY = 1, RC_SHADOW = 2 };
struct Counters {
    uint64_t rays[3] = {0, 0, 0};        // World.cast calls
    uint64_t bvh_nodes[3] = {0, 0, 0};   // BVHAggregateNode.intersect calls (one AABB slab test each)
    uint64_t bvh_prims[3] = {0, 0, 0};   // leaf object intersect calls
    uint64_t top_tests[3] = {0, 0, 0};   // top-level list object intersect calls
    uint64_t sdf_evals[3] = {0, 0, 0};   // root_sdf.distance() calls while marching
    uint64_t shaded_hits = 0;
    // diagnostic (not part of the reference's algorithm): inner nodes whose box was hit at a depth that is a multiple of
    // 2 / of 3 — the node records a 4-wide / 8-wide collapse of the same tree would fetch for the same ray
    uint64_t wide4[3] = {0, 0, 0}, wide8[3] = {0, 0, 0};
    void add(const Counters& o) {
        for (int i = 0; i < 3; ++i) { wide4[i] += o.wide4[i]; wide8[i] += o.wide8[i]; }
        for (int i = 0; i < 3; ++i) { rays[i] += o.rays[i]; bvh_nodes[i] += o.bvh_nodes[i]; bvh_prims[i] += o.bvh_prims[i]; top_tests[i] += o.top_tests[i]; sdf_evals[i] += o.sdf_evals[i]; }
        shaded_hits += o.shaded_hits;
    }
};

struct Ctx {
    uint32_t sample_key = 0;
    int rc = RC_PRIMARY;
    Counters* c = nullptr;
    // tape mode (orc_render flags bit1): the draws of one pixel sample come from ONE sequential stream in call order,
    // the way Math.random() serves the reference.  With the same stream behind Math.random (oracle/refjs.py) the
    // restatement and the reference's own sources must then agree sample by sample — which also pins the ORDER of draws.
    bool tape = false; mutable uint64_t tape_state = 0;
    double u(uint32_t node, uint32_t dim) const { return tape ? rng_tape_next(tape_state) : rng_u01(rng_node_key(sample_key, node), dim); }
};

// ---------------------------------------------------------------------------
// generic deserialised graph (src/serializer.js:76-130)
struct Node {
    enum Kind { RAW, TYPED_ARR, TYPED_OBJ } kind = RAW;
    const JV* raw = nullptr;
    std::string type;
    std::vector<Node*> arr;
    std::vector<std::pair<std::string, Node*>> fields;
    Node* get(const char* k) const { for (auto& f : fields) if (f.first == k) return f.second; return nullptr; }
    bool isNull() const { return kind == RAW && (!raw || raw->t == JV::NUL); }
    double num() const { if (kind != RAW || !raw || raw->t != JV::NUM) { if (kind == RAW && raw && raw->t == JV::NUL) return INF; throw std::runtime_error("oracle: expected number"); } return raw->num; }
    bool boolean() const { if (kind != RAW || !raw) return true; if (raw->t == JV::BOOL) return raw->b; if (raw->t == JV::NUM) return raw->num != 0; return raw->t != JV::NUL; }
};

struct Deser {
    std::vector<std::unique_ptr<Node>> pool;
    std::unordered_map<long, Node*> refs;
    std::vector<std::string> typenames;
    Node* mk() { pool.emplace_back(new Node); return pool.back().get(); }
    Node* step(const JV* j) {
        if (!j || j->t != JV::OBJ || (!j->has("_r") && !j->has("_t"))) { Node* n = mk(); n->raw = j; return n; }
        if (j->has("_r") && !j->has("_t")) {
            long id = (long)j->get("_r")->num;
            auto it = refs.find(id);
            if (it == refs.end()) throw std::runtime_error("Attempt to deserialize references out of order");
            return it->second;
        }
        const JV* t = j->get("_t");
        std::string name;
        if (t->t == JV::NUM) name = typenames.at((size_t)t->num);
        else { size_t idx = (size_t)t->arr[1]->num; if (typenames.size() <= idx) typenames.resize(idx + 1); name = typenames[idx] = t->arr[0]->s; }
        Node* n = mk();
        n->type = name;
        const JV* v = j->get("_v");
        if (v && v->t == JV::ARR) { n->kind = Node::TYPED_ARR; for (auto* e : v->arr) n->arr.push_back(step(e)); }
        else { n->kind = Node::TYPED_OBJ; if (v) for (auto& kv : v->obj) n->fields.emplace_back(kv.first, step(kv.second)); }
        if (j->has("_r")) refs[(long)j->get("_r")->num] = n;
        return n;
    }
};

static Vec toVec(const Node* n) {
    Vec r;
    if (!n) throw std::runtime_error("oracle: missing Vec");
    if (n->kind == Node::TYPED_ARR) { r.n = (int)n->arr.size(); for (int i = 0; i < r.n && i < 4; ++i) r.v[i] = (float)n->arr[i]->num(); return r; }
    if (n->kind == Node::RAW && n->raw && n->raw->t == JV::ARR) { r.n = (int)n->raw->arr.size(); for (int i = 0; i < r.n && i < 4; ++i) r.v[i] = (float)(n->raw->arr[i]->t == JV::NUL ? INF : n->raw->arr[i]->num); return r; }
    throw std::runtime_error("oracle: expected Vec");
}
static Mat4 toMat(const Node* n) {
    Mat4 m = Mat4::identity();
    if (!n || n->kind != Node::TYPED_ARR || n->arr.size() != 4) throw std::runtime_error("oracle: expected Mat");
    for (int i = 0; i < 4; ++i) {
        const JV* row = n->arr[i]->raw;
        if (!row || row->t != JV::ARR) throw std::runtime_error("oracle: expected Mat row");
        for (int j = 0; j < 4; ++j) m.m[i][j] = row->arr[j]->t == JV::NUL ? INF : row->arr[j]->num;
    }
    return m;
}

// ---------------------------------------------------------------------------
struct World;
struct Primitive;
struct WorldObject;

struct MatData {                      // the `data` object materials pass around
    Ray ray; double distance = 0; Vec position;
    bool hasNormal = false; Vec normal;
    bool hasUV = false; Vec UV;
    bool hasBase = false; Vec basecolor;
    // PhongMaterial.getBaseFactors (src/materials.js:210-238)
    Vec V, N, R; bool backside = false; double vdotn = 0;
    Vec ambient, diffusivity, specularity, reflectivity, transmissivity; double smoothness = 0;
    double kr = 0; bool hasRefr = false; Vec refractionDirection;
};

// Vec.cartesianToSpherical src/math.js:189-193
static Vec cartesianToSpherical(const Vec& n) {
    return Vec::of(0.5 + std::atan2(n[2], n[0]) / (2 * PI), 0.5 - std::asin(n[1]) / PI);
}
// Vec.spherePick src/math.js:180-188 (two draws: theta, phi)
static Vec spherePick(double u0, double u1) {
    const double theta = 2.0 * PI * u0, phi = std::acos(2.0 * u1 - 1.0);
    const double sin_phi = std::sin(phi);
    return Vec::of(std::cos(theta) * sin_phi, std::cos(phi), std::sin(theta) * sin_phi);
}

// ---------------------------------------------------------------------------
// geometry (src/geometry.js)
struct AABBox {
    Vec center, half_size;
    struct TT { bool hit; double min, max; };
    // AABB.get_intersects src/geometry.js:189-209
    TT get_intersects(const Ray& ray, double minDistance, double maxDistance) const {
        double t_min = -INF, t_max = INF;
        const Vec p = center.minus(ray.origin);
        const double epsilon = 0.0000001;
        for (int i = 0; i < 3; ++i) {
            if (std::fabs(ray.direction[i]) > epsilon) {
                double t1 = (p[i] + half_size[i]) / ray.direction[i], t2 = (p[i] - half_size[i]) / ray.direction[i];
                if (t1 > t2) { double tmp = t1; t1 = t2; t2 = tmp; }
                if (t1 > t_min) t_min = t1;
                if (t2 < t_max) t_max = t2;
                if (t_min > t_max || t_max < minDistance || t_min > maxDistance) return {false, 0, 0};
            } else if (std::fabs(p[i]) > half_size[i]) return {false, 0, 0};
        }
        return {true, t_min, t_max};
    }
};

struct Geometry {
    virtual ~Geometry() {}
    virtual double intersect(const Ray& ray, double minDistance, double maxDistance, Ctx& ctx) = 0;
    virtual void materialData(MatData& d, const Vec& direction) = 0;
    virtual Vec sampleSurface(double, double) { throw std::runtime_error("Geometry subclass has not implemented sampleSurface"); }
};

struct AABBGeom : Geometry {          // AABB / UnitBox, src/geometry.js:173-179,210-224
    AABBox box;
    double intersect(const Ray& ray, double minDistance, double maxDistance, Ctx&) override {
        auto t = box.get_intersects(ray, minDistance, maxDistance);
        if (t.hit) return (t.min >= minDistance) ? t.min : t.max;
        return -INF;
    }
    void materialData(MatData& d, const Vec&) override {
        const Vec& p = d.position;
        double norm_dist = 0; Vec norm = Vec::of(0, 0, 0, 0);
        for (int i = 0; i < 3; ++i) {
            const double comp = (p[i] - box.center[i]) / box.half_size[i], abs_comp = std::fabs(comp);
            if (abs_comp > norm_dist) { norm_dist = abs_comp; norm = Vec::of(0, 0, 0, 0); norm.v[i] = (float)js_sign(comp); }
        }
        d.hasNormal = true; d.normal = norm;
    }
};

struct SimplePlane : Geometry {       // src/geometry.js:239-255
    double planeT(const Ray& ray) const { return (ray.direction[2] != 0) ? -ray.origin[2] / ray.direction[2] : -INF; }
    double intersect(const Ray& ray, double, double, Ctx&) override { return planeT(ray); }
    void materialData(MatData& d, const Vec&) override {
        d.hasNormal = true; d.normal = Vec::of(0, 0, 1, 0);
        d.hasUV = true; d.UV = Vec::of(d.position[0], d.position[1]);
    }
};
struct Square : SimplePlane {         // src/geometry.js:280-301
    double intersect(const Ray& ray, double, double, Ctx&) override {
        const double t = planeT(ray);
        const Vec p = ray.getPoint(t);
        return (-0.5 <= p[0] && p[0] <= 0.5 && -0.5 <= p[1] && p[1] <= 0.5) ? t : -INF;
    }
    Vec sampleSurface(double u0, double u1) override { return Vec::of(u0 - 0.5, u1 - 0.5, 0, 1); }
};
struct Circle : SimplePlane {         // src/geometry.js:303-332
    double intersect(const Ray& ray, double, double, Ctx&) override {
        const double t = planeT(ray);
        const Vec p = ray.getPoint(t);
        return (p.minus(Vec::of(0, 0, 0, 1)).squarednorm() <= 1) ? t : -INF;
    }
    Vec sampleSurface(double u0, double u1) override { return Vec::of(u0 - 0.5, u1 - 0.5, 0, 1); }
};

struct Triangle : Geometry {          // src/geometry.js:334-410
    Vec ps[3]; Vec v0, v1, normal; double delta, d00, d11, d01, denom, area;
    bool hasNormals = false, hasUVs = false; Vec vnormal[3], vuv[3];
    void init() {
        v0 = ps[1].minus(ps[0]).to3();
        v1 = ps[2].minus(ps[0]).to3();
        const Vec heron = v0.cross(v1);
        area = heron.norm() / 2.0;
        normal = heron.normalized().to4(false);
        delta = normal.dot(ps[0]);
        d00 = v0.squarednorm(); d11 = v1.squarednorm(); d01 = v0.dot(v1);
        denom = d00 * d11 - d01 * d01;
    }
    Vec toBarycentric(const Vec& p) const {
        const Vec v2 = p.minus(ps[0]).to3();
        const double d20 = v2.dot(v0), d21 = v2.dot(v1),
                     v = (d11 * d20 - d01 * d21) / denom, w = (d00 * d21 - d01 * d20) / denom;
        return Vec::of(1 - v - w, v, w);
    }
    double intersect(const Ray& ray, double, double, Ctx&) override {
        const double den = normal.dot(ray.direction);
        const double distance = (den != 0) ? (delta - normal.dot(ray.origin)) / den : -INF;
        if (!std::isfinite(distance) || distance < 0) return distance;
        const Vec bary = toBarycentric(ray.getPoint(distance).to3());
        for (int i = 0; i < 3; ++i) if (!(bary[i] >= 0 && bary[i] <= 1)) return -INF;
        return distance;
    }
    static Vec blend(const Vec& bary, const Vec* data) {   // src/geometry.js:397-409
        return data[0].times(bary[0]).plus(data[1].times(bary[1])).plus(data[2].times(bary[2]));
    }
    void materialData(MatData& d, const Vec&) override {
        d.hasNormal = true; d.normal = normal;
        const Vec bary = toBarycentric(d.position);
        // `for (k in psdata)`: insertion order of the OBJ loader is UV then normal
        if (hasUVs) { d.hasUV = true; d.UV = blend(bary, vuv); }
        if (hasNormals) d.normal = blend(bary, vnormal);
    }
};

static double sphereStaticIntersect(const Ray& r, double minDistance) {   // src/geometry.js:429-442
    const double a = r.direction.squarednorm(), b = r.direction.dot(r.origin), c = r.origin.to3().squarednorm() - 1;
    double big = b * b - a * c;
    if (big < 0 || a == 0) return -INF;
    big = std::sqrt(big);
    const double t1 = (-b + big) / a, t2 = (-b - big) / a;
    if (t1 >= minDistance && t2 >= minDistance) return js_min(t1, t2);
    return (t2 < minDistance) ? t1 : t2;
}
struct Sphere : Geometry {            // src/geometry.js:412-456
    double intersect(const Ray& r, double minDistance, double, Ctx&) override { return sphereStaticIntersect(r, minDistance); }
    Vec sampleSurface(double u0, double u1) override { return spherePick(u0, u1).to4(true); }
    void materialData(MatData& d, const Vec&) override {
        const Vec n = d.position.normalized();     // sic: 4-vector with w = 1
        d.hasNormal = true; d.normal = n;
        d.hasUV = true; d.UV = cartesianToSpherical(n);
    }
};
struct Cylinder : Geometry {          // src/geometry.js:458-488
    double intersect(const Ray& r, double minDistance, double, Ctx&) override {
        if (std::fabs(r.origin[2]) > 1 && r.direction[2] != 0)
            minDistance = js_max(minDistance, -(r.origin[2] - js_sign(r.origin[2])) / r.direction[2]);
        const Vec mask = Vec::of(1, 1, 0, 1);
        const double t = sphereStaticIntersect(Ray(mask.times(r.origin), mask.times(r.direction)), minDistance);
        return (std::fabs(r.origin[2] + t * r.direction[2]) <= 1) ? t : -INF;
    }
    void materialData(MatData& d, const Vec&) override {
        d.hasNormal = true; d.normal = Vec::of(d.position[0], d.position[1], 0, 0).normalized();
        d.hasUV = true; d.UV = Vec::of(0.5 + std::atan2(d.position[1], d.position[0]) / (2 * PI), 0.5 + d.position[2]);
    }
};

// ---------------------------------------------------------------------------
// SDF tree (src/sdf.js)
struct SdfMat { bool hasBase = false; Vec base; bool hasUV = false; Vec UV; };
static SdfMat blendMaterialData(double mix, const SdfMat& a, const SdfMat& b) {   // src/sdf.js:66-73
    if (mix <= 0.0) return a;
    if (mix >= 1.0) return b;
    SdfMat r; r.hasBase = true; r.hasUV = true;
    r.base = (a.hasBase ? a.base : Vec::of(1, 1, 1)).mix(b.hasBase ? b.base : Vec::of(1, 1, 1), mix);
    r.UV = (a.hasUV ? a.UV : Vec::of(0, 0)).mix(b.hasUV ? b.UV : Vec::of(0, 0), mix);
    return r;
}
struct SDFTransformer { virtual ~SDFTransformer() {} virtual void transform(Vec& p, double& s) const = 0; };
struct SDFNode {
    virtual ~SDFNode() {}
    virtual double distance(const Vec& p) const = 0;
    virtual SdfMat getMaterialData(const Vec& p) const = 0;
};
static double smoothMin(double a, double b, double k) { const double h = js_max(k - std::fabs(a - b), 0.0) / k; return js_min(a, b) - h * h * h * k * (1.0 / 6.0); }
static double smoothMinBlend(double a, double b, double k) { const double h = js_max(k - std::fabs(a - b), 0.0) / k; const double m = h * h * h * 0.5; return (a < b) ? m : (1.0 - m); }

struct UnionSDF : SDFNode {
    std::vector<SDFNode*> children;
    double distance(const Vec& p) const override { double r = INF; for (auto* c : children) r = js_min(r, c->distance(p)); return r; }
    SdfMat getMaterialData(const Vec& p) const override {
        double mn = INF; int idx = -1;
        for (size_t i = 0; i < children.size(); ++i) { double d = children[i]->distance(p); if (d < mn) { mn = d; idx = (int)i; } }
        if (idx < 0) throw std::runtime_error("UnionSDF.getMaterialData: no finite child distance");
        return children[idx]->getMaterialData(p);
    }
};
struct IntersectionSDF : SDFNode {
    std::vector<SDFNode*> children;
    double distance(const Vec& p) const override { double r = -INF; for (auto* c : children) r = js_max(r, c->distance(p)); return r; }
    SdfMat getMaterialData(const Vec& p) const override {
        double mx = -INF; int idx = -1;
        for (size_t i = 0; i < children.size(); ++i) { double d = children[i]->distance(p); if (d > mx) { mx = d; idx = (int)i; } }
        if (idx < 0) throw std::runtime_error("IntersectionSDF.getMaterialData: no finite child distance");
        return children[idx]->getMaterialData(p);
    }
};
struct DifferenceSDF : SDFNode {
    SDFNode *positive, *negative;
    double distance(const Vec& p) const override { return js_max(positive->distance(p), -negative->distance(p)); }
    SdfMat getMaterialData(const Vec& p) const override {
        return (positive->distance(p) > -negative->distance(p)) ? positive->getMaterialData(p) : negative->getMaterialData(p);
    }
};
struct SmoothUnionSDF : SDFNode {
    SDFNode *a, *b; double k;
    double distance(const Vec& p) const override { return smoothMin(a->distance(p), b->distance(p), k); }
    SdfMat getMaterialData(const Vec& p) const override { return blendMaterialData(smoothMinBlend(a->distance(p), b->distance(p), k), a->getMaterialData(p), b->getMaterialData(p)); }
};
struct SmoothIntersectionSDF : SDFNode {
    SDFNode *a, *b; double k;
    double distance(const Vec& p) const override { return -smoothMin(-a->distance(p), -b->distance(p), k); }
    SdfMat getMaterialData(const Vec& p) const override { return blendMaterialData(1.0 - smoothMinBlend(-a->distance(p), -b->distance(p), k), a->getMaterialData(p), b->getMaterialData(p)); }
};
struct SmoothDifferenceSDF : SDFNode {
    SDFNode *positive, *negative; double k;
    double distance(const Vec& p) const override { return -smoothMin(-positive->distance(p), negative->distance(p), k); }
    SdfMat getMaterialData(const Vec& p) const override { return blendMaterialData(smoothMinBlend(-positive->distance(p), negative->distance(p), k), positive->getMaterialData(p), negative->getMaterialData(p)); }
};
struct RoundSDF : SDFNode {
    SDFNode* child; double rounding;
    double distance(const Vec& p) const override { return child->distance(p) - rounding; }
    SdfMat getMaterialData(const Vec& p) const override { return child->getMaterialData(p); }
};
struct SphereSDF : SDFNode {
    double radius; Vec basecolor;
    double distance(const Vec& p) const override { return p.to4(false).norm() - radius; }
    SdfMat getMaterialData(const Vec& p) const override { SdfMat m; m.hasBase = true; m.base = basecolor; m.hasUV = true; m.UV = cartesianToSpherical(p.to4(false).normalized()); return m; }
};
struct BoxSDF : SDFNode {
    Vec size, basecolor;
    double distance(const Vec& p) const override {     // BoxSDF.distanceComp src/sdf.js:276-279
        const Vec q = p.abs().minus(size).to4(false);
        return Vec::maxs(q, 0).norm() + js_min(js_max(js_max(q[0], q[1]), q[2]), 0);
    }
    SdfMat getMaterialData(const Vec&) const override { SdfMat m; m.hasBase = true; m.base = basecolor; return m; }
};
struct TetrahedronSDF : SDFNode {
    Vec basecolor;
    double distance(const Vec& p) const override {
        return (js_max(std::fabs(p[0] + p[1]) - p[2], std::fabs(p[0] - p[1]) + p[2]) - 1) / std::sqrt(3.0);
    }
    SdfMat getMaterialData(const Vec&) const override { SdfMat m; m.hasBase = true; m.base = basecolor; return m; }
};
struct TransformSDF : SDFNode {
    SDFNode* child; SDFTransformer* transformer;
    double distance(const Vec& p) const override { Vec pt = p; double s = 1; transformer->transform(pt, s); return child->distance(pt) * s; }
    SdfMat getMaterialData(const Vec& p) const override { return child->getMaterialData(p); }   // sic: untransformed p
};
struct RecursiveTransformUnionSDF : SDFNode {
    SDFNode* sdf; SDFTransformer* transformer; int iterations;
    double distance(const Vec& p0) const override {
        Vec p = p0; double bestDist = sdf->distance(p), s = 1;
        for (int i = 0; i < iterations; ++i) {
            double st = 1; transformer->transform(p, st);
            s = s * st;
            bestDist = js_min(sdf->distance(p) * s, bestDist);
        }
        return bestDist;
    }
    SdfMat getMaterialData(const Vec& p) const override { return sdf->getMaterialData(p); }
};
// transformers: transform(p) returns [pt, scale]; here p is updated in place and s receives the factor
struct SDFTransformerSequence : SDFTransformer {
    std::vector<SDFTransformer*> ts;
    void transform(Vec& p, double& sout) const override { double s = 1; for (auto* t : ts) { double st = 1; t->transform(p, st); s = s * st; } sout = s; }
};
struct SDFRecursiveTransformer : SDFTransformer {
    SDFTransformer* t; int iterations;
    void transform(Vec& p, double& sout) const override { double s = 1; for (int i = 0; i < iterations; ++i) { double st = 1; t->transform(p, st); s = s * st; } sout = s; }
};
struct SDFMatrixTransformer : SDFTransformer {
    Mat4 inv; double scale;
    void transform(Vec& p, double& s) const override { p = inv.times(p); s = scale; }
};
struct SDFReflectionTransformer : SDFTransformer {
    Vec normal; double delta;
    void transform(Vec& p, double& s) const override {
        const double dot = normal.dot(p) - delta;
        if (dot < 0) p = p.minus(normal.times(2 * dot));
        s = 1;
    }
};
struct SDFInfiniteRepetitionTransformer : SDFTransformer {
    Vec sizes;
    void transform(Vec& p, double& s) const override {
        double c[3];
        for (int i = 0; i < 3; ++i) c[i] = js_fmod(p[i] + sizes[i] / 2, sizes[i]) - sizes[i] / 2;
        p = Vec::of(c[0], c[1], c[2]).to4(true);
        s = 1;
    }
};

struct SDFGeometry : Geometry {       // src/sdf.js:1-51
    SDFNode* root; AABBox aabb; double max_samples, distance_epsilon, max_trace_distance, normal_step_size;
    double intersect(const Ray& ray, double minDistance, double maxDistance, Ctx& ctx) override {
        auto ib = aabb.get_intersects(ray, minDistance, maxDistance);
        if (!ib.hit) return -INF;
        minDistance = js_max(minDistance, ib.min);
        maxDistance = js_min(maxDistance, ib.max);
        double t = minDistance;
        const double rd_norm = ray.direction.norm();
        for (double i = 0; i < max_samples; ++i) {
            const Vec p = ray.getPoint(t);
            const double distance = root->distance(p);
            if (ctx.c) ctx.c->sdf_evals[ctx.rc]++;
            if (!std::isfinite(distance)) break;
            if (distance <= distance_epsilon) return t;
            t += distance / rd_norm;
            if (t < minDistance || t > maxDistance || (t - minDistance) * rd_norm > max_trace_distance) break;
        }
        return -INF;
    }
    void materialData(MatData& d, const Vec&) override {
        const double distance = root->distance(d.position);
        Vec N = Vec::of(0, 0, 0, 0);
        for (int i = 0; i < 3; ++i)
            N.v[i] = (float)((root->distance(d.position.plus(Vec::axis(i, 4, normal_step_size))) - distance) / normal_step_size);
        SdfMat m = root->getMaterialData(d.position);
        if (m.hasBase) { d.hasBase = true; d.basecolor = m.base; }
        if (m.hasUV) { d.hasUV = true; d.UV = m.UV; }
        d.hasNormal = true; d.normal = N.normalized();
    }
};

// ---------------------------------------------------------------------------
// material colours (src/materials.js:2-131)
struct MaterialColor { virtual ~MaterialColor() {} virtual Vec color(const MatData& d) const = 0; };
struct SolidMaterialColor : MaterialColor { Vec c; Vec color(const MatData&) const override { return c; } };
struct ScaledMaterialColor : MaterialColor {
    MaterialColor* mc; bool isArray = false; double s = 1; Vec sv;
    Vec color(const MatData& d) const override { return isArray ? mc->color(d).times(sv) : mc->color(d).times(s); }
};
struct CheckerboardMaterialColor : MaterialColor {
    MaterialColor *c1, *c2;
    Vec color(const MatData& d) const override {
        const double u = d.hasUV ? d.UV[0] : 0, v = d.hasUV ? d.UV[1] : 0;
        return (std::fmod(js_fmod(std::floor(u) + std::floor(v), 2), 2) < 1) ? c1->color(d) : c2->color(d);
    }
};

// TextureMaterialColor.color src/materials.js:101-130 (RGBA8 ImageData; bilinear or nearest; V flipped)
struct TextureMaterialColor : MaterialColor {
    int width = 0, height = 0; bool bilinear = true, clampU = true, clampV = true;
    std::vector<unsigned char> data;
    static double normalizeUV(double c, bool doClamp) {           // :97-100; JS `%` is fmod
        if (doClamp) return js_min(js_max(c, 0), 1);
        return std::fmod(std::fmod(c, 1) + 1, 1);
    }
    Vec color(const MatData& d) const override {
        const double U = normalizeUV(d.hasUV ? d.UV[0] : 0, clampU), V = normalizeUV(d.hasUV ? d.UV[1] : 0, clampV);
        const double fx = U * width - 0.5, fy = (1.0 - V) * height - 0.5;
        double xs[2][2], ys[2][2]; int nx = 0, ny = 0;
        if (bilinear) {
            const double mx = std::floor(fx), my = std::floor(fy);
            xs[0][0] = mx; xs[0][1] = 1 - std::fmod(fx, 1); xs[1][0] = mx + 1; xs[1][1] = std::fmod(fx, 1); nx = 2;
            ys[0][0] = my; ys[0][1] = 1 - std::fmod(fy, 1); ys[1][0] = my + 1; ys[1][1] = std::fmod(fy, 1); ny = 2;
        } else {
            xs[0][0] = std::floor(fx + 0.5); xs[0][1] = 1; nx = 1;      // Math.round
            ys[0][0] = std::floor(fy + 0.5); ys[0][1] = 1; ny = 1;
        }
        double ret[4] = {0, 0, 0, 0};
        for (int a = 0; a < nx; ++a)
            for (int b = 0; b < ny; ++b) {
                const double cy = js_min(js_max(ys[b][0], 0), height - 1), cx = js_min(js_max(xs[a][0], 0), width - 1);
                const long index = (long)(cy * width + cx);
                for (int i = 0; i < 4; ++i) ret[i] += xs[a][1] * ys[b][1] * (data[(size_t)index * 4 + i] / 255.0);
            }
        return Vec::of(ret[0], ret[1], ret[2], ret[3]);          // Vec.from(ret): a 4-vector (RGBA), f32
    }
};

// ---------------------------------------------------------------------------
struct LightSample { Vec direction, color; };
struct Light { virtual ~Light() {} virtual int sampleCount() const = 0; virtual LightSample sample(const Vec& pos, double u0, double u1) const = 0;
               virtual bool draws() const { return true; } };   // does sample() consume its two random numbers? (tape mode must not draw for point lights)
static double falloff(const Vec& delta) { return 1 / (4 * PI * delta.squarednorm()); }   // src/lights.js:21-23
struct SimplePointLight : Light {     // src/lights.js:45-53
    Vec position; MaterialColor* color_mc;
    int sampleCount() const override { return 1; }
    bool draws() const override { return false; }
    LightSample sample(const Vec& surface_position, double, double) const override {
        const Vec delta = position.minus(surface_position);
        MatData md; md.hasUV = true; md.UV = cartesianToSpherical(delta.normalized());
        return {delta, color_mc->color(md).times(falloff(delta))};
    }
};
struct RandomSampleAreaLight : Light {   // src/lights.js:80-93
    Geometry* surface_geometry; Mat4 transform, inv_transform; MaterialColor* color_mc; int samples;
    int sampleCount() const override { return samples; }
    LightSample sample(const Vec& surface_position, double u0, double u1) const override {
        const Vec local_pos = surface_geometry->sampleSurface(u0, u1);
        const Vec world_pos = transform.times(local_pos);
        const Vec delta = world_pos.minus(surface_position);
        MatData md; md.position = local_pos;
        surface_geometry->materialData(md, inv_transform.times(delta));
        const Vec nl = inv_transform.transposed().times(md.normal).to4(false).normalized();
        return {delta, color_mc->color(md).times(falloff(delta) * std::fabs(delta.normalized().dot(nl)))};
    }
};

// ---------------------------------------------------------------------------
// world objects (src/world.js, src/aggregates.js)
struct Intersection { double distance = INF; Primitive* object = nullptr; WorldObject* anc[8]; int nanc = 0; };

struct Material { virtual ~Material() {} virtual Vec color(MatData& d, World& world, int recursionDepth, uint32_t node, Ctx& ctx) = 0; };

struct WorldObject {
    Mat4 transform, inv_transform;
    virtual ~WorldObject() {}
    virtual Intersection intersect(const Ray& ray, double minDistance, double maxDistance, bool shadowCast, Ctx& ctx) = 0;
};

static Intersection getMinimumIntersection(const std::vector<WorldObject*>& objects, const Ray& ray, double minDistance,
                                           double maxDistance, bool intersectTransparent, Ctx& ctx, bool toplevel) {
    Intersection closest;             // src/world.js:7-15
    for (auto* o : objects) {
        if (ctx.c && toplevel) ctx.c->top_tests[ctx.rc]++;
        Intersection in = o->intersect(ray, minDistance, maxDistance, intersectTransparent, ctx);
        if (in.distance > minDistance && in.distance < closest.distance && in.distance < maxDistance) closest = in;
    }
    return closest;
}

struct Primitive : WorldObject {      // src/world.js:104-141
    Geometry* geometry = nullptr; Material* material = nullptr; bool does_cast_shadow = true; int prim_id = -1;
    Intersection intersect(const Ray& ray, double minDistance, double maxDistance, bool shadowCast, Ctx& ctx) override {
        Intersection r; r.object = this;
        if (!does_cast_shadow && !shadowCast) { r.distance = INF; return r; }
        r.distance = geometry->intersect(ray.getTransformed(inv_transform), minDistance, maxDistance, ctx);
        return r;
    }
    Vec color(const Ray& ray, double distance, const Mat4& ancestorInvTransform, World& world, int recursionDepth, uint32_t node, Ctx& ctx) {
        const Mat4 inv = inv_transform.times(ancestorInvTransform);
        MatData d; d.ray = ray; d.distance = distance;
        d.position = ray.getTransformed(inv).getPoint(distance);
        geometry->materialData(d, ray.direction);
        if (d.hasNormal) d.normal = inv.transposed().times(d.normal).to4(false).normalized();
        d.position = ray.getPoint(distance);
        if (ctx.c) ctx.c->shaded_hits++;
        return material->color(d, world, recursionDepth, node, ctx);
    }
};

struct Aggregate : WorldObject {      // src/aggregates.js:1-19
    std::vector<WorldObject*> objects;
    Intersection intersect(const Ray& ray, double minDistance, double maxDistance, bool shadowCast, Ctx& ctx) override {
        Intersection ret = getMinimumIntersection(objects, ray.getTransformed(inv_transform), minDistance, maxDistance, shadowCast, ctx, false);
        for (int i = ret.nanc; i > 0; --i) ret.anc[i] = ret.anc[i - 1];
        ret.anc[0] = this; ret.nanc++;
        return ret;
    }
};

struct BVHNode {                      // src/aggregates.js:63-232
    bool isLeaf = false; std::vector<WorldObject*> objects; AABBox aabb; BVHNode *lesser = nullptr, *greater = nullptr;
    void intersect(const Ray& ray, Intersection& ret, double minDist, double maxDist, bool intersectTransparent, Ctx& ctx, int depth = 0) const {
        if (ctx.c) ctx.c->bvh_nodes[ctx.rc]++;
        auto ts = aabb.get_intersects(ray, minDist, maxDist);
        if (ts.hit && ts.min <= maxDist && ts.max >= minDist && ts.min <= ret.distance) {
            if (ctx.c && !isLeaf) { if (depth % 2 == 0) ctx.c->wide4[ctx.rc]++; if (depth % 3 == 0) ctx.c->wide8[ctx.rc]++; }
            if (isLeaf) {
                for (auto* o : objects) {
                    if (ctx.c) ctx.c->bvh_prims[ctx.rc]++;
                    Intersection in = o->intersect(ray, minDist, maxDist, intersectTransparent, ctx);
                    if (in.distance > minDist && in.distance < maxDist && in.distance < ret.distance) {
                        ret.distance = in.distance; ret.object = in.object; ret.nanc = in.nanc;
                        for (int i = 0; i < in.nanc; ++i) ret.anc[i] = in.anc[i];
                    }
                }
            } else {
                greater->intersect(ray, ret, minDist, maxDist, intersectTransparent, ctx, depth + 1);
                lesser->intersect(ray, ret, minDist, maxDist, intersectTransparent, ctx, depth + 1);
            }
        }
    }
};
struct BVHAggregate : Aggregate {     // src/aggregates.js:26-61
    BVHNode* kdtree = nullptr;
    Intersection intersect(const Ray& ray, double minDist, double maxDist, bool intersectTransparent, Ctx& ctx) override {
        const Ray local_r = ray.getTransformed(inv_transform);
        Intersection ret;
        kdtree->intersect(local_r, ret, minDist, maxDist, intersectTransparent, ctx);
        for (int i = ret.nanc; i > 0; --i) ret.anc[i] = ret.anc[i - 1];
        ret.anc[0] = this; ret.nanc++;
        return ret;
    }
};

struct World {                        // src/world.js:1-42
    Vec bg_color; std::vector<WorldObject*> objects; std::vector<Light*> lights;
    Intersection cast(const Ray& ray, double minDistance, double maxDistance, bool intersectTransparent, Ctx& ctx) {
        if (ctx.c) ctx.c->rays[ctx.rc]++;
        return getMinimumIntersection(objects, ray, minDistance, maxDistance, intersectTransparent, ctx, true);
    }
    Vec color(const Ray& ray, int recursionDepth, double minDistance, uint32_t node, Ctx& ctx) {
        if (!recursionDepth) return Vec::of(0, 0, 0);
        const int saved = ctx.rc;
        ctx.rc = (node == 1) ? RC_PRIMARY : RC_SECONDARY;
        Intersection in = cast(ray, minDistance, INF, true, ctx);
        ctx.rc = saved;
        if (in.object == nullptr) return bg_color;
        Mat4 anc = Mat4::identity();
        for (int i = 0; i < in.nanc; ++i) anc = in.anc[i]->inv_transform.times(anc);
        return in.object->color(ray, in.distance, anc, *this, recursionDepth - 1, node, ctx);
    }
};

// ---------------------------------------------------------------------------
// materials (src/materials.js:145-476)
struct SolidColorMaterial : Material {
    MaterialColor* c;
    Vec color(MatData& d, World&, int, uint32_t, Ctx&) override { return c->color(d); }
};
struct PositionalUVMaterial : Material {      // src/materials.js:178-193
    Material* base; Vec origin, u_axis, v_axis;
    Vec color(MatData& d, World& world, int depth, uint32_t node, Ctx& ctx) override {
        const Vec delta = origin.minus(d.position);
        d.hasUV = true; d.UV = Vec::of(u_axis.dot(delta), v_axis.dot(delta));
        return base->color(d, world, depth, node, ctx);
    }
};
struct TransparentMaterial : Material {
    MaterialColor* c; double opacity;
    Vec color(MatData& d, World& world, int depth, uint32_t node, Ctx& ctx) override {
        return c->color(d).times(opacity).plus(world.color(Ray(d.position, d.ray.direction), depth, 0.0001, rng_child_node(node, 0), ctx).times(1 - opacity));
    }
};

struct PhongMaterial : Material {
    MaterialColor *baseColor, *ambient, *diffusivity, *specularity, *reflectivity, *transmissivity; double smoothness = 5;
    virtual void getBaseFactors(MatData& d) {          // src/materials.js:210-238
        d.V = d.ray.direction.normalized().times(-1);
        d.N = d.normal.normalized(); d.backside = false; d.vdotn = d.V.dot(d.N);
        if (d.vdotn < 0) { d.N = d.N.times(-1); d.backside = true; d.vdotn = -d.vdotn; }
        d.R = d.N.times(2 * d.vdotn).minus(d.V).normalized();
        const Vec basecolor = d.hasBase ? d.basecolor : Vec::of(1, 1, 1);
        d.ambient = basecolor.times(ambient->color(d));
        d.diffusivity = basecolor.times(diffusivity->color(d));
        d.specularity = specularity->color(d);
        d.reflectivity = reflectivity->color(d);
        d.transmissivity = transmissivity->color(d);
        d.smoothness = smoothness;
    }
    virtual Vec colorFromLightSample(const LightSample& ls, const MatData& d) {   // src/materials.js:261-269
        const Vec L = ls.direction.normalized();
        const double diffuse = js_max(L.dot(d.N), 0);
        const double specular = std::pow(js_max(L.dot(d.R), 0), smoothness);
        return ls.color.mult_pairs(d.diffusivity.times(diffuse)).plus(ls.color.mult_pairs(d.specularity.times(specular)));
    }
    Vec colorFromLights(MatData& d, World& world, uint32_t node, Ctx& ctx) {      // src/materials.js:240-259
        Vec ret = d.ambient;
        uint32_t dim = DIM_LIGHTS;
        for (auto* l : world.lights) {
            int count = 0; Vec light_color = Vec::of(0, 0, 0);
            const int ns = l->sampleCount();
            for (int s = 0; s < ns; ++s) {
                double u0 = 0, u1 = 0;
                if (!ctx.tape || l->draws()) { u0 = ctx.u(node, dim); u1 = ctx.u(node, dim + 1); }
                dim += 2;
                const LightSample ls = l->sample(d.position, u0, u1);
                ++count;
                const int saved = ctx.rc; ctx.rc = RC_SHADOW;
                const double shadowDist = world.cast(Ray(d.position, ls.direction), 0.0001, 1, false, ctx).distance;
                ctx.rc = saved;
                if (shadowDist > 0 && shadowDist < 1) continue;
                light_color = light_color.plus(colorFromLightSample(ls, d));
            }
            if (count > 0) ret = ret.plus(light_color.times(1.0 / count));
        }
        return ret;
    }
    static uint32_t scatterDimBase(World& world) { uint32_t n = 0; for (auto* l : world.lights) n += (uint32_t)l->sampleCount(); return DIM_LIGHTS + 2 * n; }
    Vec color(MatData& d, World& world, int depth, uint32_t node, Ctx& ctx) override {   // src/materials.js:271-291
        getBaseFactors(d);
        Vec surfaceColor = colorFromLights(d, world, node, ctx);
        if (d.reflectivity.squarednorm() > 0)
            surfaceColor = surfaceColor.plus(world.color(Ray(d.position, d.R), depth, 0.0001, rng_child_node(node, 0), ctx).mult_pairs(d.reflectivity));
        if (d.transmissivity.squarednorm() > 0)
            surfaceColor = surfaceColor.plus(world.color(Ray(d.position, d.ray.direction.normalized()), depth, 0.0001, rng_child_node(node, 1), ctx).mult_pairs(d.transmissivity));
        return surfaceColor;
    }
};

struct FresnelPhongMaterial : PhongMaterial {
    double refractiveIndexRatio = 1;
    double fresnelReflectionFactor(const MatData& d) const {     // src/materials.js:366-386
        if (!std::isfinite(refractiveIndexRatio)) return 1;
        const double ni = d.backside ? refractiveIndexRatio : 1, nt = d.backside ? 1 : refractiveIndexRatio;
        const double cosi = d.vdotn, sint = ni / nt * std::sqrt(js_max(0, 1 - cosi * cosi));
        if (sint >= 1) return 1;
        const double cost = std::sqrt(js_max(0, 1 - sint * sint));
        const double Rs = ((nt * cosi) - (ni * cost)) / ((nt * cosi) + (ni * cost));
        const double Rp = ((ni * cosi) - (nt * cost)) / ((ni * cosi) + (nt * cost));
        return (Rs * Rs + Rp * Rp) / 2;
    }
    void getBaseFactors(MatData& d) override {                    // src/materials.js:302-308,358-364
        PhongMaterial::getBaseFactors(d);
        d.kr = fresnelReflectionFactor(d);
        const double r = d.backside ? refractiveIndexRatio : 1 / refractiveIndexRatio, k = 1 - r * r * (1 - d.vdotn * d.vdotn);
        if (k < 0) d.hasRefr = false;
        else { d.hasRefr = true; d.refractionDirection = d.V.times(-1).times(r).plus(d.N.times(r * d.vdotn - std::sqrt(k))); }
    }
    // returns false for [null, *]; `col` is the weight
    virtual bool scatter(const Vec* R, const Vec& N, const MatData& d, uint32_t node, uint32_t dimBase, Ctx& ctx, Vec& dir, Vec& col) {
        (void)N; (void)d; (void)node; (void)dimBase; (void)ctx;
        col = Vec::of(1, 1, 1);
        if (!R) return false;
        dir = *R; return true;
    }
    Vec colorFromLightSample(const LightSample& ls, const MatData& d) override {   // src/materials.js:340-356
        const Vec L = ls.direction.normalized(); const double ldotn = L.dot(d.N);
        double diffuse = 0, specular = 0;
        if (d.kr > 0 && ldotn >= 0) { diffuse += d.kr * ldotn; specular += d.kr * std::pow(js_max(L.dot(d.R), 0), smoothness); }
        if (d.kr < 1 && ldotn <= 0) {
            diffuse += (1 - d.kr) * -ldotn;
            // L.dot(null) would throw in the reference; kr < 1 implies the direction exists up to rounding
            const double ldr = d.hasRefr ? L.dot(d.refractionDirection) : 0;
            specular += (1 - d.kr) * std::pow(js_max(ldr, 0), smoothness);
        }
        return ls.color.mult_pairs(d.diffusivity.times(diffuse)).plus(ls.color.mult_pairs(d.specularity.times(specular)));
    }
    Vec color(MatData& d, World& world, int depth, uint32_t node, Ctx& ctx) override {   // src/materials.js:309-333
        getBaseFactors(d);
        Vec surfaceColor = colorFromLights(d, world, node, ctx);
        const uint32_t sb = scatterDimBase(world);
        if (d.kr > 0) {
            Vec dir, col;
            if (scatter(&d.R, d.N, d, node, sb, ctx, dir, col))
                surfaceColor = surfaceColor.plus(world.color(Ray(d.position, dir), depth, 0.0001, rng_child_node(node, 0), ctx).times(col).times(d.reflectivity).times(d.kr));
        }
        if (d.kr < 1) {
            Vec dir, col;
            if (scatter(d.hasRefr ? &d.refractionDirection : nullptr, d.N.times(-1), d, node, sb + 4, ctx, dir, col))
                surfaceColor = surfaceColor.plus(world.color(Ray(d.position, dir), depth, 0.0001, rng_child_node(node, 1), ctx).times(col).times(d.transmissivity).times(1 - d.kr));
        }
        return surfaceColor;
    }
};

struct PhongPathTracingMaterial : FresnelPhongMaterial {
    double mirrorProbability = 0;
    bool scatter(const Vec* R, const Vec& N, const MatData& d, uint32_t node, uint32_t dimBase, Ctx& ctx, Vec& dir, Vec& col) override {
        // src/materials.js:398-412.  Draw order: mirror test, lobe choice, spherePick theta, phi.
        if (ctx.u(node, dimBase + 0) < mirrorProbability) { col = Vec::of(1, 1, 1); if (!R) return false; dir = *R; return true; }
        const double diffuseProb = d.diffusivity.average(), specularProb = d.specularity.average();
        const double probSum = diffuseProb + specularProb;
        if (probSum == 0) return false;
        if (ctx.u(node, dimBase + 1) < (diffuseProb / probSum)) {
            const double ut = ctx.u(node, dimBase + 2);      // theta first, then phi (src/math.js:181-182): sequenced for tape mode
            const double up = ctx.u(node, dimBase + 3);
            dir = N.plus(spherePick(ut, up).to4(false)).normalized();   // scatterDiffuse :438-440
            col = d.diffusivity.times(1 / PI);
            return true;
        }
        // scatterSpecular :441-445 returns R for every finite smoothness; the code
        // after it references undefined symbols and cannot run.
        if (!std::isfinite(d.smoothness)) throw std::runtime_error("scatterSpecular with infinite smoothness is not executable in the reference");
        col = d.specularity;
        if (!R) return false;
        dir = *R; return true;
    }
};

// ---------------------------------------------------------------------------
// cameras (src/cameras.js)
struct Camera {
    Mat4 transform; double tan_fov = 0, aspect = 1; bool dof = false; double focus_distance = 0, sensor_size = 0;
    Ray getRayForPixel(double x, double y, const Ctx& ctx) const {
        const Vec direction = Vec::of(x * tan_fov * aspect, y * tan_fov, -1, 0);
        Ray ray(transform.column(3), transform.times(direction));
        if (dof) {
            // Vec.circlePick src/math.js:175-179
            const double a = ctx.u(1, DIM_LENS_A) * 2 * PI, r = std::sqrt(ctx.u(1, DIM_LENS_R));
            const Vec offset = transform.times(Vec::of(r * std::cos(a), r * std::sin(a)).times(sensor_size).to4(false));
            ray.origin = ray.origin.plus(offset);
            ray.direction = ray.direction.times(focus_distance).minus(offset).normalized();
        }
        return ray;
    }
};

// ---------------------------------------------------------------------------
// scene = {renderer:{world,camera,maxRecursionDepth,samplesPerPixel}, width, height}
struct Scene {
    std::unique_ptr<JsonParser> parser;
    Deser deser;
    std::vector<std::unique_ptr<Geometry>> geoms; std::vector<std::unique_ptr<SDFNode>> sdfs; std::vector<std::unique_ptr<SDFTransformer>> sdfts;
    std::vector<std::unique_ptr<MaterialColor>> mcs; std::vector<std::unique_ptr<Material>> mats; std::vector<std::unique_ptr<Light>> lights;
    std::vector<std::unique_ptr<WorldObject>> wobjs; std::vector<std::unique_ptr<BVHNode>> bvhnodes;
    std::unordered_map<const Node*, void*> memo;
    World world; Camera camera; int maxRecursionDepth = 3, samplesPerPixel = 1, width = 0, height = 0; bool jitter = true;
    int nprims = 0; std::vector<Primitive*> prims_by_id;
    std::string renderer_type;

    Vec vec(const Node* n) { return toVec(n); }
    double numOr(const Node* n, double dflt) { return (n && !(n->kind == Node::RAW && !n->raw)) ? n->num() : dflt; }

    Geometry* geometry(const Node* n);
    SDFNode* sdf(const Node* n);
    SDFTransformer* sdft(const Node* n);
    MaterialColor* mcolor(const Node* n);
    Material* material(const Node* n);
    WorldObject* wobject(const Node* n);
    BVHNode* bvhnode(const Node* n);
    Light* light(const Node* n);
    AABBox aabb(const Node* n) { AABBox b; b.center = vec(n->get("center")); b.half_size = vec(n->get("half_size")); return b; }
    void assignIds(const std::vector<WorldObject*>& objs);
};

Geometry* Scene::geometry(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (Geometry*)it->second;
    Geometry* g = nullptr; const std::string& t = n->type;
    if (t == "Plane" || t == "SimplePlane") g = new SimplePlane;
    else if (t == "Square") g = new Square;
    else if (t == "Circle") g = new Circle;
    else if (t == "Sphere") g = new Sphere;
    else if (t == "Cylinder") g = new Cylinder;
    else if (t == "UnitBox") { auto* b = new AABBGeom; b->box.center = Vec::of(0, 0, 0, 1); b->box.half_size = Vec::of(0.5, 0.5, 0.5, 0); g = b; }   // UnitBox.deserialize src/geometry.js:234-236
    else if (t == "AABB") { auto* b = new AABBGeom; b->box = aabb(n); g = b; }
    else if (t == "Triangle") {
        auto* tr = new Triangle; const Node* ps = n->get("ps");
        for (int i = 0; i < 3; ++i) tr->ps[i] = vec(ps->arr.at(i));
        tr->init();
        const Node* pd = n->get("psdata");
        if (pd && pd->kind == Node::TYPED_OBJ) {   // {UV:[...], normal:[...]}; the reference's lossy serialize gives an Array here
            if (const Node* uv = pd->get("UV")) { tr->hasUVs = true; for (int i = 0; i < 3; ++i) tr->vuv[i] = vec(uv->arr.at(i)); }
            if (const Node* nn = pd->get("normal")) { tr->hasNormals = true; for (int i = 0; i < 3; ++i) tr->vnormal[i] = vec(nn->arr.at(i)); }
        }
        g = tr;
    } else if (t == "SDFGeometry") {
        auto* s = new SDFGeometry; s->root = sdf(n->get("root_sdf")); s->aabb = aabb(n->get("aabb"));
        s->max_samples = n->get("max_samples")->num(); s->distance_epsilon = n->get("distance_epsilon")->num();
        s->max_trace_distance = n->get("max_trace_distance")->num(); s->normal_step_size = n->get("normal_step_size")->num();
        g = s;
    } else throw std::runtime_error("oracle: unsupported geometry " + t);
    geoms.emplace_back(g); memo[n] = g; return g;
}

SDFTransformer* Scene::sdft(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (SDFTransformer*)it->second;
    SDFTransformer* r = nullptr; const std::string& t = n->type;
    if (t == "SDFTransformerSequence") { auto* s = new SDFTransformerSequence; for (auto* c : n->get("transformers")->arr) s->ts.push_back(sdft(c)); r = s; }
    else if (t == "SDFRecursiveTransformer") { auto* s = new SDFRecursiveTransformer; s->t = sdft(n->get("transformer")); s->iterations = (int)n->get("iterations")->num(); r = s; }
    else if (t == "SDFMatrixTransformer") { auto* s = new SDFMatrixTransformer; s->inv = toMat(n->get("_inv_transform")); s->scale = n->get("_scale")->num(); r = s; }
    else if (t == "SDFReflectionTransformer") { auto* s = new SDFReflectionTransformer; s->normal = vec(n->get("normal")); s->delta = n->get("delta")->num(); r = s; }
    else if (t == "SDFInfiniteRepetitionTransformer") { auto* s = new SDFInfiniteRepetitionTransformer; s->sizes = vec(n->get("sizes")); r = s; }
    else throw std::runtime_error("oracle: unsupported SDF transformer " + t);
    sdfts.emplace_back(r); memo[n] = r; return r;
}

SDFNode* Scene::sdf(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (SDFNode*)it->second;
    SDFNode* r = nullptr; const std::string& t = n->type;
    auto base = [&](const Node* b) { return b ? vec(b) : Vec::of(1, 1, 1); };
    if (t == "UnionSDF") { auto* s = new UnionSDF; for (auto* c : n->get("children")->arr) s->children.push_back(sdf(c)); r = s; }
    else if (t == "IntersectionSDF") { auto* s = new IntersectionSDF; for (auto* c : n->get("children")->arr) s->children.push_back(sdf(c)); r = s; }
    else if (t == "DifferenceSDF") { auto* s = new DifferenceSDF; s->positive = sdf(n->get("positive")); s->negative = sdf(n->get("negative")); r = s; }
    else if (t == "SmoothUnionSDF") { auto* s = new SmoothUnionSDF; s->k = n->get("k")->num(); s->a = sdf(n->get("childA")); s->b = sdf(n->get("childB")); r = s; }
    else if (t == "SmoothIntersectionSDF") { auto* s = new SmoothIntersectionSDF; s->k = n->get("k")->num(); s->a = sdf(n->get("childA")); s->b = sdf(n->get("childB")); r = s; }
    else if (t == "SmoothDifferenceSDF") { auto* s = new SmoothDifferenceSDF; s->k = n->get("k")->num(); s->positive = sdf(n->get("positive")); s->negative = sdf(n->get("negative")); r = s; }
    else if (t == "RoundSDF") { auto* s = new RoundSDF; s->child = sdf(n->get("child_sdf")); s->rounding = n->get("rounding")->num(); r = s; }
    else if (t == "SphereSDF") { auto* s = new SphereSDF; s->radius = n->get("radius")->num(); s->basecolor = base(n->get("basecolor")); r = s; }
    else if (t == "BoxSDF") { auto* s = new BoxSDF; s->size = vec(n->get("size")); s->basecolor = base(n->get("basecolor")); r = s; }
    else if (t == "TetrahedronSDF") { auto* s = new TetrahedronSDF; s->basecolor = base(n->get("basecolor")); r = s; }
    else if (t == "TransformSDF") { auto* s = new TransformSDF; s->child = sdf(n->get("child_sdf")); s->transformer = sdft(n->get("transformer")); r = s; }
    else if (t == "RecursiveTransformUnionSDF") { auto* s = new RecursiveTransformUnionSDF; s->sdf = sdf(n->get("sdf")); s->transformer = sdft(n->get("transformer")); s->iterations = (int)n->get("iterations")->num(); r = s; }
    else throw std::runtime_error("oracle: unsupported SDF node " + t);
    sdfs.emplace_back(r); memo[n] = r; return r;
}

MaterialColor* Scene::mcolor(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (MaterialColor*)it->second;
    MaterialColor* r = nullptr; const std::string& t = n->type;
    if (t == "SolidMaterialColor") { auto* s = new SolidMaterialColor; s->c = vec(n->get("_color")); r = s; }
    else if (t == "ScaledMaterialColor") {
        auto* s = new ScaledMaterialColor; s->mc = mcolor(n->get("_mc"));
        const Node* sc = n->get("_scale");
        if (sc->kind == Node::RAW && sc->raw && sc->raw->t == JV::NUM) s->s = sc->raw->num;
        else { s->isArray = true; s->sv = vec(sc); }
        r = s;
    } else if (t == "CheckerboardMaterialColor") { auto* s = new CheckerboardMaterialColor; s->c1 = mcolor(n->get("color1")); s->c2 = mcolor(n->get("color2")); r = s; }
    else if (t == "TextureMaterialColor") {
        auto* s = new TextureMaterialColor;
        const Node* img = n->get("_imgdata");
        s->width = (int)img->get("width")->num(); s->height = (int)img->get("height")->num();
        const Node* dn = img->get("data");
        if (!dn || dn->kind != Node::RAW || !dn->raw || dn->raw->t != JV::ARR || (long)dn->raw->arr.size() != (long)s->width * s->height * 4)
            throw std::runtime_error("oracle: TextureMaterialColor needs _imgdata.data as an array of width*height*4 numbers");
        s->data.reserve(dn->raw->arr.size());
        for (const JV* x : dn->raw->arr) s->data.push_back((unsigned char)x->num);
        const Node* mode = n->get("mode");
        s->bilinear = !(mode && mode->kind == Node::RAW && mode->raw && mode->raw->t == JV::STR && mode->raw->s == "nearest");
        s->clampU = n->get("clampU") ? n->get("clampU")->boolean() : true;
        s->clampV = n->get("clampV") ? n->get("clampV")->boolean() : true;
        r = s;
    }
    else throw std::runtime_error("oracle: unsupported MaterialColor " + t);
    mcs.emplace_back(r); memo[n] = r; return r;
}

Material* Scene::material(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (Material*)it->second;
    Material* r = nullptr; const std::string& t = n->type;
    if (t == "PhongMaterial" || t == "FresnelPhongMaterial" || t == "PhongPathTracingMaterial") {
        PhongMaterial* p;
        if (t == "PhongMaterial") p = new PhongMaterial;
        else if (t == "FresnelPhongMaterial") { auto* f = new FresnelPhongMaterial; f->refractiveIndexRatio = numOr(n->get("refractiveIndexRatio"), INF); p = f; }
        else { auto* f = new PhongPathTracingMaterial; f->refractiveIndexRatio = numOr(n->get("refractiveIndexRatio"), INF); f->mirrorProbability = numOr(n->get("mirrorProbability"), 0); p = f; }
        p->baseColor = mcolor(n->get("baseColor")); p->ambient = mcolor(n->get("ambient")); p->diffusivity = mcolor(n->get("diffusivity"));
        p->specularity = mcolor(n->get("specularity")); p->reflectivity = mcolor(n->get("reflectivity")); p->transmissivity = mcolor(n->get("transmissivity"));
        p->smoothness = n->get("smoothness")->num();
        r = p;
    } else if (t == "SolidColorMaterial") { auto* s = new SolidColorMaterial; s->c = mcolor(n->get("_color")); r = s; }
    else if (t == "TransparentMaterial") { auto* s = new TransparentMaterial; s->c = mcolor(n->get("_color")); s->opacity = n->get("_opacity")->num(); r = s; }
    else if (t == "PositionalUVMaterial") {
        auto* s = new PositionalUVMaterial; s->base = material(n->get("baseMaterial"));
        s->origin = vec(n->get("origin")); s->u_axis = vec(n->get("u_axis")); s->v_axis = vec(n->get("v_axis")); r = s;
    }
    else throw std::runtime_error("oracle: unsupported material " + t);
    mats.emplace_back(r); memo[n] = r; return r;
}

BVHNode* Scene::bvhnode(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (BVHNode*)it->second;
    auto* b = new BVHNode; bvhnodes.emplace_back(b); memo[n] = b;
    b->isLeaf = n->get("isLeaf")->boolean();
    b->aabb = aabb(n->get("aabb"));
    if (b->isLeaf) { if (const Node* o = n->get("objects")) for (auto* c : o->arr) b->objects.push_back(wobject(c)); }
    else { b->lesser = bvhnode(n->get("lesser_node")); b->greater = bvhnode(n->get("greater_node")); }
    return b;
}

WorldObject* Scene::wobject(const Node* n) {
    auto it = memo.find(n); if (it != memo.end()) return (WorldObject*)it->second;
    WorldObject* r = nullptr; const std::string& t = n->type;
    if (t == "Primitive") {
        auto* p = new Primitive; p->geometry = geometry(n->get("geometry")); p->material = material(n->get("material"));
        p->does_cast_shadow = n->get("does_cast_shadow") ? n->get("does_cast_shadow")->boolean() : true; r = p;
    } else if (t == "Aggregate") { auto* a = new Aggregate; for (auto* c : n->get("objects")->arr) a->objects.push_back(wobject(c)); r = a; }
    else if (t == "BVHAggregate") {
        auto* a = new BVHAggregate; wobjs.emplace_back(a); memo[n] = a;
        a->transform = toMat(n->get("transform")); a->inv_transform = toMat(n->get("inv_transform"));
        for (auto* c : n->get("objects")->arr) a->objects.push_back(wobject(c));
        a->kdtree = bvhnode(n->get("kdtree"));
        return a;
    } else throw std::runtime_error("oracle: unsupported world object " + t);
    r->transform = toMat(n->get("transform")); r->inv_transform = toMat(n->get("inv_transform"));
    wobjs.emplace_back(r); memo[n] = r; return r;
}

Light* Scene::light(const Node* n) {
    Light* r = nullptr; const std::string& t = n->type;
    if (t == "SimplePointLight") { auto* l = new SimplePointLight; l->position = vec(n->get("position")); l->color_mc = mcolor(n->get("color_mc")); r = l; }
    else if (t == "RandomSampleAreaLight") {
        auto* l = new RandomSampleAreaLight; l->surface_geometry = geometry(n->get("surface_geometry"));
        l->transform = toMat(n->get("transform")); l->inv_transform = toMat(n->get("inv_transform"));
        l->color_mc = mcolor(n->get("color_mc")); l->samples = (int)n->get("samples")->num(); r = l;
    } else throw std::runtime_error("oracle: unsupported light " + t);
    lights.emplace_back(r); return r;
}

// prim_id = index of the Primitive's first appearance in a depth-first walk of
// world.objects, descending into Aggregate.objects.
void Scene::assignIds(const std::vector<WorldObject*>& objs) {
    for (auto* o : objs) {
        if (auto* p = dynamic_cast<Primitive*>(o)) { if (p->prim_id < 0) { p->prim_id = nprims++; prims_by_id.push_back(p); } }
        else if (auto* a = dynamic_cast<Aggregate*>(o)) assignIds(a->objects);
    }
}

static Scene* loadScene(const char* text, size_t len) {
    std::unique_ptr<Scene> s(new Scene);
    s->parser.reset(new JsonParser(text, len));
    const JV* rootj = s->parser->parse();
    const Node* root = s->deser.step(rootj);
    const Node* rend = root->get("renderer");
    if (!rend) throw std::runtime_error("oracle: no renderer in scene");
    s->renderer_type = rend->type;
    s->jitter = rend->type != "SimpleRenderer";
    s->width = (int)root->get("width")->num(); s->height = (int)root->get("height")->num();
    s->maxRecursionDepth = (int)rend->get("maxRecursionDepth")->num();
    s->samplesPerPixel = rend->get("samplesPerPixel") ? (int)rend->get("samplesPerPixel")->num() : 1;
    const Node* w = rend->get("world");
    s->world.bg_color = toVec(w->get("bg_color"));
    for (auto* o : w->get("objects")->arr) s->world.objects.push_back(s->wobject(o));
    for (auto* l : w->get("lights")->arr) s->world.lights.push_back(s->light(l));
    s->assignIds(s->world.objects);
    const Node* c = rend->get("camera");
    s->camera.transform = toMat(c->get("transform"));
    s->camera.tan_fov = c->get("tan_fov")->num(); s->camera.aspect = c->get("aspect")->num();
    if (c->type == "DepthOfFieldPerspectiveCamera") { s->camera.dof = true; s->camera.focus_distance = c->get("focus_distance")->num(); s->camera.sensor_size = c->get("sensor_size")->num(); }
    return s.release();
}

static void parallelFor(int nthreads, long n, const std::function<void(long, long, int)>& fn) {
    if (nthreads <= 1) { fn(0, n, 0); return; }
    std::atomic<long> next(0);
    const long chunk = 256;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() { for (;;) { long b = next.fetch_add(chunk); if (b >= n) break; fn(b, std::min(n, b + chunk), t); } });
    for (auto& x : th) x.join();
}

}  // namespace orc

// ---------------------------------------------------------------------------
using namespace orc;
static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_load(const char* json, size_t len) {
    try { return loadScene(json, len); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_free(void* h) { delete (Scene*)h; }

// out[0..7] = width, height, samplesPerPixel, maxRecursionDepth, nprims, jitter, nlights, ntop
int orc_info(void* h, int* out) {
    Scene* s = (Scene*)h;
    out[0] = s->width; out[1] = s->height; out[2] = s->samplesPerPixel; out[3] = s->maxRecursionDepth; out[4] = s->nprims;
    out[5] = s->jitter; out[6] = (int)s->world.lights.size(); out[7] = (int)s->world.objects.size();
    return 0;
}

// Un-jittered primary rays (SimpleRenderer sampling, src/renderers.js:21-25):
// prim_id (-1 = miss) and hit distance per pixel, row-major [y*W+x].
int orc_primary_hits(void* h, int W, int H, int32_t* prim_id, double* tout, int nthreads, uint64_t* counters) {
    Scene* s = (Scene*)h;
    try {
        std::vector<Counters> cs(std::max(1, nthreads));
        Camera pinhole = s->camera; pinhole.dof = false;   // the probe ignores the lens (PerspectiveCamera.getRayForPixel)
        parallelFor(nthreads, (long)W * H, [&](long b, long e, int tid) {
            Ctx ctx; ctx.c = &cs[tid];
            for (long i = b; i < e; ++i) {
                const int px = (int)(i % W), py = (int)(i / W);
                const double x = 2 * ((double)px / W) - 1, y = -2 * ((double)py / H) + 1;
                ctx.rc = RC_PRIMARY;
                Intersection in = s->world.cast(pinhole.getRayForPixel(x, y, ctx), 0, INF, true, ctx);
                prim_id[i] = in.object ? in.object->prim_id : -1;
                tout[i] = in.distance;
            }
        });
        if (counters) { Counters t; for (auto& c : cs) t.add(c); for (int i = 0; i < 3; ++i) { counters[i] = t.rays[i]; counters[3 + i] = t.bvh_nodes[i]; counters[6 + i] = t.bvh_prims[i]; counters[9 + i] = t.top_tests[i]; counters[12 + i] = t.sdf_evals[i]; } counters[15] = t.shaded_hits; for (int i = 0; i < 3; ++i) { counters[16 + i] = t.wide4[i]; counters[19 + i] = t.wide8[i]; } }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// IncrementalMultisamplingRenderer.render (src/renderers.js:70-117) for passes
// [first_pass, first_pass+n_passes): `sum` (W*H*3 f32, row-major) is the
// reference's `buffer[px][py]` and is accumulated in f32 like `Vec.plus`.
// flags bit0: no pixel jitter (SimpleRenderer sampling, :21-25).  bit1: tape mode (see Ctx).
// bit2: RandomMultisamplingRenderer (:47-63) — `sum` then holds each pixel's colour, not a sum over passes.
// counters[22]: rays[3], bvh_nodes[3], bvh_prims[3], top_tests[3], sdf_evals[3], shaded_hits, wide4[3], wide8[3]
int orc_render(void* h, int W, int H, int first_pass, int n_passes, uint64_t seed, int flags, int x_offset, int x_delt,
               float* sum, int nthreads, uint64_t* counters) {
    Scene* s = (Scene*)h;
    try {
        const bool jitter = !(flags & 1);
        const double pixel_width = 2.0 / W, pixel_height = 2.0 / H;
        std::vector<Counters> cs(std::max(1, nthreads));
        if (x_delt < 1) x_delt = 1;
        const int ncols = (W - x_offset + x_delt - 1) / x_delt;
        parallelFor(nthreads, (long)ncols * H, [&](long b, long e, int tid) {
            Ctx ctx; ctx.c = &cs[tid];
            for (long i = b; i < e; ++i) {
                const int px = x_offset + (int)(i % ncols) * x_delt, py = (int)(i / ncols);
                const uint32_t pixel = (uint32_t)(py * W + px);
                const double x = 2 * ((double)px / W) - 1, y = -2 * ((double)py / H) + 1;
                float* acc = sum + (size_t)pixel * 3;
                if (flags & 4) {
                    // RandomMultisamplingRenderer.getPixelColor (src/renderers.js:52-62): all samples of a pixel inside one call —
                    // `color = color.plus(world.color(...).times(1 / spp))`, every Vec op storing f32 — then ONE setColor; in tape
                    // mode the pixel's samples share one stream, the way consecutive Math.random() calls do.  `sum` receives the
                    // pixel's colour itself.
                    if (flags & 2) { ctx.tape = true; ctx.tape_state = rng_tape_seed(seed, pixel, (uint32_t)first_pass); }
                    float col[3] = {0, 0, 0};
                    const double inv = 1.0 / n_passes;
                    for (int iter = first_pass; iter < first_pass + n_passes; ++iter) {
                        ctx.sample_key = rng_sample_key(seed, pixel, (uint32_t)iter);
                        const double jx = ctx.u(1, DIM_JITTER_X), jy = ctx.u(1, DIM_JITTER_Y);
                        const double sx = x + pixel_width * (jx - 0.5), sy = y + pixel_height * (jy - 0.5);
                        ctx.rc = RC_PRIMARY;
                        const Vec color = s->world.color(s->camera.getRayForPixel(sx, sy, ctx), s->maxRecursionDepth, 0, 1, ctx);
                        for (int k = 0; k < 3; ++k) col[k] = (float)((double)col[k] + (double)(float)(color.at(k) * inv));
                    }
                    for (int k = 0; k < 3; ++k) acc[k] = col[k];
                    continue;
                }
                for (int iter = first_pass; iter < first_pass + n_passes; ++iter) {
                    ctx.sample_key = rng_sample_key(seed, pixel, (uint32_t)iter);
                    if (flags & 2) { ctx.tape = true; ctx.tape_state = rng_tape_seed(seed, pixel, (uint32_t)iter); }
                    double sx = x, sy = y;
                    if (jitter) { sx = x + pixel_width * (ctx.u(1, DIM_JITTER_X) - 0.5); sy = y + pixel_height * (ctx.u(1, DIM_JITTER_Y) - 0.5); }
                    ctx.rc = RC_PRIMARY;
                    const Vec color = s->world.color(s->camera.getRayForPixel(sx, sy, ctx), s->maxRecursionDepth, 0, 1, ctx);
                    for (int k = 0; k < 3; ++k) acc[k] = (float)((double)acc[k] + color.at(k));
                }
            }
        });
        if (counters) { Counters t; for (auto& c : cs) t.add(c); for (int i = 0; i < 3; ++i) { counters[i] = t.rays[i]; counters[3 + i] = t.bvh_nodes[i]; counters[6 + i] = t.bvh_prims[i]; counters[9 + i] = t.top_tests[i]; counters[12 + i] = t.sdf_evals[i]; } counters[15] = t.shaded_hits; for (int i = 0; i < 3; ++i) { counters[16 + i] = t.wide4[i]; counters[19 + i] = t.wide8[i]; } }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// PixelBuffer.setColor (src/pixelbuffer.js:39-49) applied to sum/(passes):
// `buffer.times(1/(iter+1))` -> clamp -> Math.round(255*c); alpha 255.
int orc_resolve_rgba8(const float* sum, int npix, int passes, uint8_t* out) {
    const double inv = 1.0 / passes;
    for (int i = 0; i < npix; ++i) {
        for (int k = 0; k < 3; ++k) {
            const float c = (float)((double)sum[i * 3 + k] * inv);
            double comp = js_min(js_max((double)c, 0), 1);
            double v = std::floor(255 * comp + 0.5);
            out[i * 4 + k] = (uint8_t)(v != v ? 0 : v);
        }
        out[i * 4 + 3] = 255;
    }
    return 0;
}

// ---- known-answer-test entry points (tests/test_oracle_kat.py) -------------
// kind: 0 plane, 1 square, 2 circle, 3 unitbox, 4 sphere, 5 cylinder, 6 triangle(ps = 12 doubles in `extra`)
double orc_kat_intersect(int kind, const double* o, const double* d, double minD, double maxD, const double* extra) {
    Ray r(Vec::of(o[0], o[1], o[2], o[3]), Vec::of(d[0], d[1], d[2], d[3])); Ctx ctx;
    switch (kind) {
        case 0: { SimplePlane g; return g.intersect(r, minD, maxD, ctx); }
        case 1: { Square g; return g.intersect(r, minD, maxD, ctx); }
        case 2: { Circle g; return g.intersect(r, minD, maxD, ctx); }
        case 3: { AABBGeom g; g.box.center = Vec::of(0, 0, 0, 1); g.box.half_size = Vec::of(0.5, 0.5, 0.5, 0); return g.intersect(r, minD, maxD, ctx); }
        case 4: { Sphere g; return g.intersect(r, minD, maxD, ctx); }
        case 5: { Cylinder g; return g.intersect(r, minD, maxD, ctx); }
        case 6: { Triangle g; for (int i = 0; i < 3; ++i) g.ps[i] = Vec::of(extra[4 * i], extra[4 * i + 1], extra[4 * i + 2], extra[4 * i + 3]); g.init(); return g.intersect(r, minD, maxD, ctx); }
    }
    return std::nan("");
}
// out[0..2] normal xyz, out[3..4] UV (NaN if absent)
int orc_kat_material_data(int kind, const double* pos, const double* extra, double* out) {
    MatData d; d.position = Vec::of(pos[0], pos[1], pos[2], pos[3]); Vec dir = Vec::of(0, 0, -1, 0);
    switch (kind) {
        case 0: case 1: case 2: { SimplePlane g; g.materialData(d, dir); break; }
        case 3: { AABBGeom g; g.box.center = Vec::of(0, 0, 0, 1); g.box.half_size = Vec::of(0.5, 0.5, 0.5, 0); g.materialData(d, dir); break; }
        case 4: { Sphere g; g.materialData(d, dir); break; }
        case 5: { Cylinder g; g.materialData(d, dir); break; }
        case 6: { Triangle g; for (int i = 0; i < 3; ++i) g.ps[i] = Vec::of(extra[4 * i], extra[4 * i + 1], extra[4 * i + 2], extra[4 * i + 3]); g.init(); g.materialData(d, dir); break; }
        default: return 1;
    }
    for (int i = 0; i < 3; ++i) out[i] = d.normal[i];
    out[3] = d.hasUV ? d.UV[0] : std::nan(""); out[4] = d.hasUV ? d.UV[1] : std::nan("");
    return 0;
}
double orc_kat_fresnel(double ior, double vdotn, int backside, double* refr_k) {
    FresnelPhongMaterial m; m.refractiveIndexRatio = ior; MatData d; d.vdotn = vdotn; d.backside = backside != 0;
    const double r = d.backside ? ior : 1 / ior; if (refr_k) *refr_k = 1 - r * r * (1 - vdotn * vdotn);
    return m.fresnelReflectionFactor(d);
}
double orc_kat_fmod(double a, double b) { return js_fmod(a, b); }
double orc_kat_rng(uint64_t seed, uint32_t pixel, uint32_t pass, uint32_t node, uint32_t dim) { return rng_u01(rng_node_key(rng_sample_key(seed, pixel, pass), node), dim); }
// SDF distance / material probe on a loaded scene: prim_id must be an SDFGeometry primitive
int orc_sdf_probe(void* h, int prim_id, const double* p, double* dist, double* normal_uv_base) {
    Scene* s = (Scene*)h;
    try {
        auto* g = dynamic_cast<SDFGeometry*>(s->prims_by_id.at(prim_id)->geometry);
        if (!g) { g_err = "not an SDF primitive"; return 1; }
        const Vec P = Vec::of(p[0], p[1], p[2], 1);
        *dist = g->root->distance(P);
        if (normal_uv_base) {
            MatData d; d.position = P; g->materialData(d, Vec::of(0, 0, -1, 0));
            for (int i = 0; i < 3; ++i) normal_uv_base[i] = d.normal[i];
            normal_uv_base[3] = d.hasUV ? d.UV[0] : std::nan(""); normal_uv_base[4] = d.hasUV ? d.UV[1] : std::nan("");
            for (int i = 0; i < 3; ++i) normal_uv_base[5 + i] = d.hasBase ? d.basecolor[i] : 1.0;
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// Single-ray probe in world space: returns prim id, writes t; rc_shadow selects shadow-cast semantics
int orc_cast(void* h, const double* o, const double* d, double minD, double maxD, int shadow, double* t) {
    Scene* s = (Scene*)h; Ctx ctx;
    Intersection in = s->world.cast(Ray(Vec::of(o[0], o[1], o[2], 1), Vec::of(d[0], d[1], d[2], 0)), minD, maxD, !shadow, ctx);
    *t = in.distance; return in.object ? in.object->prim_id : -1;
}

}  // extern "C"
