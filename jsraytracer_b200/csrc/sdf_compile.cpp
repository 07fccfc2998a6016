// SDF operator tree (reference src/sdf.js:53-477) -> straight-line bytecode for
// the register-stack interpreter in kernels.cu.
//
// `node.distance(p)` is compiled to code that, with the current point on top of
// the point stack, pushes one distance.  Loops with fixed trip counts
// (RecursiveTransformUnionSDF.iterations, SDFRecursiveTransformer.iterations)
// are unrolled, so a program has no branches: every lane of a warp executes the
// same instruction stream.
//   TransformSDF        : PUSHP; T...; child; MULS; POPP        (src/sdf.js:330-333)
//   RecursiveTransformU.: PUSHP; sdf; { T...; sdf; MULS; MIN }*n; POPP   (:349-357)
//   Union/Intersection  : c0; c1; MIN|MAX; c2; MIN|MAX ...      (:83-101)
//   Difference          : pos; neg; NEG; MAX                    (:116-118)
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

#include <algorithm>

#include "host_scene.h"

namespace jsrt {

static const double kInf = std::numeric_limits<double>::infinity();

namespace {

struct SdfCompiler {
    const WireDoc& doc;
    HostScene& out;
    int depth_points = 1, max_points = 1, depth_dist = 0, max_dist = 0, depth_scale = 1, max_scale = 1;
    bool base_set = false, base_uniform = true, any_sphere_leaf = false;
    float base[3] = {1, 1, 1};

    SdfCompiler(const WireDoc& d, HostScene& o) : doc(d), out(o) {}

    // guards against malformed blobs: cyclic `_r` references (NestGuard) and iteration counts that would unroll into
    // an unbounded program (the reference's scenes: 22..41 instructions per distance query, <= 10 iterations)
    static constexpr int kMaxNest = 256;
    static constexpr size_t kMaxInstrs = 1u << 20;
    int nest = 0;

    // Peephole fusion (JSRT_SDF_FUSE=0 emits the plain code): the interpreter pays a fetch + dispatch per instruction,
    // and on the fractal scenes two thirds of the stream were stack shuffles.
    //   leaf; MIN|MAX      -> leaf with idx = 1|2 (folds into the distance below instead of pushing)
    //   MULS; MIN          -> MULS_MIN
    //   SBEGIN ... SEND    -> elided when the group cannot round differently: it multiplies the scale at most once
    //                         (1 * a = a exactly), or the enclosing scale is still exactly 1 (src/sdf.js:387-394)
    bool fuse = true;
    size_t prog_first = 0;                 // first instruction of the program being emitted (no fusion across programs)
    bool scale_is_one = true;              // the current scale accumulator has not been multiplied yet
    std::vector<bool> one_stack;
    void emit(int op, double a0 = 0, float f0 = 0, float f1 = 0, float f2 = 0, int idx = 0) {
        if (out.sdf_code.size() >= kMaxInstrs) fail("jsrt: SDF program too long (more than 2^20 instructions after unrolling)");
        if (fuse && out.sdf_code.size() > prog_first) {
            SdfInstr& last = out.sdf_code.back();
            if ((op == S_MIN || op == S_MAX) && ((last.op >= S_SPHERE && last.op <= S_TETRA) || last.op == S_CROSS) && last.idx == 0) {
                last.idx = (op == S_MIN) ? 1 : 2; --depth_dist;
                fuseCross();
                return;
            }
            if (op == S_MIN && last.op == S_MULS) { last.op = S_MULS_MIN; --depth_dist; return; }
        }
        SdfInstr i{}; i.op = op; i.idx = idx; i.a0 = a0; i.f[0] = f0; i.f[1] = f1; i.f[2] = f2; i.f[3] = 0;
        out.sdf_code.push_back(i);
        switch (op) {
            case S_SPHERE: case S_BOX: case S_TETRA: if (++depth_dist > max_dist) max_dist = depth_dist; break;
            case S_MIN: case S_MAX: case S_SMIN: case S_SMIN_NEGA: case S_SMIN_NEGAB: --depth_dist; break;
            case S_PUSHP: if (++depth_points > max_points) max_points = depth_points; if (++depth_scale > max_scale) max_scale = depth_scale; break;
            case S_POPP: --depth_points; --depth_scale; break;
            case S_SBEGIN: if (++depth_scale > max_scale) max_scale = depth_scale; break;
            case S_SEND: --depth_scale; break;
            default: break;
        }
        if (out.sdf_code.size() > 65536) fail("jsrt: SDF program longer than 65536 instructions after unrolling");
    }
    //   BOX(Inf,a,a); BOX(a,Inf,a) min; BOX(a,a,Inf) min  -> CROSS(a): the union of the three axis bars (the Menger sponge's
    //                         building block) from one q = |p| - a, bit for bit the three BoxSDF distances and their
    //                         Math.min (device_math.cuh); two fetch + dispatch rounds and two thirds of the box code saved
    void fuseCross() {
        const size_t n = out.sdf_code.size();
        if (n < prog_first + 3) return;
        SdfInstr& b0 = out.sdf_code[n - 3]; const SdfInstr& b1 = out.sdf_code[n - 2]; const SdfInstr& b2 = out.sdf_code[n - 1];
        if (b0.op != S_BOX || b1.op != S_BOX || b2.op != S_BOX || b1.idx != 1 || b2.idx != 1) return;
        const float a = b0.f[1];
        auto inf = [](float v) { return std::isinf(v) && v > 0; };
        if (!(std::isfinite(a) && a >= 0)) return;
        if (!(inf(b0.f[0]) && b0.f[1] == a && b0.f[2] == a && b1.f[0] == a && inf(b1.f[1]) && b1.f[2] == a && b2.f[0] == a && b2.f[1] == a && inf(b2.f[2]))) return;
        b0.op = S_CROSS; b0.f[0] = a; b0.f[1] = a; b0.f[2] = a;        // idx stays: 0 = push, 1 / 2 = fold into the distance below
        out.sdf_code.pop_back(); out.sdf_code.pop_back();
    }
    //   n x { XFORM; REP (power-of-two period); CROSS; MULS_MIN }  -> RTU_CROSS: the whole loop of the Menger recursion as one
    //                         instruction (same operations in the same order; the interpreter's fetch + decode + dispatch
    //                         was a quarter of its instructions, profiles/r2_ab.md §4)
    void fuseRtuCross(size_t first, int iterations) {
        if (const char* e = getenv("JSRT_SDF_RTU")) if (atoi(e) == 0) return;          // A/B switch
        if (!fuse || iterations < 2 || iterations > 64 || out.sdf_code.size() != first + 4 * (size_t)iterations || first < prog_first) return;
        const SdfInstr x = out.sdf_code[first], r = out.sdf_code[first + 1], c = out.sdf_code[first + 2], m = out.sdf_code[first + 3];
        if (x.op != S_XFORM || r.op != S_REP || r.idx != 1 || c.op != S_CROSS || c.idx != 0 || m.op != S_MULS_MIN) return;
        for (int i = 1; i < iterations; ++i) {
            // (every unrolled iteration pushed its own copy of the transformer's matrix: compare the matrices, not their indices)
            SdfInstr xi = out.sdf_code[first + 4 * i];
            if (xi.op != S_XFORM || memcmp(&out.xforms64[xi.idx], &out.xforms64[x.idx], sizeof(Xform64)) != 0) return;
            xi.idx = x.idx;
            if (memcmp(&xi, &x, sizeof(SdfInstr)) != 0) return;
            for (int k = 1; k < 4; ++k)
                if (memcmp(&out.sdf_code[first + 4 * i + k], &out.sdf_code[first + k], sizeof(SdfInstr)) != 0) return;
        }
        SdfInstr f{}; f.op = S_RTU_CROSS; f.idx = x.idx; f.a0 = x.a0;
        f.f[0] = r.f[0]; f.f[1] = c.f[0]; f.f[2] = (float)iterations; f.f[3] = x.f[0];
        out.sdf_code.resize(first);
        out.sdf_code.push_back(f);
    }
    void noteBase(const Val* n) {
        double c[4] = {1, 1, 1, 0};
        if (const Val* b = doc.field(n, "basecolor")) doc.vec(b, c);
        if (!base_set) { base_set = true; for (int i = 0; i < 3; ++i) base[i] = (float)c[i]; }
        else for (int i = 0; i < 3; ++i) if (base[i] != (float)c[i]) base_uniform = false;
    }

    // ---- material program -------------------------------------------------------------
    std::vector<SdfInstr> mat;                          // built aside, appended after the distance programs
    std::unordered_map<const Val*, int> dist_prog;      // node -> first instruction of its standalone distance program
    int mat_depth = 0, mat_max = 0;
    void memit(int op, int idx = 0, double a0 = 0, float f0 = 0, float f1 = 0, float f2 = 0) {
        SdfInstr i{}; i.op = op; i.idx = idx; i.a0 = a0; i.f[0] = f0; i.f[1] = f1; i.f[2] = f2;
        mat.push_back(i);
        if (op == MP_LEAF) { if (++mat_depth > mat_max) mat_max = mat_depth; }
        else if (op >= MP_SELMIN && op <= MP_BLEND_D) --mat_depth;
    }
    int distanceProgram(const Val* n) {
        n = doc.resolve(n);
        auto it = dist_prog.find(n);
        if (it != dist_prog.end()) return it->second;
        const int first = (int)out.sdf_code.size();
        depth_points = 1; depth_dist = 0; depth_scale = 1;
        prog_first = out.sdf_code.size(); scale_is_one = true; one_stack.clear();
        node(n); emit(S_END);
        return dist_prog[n] = first;
    }
    void attach(const Val* n) { memit(MP_ATTACH, distanceProgram(n)); }
    // pushes node.getMaterialData(p) (the point is NOT transformed on the way down: src/sdf.js:334-336,358-361)
    void material(const Val* n) {
        NestGuard guard(nest, kMaxNest);
        n = doc.resolve(n);
        const std::string& ty = doc.typeName(n);
        auto base = [&](float c[3]) { double v[4] = {1, 1, 1, 0}; if (const Val* b = doc.field(n, "basecolor")) doc.vec(b, v); for (int i = 0; i < 3; ++i) c[i] = (float)v[i]; };
        if (ty == "SphereSDF") { float c[3]; base(c); memit(MP_LEAF, 1, 0, c[0], c[1], c[2]); }
        else if (ty == "BoxSDF" || ty == "TetrahedronSDF") { float c[3]; base(c); memit(MP_LEAF, 0, 0, c[0], c[1], c[2]); }
        else if (ty == "UnionSDF" || ty == "IntersectionSDF") {
            const Val* cs = doc.field(n, "children");
            for (uint32_t i = 0; i < doc.length(cs); ++i) { material(doc.at(cs, i)); attach(doc.at(cs, i)); if (i) memit(ty == "UnionSDF" ? MP_SELMIN : MP_SELMAX); }
        }
        else if (ty == "DifferenceSDF") { material(doc.field(n, "positive")); attach(doc.field(n, "positive")); material(doc.field(n, "negative")); attach(doc.field(n, "negative")); memit(MP_DIFF); }
        else if (ty == "SmoothUnionSDF" || ty == "SmoothIntersectionSDF") {
            material(doc.field(n, "childA")); attach(doc.field(n, "childA")); material(doc.field(n, "childB")); attach(doc.field(n, "childB"));
            memit(ty == "SmoothUnionSDF" ? MP_BLEND_U : MP_BLEND_I, 0, doc.number(doc.field(n, "k"), 1));
        }
        else if (ty == "SmoothDifferenceSDF") {
            material(doc.field(n, "positive")); attach(doc.field(n, "positive")); material(doc.field(n, "negative")); attach(doc.field(n, "negative"));
            memit(MP_BLEND_D, 0, doc.number(doc.field(n, "k"), 1));
        }
        else if (ty == "RoundSDF" || ty == "TransformSDF") material(doc.field(n, "child_sdf"));
        else if (ty == "RecursiveTransformUnionSDF") material(doc.field(n, "sdf"));
        else fail("jsrt: unsupported SDF node '" + ty + "'");
    }

    // how many times transformer `t` multiplies its caller's scale accumulator by something other than a literal 1
    int scaleMultiplications(const Val* t) {
        NestGuard guard(nest, kMaxNest);
        t = doc.resolve(t);
        const std::string& ty = doc.typeName(t);
        if (ty == "SDFTransformerSequence") { int m = 0; const Val* ts = doc.field(t, "transformers"); for (uint32_t i = 0; i < doc.length(ts); ++i) m += scaleMultiplications(doc.at(ts, i)); return m; }
        if (ty == "SDFRecursiveTransformer") {
            const double m = std::min(doc.number(doc.field(t, "iterations"), 0), 2e9) * (double)scaleMultiplications(doc.field(t, "transformer"));
            return (int)std::min(m, 2e9);
        }
        return ty == "SDFMatrixTransformer" ? 1 : 0;
    }
    void transformer(const Val* t) {
        NestGuard guard(nest, kMaxNest);
        t = doc.resolve(t);
        const std::string& ty = doc.typeName(t);
        if (ty == "SDFTransformerSequence" || ty == "SDFRecursiveTransformer") {
            // `let s = 1; for (...) s = s * st; return [p, s]` (src/sdf.js:387-394,408-415): a local scale accumulator
            const bool seq = ty == "SDFTransformerSequence";
            const Val* ts = seq ? doc.field(t, "transformers") : nullptr;
            const int n = seq ? (int)doc.length(ts) : (int)std::min(doc.number(doc.field(t, "iterations"), 0), 2e9);
            if (n > (int)kMaxInstrs) fail("jsrt: SDF transformer iteration count out of range");
            int mults = 0;
            for (int i = 0; i < n; ++i) mults += scaleMultiplications(seq ? doc.at(ts, i) : doc.field(t, "transformer"));
            const bool elide = fuse && (mults <= 1 || scale_is_one);
            if (!elide) { emit(S_SBEGIN); one_stack.push_back(scale_is_one); scale_is_one = true; }
            for (int i = 0; i < n; ++i) transformer(seq ? doc.at(ts, i) : doc.field(t, "transformer"));
            if (!elide) { emit(S_SEND); scale_is_one = one_stack.back() && scale_is_one; one_stack.pop_back(); }
        } else if (ty == "SDFMatrixTransformer") {
            double m[16]; doc.mat4(doc.field(t, "_inv_transform"), m);
            Xform x; Xform64 y; for (int i = 0; i < 12; ++i) { x.m[i] = (float)m[i]; y.m[i] = m[i]; }
            out.xforms.push_back(x); out.xforms64.push_back(y);
            // every entry an f32 value: its product with an f32 coordinate is exact in f64 (device_math.cuh, xf64_apply_exact)
            bool exact = fuse;
            for (int i = 0; i < 12; ++i) if ((double)(float)m[i] != m[i]) exact = false;
            emit(S_XFORM, doc.number(doc.field(t, "_scale"), 1), exact ? 1.f : 0.f, 0, 0, (int)out.xforms.size() - 1);
            scale_is_one = false;
        } else if (ty == "SDFReflectionTransformer") {
            double n[4]; doc.vec(doc.field(t, "normal"), n);
            emit(S_REFL, doc.number(doc.field(t, "delta"), 0), (float)n[0], (float)n[1], (float)n[2]);
        } else if (ty == "SDFInfiniteRepetitionTransformer") {
            double s[4]; doc.vec(doc.field(t, "sizes"), s, kInf);
            // one power-of-two period on all axes: the interpreter multiplies by its exact reciprocal (a0) instead of dividing
            int e = 0;
            const float s0 = (float)s[0];
            const bool pow2 = fuse && (float)s[1] == s0 && (float)s[2] == s0 && std::isfinite(s0) && s0 > 0 && std::frexp((double)s0, &e) == 0.5;
            emit(S_REP, pow2 ? 1.0 / (double)s0 : 0.0, s0, (float)s[1], (float)s[2], pow2 ? 1 : 0);
        } else fail("jsrt: unsupported SDF transformer '" + ty + "'");
    }

    void node(const Val* n) {
        NestGuard guard(nest, kMaxNest);
        n = doc.resolve(n);
        const std::string& ty = doc.typeName(n);
        if (ty == "SphereSDF") { noteBase(n); any_sphere_leaf = true; emit(S_SPHERE, doc.number(doc.field(n, "radius"), kInf)); }
        else if (ty == "BoxSDF") { noteBase(n); double s[4]; doc.vec(doc.field(n, "size"), s, kInf); emit(S_BOX, 0, (float)s[0], (float)s[1], (float)s[2]); }
        else if (ty == "TetrahedronSDF") { noteBase(n); emit(S_TETRA); }
        else if (ty == "UnionSDF" || ty == "IntersectionSDF") {
            const Val* cs = doc.field(n, "children");
            const uint32_t cnt = doc.length(cs);
            if (!cnt) fail("jsrt: empty " + ty);
            for (uint32_t i = 0; i < cnt; ++i) { node(doc.at(cs, i)); if (i) emit(ty == "UnionSDF" ? S_MIN : S_MAX); }
        }
        else if (ty == "DifferenceSDF") { node(doc.field(n, "positive")); node(doc.field(n, "negative")); emit(S_NEG); emit(S_MAX); }
        else if (ty == "SmoothUnionSDF") { node(doc.field(n, "childA")); node(doc.field(n, "childB")); emit(S_SMIN, doc.number(doc.field(n, "k"), 1)); }
        else if (ty == "SmoothIntersectionSDF") { node(doc.field(n, "childA")); node(doc.field(n, "childB")); emit(S_SMIN_NEGAB, doc.number(doc.field(n, "k"), 1)); }
        else if (ty == "SmoothDifferenceSDF") { node(doc.field(n, "positive")); node(doc.field(n, "negative")); emit(S_SMIN_NEGA, doc.number(doc.field(n, "k"), 1)); }
        else if (ty == "RoundSDF") { node(doc.field(n, "child_sdf")); emit(S_ADDC, -doc.number(doc.field(n, "rounding"), 0)); }
        else if (ty == "TransformSDF") {
            emit(S_PUSHP); one_stack.push_back(scale_is_one); scale_is_one = true;
            transformer(doc.field(n, "transformer")); node(doc.field(n, "child_sdf")); emit(S_MULS); emit(S_POPP);
            scale_is_one = one_stack.back(); one_stack.pop_back();
        }
        else if (ty == "RecursiveTransformUnionSDF") {
            const int it = (int)std::min(doc.number(doc.field(n, "iterations"), 0), 2e9);
            if (it > (int)kMaxInstrs) fail("jsrt: RecursiveTransformUnionSDF iteration count out of range");
            emit(S_PUSHP); one_stack.push_back(scale_is_one); scale_is_one = true;
            node(doc.field(n, "sdf"));
            const size_t loop_first = out.sdf_code.size();
            for (int i = 0; i < it; ++i) { transformer(doc.field(n, "transformer")); node(doc.field(n, "sdf")); emit(S_MULS); emit(S_MIN); }
            fuseRtuCross(loop_first, it);
            emit(S_POPP);
            scale_is_one = one_stack.back(); one_stack.pop_back();
        }
        else fail("jsrt: unsupported SDF node '" + ty + "'");
    }
};

}  // namespace

int compileSdf(const WireDoc& doc, const Val* g, HostScene& out) {
    SdfCompiler c(doc, out);
    if (const char* e = getenv("JSRT_SDF_FUSE")) c.fuse = atoi(e) != 0;
    SdfProgram p{};
    p.first_instr = (int)out.sdf_code.size();
    c.prog_first = out.sdf_code.size();
    c.node(doc.field(g, "root_sdf"));
    c.emit(S_END);
    p.instr_count = (int)out.sdf_code.size() - p.first_instr;
    if (c.max_points > 8 || c.max_dist > 8 || c.max_scale > 12) fail("jsrt: SDF tree needs more interpreter stack slots than the 8 / 8 / 12 available");
    const double ms = doc.number(doc.field(g, "max_samples"), 1000);
    p.max_samples = ms > 2147483647.0 ? 2147483647 : (int)ms;
    p.distance_epsilon = doc.number(doc.field(g, "distance_epsilon"), 0.0001);
    p.max_trace_distance = doc.number(doc.field(g, "max_trace_distance"), 1000);
    p.normal_step_size = doc.number(doc.field(g, "normal_step_size"), 0.001);
    const Val* box = doc.field(g, "aabb");
    double ce[4], h[4];
    doc.vec(doc.field(box, "center"), ce); doc.vec(doc.field(box, "half_size"), h, kInf);
    p.cx = (float)ce[0]; p.cy = (float)ce[1]; p.cz = (float)ce[2]; p.hx = (float)h[0]; p.hy = (float)h[1]; p.hz = (float)h[2];
    p.uniform_base = c.base_uniform ? 1 : 0;
    for (int i = 0; i < 3; ++i) p.base[i] = c.base[i];
    p.mat_first = -1;
    if (!c.base_uniform || c.any_sphere_leaf) {
        // leaves disagree on basecolor, or a SphereSDF leaf supplies UVs: compile getMaterialData
        c.material(doc.field(g, "root_sdf"));
        c.memit(MP_END);
        if (c.mat_max > 8) fail("jsrt: SDF material program needs more than 8 stack slots");
        p.mat_first = (int)out.sdf_code.size();
        out.sdf_code.insert(out.sdf_code.end(), c.mat.begin(), c.mat.end());
        p.uniform_base = 0;
    }
    out.sdfs.push_back(p);
    return (int)out.sdfs.size() - 1;
}

}  // namespace jsrt
