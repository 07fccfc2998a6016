// Serializer object graph -> flattened scene (scene_types.h).
//
// Walks `{renderer:{world,camera,maxRecursionDepth[,samplesPerPixel]},width,height}`
// exactly as the reference's classes lay it out on the wire (key sets in
// SURVEY.md §8b; src/world.js, src/aggregates.js, src/materials.js,
// src/lights.js, src/cameras.js, src/geometry.js constructors).  Setup code:
// runs once per scene on the host, never per ray.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <unordered_map>

#include "host_scene.h"

namespace jsrt {

static const double kInf = std::numeric_limits<double>::infinity();

namespace {

struct Folded { float c1[3]; float c2[3]; bool checker; int tex = -1; float alpha_scale = 1.f; };

struct Flattener {
    const WireDoc& doc;
    HostScene& out;
    std::unordered_map<const Val*, int> ext_id;        // Primitive -> prim_id
    std::unordered_map<const Val*, int> material_of;   // Material -> index
    std::unordered_map<const Val*, int> tri_of;        // Triangle geometry -> index
    std::unordered_map<const Val*, int> sdf_of;        // SDFGeometry -> program index
    // One distinct kdtree (shared-tree instancing: several BVHAggregates may point at it).  Its node layouts are kept
    // tree-relative in depth-first visit order (hit successor = i + 1, skip = relative index, leaf = -1 for inner
    // nodes) until assembleNodes() gives every node its place in the scene's node array.
    struct TreeBuild { std::vector<std::vector<BvhNode>> layouts; int first_prim = 0, prim_count = 0, tri_base = -1, root = 0, stride = 0, n_layouts = 1; };
    std::vector<TreeBuild> trees;
    std::unordered_map<const Val*, int> tree_of;       // kdtree root -> index into trees
    std::vector<int> top_tree;                         // per entry of out.tops: its tree, or -1
    std::unordered_map<const Val*, int> tree_pure;     // kdtree root -> 1 if every leaf object is a Primitive
    bool any_tri_data = false;

    Flattener(const WireDoc& d, HostScene& o) : doc(d), out(o) {}

    // ---- transforms -----------------------------------------------------------
    static void mul44(const double a[16], const double b[16], double r[16]) {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j]; r[i * 4 + j] = s; }
    }
    int addXform(const double m[16], bool* is_identity = nullptr) {
        if (std::fabs(m[12]) > 1e-9 || std::fabs(m[13]) > 1e-9 || std::fabs(m[14]) > 1e-9 || std::fabs(m[15] - 1) > 1e-9)
            fail("jsrt: non-affine transform (fourth row is not 0 0 0 1) is not supported");
        bool ident = true;
        for (int i = 0; i < 12; ++i) if (m[i] != ((i % 5 == 0) ? 1.0 : 0.0)) ident = false;
        if (is_identity) *is_identity = ident;
        if (ident) return 0;
        Xform x; Xform64 y; for (int i = 0; i < 12; ++i) { x.m[i] = (float)m[i]; y.m[i] = m[i]; }
        out.xforms.push_back(x); out.xforms64.push_back(y);
        return (int)out.xforms.size() - 1;
    }

    // ---- material colours (src/materials.js:2-76) ------------------------------
    // nesting limits (NestGuard): material-colour / aggregate nesting, BVH depth (the reference's trees are ~log2 N deep;
    // the dragon's is 24)
    static constexpr int kMaxNest = 256, kMaxTreeDepth = 1024;
    int nest = 0, tree_nest = 0;
    Folded fold(const Val* mc) {
        NestGuard guard(nest, kMaxNest);
        mc = doc.resolve(mc);
        const std::string& t = doc.typeName(mc);
        Folded f{}; f.checker = false;
        if (t == "SolidMaterialColor") {
            double c[4]; doc.vec(doc.field(mc, "_color"), c);
            for (int i = 0; i < 3; ++i) f.c1[i] = f.c2[i] = (float)c[i];
        } else if (t == "ScaledMaterialColor") {
            f = fold(doc.field(mc, "_mc"));
            const Val* s = doc.field(mc, "_scale");
            double sv[4] = {1, 1, 1, 1};
            if (s && (s->type == Val::NUM || s->type == Val::NIL)) { double x = doc.number(s, kInf); sv[0] = sv[1] = sv[2] = x; }
            else if (s) doc.vec(s, sv, kInf);
            // `_mc.color(data).times(_scale)`: f64 product of the f32 colour, stored f32
            for (int i = 0; i < 3; ++i) { f.c1[i] = (float)((double)f.c1[i] * sv[i]); f.c2[i] = (float)((double)f.c2[i] * sv[i]); }
            // a texture colour is RGBA: a number also scales alpha, an array leaves alpha * undefined = NaN
            if (f.tex >= 0) f.alpha_scale = (s && (s->type == Val::NUM || s->type == Val::NIL)) ? (float)((double)f.alpha_scale * sv[0]) : NAN;
        } else if (t == "CheckerboardMaterialColor") {
            // Nested checkerboards see the same UV, so the outer predicate selects
            // the same side of the inner one.
            Folded a = fold(doc.field(mc, "color1")), b = fold(doc.field(mc, "color2"));
            if (a.tex >= 0 || b.tex >= 0) fail("jsrt: a TextureMaterialColor inside a CheckerboardMaterialColor is not supported");
            for (int i = 0; i < 3; ++i) { f.c1[i] = a.c1[i]; f.c2[i] = b.checker ? b.c2[i] : b.c1[i]; }
            f.checker = true;
        } else if (t == "TextureMaterialColor") {
            f.tex = texture(mc);
            for (int i = 0; i < 3; ++i) f.c1[i] = f.c2[i] = 1.f;
        } else fail("jsrt: unknown MaterialColor type '" + t + "'");
        return f;
    }
    static Color toColor(const Folded& f) {
        Color c{}; for (int i = 0; i < 3; ++i) { c.c1[i] = f.c1[i]; c.c2[i] = f.checker ? f.c2[i] : f.c1[i]; }
        c.checker = f.checker ? CK_CHECKER : CK_SOLID;
        if (f.tex >= 0) { c.checker = CK_TEXTURE; c.tex = f.tex; c.c2[0] = f.alpha_scale; }
        return c;
    }
    // may `color(data).squarednorm() > 0` hold?  (a texture's alpha counts: src/materials.js:277,284)
    static bool nonzero(const Color& c) { if (c.checker == CK_TEXTURE) return true; for (int i = 0; i < 3; ++i) if (c.c1[i] != 0 || c.c2[i] != 0) return true; return false; }

    // TextureMaterialColor -> texture table entry; `_imgdata` = {width, height, data} with data a msgpack bin or an
    // array of numbers (the wire form of ImageData is defined by this repo's glue, see jsraytracer_b200/materials.py)
    std::unordered_map<const Val*, int> texture_of;
    int texture(const Val* mc) {
        auto it = texture_of.find(mc);
        if (it != texture_of.end()) return it->second;
        const Val* img = doc.field(mc, "_imgdata");
        if (!img) fail("jsrt: TextureMaterialColor without _imgdata");
        Texture t{};
        const double tw = doc.number(doc.field(img, "width"), 0), th = doc.number(doc.field(img, "height"), 0);
        if (tw > 32768 || th > 32768) fail("jsrt: TextureMaterialColor larger than 32768 texels on a side");
        t.width = tw >= 1 ? (int)tw : 0; t.height = th >= 1 ? (int)th : 0;
        if (t.width <= 0 || t.height <= 0) fail("jsrt: TextureMaterialColor with an empty image (a browser ImageData serialises empty: SURVEY.md §8b)");
        const size_t nbytes = (size_t)t.width * t.height * 4;
        while (out.texels.size() % 16) out.texels.push_back(0);
        t.offset = out.texels.size();
        const Val* data = doc.payload(doc.field(img, "data"));
        if (!data) fail("jsrt: TextureMaterialColor: _imgdata.data is missing");
        if (data->type == Val::STR) {
            if (data->count != nbytes) fail("jsrt: TextureMaterialColor: _imgdata.data must hold width*height*4 bytes");
            out.texels.insert(out.texels.end(), (const unsigned char*)data->str, (const unsigned char*)data->str + nbytes);
        } else if (data->type == Val::ARR) {
            if (doc.length(data) != nbytes) fail("jsrt: TextureMaterialColor: _imgdata.data must hold width*height*4 numbers");
            for (uint32_t i = 0; i < nbytes; ++i) out.texels.push_back((unsigned char)doc.number(doc.at(data, i), 0));
        } else fail("jsrt: TextureMaterialColor: unsupported _imgdata.data encoding");
        const Val* mode = doc.payload(doc.field(mc, "mode"));
        if (mode && mode->type == Val::STR) {
            const std::string m(mode->str, mode->count);
            if (m == "nearest") t.flags |= TF_NEAREST;
            else if (m != "bilinear") fail("jsrt: Unsupported texture mode " + m);       // src/materials.js:119
        }
        const Val* cu = doc.field(mc, "clampU"); const Val* cv = doc.field(mc, "clampV");
        if (!cu || doc.truthy(cu)) t.flags |= TF_CLAMP_U;
        if (!cv || doc.truthy(cv)) t.flags |= TF_CLAMP_V;
        out.textures.push_back(t);
        return texture_of[mc] = (int)out.textures.size() - 1;
    }

    int material(const Val* m) {
        m = doc.resolve(m);
        auto it = material_of.find(m);
        if (it != material_of.end()) return it->second;
        const std::string& t = doc.typeName(m);
        Material d{};
        if (t == "PhongMaterial" || t == "FresnelPhongMaterial" || t == "PhongPathTracingMaterial") {
            d.kind = (t == "PhongMaterial") ? M_PHONG : (t == "FresnelPhongMaterial" ? M_FRESNEL : M_PATH);
            d.ambient = toColor(fold(doc.field(m, "ambient")));
            d.diffusivity = toColor(fold(doc.field(m, "diffusivity")));
            d.specularity = toColor(fold(doc.field(m, "specularity")));
            d.reflectivity = toColor(fold(doc.field(m, "reflectivity")));
            d.transmissivity = toColor(fold(doc.field(m, "transmissivity")));
            d.smoothness = (float)doc.number(doc.field(m, "smoothness"), kInf);
            d.ior = (float)(d.kind == M_PHONG ? kInf : doc.number(doc.field(m, "refractiveIndexRatio"), kInf));
            d.mirror_prob = (float)(d.kind == M_PATH ? doc.number(doc.field(m, "mirrorProbability"), 0) : 0);
            if (d.kind == M_PATH && !std::isfinite(d.smoothness))
                fail("jsrt: PhongPathTracingMaterial with infinite smoothness is not executable in the reference (src/materials.js:447-474)");
            if (d.kind == M_PHONG) { if (nonzero(d.reflectivity) && nonzero(d.transmissivity)) out.fanout = 2; }
            else if (std::isfinite(d.ior)) out.fanout = 2;
        } else if (t == "SolidColorMaterial") {
            d.kind = M_SOLID; d.ambient = toColor(fold(doc.field(m, "_color")));
        } else if (t == "TransparentMaterial") {
            d.kind = M_TRANSPARENT; d.ambient = toColor(fold(doc.field(m, "_color")));
            d.smoothness = (float)doc.number(doc.field(m, "_opacity"), 1);
        } else if (t == "PositionalUVMaterial") {
            // the innermost PositionalUVMaterial of a chain wins (each overwrites data.UV before calling its base)
            d = out.materials[material(doc.field(m, "baseMaterial"))];
            if (!d.uv_from_position) {
                d.uv_from_position = 1;
                double o[4] = {0, 0, 0, 0}, u[4] = {1, 0, 0, 0}, v[4] = {0, 0, 1, 0};
                doc.vec(doc.field(m, "origin"), o); doc.vec(doc.field(m, "u_axis"), u); doc.vec(doc.field(m, "v_axis"), v);
                for (int i = 0; i < 3; ++i) { d.uv_origin[i] = (float)o[i]; d.u_axis[i] = (float)u[i]; d.v_axis[i] = (float)v[i]; }
            }
        } else fail("jsrt: unsupported material type '" + t + "'");
        out.materials.push_back(d);
        return material_of[m] = (int)out.materials.size() - 1;
    }

    // ---- geometry -----------------------------------------------------------------
    int triangle(const Val* g, int* flags) {
        auto it = tri_of.find(g);
        int idx;
        if (it != tri_of.end()) idx = it->second;
        else {
            const Val* ps = doc.field(g, "ps");
            if (doc.length(ps) != 3) fail("jsrt: Triangle.ps must hold 3 points");
            float p[3][4];
            for (int i = 0; i < 3; ++i) { double v[4]; doc.vec(doc.at(ps, i), v); for (int k = 0; k < 4; ++k) p[i][k] = (float)v[k]; }
            // Triangle constructor, src/geometry.js:341-353 (f64 ops, f32 stores)
            float v0[3], v1[3], h[3], n[3];
            for (int k = 0; k < 3; ++k) { v0[k] = (float)((double)p[1][k] - (double)p[0][k]); v1[k] = (float)((double)p[2][k] - (double)p[0][k]); }
            h[0] = (float)((double)v0[1] * v1[2] - (double)v0[2] * v1[1]);
            h[1] = (float)((double)v0[2] * v1[0] - (double)v0[0] * v1[2]);
            h[2] = (float)((double)v0[0] * v1[1] - (double)v0[1] * v1[0]);
            const double nn = std::sqrt((double)h[0] * h[0] + (double)h[1] * h[1] + (double)h[2] * h[2]);
            for (int k = 0; k < 3; ++k) n[k] = (nn > 0.00001) ? (float)((double)h[k] * (1 / nn)) : h[k];
            Tri t{};
            t.nx = n[0]; t.ny = n[1]; t.nz = n[2];
            t.delta = (float)((double)n[0] * p[0][0] + (double)n[1] * p[0][1] + (double)n[2] * p[0][2] + 0.0 * p[0][3]);
            t.p0x = p[0][0]; t.p0y = p[0][1]; t.p0z = p[0][2];
            t.v0x = v0[0]; t.v0y = v0[1]; t.v0z = v0[2];
            t.v1x = v1[0]; t.v1y = v1[1]; t.v1z = v1[2];
            t.d00 = (double)v0[0] * v0[0] + (double)v0[1] * v0[1] + (double)v0[2] * v0[2];
            t.d11 = (double)v1[0] * v1[0] + (double)v1[1] * v1[1] + (double)v1[2] * v1[2];
            t.d01 = (double)v0[0] * v1[0] + (double)v0[1] * v1[1] + (double)v0[2] * v1[2];
            t.inv_denom = 1.0 / (t.d00 * t.d11 - t.d01 * t.d01);
            // f32 fast path only for well-conditioned triangles: sin^2 of the corner angle = denom / (d00 d11)
            const double cond = (t.d00 * t.d11 - t.d01 * t.d01) / (t.d00 * t.d11);
            for (int k = 0; k < 3; ++k) {
                (&t.ax)[k] = (float)((t.d11 * v0[k] - t.d01 * v1[k]) * t.inv_denom);
                (&t.bx)[k] = (float)((t.d00 * v1[k] - t.d01 * v0[k]) * t.inv_denom);
            }
            if (!(cond > 1e-2)) t.ax = t.bx = std::numeric_limits<float>::quiet_NaN();      // u, v, w all NaN: every comparison fails
            TriShade s{};
            // psdata: {UV:[..], normal:[..]} — an Array here means the reference's
            // lossy Triangle.serialize (src/geometry.js:355-357) wrote `ps` twice.
            const Val* pd = doc.field(g, "psdata");
            int fl = 0;
            if (pd && doc.payload(pd) && doc.payload(pd)->type == Val::MAP) {
                if (const Val* nv = doc.field(pd, "normal")) {
                    if (doc.length(nv) == 3) { fl |= PF_HAS_VNORMALS; for (int i = 0; i < 3; ++i) { double v[4]; doc.vec(doc.at(nv, i), v); for (int k = 0; k < 4; ++k) s.n[i][k] = (float)v[k]; } }
                }
                if (const Val* uv = doc.field(pd, "UV")) {
                    if (doc.length(uv) == 3) { fl |= PF_HAS_UVS; for (int i = 0; i < 3; ++i) { double v[4]; doc.vec(doc.at(uv, i), v); s.uv[i][0] = (float)v[0]; s.uv[i][1] = (float)v[1]; } }
                }
            }
            s.pad[0] = (float)fl;
            if (fl) any_tri_data = true;
            out.tris.push_back(t);
            out.tri_shade.push_back(s);
            idx = tri_of[g] = (int)out.tris.size() - 1;
        }
        *flags |= (int)out.tri_shade[idx].pad[0];
        return idx;
    }

    // Appends one placed primitive.  `outer` (may be null) is an extra inverse
    // transform applied before the primitive's own (flattened nested Aggregates).
    void placePrim(const Val* p, const double* outer) {
        p = doc.resolve(p);
        if (doc.typeName(p) != "Primitive")
            fail("jsrt: '" + doc.typeName(p) + "' inside an aggregate is not supported (only Primitive members)");
        Prim d{};
        const Val* g = doc.field(p, "geometry");
        const std::string& gt = doc.typeName(g);
        d.geom_index = -1;
        if (gt == "Plane" || gt == "SimplePlane") d.geom_kind = G_PLANE;
        else if (gt == "Square") d.geom_kind = G_SQUARE;
        else if (gt == "Circle") d.geom_kind = G_CIRCLE;
        else if (gt == "UnitBox") d.geom_kind = G_BOX;
        else if (gt == "AABB") {
            d.geom_kind = G_BOX; double c[4], h[4];
            doc.vec(doc.field(g, "center"), c); doc.vec(doc.field(g, "half_size"), h, kInf);
            d.geom_index = (int)out.boxes.size() / 8;
            for (int i = 0; i < 4; ++i) out.boxes.push_back((float)c[i]);
            for (int i = 0; i < 4; ++i) out.boxes.push_back((float)h[i]);
        }
        else if (gt == "Sphere") d.geom_kind = G_SPHERE;
        else if (gt == "Cylinder") d.geom_kind = G_CYLINDER;
        else if (gt == "Triangle") { d.geom_kind = G_TRIANGLE; d.geom_index = triangle(g, &d.flags); }
        else if (gt == "SDFGeometry") {
            d.geom_kind = G_SDF;
            auto it = sdf_of.find(g);
            d.geom_index = (it != sdf_of.end()) ? it->second : (sdf_of[g] = compileSdf(doc, g, out));
        }
        else if (gt == "OriginPoint" || gt == "UnitLine") d.geom_kind = G_NEVER;   // intersect() returns -Infinity (src/geometry.js:34-36,59-61)
        else fail("jsrt: unsupported geometry type '" + gt + "'");
        d.material = material(doc.field(p, "material"));
        double inv[16]; doc.mat4(doc.field(p, "inv_transform"), inv);
        bool ident = false;
        if (outer) { double c[16]; mul44(inv, outer, c); d.xform = addXform(c, &ident); }
        else d.xform = addXform(inv, &ident);
        if (ident) d.flags |= PF_IDENTITY_XFORM;
        const Val* cs = doc.field(p, "does_cast_shadow");
        if (!cs || doc.truthy(cs)) d.flags |= PF_CASTS_SHADOW;
        auto e = ext_id.find(p);
        if (e == ext_id.end()) fail("jsrt: primitive reachable from the BVH but not from the aggregate's objects array");
        d.ext_id = e->second;
        out.prims.push_back(d);
    }

    // prim_id assignment: DFS of world.objects descending into Aggregate.objects
    void assignIds(const Val* objects) {
        NestGuard guard(nest, kMaxNest);
        for (uint32_t i = 0; i < doc.length(objects); ++i) {
            const Val* o = doc.at(objects, i);
            const std::string& t = doc.typeName(o);
            if (t == "Primitive") { if (!ext_id.count(o)) { int id = (int)ext_id.size(); ext_id[o] = id; } }
            else if (t == "Aggregate" || t == "BVHAggregate") assignIds(doc.field(o, "objects"));
        }
    }

    // ---- BVH layout in reference visit order (src/aggregates.js:207-225) ------------
    // Number of objects below node `n` if they are all Triangles and there are at most `cap` of them, else cap + 1.
    int smallTriSubtree(const Val* n, int cap) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (doc.truthy(doc.field(n, "isLeaf"))) {
            const Val* objs = doc.field(n, "objects");
            const uint32_t cnt = objs ? doc.length(objs) : 0;
            if ((int)cnt > cap) return cap + 1;
            for (uint32_t i = 0; i < cnt; ++i) {
                const Val* p = doc.resolve(doc.at(objs, i));
                if (doc.typeName(p) != "Primitive" || doc.typeName(doc.field(p, "geometry")) != "Triangle") return cap + 1;
            }
            return (int)cnt;
        }
        const int g = smallTriSubtree(doc.field(n, "greater_node"), cap);
        if (g > cap) return cap + 1;
        const int l = smallTriSubtree(doc.field(n, "lesser_node"), cap - g);
        return (g + l > cap) ? cap + 1 : g + l;
    }
    // the objects below `n` in the reference's visit order (greater child first)
    void placeSubtree(const Val* n) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (doc.truthy(doc.field(n, "isLeaf"))) {
            const Val* objs = doc.field(n, "objects");
            const uint32_t cnt = objs ? doc.length(objs) : 0;
            for (uint32_t i = 0; i < cnt; ++i) placePrim(doc.at(objs, i), nullptr);
        } else {
            placeSubtree(doc.field(n, "greater_node"));
            placeSubtree(doc.field(n, "lesser_node"));
        }
    }
    // The reference's leaves hold one object each (src/aggregates.js:66,87-89), so half of its nodes are boxes
    // around single triangles.  A subtree with at most `leaf_tris` triangles is emitted as ONE leaf (its root's
    // box, its triangles in visit order): the walk tests those few triangles directly instead of first walking
    // 2k-1 more boxes.  Same hits (every triangle the reference would test is still tested, ties still go to the
    // earlier-visited one); fewer dependent node fetches.  JSRT_LEAF_TRIS=1 keeps the reference topology.
    int leaf_tris = 2;                 // measured: profiles/r1_ab.md (2 is the best of 1..8 on bunny_path, dragon and starwars)
    int collapsed_nodes = 0;           // reference nodes folded into multi-triangle leaves
    void refStats(const Val* n, int depth) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (depth > out.max_bvh_depth) out.max_bvh_depth = depth;
        if (doc.truthy(doc.field(n, "isLeaf"))) return;
        collapsed_nodes += 2;
        refStats(doc.field(n, "greater_node"), depth + 1);
        refStats(doc.field(n, "lesser_node"), depth + 1);
    }
    void layoutNode(const Val* n, std::vector<BvhNode>& nodes, int first_prim, int depth) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (depth > out.max_bvh_depth) out.max_bvh_depth = depth;
        const int me = (int)nodes.size();
        BvhNode b{};
        const Val* box = doc.field(n, "aabb");
        double c[4], h[4];
        doc.vec(doc.field(box, "center"), c); doc.vec(doc.field(box, "half_size"), h, kInf);
        b.cx = (float)c[0]; b.cy = (float)c[1]; b.cz = (float)c[2]; b.hx = (float)h[0]; b.hy = (float)h[1]; b.hz = (float)h[2];
        b.leaf = -1;
        nodes.push_back(b);
        const bool is_leaf = doc.truthy(doc.field(n, "isLeaf"));
        if (is_leaf || (leaf_tris > 1 && smallTriSubtree(n, leaf_tris) <= leaf_tris)) {
            const int first = (int)out.prims.size() - first_prim;
            if (first >= (1 << 24)) fail("jsrt: more than 16M primitives in one BVH");
            placeSubtree(n);
            if (!is_leaf) refStats(n, depth);          // Info.n_nodes / max_bvh_depth keep describing the reference's tree
            const uint32_t cnt = (uint32_t)((int)out.prims.size() - first_prim - first);
            if (cnt > 127) fail("jsrt: BVH leaf with more than 127 objects");
            nodes[me].leaf = (int)((cnt << 24) | (uint32_t)first);
        } else {
            layoutNode(doc.field(n, "greater_node"), nodes, first_prim, depth + 1);
            layoutNode(doc.field(n, "lesser_node"), nodes, first_prim, depth + 1);
        }
        nodes[me].skip = (int)nodes.size();
    }

    // Octant layouts: re-emit the tree rooted at layout-0 node `i` (children: i+1 = greater,
    // skip(i+1) = lesser) visiting first the child that a ray of direction octant `q` meets
    // first.  The split axis is not on the wire; the axis along which the children's box
    // centres differ most stands in for it (any order is valid, this one prunes well).
    void emitOctant(const std::vector<BvhNode>& ref, int i, int q, std::vector<BvhNode>& dst) {
        const int me = (int)dst.size();
        dst.push_back(ref[i]);
        if (ref[i].leaf == -1) {
            const int g = i + 1, l = ref[g].skip;
            const float dc[3] = {ref[g].cx - ref[l].cx, ref[g].cy - ref[l].cy, ref[g].cz - ref[l].cz};
            int ax = 0; if (std::fabs(dc[1]) > std::fabs(dc[ax])) ax = 1; if (std::fabs(dc[2]) > std::fabs(dc[ax])) ax = 2;
            const bool neg = (q >> ax) & 1;                 // ray travels towards -axis
            const bool greater_is_high = dc[ax] >= 0;
            const bool greater_first = (neg == greater_is_high);
            emitOctant(ref, greater_first ? g : l, q, dst);
            emitOctant(ref, greater_first ? l : g, q, dst);
        }
        dst[me].skip = (int)dst.size();
    }

    // ---- world.objects -> flattened top-level entries ------------------------------------
    // The reference nests freely: an Aggregate's or a BVHAggregate's members may be Primitives or further aggregates
    // (Aggregate.intersect / BVHAggregate.intersect push themselves onto `ancestors`, src/aggregates.js:14-18,43-49), and
    // every level maps the ray by its own inv_transform.  On the device there are only three kinds of entries — a
    // Primitive, a run of Primitives sharing an enclosing Aggregate (T_LIST), and a BVHAggregate whose leaves are all
    // Primitives (T_BVH) — emitted in the reference's visit order (world.objects order, members in array order, the
    // objects of a BVHAggregate with non-Primitive members in its tree's leaf visit order).  Visit order is what decides
    // exact ties (`<` in src/world.js:9-13, src/aggregates.js:213), so "lower entry index, then lower primitive index"
    // stays the tie rule.  Transforms of enclosing aggregates are folded into one matrix per entry / primitive.
    bool treeIsPure(const Val* n) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (doc.truthy(doc.field(n, "isLeaf"))) {
            const Val* objs = doc.field(n, "objects");
            for (uint32_t i = 0; objs && i < doc.length(objs); ++i)
                if (doc.typeName(doc.at(objs, i)) != "Primitive") return false;
            return true;
        }
        return treeIsPure(doc.field(n, "greater_node")) && treeIsPure(doc.field(n, "lesser_node"));
    }
    bool pureBvh(const Val* agg) {
        const Val* tree = doc.field(agg, "kdtree");
        if (!tree) fail("jsrt: BVHAggregate without kdtree");
        auto it = tree_pure.find(tree);
        if (it == tree_pure.end()) it = tree_pure.emplace(tree, treeIsPure(tree) ? 1 : 0).first;
        return it->second != 0;
    }
    int world_index = 0;               // entry of world.objects being flattened
    int open_list = -1;                // index into out.tops of the T_LIST run that is collecting primitives, or -1
    void closeList() {
        if (open_list >= 0) out.tops[open_list].prim_count = (int)out.prims.size() - out.tops[open_list].first_prim;
        open_list = -1;
    }
    void pushTop(const Top& t, int tree) { out.tops.push_back(t); top_tree.push_back(tree); out.top_world.push_back(world_index); }

    // One BVHAggregate whose leaf objects are all Primitives.  `inv` = its inv_transform composed with every enclosing aggregate's.
    void emitBvh(const Val* o, const double inv[16]) {
        closeList();
        Top top{}; top.kind = T_BVH; top.tri_base = -1; top.layouts = 1;
        top.xform = addXform(inv);
        const Val* tree = doc.field(o, "kdtree");
        auto it = tree_of.find(tree);
        if (it == tree_of.end()) {
            TreeBuild tb;
            tb.first_prim = (int)out.prims.size();
            const int first_tri = (int)out.tris.size();
            collapsed_nodes = 0;
            tb.layouts.emplace_back();
            layoutNode(tree, tb.layouts[0], tb.first_prim, 0);
            const int node_count = (int)tb.layouts[0].size();
            tb.prim_count = (int)out.prims.size() - tb.first_prim;
            // mesh fast path: all leaf objects are fresh identity-transform shadow-casting triangles in leaf order
            bool pure = tb.prim_count > 0 && (int)out.tris.size() - first_tri == tb.prim_count;
            for (int k = 0; pure && k < tb.prim_count; ++k) {
                const Prim& p = out.prims[tb.first_prim + k];
                pure = p.geom_kind == G_TRIANGLE && p.geom_index == first_tri + k && (p.flags & PF_IDENTITY_XFORM) && (p.flags & PF_CASTS_SHADOW);
            }
            tb.tri_base = pure ? first_tri : -1;
            out.tree_node_count += node_count + collapsed_nodes; collapsed_nodes = 0;
            // Eight octant layouts (closest-hit rays use the layout of their direction octant, shadow rays layout 7,
            // which is the reference's greater-child-first order).  Pays on big meshes (dragon 1080p: extend 11.9 -> 9.8 ms per 16
            // passes); on small ones the reference order is as good and 8 layouts only cost L1 hits (bunny:
            // 10.6 -> 10.8 ms).  Default: trees with >= 32768 nodes; JSRT_OCTANT_LAYOUTS=0/1 forces it.
            bool octants = node_count >= 32768;
            if (const char* e = getenv("JSRT_OCTANT_LAYOUTS")) octants = atoi(e) != 0;
            if (node_count > 1 && octants) {
                const std::vector<BvhNode> ref = tb.layouts[0];
                tb.layouts.clear();
                for (int q = 0; q < 8; ++q) {
                    tb.layouts.emplace_back(); tb.layouts.back().reserve(ref.size());
                    emitOctant(ref, 0, q, tb.layouts.back());
                }
            }
            trees.push_back(std::move(tb));
            it = tree_of.emplace(tree, (int)trees.size() - 1).first;
        }
        const TreeBuild& tb = trees[it->second];
        top.first_prim = tb.first_prim; top.prim_count = tb.prim_count; top.tri_base = tb.tri_base;
        top.node_count = (int)tb.layouts[0].size();
        pushTop(top, it->second);
    }

    // the objects below node `n` of a BVHAggregate with non-Primitive members, in the reference's visit order
    void mixedTreeMembers(const Val* n, const double* rel, const double* root_inv) {
        NestGuard guard(tree_nest, kMaxTreeDepth);
        n = doc.resolve(n);
        if (doc.truthy(doc.field(n, "isLeaf"))) {
            const Val* objs = doc.field(n, "objects");
            for (uint32_t i = 0; objs && i < doc.length(objs); ++i) member(doc.at(objs, i), rel, root_inv);
        } else {
            mixedTreeMembers(doc.field(n, "greater_node"), rel, root_inv);
            mixedTreeMembers(doc.field(n, "lesser_node"), rel, root_inv);
        }
    }
    // One member of an aggregate.  `root_inv`: inv_transform of the outermost enclosing aggregate (kept apart, like the
    // reference's first ray.getTransformed, for the primitives of its T_LIST runs); `rel`: the inv_transforms of the
    // aggregates between that one and this member, composed (null = identity).
    void member(const Val* o, const double* rel, const double* root_inv) {
        NestGuard guard(nest, kMaxNest);
        const std::string& t = doc.typeName(o);
        if (t == "Primitive") {
            if (open_list < 0) {
                Top top{}; top.kind = T_LIST; top.tri_base = -1; top.layouts = 1;
                top.xform = addXform(root_inv);
                top.first_prim = (int)out.prims.size();
                pushTop(top, -1);
                open_list = (int)out.tops.size() - 1;
            }
            placePrim(o, rel);
        } else if (t == "Aggregate" || t == "BVHAggregate") {
            double inner[16], c[16]; doc.mat4(doc.field(o, "inv_transform"), inner);
            if (rel) mul44(inner, rel, c); else memcpy(c, inner, sizeof c);
            if (t == "BVHAggregate" && pureBvh(o)) { double full[16]; mul44(c, root_inv, full); emitBvh(o, full); }
            else if (t == "BVHAggregate") mixedTreeMembers(doc.field(o, "kdtree"), c, root_inv);
            else { const Val* objs = doc.field(o, "objects"); for (uint32_t i = 0; objs && i < doc.length(objs); ++i) member(doc.at(objs, i), c, root_inv); }
        } else fail("jsrt: '" + t + "' inside an aggregate is not supported");
    }

    // Gives every node of every tree layout its place in the scene's node array and turns the tree-relative links into
    // absolute ones.  Front block: the first `take` nodes of every layout in breadth-first order — the top levels every
    // ray walks through, staged in shared memory by bvh_kernel (DeviceScene::n_staged); behind it the remaining nodes of
    // each layout in depth-first order (siblings' subtrees stay close together for the L1 / L2 lines).
    void assembleNodes() {
        size_t n_layouts = 0, total = 0;
        for (const TreeBuild& tb : trees) { n_layouts += tb.layouts.size(); for (auto& l : tb.layouts) total += l.size(); }
        // Nodes in the staged block (JSRT_STAGE_NODES overrides): everything when the scene's trees fit the 224 KB a CTA can
        // have (bunny_path: 5.5 k nodes, +2 % over a 4 096-node block), otherwise 4 096 = 128 KB, which leaves the other
        // half of the SM's unified array to L1 for the deep nodes (dragon, 200 k nodes: 7 168 staged is 2.3 % slower than
        // 4 096, 2 048 within noise of it: profiles/r2/ab_r2j_stage_*)
        int budget = total <= 7168 ? 7168 : 4096;
        if (const char* e = getenv("JSRT_STAGE_NODES")) { const int v = atoi(e); if (v >= 0 && v <= 7168) budget = v; }
        if (total >= (size_t)kNodeEnd) fail("jsrt: more than 2^31 BVH nodes");
        out.nodes.assign(total, BvhNode{});
        // every layout owns at least its root in the front block (the root of layout q is found at root + q * stride)
        std::vector<int> take(trees.size(), 0);
        {
            long long left = budget;
            for (size_t t = 0; t < trees.size(); ++t) { take[t] = 1; left -= (long long)trees[t].layouts.size(); }
            // hand the rest out in rounds, one level-ish at a time, so that small trees do not strand budget
            bool grew = true;
            while (left > 0 && grew) {
                grew = false;
                for (size_t t = 0; t < trees.size() && left > 0; ++t) {
                    const int size = (int)trees[t].layouts[0].size(), nl = (int)trees[t].layouts.size();
                    const int want = std::min(size, take[t] * 2 + 1) - take[t];
                    const int can = (int)std::min<long long>(want, left / nl);
                    if (can > 0) { take[t] += can; left -= (long long)can * nl; grew = true; }
                }
            }
        }
        size_t front = 0;
        for (size_t t = 0; t < trees.size(); ++t) { trees[t].root = (int)front; trees[t].stride = take[t]; front += (size_t)take[t] * trees[t].layouts.size(); }
        out.n_staged = (int)front;
        size_t rest = front;
        std::vector<int> place, bfs;
        for (size_t t = 0; t < trees.size(); ++t) {
            TreeBuild& tb = trees[t];
            for (size_t q = 0; q < tb.layouts.size(); ++q) {
                const std::vector<BvhNode>& L = tb.layouts[q];
                const int n = (int)L.size();
                place.assign(n, -1); bfs.clear();
                bfs.push_back(0);
                for (size_t h = 0; h < bfs.size() && (int)bfs.size() < take[t]; ++h) {
                    const int i = bfs[h];
                    if (L[i].leaf != -1) continue;
                    bfs.push_back(i + 1);
                    if ((int)bfs.size() < take[t]) bfs.push_back(L[i + 1].skip);
                }
                const int base = tb.root + (int)q * tb.stride;
                for (size_t k = 0; k < bfs.size(); ++k) place[bfs[k]] = base + (int)k;
                // (bfs.size() == take[t] unless the loop ran out of inner nodes, which cannot happen for take <= n)
                for (int i = 0; i < n; ++i) if (place[i] < 0) place[i] = (int)rest++;
                for (int i = 0; i < n; ++i) {
                    BvhNode b = L[i];
                    b.skip = (L[i].skip >= n) ? kNodeEnd : place[L[i].skip];
                    if (L[i].leaf == -1) b.leaf = kNodeInner | place[i + 1];
                    out.nodes[place[i]] = b;
                }
            }
            tb.n_layouts = (int)tb.layouts.size();
            tb.layouts.clear(); tb.layouts.shrink_to_fit();
        }
        for (size_t i = 0; i < out.tops.size(); ++i) {
            if (top_tree[i] < 0) continue;
            const TreeBuild& tb = trees[top_tree[i]];
            out.tops[i].first_node = tb.root;
        }
    }

    void run() {
        if (const char* e = getenv("JSRT_LEAF_TRIS")) { const int v = atoi(e); if (v >= 1 && v <= 64) leaf_tris = v; }
        const Val* root = doc.resolve(doc.root());
        const Val* rend = doc.field(root, "renderer");
        if (!rend) fail("jsrt: scene blob has no 'renderer' (expected Serializer({renderer,width,height}))");
        out.renderer_type = doc.typeName(rend);
        out.jitter = out.renderer_type != "SimpleRenderer";
        // header numbers are range-checked before the casts (a double outside int's range is undefined behaviour in C++)
        auto ranged = [&](const Val* v, double dflt, double lo, double hi, const char* what) {
            const double x = doc.number(v, dflt);
            if (!(x >= lo && x <= hi)) fail(std::string("jsrt: ") + what + " out of range");
            return (int)x;
        };
        if (!doc.field(root, "width") || !doc.field(root, "height")) fail("jsrt: scene blob has no positive width/height");
        out.width = ranged(doc.field(root, "width"), 0, 1, 65536, "width");
        out.height = ranged(doc.field(root, "height"), 0, 1, 65536, "height");
        if ((long long)out.width * out.height > (1LL << 28)) fail("jsrt: image larger than 2^28 pixels");
        out.max_depth = ranged(doc.field(rend, "maxRecursionDepth"), 3, -1e9, 1e9, "maxRecursionDepth");      // jsrt_render accepts 1..255
        out.samples_per_pixel = doc.field(rend, "samplesPerPixel") ? ranged(doc.field(rend, "samplesPerPixel"), 1, -1e9, 1e9, "samplesPerPixel") : 1;

        const Val* cam = doc.field(rend, "camera");
        if (!cam) fail("jsrt: renderer has no camera");
        double t[16]; doc.mat4(doc.field(cam, "transform"), t);
        for (int i = 0; i < 12; ++i) out.camera.t[i] = t[i];
        out.camera.tan_fov = doc.number(doc.field(cam, "tan_fov"), 0);
        out.camera.aspect = doc.number(doc.field(cam, "aspect"), 1);
        out.camera.dof = doc.typeName(cam) == "DepthOfFieldPerspectiveCamera";
        if (out.camera.dof) {
            out.camera.focus_distance = doc.number(doc.field(cam, "focus_distance"), 0);
            out.camera.sensor_size = doc.number(doc.field(cam, "sensor_size"), 0);
        }

        const Val* world = doc.field(rend, "world");
        if (!world) fail("jsrt: renderer has no world");
        double bg[4]; doc.vec(doc.field(world, "bg_color"), bg);
        for (int i = 0; i < 3; ++i) out.bg[i] = (float)bg[i];

        Xform id{}; id.m[0] = id.m[5] = id.m[10] = 1; out.xforms.push_back(id);
        Xform64 id64{}; id64.m[0] = id64.m[5] = id64.m[10] = 1; out.xforms64.push_back(id64);

        const Val* objects = doc.field(world, "objects");
        assignIds(objects);
        out.ext_prim_count = (int)ext_id.size();
        out.world_object_count = (int)doc.length(objects);
        for (uint32_t i = 0; i < doc.length(objects); ++i) {
            const Val* o = doc.at(objects, i);
            const std::string& ty = doc.typeName(o);
            world_index = (int)i;
            if (ty == "Primitive") {
                Top top{}; top.tri_base = -1; top.layouts = 1;
                top.kind = T_PRIM; top.first_prim = (int)out.prims.size(); top.prim_count = 1;
                placePrim(o, nullptr);
                if (out.prims.back().geom_kind == G_SDF) top.kind = T_SDF;
                pushTop(top, -1);
            } else if (ty == "BVHAggregate" || ty == "Aggregate") {
                double inv[16]; doc.mat4(doc.field(o, "inv_transform"), inv);
                if (ty == "BVHAggregate" && pureBvh(o)) emitBvh(o, inv);
                else if (ty == "BVHAggregate") mixedTreeMembers(doc.field(o, "kdtree"), nullptr, inv);
                else { const Val* objs = doc.field(o, "objects"); for (uint32_t k = 0; objs && k < doc.length(objs); ++k) member(doc.at(objs, k), nullptr, inv); }
                closeList();
                // (an aggregate without members leaves no entry: it can never be hit)
            } else fail("jsrt: unsupported world object type '" + ty + "'");
        }
        assembleNodes();
        // the layout word of every BVH entry: 8 link orders or 1, and the distance between their roots
        for (size_t i = 0; i < out.tops.size(); ++i)
            if (top_tree[i] >= 0) {
                const TreeBuild& tb = trees[top_tree[i]];
                const int nl = (out.tops[i].node_count > 1 && tb.stride > 0 && tb.n_layouts == 8) ? 8 : 1;
                out.tops[i].layouts = nl | (tb.stride << 8);
            }
        if (!any_tri_data) out.tri_shade.clear();

        const Val* lights = doc.field(world, "lights");
        for (uint32_t i = 0; lights && i < doc.length(lights); ++i) {
            const Val* l = doc.at(lights, i);
            const std::string& ty = doc.typeName(l);
            Light d{};
            Folded col = fold(doc.field(l, "color_mc"));
            if (col.checker) fail("jsrt: checkerboard light colours are not supported");
            for (int k = 0; k < 3; ++k) d.color[k] = col.c1[k];
            if (ty == "SimplePointLight") {
                d.kind = L_POINT; d.samples = 1;
                double p[4]; doc.vec(doc.field(l, "position"), p);
                for (int k = 0; k < 4; ++k) d.pos[k] = (float)p[k];
            } else if (ty == "RandomSampleAreaLight") {
                const double ns = doc.number(doc.field(l, "samples"), 1);
                if (!(ns >= 0 && ns <= 65536)) fail("jsrt: RandomSampleAreaLight.samples out of range");
                d.kind = L_AREA; d.samples = (int)ns;
                const std::string& gt = doc.typeName(doc.field(l, "surface_geometry"));
                if (gt == "Square") d.geom = G_SQUARE; else if (gt == "Circle") d.geom = G_CIRCLE; else if (gt == "Sphere") d.geom = G_SPHERE;
                else fail("jsrt: area light surface '" + gt + "' has no sampleSurface in the reference");
                double m[16];
                doc.mat4(doc.field(l, "transform"), m); for (int k = 0; k < 12; ++k) d.xf.m[k] = (float)m[k];
                doc.mat4(doc.field(l, "inv_transform"), m); for (int k = 0; k < 12; ++k) d.inv.m[k] = (float)m[k];
            } else fail("jsrt: unsupported light type '" + ty + "'");
            out.light_samples += d.samples;
            out.lights.push_back(d);
        }
    }
};

}  // namespace

void flattenScene(const WireDoc& doc, HostScene& out) {
    Flattener f(doc, out);
    f.run();
    // a top-level BVH over the aggregates once there are enough of them for it to beat the linear walk (JSRT_TLAS_MIN, 0 = never)
    int n_bvh = 0;
    for (const Top& t : out.tops) if (t.kind == T_BVH && t.node_count > 0) ++n_bvh;
    // Measured on the 27-instance dragon grid (profiles/r2_ab.md): the in-walk top-level BVH is 4 % SLOWER than the linear walk
    // behind the world-box reject (27 cheap, coherent slab tests in prims_kernel against divergent level changes in the
    // walk, each re-reading the world ray), so it is only used where the linear cost must lose: from 48 aggregates on.
    int tlas_min = 48;
    if (const char* e = getenv("JSRT_TLAS_MIN")) tlas_min = atoi(e);
    if (tlas_min > 0 && n_bvh >= tlas_min) out.tlas_root = buildTlas(out);
}

// The root box (centre c, half h, aggregate space) under the aggregate's transform M = inv_transform^-1 has centre M c and
// half |M3x3| h; the pad (1e-3 of the box, 1e-4 of its distance from the origin) is far above the f32 rounding of
// either the world-space or the reference's local test, so a ray that misses the padded box misses the local one.
// A singular / non-finite transform gives an infinite box (never rejects).
static void worldBoxOf(const double* a /* rows of the affine inv_transform */, const float cf[3], const float hf[3], std::vector<float>& out) {
    const double det = a[0] * (a[5] * a[10] - a[6] * a[9]) - a[1] * (a[4] * a[10] - a[6] * a[8]) + a[2] * (a[4] * a[9] - a[5] * a[8]);
    const double m[9] = {(a[5] * a[10] - a[6] * a[9]) / det, (a[2] * a[9] - a[1] * a[10]) / det, (a[1] * a[6] - a[2] * a[5]) / det,
                         (a[6] * a[8] - a[4] * a[10]) / det, (a[0] * a[10] - a[2] * a[8]) / det, (a[2] * a[4] - a[0] * a[6]) / det,
                         (a[4] * a[9] - a[5] * a[8]) / det, (a[1] * a[8] - a[0] * a[9]) / det, (a[0] * a[5] - a[1] * a[4]) / det};
    const double c[3] = {cf[0] - a[3], cf[1] - a[7], cf[2] - a[11]};      // M c = A^-1 (c - translation of inv_transform)
    const double h[3] = {hf[0], hf[1], hf[2]};
    double wc[3], wh[3], hmax = 0, cmax = 0;
    bool finite = std::isfinite(det) && det != 0.0;
    for (int i = 0; i < 3; ++i) {
        wc[i] = m[3 * i] * c[0] + m[3 * i + 1] * c[1] + m[3 * i + 2] * c[2];
        wh[i] = std::fabs(m[3 * i]) * h[0] + std::fabs(m[3 * i + 1]) * h[1] + std::fabs(m[3 * i + 2]) * h[2];
        finite = finite && std::isfinite(wc[i]) && std::isfinite(wh[i]);
        hmax = std::max(hmax, wh[i]); cmax = std::max(cmax, std::fabs(wc[i]));
    }
    const double pad = 1e-3 * hmax + 1e-4 * cmax + 1e-6;
    for (int i = 0; i < 3; ++i) out.push_back(finite ? (float)wc[i] : 0.f);
    out.push_back(0.f);
    for (int i = 0; i < 3; ++i) out.push_back(finite ? (float)(wh[i] + pad) : INFINITY);
    out.push_back(0.f);
}
void computeWorldBoxes(const HostScene& hs, std::vector<float>& out) {
    out.clear();
    for (const Top& t : hs.tops) {
        if (t.kind != T_BVH || t.node_count <= 0) continue;
        const BvhNode& root = hs.nodes[t.first_node];
        const float c[3] = {root.cx, root.cy, root.cz}, h[3] = {root.hx, root.hy, root.hz};
        worldBoxOf(hs.xforms64[t.xform].m, c, h, out);
    }
}
// The same for the top-level SDF primitives (SDFGeometry.aabb under the primitive's transform): sdf_kernel's hand-over
// rejects the rays that miss the padded box with one FP32 slab test before it sets up the march in the reference's f64
// arithmetic (most rays of an SDF scene never meet the SDF).  An infinite box (SDF_SphereRepetition) never rejects.
void computeSdfWorldBoxes(const HostScene& hs, std::vector<float>& out) {
    out.clear();
    for (const Top& t : hs.tops) {
        if (t.kind != T_SDF) continue;
        const Prim& p = hs.prims[t.first_prim];
        const SdfProgram& pr = hs.sdfs[p.geom_index];
        const float c[3] = {pr.cx, pr.cy, pr.cz}, h[3] = {pr.hx, pr.hy, pr.hz};
        worldBoxOf(hs.xforms64[p.xform].m, c, h, out);
    }
}

// Top-level BVH over the BVHAggregates of a scene with many of them (27 dragon instances: the reference's linear walk over
// world.objects costs 27 ray transforms + root-box tests per ray, 8.7 after the world-box reject).  Built here, by a
// plain surface-area-heuristic sweep over the instances' padded world boxes (computeWorldBoxes); the reference has no
// counterpart and needs none, because the structure only prunes: every aggregate whose box the ray reaches within the
// current bound is still entered, and exact ties are settled by the static rank (entry index, primitive index) whatever
// the order of the visits.  Nodes use the device's stackless format with leaves = (1 << 24) | aggregate ordinal.
namespace {
struct TlasItem { double lo[3], hi[3]; int ordinal; };
void tlasEmit(std::vector<TlasItem>& items, int begin, int end, std::vector<BvhNode>& out) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int i = begin; i < end; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], items[i].lo[a]); hi[a] = std::max(hi[a], items[i].hi[a]);
            const double c = 0.5 * (items[i].lo[a] + items[i].hi[a]);
            clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
        }
    const int me = (int)out.size();
    BvhNode b{};
    const double c[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    b.cx = (float)c[0]; b.cy = (float)c[1]; b.cz = (float)c[2];
    // half sizes rounded outwards so that the f32 box contains the f64 union
    auto up = [](double h, double cc, float cf) { const double need = h + std::fabs(cc - (double)cf); float f = (float)need; while ((double)f < need) f = std::nextafter(f, INFINITY); return f; };
    b.hx = up(0.5 * (hi[0] - lo[0]), c[0], b.cx); b.hy = up(0.5 * (hi[1] - lo[1]), c[1], b.cy); b.hz = up(0.5 * (hi[2] - lo[2]), c[2], b.cz);
    b.leaf = -1;
    out.push_back(b);
    if (end - begin == 1) { out[me].leaf = (1 << 24) | items[begin].ordinal; out[me].skip = (int)out.size(); return; }
    int ax = 0; for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > chi[ax] - clo[ax]) ax = a;
    std::sort(items.begin() + begin, items.begin() + end, [ax](const TlasItem& x, const TlasItem& y) { return x.lo[ax] + x.hi[ax] < y.lo[ax] + y.hi[ax]; });
    // SAH sweep along that axis
    const int n = end - begin;
    auto area = [](const double l[3], const double h[3]) { const double d[3] = {std::max(0.0, h[0] - l[0]), std::max(0.0, h[1] - l[1]), std::max(0.0, h[2] - l[2])}; return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]; };
    std::vector<double> right(n, 0.0);
    { double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
      for (int i = n - 1; i > 0; --i) { for (int a = 0; a < 3; ++a) { l[a] = std::min(l[a], items[begin + i].lo[a]); h[a] = std::max(h[a], items[begin + i].hi[a]); } right[i] = area(l, h); } }
    int best = n / 2; double best_cost = 1e300;
    { double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
      for (int i = 1; i < n; ++i) {
          for (int a = 0; a < 3; ++a) { l[a] = std::min(l[a], items[begin + i - 1].lo[a]); h[a] = std::max(h[a], items[begin + i - 1].hi[a]); }
          const double cost = area(l, h) * i + right[i] * (n - i);
          if (cost < best_cost) { best_cost = cost; best = i; }
      } }
    tlasEmit(items, begin, begin + best, out);
    tlasEmit(items, begin + best, end, out);
    out[me].skip = (int)out.size();
}
}  // namespace

int buildTlas(HostScene& hs) {
    std::vector<float> wb;
    computeWorldBoxes(hs, wb);
    const int n = (int)(wb.size() / 8);
    std::vector<TlasItem> items;
    for (int k = 0; k < n; ++k) {
        TlasItem it; it.ordinal = k;
        for (int a = 0; a < 3; ++a) {
            const double c = wb[8 * k + a], h = wb[8 * k + 4 + a];
            if (!std::isfinite(c) || !std::isfinite(h)) return -1;          // an unbounded aggregate: keep the linear walk
            it.lo[a] = c - h; it.hi[a] = c + h;
        }
        items.push_back(it);
    }
    if (items.empty() || items.size() >= (1u << 24)) return -1;
    std::vector<BvhNode> rel;
    tlasEmit(items, 0, n, rel);
    const int base = (int)hs.nodes.size();
    if ((size_t)base + rel.size() >= (size_t)kNodeEnd) return -1;
    for (size_t i = 0; i < rel.size(); ++i) {
        BvhNode b = rel[i];
        b.skip = (rel[i].skip >= (int)rel.size()) ? kNodeEnd : base + rel[i].skip;
        if (rel[i].leaf == -1) b.leaf = kNodeInner | (base + (int)i + 1);
        hs.nodes.push_back(b);
    }
    return base;
}

}  // namespace jsrt
