// Ray-scene intersection on the device: the flattened equivalents of
// World.getMinimumIntersection (src/world.js:7-15), Primitive.intersect
// (:116-124), Aggregate.intersect / BVHAggregate.intersect
// (src/aggregates.js:14-18,43-49) and BVHAggregateNode.intersect (:207-225).
#pragma once
#include "device_math.cuh"

namespace jsrt {

struct DeviceScene {
    const Top* tops;
    const Prim* prims;
    const Xform* xforms;
    const Xform64* xforms64;
    const BvhNode* nodes;
    const Tri* tris;
    const TriShade* tri_shade;
    const float* boxes;
    const Material* materials;
    const Light* lights;
    const SdfProgram* sdfs;
    const SdfInstr* sdf_code;
    int n_top, n_lights, light_samples, max_depth;
    float bg[3];
    int pad;
};

struct Hit { float t; int prim; int top; float t_lo; };   // t_lo: low-order part of an f64 hit distance (SDF hits)

// Work counters of one ray (only maintained by the COUNT instantiations, which back
// the roofline's algorithmic bytes: SURVEY.md §8d).
struct Work { unsigned nodes = 0, leaf_prims = 0, top_prims = 0; unsigned long long sdf_evals = 0; };

// geometry.intersect(localRay, minDistance, maxDistance) for one placed primitive.
// `best` is the caller's current closest distance (only used to skip work that the
// caller's acceptance test would reject anyway).
JSRT_DEV float prim_intersect(const DeviceScene& sc, const int4 pa, float3 o, float3 d, float minD, float maxD, float best, unsigned long long* sdf_evals = nullptr) {
    switch (pa.x) {   // geom_kind
        case G_TRIANGLE: return triangle_intersect(sc.tris, pa.y, o, d, minD, fminf(maxD, best));
        case G_PLANE: return plane_t(o, d);
        case G_SQUARE: {                                   // src/geometry.js:287-291
            const float t = plane_t(o, d);
            const float3 p = ray_point(o, d, t);
            return (-0.5f <= p.x && p.x <= 0.5f && -0.5f <= p.y && p.y <= 0.5f) ? t : -CUDART_INF_F;
        }
        case G_CIRCLE: {                                   // src/geometry.js:310-314
            const float t = plane_t(o, d);
            const float3 p = ray_point(o, d, t);
            const float w = __fadd_rn(1.f, __fmul_rn(0.f, t)) - 1.f;   // (origin.w + direction.w * t) - 1: NaN for infinite t
            return (p.x * p.x + p.y * p.y + p.z * p.z + w * w <= 1.f) ? t : -CUDART_INF_F;
        }
        case G_BOX: {                                      // AABB.intersect src/geometry.js:173-179
            float3 c = f3(0.f, 0.f, 0.f), h = f3(0.5f, 0.5f, 0.5f);
            if (pa.y >= 0) { const float* b = sc.boxes + 8 * pa.y; c = f3(b[0], b[1], b[2]); h = f3(b[4], b[5], b[6]); }
            float t0, t1;
            if (!aabb_intersects(c, h, o, d, minD, maxD, t0, t1)) return -CUDART_INF_F;
            return (t0 >= minD) ? t0 : t1;
        }
        case G_SPHERE: return sphere_intersect(o, d, minD);
        case G_CYLINDER: {                                 // src/geometry.js:473-478
            float md = minD;
            if (fabsf(o.z) > 1.f && d.z != 0.f) md = js_max(md, -(o.z - js_sign(o.z)) / d.z);
            const float t = sphere_intersect(f3(o.x, o.y, 0.f), f3(d.x, d.y, 0.f), md);
            return (fabsf(o.z + t * d.z) <= 1.f) ? t : -CUDART_INF_F;
        }
        default: return -CUDART_INF_F;
    }
}

// One placed primitive against a ray given in its parent's space.
// `t_lo` receives the low-order part of the distance for SDF hits (their f64 distance
// is carried as t + t_lo so that shading recomputes the reference's hit point exactly).
JSRT_DEV float placed_prim_intersect(const DeviceScene& sc, int prim_index, float3 o, float3 d, float minD, float maxD, float best, bool shadow_ray,
                                     unsigned long long* sdf_evals = nullptr, float* t_lo = nullptr) {
    const int4* pp = reinterpret_cast<const int4*>(sc.prims + prim_index);
    const int4 pa = __ldg(pp);          // geom_kind, geom_index, material, xform
    const int flags = __ldg(reinterpret_cast<const int*>(pp + 1));
    if (shadow_ray && !(flags & PF_CASTS_SHADOW)) return CUDART_INF_F;     // src/world.js:117-118
    if (pa.x == G_SDF) {
        // ray.getTransformed(inv_transform) with the f64 matrix (src/math.js:392-397), then SDFGeometry.intersect
        const double* m = sc.xforms64[pa.w].m;
        const float3 lo = xf64_apply(m, o, 1.0), ld = xf64_apply(m, d, 0.0);
        const double t = sdf_intersect(sc.sdfs[pa.y], sc.sdf_code, sc.xforms64, lo, ld, (double)minD, (double)maxD, sdf_evals);
        const float tf = (float)t;
        if (t_lo) *t_lo = isfinite(t) ? (float)(t - (double)tf) : 0.f;
        return tf;
    }
    if (!(flags & PF_IDENTITY_XFORM)) {                                    // ray.getTransformed(inv_transform), src/world.js:120
        const XformReg m = load_xform(sc.xforms, pa.w);
        const float3 lo = xf_point(m, o), ld = xf_dir(m, d);
        return prim_intersect(sc, pa, lo, ld, minD, maxD, best, sdf_evals);
    }
    return prim_intersect(sc, pa, o, d, minD, maxD, best, sdf_evals);
}

// World.cast: closest hit over the top-level list in order, strict `<` so the
// earliest object wins exact ties (src/world.js:9-13).  ANY_HIT: shadow-ray
// semantics of materials.js:250-252 — only "is there a hit with minD < t < maxD"
// matters, so the walk stops at the first accepted hit (result-identical).
template <bool ANY_HIT, bool COUNT = false>
JSRT_DEV Hit trace_ray(const DeviceScene& sc, float3 o, float3 d, float minD, float maxD, Work* work = nullptr) {
    unsigned long long* const se = COUNT ? &work->sdf_evals : nullptr;
    Hit best; best.t = CUDART_INF_F; best.prim = -1; best.top = -1; best.t_lo = 0.f;
    for (int ti = 0; ti < sc.n_top; ++ti) {
        const int4* tp = reinterpret_cast<const int4*>(sc.tops + ti);
        const int4 ta = __ldg(tp);          // kind, xform, first_prim, prim_count
        if (ta.x == T_PRIM) {
            if (COUNT) ++work->top_prims;
            float tl = 0.f;
            const float t = placed_prim_intersect(sc, ta.z, o, d, minD, maxD, best.t, ANY_HIT, se, &tl);
            if (t > minD && t < best.t && t < maxD) { best.t = t; best.prim = ta.z; best.top = ti; best.t_lo = tl; if (ANY_HIT) return best; }
            continue;
        }
        const XformReg m = load_xform(sc.xforms, ta.y);
        const float3 lo = xf_point(m, o), ld = xf_dir(m, d);     // ray.getTransformed(this.getInvTransform())
        if (ta.x == T_LIST) {
            for (int k = 0; k < ta.w; ++k) {
                if (COUNT) ++work->top_prims;
                float tl = 0.f;
                const float t = placed_prim_intersect(sc, ta.z + k, lo, ld, minD, maxD, best.t, ANY_HIT, se, &tl);
                if (t > minD && t < best.t && t < maxD) { best.t = t; best.prim = ta.z + k; best.top = ti; best.t_lo = tl; if (ANY_HIT) return best; }
            }
            continue;
        }
        // T_BVH: stackless walk of the tree laid out in the reference's visit order.
        // The aggregate starts from its own `ret` (distance = Infinity,
        // src/aggregates.js:45) and the caller keeps it only if it beats the running
        // best with strict `<`; pruning with min(local, running) best gives the same
        // answer because a later-equal hit never replaces an earlier one.
        const int4 tb = __ldg(tp + 1);      // first_node, node_count, pad, pad
        const float4* nodes = reinterpret_cast<const float4*>(sc.nodes + tb.x);
        float local_best = CUDART_INF_F, local_lo = 0.f; int local_prim = -1;
        int i = 0;
        while (i < tb.y) {
            const float4 n0 = __ldg(nodes + 2 * i), n1 = __ldg(nodes + 2 * i + 1);
            const int skip = __float_as_int(n1.z), leaf = __float_as_int(n1.w);
            float b0, b1;
            if (COUNT) ++work->nodes;
            // src/aggregates.js:208-209
            if (aabb_intersects(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), lo, ld, minD, maxD, b0, b1) && b0 <= maxD && b1 >= minD && b0 <= local_best) {
                if (leaf != -1) {
                    const int cnt = (int)((unsigned)leaf >> 24), first = ta.z + (leaf & 0xffffff);
                    for (int k = 0; k < cnt; ++k) {
                        if (COUNT) ++work->leaf_prims;
                        float tl = 0.f;
                        const float t = placed_prim_intersect(sc, first + k, lo, ld, minD, maxD, fminf(local_best, best.t), ANY_HIT, se, &tl);
                        if (t > minD && t < maxD && t < local_best) { local_best = t; local_prim = first + k; local_lo = tl; }   // :213
                    }
                    if (ANY_HIT && local_prim >= 0) break;
                    i = skip;
                } else ++i;
            } else i = skip;
        }
        if (local_best > minD && local_best < best.t && local_best < maxD) { best.t = local_best; best.prim = local_prim; best.top = ti; best.t_lo = local_lo; if (ANY_HIT) return best; }
    }
    return best;
}

}  // namespace jsrt
