// Ray-scene intersection on the device: the flattened equivalents of
// World.getMinimumIntersection (src/world.js:7-15), Primitive.intersect
// (:116-124), Aggregate.intersect / BVHAggregate.intersect
// (src/aggregates.js:14-18,43-49) and BVHAggregateNode.intersect (:207-225).
#pragma once
#include "device_math.cuh"
#include "rng.h"

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
    const int* bvh_tops;       // indices of the T_BVH entries of `tops`
    const int* sdf_tops;       // indices of the T_SDF entries of `tops`
    const Texture* textures;   // TextureMaterialColor table + RGBA8 texels
    const unsigned char* texels;
    // Analytic-primitive table (prims_wave): every Primitive of world.objects that is not inside a BVHAggregate and is
    // not a top-level SDF, grouped by geometry kind so that each group runs its own tight loop.  Two copies: all of them
    // (closest-hit rays) and the shadow casters only (shadow rays, src/world.js:117-118).
    const APrim* atab;
    int atab_end[2][AG_COUNT]; // [0: closest-hit table | 1: shadow table][group]: end offset of the group within `atab`
    // World-space boxes of the BVHAggregates (two float4 each: centre | half x, half y | half z), padded: a cheap
    // reject in front of the per-aggregate ray transform + root-box test when a scene holds several aggregates
    // (use_wbox).  Conservative — a ray that misses the padded world box misses the local root box — so results are
    // those of the reference's linear walk over world.objects (src/world.js:7-15).
    const float4* wboxes;
    const float4* sdf_wboxes;  // the same for the top-level SDF primitives (sdf_tops order): sdf_wave's cheap reject in front of the f64 set-up
    int n_sdf_tops, use_wbox;
    int n_staged;              // nodes[0, n_staged): the top levels of every tree, staged in shared memory by bvh_kernel
    int tlas_root;             // >= 0: root node of the top-level BVH over the BVHAggregates (scene_flatten.cpp: buildTlas)
    int n_top, n_lights, light_samples, max_depth;
    float bg[3];
    int n_bvh;
};

struct Hit { float t; int prim; int top; float t_lo; };   // t_lo: low-order part of an f64 hit distance (SDF hits)

// Work counters of one ray (only maintained by the COUNT instantiations, which back
// the roofline's algorithmic bytes: SURVEY.md §8d).
struct Work { unsigned nodes = 0, leaf_prims = 0, top_prims = 0; unsigned long long sdf_evals = 0; };

// geometry.intersect(localRay, minDistance, maxDistance) for one placed primitive.
// `best` is the caller's current closest distance (only used to skip work that the
// caller's acceptance test would reject anyway).
JSRT_DEV float prim_intersect(const DeviceScene& sc, const int4 pa, float3 o, float3 d, float minD, float maxD, float best, float* t_lo = nullptr) {
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
template <bool HAS_SDF>
JSRT_DEV float placed_prim_intersect(const DeviceScene& sc, int prim_index, float3 o, float3 d, float minD, float maxD, float best, bool shadow_ray,
                                     unsigned long long* sdf_evals = nullptr, float* t_lo = nullptr) {
    const int4* pp = reinterpret_cast<const int4*>(sc.prims + prim_index);
    const int4 pa = __ldg(pp);          // geom_kind, geom_index, material, xform
    const int flags = __ldg(reinterpret_cast<const int*>(pp + 1));
    if (shadow_ray && !(flags & PF_CASTS_SHADOW)) return CUDART_INF_F;     // src/world.js:117-118
    if (HAS_SDF && pa.x == G_SDF) {
        // ray.getTransformed(inv_transform) with the f64 matrix (src/math.js:392-397), then SDFGeometry.intersect
        const double* m = sc.xforms64[pa.w].m;
        const float3 lo = xf64_apply(m, o, 1.0), ld = xf64_apply(m, d, 0.0);
        return split_t(sdf_intersect(sc.sdfs[pa.y], sc.sdf_code, sc.xforms64, lo, ld, (double)minD, (double)maxD, sdf_evals), t_lo);
    }
    if (!(flags & PF_IDENTITY_XFORM)) {                                    // ray.getTransformed(inv_transform), src/world.js:120
        const XformReg m = load_xform(sc.xforms, pa.w);
        const float3 lo = xf_point(m, o), ld = xf_dir(m, d);
        return prim_intersect(sc, pa, lo, ld, minD, maxD, best, t_lo);
    }
    return prim_intersect(sc, pa, o, d, minD, maxD, best, t_lo);
}

// ---------------------------------------------------------------------------------
// World.cast for a whole queue of rays, in two kernels per wave:
//
//   prims_wave  every top-level Primitive / plain Aggregate of world.objects, one ray per
//               thread: all lanes run the same short loop, no divergence to speak of.
//   bvh_wave    the BVHAggregates, persistent threads, one ray per lane, lanes refilled
//               individually from a warp-local pool (one atomicAdd per 128 rays): a
//               lane whose walk ends takes the next ray instead of idling while its
//               neighbours finish (walk lengths differ by two orders of magnitude: most
//               rays miss the mesh's root box after one node).
//
// Reference semantics being reproduced (src/world.js:7-15, src/aggregates.js:43-49,207-225):
// the closest hit over world.objects with strict `<`, so the earliest object wins exact
// ties; inside a BVHAggregate the reference walks greater-child-first and also keeps the
// first of equal hits.  Both orders are ray-independent, so "first visited" is a static
// rank: (top-level index, leaf rank).  Placed primitives are stored in leaf-rank order,
// hence any evaluation order gives the reference's answer as long as ties go to the lower
// (top index, primitive index) — which is what lets the analytic primitives run before
// the BVHs and lets the walk pick, per ray-direction octant, one of eight stackless
// (hit/miss-link) layouts of the same tree that visit the nearer child first.
// tuning constants of bvh_wave (measured on bunny_path / dragon 1080p, profiles/r1_ncu_summary.md)
#ifndef JSRT_REFILL_T
#define JSRT_REFILL_T 12
#endif
#ifndef JSRT_LEAF_T
#define JSRT_LEAF_T 12
#endif
#ifndef JSRT_PROG_T
#define JSRT_PROG_T 20     // parked leaves are tested when 12 lanes hold one or fewer than 20 can still walk (profiles/r2/ab_r2j_*: grid over 16..33 x 8..16 x refill 8..16)
#endif
#ifndef JSRT_NODE_LDG256
#define JSRT_NODE_LDG256 0
#endif
#ifndef JSRT_POOL_BATCH
#define JSRT_POOL_BATCH 64
#endif
#ifndef JSRT_NODE_STEPS
#define JSRT_NODE_STEPS 8
#endif
// FP32 near-ties between two triangles of a mesh settled in the reference's f64 (tie_wave); 0 = off (A/B runs).
#ifndef JSRT_TRI_TIE
#define JSRT_TRI_TIE 1
#endif
enum TraceMode { TM_EXTEND = 0, TM_SHADOW = 1 };

struct TraceIO {
    const float4* __restrict__ o;       // origin.xyz | pixel
    const float4* __restrict__ d;       // direction.xyz | node id (extend)
    const float4* __restrict__ c;       // shadow: contribution
    float4* __restrict__ hits;          // t | prim | top | t_lo   (partial result between the two kernels)
    float4* __restrict__ accum;         // shadow: pixel sums
    const int* __restrict__ count;
    int cap;
    int* cursor;
    unsigned long long* stats;
    float4* __restrict__ aux;           // extend + SDF scenes: local SDF normal of the hit (xyz, w = 1) or w = 0
    int final_pass;                     // 1 if no further tracing kernel follows for this wave (the last one writes results)
    int2* __restrict__ list;            // BVH work list: (ray index, first BVH whose root box the ray hits), written by
    int* list_count;                    //   prims_wave with warp-aggregated appends, consumed by bvh_wave
    // JSRT_FLAG_AOV renders: radiance goes to a per-sample buffer (`accum` then points at it) so that the per-pixel
    // variance can be formed from whole samples; slot = pixel + (pass - pass0) * accum_stride.  0 otherwise.
    int accum_stride, pass0;
    int4* __restrict__ tie_list; int* tie_count; int tie_cap;      // extend: FP32 near-ties between triangles (tie_wave)
};

// One radiance term into the pixel's sum.  sm_90+ has a 128-bit vector reduction (red.global.add.v4.f32): one L2
// operation per term instead of three scalar ones (JSRT_ACCUM_VEC4=0 keeps the scalar form for A/B runs).
#ifndef JSRT_ACCUM_VEC4
#define JSRT_ACCUM_VEC4 1
#endif
JSRT_DEV void accum_add3(float4* accum, uint32_t pixel, float3 c) {
#if JSRT_ACCUM_VEC4
    if (c.x != 0.f || c.y != 0.f || c.z != 0.f)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(accum + pixel), "f"(c.x), "f"(c.y), "f"(c.z), "f"(0.f) : "memory");
#else
    float* a = reinterpret_cast<float*>(accum + pixel);
    if (c.x != 0.f) atomicAdd(a + 0, c.x);
    if (c.y != 0.f) atomicAdd(a + 1, c.y);
    if (c.z != 0.f) atomicAdd(a + 2, c.z);
#endif
}

JSRT_DEV bool better_hit(float t, int top, const Hit& best) { return t < best.t || (t == best.t && top < best.top); }

// AABB.get_intersects (src/geometry.js:189-209) for a node stored as (centre, half) with the ray's reciprocal
// direction: (p_i +- h_i) * (1 / d_i); axes with |d_i| <= 1e-7 follow the reference's parallel rule (:194,205).
// The per-axis early returns of the reference are equivalent to one final test because t_min only grows and
// t_max only shrinks.  Returns the entry distance in b0.
JSRT_DEV bool slab_test(const float4 n0, const float4 n1, float3 lo, float3 inv, bool parx, bool pary, bool parz, float minD, float maxD, float& b0) {
    const float px = n0.x - lo.x, py = n0.y - lo.y, pz = n0.z - lo.z;
    const float ax = (px + n0.w) * inv.x, bx = (px - n0.w) * inv.x;
    const float ay = (py + n1.x) * inv.y, by = (py - n1.x) * inv.y;
    const float az = (pz + n1.y) * inv.z, bz = (pz - n1.y) * inv.z;
    float b1 = CUDART_INF_F; b0 = -CUDART_INF_F;
    bool miss = false;
    if (parx) miss = fabsf(px) > n0.w; else { b0 = fminf(ax, bx); b1 = fmaxf(ax, bx); }
    if (pary) miss = miss || fabsf(py) > n1.x; else { b0 = fmaxf(b0, fminf(ay, by)); b1 = fminf(b1, fmaxf(ay, by)); }
    if (parz) miss = miss || fabsf(pz) > n1.y; else { b0 = fmaxf(b0, fminf(az, bz)); b1 = fminf(b1, fmaxf(az, bz)); }
    return !miss && !(b0 > b1) && !(b1 < minD) && !(b0 > maxD);
}

// four components at once (first-hit AOVs: normal.xyz, distance)
JSRT_DEV void accum_add3w(float4* buf, uint32_t pixel, float3 c, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(buf + pixel), "f"(c.x), "f"(c.y), "f"(c.z), "f"(w) : "memory");
}

template <int MODE>
JSRT_DEV void ray_window(const float4 d4, float& minD, float& maxD, bool& primary) {
    if (MODE == TM_EXTEND) {
        primary = __float_as_int(d4.w) == 1;
        minD = primary ? 0.f : 0.0001f; maxD = CUDART_INF_F;       // src/world.js:31-34, src/materials.js:279,286,319,328
    } else { primary = false; minD = 0.0001f; maxD = 1.0f; }        // src/materials.js:250
}

template <int MODE>
JSRT_DEV void finish_ray(const TraceIO& io, int i, const Hit& best, const float4 o4) {
    if (MODE == TM_EXTEND) io.hits[i] = make_float4(best.t, __int_as_float(best.prim), __int_as_float(best.top), best.t_lo);
    else if (best.prim < 0) {            // unoccluded: the light sample counts (src/materials.js:251-253)
        const float4 c4 = io.c[i];
        uint32_t slot = (uint32_t)__float_as_int(o4.w);
        if (io.accum_stride) slot += (uint32_t)(__float_as_int(io.d[i].w) - io.pass0) * (uint32_t)io.accum_stride;   // shadow rays carry their pass in d.w
        accum_add3(io.accum, slot, f3(c4.x, c4.y, c4.z));
    }
}

// Per-ray constants of a walk through one BVHAggregate: the ray in the aggregate's space
// (Aggregate.intersect, src/aggregates.js:15) and what the slab test needs.
// The slab tests multiply by 1 / d; with JSRT_FAST_RCP that reciprocal comes from the SFU (1 ulp) instead of
// the IEEE sequence: three divisions per ray and BVH in every refill and in shade_kernel's fused root-box tests.
#ifndef JSRT_FAST_RCP
#define JSRT_FAST_RCP 1
#endif
#if JSRT_FAST_RCP
#define JSRT_RCP(x) rcp_fast(x)
#else
#define JSRT_RCP(x) (1.0f / (x))
#endif
struct LocalRay {
    float3 lo, ld, inv, sgn;      // origin, direction, 1 / direction, sign of the direction (+-1)
    bool par;                     // some |d_i| <= 1e-7: AABB.get_intersects' parallel rule applies (src/geometry.js:194,205)
};
JSRT_DEV LocalRay make_local_ray(const XformReg& m, float3 o, float3 d) {
    LocalRay r;
    r.lo = xf_point(m, o); r.ld = xf_dir(m, d);              // ray.getTransformed(this.getInvTransform())
    r.par = !(fabsf(r.ld.x) > 0.0000001f) || !(fabsf(r.ld.y) > 0.0000001f) || !(fabsf(r.ld.z) > 0.0000001f);
    r.inv = f3(JSRT_RCP(r.ld.x), JSRT_RCP(r.ld.y), JSRT_RCP(r.ld.z));
    r.sgn = f3(r.ld.x < 0.f ? -1.f : 1.f, r.ld.y < 0.f ? -1.f : 1.f, r.ld.z < 0.f ? -1.f : 1.f);
    return r;
}
// the same for a ray that is already in the right space (the top-level BVH is walked with the world ray)
JSRT_DEV LocalRay make_ray(float3 o, float3 d) {
    LocalRay r;
    r.lo = o; r.ld = d;
    r.par = !(fabsf(d.x) > 0.0000001f) || !(fabsf(d.y) > 0.0000001f) || !(fabsf(d.z) > 0.0000001f);
    r.inv = f3(JSRT_RCP(d.x), JSRT_RCP(d.y), JSRT_RCP(d.z));
    r.sgn = f3(d.x < 0.f ? -1.f : 1.f, d.y < 0.f ? -1.f : 1.f, d.z < 0.f ? -1.f : 1.f);
    return r;
}
// AABB.get_intersects (src/geometry.js:189-209) + the visit condition of BVHAggregateNode.intersect (:209) for
// a ray without parallel axes.  The reference sorts (p+h)/d and (p-h)/d per axis; which of the two is the
// smaller is known from the sign of d, so the entry distance of an axis is (p - sgn*h) * (1/d) and the exit
// distance (p + sgn*h) * (1/d): the same two products, no min/max.  `hi` = min(maxDistance, closest hit so far).
JSRT_DEV bool slab_fast(const float4 n0, const float4 n1, const LocalRay& r, float minD, float hi) {
    const float px = n0.x - r.lo.x, py = n0.y - r.lo.y, pz = n0.z - r.lo.z;
    const float nx = fmaf(-r.sgn.x, n0.w, px) * r.inv.x, fx = fmaf(r.sgn.x, n0.w, px) * r.inv.x;
    const float ny = fmaf(-r.sgn.y, n1.x, py) * r.inv.y, fy = fmaf(r.sgn.y, n1.x, py) * r.inv.y;
    const float nz = fmaf(-r.sgn.z, n1.y, pz) * r.inv.z, fz = fmaf(r.sgn.z, n1.y, pz) * r.inv.z;
    const float b0 = fmaxf(fmaxf(nx, ny), nz), b1 = fminf(fminf(fx, fy), fz);
    return b0 <= b1 && b1 >= minD && b0 <= hi;
}
// the general form (parallel axes): as the reference writes it
JSRT_DEV bool slab_general(const float4 n0, const float4 n1, const LocalRay& r, float minD, float maxD, float bound) {
    const bool parx = !(fabsf(r.ld.x) > 0.0000001f), pary = !(fabsf(r.ld.y) > 0.0000001f), parz = !(fabsf(r.ld.z) > 0.0000001f);
    float b0;
    return slab_test(n0, n1, r.lo, r.inv, parx, pary, parz, minD, maxD, b0) && b0 <= bound;
}
JSRT_DEV bool slab_any(const float4 n0, const float4 n1, const LocalRay& r, float minD, float maxD, float bound) {
    return r.par ? slab_general(n0, n1, r, minD, maxD, bound) : slab_fast(n0, n1, r, minD, fminf(maxD, bound));
}

// AABB.intersect (src/geometry.js:173-179) over AABB.get_intersects (:189-209) for a box primitive.  Rays without a
// parallel axis take the sign-ordered form of slab_fast (the per-axis sort of (p+h)/d, (p-h)/d is known from the sign
// of d; products with 1/d instead of six divisions); the others the reference's general form.
JSRT_DEV float box_prim_intersect(float3 c, float3 h, float3 o, float3 d, float minD, float maxD) {
    float t0, t1;
    if (!(fabsf(d.x) > 0.0000001f) || !(fabsf(d.y) > 0.0000001f) || !(fabsf(d.z) > 0.0000001f)) {
        if (!aabb_intersects(c, h, o, d, minD, maxD, t0, t1)) return -CUDART_INF_F;
    } else {
        const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
        const float px = c.x - o.x, py = c.y - o.y, pz = c.z - o.z;
        const float sx = copysignf(h.x, d.x), sy = copysignf(h.y, d.y), sz = copysignf(h.z, d.z);
        t0 = fmaxf(fmaxf((px - sx) * ix, (py - sy) * iy), (pz - sz) * iz);
        t1 = fminf(fminf((px + sx) * ix, (py + sy) * iy), (pz + sz) * iz);
        if (!(t0 <= t1) || t1 < minD || t0 > maxD) return -CUDART_INF_F;
    }
    return (t0 >= minD) ? t0 : t1;
}

// World-space reject for BVHAggregate number b (DeviceScene::wboxes): slab test with the world ray's reciprocal
// direction against the same visit window as BVHAggregateNode.intersect (src/aggregates.js:209).  NaNs (0 * inf:
// origin exactly on a padded slab plane of an axis the ray is parallel to) drop out of fminf / fmaxf, which then
// reports a miss — correct, since the true box lies strictly inside the padded one.
JSRT_DEV bool wbox_hit(const float4* __restrict__ wb, int b, float3 o, float3 inv, float minD, float hi) {
    const float4 a = __ldg(wb + 2 * b), c = __ldg(wb + 2 * b + 1);
    const float px = a.x - o.x, py = a.y - o.y, pz = a.z - o.z;
    const float x0 = (px - a.w) * inv.x, x1 = (px + a.w) * inv.x;
    const float y0 = (py - c.x) * inv.y, y1 = (py + c.x) * inv.y;
    const float z0 = (pz - c.y) * inv.z, z1 = (pz + c.y) * inv.z;
    const float n = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
    const float f = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
    return n <= f && f >= minD && n <= hi;
}

// World.getMinimumIntersection (src/world.js:7-15) over the analytic primitives — every Primitive of world.objects
// outside the BVHAggregates — one tight loop per geometry kind (a single loop over world.objects with a switch cost 285
// instructions per test on cornell_box_path, the compiler having hoisted the sphere's f64 set-up in front of the
// switch: profiles/r1_s3).  The reference's "first object wins exact ties" is the static rank (top, prim), so the order
// of the tests is free.  Called together by the lanes of `mask` (a converged subset of the warp).
// Shadow rays (ANY_HIT): no per-lane early exit inside the loops.  A lane that leaves a loop on its own runs the rest of
// the body apart from its warp (measured on cornell_box_path: 8.7 of 32 lanes active in the box loop, each group of
// lanes running it separately); instead every lane runs every test of a group and the lanes of `mask` skip the
// remaining groups together once all of their rays are occluded.
// PLANES_ONLY (shade_kernel's lean build): the table holds planes only, the loops of the other kinds are compiled out.
template <bool ANY_HIT, bool COUNT, bool HAS_SDF, bool PLANES_ONLY = false>
JSRT_DEV void analytic_hits(const DeviceScene& sc, const float3 o, const float3 d, const float minD, const float maxD, const unsigned mask, Hit& best, Work* work) {
    const APrim* const tab = sc.atab;
    const int tb = ANY_HIT ? 1 : 0;
    int k = (ANY_HIT ? sc.atab_end[0][AG_COUNT - 1] : 0);
    #define JSRT_ACCEPT(T, TL)                                                                                              \
        if ((T) > minD && (T) < maxD && ((T) < best.t || ((T) == best.t && (e0.y < best.top || (e0.y == best.top && e0.x < best.prim))))) { \
            best.t = (T); best.prim = e0.x; best.top = e0.y; best.t_lo = (TL); }
    #define JSRT_ENTRY()                                                                                                    \
        const int4 e0 = __ldg(reinterpret_cast<const int4*>(tab + k));     /* prim, top, agg_xform, xform */                \
        const int4 e1 = __ldg(reinterpret_cast<const int4*>(tab + k) + 1); /* geom_index, flags */                          \
        if (COUNT) ++work->top_prims;
    // ray.getTransformed(inv_transform) (src/world.js:120), after the enclosing Aggregate's own map if there is one
    #define JSRT_LOCAL_RAY()                                                                                                \
        float3 lo = o, ld = d;                                                                                              \
        if (e0.z >= 0) { const XformReg ma = load_xform(sc.xforms, e0.z); lo = xf_point(ma, o); ld = xf_dir(ma, d); }         \
        if (!(e1.y & PF_IDENTITY_XFORM)) { const XformReg m = load_xform(sc.xforms, e0.w); const float3 a = xf_point(m, lo), b = xf_dir(m, ld); lo = a; ld = b; }
    #define JSRT_GROUP_DONE() (ANY_HIT && __all_sync(mask, best.prim >= 0))
    for (; k < sc.atab_end[tb][AG_PLANE]; ++k) {                           // SimplePlane.intersect src/geometry.js:246-248
        JSRT_ENTRY()
        float oz, dz;
        if (e0.z < 0 && !(e1.y & PF_IDENTITY_XFORM)) {                     // only the z row of the local ray is needed
            const float4 r2 = __ldg(reinterpret_cast<const float4*>(sc.xforms + e0.w) + 2);
            oz = r2.x * o.x + r2.y * o.y + r2.z * o.z + r2.w; dz = r2.x * d.x + r2.y * d.y + r2.z * d.z;
        } else { JSRT_LOCAL_RAY() oz = lo.z; dz = ld.z; }
        // (shadow rays only compare t with their window: SFU division)
        const float t = (dz != 0.f) ? (ANY_HIT ? __fdividef(-oz, dz) : -oz / dz) : -CUDART_INF_F;
        JSRT_ACCEPT(t, 0.f)
    }
    if (!PLANES_ONLY && !JSRT_GROUP_DONE()) for (; k < sc.atab_end[tb][AG_SQUARE]; ++k) {   // Square.intersect src/geometry.js:287-291
        JSRT_ENTRY()
        JSRT_LOCAL_RAY()
        const float t = plane_t(lo, ld);
        const float x = __fadd_rn(lo.x, __fmul_rn(ld.x, t)), y = __fadd_rn(lo.y, __fmul_rn(ld.y, t));
        if (-0.5f <= x && x <= 0.5f && -0.5f <= y && y <= 0.5f) { JSRT_ACCEPT(t, 0.f) }
    }
    if (!PLANES_ONLY && !JSRT_GROUP_DONE()) for (k = sc.atab_end[tb][AG_SQUARE]; k < sc.atab_end[tb][AG_BOX]; ++k) {      // AABB.intersect src/geometry.js:173-179
        JSRT_ENTRY()
        JSRT_LOCAL_RAY()
        float3 c = f3(0.f, 0.f, 0.f), h = f3(0.5f, 0.5f, 0.5f);
        if (e1.x >= 0) { const float* b = sc.boxes + 8 * e1.x; c = f3(__ldg(b), __ldg(b + 1), __ldg(b + 2)); h = f3(__ldg(b + 4), __ldg(b + 5), __ldg(b + 6)); }
        const float t = box_prim_intersect(c, h, lo, ld, minD, maxD);
        JSRT_ACCEPT(t, 0.f)
    }
    if (!PLANES_ONLY && !JSRT_GROUP_DONE()) for (k = sc.atab_end[tb][AG_BOX]; k < sc.atab_end[tb][AG_SPHERE]; ++k) {      // Sphere.staticIntersect src/geometry.js:429-442
        JSRT_ENTRY()
        JSRT_LOCAL_RAY()
        const float t = sphere_intersect(lo, ld, minD);
        JSRT_ACCEPT(t, 0.f)
    }
    if (!PLANES_ONLY && !JSRT_GROUP_DONE()) for (k = sc.atab_end[tb][AG_SPHERE]; k < sc.atab_end[tb][AG_OTHER]; ++k) {    // every other geometry: the general code
        if (ANY_HIT && best.prim >= 0) break;
        JSRT_ENTRY()
        float3 lo = o, ld = d;
        if (e0.z >= 0) { const XformReg ma = load_xform(sc.xforms, e0.z); lo = xf_point(ma, o); ld = xf_dir(ma, d); }
        float tl = 0.f;
        const float t = placed_prim_intersect<HAS_SDF>(sc, e0.x, lo, ld, minD, maxD, best.t, ANY_HIT, COUNT ? &work->sdf_evals : nullptr, &tl);
        JSRT_ACCEPT(t, tl)
    }
    #undef JSRT_GROUP_DONE
    #undef JSRT_ACCEPT
    #undef JSRT_ENTRY
    #undef JSRT_LOCAL_RAY
}

// The root box of every BVHAggregate in world.objects order (the first test of BVHAggregateNode.intersect,
// src/aggregates.js:208-209), behind the padded world-space reject when the scene holds several aggregates.
// Returns the ordinal of the first aggregate whose root box the ray hits within (minD, min(maxD, best_t)], or -1;
// `lr` receives the ray in that aggregate's space.
template <bool COUNT>
JSRT_DEV int first_bvh_hit(const DeviceScene& sc, const float3 o, const float3 d, const float minD, const float maxD, const float best_t, Work* work, LocalRay& lr) {
    if (sc.tlas_root >= 0) {           // many aggregates: the walk starts at the top-level BVH; here only its root box
        const LocalRay r = make_ray(o, d);
        const float4* root = reinterpret_cast<const float4*>(sc.nodes + sc.tlas_root);
        if (COUNT) ++work->nodes;
        if (slab_any(__ldg(root), __ldg(root + 1), r, minD, maxD, best_t)) { lr = r; return 0; }
        return -1;
    }
    float3 winv = f3(0.f, 0.f, 0.f);
    if (sc.use_wbox) winv = f3(JSRT_RCP(d.x), JSRT_RCP(d.y), JSRT_RCP(d.z));      // (the world boxes are padded by 1e-3)
    for (int b = 0; b < sc.n_bvh; ++b) {
        if (sc.use_wbox && !wbox_hit(sc.wboxes, b, o, winv, minD, fminf(maxD, best_t))) continue;
        const int4* tp = reinterpret_cast<const int4*>(sc.tops + __ldg(sc.bvh_tops + b));
        const int4 ta = __ldg(tp); const int first_node = __ldg(reinterpret_cast<const int*>(tp + 1));
        const LocalRay r = make_local_ray(load_xform(sc.xforms, ta.y), o, d);
        const float4* root = reinterpret_cast<const float4*>(sc.nodes + first_node);
        if (COUNT) ++work->nodes;
        if (slab_any(__ldg(root), __ldg(root + 1), r, minD, maxD, best_t)) { lr = r; return b; }
    }
    return -1;
}

// Camera rays (generate): what prims_kernel<extend, GEN> computes itself instead of reading a queue entry.
struct GenParams {
    Camera cam;
    int width, height, x_offset, x_delt, ncols, npix_active;
    int first_pass, jitter, max_depth, use_lens, count_samples;
    unsigned long long seed;
    long long first_sample;      // index of the batch's first sample within the call
    int n_samples;               // samples in this batch
    float4 *qo, *qd, *qw;        // the level-0 ray queue (written), the pixel sums (sample counts)
    float4* accum;
};

// camera.getRayForPixel for sample s of the batch (src/cameras.js:29-34,46-52) with the pixel / jitter arithmetic of
// src/renderers.js:89-96, all in f64 so the primary ray is the reference's bit for bit.
JSRT_DEV void camera_ray(const GenParams& g, int px, int py, uint32_t sample_key, bool jitter, bool use_lens, float3& o, float3& d) {
    const uint32_t nk = rng_node_key(sample_key, 1);
    double x = dsub(dmul(2.0, (double)px / (double)g.width), 1.0);
    double y = dadd(dmul(-2.0, (double)py / (double)g.height), 1.0);
    if (jitter) {
        x = dadd(x, dmul(2.0 / (double)g.width, dsub((double)rng_u01(nk, DIM_JITTER_X), 0.5)));
        y = dadd(y, dmul(2.0 / (double)g.height, dsub((double)rng_u01(nk, DIM_JITTER_Y), 0.5)));
    }
    const Camera& c = g.cam;
    const float dx = (float)dmul(dmul(x, c.tan_fov), c.aspect), dy = (float)dmul(y, c.tan_fov), dz = -1.f;   // Vec.of(...) stores f32
    const double* t = c.t;
    // transform.times(direction): f64 dot of the f32 vector with each row, stored f32 (w = 0)
    float3 dir = f3((float)ddot4(dx, dy, dz, 0.0, t[0], t[1], t[2], t[3]), (float)ddot4(dx, dy, dz, 0.0, t[4], t[5], t[6], t[7]),
                    (float)ddot4(dx, dy, dz, 0.0, t[8], t[9], t[10], t[11]));
    float3 org = f3((float)t[3], (float)t[7], (float)t[11]);       // transform.column(3)
    if (c.dof && use_lens) {
        // Vec.circlePick src/math.js:175-179, then DepthOfFieldPerspectiveCamera.getRayForPixel (src/cameras.js:46-52)
        const double a = dmul(dmul((double)rng_u01(nk, DIM_LENS_A), 2.0), 3.141592653589793), r = sqrt((double)rng_u01(nk, DIM_LENS_R));
        const float cx = (float)dmul(r, cos(a)), cy = (float)dmul(r, sin(a));
        const float sx = (float)dmul(cx, c.sensor_size), sy = (float)dmul(cy, c.sensor_size);
        const float3 off = f3((float)ddot4(sx, sy, 0.0, 0.0, t[0], t[1], t[2], t[3]), (float)ddot4(sx, sy, 0.0, 0.0, t[4], t[5], t[6], t[7]),
                              (float)ddot4(sx, sy, 0.0, 0.0, t[8], t[9], t[10], t[11]));
        org = f3((float)dadd(org.x, off.x), (float)dadd(org.y, off.y), (float)dadd(org.z, off.z));
        const float fx = (float)dmul(dir.x, c.focus_distance), fy = (float)dmul(dir.y, c.focus_distance), fz = (float)dmul(dir.z, c.focus_distance);
        const float mx = (float)dsub(fx, off.x), my = (float)dsub(fy, off.y), mz = (float)dsub(fz, off.z);
        const double nn = sqrt(ddot4(mx, my, mz, 0.0, mx, my, mz, 0.0));
        if (nn > 0.00001) { const double inv = 1.0 / nn; dir = f3((float)dmul(mx, inv), (float)dmul(my, inv), (float)dmul(mz, inv)); }
        else dir = f3(mx, my, mz);
    }
    o = org; d = dir;
}
// sample s of the batch -> (pixel, pass) and the three queue words of its camera ray
JSRT_DEV void generate_sample(const GenParams& g, int s, float4& o4, float4& d4, float4& w4, uint32_t& pixel) {
    const long long gs = g.first_sample + s;
    const int pass = g.first_pass + (int)(gs / g.npix_active);
    const int idx = (int)(gs % g.npix_active);
    const int py = idx / g.ncols, px = g.x_offset + (idx % g.ncols) * g.x_delt;
    pixel = (uint32_t)(py * g.width + px);
    const uint32_t key = rng_sample_key(g.seed, pixel, (uint32_t)pass);
    float3 o, d;
    camera_ray(g, px, py, key, g.jitter != 0, g.use_lens != 0, o, d);
    o4 = make_float4(o.x, o.y, o.z, __int_as_float((int)pixel));
    d4 = make_float4(d.x, d.y, d.z, __int_as_float(1));
    w4 = make_float4(1.f, 1.f, 1.f, __int_as_float((pass << 8) | g.max_depth));
}

// prims_wave: every top-level Primitive / plain Aggregate against every ray of the queue, one ray per thread,
// then the root box of every BVHAggregate (the first test of BVHAggregateNode.intersect, src/aggregates.js:208-209).
// Rays that hit no root box are finished here (most rays: 92 % of bunny_path's camera rays miss the mesh's
// box); the others are appended to the BVH work list, compacted with a warp ballot + prefix popcount and one
// atomicAdd per warp, so that bvh_wave only ever sees rays that walk.
// GEN (level 0 of a batch): the ray of entry i is the camera ray of sample i, computed here and written to the queue for
// the later kernels, instead of being written by a generate kernel and read back (92 % of them only ever meet the plane).
template <int MODE, bool COUNT, bool HAS_SDF, bool GEN>
JSRT_DEV void prims_wave(const DeviceScene& sc, const TraceIO& io, const GenParams* gen, Work* work_primary, Work* work_other) {
    constexpr bool ANY_HIT = (MODE == TM_SHADOW);
    const unsigned FULL = 0xffffffffu;
    const int n = GEN ? gen->n_samples : min(*io.count, io.cap);
    const int stride = gridDim.x * blockDim.x;
    const int n_round = (n + 31) & ~31;                // warp-uniform trip count: every lane reaches the ballot
    const int lane = threadIdx.x & 31;
    // The queue is streamed once (32 B per ray in, 16-24 B out) and the arithmetic per ray is short, so the kernel lives
    // on bytes in flight: the next ray of this thread is fetched before the current one is processed
    // (JSRT_PRIMS_PREFETCH=0: plain loads; measured in profiles/r1_s3).
#ifndef JSRT_PRIMS_PREFETCH
#define JSRT_PRIMS_PREFETCH 1      // 2: the same with evict-first (ld.global.cs) loads
#endif
#define JSRT_PRIMS_LD(p) ((JSRT_PRIMS_PREFETCH == 2) ? __ldcs(p) : *(p))
    constexpr bool PREFETCH = JSRT_PRIMS_PREFETCH && !GEN;
    float4 no4 = make_float4(0, 0, 0, 0), nd4 = no4;
    {
        const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (PREFETCH && i0 < n) { no4 = JSRT_PRIMS_LD(io.o + i0); nd4 = JSRT_PRIMS_LD(io.d + i0); }
    }
    bool p_any = false; int p_base = 0, p_rank = -1; int2 p_entry = make_int2(0, 0);      // the previous iteration's pending append
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int first_bvh = -1;
        Hit best; best.t = CUDART_INF_F; best.prim = -1; best.top = -1; best.t_lo = 0.f;
        float4 o4 = make_float4(0, 0, 0, 0);
        const unsigned live = __ballot_sync(FULL, i < n);
        float4 cur_d4 = nd4;
        if (PREFETCH) {
            o4 = no4;
            const int inext = i + stride;
            if (inext < n) { no4 = JSRT_PRIMS_LD(io.o + inext); nd4 = JSRT_PRIMS_LD(io.d + inext); }
        }
        if (i < n) {
            if (GEN) {
                float4 w4; uint32_t pixel;
                generate_sample(*gen, i, o4, cur_d4, w4, pixel);
                gen->qo[i] = o4; gen->qd[i] = cur_d4; gen->qw[i] = w4;
                if (gen->count_samples) atomicAdd(&gen->accum[pixel].w, 1.0f);      // samples taken for this pixel
            } else if (!PREFETCH) { o4 = io.o[i]; cur_d4 = io.d[i]; }
            const float4 d4 = cur_d4;
            const float3 o = f3(o4.x, o4.y, o4.z), d = f3(d4.x, d4.y, d4.z);
            float minD, maxD; bool primary;
            ray_window<MODE>(d4, minD, maxD, primary);
            Work* work = (COUNT && primary) ? work_primary : work_other;
            analytic_hits<ANY_HIT, COUNT, HAS_SDF>(sc, o, d, minD, maxD, live, best, work);
            if (!(ANY_HIT && best.prim >= 0)) { LocalRay lr; first_bvh = first_bvh_hit<COUNT>(sc, o, d, minD, maxD, best.t, work, lr); }
        }
        // Work-list append, software-pipelined: the atomicAdd that reserves this iteration's slots is issued here, its result
        // is only consumed — and the entries stored — at the same point of the NEXT iteration, so the counter's round trip
        // (~1 us under contention: it was 52 % of this kernel's stall samples, profiles/r2_ab.md §6) hides behind a whole
        // iteration of work instead of stalling the warp.
        const unsigned walkers = __ballot_sync(FULL, first_bvh >= 0);
        if (p_any) {
            const int b = __shfl_sync(FULL, p_base, 0);
            if (p_rank >= 0) io.list[b + p_rank] = p_entry;
        }
        p_any = walkers != 0u;
        if (p_any) {
            if (lane == 0) p_base = atomicAdd(io.list_count, __popc(walkers));
            p_rank = (first_bvh >= 0) ? __popc(walkers & ((1u << lane) - 1u)) : -1;
            p_entry = make_int2(i, first_bvh);
        }
        if (i < n) {
            if (io.final_pass && first_bvh < 0) finish_ray<MODE>(io, i, best, o4);
            else io.hits[i] = make_float4(best.t, __int_as_float(best.prim), __int_as_float(best.top), best.t_lo);
        }
    }
    if (p_any) {                                          // the last iteration's append
        const int b = __shfl_sync(FULL, p_base, 0);
        if (p_rank >= 0) io.list[b + p_rank] = p_entry;
    }
}

// One 32-byte node: two 128-bit loads, or (JSRT_NODE_LDG256) one 256-bit load (sm_100: LDG.E.256).
JSRT_DEV void load_node(const float4* p, float4& a, float4& b) {
#if JSRT_NODE_LDG256
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = __ldg(p); b = __ldg(p + 1);
#endif
}

// the same from the staged block in shared memory (32-bit shared-window address)
#ifndef JSRT_LDS_ASM
#define JSRT_LDS_ASM 1
#endif
#ifndef JSRT_STAGE_SOA
#define JSRT_STAGE_SOA 1          // staged block stored as two arrays of halves (render.cu: bvh_kernel)
#endif
JSRT_DEV void lds_node(unsigned addr, float4& a, float4& b) {
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(addr));
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(addr));
}
JSRT_DEV void lds_node_soa(unsigned addr_lo, unsigned addr_hi, float4& a, float4& b) {
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(addr_lo));
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(addr_hi));
}

// Two triangles of one mesh whose FP32 hit distances agree to the last bits (a ray through coincident or overlapping
// coplanar faces, or through a shared edge: both pass the inside test).  The reference decides `t < ret.distance`
// (src/aggregates.js:213) on f64 distances computed from f32 vectors, so the winner is settled by rounding noise ~1e-8
// that FP32 cannot see (x-wing at 256 x 256: 18 of 65 536 camera rays picked the neighbour).  bvh_wave keeps its FP32
// winner and appends (ray, BVH, winner, loser) to a tie list; tie_wave — a few dozen entries per frame — re-derives the
// local ray exactly as the reference does (f64 matrix times f32 vector, rounded to f32: src/math.js:294-296,392-397),
// forms both distances as Triangle.intersect forms them (src/geometry.js:345,368-370, un-contracted) and corrects the
// hit record.  Settling it inside the walk instead cost 6 % of bvh_kernel<extend> (one more live value in a 64-register
// kernel: profiles/r1_s4).
JSRT_DEV double triangle_t64(const Tri* __restrict__ tris, int idx, float3 o, float3 d) {
    const float4* tp = reinterpret_cast<const float4*>(tris + idx);
    const float4 a = __ldg(tp), b = __ldg(tp + 1);                      // normal | . , p0 | .
    const double den = ddot4(a.x, a.y, a.z, 0.0, d.x, d.y, d.z, 0.0);
    const double delta = ddot4(a.x, a.y, a.z, 0.0, b.x, b.y, b.z, 1.0);   // this.normal.dot(ps[0])
    return (den != 0.0) ? dsub(delta, ddot4(a.x, a.y, a.z, 0.0, o.x, o.y, o.z, 1.0)) / den : -CUDART_INF;
}
// Every reported candidate challenges the ray's current hit under the total order (f64 distance, rank) with a 64-bit
// compare-and-swap on (t, prim), so the outcome is the minimum over all candidates whatever the order of the entries
// (three-way ties on a shared vertex or on stacked duplicate faces included).
JSRT_DEV void tie_wave(const DeviceScene& sc, const TraceIO& io) {
    const int n = min(*io.tie_count, io.tie_cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 e = io.tie_list[i];                                  // ray, BVH, FP32 winner, FP32 loser (placed primitives)
        const int top_i = __ldg(sc.bvh_tops + e.y);
        const int4* tp = reinterpret_cast<const int4*>(sc.tops + top_i);
        const int4 ta = __ldg(tp), tb = __ldg(tp + 1);                  // kind, xform, first_prim, prim_count | first_node, node_count, tri_base, n_layouts
        if (__float_as_int(io.hits[e.x].z) != top_i) continue;          // the closest hit lies in another object by now
        const float4 o4 = io.o[e.x], d4 = io.d[e.x];
        const double* m = sc.xforms64[ta.y].m;
        const float3 lo = xf64_apply(m, f3(o4.x, o4.y, o4.z), 1.0), ld = xf64_apply(m, f3(d4.x, d4.y, d4.z), 0.0);
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(io.hits + e.x);      // low word: t, high word: prim
        for (int c = 0; c < 2; ++c) {
            const int cand = c ? e.w : e.z;
            const double t_c = triangle_t64(sc.tris, tb.z + (cand - ta.z), lo, ld);
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(slot);
            for (;;) {
                const int champ = (int)(unsigned)(cur >> 32);
                if (champ == cand || champ < ta.z || champ >= ta.z + ta.w) break;
                const double t_h = triangle_t64(sc.tris, tb.z + (champ - ta.z), lo, ld);
                if (!(t_c < t_h || (t_c == t_h && cand < champ))) break;   // strict `<`; the first in the reference's order wins exact ties
                const unsigned long long nv = ((unsigned long long)(unsigned)cand << 32) | (unsigned long long)__float_as_uint((float)t_c);
                const unsigned long long old = atomicCAS(slot, cur, nv);
                if (old == cur) break;
                cur = old;
            }
        }
    }
}

// bvh_wave: BVHAggregateNode.intersect (src/aggregates.js:207-225) for the rays of the work list.
// Persistent threads, one ray per lane, lanes refilled individually from a warp-local pool of list
// entries (one atomicAdd per JSRT_POOL_BATCH rays): a lane whose walk ends takes the next ray instead of
// idling while its neighbours finish (walk lengths differ by two orders of magnitude).
//
// Shared-memory staging.  The first sc.n_staged nodes of the scene's node array — the top levels of every tree, in
// breadth-first order (scene_flatten.cpp: assembleNodes) — are copied into shared memory when the CTA starts; a node
// step reads its 32 bytes from there when the index is below n_staged and through L1 otherwise.  Every ray walks
// through the top levels, and a warp's 32 lanes fetch 32 different nodes: in L1 that is one wavefront per 128-byte
// line touched (32 per load instruction, l1tex was at 51-60 % of its peak in round 1) and an L2 round trip for each
// of the 13-17 % that miss; in shared memory it is a bank-conflicted but local access.  With one 1 024-thread CTA per
// SM and 128 KB staged, bunny_path's tree (5.5 k nodes with two-triangle leaves) is three quarters resident.
//
// DIRECT (shadow rays of scenes without SDFs): the queue holds walkers only — shade_kernel has already run the analytic
// primitives and the root boxes and accumulated or dropped every ray that needs no walk — so entry i of the queue is
// ray i: o.xyz | pixel, d.xyz | pass, contribution.rgb | first BVH.  No work list, no partial-hit buffer.
// MESH: every BVHAggregate of the scene is a pure identity-transform triangle mesh (tri_base >= 0), so the leaf code of the
// general primitives — a call into every geometry's intersect with its own ray transform — is compiled out: no spills
// left in the kernel, bunny_path +4.9 %, dragon +4.5 % (profiles/r2_ab.md §2).
template <int MODE, bool COUNT, bool HAS_SDF, bool DIRECT, bool TLAS, bool MESH>
JSRT_DEV void bvh_wave(const DeviceScene& sc, const TraceIO& io, Work* work_primary, Work* work_other, const float4* __restrict__ s_nodes) {
    constexpr bool ANY_HIT = (MODE == TM_SHADOW);
    constexpr int BATCH = JSRT_POOL_BATCH;      // list entries fetched per atomicAdd
    // A warp runs three phases per iteration, each only when enough lanes need it, so that the
    // rarely-needed code (ray hand-over, leaf tests) executes with many lanes instead of one or two:
    constexpr int REFILL_T = JSRT_REFILL_T;     // finish + refill when this many lanes are idle (or nothing else is left)
    constexpr int LEAF_T = JSRT_LEAF_T;         // test postponed leaves when this many lanes hold one (or one must be flushed)
    constexpr int NODE_STEPS = JSRT_NODE_STEPS; // nodes walked per iteration between the warp votes
    constexpr int PROG_T = JSRT_PROG_T;         // ... or when fewer than this many lanes can still walk
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = DIRECT ? min(*io.count, io.cap) : min(*io.list_count, io.cap);
    const int n_staged = sc.n_staged;
    int pool_next = 0, pool_end = 0;      // warp-uniform
    bool exhausted = false;               // warp-uniform
    // per-lane ray state, kept small (64 registers): the world-space ray is not kept — it is re-read from the queue in
    // the rare case that a second BVH has to be entered — and node links are absolute, so no per-tree base is carried.
    int cur = -1;                         // >= 0: ray index, walking; -1: none; <= -2: ray -2 - cur has finished, result not yet written
    int pending = -1;                     // postponed leaf word; -(leaf + 2): a second leaf is waiting behind it
    LocalRay r; r.lo = f3(0, 0, 0); r.ld = f3(0, 0, 1); r.inv = r.ld; r.sgn = r.ld; r.par = false;
    float minD_v = 0.f;                   // extend only (the shadow window is a constant)
    #define JSRT_MIND ((MODE == TM_SHADOW) ? 0.0001f : minD_v)
    #define JSRT_MAXD ((MODE == TM_SHADOW) ? 1.0f : CUDART_INF_F)
    float hi = CUDART_INF_F;              // min(maxD, local_best, best.t): the pruning bound of :209
    Hit best; best.t = CUDART_INF_F; best.prim = -1; best.top = -1; best.t_lo = 0.f;
    int bi = 0;
    int node_i = kNodeEnd, first_prim = 0, tri_base = -1;
    // Top-level BVH (scenes with many aggregates): the lane walks it with the world ray (tri_base == kTlasLevel marks that
    // level); a leaf names an aggregate (`want`), which is entered like any BVH, and when that tree is finished the walk
    // resumes at the top-level node `resume`.  One level of nesting, one extra register.
    constexpr int kTlasLevel = -2;
    constexpr bool use_tlas = TLAS;          // (a template flag: the extra state must not cost the single-mesh kernels registers)
    int want = -1, resume = kNodeEnd;
    float local_best = CUDART_INF_F, local_lo = 0.f; int local_prim = -1;
    Work* work = work_other;
    const float4* const all_nodes = reinterpret_cast<const float4*>(sc.nodes);
    // 32-bit shared-window address of the staged block, formed once (the generic pointer made the compiler rebuild it
    // from SR_CgaCtaId at every node step: 4 of the 52 instructions of the loop)
    unsigned s_base = (unsigned)__cvta_generic_to_shared(s_nodes);
    asm volatile("" : "+r"(s_base));      // opaque: otherwise it is rebuilt from SR_CgaCtaId (an S2R + two LEAs) at every node step
    const unsigned s_base_hi = s_base + 16u * (unsigned)n_staged;      // second halves of the staged records (JSRT_STAGE_SOA)

    // enter BVH number `bi` of the scene (Aggregate.intersect / BVHAggregate.intersect, src/aggregates.js:43-46)
    auto enter = [&](float3 o, float3 d) {
        const int4* tp = reinterpret_cast<const int4*>(sc.tops + __ldg(sc.bvh_tops + bi));
        const int4 ta = __ldg(tp), tb = __ldg(tp + 1);      // kind, xform, first_prim, prim_count | first_node, node_count, tri_base, layouts
        r = make_local_ray(load_xform(sc.xforms, ta.y), o, d);
        // closest-hit rays pick the link order that visits the nearer child first; any-hit (shadow) rays gain nothing from
        // it (measured: +18 % nodes per shadow ray on bunny_path) and keep the reference order
        // (layout 7 = higher child first on every axis = the reference's greater-child-first order)
        const int octant = ((tb.w & 0xff) != 8) ? 0 : ANY_HIT ? 7 : ((r.ld.x < 0.f ? 1 : 0) | (r.ld.y < 0.f ? 2 : 0) | (r.ld.z < 0.f ? 4 : 0));
        node_i = tb.x + octant * (tb.w >> 8); first_prim = ta.z; tri_base = tb.z;
        local_best = CUDART_INF_F; local_prim = -1; local_lo = 0.f; pending = -1;
        hi = fminf(JSRT_MAXD, best.t);
    };

    for (;;) {
        // ---- phase 1: write finished rays, hand out new ones ------------------------------
        {
            const unsigned idle_mask = __ballot_sync(FULL, cur < 0);
            const int n_idle = __popc(idle_mask);
            if (n_idle >= REFILL_T || idle_mask == FULL) {
                if (cur < -1) {
                    const int ray = -2 - cur;
                    if (DIRECT || io.final_pass) finish_ray<MODE>(io, ray, best, io.o[ray]);
                    else io.hits[ray] = make_float4(best.t, __int_as_float(best.prim), __int_as_float(best.top), best.t_lo);
                    cur = -1;
                }
                if (pool_next >= pool_end && !exhausted) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(io.cursor, BATCH);
                    base = __shfl_sync(FULL, base, 0);
                    pool_next = base; pool_end = min(base + BATCH, n);
                    if (base >= n) { exhausted = true; pool_next = pool_end = 0; }
                }
                const int avail = pool_end - pool_next;
                if (avail <= 0) { if (exhausted && __all_sync(FULL, cur < 0)) break; }
                else {
                    const int rank = __popc(idle_mask & ((1u << lane) - 1u));
                    if (rank < avail) {             // only idle lanes have a rank that is meaningful: cur < 0 here for them
                        if (cur < 0) {
                            float4 o4, d4;
                            if (DIRECT) {
                                cur = pool_next + rank;
                                o4 = io.o[cur]; d4 = io.d[cur];
                                bi = __float_as_int(io.c[cur].w);
                                best.t = CUDART_INF_F; best.prim = -1; best.top = -1; best.t_lo = 0.f;
                            } else {
                                const int2 e = __ldg(reinterpret_cast<const int2*>(io.list) + pool_next + rank);
                                cur = e.x; bi = e.y;
                                o4 = io.o[cur]; d4 = io.d[cur];
                                const float4 h4 = io.hits[cur];
                                bool primary; float mx;
                                ray_window<MODE>(d4, minD_v, mx, primary);
                                if (COUNT) work = primary ? work_primary : work_other;
                                best.t = h4.x; best.prim = __float_as_int(h4.y); best.top = __float_as_int(h4.z); best.t_lo = h4.w;
                            }
                            if (use_tlas) {
                                r = make_ray(f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
                                node_i = sc.tlas_root; tri_base = kTlasLevel; want = -1; resume = kNodeEnd; pending = -1;
                                local_best = CUDART_INF_F; local_prim = -1; local_lo = 0.f;
                                hi = fminf(JSRT_MAXD, best.t);
                            } else enter(f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
                            // the producer has tested this tree's root box; an inner root needs no second test
                            const int root_word = __float_as_int(__ldg(all_nodes + 2 * node_i + 1).w);
                            if (root_word < 0) node_i = root_word & 0x7fffffff;
                        }
                    }
                    pool_next += min(avail, n_idle);
                }
            }
        }
        const bool active = cur >= 0;

        // ---- phase 2: postponed leaf tests (src/aggregates.js:211-218) -------------------------
        // A lane that reaches a leaf parks it and keeps walking; the leaf is tested when enough
        // lanes hold one, when the lane reaches its next leaf, or when its walk ends.  Testing later
        // only delays the `ret.distance` update that prunes the walk; the result is the same.
        // Policy: run the leaf phase when LEAF_T lanes hold a leaf, or when fewer than PROG_T lanes could still
        // walk (the others are parked, blocked or idle): a blocked or finished lane waits for company instead of
        // dragging the whole warp through the triangle code with a handful of lanes.
        const unsigned pend_mask = __ballot_sync(FULL, active && pending != -1);
        const unsigned prog_mask = __ballot_sync(FULL, active && node_i != kNodeEnd && pending >= -1);
        if (pend_mask && (__popc(pend_mask) >= LEAF_T || __popc(prog_mask) < PROG_T)) {
            if (active && pending != -1) {
                const int leaf = (pending < -1) ? -(pending + 2) : pending;      // a blocked lane stores -(leaf + 2)
                const int cnt = (int)((unsigned)leaf >> 24), rel = leaf & 0xffffff;
                // (two loops, so that the compiler cannot hoist the general primitives' ray setup in front of the
                // triangle path: it did, 140 instructions with 25 FP64 ones per leaf — profiles/r2_ab.md)
                if (MESH || tri_base >= 0) {
                    for (int k = 0; k < cnt; ++k) {
                        if (COUNT) ++work->leaf_prims;
                        const int pi = first_prim + rel + k;
                        const float t = triangle_intersect(sc.tris, tri_base + rel + k, r.lo, r.ld, JSRT_MIND, hi);
                        // :213 with the rank tie rule (see the header comment).  Closest-hit rays report FP32 near-ties between
                        // two triangles to the tie list (rare: one atomic + one store); tie_wave settles them in f64 afterwards
                        if (t > JSRT_MIND && t < JSRT_MAXD) {
                            const bool take = t < local_best || (t == local_best && pi < local_prim);
#if JSRT_TRI_TIE
                            if (!ANY_HIT && fabsf(t - local_best) <= 1e-6f * fabsf(t)) {        // (never true while local_best is +Inf)
                                const int e = atomicAdd(io.tie_count, 1);
                                if (e < io.tie_cap) io.tie_list[e] = make_int4(cur, bi, take ? pi : local_prim, take ? local_prim : pi);
                            }
#endif
                            if (take) { local_best = t; local_prim = pi; local_lo = 0.f; }
                        }
                    }
                } else if (!MESH) {
                    for (int k = 0; k < cnt; ++k) {
                        if (COUNT) ++work->leaf_prims;
                        const int pi = first_prim + rel + k;
                        float tl = 0.f;
                        const float t = placed_prim_intersect<HAS_SDF>(sc, pi, r.lo, r.ld, JSRT_MIND, JSRT_MAXD, hi, ANY_HIT, COUNT ? &work->sdf_evals : nullptr, &tl);
                        if (t > JSRT_MIND && t < JSRT_MAXD && (t < local_best || (t == local_best && pi < local_prim))) { local_best = t; local_prim = pi; local_lo = tl; }
                    }
                }
                pending = -1;
                // (closest-hit: the bound keeps 2 ppm of slack so that a candidate tied with the current hit in FP32 still
                // reaches the tie-break above; acceptance itself compares against local_best)
                hi = fminf(hi, (ANY_HIT || !JSRT_TRI_TIE) ? local_best : fmaf(fabsf(local_best), 2e-6f, local_best));
                if (ANY_HIT && local_prim >= 0) node_i = kNodeEnd;
            }
        }

        if (active) {
            if (use_tlas && want >= 0) {
                // ---- a leaf of the top-level BVH: enter that aggregate -------------------------
                const float4 o4 = io.o[cur], d4 = io.d[cur];
                bi = want; want = -1;
                enter(f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
            } else if (use_tlas && node_i == kNodeEnd && pending == -1) {
                if (tri_base != kTlasLevel) {
                    // ---- an aggregate's tree is finished: merge, back to the top-level walk ----
                    const int top_i = __ldg(sc.bvh_tops + bi);
                    if (local_best > JSRT_MIND && local_best < JSRT_MAXD && better_hit(local_best, top_i, best)) {
                        best.t = local_best; best.prim = local_prim; best.top = top_i; best.t_lo = local_lo;
                    }
                    if (resume == kNodeEnd || (ANY_HIT && best.prim >= 0)) cur = -2 - cur;
                    else {
                        const float4 o4 = io.o[cur], d4 = io.d[cur];
                        r = make_ray(f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
                        node_i = resume; resume = kNodeEnd; tri_base = kTlasLevel;
                        local_best = CUDART_INF_F; local_prim = -1; local_lo = 0.f;
                        hi = fminf(JSRT_MAXD, best.t);
                    }
                } else cur = -2 - cur;              // the top-level walk is over
            } else if (node_i == kNodeEnd && pending == -1) {
                // ---- tree finished: merge into the running closest hit, next BVH or done -------
                const int top_i = __ldg(sc.bvh_tops + bi);
                if (local_best > JSRT_MIND && local_best < JSRT_MAXD && better_hit(local_best, top_i, best)) {
                    best.t = local_best; best.prim = local_prim; best.top = top_i; best.t_lo = local_lo;
                }
                ++bi;
                if (bi >= sc.n_bvh || (ANY_HIT && best.prim >= 0)) cur = -2 - cur;
                else {
                    const float4 o4 = io.o[cur], d4 = io.d[cur];
                    if (sc.use_wbox) {          // skip the aggregates whose world box the ray misses (or reaches behind the closest hit)
                        const float3 winv = f3(JSRT_RCP(d4.x), JSRT_RCP(d4.y), JSRT_RCP(d4.z));
                        const float whi = fminf(JSRT_MAXD, best.t);
                        while (bi < sc.n_bvh && !wbox_hit(sc.wboxes, bi, f3(o4.x, o4.y, o4.z), winv, JSRT_MIND, whi)) ++bi;
                    }
                    if (bi >= sc.n_bvh) cur = -2 - cur;
                    else enter(f3(o4.x, o4.y, o4.z), f3(d4.x, d4.y, d4.z));
                }
            } else {
                // ---- phase 3: up to NODE_STEPS nodes of BVHAggregateNode.intersect (src/aggregates.js:207-225)
                // per iteration, so the warp votes of phases 1-2 are paid once per few nodes
                // The step is branch-free after the box test (selects instead of the nested if / else: the compiler's version
                // spent 25 issue slots per step on BSSY / BRA / BSYNC with 3-13 lanes, profiles/r2/ncu_r2h_*), and rays with a
                // parallel axis (AABB.get_intersects' `else` rule, src/geometry.js:194,205) take their own copy of the loop so
                // that the common one carries no test for them.
                #define JSRT_NODE_STEP(SLAB)                                                                                       \
                    {                                                                                                              \
                        float4 n0, n1;                                                                                             \
                        if (node_i < n_staged) {                                                                                   \
                            if (JSRT_STAGE_SOA) lds_node_soa(s_base + 16u * (unsigned)node_i, s_base_hi + 16u * (unsigned)node_i, n0, n1);  \
                            else lds_node(s_base + 32u * (unsigned)node_i, n0, n1);                                                \
                        }                                                                                                          \
                        else load_node(all_nodes + 2 * node_i, n0, n1);                                                            \
                        const int skip = __float_as_int(n1.z), word = __float_as_int(n1.w);                                        \
                        if (COUNT) ++work->nodes;                                                                                  \
                        const bool hit_box = SLAB;                                                                                 \
                        const bool leaf_hit = hit_box && word >= 0;                                                                \
                        int next = hit_box ? (word & 0x7fffffff) : skip;              /* inner: the hit link; missed: the skip link */ \
                        if (use_tlas && tri_base == kTlasLevel) {                                                                  \
                            if (leaf_hit) { want = word & 0xffffff; resume = skip; next = kNodeEnd; }   /* an aggregate: entered after the loop */ \
                        } else {                                                                                                   \
                            /* leaf: park it and go on; one leaf already parked: block here (same node again) until it is tested */ \
                            const bool park = pending == -1;                                                                       \
                            const int parked = park ? word : -(pending + 2);                                                       \
                            const int after = park ? skip : node_i;                                                                \
                            pending = leaf_hit ? parked : pending;                                                                 \
                            next = leaf_hit ? after : next;                                                                        \
                        }                                                                                                          \
                        node_i = next;                                                                                             \
                    }
                if (!r.par) {
                    #pragma unroll 1
                    for (int rep = 0; rep < NODE_STEPS && node_i != kNodeEnd && pending >= -1; ++rep) JSRT_NODE_STEP(slab_fast(n0, n1, r, JSRT_MIND, hi))
                } else {
                    #pragma unroll 1
                    for (int rep = 0; rep < NODE_STEPS && node_i != kNodeEnd && pending >= -1; ++rep) JSRT_NODE_STEP(slab_general(n0, n1, r, JSRT_MIND, JSRT_MAXD, hi))
                }
                #undef JSRT_NODE_STEP
            }
        }
    }
    #undef JSRT_MIND
    #undef JSRT_MAXD
}

// ---------------------------------------------------------------------------------
// SDFGeometry.intersect (src/sdf.js:12-40) for the top-level SDF primitives of a whole
// queue: persistent threads, one ray per lane, one sphere-tracing step (one
// root_sdf.distance evaluation) per iteration, lanes refilled in batches.  Step counts
// range from 0 (ray misses the SDF's box) to max_samples, and the bytecode itself is
// branch-free, so with per-lane refill every lane of a warp executes the interpreter
// in lockstep.  Arithmetic: the reference's (device_math.cuh).
// lanes idle before the warp stops to hand out new rays: one distance evaluation costs ~2 000 instructions and the
// hand-over ~200, so a warp refills as soon as a few lanes are free (measured: profiles/r2/ab_r2l_*)
#ifndef JSRT_SDF_WBOX
#define JSRT_SDF_WBOX 1       // FP32 reject against the SDF's padded world box in front of the f64 set-up (0: A/B)
#endif
#ifndef JSRT_SDF_REFILL_T
#define JSRT_SDF_REFILL_T 8
#endif
template <int MODE, bool COUNT, bool RTU>
JSRT_DEV void sdf_wave(const DeviceScene& sc, const TraceIO& io, Work* work_primary, Work* work_other) {
    constexpr bool ANY_HIT = (MODE == TM_SHADOW);
    constexpr int BATCH = 64, REFILL_T = JSRT_SDF_REFILL_T;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = min(*io.count, io.cap);
    int pool_next = 0, pool_end = 0; bool exhausted = false;       // warp-uniform

    int cur = -1; bool done = false;
    float4 o4 = make_float4(0, 0, 0, 0);
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1), lo = o, ld = d;
    float minD = 0.f, maxD = CUDART_INF_F;
    Hit best; best.t = CUDART_INF_F; best.prim = -1; best.top = -1; best.t_lo = 0.f;
    int si = 0, top_i = 0, prim_i = 0, step = 0, max_steps = 0;
    bool marching = false;
    double t = 0, t_lo = 0, t_hi = 0, rd_norm = 1, eps = 0, max_trace = 0, nstep = 0;
    const SdfInstr* prog = nullptr;
    Work* work = work_other;
    // SDFGeometry.materialData's forward-difference normal (src/sdf.js:42-46) is computed here, as four more
    // lockstep evaluations by the lane that found the hit, instead of in the shade kernel where only the
    // lanes with an SDF hit would run the interpreter (measured: 114 of 331 ms on SDF_Menger at 1080p).
    int nphase = -1;                      // -1: marching; 0..3: evaluating d0, dx, dy, dz
    float3 hp = f3(0, 0, 0), nrm = f3(0, 0, 0); bool has_nrm = false;
    const SdfInstr* hprog = nullptr; double hstep = 0, nd0 = 0, ndx = 0, ndy = 0;

    // set up the march through SDF primitive number `si` (Primitive.intersect + the head of SDFGeometry.intersect)
    auto enter = [&]() {
        marching = false;
        const float3 winv = f3(JSRT_RCP(d.x), JSRT_RCP(d.y), JSRT_RCP(d.z));      // (three SFU reciprocals per call: cheaper than three registers kept across the march)
        while (si < sc.n_sdf_tops && !marching) {
            // FP32 reject against the padded world box first: most rays of an SDF scene never meet the SDF, and the exact
            // test below costs two f64 matrix products and six f64 divisions per ray
            if (JSRT_SDF_WBOX && !wbox_hit(sc.sdf_wboxes, si, o, winv, minD, fminf(maxD, best.t))) { ++si; continue; }
            top_i = __ldg(sc.sdf_tops + si);
            prim_i = __ldg(&sc.tops[top_i].first_prim);
            const int4* pp = reinterpret_cast<const int4*>(sc.prims + prim_i);
            const int4 pa = __ldg(pp); const int flags = __ldg(reinterpret_cast<const int*>(pp + 1));
            if (!(ANY_HIT && !(flags & PF_CASTS_SHADOW))) {                   // src/world.js:117-118
                const double* m = sc.xforms64[pa.w].m;
                lo = xf64_apply(m, o, 1.0); ld = xf64_apply(m, d, 0.0);        // ray.getTransformed(inv_transform)
                const SdfProgram& pr = sc.sdfs[pa.y];
                double b0, b1;
                // a hit beyond the running closest hit cannot win, so the march may stop there
                const double cap = fmin((double)maxD, (double)best.t);
                if (aabb_intersects_f64(f3(pr.cx, pr.cy, pr.cz), f3(pr.hx, pr.hy, pr.hz), lo, ld, (double)minD, (double)maxD, b0, b1)) {
                    t_lo = jsd_max((double)minD, b0); t_hi = jsd_min((double)maxD, b1);
                    if (t_lo <= cap) {
                        t = t_lo; step = 0; max_steps = pr.max_samples; eps = pr.distance_epsilon; max_trace = pr.max_trace_distance; nstep = pr.normal_step_size;
                        if (cap < t_hi) t_hi = cap;
                        rd_norm = sqrt(ddot4(ld.x, ld.y, ld.z, 0.0, ld.x, ld.y, ld.z, 0.0));
                        prog = sc.sdf_code + pr.first_instr;
                        marching = max_steps > 0;
                    }
                }
            }
            if (!marching) ++si;
        }
        if (!marching) {
            if (MODE == TM_EXTEND && hprog != nullptr) nphase = 0;       // closest hit is an SDF found here: normal next
            else done = true;
        }
    };

    for (;;) {
        // ---- write finished rays, hand out new ones.  Repeated until fewer than REFILL_T lanes are idle or the queue is
        // empty: a new ray that misses the SDF's box (or needs no march) is done at once, and one evaluation of the
        // interpreter costs thousands of instructions, so no lane should enter it empty-handed while rays are waiting
        // (measured before: 17 of 32 lanes active, profiles/r1_s3) ----
        bool all_done = false;
        for (;;) {
        const unsigned idle_mask = __ballot_sync(FULL, cur < 0 || done);
        const int n_idle = __popc(idle_mask);
        if (!(n_idle >= REFILL_T || idle_mask == FULL)) break;
        {
            if (cur >= 0 && done) {
                finish_ray<MODE>(io, cur, best, o4);
                if (MODE == TM_EXTEND && io.aux) io.aux[cur] = make_float4(nrm.x, nrm.y, nrm.z, has_nrm ? 1.f : 0.f);
                cur = -1; done = false;
            }
            if (pool_next >= pool_end && !exhausted) {
                int base = 0;
                if (lane == 0) base = atomicAdd(io.cursor, BATCH);
                base = __shfl_sync(FULL, base, 0);
                pool_next = base; pool_end = min(base + BATCH, n);
                if (base >= n) { exhausted = true; pool_next = pool_end = 0; }
            }
            const int avail = pool_end - pool_next;
            if (avail > 0) {
                const int rank = __popc(idle_mask & ((1u << lane) - 1u));
                if (cur < 0 && rank < avail) {
                    cur = pool_next + rank;
                    o4 = io.o[cur];
                    const float4 d4 = io.d[cur], h4 = io.hits[cur];
                    o = f3(o4.x, o4.y, o4.z); d = f3(d4.x, d4.y, d4.z);
                    bool primary;
                    ray_window<MODE>(d4, minD, maxD, primary);
                    if (COUNT) work = primary ? work_primary : work_other;
                    best.t = h4.x; best.prim = __float_as_int(h4.y); best.top = __float_as_int(h4.z); best.t_lo = h4.w;
                    si = 0; done = false; nphase = -1; hprog = nullptr; has_nrm = false; nrm = f3(0, 0, 0);
                    if (ANY_HIT && best.prim >= 0) cur = -1;
                    else enter();
                }
                pool_next += min(avail, n_idle);
            } else { all_done = exhausted && __all_sync(FULL, cur < 0); break; }
        }
        }
        if (all_done) break;
        if (cur >= 0 && !done) {
            // ---- one distance evaluation per iteration: a step of the sphere-tracing loop (src/sdf.js:22-38)
            // or one of the four samples of the forward-difference normal (src/sdf.js:42-46) -----------------
            float3 pt; const SdfInstr* pg;
            if (nphase < 0) { pt = ray_point_f64(lo, ld, t); pg = prog; }
            else {
                const float fs = (float)hstep;            // Vec.axis(i, 4, step) stores the step as f32
                pt = f3(nphase == 1 ? (float)dadd(hp.x, fs) : hp.x, nphase == 2 ? (float)dadd(hp.y, fs) : hp.y, nphase == 3 ? (float)dadd(hp.z, fs) : hp.z);
                pg = hprog;
            }
            const double dist = sdf_eval<RTU ? 1 : 0>(pg, sc.xforms64, pt);
            if (COUNT) ++work->sdf_evals;
            if (nphase >= 0) {
                if (nphase == 0) nd0 = dist; else if (nphase == 1) ndx = dist; else if (nphase == 2) ndy = dist;
                if (nphase == 3) {
                    const float nx = (float)(dsub(ndx, nd0) / hstep), ny = (float)(dsub(ndy, nd0) / hstep), nz = (float)(dsub(dist, nd0) / hstep);
                    const double nn = sqrt(ddot4(nx, ny, nz, 0.0, nx, ny, nz, 0.0));
                    nrm = (nn > 0.00001) ? f3((float)dmul(nx, 1.0 / nn), (float)dmul(ny, 1.0 / nn), (float)dmul(nz, 1.0 / nn)) : f3(nx, ny, nz);
                    has_nrm = true; nphase = -1; done = true;
                } else ++nphase;
            } else {
                bool over = false;
                if (!isfinite(dist)) over = true;
                else if (dist <= eps) {
                    const float tf = (float)t;
                    if (tf > minD && tf < maxD && better_hit(tf, top_i, best)) {
                        best.t = tf; best.prim = prim_i; best.top = top_i; best.t_lo = (float)(t - (double)tf);
                        // the local hit point exactly as Primitive.color recomputes it (src/world.js:127-131)
                        hp = ray_point_f64(lo, ld, (double)best.t + (double)best.t_lo); hprog = prog; hstep = nstep;
                    }
                    over = true;
                } else {
                    t = dadd(t, dist / rd_norm);
                    if (t < t_lo || t > t_hi || dmul(dsub(t, t_lo), rd_norm) > max_trace || ++step >= max_steps) over = true;
                }
                if (over) {
                    ++si;
                    if (ANY_HIT && best.prim >= 0) done = true; else enter();
                }
            }
        }
    }
}

}  // namespace jsrt
