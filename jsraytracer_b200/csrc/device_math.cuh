// Device-side ray/geometry math: FP32 restatement of the reference's
// per-primitive `intersect` / `materialData` / `sampleSurface`
// (src/geometry.js, src/sdf.js) and the AABB slab test (src/geometry.js:189-209).
// The reference computes in f64 on f32-stored vectors; here everything is FP32
// with the same formulas in the same order, and explicit __fmul_rn/__fadd_rn
// where the reference rounds to f32 between a product and a sum
// (`origin.plus(direction.times(t))`, src/math.js:297-299).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "scene_types.h"

namespace jsrt {

#define JSRT_DEV __device__ __forceinline__

JSRT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
JSRT_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
JSRT_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
JSRT_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
JSRT_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
JSRT_DEV float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
JSRT_DEV float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Vec.normalized(): unchanged when the norm is <= 1e-5 (src/math.js:242-245)
// (FP32 with the SFU reciprocal square root, ~2 ulp, instead of IEEE sqrt + IEEE division: 26 -> 10 instructions, and
// shade_kernel normalises ten vectors per hit — 8.5 % of its instructions, profiles/r2/ncu_r2j_*.  JSRT_FAST_NORMALIZE=0: A/B)
#ifndef JSRT_FAST_NORMALIZE
#define JSRT_FAST_NORMALIZE 1
#endif
JSRT_DEV float3 normalized3(float3 a) {
#if JSRT_FAST_NORMALIZE
    const float n2 = dot3(a, a);
    return (n2 > 0.00001f * 0.00001f) ? a * rsqrtf(n2) : a;
#else
    const float n = sqrtf(dot3(a, a)); return (n > 0.00001f) ? a * (1.0f / n) : a;
#endif
}
// 1 / x on the SFU (1 ulp, denormals flushed): reciprocal ray directions of the slab tests
JSRT_DEV float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// Ray.getPoint: two f32 roundings (product, then sum)
JSRT_DEV float3 ray_point(float3 o, float3 d, float t) {
    return f3(__fadd_rn(o.x, __fmul_rn(d.x, t)), __fadd_rn(o.y, __fmul_rn(d.y, t)), __fadd_rn(o.z, __fmul_rn(d.z, t)));
}
JSRT_DEV float js_sign(float x) { return (x > 0.f) ? 1.f : (x < 0.f ? -1.f : x); }
// Math.min / Math.max propagate NaN
JSRT_DEV float js_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }
JSRT_DEV float js_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }

// Un-contracted f64 helpers: the compiler must not fuse a*b+c here (JS has no FMA).
JSRT_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
JSRT_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
JSRT_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
// Vec.dot for 3 / 4 components: products summed left to right (src/math.js:252-254)
JSRT_DEV double ddot3(double ax, double ay, double az, double bx, double by, double bz) { return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz)); }
JSRT_DEV double ddot4(double ax, double ay, double az, double aw, double bx, double by, double bz, double bw) {
    return dadd(dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz)), dmul(aw, bw));
}
JSRT_DEV double jsd_min(double a, double b) { return (a != a || b != b) ? CUDART_NAN : fmin(a, b); }
JSRT_DEV double jsd_max(double a, double b) { return (a != a || b != b) ? CUDART_NAN : fmax(a, b); }

struct XformReg { float4 r0, r1, r2; };
JSRT_DEV XformReg load_xform(const Xform* __restrict__ xs, int i) {
    const float4* p = reinterpret_cast<const float4*>(xs + i);
    XformReg x; x.r0 = __ldg(p); x.r1 = __ldg(p + 1); x.r2 = __ldg(p + 2); return x;
}
// Mat.times(Vec) with w = 1 / w = 0 (src/math.js:392-397)
JSRT_DEV float3 xf_point(const XformReg& m, float3 p) {
    return f3(m.r0.x * p.x + m.r0.y * p.y + m.r0.z * p.z + m.r0.w,
              m.r1.x * p.x + m.r1.y * p.y + m.r1.z * p.z + m.r1.w,
              m.r2.x * p.x + m.r2.y * p.y + m.r2.z * p.z + m.r2.w);
}
JSRT_DEV float3 xf_dir(const XformReg& m, float3 d) {
    return f3(m.r0.x * d.x + m.r0.y * d.y + m.r0.z * d.z,
              m.r1.x * d.x + m.r1.y * d.y + m.r1.z * d.z,
              m.r2.x * d.x + m.r2.y * d.y + m.r2.z * d.z);
}
// inv_transform.transposed().times(n) for a direction n (w = 0)
JSRT_DEV float3 xf_normal(const XformReg& m, float3 n) {
    return f3(m.r0.x * n.x + m.r1.x * n.y + m.r2.x * n.z,
              m.r0.y * n.x + m.r1.y * n.y + m.r2.y * n.z,
              m.r0.z * n.x + m.r1.z * n.y + m.r2.z * n.z);
}
// a * b for two affine maps (a applied after b): prim.inv_transform.times(ancestorInvTransform), src/world.js:126
JSRT_DEV XformReg xf_compose(const XformReg& a, const XformReg& b) {
    XformReg r;
    r.r0 = make_float4(a.r0.x * b.r0.x + a.r0.y * b.r1.x + a.r0.z * b.r2.x, a.r0.x * b.r0.y + a.r0.y * b.r1.y + a.r0.z * b.r2.y,
                       a.r0.x * b.r0.z + a.r0.y * b.r1.z + a.r0.z * b.r2.z, a.r0.x * b.r0.w + a.r0.y * b.r1.w + a.r0.z * b.r2.w + a.r0.w);
    r.r1 = make_float4(a.r1.x * b.r0.x + a.r1.y * b.r1.x + a.r1.z * b.r2.x, a.r1.x * b.r0.y + a.r1.y * b.r1.y + a.r1.z * b.r2.y,
                       a.r1.x * b.r0.z + a.r1.y * b.r1.z + a.r1.z * b.r2.z, a.r1.x * b.r0.w + a.r1.y * b.r1.w + a.r1.z * b.r2.w + a.r1.w);
    r.r2 = make_float4(a.r2.x * b.r0.x + a.r2.y * b.r1.x + a.r2.z * b.r2.x, a.r2.x * b.r0.y + a.r2.y * b.r1.y + a.r2.z * b.r2.y,
                       a.r2.x * b.r0.z + a.r2.y * b.r1.z + a.r2.z * b.r2.z, a.r2.x * b.r0.w + a.r2.y * b.r1.w + a.r2.z * b.r2.w + a.r2.w);
    return r;
}

// AABB.get_intersects (src/geometry.js:189-209), centre / half-size form.
JSRT_DEV bool aabb_intersects(float3 c, float3 h, float3 o, float3 d, float minD, float maxD, float& t_min, float& t_max) {
    t_min = -CUDART_INF_F; t_max = CUDART_INF_F;
    const float3 p = c - o;
    const float eps = 0.0000001f;
#define JSRT_SLAB(PI, HI, DI)                                                   \
    if (fabsf(DI) > eps) {                                                      \
        float t1 = (PI + HI) / DI, t2 = (PI - HI) / DI;                         \
        if (t1 > t2) { const float tmp = t1; t1 = t2; t2 = tmp; }               \
        if (t1 > t_min) t_min = t1;                                             \
        if (t2 < t_max) t_max = t2;                                             \
        if (t_min > t_max || t_max < minD || t_min > maxD) return false;        \
    } else if (fabsf(PI) > HI) return false;
    JSRT_SLAB(p.x, h.x, d.x)
    JSRT_SLAB(p.y, h.y, d.y)
    JSRT_SLAB(p.z, h.z, d.z)
#undef JSRT_SLAB
    return true;
}

// SimplePlane.intersect src/geometry.js:246-248
JSRT_DEV float plane_t(float3 o, float3 d) { return (d.z != 0.f) ? -o.z / d.z : -CUDART_INF_F; }

// Sphere.staticIntersect src/geometry.js:429-442.  The reference solves the quadratic in f64, where
// its textbook root formula is harmless; in FP32 it is not, and the place it breaks is systematic: a ray
// that starts on the sphere (every reflection / refraction / shadow ray leaving one) has c = |o|^2 - 1
// equal to the rounding residue of its f32 origin (~3e-7), an FP32 dot product adds an error of the same
// size, and at grazing angles the near root c / 2b then crosses the 1e-4 self-hit threshold more often than
// in the reference (measured: silhouette pixels of tests/refraction systematically darker, 47 dB).
// So: c is accumulated in f64 (three DFMAs — it is exact for f32 inputs up to the final rounding), and
// the roots come from the cancellation-free form q = -(b + sign(b) sqrt(disc)), {q / a, c / q}, which
// returns the same two numbers as (-b +- sqrt(disc)) / a would in exact arithmetic.
JSRT_DEV float sphere_intersect(float3 o, float3 d, float minD) {
    const float a = dot3(d, d), b = dot3(d, o);
    const float c = (float)fma((double)o.x, (double)o.x, fma((double)o.y, (double)o.y, fma((double)o.z, (double)o.z, -1.0)));
    float disc = fmaf(b, b, -a * c);
    if (fabsf(disc) < 1e-5f * fmaf(b, b, fabsf(a * c))) {
        // near the silhouette b^2 and a*c cancel; with strongly anisotropic transforms (a bolt scaled 0.01 x 0.01 x 1.5
        // in tests/starwars) the FP32 residue decides hit or miss for the whole primitive.  Redo the discriminant in f64.
        const double a64 = ddot3(d.x, d.y, d.z, d.x, d.y, d.z), b64 = ddot3(d.x, d.y, d.z, o.x, o.y, o.z);
        const double c64 = fma((double)o.x, (double)o.x, fma((double)o.y, (double)o.y, fma((double)o.z, (double)o.z, -1.0)));
        disc = (float)dsub(dmul(b64, b64), dmul(a64, c64));
    }
    if (disc < 0.f || a == 0.f) return -CUDART_INF_F;
    const float big = sqrtf(disc);
    const float q = -(b + copysignf(big, b));
    const float ra = q / a, rb = (q != 0.f) ? c / q : 0.f;
    // t1 = (-b + big) / a, t2 = (-b - big) / a
    const float t1 = (b < 0.f) ? ra : rb, t2 = (b < 0.f) ? rb : ra;
    if (t1 >= minD && t2 >= minD) return js_min(t1, t2);
    return (t2 < minD) ? t1 : t2;
}

// The same quadratic solved like the reference does, in f64 (src/geometry.js:429-442).  Used once per *shaded*
// sphere / cylinder hit to recover the reference's hit distance to f64 accuracy before the hit point is formed:
// the point's distance from the surface (the `c` of the next ray's quadratic) is what decides self-hits at
// grazing angles, and a point built from an f32 distance sits up to 1e-6 further off the surface.
JSRT_DEV double sphere_intersect64(float3 o, float3 d, double md) {
    const double a = ddot3(d.x, d.y, d.z, d.x, d.y, d.z), b = ddot3(d.x, d.y, d.z, o.x, o.y, o.z), c = dsub(ddot3(o.x, o.y, o.z, o.x, o.y, o.z), 1.0);
    double big = dsub(dmul(b, b), dmul(a, c));
    if (big < 0.0 || a == 0.0) return -CUDART_INF;
    big = sqrt(big);
    const double t1 = dsub(-b, -big) / a, t2 = dsub(-b, big) / a;      // (-b + big) / a, (-b - big) / a
    if (t1 >= md && t2 >= md) return jsd_min(t1, t2);
    return (t2 < md) ? t1 : t2;
}

// f64 distance -> (f32 value, f32 remainder): hit records carry t + t_lo so that shading recomputes the
// reference's hit point (origin.plus(direction.times(t)) with an f64 t) instead of one rounded through f32
JSRT_DEV float split_t(double t, float* t_lo) {
    const float tf = (float)t;
    if (t_lo) *t_lo = isfinite(t) ? (float)(t - (double)tf) : 0.f;
    return tf;
}

// Triangle.toBarycentric src/geometry.js:389-396, in the reference's arithmetic: v2 is an f32 vector, the
// dot products and the Cramer solve are f64 (see the note on struct Tri), the result is stored f32.
JSRT_DEV float3 triangle_bary(const Tri* __restrict__ tris, int idx, float3 P) {
    const float4* tp = reinterpret_cast<const float4*>(tris + idx);
    const float4 b = __ldg(tp + 1), c = __ldg(tp + 4), e = __ldg(tp + 5);
    const double2 g = __ldg(reinterpret_cast<const double2*>(tp + 6)), h = __ldg(reinterpret_cast<const double2*>(tp + 7));
    const float v2x = (float)dsub(P.x, b.x), v2y = (float)dsub(P.y, b.y), v2z = (float)dsub(P.z, b.z);
    const double d20 = ddot3(v2x, v2y, v2z, c.x, c.y, c.z), d21 = ddot3(v2x, v2y, v2z, e.x, e.y, e.z);
    const double v = dmul(dsub(dmul(g.y, d20), dmul(h.x, d21)), h.y), w = dmul(dsub(dmul(g.x, d21), dmul(h.x, d20)), h.y);
    return f3((float)dsub(dsub(1.0, v), w), (float)v, (float)w);
}
// Triangle.intersect src/geometry.js:368-375 with the constructor constants of :341-353.
// `accept_lo` / `accept_hi` are the caller's acceptance window (t > lo && t <= hi; the caller applies
// its own strict tests and the tie rule); the reference evaluates the barycentrics regardless and the
// caller filters, which gives the same result as skipping them for a t that will be rejected anyway.
JSRT_DEV float triangle_intersect(const Tri* __restrict__ tris, int idx, float3 o, float3 d, float accept_lo, float accept_hi) {
    const float4* tp = reinterpret_cast<const float4*>(tris + idx);
    const float4 a = __ldg(tp);
    const float den = a.x * d.x + a.y * d.y + a.z * d.z;
    // (delta - n.o) / den with the SFU reciprocal (2 ulp) instead of the IEEE sequence (13 instructions with a branch):
    // distances already carry FP32 noise of that size, near-ties are settled in f64 by tie_wave (1e-6 window)
    const float t = (den != 0.f) ? __fdividef(a.w - (a.x * o.x + a.y * o.y + a.z * o.z), den) : -CUDART_INF_F;
    if (!(t > accept_lo && t <= accept_hi) || t < 0.f || isinf(t)) return -CUDART_INF_F;
    const float3 P = ray_point(o, d, t);
    {
        // FP32 fast path: the reference tests the f32-rounded barycentrics against [0, 1]; for a well-conditioned
        // triangle the FP32 solve (two dot products with host-folded vectors, struct Tri) is within ~1e-6 of them, so
        // any point whose barycentrics clear 0 and 1 by 1e-4 is decided here (hit or miss); everything else (edges,
        // slivers via the NaNs in ax / bx) falls through to the reference's arithmetic below.
        const float4 b = __ldg(tp + 1), A = __ldg(tp + 2), B = __ldg(tp + 3);
        const float v2x = P.x - b.x, v2y = P.y - b.y, v2z = P.z - b.z;
        const float v = v2x * A.x + v2y * A.y + v2z * A.z, w = v2x * B.x + v2y * B.y + v2z * B.z, u = 1.f - v - w;
        const float m = 1e-4f;
        // (all three NaN for a sliver: both tests fail and the f64 path decides)
        const float lo3 = fminf(fminf(u, v), w), hi3 = fmaxf(fmaxf(u, v), w);
        if (lo3 > m && hi3 < 1.f - m) return t;
        if (lo3 < -m || hi3 > 1.f + m) return -CUDART_INF_F;
    }
    const float3 bary = triangle_bary(tris, idx, P);
    return (bary.x >= 0.f && bary.x <= 1.f && bary.y >= 0.f && bary.y <= 1.f && bary.z >= 0.f && bary.z <= 1.f) ? t : -CUDART_INF_F;
}

// ---------------------------------------------------------------------------------
// SDF bytecode interpreter: register stacks of points / scales / distances.
// The program is straight-line (sdf_compile.cpp unrolls every loop), so all lanes
// of a warp run the same instruction stream; only the REFL fold is predicated.
//
// Arithmetic is the reference's, bit for bit: points are f32 vectors, every scalar
// (distance, scale, dot product, matrix entry) is f64, and each Vec-returning step
// rounds to f32.  Sphere tracing stops on `distance <= epsilon` (src/sdf.js:32) and
// the normal is a forward difference over a 1e-3 step (src/sdf.js:42-46); both turn a
// one-ulp difference into a visibly different pixel on fractal SDFs, so plain FP32
// evaluation cannot meet the parity bar here (measured: 47.8 dB on SDF_Menger).
// What keeps the f64 work small without changing a bit:
//  * a single +, -, * of two f32 values computed in f64 and rounded to f32 equals the FP32
//    operation (53 >= 2*24 + 2 bits: double rounding is innocuous), so those steps are FP32;
//  * the square of an f32 value, and the product of an f32 value with a matrix entry that is
//    itself representable in f32, are exact in f64, so mul + add chains over them are DFMAs;
//  * the top of every stack lives in registers; leaf + min/max and scale + min are single
//    instructions (sdf_compile.cpp fuses them), scale groups that cannot round are elided.
// Mat.times(Vec) with an f64 3x4 matrix: result[r] = b.dot(row r) stored f32 (src/math.js:392-397)
JSRT_DEV float3 xf64_apply(const double* __restrict__ m, float3 p, double w) {
    return f3((float)ddot4(p.x, p.y, p.z, w, __ldg(m + 0), __ldg(m + 1), __ldg(m + 2), __ldg(m + 3)),
              (float)ddot4(p.x, p.y, p.z, w, __ldg(m + 4), __ldg(m + 5), __ldg(m + 6), __ldg(m + 7)),
              (float)ddot4(p.x, p.y, p.z, w, __ldg(m + 8), __ldg(m + 9), __ldg(m + 10), __ldg(m + 11)));
}
// The same for a point (w = 1) and a matrix whose twelve entries are all representable in f32: every product
// is exact, so ((x m0 + y m1) + z m2) + m3 rounds exactly like the fused chain below.
JSRT_DEV float3 xf64_apply_exact(const double* __restrict__ m, float3 p) {
    const double2* r = reinterpret_cast<const double2*>(m);
    const double2 a0 = __ldg(r), a1 = __ldg(r + 1), b0 = __ldg(r + 2), b1 = __ldg(r + 3), c0 = __ldg(r + 4), c1 = __ldg(r + 5);
    const double x = p.x, y = p.y, z = p.z;
    return f3((float)dadd(fma(z, a1.x, fma(y, a0.y, dmul(x, a0.x))), a1.y),
              (float)dadd(fma(z, b1.x, fma(y, b0.y, dmul(x, b0.x))), b1.y),
              (float)dadd(fma(z, c1.x, fma(y, c0.y, dmul(x, c0.x))), c1.y));
}
JSRT_DEV double sdf_smooth_min(double a, double b, double k) {   // src/sdf.js:128-131
    const double h = jsd_max(dsub(k, fabs(dsub(a, b))), 0.0) / k;
    return dsub(jsd_min(a, b), dmul(dmul(dmul(dmul(h, h), h), k), (1.0 / 6.0)));
}
// Number(x.toPrecision(8)) (Math.fmod, src/math.js:27): round to 8 significant decimal digits.
// N = round(|x| * 10^k) with k chosen so that N has 8 digits, result = N / 10^k (or N * 10^-k for
// |x| >= 1e8); 10^|k| is exact in f64 for |k| <= 22 and the final division is correctly rounded, so this
// equals the decimal round trip except when |x| * 10^k lands within an ulp of a half-way point.
// Exact half-way cases go UP in magnitude ("pick the larger n", ECMA-262 21.1.3.5; printf / rint would go to even):
// they are common here, because a coordinate that came from an f32 has few significant bits (4.05078125 ->
// 4.0507813).  js_round_half_up also looks at the rounding error of the product, so a product that merely ROUNDED to
// n + 0.5 from below is not bumped.
__device__ const double kPow10[32] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                      1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1e23, 1e24, 1e25, 1e26, 1e27, 1e28, 1e29, 1e30, 1e31};
// correctly rounded reciprocals of the exact powers (the compiler folds these IEEE divisions)
__device__ const double kInvPow10[23] = {1.0 / 1e0, 1.0 / 1e1, 1.0 / 1e2, 1.0 / 1e3, 1.0 / 1e4, 1.0 / 1e5, 1.0 / 1e6, 1.0 / 1e7, 1.0 / 1e8, 1.0 / 1e9,
                                         1.0 / 1e10, 1.0 / 1e11, 1.0 / 1e12, 1.0 / 1e13, 1.0 / 1e14, 1.0 / 1e15, 1.0 / 1e16, 1.0 / 1e17, 1.0 / 1e18,
                                         1.0 / 1e19, 1.0 / 1e20, 1.0 / 1e21, 1.0 / 1e22};
JSRT_DEV double js_round_half_up(double y, double ax, double p10) {   // y = rn(ax * p10), 0 < y < 2^27
    double nn = floor(dadd(y, 0.5));
    if (dsub(nn, y) == 0.5 && fma(ax, p10, -y) < 0.0) nn = dsub(nn, 1.0);
    return nn;
}
// (out of line: never taken by coordinates an SDF scene produces, and its exp10 / log10 would sit in the interpreter's hot code)
__device__ __noinline__ double js_to_precision8_slow(double x) {
    const double ax = fabs(x);
    int e = (int)floor((double)(ilogb(ax)) * 0.30102999566398120);      // floor(log10(ax)) or one less
    int k = 7 - e;
    if (k > 30 || k < -30) {                                             // far outside the table: generic path
        e = (int)floor(log10(ax));
        if (ax >= exp10((double)(e + 1))) ++e; else if (ax < exp10((double)e)) --e;
        k = 7 - e;
        const double r = (k >= 0) ? rint(dmul(ax, exp10((double)k))) / exp10((double)k) : dmul(rint(ax / exp10((double)-k)), exp10((double)-k));
        return copysign(r, x);
    }
    double y = (k >= 0) ? dmul(ax, kPow10[k]) : ax / kPow10[-k];
    if (y >= 1e8) { --k; y = (k >= 0) ? dmul(ax, kPow10[k]) : ax / kPow10[-k]; }
    const double nn = (k >= 0) ? js_round_half_up(y, ax, kPow10[k]) : floor(dadd(y, 0.5));
    const double r = (k >= 0) ? nn / kPow10[k] : dmul(nn, kPow10[-k]);
    return copysign(r, x);
}
JSRT_DEV double js_to_precision8(double x) {
    if (x == 0.0 || !isfinite(x)) return x;
    // fast path, 1e-15 <= |x| < 1e8 (every coordinate an SDF scene produces): binary exponent from the bits,
    // floor(e2 log10 2) in integers (exact for |e2| <= 64), and N / 10^k as a Markstein division with the
    // tabulated reciprocal: q0 = N y, r = N - q0 10^k (exact, fused), q = q0 + r y — the correctly rounded
    // quotient, which is what the IEEE division of the general path returns.
    const int e2 = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
    if (e2 < -46 || e2 > 26) return js_to_precision8_slow(x);                // keeps k within the exact powers 10^0 .. 10^21
    const double ax = fabs(x);
    int k = 7 - ((e2 * 1233) >> 12);                                     // 7 - floor(log10(2^e2)); one too large when the mantissa crosses a decade
    double y = dmul(ax, __ldg(kPow10 + k));
    if (y >= 1e8) { --k; y = dmul(ax, __ldg(kPow10 + k)); }
    if (k < 0) return js_to_precision8_slow(x);
    const double p10 = __ldg(kPow10 + k), inv = __ldg(kInvPow10 + k);
    const double nn = js_round_half_up(y, ax, p10);
    const double q0 = dmul(nn, inv);
    const double q = fma(fma(-q0, p10, nn), inv, q0);
    return copysign(q, x);
}
JSRT_DEV double js_fmod(double a, double b) {
    // a / b: for a power-of-two period the product with the (exact) reciprocal is the same number
    int e; const bool pow2 = frexp(b, &e) == 0.5;
    const double q = pow2 ? dmul(a, 1.0 / b) : a / b;
    return js_to_precision8(dsub(a, dmul(floor(q), b)));
}
// the same with the exact reciprocal of a power-of-two period supplied by the compiler (sdf_compile.cpp)
JSRT_DEV double js_fmod_pow2(double a, double b, double inv_b) { return js_to_precision8(dsub(a, dmul(floor(dmul(a, inv_b)), b))); }

// min over BoxSDF(Inf, a, a), BoxSDF(a, Inf, a), BoxSDF(a, a, Inf) (src/sdf.js:276-279 under :83-88): the component of q
// on a bar's infinite axis is -Inf, so max(q, 0) is 0 there and the bar's distance is the 2-D box distance of the other
// two — exactly the S_BOX arithmetic with that component dropped (adding 0 * 0 to the sum of squares does not round) —
// and all three share q = |p| - a.
JSRT_DEV double sdf_cross(float3 p, float a) {
    const float qx = __fsub_rn(fabsf(p.x), a), qy = __fsub_rn(fabsf(p.y), a), qz = __fsub_rn(fabsf(p.z), a);
    const float mx = fmaxf(qx, 0.f), my = fmaxf(qy, 0.f), mz = fmaxf(qz, 0.f);
    const int npos = (mx > 0.f) + (my > 0.f) + (mz > 0.f);
    double lx, ly, lz;                    // |max(q, 0)| of the bar along x, y, z
    if (npos <= 1) { lx = (double)(my + mz); ly = (double)(mx + mz); lz = (double)(mx + my); }      // at most one term: exact
    else {
        const double x2 = dmul((double)mx, (double)mx), y2 = dmul((double)my, (double)my);
        lx = (my > 0.f && mz > 0.f) ? sqrt(fma((double)mz, (double)mz, y2)) : (double)(my + mz);
        ly = (mx > 0.f && mz > 0.f) ? sqrt(fma((double)mz, (double)mz, x2)) : (double)(mx + mz);
        lz = (mx > 0.f && my > 0.f) ? sqrt(fma((double)my, (double)my, x2)) : (double)(mx + my);
    }
    const double bx = dadd(lx, (double)fminf(fmaxf(qy, qz), 0.f)), by = dadd(ly, (double)fminf(fmaxf(qx, qz), 0.f)),
                 bz = dadd(lz, (double)fminf(fmaxf(qx, qy), 0.f));
    // (NaN like the three boxes: a NaN coordinate, or an infinite one — |p| - Inf on that bar's own axis)
    return (qx != qx || qy != qy || qz != qz || isinf(qx) || isinf(qy) || isinf(qz)) ? CUDART_NAN : jsd_min(jsd_min(bx, by), bz);
}
// S_RTU_CROSS: the loop of RecursiveTransformUnionSDF.distance (src/sdf.js:349-357) for the step { matrix transformer (:433-435);
// infinite repetition with one power-of-two period (:471-473); cross; bestDist = min(d * s, bestDist) } — the Menger recursion.
// The case costs the interpreter registers (with it compiled in, SDF_Sierpinski fell from 1 090 to 700-840 Mrays/s, inline or
// as a call: profiles/r2_ab.md §4), so the marching kernel exists in two builds and only scenes whose programs contain the
// instruction run the one with it (sdf_eval<true>; render.cu: has_rtu).
JSRT_DEV void sdf_rtu_cross(const double* __restrict__ m, bool exact, double scale, float period, float a, int n, float3& p, double& s, double& dtop) {
    const double sx = period, hh = sx / 2, inv = 1.0 / sx;                                                // (a power of two: exact)
    #pragma unroll 1
    for (int it = 0; it < n; ++it) {
        p = exact ? xf64_apply_exact(m, p) : xf64_apply(m, p, 1.0);
        s = dmul(s, scale);
        p = f3((float)dsub(js_fmod_pow2(dadd(p.x, hh), sx, inv), hh), (float)dsub(js_fmod_pow2(dadd(p.y, hh), sx, inv), hh),
               (float)dsub(js_fmod_pow2(dadd(p.z, hh), sx, inv), hh));
        dtop = jsd_min(dmul(sdf_cross(p, a), s), dtop);
    }
}
// the same out of line, for the callers that must accept every program but almost never meet this instruction (shading's
// material programs and fallback normals, SDF primitives inside aggregates)
__device__ __noinline__ void sdf_rtu_cross_call(const double* __restrict__ m, bool exact, double scale, float period, float a, int n,
                                                float3* p_io, double* s_io, double* d_io) {
    float3 p = *p_io; double s = *s_io, dtop = *d_io;
    sdf_rtu_cross(m, exact, scale, period, a, n, p, s, dtop);
    *p_io = p; *s_io = s; *d_io = dtop;
}
// RTU: 1 = S_RTU_CROSS inline (the marching kernel of scenes that use it), 0 = programs are known not to contain it,
// -1 = out-of-line call (everything else)
template <int RTU = -1>
JSRT_DEV double sdf_eval(const SdfInstr* __restrict__ code, const Xform64* __restrict__ xforms64, float3 p0) {
    // the tops of the three stacks are registers (p, s, dtop); the arrays hold what lies below them
    float3 P[8]; double S[12]; double D[8];
    int sp = 0, ss = 0, dp = 0;           // dp: distances on the stack (dtop is element dp - 1)
    float3 p = p0; double s = 1.0, dtop = 0.0;
    int4 i0 = __ldg(reinterpret_cast<const int4*>(code));
    float4 fv = __ldg(reinterpret_cast<const float4*>(code) + 1);
    for (int pc = 0;; ++pc) {
        const int op = i0.x, idx = i0.y;
        const double a0 = __hiloint2double(i0.w, i0.z);
        const float4 f = fv;
        // the next instruction is fetched while this one executes (straight-line code: pc + 1 always exists before S_END)
        if (op != S_END) { i0 = __ldg(reinterpret_cast<const int4*>(code + pc + 1)); fv = __ldg(reinterpret_cast<const float4*>(code + pc + 1) + 1); }
        if (op <= S_TETRA || op == S_CROSS) {
            if (op == S_END) return dtop;
            double v;
            if (op == S_CROSS) {
                v = sdf_cross(p, f.x);
            } else if (op == S_BOX) {                                                                            // src/sdf.js:276-279
                // q = |p| - size: one f64 subtraction of f32 values stored f32 = the FP32 subtraction
                const float qx = __fsub_rn(fabsf(p.x), f.x), qy = __fsub_rn(fabsf(p.y), f.y), qz = __fsub_rn(fabsf(p.z), f.z);
                const float mx = fmaxf(qx, 0.f), my = fmaxf(qy, 0.f), mz = fmaxf(qz, 0.f);                    // Vec.max(q, 0): q.w = 0
                // |max(q, 0)|: with at most one positive component the f64 square of an f32 value is exact and
                // its correctly rounded root is that value again, so the sqrt can be skipped without changing a bit;
                // otherwise the squares are exact in f64 and the sum rounds like the fused chain
                double len;
                if (my == 0.f && mz == 0.f) len = mx;
                else if (mx == 0.f && mz == 0.f) len = my;
                else if (mx == 0.f && my == 0.f) len = mz;
                else { const double x = mx, y = my, z = mz; len = sqrt(fma(z, z, fma(y, y, dmul(x, x)))); }
                const float inner = fminf(fmaxf(fmaxf(qx, qy), qz), 0.f);
                v = (qx != qx || qy != qy || qz != qz) ? CUDART_NAN : dadd(len, (double)inner);            // Math.max / min propagate NaN
            } else if (op == S_SPHERE) {                                                                  // src/sdf.js:232-234: p.to4(0).norm() - radius
                const double x = p.x, y = p.y, z = p.z;
                v = dsub(sqrt(fma(z, z, fma(y, y, dmul(x, x)))), a0);
            } else {                                                                                      // S_TETRA, src/sdf.js:305-308
                const double a = dsub(fabs(dadd(p.x, p.y)), p.z), b = dadd(fabs(dsub(p.x, p.y)), p.z);
                v = dsub(jsd_max(a, b), 1.0) / sqrt(3.0);
            }
            // idx: fused Union / Intersection step with the distance below (FOLD_MIN / FOLD_MAX), else push
            if (idx == 1) dtop = jsd_min(dtop, v);
            else if (idx == 2) dtop = jsd_max(dtop, v);
            else { if (dp > 0) D[dp - 1] = dtop; dtop = v; ++dp; }
            continue;
        }
        switch (op) {
            case S_MIN: { --dp; dtop = jsd_min(D[dp - 1], dtop); break; }
            case S_MAX: { --dp; dtop = jsd_max(D[dp - 1], dtop); break; }
            case S_NEG: dtop = -dtop; break;
            case S_ADDC: dtop = dadd(dtop, a0); break;
            case S_SMIN: { --dp; dtop = sdf_smooth_min(D[dp - 1], dtop, a0); break; }
            case S_SMIN_NEGA: { --dp; dtop = -sdf_smooth_min(-D[dp - 1], dtop, a0); break; }
            case S_SMIN_NEGAB: { --dp; dtop = -sdf_smooth_min(-D[dp - 1], -dtop, a0); break; }
            case S_PUSHP: { P[sp++] = p; S[ss++] = s; s = 1.0; break; }
            case S_POPP: { p = P[--sp]; s = S[--ss]; break; }
            case S_SBEGIN: { S[ss++] = s; s = 1.0; break; }
            case S_SEND: { s = dmul(S[--ss], s); break; }
            case S_MULS: dtop = dmul(dtop, s); break;
            case S_MULS_MIN: { --dp; dtop = jsd_min(dmul(dtop, s), D[dp - 1]); break; }                     // src/sdf.js:353-354
            case S_XFORM: {                                                                               // src/sdf.js:433-435
                p = (f.x != 0.f) ? xf64_apply_exact(xforms64[idx].m, p) : xf64_apply(xforms64[idx].m, p, 1.0);
                s = dmul(s, a0);
                break;
            }
            case S_RTU_CROSS: {
                if (RTU == 1) sdf_rtu_cross(xforms64[idx].m, f.w != 0.f, a0, f.x, f.y, (int)f.z, p, s, dtop);
                else if (RTU == -1) sdf_rtu_cross_call(xforms64[idx].m, f.w != 0.f, a0, f.x, f.y, (int)f.z, &p, &s, &dtop);
                else return CUDART_NAN;
                break;
            }
            case S_REFL: {                                                                                // src/sdf.js:450-455
                const double dt = dsub(ddot4(f.x, f.y, f.z, 0.0, p.x, p.y, p.z, 1.0), a0);
                if (dt < 0.0) {
                    const double k = dmul(2.0, dt);
                    const float nx = (float)dmul(f.x, k), ny = (float)dmul(f.y, k), nz = (float)dmul(f.z, k);
                    p = f3(__fsub_rn(p.x, nx), __fsub_rn(p.y, ny), __fsub_rn(p.z, nz));
                }
                break;
            }
            case S_REP: {                                                                                 // src/sdf.js:471-473
                const double sx = f.x, sy = f.y, sz = f.z;
                if (idx == 1) {                                                                           // one power-of-two period, a0 = 1 / period
                    const double h = sx / 2;
                    p = f3((float)dsub(js_fmod_pow2(dadd(p.x, h), sx, a0), h), (float)dsub(js_fmod_pow2(dadd(p.y, h), sx, a0), h),
                           (float)dsub(js_fmod_pow2(dadd(p.z, h), sx, a0), h));
                    break;
                }
                p = f3((float)dsub(js_fmod(dadd(p.x, sx / 2), sx), sx / 2), (float)dsub(js_fmod(dadd(p.y, sy / 2), sy), sy / 2),
                       (float)dsub(js_fmod(dadd(p.z, sz / 2), sz), sz / 2));
                break;
            }
            default: return CUDART_NAN;
        }
    }
}

// AABB.get_intersects in the reference's arithmetic (f32 vectors, f64 scalars); the SDF
// march starts exactly at its t_min (src/sdf.js:13-20).
JSRT_DEV bool aabb_intersects_f64(float3 c, float3 h, float3 o, float3 d, double minD, double maxD, double& t_min, double& t_max) {
    t_min = -CUDART_INF; t_max = CUDART_INF;
    const float3 p = f3((float)dsub(c.x, o.x), (float)dsub(c.y, o.y), (float)dsub(c.z, o.z));
    const double eps = 0.0000001;
#define JSRT_SLAB64(PI, HI, DI)                                                 \
    if (fabs((double)DI) > eps) {                                               \
        double t1 = dadd(PI, HI) / (double)DI, t2 = dsub(PI, HI) / (double)DI;  \
        if (t1 > t2) { const double tmp = t1; t1 = t2; t2 = tmp; }              \
        if (t1 > t_min) t_min = t1;                                             \
        if (t2 < t_max) t_max = t2;                                             \
        if (t_min > t_max || t_max < minD || t_min > maxD) return false;        \
    } else if (fabs((double)PI) > (double)HI) return false;
    JSRT_SLAB64(p.x, h.x, d.x)
    JSRT_SLAB64(p.y, h.y, d.y)
    JSRT_SLAB64(p.z, h.z, d.z)
#undef JSRT_SLAB64
    return true;
}

JSRT_DEV float3 ray_point_f64(float3 o, float3 d, double t) {   // origin.plus(direction.times(t)) with an f64 t
    // (the sum of two f32 values rounded through f64 is the FP32 sum)
    return f3(__fadd_rn(o.x, (float)dmul(d.x, t)), __fadd_rn(o.y, (float)dmul(d.y, t)), __fadd_rn(o.z, (float)dmul(d.z, t)));
}

// SDFGeometry.intersect src/sdf.js:12-40.  Returns the hit distance as f64 (NaN-free:
// -inf for a miss).
JSRT_DEV double sdf_intersect(const SdfProgram& pr, const SdfInstr* __restrict__ code, const Xform64* __restrict__ xforms64,
                              float3 o, float3 d, double minD, double maxD, unsigned long long* evals) {
    double bt0, bt1;
    if (!aabb_intersects_f64(f3(pr.cx, pr.cy, pr.cz), f3(pr.hx, pr.hy, pr.hz), o, d, minD, maxD, bt0, bt1)) return -CUDART_INF;
    const double lo = jsd_max(minD, bt0), hi = jsd_min(maxD, bt1);
    double t = lo;
    const double rd_norm = sqrt(ddot4(d.x, d.y, d.z, 0.0, d.x, d.y, d.z, 0.0));
    const SdfInstr* prog = code + pr.first_instr;
    for (int i = 0; i < pr.max_samples; ++i) {
        const double dist = sdf_eval(prog, xforms64, ray_point_f64(o, d, t));
        if (evals) ++*evals;
        if (!isfinite(dist)) break;
        if (dist <= pr.distance_epsilon) return t;
        t = dadd(t, dist / rd_norm);
        if (t < lo || t > hi || dmul(dsub(t, lo), rd_norm) > pr.max_trace_distance) break;
    }
    return -CUDART_INF;
}

}  // namespace jsrt
