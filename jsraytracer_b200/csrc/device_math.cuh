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
JSRT_DEV float3 normalized3(float3 a) { const float n = sqrtf(dot3(a, a)); return (n > 0.00001f) ? a * (1.0f / n) : a; }
// Ray.getPoint: two f32 roundings (product, then sum)
JSRT_DEV float3 ray_point(float3 o, float3 d, float t) {
    return f3(__fadd_rn(o.x, __fmul_rn(d.x, t)), __fadd_rn(o.y, __fmul_rn(d.y, t)), __fadd_rn(o.z, __fmul_rn(d.z, t)));
}
JSRT_DEV float js_sign(float x) { return (x > 0.f) ? 1.f : (x < 0.f ? -1.f : x); }
// Math.min / Math.max propagate NaN
JSRT_DEV float js_min(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fminf(a, b); }
JSRT_DEV float js_max(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }

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

// Sphere.staticIntersect src/geometry.js:429-442
JSRT_DEV float sphere_intersect(float3 o, float3 d, float minD) {
    const float a = dot3(d, d), b = dot3(d, o), c = dot3(o, o) - 1.f;
    float big = b * b - a * c;
    if (big < 0.f || a == 0.f) return -CUDART_INF_F;
    big = sqrtf(big);
    const float t1 = (-b + big) / a, t2 = (-b - big) / a;
    if (t1 >= minD && t2 >= minD) return js_min(t1, t2);
    return (t2 < minD) ? t1 : t2;
}

// Triangle.intersect src/geometry.js:368-375 with the constructor constants of :341-353.
// `accept_lo` / `accept_hi` are the caller's acceptance window (t > lo && t < hi);
// the reference evaluates the barycentrics regardless and the caller filters, which
// gives the same result as skipping them for a t that will be rejected anyway.
JSRT_DEV float triangle_intersect(const Tri* __restrict__ tris, int idx, float3 o, float3 d, float accept_lo, float accept_hi) {
    const float4* tp = reinterpret_cast<const float4*>(tris + idx);
    const float4 a = __ldg(tp);
    const float den = a.x * d.x + a.y * d.y + a.z * d.z;
    const float t = (den != 0.f) ? (a.w - (a.x * o.x + a.y * o.y + a.z * o.z)) / den : -CUDART_INF_F;
    if (!(t > accept_lo && t < accept_hi) || t < 0.f || isinf(t)) return -CUDART_INF_F;
    const float4 b = __ldg(tp + 1), c = __ldg(tp + 2), e = __ldg(tp + 3);
    const float3 P = ray_point(o, d, t);
    const float3 v2 = f3(P.x - b.x, P.y - b.y, P.z - b.z);
    const float d20 = v2.x * c.x + v2.y * c.y + v2.z * c.z, d21 = v2.x * e.x + v2.y * e.y + v2.z * e.z;
    const float d00 = c.w, d11 = e.w, d01 = b.w;
    const float denom = d00 * d11 - d01 * d01;
    const float v = (d11 * d20 - d01 * d21) / denom, w = (d00 * d21 - d01 * d20) / denom;
    const float u = 1.f - v - w;
    return (u >= 0.f && u <= 1.f && v >= 0.f && v <= 1.f && w >= 0.f && w <= 1.f) ? t : -CUDART_INF_F;
}
// Triangle.toBarycentric src/geometry.js:389-396
JSRT_DEV float3 triangle_bary(const Tri* __restrict__ tris, int idx, float3 P) {
    const float4* tp = reinterpret_cast<const float4*>(tris + idx);
    const float4 b = __ldg(tp + 1), c = __ldg(tp + 2), e = __ldg(tp + 3);
    const float3 v2 = f3(P.x - b.x, P.y - b.y, P.z - b.z);
    const float d20 = v2.x * c.x + v2.y * c.y + v2.z * c.z, d21 = v2.x * e.x + v2.y * e.y + v2.z * e.z;
    const float d00 = c.w, d11 = e.w, d01 = b.w;
    const float denom = d00 * d11 - d01 * d01;
    const float v = (d11 * d20 - d01 * d21) / denom, w = (d00 * d21 - d01 * d20) / denom;
    return f3(1.f - v - w, v, w);
}

// ---------------------------------------------------------------------------------
// SDF bytecode interpreter: register stacks of points / scales / distances.
// The program is straight-line (sdf_compile.cpp unrolls every loop), so all lanes
// of a warp run the same instruction stream; only the REFL fold is predicated.
JSRT_DEV float sdf_smooth_min(float a, float b, float k) {   // src/sdf.js:128-131
    const float h = js_max(k - fabsf(a - b), 0.0f) / k;
    return js_min(a, b) - h * h * h * k * (1.0f / 6.0f);
}
JSRT_DEV float sdf_eval(const SdfInstr* __restrict__ code, const Xform* __restrict__ xforms, float3 p0) {
    float3 P[8]; float S[8]; float D[8];
    int sp = 0, dp = 0;
    P[0] = p0; S[0] = 1.f;
    for (int pc = 0;; ++pc) {
        const float4* ip = reinterpret_cast<const float4*>(code + pc);
        const float4 i0 = __ldg(ip);
        const int op = __float_as_int(i0.x);
        const float a0 = i0.y, a1 = i0.z, a2 = i0.w;
        switch (op) {
            case S_END: return D[0];
            case S_SPHERE: { const float3 p = P[sp]; D[dp++] = sqrtf(dot3(p, p)) - a0; break; }          // src/sdf.js:232-234
            case S_BOX: {                                                                                 // src/sdf.js:276-279
                const float3 p = P[sp];
                const float qx = fabsf(p.x) - a0, qy = fabsf(p.y) - a1, qz = fabsf(p.z) - a2;
                const float mx = fmaxf(qx, 0.f), my = fmaxf(qy, 0.f), mz = fmaxf(qz, 0.f);
                D[dp++] = sqrtf(mx * mx + my * my + mz * mz) + fminf(fmaxf(fmaxf(qx, qy), qz), 0.f);
                break;
            }
            case S_TETRA: { const float3 p = P[sp]; D[dp++] = (fmaxf(fabsf(p.x + p.y) - p.z, fabsf(p.x - p.y) + p.z) - 1.f) / 1.7320508075688772f; break; }   // src/sdf.js:305-308
            case S_MIN: { --dp; D[dp - 1] = js_min(D[dp - 1], D[dp]); break; }
            case S_MAX: { --dp; D[dp - 1] = js_max(D[dp - 1], D[dp]); break; }
            case S_NEG: D[dp - 1] = -D[dp - 1]; break;
            case S_ADDC: D[dp - 1] += a0; break;
            case S_SMIN: { --dp; D[dp - 1] = sdf_smooth_min(D[dp - 1], D[dp], a0); break; }
            case S_SMIN_NEGA: { --dp; D[dp - 1] = -sdf_smooth_min(-D[dp - 1], D[dp], a0); break; }
            case S_SMIN_NEGAB: { --dp; D[dp - 1] = -sdf_smooth_min(-D[dp - 1], -D[dp], a0); break; }
            case S_PUSHP: { P[sp + 1] = P[sp]; S[sp + 1] = 1.f; ++sp; break; }
            case S_POPP: --sp; break;
            case S_MULS: D[dp - 1] *= S[sp]; break;
            case S_XFORM: { const XformReg m = load_xform(xforms, __float_as_int(a0)); P[sp] = xf_point(m, P[sp]); S[sp] *= a1; break; }   // src/sdf.js:433-435
            case S_REFL: {                                                                                // src/sdf.js:450-455
                const float a3 = __ldg(reinterpret_cast<const float*>(ip + 1));
                const float3 p = P[sp];
                const float dt = a0 * p.x + a1 * p.y + a2 * p.z - a3;
                if (dt < 0.f) { const float k = 2.f * dt; P[sp] = f3(p.x - a0 * k, p.y - a1 * k, p.z - a2 * k); }
                break;
            }
            case S_REP: {                                                                                 // src/sdf.js:471-473, Math.fmod src/math.js:27
                const float3 p = P[sp];
                const float ax = p.x + a0 * 0.5f, ay = p.y + a1 * 0.5f, az = p.z + a2 * 0.5f;
                P[sp] = f3(ax - floorf(ax / a0) * a0 - a0 * 0.5f, ay - floorf(ay / a1) * a1 - a1 * 0.5f, az - floorf(az / a2) * a2 - a2 * 0.5f);
                break;
            }
            default: return CUDART_NAN_F;
        }
    }
}

// SDFGeometry.intersect src/sdf.js:12-40.  t advances in f64 like the reference's
// scalar, the marched point is f32 like its Vec.
JSRT_DEV float sdf_intersect(const SdfProgram& pr, const SdfInstr* __restrict__ code, const Xform* __restrict__ xforms,
                             float3 o, float3 d, float minD, float maxD, unsigned long long* evals) {
    float bt0, bt1;
    if (!aabb_intersects(f3(pr.cx, pr.cy, pr.cz), f3(pr.hx, pr.hy, pr.hz), o, d, minD, maxD, bt0, bt1)) return -CUDART_INF_F;
    const double lo = fmax((double)minD, (double)bt0), hi = fmin((double)maxD, (double)bt1);
    double t = lo;
    const double rd_norm = sqrt((double)d.x * d.x + (double)d.y * d.y + (double)d.z * d.z);
    const SdfInstr* prog = code + pr.first_instr;
    for (int i = 0; i < pr.max_samples; ++i) {
        const float3 p = f3((float)((double)o.x + (double)(float)((double)d.x * t)), (float)((double)o.y + (double)(float)((double)d.y * t)),
                            (float)((double)o.z + (double)(float)((double)d.z * t)));
        const float dist = sdf_eval(prog, xforms, p);
        if (evals) ++*evals;
        if (!isfinite(dist)) break;
        if (dist <= pr.distance_epsilon) return (float)t;
        t += (double)dist / rd_norm;
        if (t < lo || t > hi || (t - lo) * rd_norm > (double)pr.max_trace_distance) break;
    }
    return -CUDART_INF_F;
}

}  // namespace jsrt
