// Shading on the device: Primitive.color (src/world.js:125-137), geometry
// materialData (src/geometry.js, src/sdf.js:41-47), MaterialColor evaluation
// (src/materials.js:27-76), PhongMaterial / FresnelPhongMaterial /
// PhongPathTracingMaterial (src/materials.js:195-476) and the light samplers
// (src/lights.js:45-53,80-93).
//
// The reference evaluates the bounce tree depth-first and multiplies child
// colours on the way back up; a wavefront instead carries the product of the
// weights on the way down (`throughput`) and adds every term straight into the
// pixel: ambient immediately, each light sample when its shadow ray comes back
// unoccluded, children as new rays.  Same sum, different association.
#pragma once
#include "rng.h"
#include "trace.cuh"

namespace jsrt {

struct SurfaceData {           // the `data` object of materials.js
    float3 position;           // world
    float3 normal;             // world, normalised (src/world.js:134)
    float2 uv; bool has_uv;
    float3 basecolor;          // SDF only; (1,1,1) otherwise
};

// TextureMaterialColor.color (src/materials.js:101-130) in the reference's arithmetic: f64 weights and sums, the
// result stored as an f32 RGBA vector, then the folded ScaledMaterialColor factor.  `alpha` (optional) receives
// the fourth component, which only matters where the reference takes squarednorm() of a colour (:277,284).
JSRT_DEV double tex_normalize_uv(double c, bool clamp) {
    if (clamp) return (c != c) ? c : fmin(fmax(c, 0.0), 1.0);
    return fmod(fmod(c, 1.0) + 1.0, 1.0);
}
// (noinline, scalar arguments: the rare texture path must not cost the common path registers or a stack frame)
__device__ __noinline__ float4 texture_eval(const Texture* __restrict__ textures, const unsigned char* __restrict__ texels, int tex_index,
                                            float u, float v, float sr, float sg, float sb, float sa) {
    const Texture t = textures[tex_index];
    const double U = tex_normalize_uv((double)u, t.flags & TF_CLAMP_U), V = tex_normalize_uv((double)v, t.flags & TF_CLAMP_V);
    const double fx = dsub(dmul(U, (double)t.width), 0.5), fy = dsub(dmul(dsub(1.0, V), (double)t.height), 0.5);
    double xs[2], wx[2], ys[2], wy[2]; int n = 2;
    if (t.flags & TF_NEAREST) { xs[0] = floor(fx + 0.5); wx[0] = 1.0; ys[0] = floor(fy + 0.5); wy[0] = 1.0; xs[1] = xs[0]; ys[1] = ys[0]; wx[1] = wy[1] = 0.0; n = 1; }
    else {
        xs[0] = floor(fx); xs[1] = xs[0] + 1.0; wx[1] = fmod(fx, 1.0); wx[0] = dsub(1.0, wx[1]);
        ys[0] = floor(fy); ys[1] = ys[0] + 1.0; wy[1] = fmod(fy, 1.0); wy[0] = dsub(1.0, wy[1]);
    }
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    const uchar4* tex = reinterpret_cast<const uchar4*>(texels + t.offset);
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            const double cy = fmin(fmax(ys[b], 0.0), (double)(t.height - 1)), cx = fmin(fmax(xs[a], 0.0), (double)(t.width - 1));
            const long long index = (long long)(cy * (double)t.width + cx);
            const uchar4 p = (cx == cx && cy == cy) ? tex[index] : make_uchar4(0, 0, 0, 0);
            const double w = dmul(wx[a], wy[b]);
            r[0] = dadd(r[0], dmul(w, (double)p.x / 255.0)); r[1] = dadd(r[1], dmul(w, (double)p.y / 255.0));
            r[2] = dadd(r[2], dmul(w, (double)p.z / 255.0)); r[3] = dadd(r[3], dmul(w, (double)p.w / 255.0));
        }
    return make_float4((float)((double)(float)r[0] * (double)sr), (float)((double)(float)r[1] * (double)sg), (float)((double)(float)r[2] * (double)sb),
                       (float)((double)(float)r[3] * (double)sa));
}

// LEAN (here and below): builds of shade_kernel with the code of features the scene does not use compiled out (render.cu:
// leanLevel).  1: no textures, positional UVs, solid or transparent materials (most scenes) — the texture path is a call,
// and a call site costs the whole kernel registers.  2: level 1 and no area lights, round or box primitives, with planes as
// the only analytic primitives — the ground-plane + mesh + point-light scenes.
template <int LEAN = 0>
JSRT_DEV float3 color_eval(const DeviceScene& sc, const Color& c, const SurfaceData& s, float* alpha = nullptr) {
    if (alpha) *alpha = 0.f;
    if (c.checker == CK_SOLID) return f3(c.c1[0], c.c1[1], c.c1[2]);
    if (LEAN < 1 && c.checker == CK_TEXTURE) {
        const float4 t = texture_eval(sc.textures, sc.texels, c.tex, s.has_uv ? s.uv.x : 0.f, s.has_uv ? s.uv.y : 0.f, c.c1[0], c.c1[1], c.c1[2], c.c2[0]);
        if (alpha) *alpha = t.w;
        return f3(t.x, t.y, t.z);
    }
    // CheckerboardMaterialColor.color src/materials.js:72-75 (f64: UV can be huge towards the horizon)
    const double u = s.has_uv ? (double)s.uv.x : 0.0, v = s.has_uv ? (double)s.uv.y : 0.0;
    const double a = floor(u) + floor(v);
    // Math.fmod(a, 2) % 2 for an integer-valued a: a - floor(a / 2) * 2 is exactly 0 or 1 (NaN for an infinite a), which
    // neither toPrecision(8) nor the second `% 2` changes
    const double m = a - floor(a * 0.5) * 2.0;
    return (m < 1.0) ? f3(c.c1[0], c.c1[1], c.c1[2]) : f3(c.c2[0], c.c2[1], c.c2[2]);
}

// Vec.cartesianToSpherical src/math.js:189-193
JSRT_DEV float2 cartesian_to_spherical(float3 n) {
    return make_float2(0.5f + atan2f(n.z, n.x) / (2.f * CUDART_PI_F), 0.5f - asinf(n.y) / CUDART_PI_F);
}
// Vec.spherePick src/math.js:180-188: theta = 2 pi u0, phi = acos(2 u1 - 1), (cos theta sin phi, cos phi, sin theta sin phi).
// cos(acos x) = x and sin(acos x) = sqrt(1 - x^2), and sin / cos of 2 pi u0 come from the pi-scaled routine (exact argument
// reduction, no slow path): the same point to FP32 rounding, a fifth of the instructions of acosf + sincosf + sinf + cosf
// (shade_kernel lost 40 % of its stall samples to instruction fetch: profiles/r2_ab.md).
JSRT_DEV float3 sphere_pick(float u0, float u1) {
    const float c = 2.0f * u1 - 1.0f, sin_phi = sqrtf(fmaxf(0.f, 1.0f - c * c));
    float st, ct; sincospif(2.0f * u0, &st, &ct);
    return f3(ct * sin_phi, c, st * sin_phi);
}
// Math.pow(base, exponent) for base >= 0 (a clamped cosine) and a finite exponent >= 0 (smoothness): exp2(e log2 b) on the
// special-function unit.  Relative error ~ e * 2^-22 (1e-5 at smoothness 100), far below the image tolerance; the general
// powf is ~150 instructions with several slow paths.  Math.pow(x, 0) = 1 for every x including 0 (src/materials.js:390
// defaults smoothness to 0), Math.pow(0, e > 0) = 0.
#ifndef JSRT_SHADE_OUTLINE
#define JSRT_SHADE_OUTLINE 0      // 1: rare paths of shade_kernel (library powf, sphere / cylinder UVs, the f64 re-solve of round hits) out of
                                  // line.  Measured 0.3-1.4 % SLOWER (bunny_path 7 275 -> 7 257, cornell_box_path 10 400 -> 10 344, refraction_path
                                  // 15 547 -> 15 329: profiles/r2/ab_r2p_*): ptxas already lays those blocks out behind the hot path, and the
                                  // call sequence costs registers; kept as a switch for A/B runs
#endif
#if JSRT_SHADE_OUTLINE
#define JSRT_RARE __device__ __noinline__
#else
#define JSRT_RARE JSRT_DEV
#endif
JSRT_RARE float pow_library(float b, float e) { return powf(b, e); }
JSRT_DEV float pow_clamped(float b, float e) {
    if (e == 0.f) return 1.f;
    if (!(b > 0.f)) return (b == 0.f) ? 0.f : pow_library(b, e);   // NaN / negative: the library's answer
    if (!(e < 3.0e38f)) return pow_library(b, e);                    // infinite smoothness
    return exp2f(e * __log2f(b));
}
// Sphere / Cylinder.materialData (src/geometry.js:449-455,479-487): normal + spherical / cylindrical UV (atan2, asin)
JSRT_RARE void round_material_data(int geom_kind, float lx, float ly, float lz, float* out /* n.xyz, u, v */) {
    const float3 lp = f3(lx, ly, lz);
    float3 n; float2 uv;
    if (geom_kind == G_SPHERE) {                           // position is a 4-vector with w = 1 (sic)
        const float nn = sqrtf(lp.x * lp.x + lp.y * lp.y + lp.z * lp.z + 1.f);
        n = (nn > 0.00001f) ? lp * (1.f / nn) : lp;
        uv = cartesian_to_spherical(n);
    } else {
        n = normalized3(f3(lp.x, lp.y, 0.f));
        uv = make_float2(0.5f + atan2f(lp.y, lp.x) / (2.f * CUDART_PI_F), 0.5f + lp.z);
    }
    out[0] = n.x; out[1] = n.y; out[2] = n.z; out[3] = uv.x; out[4] = uv.y;
}

// SDF material program (sdf_compile.cpp): root_sdf.getMaterialData(p) -> basecolor, UV.
struct SdfMat { double d; float3 base; float2 uv; bool has_uv; };
JSRT_DEV double smooth_min_blend(double a, double b, double k) {     // src/sdf.js:133-137
    const double h = jsd_max(dsub(k, fabs(dsub(a, b))), 0.0) / k;
    const double m = dmul(dmul(dmul(h, h), h), 0.5);
    return (a < b) ? m : dsub(1.0, m);
}
JSRT_DEV SdfMat sdf_blend(double mix, const SdfMat& a, const SdfMat& b) {   // SDF.blendMaterialData src/sdf.js:66-73
    if (mix <= 0.0) return a;
    if (mix >= 1.0) return b;
    SdfMat r; r.d = 0; r.has_uv = true;
    const float2 ua = a.has_uv ? a.uv : make_float2(0.f, 0.f), ub = b.has_uv ? b.uv : make_float2(0.f, 0.f);
    const double w = dsub(1.0, mix);
    r.base = f3((float)dadd(dmul(w, a.base.x), dmul(mix, b.base.x)), (float)dadd(dmul(w, a.base.y), dmul(mix, b.base.y)), (float)dadd(dmul(w, a.base.z), dmul(mix, b.base.z)));
    r.uv = make_float2((float)dadd(dmul(w, ua.x), dmul(mix, ub.x)), (float)dadd(dmul(w, ua.y), dmul(mix, ub.y)));
    return r;
}
JSRT_DEV void sdf_material(const DeviceScene& sc, int first, float3 p, float3& base, float2& uv, bool& has_uv) {
    SdfMat M[8]; int sp = 0;
    for (int pc = first;; ++pc) {
        const int4 i0 = __ldg(reinterpret_cast<const int4*>(sc.sdf_code + pc));
        const float4 fv = __ldg(reinterpret_cast<const float4*>(sc.sdf_code + pc) + 1);
        const double a0 = __hiloint2double(i0.w, i0.z);
        switch (i0.x) {
            case MP_LEAF: {
                SdfMat m; m.d = 0; m.base = f3(fv.x, fv.y, fv.z); m.has_uv = false; m.uv = make_float2(0.f, 0.f);
                if (i0.y == 1) {          // SphereSDF.getMaterialData src/sdf.js:235-240
                    const double nn = sqrt(ddot4(p.x, p.y, p.z, 0.0, p.x, p.y, p.z, 0.0));
                    float3 n = p;
                    if (nn > 0.00001) n = f3((float)dmul(p.x, 1.0 / nn), (float)dmul(p.y, 1.0 / nn), (float)dmul(p.z, 1.0 / nn));
                    m.uv = make_float2((float)dadd(0.5, atan2((double)n.z, (double)n.x) / dmul(2.0, 3.141592653589793)),
                                       (float)dsub(0.5, asin((double)n.y) / 3.141592653589793));
                    m.has_uv = true;
                }
                M[sp++] = m; break;
            }
            case MP_ATTACH: M[sp - 1].d = sdf_eval(sc.sdf_code + i0.y, sc.xforms64, p); break;
            case MP_SELMIN: { --sp; if (M[sp].d < M[sp - 1].d) M[sp - 1] = M[sp]; break; }
            case MP_SELMAX: { --sp; if (M[sp].d > M[sp - 1].d) M[sp - 1] = M[sp]; break; }
            case MP_DIFF: { --sp; if (!(M[sp - 1].d > -M[sp].d)) M[sp - 1] = M[sp]; break; }
            case MP_BLEND_U: { --sp; M[sp - 1] = sdf_blend(smooth_min_blend(M[sp - 1].d, M[sp].d, a0), M[sp - 1], M[sp]); break; }
            case MP_BLEND_I: { --sp; M[sp - 1] = sdf_blend(dsub(1.0, smooth_min_blend(-M[sp - 1].d, -M[sp].d, a0)), M[sp - 1], M[sp]); break; }
            case MP_BLEND_D: { --sp; M[sp - 1] = sdf_blend(smooth_min_blend(-M[sp - 1].d, M[sp].d, a0), M[sp - 1], M[sp]); break; }
            default: base = M[0].base; uv = M[0].uv; has_uv = M[0].has_uv; return;    // MP_END
        }
    }
}

// geometry.materialData in the primitive's local space -> local normal, UV, basecolor.
template <bool HAS_SDF, int LEAN = 0>
JSRT_DEV void material_data(const DeviceScene& sc, int geom_kind, int geom_index, int flags, float3 lp,
                            float3& n, float2& uv, bool& has_uv, float3& base, const float4* sdf_normal = nullptr) {
    has_uv = false; uv = make_float2(0.f, 0.f); base = f3(1.f, 1.f, 1.f); n = f3(0.f, 0.f, 1.f);
    switch (geom_kind) {
        case G_PLANE: case G_SQUARE: case G_CIRCLE:       // src/geometry.js:249-254
            n = f3(0.f, 0.f, 1.f); uv = make_float2(lp.x, lp.y); has_uv = true; break;
        case G_BOX: if (LEAN < 2) {                                     // AABB.materialData src/geometry.js:210-224
            float3 c = f3(0.f, 0.f, 0.f), h = f3(0.5f, 0.5f, 0.5f);
            if (geom_index >= 0) { const float* b = sc.boxes + 8 * geom_index; c = f3(b[0], b[1], b[2]); h = f3(b[4], b[5], b[6]); }
            float norm_dist = 0.f; n = f3(0.f, 0.f, 0.f);
            const float cx = (lp.x - c.x) / h.x, cy = (lp.y - c.y) / h.y, cz = (lp.z - c.z) / h.z;
            if (fabsf(cx) > norm_dist) { norm_dist = fabsf(cx); n = f3(js_sign(cx), 0.f, 0.f); }
            if (fabsf(cy) > norm_dist) { norm_dist = fabsf(cy); n = f3(0.f, js_sign(cy), 0.f); }
            if (fabsf(cz) > norm_dist) { norm_dist = fabsf(cz); n = f3(0.f, 0.f, js_sign(cz)); }
            break;
        }
        case G_SPHERE: case G_CYLINDER: if (LEAN < 2) {                 // src/geometry.js:449-455,479-487
            float r[5];
            round_material_data(geom_kind, lp.x, lp.y, lp.z, r);
            n = f3(r[0], r[1], r[2]); uv = make_float2(r[3], r[4]); has_uv = true; break;
        }
        case G_TRIANGLE: {                                // src/geometry.js:376-385,397-409
            const float4 a = __ldg(reinterpret_cast<const float4*>(sc.tris + geom_index));
            n = f3(a.x, a.y, a.z);
            if (flags & (PF_HAS_VNORMALS | PF_HAS_UVS)) {
                const float3 bary = triangle_bary(sc.tris, geom_index, lp);
                const TriShade* ts = sc.tri_shade + geom_index;
                if (flags & PF_HAS_UVS) {
                    uv = make_float2(ts->uv[0][0] * bary.x + ts->uv[1][0] * bary.y + ts->uv[2][0] * bary.z,
                                     ts->uv[0][1] * bary.x + ts->uv[1][1] * bary.y + ts->uv[2][1] * bary.z);
                    has_uv = true;
                }
                if (flags & PF_HAS_VNORMALS)
                    n = f3(ts->n[0][0] * bary.x + ts->n[1][0] * bary.y + ts->n[2][0] * bary.z,
                           ts->n[0][1] * bary.x + ts->n[1][1] * bary.y + ts->n[2][1] * bary.z,
                           ts->n[0][2] * bary.x + ts->n[1][2] * bary.y + ts->n[2][2] * bary.z);
            }
            break;
        }
        case G_SDF: if (HAS_SDF) {                        // src/sdf.js:41-47: forward differences, reference arithmetic
            const SdfProgram& pr = sc.sdfs[geom_index];
            base = f3(pr.base[0], pr.base[1], pr.base[2]);
            if (pr.mat_first >= 0) sdf_material(sc, pr.mat_first, lp, base, uv, has_uv);
            if (sdf_normal && sdf_normal->w != 0.f) { n = f3(sdf_normal->x, sdf_normal->y, sdf_normal->z); break; }   // computed by sdf_kernel
            const SdfInstr* prog = sc.sdf_code + pr.first_instr;
            const double step = pr.normal_step_size;
            const float fs = (float)step;                 // Vec.axis(i, 4, step) stores the step as f32
            const double d0 = sdf_eval(prog, sc.xforms64, lp);
            const double dx = sdf_eval(prog, sc.xforms64, f3((float)dadd(lp.x, fs), lp.y, lp.z));
            const double dy = sdf_eval(prog, sc.xforms64, f3(lp.x, (float)dadd(lp.y, fs), lp.z));
            const double dz = sdf_eval(prog, sc.xforms64, f3(lp.x, lp.y, (float)dadd(lp.z, fs)));
            const float nx = (float)(dsub(dx, d0) / step), ny = (float)(dsub(dy, d0) / step), nz = (float)(dsub(dz, d0) / step);
            const double nn = sqrt(ddot4(nx, ny, nz, 0.0, nx, ny, nz, 0.0));
            n = (nn > 0.00001) ? f3((float)dmul(nx, 1.0 / nn), (float)dmul(ny, 1.0 / nn), (float)dmul(nz, 1.0 / nn)) : f3(nx, ny, nz);
            break;
        }
        default: break;
    }
}

struct LightSample { float3 direction; float3 color; };

// Light.sampleIterator for sample `u0,u1` (src/lights.js:45-53,80-93)
template <int LEAN = 0>
JSRT_DEV LightSample light_sample(const Light& l, float3 P, float u0, float u1) {
    LightSample s;
    const float3 lc = f3(l.color[0], l.color[1], l.color[2]);
    if (LEAN >= 2 || l.kind == L_POINT) {
        s.direction = f3(l.pos[0], l.pos[1], l.pos[2]) - P;
        s.color = lc * (1.f / (4.f * CUDART_PI_F * dot3(s.direction, s.direction)));     // Light.falloff :21-23
        return s;
    }
    float3 lp, ln;
    if (l.geom == G_SPHERE) {                              // Sphere.sampleSurface + Sphere.materialData (w = 1 quirk)
        lp = sphere_pick(u0, u1);
        const float nn = sqrtf(dot3(lp, lp) + 1.f);
        ln = lp * (1.f / nn);
    } else {                                               // Square / Circle.sampleSurface src/geometry.js:295-300,326-331
        lp = f3(u0 - 0.5f, u1 - 0.5f, 0.f); ln = f3(0.f, 0.f, 1.f);
    }
    const XformReg xf = load_xform(&l.xf, 0), inv = load_xform(&l.inv, 0);
    const float3 wp = xf_point(xf, lp);
    s.direction = wp - P;
    const float3 nl = normalized3(xf_normal(inv, ln));
    const float fall = 1.f / (4.f * CUDART_PI_F * dot3(s.direction, s.direction));
    s.color = lc * (fall * fabsf(dot3(normalized3(s.direction), nl)));
    return s;
}

struct PhongFactors {          // PhongMaterial.getBaseFactors src/materials.js:210-238 (+ Fresnel :302-308)
    float3 V, N, R;
    bool backside; float vdotn;
    float3 ambient, diffusivity, specularity, reflectivity, transmissivity;
    float smoothness;
    float kr; bool has_refr; float3 refr;
    float refl_alpha, trans_alpha;      // texture colours are RGBA: alpha takes part in the reference's `squarednorm() > 0` tests
};

template <int LEAN = 0>
JSRT_DEV void base_factors(const DeviceScene& sc, const Material& m, const SurfaceData& s, float3 ray_dir, PhongFactors& f) {
    f.V = normalized3(ray_dir) * -1.f;
    f.N = normalized3(s.normal); f.backside = false; f.vdotn = dot3(f.V, f.N);
    if (f.vdotn < 0.f) { f.N = f.N * -1.f; f.backside = true; f.vdotn = -f.vdotn; }
    f.R = normalized3(f.N * (2.f * f.vdotn) - f.V);
    f.ambient = s.basecolor * color_eval<LEAN>(sc, m.ambient, s);
    f.diffusivity = s.basecolor * color_eval<LEAN>(sc, m.diffusivity, s);
    f.specularity = color_eval<LEAN>(sc, m.specularity, s);
    f.reflectivity = color_eval<LEAN>(sc, m.reflectivity, s, &f.refl_alpha);
    f.transmissivity = color_eval<LEAN>(sc, m.transmissivity, s, &f.trans_alpha);
    f.smoothness = m.smoothness;
    f.kr = 1.f; f.has_refr = false; f.refr = f3(0.f, 0.f, 0.f);
    if (m.kind == M_FRESNEL || m.kind == M_PATH) {
        const float ior = m.ior;
        // fresnelReflectionFactor src/materials.js:366-386
        if (isfinite(ior)) {
            const float ni = f.backside ? ior : 1.f, nt = f.backside ? 1.f : ior;
            const float cosi = f.vdotn, sint = ni / nt * sqrtf(fmaxf(0.f, 1.f - cosi * cosi));
            if (sint >= 1.f) f.kr = 1.f;
            else {
                const float cost = sqrtf(fmaxf(0.f, 1.f - sint * sint));
                const float Rs = ((nt * cosi) - (ni * cost)) / ((nt * cosi) + (ni * cost));
                const float Rp = ((ni * cosi) - (nt * cost)) / ((ni * cosi) + (nt * cost));
                f.kr = (Rs * Rs + Rp * Rp) / 2.f;
            }
        }
        // getRefractionDirection src/materials.js:358-364
        const float r = f.backside ? ior : 1.f / ior, k = 1.f - r * r * (1.f - f.vdotn * f.vdotn);
        if (!(k < 0.f)) { f.has_refr = true; f.refr = (f.V * -1.f) * r + f.N * (r * f.vdotn - sqrtf(k)); }
    }
}

// colorFromLightSample: PhongMaterial src/materials.js:261-269, FresnelPhongMaterial :340-356
JSRT_DEV float3 color_from_light_sample(const Material& m, const PhongFactors& f, const LightSample& ls) {
    const float3 L = normalized3(ls.direction);
    float diffuse, specular;
    if (m.kind == M_PHONG) {
        diffuse = fmaxf(dot3(L, f.N), 0.f);
        specular = pow_clamped(fmaxf(dot3(L, f.R), 0.f), f.smoothness);
    } else {
        const float ldotn = dot3(L, f.N);
        diffuse = 0.f; specular = 0.f;
        if (f.kr > 0.f && ldotn >= 0.f) { diffuse += f.kr * ldotn; specular += f.kr * pow_clamped(fmaxf(dot3(L, f.R), 0.f), f.smoothness); }
        if (f.kr < 1.f && ldotn <= 0.f) {
            diffuse += (1.f - f.kr) * -ldotn;
            specular += (1.f - f.kr) * pow_clamped(fmaxf(f.has_refr ? dot3(L, f.refr) : 0.f, 0.f), f.smoothness);
        }
    }
    return ls.color * (f.diffusivity * diffuse) + ls.color * (f.specularity * specular);
}

// scatter(): FresnelPhongMaterial src/materials.js:335-337, PhongPathTracingMaterial :398-412.
// Returns false for a null direction.  dims: +0 mirror test, +1 lobe choice, +2,+3 spherePick.
JSRT_DEV bool scatter(const Material& m, const PhongFactors& f, bool has_R, float3 R, float3 N, uint32_t node_key, uint32_t dim_base,
                      float3& dir, float3& col) {
    col = f3(1.f, 1.f, 1.f);
    if (m.kind != M_PATH) { dir = R; return has_R; }
    if (rng_u01(node_key, dim_base + 0) < m.mirror_prob) { dir = R; return has_R; }
    const float diffuseProb = (f.diffusivity.x + f.diffusivity.y + f.diffusivity.z) / 3.f,
                specularProb = (f.specularity.x + f.specularity.y + f.specularity.z) / 3.f;
    const float probSum = diffuseProb + specularProb;
    if (probSum == 0.f) return false;
    if (rng_u01(node_key, dim_base + 1) < (diffuseProb / probSum)) {
        dir = normalized3(N + sphere_pick(rng_u01(node_key, dim_base + 2), rng_u01(node_key, dim_base + 3)));   // scatterDiffuse :438-440
        col = f.diffusivity * (1.f / CUDART_PI_F);
        return true;
    }
    col = f.specularity;       // scatterSpecular :441-445 returns R for finite smoothness
    dir = R;
    return has_R;
}

}  // namespace jsrt
