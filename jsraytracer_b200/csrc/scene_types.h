// Flattened scene: the layout the CUDA kernels read from HBM/L2.  Built on the
// host by scene_flatten.cpp from the serializer object graph and uploaded once
// per scene.  All hot arrays are 16-byte records read with 128-bit loads.
#pragma once
#include <cstdint>

namespace jsrt {

// geometry kinds (reference src/geometry.js, src/sdf.js)
enum GeomKind : int { G_PLANE = 0, G_SQUARE = 1, G_CIRCLE = 2, G_BOX = 3, G_SPHERE = 4, G_CYLINDER = 5, G_TRIANGLE = 6, G_SDF = 7, G_NEVER = 8 };
// material kinds (reference src/materials.js)
enum MatKind : int { M_PHONG = 0, M_FRESNEL = 1, M_PATH = 2, M_SOLID = 3, M_TRANSPARENT = 4 };
// top-level object kinds (reference src/world.js Primitive, src/aggregates.js Aggregate / BVHAggregate)
enum TopKind : int { T_PRIM = 0, T_BVH = 1, T_LIST = 2, T_SDF = 3 };   // T_SDF: a top-level Primitive whose geometry is an SDFGeometry (marched by its own kernel)
enum LightKind : int { L_POINT = 0, L_AREA = 1 };

enum PrimFlags : int { PF_CASTS_SHADOW = 1, PF_IDENTITY_XFORM = 2, PF_HAS_VNORMALS = 4, PF_HAS_UVS = 8 };

// 3x4 affine map, row-major: the upper three rows of a reference `inv_transform`
// (the fourth row of every transform the scene API can build is 0 0 0 1).
struct Xform { float m[12]; };
// The same map in f64 (the reference keeps matrices as f64, src/math.js:303).  Only the
// SDF path reads these: sphere tracing stops on a hard `distance <= epsilon` test
// (src/sdf.js:32), so its inputs must be rounded exactly like the reference's.
struct Xform64 { double m[12]; };

// One placed Primitive (src/world.js:104-141).  32 bytes = two 128-bit loads.
struct Prim {
    int geom_kind;     // GeomKind
    int geom_index;    // G_TRIANGLE: index into tris; G_SDF: index into sdfs; G_BOX with a non-unit AABB: index into boxes, else -1
    int material;      // index into materials
    int xform;         // index into xforms: the primitive's own inv_transform
    int flags;         // PrimFlags
    int ext_id;        // prim_id of the drop-in contract: first appearance in a DFS of world.objects
    int pad0, pad1;
};

// Triangle constants, computed like the reference constructor (src/geometry.js:335-354) in f64 from
// the f32 vertices.  128 bytes = one cache line, eight 128-bit loads; the plane test needs only the first, the FP32
// inside test of a candidate the next three, the reference-arithmetic test (edges, slivers) the rest.  The
// barycentric constants stay f64: for sliver triangles (d00 d11 - d01^2 -> 0) their f32 roundings alone
// are larger than the determinant, and real meshes are full of slivers (x_wing_fighter.obj).
struct Tri {
    float nx, ny, nz, delta;      // normal (normalised v0 x v1), delta = normal . p0
    float p0x, p0y, p0z, pad0;
    // FP32 fast path: the Cramer solve of Triangle.toBarycentric folded into two vectors on the host (in f64),
    // v = (P - p0) . A, w = (P - p0) . B with A = (d11 v0 - d01 v1) / denom, B = (d00 v1 - d01 v0) / denom.
    // ax and bx are NaN for ill-conditioned (sliver) triangles, which sends every test of such a triangle down the f64 path
    float ax, ay, az, pad1;
    float bx, by, bz, pad2;
    float v0x, v0y, v0z, pad3;
    float v1x, v1y, v1z, pad4;
    double d00, d11;
    double d01, inv_denom;        // 1 / (d00 d11 - d01^2)
};
struct TriShade {                 // per-vertex shading data (only read for shaded hits)
    float n[3][4];                // psdata.normal (w = 0)
    float uv[3][2];               // psdata.UV
    float pad[2];
};

// BVH node, 32 bytes = two 128-bit loads.  The walk is stackless: box hit -> the node's hit link (inner) or its leaf
// (then the skip link), box missed -> the skip link.  Both links are explicit absolute indices, so the nodes may be
// stored in any order; scene_flatten.cpp puts the top levels of every tree (breadth-first) in one block at the front of
// the scene's node array — the block bvh_kernel stages in shared memory — and the rest of each tree depth-first.
// Placed primitives are stored in the reference's visit order (greater child first, src/aggregates.js:221-222), which
// makes "lower primitive index" the reference's tie rule; big trees exist in eight link orders, one per ray-direction
// octant, each visiting the nearer child first (trace.cuh).
struct BvhNode {
    float cx, cy, cz, hx;         // AABB centre, half-size x   (centre/half form: src/geometry.js:189-209)
    float hy, hz;
    int skip;                     // next node when this subtree is done (absolute index), kNodeEnd when the walk is over
    int leaf;                     // inner: kNodeInner | hit link (absolute index of the first child to visit);
                                  // leaf: (count << 24) | first placed-primitive index (relative to the aggregate's first primitive), count <= 127
};
constexpr int kNodeEnd = 0x7fffffff;
constexpr int kNodeInner = (int)0x80000000u;

struct Top {                      // one entry of world.objects (src/world.js:7-15 walks them in order)
    int kind;                     // TopKind
    int xform;                    // T_BVH / T_LIST: the aggregate's inv_transform
    int first_prim;               // first placed primitive
    int prim_count;               // T_PRIM: 1
    int first_node;               // T_BVH: root BvhNode of layout 0; the root of layout q is first_node + q * (layouts >> 8)
    int node_count;               // nodes of one layout (host-side bookkeeping)
    int tri_base;                 // >= 0: every leaf object is an identity-transform, shadow-casting Triangle and
                                  // triangle index = tri_base + (placed primitive index - first_prim); else -1
    int layouts;                  // low byte: 8 = one link order per ray-direction octant (near child first), 1 = reference order only;
                                  // upper bits: distance between the roots of consecutive layouts
};

// One row of the analytic-primitive table: 32 bytes, read with two warp-uniform 128-bit loads.
struct APrim {
    int prim;        // placed primitive index (ties between equal distances go to the lower (top, prim): the reference's order)
    int top;         // index into world.objects
    int agg_xform;   // >= 0: member of a plain Aggregate, the ray is first mapped by this inv_transform (src/aggregates.js:15)
    int xform;       // the primitive's own inv_transform
    int geom_index;  // G_BOX with a non-unit AABB: index into boxes, else -1
    int flags;       // PrimFlags
    int pad0, pad1;
};
enum AGroup : int { AG_PLANE = 0, AG_SQUARE = 1, AG_BOX = 2, AG_SPHERE = 3, AG_OTHER = 4, AG_COUNT = 5 };

struct Color {                    // a MaterialColor folded to Solid, Checkerboard or scaled Texture (src/materials.js:27-131)
    float c1[3];                  // solid colour / checker colour 1 / texture: the folded ScaledMaterialColor factor
    int checker;                  // ColorKind
    float c2[3];                  // checker colour 2; texture: c2[0] = factor applied to alpha (NaN after an array scale)
    int tex;                      // texture index
};
enum ColorKind : int { CK_SOLID = 0, CK_CHECKER = 1, CK_TEXTURE = 2 };

struct Texture {                  // TextureMaterialColor (src/materials.js:77-131): RGBA8 texels in `texels` at `offset`
    int width, height;
    int flags;                    // TextureFlags
    int pad;
    unsigned long long offset;    // byte offset of texel (0, 0)
    unsigned long long pad2;
};
enum TextureFlags : int { TF_NEAREST = 1, TF_CLAMP_U = 2, TF_CLAMP_V = 4 };

struct Material {                 // src/materials.js:145-476
    int kind;                     // MatKind
    float smoothness;
    float ior;                    // refractiveIndexRatio (may be +Inf)
    float mirror_prob;
    Color ambient, diffusivity, specularity, reflectivity, transmissivity;   // M_SOLID/M_TRANSPARENT: colour in `ambient`, opacity in `smoothness`
    // PositionalUVMaterial (src/materials.js:178-193) folded into its base material: UV = (u_axis . delta, v_axis . delta),
    // delta = origin - position
    int uv_from_position;
    float uv_origin[3], u_axis[3], v_axis[3];
    int pad[2];
};

struct Light {                    // src/lights.js
    int kind;                     // LightKind
    int samples;
    int geom;                     // L_AREA: GeomKind of surface_geometry (G_SQUARE, G_CIRCLE, G_SPHERE)
    int pad;
    float pos[4];                 // L_POINT
    float color[4];               // folded solid colour (colour * intensity)
    Xform xf;                     // L_AREA transform
    Xform inv;                    // L_AREA inv_transform
};

// SDF programs: see sdf_compile.cpp / device_math.cuh for the bytecode.  The
// interpreter mirrors the reference's arithmetic exactly — points are f32 (`Vec`),
// distances / scales / matrix entries are f64 — because sphere tracing's stop test is
// a hard threshold and forward-difference normals amplify one-ulp differences.
enum SdfOp : int {
    S_END = 0,
    // the three leaves come first; their idx = 1 / 2 folds the new distance into the one below with min / max instead of pushing
    S_SPHERE,      // a0 = radius                        push |p| - r
    S_BOX,         // f32 size in (a1,a2) as 4 packed floats   push box distance
    S_TETRA,       //                                     push tetrahedron distance
    S_MIN, S_MAX,  //                                     pop 2, push
    S_NEG,         //                                     negate top
    S_ADDC,        // a0                                  top += a0
    S_SMIN,        // a0 = k                              pop b, a; push smoothMin(a, b, k)
    S_SMIN_NEGA,   // a0 = k                              pop b, a; push -smoothMin(-a, b, k)   (SmoothDifference)
    S_SMIN_NEGAB,  // a0 = k                              pop b, a; push -smoothMin(-a, -b, k)  (SmoothIntersection)
    S_PUSHP,       //                                     duplicate the point, push scale 1
    S_POPP,        //                                     drop the point and scale
    S_MULS,        //                                     top distance *= current scale
    S_XFORM,       // idx = Xform64 index, a0 = scale     p = M p; scale *= a0   (f[0] != 0: every matrix entry is an f32 value)
    S_REFL,        // f32 normal in (a1,a2), a0 = delta
    S_REP,         // f32 sizes in (a1,a2)
    S_SBEGIN,      //                                     push a local scale accumulator (= 1)
    S_SEND,        //                                     pop it; enclosing scale *= popped   (Sequence / Recursive transformers
                   //                                     return their own product, src/sdf.js:387-394,408-415)
    S_MULS_MIN,    //                                     pop d; top = min(top, d * scale)   (RecursiveTransformUnion step, src/sdf.js:353-354)
    S_CROSS,       // f[0] = a: UnionSDF of BoxSDF(Inf, a, a), BoxSDF(a, Inf, a), BoxSDF(a, a, Inf) — the Menger "cross" — as one leaf
                   //           (idx = 1 / 2 folds like the other leaves): three box distances from one q = |p| - a
    S_RTU_CROSS,   // idx = Xform64 index, a0 = its scale, f = (repetition period, cross a, iterations, xform-exact flag): the loop of a
                   //           RecursiveTransformUnionSDF whose step is { XFORM; REP (one power-of-two period); CROSS; MULS_MIN } —
                   //           the Menger sponge's recursion — run by one instruction instead of 4 x iterations dispatches
    // ---- material program (getMaterialData, src/sdf.js:86-88,102-104,119-121,149-154,...): straight-line code over a
    // stack of {distance, basecolor, UV}; every leaf is evaluated, selections / blends fold them bottom-up.
    MP_END = 32,
    MP_LEAF,       // f[0..2] = basecolor, idx = 1: UV = cartesianToSpherical(normalize(p)) (SphereSDF), 0: no UV
    MP_ATTACH,     // idx = first instruction of the node's own distance program: top.d = distance(p)
    MP_SELMIN,     // pop b, a: keep a unless b.d < a.d          (UnionSDF: Math.indexOfMin, first minimum wins)
    MP_SELMAX,     // pop b, a: keep a unless b.d > a.d          (IntersectionSDF)
    MP_DIFF,       // pop neg, pos: pos.d > -neg.d ? pos : neg   (DifferenceSDF)
    MP_BLEND_U,    // a0 = k: SmoothUnion        mix = smoothMinBlend(a.d, b.d, k)
    MP_BLEND_I,    // a0 = k: SmoothIntersection mix = 1 - smoothMinBlend(-a.d, -b.d, k)
    MP_BLEND_D,    // a0 = k: SmoothDifference   mix = smoothMinBlend(-pos.d, neg.d, k)
};
struct SdfInstr { int op; int idx; double a0; float f[4]; };   // 32 bytes: two 128-bit loads
struct SdfProgram {
    int first_instr, instr_count;
    int max_samples;
    int uniform_base;             // 1: every leaf has the same basecolor (in `base`)
    double distance_epsilon, max_trace_distance, normal_step_size;
    float cx, cy, cz, hx, hy, hz; // SDFGeometry.aabb
    float base[3];
    int mat_first;                // first instruction of the material program, -1 if uniform_base
    float pad1[2];
};

struct Camera {                   // src/cameras.js; f64 so primary rays match the reference bit for bit
    double t[12];                 // transform rows 0..2
    double tan_fov, aspect, focus_distance, sensor_size;
    int dof, pad;
};

}  // namespace jsrt
