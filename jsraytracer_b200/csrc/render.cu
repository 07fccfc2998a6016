// Wavefront renderer: generate / extend / shade / shadow kernels over ray queues
// that live in HBM, plus the accumulation buffer and its 8-bit resolve.
//
// Replaces the reference's per-pixel recursion
//   IncrementalMultisamplingRenderer.render -> World.color -> Primitive.color ->
//   Material.color -> World.color ...   (src/renderers.js:87-98, src/world.js:31-41,
//   src/materials.js:271-333)
// with level-synchronous waves: level L holds every ray of path-tree depth L of
// the camera samples in the current batch.  A shaded hit appends up to two
// children (reflection, transmission/refraction) to the next level's queue and
// one shadow ray per light sample to the shadow queue; both appends are
// compacted with a warp ballot + prefix popcount and one atomicAdd per warp.
//
// Queues are SoA of float4 (128-bit coalesced loads/stores):
//   ray:    o.xyz | pixel      d.xyz | node id      throughput.rgb | pass<<8 | depth_remaining
//   hit:    t | placed prim | top-level object | -
//   shadow: o.xyz | pixel      d.xyz (unnormalised, light at t=1) | -      contribution.rgb | -
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "render.h"
#include "shade.cuh"

namespace jsrt {

namespace {

constexpr int kBlock = 256;
// threads per CTA of shade_kernel (its warps run in lock-step, see there) and CTAs per SM
#ifndef JSRT_SHADE_BLOCK
#define JSRT_SHADE_BLOCK 256
#endif
constexpr int kShadeBlock = JSRT_SHADE_BLOCK;
#ifndef JSRT_SHADE_MIN_BLOCKS
#define JSRT_SHADE_MIN_BLOCKS 3
#endif
// SDF march: the interpreter stalls on fixed-latency f64 dependency chains ("wait" 37 % of stall samples at 2 CTAs / SM,
// FP64 pipe 7.5 % busy), so occupancy beats registers: SDF_Menger 1080p extend 177.7 / 134.6 / 125.5 ms and shadow
// 68.8 / 56.5 / 61.2 ms per 16 passes at 2 / 3 / 4 CTAs per SM (128 / 80 / 64 registers; profiles/r1_s3)
#ifndef JSRT_SDF_MIN_BLOCKS
#define JSRT_SDF_MIN_BLOCKS 4
#endif
#ifndef JSRT_SDF_RTU_MIN_BLOCKS
#define JSRT_SDF_RTU_MIN_BLOCKS 4      // the build with S_RTU_CROSS: 2 / 3 / 4 CTAs per SM = 742 / 761 / 787 Mrays/s on SDF_Menger (profiles/r2_ab.md §4)
#endif
#ifndef JSRT_SDF_SHADOW_MIN_BLOCKS
#define JSRT_SDF_SHADOW_MIN_BLOCKS 3
#endif
#ifndef JSRT_BVH_BLOCK
#define JSRT_BVH_BLOCK 1024        // threads per CTA of bvh_kernel (see there)
#endif
#ifndef JSRT_BVH_MIN_BLOCKS
#define JSRT_BVH_MIN_BLOCKS 1      // x 1024 threads = 64 registers each, the same 32 warps per SM as round 1's 4 x 256
#endif

struct RayQueue { float4* o; float4* d; float4* w; };
struct ShadowQueue { float4* o; float4* d; float4* c; };

// counters[0], [1]: ray queue counts (ping-pong); [2]: shadow queue count
// Every counter sits in a 128-byte line of its own (JSRT_COUNTER_PAD=0: packed as in round 1): the queue counters that one
// kernel bumps concurrently — shade_kernel's children and shadow counts, a trace wave's work-list count and cursor — shared
// one 32-byte sector, i.e. one L2 atomic unit serialised all of them (profiles/r2_ab.md §6).
#ifndef JSRT_COUNTER_PAD
#define JSRT_COUNTER_PAD 1
#endif
struct alignas(JSRT_COUNTER_PAD ? 128 : 4) Cnt { int v; };
constexpr int kCntStride = (int)(sizeof(Cnt) / sizeof(int));      // distance between consecutive counters, in ints
struct Counters { Cnt ray[2], shadow, ties, cursor_extend, cursor_shadow, cursor_extend_sdf, cursor_shadow_sdf, list_extend, list_shadow; unsigned long long stats[24]; };
enum { ST_PRIMARY = 0, ST_SECONDARY = 1, ST_SHADOW = 2, ST_SHADED = 3, ST_SAMPLES = 4,
       ST_NODES = 8, ST_LEAF_PRIMS = 11, ST_TOP_PRIMS = 14, ST_SDF_EVALS = 17 };   // + ray class (0 primary, 1 secondary, 2 shadow)

__device__ __forceinline__ void flush_work(unsigned long long* stats, int cls, const Work& w) {
    unsigned long long v[4] = {w.nodes, w.leaf_prims, w.top_prims, w.sdf_evals};
    const int slot[4] = {ST_NODES, ST_LEAF_PRIMS, ST_TOP_PRIMS, ST_SDF_EVALS};
    for (int k = 0; k < 4; ++k) {
        unsigned long long x = v[k];
        for (int off = 16; off; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(stats + slot[k] + cls, x);
    }
}

// ---------------------------------------------------------------------------------
// generate: the camera rays of a batch as a kernel of its own (jsrt_primary_hits, JSRT_FUSE_GEN=0); renders normally
// let prims_kernel<extend, GEN> compute them (trace.cuh: generate_sample).
__global__ void __launch_bounds__(kBlock) generate_kernel(const __grid_constant__ GenParams g) {
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < g.n_samples; s += stride) {
        float4 o4, d4, w4; uint32_t pixel;
        generate_sample(g, s, o4, d4, w4, pixel);
        g.qo[s] = o4; g.qd[s] = d4; g.qw[s] = w4;
        if (g.count_samples) atomicAdd(&g.accum[pixel].w, 1.0f);      // samples taken for this pixel
    }
}

// ---------------------------------------------------------------------------------
// extend: closest hit for every ray of the level (World.cast, src/world.js:28-30);
// shadow: `world.cast(new Ray(position, direction), 0.0001, 1, false)`, the sample is
// dropped iff 0 < t < 1 (src/materials.js:250-252).  Two kernels each: trace.cuh.
// prims_kernel streams the queue.  Without a bound ptxas gives it 72 registers = 3 CTAs / SM; measured
// (profiles/r1_s4/ab_s4_prims_*.jsonl, 3 / 4 / 5 / 6 / 8 CTAs per SM): cornell_box_path 9 924 / 10 545 / 10 627 / 10 411 /
// 10 036 Mrays/s, bunny_path 5 705 / 5 723 / 5 540 / 5 388 / 5 320 (spills from 5 on).  4 = 64 registers, no spills.
#ifndef JSRT_PRIMS_MIN_BLOCKS
#define JSRT_PRIMS_MIN_BLOCKS 4
#endif
template <int MODE, bool COUNT, bool HAS_SDF, bool GEN>
__global__ void __launch_bounds__(kBlock, (JSRT_PRIMS_MIN_BLOCKS > 0 && !HAS_SDF && !COUNT && !GEN) ? JSRT_PRIMS_MIN_BLOCKS : 1) prims_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TraceIO io, const __grid_constant__ GenParams gen) {
    Work wp, ws;
    prims_wave<MODE, COUNT, HAS_SDF, GEN>(sc, io, &gen, &wp, &ws);
    if (COUNT) { if (MODE == TM_EXTEND) { flush_work(io.stats, 0, wp); flush_work(io.stats, 1, ws); } else flush_work(io.stats, 2, ws); }
}
// RTU: the build whose interpreter knows S_RTU_CROSS (scenes whose programs contain it; device_math.cuh: sdf_rtu_cross)
template <int MODE, bool COUNT, bool RTU>
__global__ void __launch_bounds__(kBlock, RTU ? JSRT_SDF_RTU_MIN_BLOCKS : MODE == TM_SHADOW ? JSRT_SDF_SHADOW_MIN_BLOCKS : JSRT_SDF_MIN_BLOCKS) sdf_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TraceIO io) {
    Work wp, ws;
    sdf_wave<MODE, COUNT, RTU>(sc, io, &wp, &ws);
    if (COUNT) { if (MODE == TM_EXTEND) { flush_work(io.stats, 0, wp); flush_work(io.stats, 1, ws); } else flush_work(io.stats, 2, ws); }
}
// One big CTA per SM (JSRT_BVH_BLOCK threads = 64 registers each) so that the staged top levels of the trees exist once
// per SM: up to 7 168 nodes = 224 KB of shared memory, 4 096 = 128 KB by default (the other half of the unified array
// stays L1 for the deep nodes, the triangles and the ray records).
extern __shared__ float4 s_staged_nodes[];
template <int MODE, bool COUNT, bool HAS_SDF, bool DIRECT, bool TLAS, bool MESH = false>
__global__ void __launch_bounds__(JSRT_BVH_BLOCK, JSRT_BVH_MIN_BLOCKS) bvh_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TraceIO io) {
    {
        const float4* src = reinterpret_cast<const float4*>(sc.nodes);
        // Stored as two arrays (first halves | second halves).  A lane's LDS.128 then lands on 16-byte slot (node mod 8) of
        // the 128-byte bank row instead of only the even (or only the odd) slots of a 32-byte record, which halves the
        // structural bank conflicts of a warp whose 32 lanes read 32 different nodes (JSRT_STAGE_SOA=0: records as in HBM).
        const int ns = sc.n_staged;
        for (int k = threadIdx.x; k < 2 * ns; k += blockDim.x) s_staged_nodes[JSRT_STAGE_SOA ? ((k >> 1) + (k & 1) * ns) : k] = __ldg(src + k);
        __syncthreads();
    }
    Work wp, ws;
    bvh_wave<MODE, COUNT, HAS_SDF, DIRECT, TLAS, MESH>(sc, io, &wp, &ws, s_staged_nodes);
    if (COUNT) { if (MODE == TM_EXTEND) { flush_work(io.stats, 0, wp); flush_work(io.stats, 1, ws); } else flush_work(io.stats, 2, ws); }
}

__global__ void __launch_bounds__(kBlock) tie_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ TraceIO io) { tie_wave(sc, io); }

__device__ __forceinline__ void accum_add(float4* accum, uint32_t pixel, float3 c) { accum_add3(accum, pixel, c); }

// ---------------------------------------------------------------------------------
// shade: World.color's miss / hit handling (src/world.js:31-41), Primitive.color
// (:125-137), Material.color (src/materials.js).  Emits ambient, pushes shadow rays
// and children.
#ifndef JSRT_SHADE_LOCKSTEP
#define JSRT_SHADE_LOCKSTEP 2
#endif
// Shaded sphere / cylinder hits: the accepted root re-solved in f64 (device_math.cuh: sphere_intersect64); the f32 value is
// kept if the two disagree about which root it was.  Out of line: f64 sqrt + two divisions that mesh scenes never execute.
JSRT_RARE double round_hit_distance64(bool sphere, float lox, float loy, float loz, float ldx, float ldy, float ldz, double md, float t, double td) {
    const double t64 = sphere ? sphere_intersect64(f3(lox, loy, loz), f3(ldx, ldy, ldz), md) : sphere_intersect64(f3(lox, loy, 0.f), f3(ldx, ldy, 0.f), md);
    return (fabs(t64 - (double)t) <= 1e-4 * fabs((double)t)) ? t64 : td;
}
struct ShadeIO {
    RayQueue q; const int* count; const float4* hits;
    RayQueue next; int* next_count; int next_cap;
    ShadowQueue sq; int* shadow_count; int shadow_cap;
    float4* accum;                      // radiance destination: the pixel sums, or the per-sample buffer (JSRT_FLAG_AOV), or the batch scratch
    unsigned long long seed; unsigned long long* stats; int* overflow;
    const float4* sdf_normals;
    int accum_stride, pass0;
    float4* aov_nd; float4* aov_var;
};
// FUSE (scenes without SDFs): the shadow ray of every light sample is tested here against everything that needs no
// walk — the analytic shadow casters and the root boxes of the BVHAggregates (trace.cuh: analytic_hits,
// first_bvh_hit).  An occluded sample is dropped, an unoccluded one that reaches no BVH is added to its pixel at once,
// and only the walkers go to the shadow queue (with their first BVH), which bvh_kernel<shadow, DIRECT> consumes as its
// work list.  Round 1 wrote all of them (64 B each) for prims_kernel<shadow> to read back and find that most need no walk
// (bunny_path: 2/3 of the shadow rays; cornell_box_path: all of them — its shadow queue is never touched now).
template <bool HAS_SDF, bool SORT, bool FUSE, bool COUNT, int LEAN = 0>
__global__ void __launch_bounds__(kShadeBlock, HAS_SDF ? 1 : JSRT_SHADE_MIN_BLOCKS) shade_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ ShadeIO io) {
    const RayQueue& q = io.q; const RayQueue& next = io.next; const ShadowQueue& sq = io.sq;
    const float4* __restrict__ hits = io.hits;
    float4* __restrict__ accum = io.accum;
    const int n = *io.count;
    const int stride = gridDim.x * blockDim.x;
    unsigned long long my_shaded = 0, my_shadow = 0;
    Work ws;
    // Shading sorted by material (SORT): the kBlock rays of a tile are permuted inside the CTA by a counting sort on
    // (miss | material index) before they are shaded, so that a warp runs one material's code path instead of the
    // union of its 32 rays' paths.  The permutation stays inside a 4 KB window of each queue array, so the loads
    // still hit whole sectors, and results do not depend on the order (the RNG is keyed by pixel / pass / path
    // node).  The price is three CTA barriers per tile, which tie the fast warps (misses) to the slow ones; measured
    // (profiles/r1_ab.md): -6 % shade time on cornell_box_path (closed box, 12 analytic primitives with their own
    // materials, depth 8), +5..10 % on bunny_path / dragon / starwars, whose camera and shadow-side rays are coherent
    // already.  The host turns it on for scenes without BVH aggregates (JSRT_SHADE_SORT=0/1 overrides).
    __shared__ int s_hist[64], s_off[64];
    __shared__ unsigned short s_perm[kShadeBlock];
    if (SORT) { if (threadIdx.x < 64) s_hist[threadIdx.x] = 0; __syncthreads(); }
    // Lockstep (JSRT_SHADE_LOCKSTEP, unsorted tiles): this kernel is 8 192 SASS instructions of which a ray executes ~1 500 once,
    // and instruction fetch was 28-44 % of its stall samples (profiles/r2_ab.md §3).  CTA barriers at the loop top (1) and
    // before the light loop (2; 3 = also at every light sample) keep the 8 warps of a CTA in the same stretch of the code,
    // so that one instruction-cache fill serves them all: bunny_path shade 6.84 -> 6.60 / 6.48 / 6.58 ms, dragon and
    // cornell_box_path (sorted tiles have their own barriers) within noise (profiles/r2/ab_r2o_*).
    constexpr int LOCKSTEP = SORT ? 0 : JSRT_SHADE_LOCKSTEP;
    const int n_round = (SORT || LOCKSTEP) ? ((n + kShadeBlock - 1) / kShadeBlock) * kShadeBlock : ((n + 31) & ~31);      // block- / warp-uniform trip count
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n_round; i0 += stride) {
        int i = i0;
        if (LOCKSTEP >= 1) __syncthreads();
        if (SORT) {
            const int tile = i0 - (int)threadIdx.x;
            const int j = tile + threadIdx.x, lane_s = threadIdx.x & 31;
            int key = 63;
            if (j < n) {
                const int prim = __float_as_int(hits[j].y);
                key = prim < 0 ? 0 : min(1 + __ldg(&sc.prims[prim].material), 62);
            }
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane_s == leader) base = atomicAdd(&s_hist[key], __popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader) + __popc(peers & ((1u << lane_s) - 1u));
            __syncthreads();
            if (threadIdx.x < 32) {
                const int v0 = s_hist[lane_s], v1 = s_hist[lane_s + 32];
                int a = v0, b = v1;
                for (int off = 1; off < 32; off <<= 1) {
                    const int ta = __shfl_up_sync(0xffffffffu, a, off), tb = __shfl_up_sync(0xffffffffu, b, off);
                    if (lane_s >= off) { a += ta; b += tb; }
                }
                const int total0 = __shfl_sync(0xffffffffu, a, 31);
                s_off[lane_s] = a - v0; s_off[lane_s + 32] = total0 + b - v1;
                s_hist[lane_s] = 0; s_hist[lane_s + 32] = 0;
            }
            __syncthreads();
            s_perm[s_off[key] + base] = (unsigned short)threadIdx.x;
            __syncthreads();
            i = tile + s_perm[threadIdx.x];
        }
        const bool active = i < n;
        bool hit = false;
        uint32_t pixel = 0, node = 0, slot = 0; int depth_rem = 0, pass = 0;
        float3 o = f3(0, 0, 0), d = f3(0, 0, 1), thr = f3(0, 0, 0);
        SurfaceData s; s.position = f3(0, 0, 0); s.normal = f3(0, 0, 1); s.uv = make_float2(0, 0); s.has_uv = false; s.basecolor = f3(1, 1, 1);
        PhongFactors f;
        const Material* mat = sc.materials;
        uint32_t node_key = 0;
        if (active) {
            const float4 o4 = q.o[i], d4 = q.d[i], w4 = q.w[i], h4 = hits[i];
            o = f3(o4.x, o4.y, o4.z); d = f3(d4.x, d4.y, d4.z); thr = f3(w4.x, w4.y, w4.z);
            pixel = (uint32_t)__float_as_int(o4.w); node = (uint32_t)__float_as_int(d4.w);
            const int packed = __float_as_int(w4.w); depth_rem = packed & 0xff; pass = packed >> 8;
            slot = pixel + (uint32_t)(pass - io.pass0) * (uint32_t)io.accum_stride;      // JSRT_FLAG_AOV: radiance slot of this sample (stride 0 otherwise)
            const int prim = __float_as_int(h4.y), top = __float_as_int(h4.z);
            const float t = h4.x;
            if (prim < 0) {
                accum_add(accum, slot, thr * f3(sc.bg[0], sc.bg[1], sc.bg[2]));      // `return this.bg_color`
            } else {
                hit = true; ++my_shaded;
                const int4* pp = reinterpret_cast<const int4*>(sc.prims + prim);
                const int4 pa = __ldg(pp); const int flags = __ldg(reinterpret_cast<const int*>(pp + 1));
                // inv_transform = prim.inv * ancestorInvTransform (src/world.js:126)
                XformReg inv = load_xform(sc.xforms, pa.w);
                const int4 ta = __ldg(reinterpret_cast<const int4*>(sc.tops + top));
                if (ta.x == T_BVH || ta.x == T_LIST) {
                    const XformReg anc = load_xform(sc.xforms, ta.y);
                    inv = (flags & PF_IDENTITY_XFORM) ? anc : xf_compose(inv, anc);
                }
                float3 lp;
                double td = (double)t + (double)h4.w;
                if (HAS_SDF && pa.x == G_SDF && ta.x == T_SDF) {
                    // the SDF normal is a forward difference: recompute the reference's local hit point exactly
                    // (f64 matrix, f64 distance carried as t + t_lo)
                    const double* m64 = sc.xforms64[pa.w].m;
                    lp = ray_point_f64(xf64_apply(m64, o, 1.0), xf64_apply(m64, d, 0.0), (double)t + (double)h4.w);
                } else {
                    const float3 lo = xf_point(inv, o), ld = xf_dir(inv, d);
                    if (LEAN < 2 && (pa.x == G_SPHERE || pa.x == G_CYLINDER)) {
                        // re-solve the accepted hit in f64 (see sphere_intersect64); keep the f32 value if the
                        // two disagree about which root it was
                        td = round_hit_distance64(pa.x == G_SPHERE, lo.x, lo.y, lo.z, ld.x, ld.y, ld.z, (node == 1u) ? 0.0 : 0.0001, t, td);
                    }
                    lp = ray_point_f64(lo, ld, td);
                }
                float3 ln; float4 sn = make_float4(0, 0, 0, 0);
                if (HAS_SDF && io.sdf_normals && ta.x == T_SDF) sn = io.sdf_normals[i];
                material_data<HAS_SDF, LEAN>(sc, pa.x, pa.y, flags, lp, ln, s.uv, s.has_uv, s.basecolor, &sn);
                s.normal = normalized3(xf_normal(inv, ln));
                s.position = ray_point_f64(o, d, td);
                if (io.aov_nd && node == 1u) {
                    // first-hit AOVs of the GL path (gl/src/WebGLRendererAdapter.js:376-379: `initial_intersection_distance =
                    // length(r.o - intersect_position)`, `first_hit_normal.xyz = intersect_normal`), summed per pixel
                    const float3 e = s.position - o;
                    accum_add3w(io.aov_nd, pixel, s.normal, sqrtf(dot3(e, e)));
                    atomicAdd(&io.aov_var[pixel].w, 1.0f);
                }
                mat = sc.materials + pa.z;
                if (LEAN < 1 && mat->uv_from_position) {          // PositionalUVMaterial.color src/materials.js:188-192
                    const float dx = (float)dsub(mat->uv_origin[0], s.position.x), dy = (float)dsub(mat->uv_origin[1], s.position.y), dz = (float)dsub(mat->uv_origin[2], s.position.z);
                    s.uv = make_float2((float)ddot3(mat->u_axis[0], mat->u_axis[1], mat->u_axis[2], dx, dy, dz), (float)ddot3(mat->v_axis[0], mat->v_axis[1], mat->v_axis[2], dx, dy, dz));
                    s.has_uv = true;
                }
                node_key = rng_node_key(rng_sample_key(io.seed, pixel, (uint32_t)pass), node);
                if (LEAN < 1 && mat->kind == M_SOLID) {
                    accum_add(accum, slot, thr * color_eval<LEAN>(sc, mat->ambient, s));
                    hit = false;                 // no lights, no children (src/materials.js:153-155)
                } else if (LEAN < 1 && mat->kind == M_TRANSPARENT) {
                    accum_add(accum, slot, thr * (color_eval<LEAN>(sc, mat->ambient, s) * mat->smoothness));
                } else {
                    base_factors<LEAN>(sc, *mat, s, d, f);
                    accum_add(accum, slot, thr * f.ambient);      // `let ret = data.ambient`
                }
            }
        }
        if (LOCKSTEP >= 2) __syncthreads();
        const bool lit = hit && mat->kind != M_TRANSPARENT;
        // ---- children (src/materials.js:277-288, 315-330, 169-172).  world.color with
        // recursionDepth - 1 == 0 returns black without casting (src/world.js:32-33).
        bool want0 = false, want1 = false; float3 dir0 = f3(0, 0, 1), dir1 = f3(0, 0, 1), w0 = f3(0, 0, 0), w1 = f3(0, 0, 0);
        if (hit && depth_rem - 1 > 0) {
            if (mat->kind == M_PHONG) {
                if (dot3(f.reflectivity, f.reflectivity) + f.refl_alpha * f.refl_alpha > 0.f) { want0 = true; dir0 = f.R; w0 = thr * f.reflectivity; }
                if (dot3(f.transmissivity, f.transmissivity) + f.trans_alpha * f.trans_alpha > 0.f) { want1 = true; dir1 = normalized3(d); w1 = thr * f.transmissivity; }
            } else if (mat->kind == M_TRANSPARENT) {
                want0 = true; dir0 = d; w0 = thr * (1.f - mat->smoothness);
            } else {
                const uint32_t sb = DIM_LIGHTS + 2u * (uint32_t)sc.light_samples;
                float3 col;
                if (f.kr > 0.f && scatter(*mat, f, true, f.R, f.N, node_key, sb, dir0, col)) { want0 = true; w0 = thr * (col * f.reflectivity * f.kr); }
                if (f.kr < 1.f && scatter(*mat, f, f.has_refr, f.refr, f.N * -1.f, node_key, sb + 4, dir1, col)) { want1 = true; w1 = thr * (col * f.transmissivity * (1.f - f.kr)); }
            }
        }
        // ---- queue space: ONE atomicAdd per warp per queue per iteration.  (One atomic per appended item made the return
        // latency of the contended counter 60 % of this kernel's stall samples: profiles/r1_ncu_summary.md.)
        const int lane = threadIdx.x & 31;
        const unsigned lit_mask = __ballot_sync(0xffffffffu, lit), m0 = __ballot_sync(0xffffffffu, want0), m1 = __ballot_sync(0xffffffffu, want1);
        const int n_lit = __popc(lit_mask), n0 = __popc(m0), n1 = __popc(m1);
        int base_s = 0, base_c = 0;
        if (lane == 0) {
            if (!FUSE && n_lit) base_s = atomicAdd(io.shadow_count, n_lit * sc.light_samples);
            if (n0 + n1) base_c = atomicAdd(io.next_count, n0 + n1);
        }
        const unsigned lt = (1u << lane) - 1u;
        const int rank_s = __popc(lit_mask & lt), rank0 = __popc(m0 & lt), rank1 = n0 + __popc(m1 & lt);
        const int packed_next = (pass << 8) | (depth_rem - 1);
        auto write_children = [&]() {
            base_c = __shfl_sync(0xffffffffu, base_c, 0);
            if (want0) {
                const int cs = base_c + rank0;
                if (cs < io.next_cap) {
                    next.o[cs] = make_float4(s.position.x, s.position.y, s.position.z, __int_as_float((int)pixel));
                    next.d[cs] = make_float4(dir0.x, dir0.y, dir0.z, __int_as_float((int)rng_child_node(node, 0u)));
                    next.w[cs] = make_float4(w0.x, w0.y, w0.z, __int_as_float(packed_next));
                } else *io.overflow = 1;
            }
            if (want1) {
                const int cs = base_c + rank1;
                if (cs < io.next_cap) {
                    next.o[cs] = make_float4(s.position.x, s.position.y, s.position.z, __int_as_float((int)pixel));
                    next.d[cs] = make_float4(dir1.x, dir1.y, dir1.z, __int_as_float((int)rng_child_node(node, 1u)));
                    next.w[cs] = make_float4(w1.x, w1.y, w1.z, __int_as_float(packed_next));
                } else *io.overflow = 1;
            }
        };
        // FUSE: the children leave before the light loop (their twelve values would otherwise stay live across the
        // intersection tests); otherwise after it, so that the counter's return latency hides behind the light arithmetic
        if (FUSE) write_children();
        if (lit) my_shadow += (unsigned long long)sc.light_samples;

        // ---- shadow rays: one per light sample (src/materials.js:244-257)
        uint32_t dim = DIM_LIGHTS;
        int j = 0;
        bool have_base = false;
        for (int li = 0; li < sc.n_lights; ++li) {
            const Light& L = sc.lights[li];
            const int ns = L.samples;
            for (int k = 0; k < ns; ++k, dim += 2, ++j) {
                if (LOCKSTEP >= 3) __syncthreads();
                LightSample ls; ls.direction = f3(0, 0, 1); float3 contrib = f3(0, 0, 0);
                if (lit) {
                    // (a point light draws no random numbers: the two hashes are only evaluated for area lights)
                    const bool area = L.kind != L_POINT;
                    ls = light_sample<LEAN>(L, s.position, area ? rng_u01(node_key, dim) : 0.f, area ? rng_u01(node_key, dim + 1) : 0.f);
                    contrib = thr * (color_from_light_sample(*mat, f, ls) * (1.0f / (float)ns));
                }
                if (FUSE) {
                    // world.cast(new Ray(position, direction), 0.0001, 1, false) as far as it goes without a walk
                    int fb = -1; bool walker = false;
                    if (lit) {
                        Hit sb; sb.t = CUDART_INF_F; sb.prim = -1; sb.top = -1; sb.t_lo = 0.f;
                        analytic_hits<true, COUNT, false, (LEAN >= 2)>(sc, s.position, ls.direction, 0.0001f, 1.0f, lit_mask, sb, &ws);
                        if (sb.prim < 0) {
                            LocalRay lr;
                            fb = first_bvh_hit<COUNT>(sc, s.position, ls.direction, 0.0001f, 1.0f, CUDART_INF_F, &ws, lr);
                            if (fb < 0) accum_add(accum, slot, contrib);       // unoccluded: the light sample counts (src/materials.js:251-253)
                            else walker = true;
                        }
                    }
                    const unsigned wm = __ballot_sync(0xffffffffu, walker);
                    if (wm) {
                        const int leader = __ffs(wm) - 1;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(io.shadow_count, __popc(wm));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (walker) {
                            const int e = base + __popc(wm & lt);
                            if (e < io.shadow_cap) {
                                sq.o[e] = make_float4(s.position.x, s.position.y, s.position.z, __int_as_float((int)pixel));
                                sq.d[e] = make_float4(ls.direction.x, ls.direction.y, ls.direction.z, __int_as_float(pass));
                                sq.c[e] = make_float4(contrib.x, contrib.y, contrib.z, __int_as_float(fb));
                            } else *io.overflow = 1;
                        }
                    }
                } else {
                    // sample j of the warp's lit lanes occupies slots [base_s + j * n_lit, base_s + (j + 1) * n_lit): coalesced per sample
                    if (!have_base) { base_s = __shfl_sync(0xffffffffu, base_s, 0); have_base = true; }
                    if (lit) {
                        const int e = base_s + j * n_lit + rank_s;
                        if (e < io.shadow_cap) {
                            sq.o[e] = make_float4(s.position.x, s.position.y, s.position.z, __int_as_float((int)pixel));
                            sq.d[e] = make_float4(ls.direction.x, ls.direction.y, ls.direction.z, __int_as_float(pass));
                            sq.c[e] = make_float4(contrib.x, contrib.y, contrib.z, 0.f);
                        } else *io.overflow = 1;
                    }
                }
            }
        }
        if (!FUSE) write_children();
    }
    // per-warp reduction of the shaded-hit / shadow-ray counts, one atomic per warp
    for (int off = 16; off; off >>= 1) { my_shaded += __shfl_down_sync(0xffffffffu, my_shaded, off); my_shadow += __shfl_down_sync(0xffffffffu, my_shadow, off); }
    if ((threadIdx.x & 31) == 0 && my_shaded) atomicAdd(io.stats + ST_SHADED, my_shaded);
    if (FUSE && (threadIdx.x & 31) == 0 && my_shadow) atomicAdd(io.stats + ST_SHADOW, my_shadow);
    if (FUSE && COUNT) flush_work(io.stats, 2, ws);
}

// JSRT_FLAG_AOV: the batch's per-sample radiance -> pixel sums + running variance, one thread per pixel, samples of a
// pixel in pass order.  The recurrence is the GL path's (gl/src/WebGLRendererAdapter.js:352-356), evaluated in FP32 like
// its shader: count1 = n + 1; mean = sum / count1 (sum and n before this sample); delta = sample - mean;
// delta2 = sample - (delta / count1 + mean); variance += delta * delta2; sum += sample.
__global__ void __launch_bounds__(kBlock) fold_samples_kernel(const __grid_constant__ GenParams g, int pass0, int span, float4* __restrict__ sample_rad,
                                                            float4* __restrict__ accum, float4* __restrict__ var, int npix) {
    const int stride = gridDim.x * blockDim.x;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < g.npix_active; idx += stride) {
        const int py = idx / g.ncols, px = g.x_offset + (idx % g.ncols) * g.x_delt;
        const int pixel = py * g.width + px;
        int members = 0;
        for (int p = 0; p < span; ++p) {
            const long long gs = (long long)(pass0 + p - g.first_pass) * g.npix_active + idx;
            if (gs >= g.first_sample && gs < g.first_sample + g.n_samples) ++members;
        }
        if (!members) continue;
        float4 a = accum[pixel], v = var[pixel];
        float n = a.w - (float)members;             // generate_kernel has already counted this batch's samples
        for (int p = 0; p < span; ++p) {
            const long long gs = (long long)(pass0 + p - g.first_pass) * g.npix_active + idx;
            if (!(gs >= g.first_sample && gs < g.first_sample + g.n_samples)) continue;
            const size_t slot = (size_t)p * npix + pixel;
            const float4 L = sample_rad[slot];
            sample_rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float count1 = n + 1.f;
            const float mx = a.x / count1, my = a.y / count1, mz = a.z / count1;
            const float dx = L.x - mx, dy = L.y - my, dz = L.z - mz;
            v.x += dx * (L.x - (dx / count1 + mx)); v.y += dy * (L.y - (dy / count1 + my)); v.z += dz * (L.z - (dz / count1 + mz));
            a.x += L.x; a.y += L.y; a.z += L.z;
            n += 1.f;
        }
        accum[pixel] = a; var[pixel] = v;
    }
}

// bookkeeping between levels: fold queue sizes into the ray statistics and recycle the counters
// (fused: shade_kernel<FUSE> has counted the shadow rays itself — the queue only holds the walkers)
__global__ void level_end_kernel(Counters* c, int cur, int level, int next_cap, int shadow_cap, int fused) {
    c->stats[level == 0 ? ST_PRIMARY : ST_SECONDARY] += (unsigned long long)c->ray[cur].v;
    if (!fused) c->stats[ST_SHADOW] += (unsigned long long)min(c->shadow.v, shadow_cap);
    c->ray[cur].v = 0;
    c->shadow.v = 0;
    c->cursor_extend.v = 0; c->cursor_shadow.v = 0; c->cursor_extend_sdf.v = 0; c->cursor_shadow_sdf.v = 0; c->list_extend.v = 0; c->list_shadow.v = 0; c->ties.v = 0;
    if (c->ray[cur ^ 1].v > next_cap) c->ray[cur ^ 1].v = next_cap;
}
__global__ void set_count_kernel(Counters* c, int which, int n, int count_samples) {
    c->ray[which].v = n; c->ray[which ^ 1].v = 0; c->shadow.v = 0; c->cursor_extend.v = 0; c->cursor_shadow.v = 0; c->cursor_extend_sdf.v = 0; c->cursor_shadow_sdf.v = 0; c->list_extend.v = 0; c->list_shadow.v = 0; c->ties.v = 0;
    if (count_samples) c->stats[ST_SAMPLES] += (unsigned long long)n;
}

// Accumulation buffers of the other GPUs that render passes of the same frame (jsrt_scene_create with ndev > 1:
// peer pointers of this process; jsrt_accum_attach: CUDA IPC mappings of other processes' buffers).  The kernels below
// read them straight over NVLink: the cross-GPU sum is fused into the resolve / read-back, there is no staging copy
// and no separate reduction pass.  Replaces the compositing of src/raytrace_launcher.js:92-97.
constexpr int kMaxPeers = 15;
struct PeerAccum { const float4* p[kMaxPeers]; int n; };
__device__ __forceinline__ float4 summed_pixel(const float4* __restrict__ accum, const PeerAccum& peers, int i) {
    float4 a = accum[i];
    for (int k = 0; k < peers.n; ++k) { const float4 b = __ldcs(peers.p[k] + i); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    return a;
}
// PixelBuffer.setColor on buffer.times(1/(iter+1)) (src/renderers.js:98, src/pixelbuffer.js:39-49)
__global__ void resolve_kernel(const float4* __restrict__ accum, const __grid_constant__ PeerAccum peers, uchar4* __restrict__ out, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 a = summed_pixel(accum, peers, i);
    if (!(a.w > 0.f)) { out[i] = make_uchar4(0, 0, 0, 0); return; }
    const double inv = 1.0 / (double)a.w;
    float c[3] = {(float)((double)a.x * inv), (float)((double)a.y * inv), (float)((double)a.z * inv)};
    unsigned char r[3];
    for (int k = 0; k < 3; ++k) {
        double comp = fmin(fmax((double)c[k], 0.0), 1.0);     // Math.min(Math.max(c, 0), 1): NaN stays NaN -> 0 in a Uint8ClampedArray
        if (c[k] != c[k]) comp = 0.0;
        r[k] = (unsigned char)floor(255.0 * comp + 0.5);      // Math.round
    }
    out[i] = make_uchar4(r[0], r[1], r[2], 255);
}
__global__ void sum_peers_kernel(const float4* __restrict__ accum, const __grid_constant__ PeerAccum peers, float4* __restrict__ out, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) out[i] = summed_pixel(accum, peers, i);
}
// The GL path's display pass with its variance-guided denoiser (gl/src/WebGLRendererAdapter.js:183-246, `smartDeNoise` +
// `main` of the passthrough shader), over the HBM-resident buffers: `accum` = the shader's uSampleSumTexture (colour sums;
// w = sample count here), `var` = uVarianceTexture (the running variance sums written by fold_samples_kernel).  FP32 like
// the shader.  Per pixel: a circular Gaussian window of radius round(kSigma * sigma) whose taps are weighted down by their
// OWN standard deviation (`exp(-|std|^2 / (2 threshold^2))`: the shader's active line :210), textures sampled NEAREST with
// coordinates clamped to [0, 1] (gl/src/WebGLUtilHelpers.js:452-453).  The row offsets follow the shader literally:
// d.y runs from -sqrt(r^2 - d.x^2) in steps of 1, so they are not integers off the centre column.
__global__ void denoise_kernel(const float4* __restrict__ accum, const float4* __restrict__ var, float4* __restrict__ out, int W, int H,
                               float sigma, float k_sigma, float threshold, float color_log_scale) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float INV_PI = 0.31830988618379067153776752674503f, INV_SQRT_OF_2PI = 0.39894228040143267793994605993439f;
    const float radius = rintf(k_sigma * sigma), radQ = radius * radius;
    const float invSigmaQx2 = .5f / (sigma * sigma), invSigmaQx2PI = INV_PI * invSigmaQx2;
    const float invThresholdSqx2 = .5f / (threshold * threshold), invThresholdSqrt2PI = INV_SQRT_OF_2PI / threshold;
    const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
    float zBuff = 0.f; float3 aBuff = f3(0.f, 0.f, 0.f); float aW = 0.f;
    for (float dx = -radius; dx <= radius; dx += 1.f) {
        const float pt = sqrtf(radQ - dx * dx);
        for (float dy = -pt; dy <= pt; dy += 1.f) {
            const float blurFactor = expf(-(dx * dx + dy * dy) * invSigmaQx2) * invSigmaQx2PI;
            const float cu = fmaxf(fminf(u + dx / (float)W, 1.0f), 0.0f), cv = fmaxf(fminf(v + dy / (float)H, 1.0f), 0.0f);
            const int tx = min(max((int)floorf(cu * (float)W), 0), W - 1), ty = min(max((int)floorf(cv * (float)H), 0), H - 1);
            const float4 a = accum[ty * W + tx], q = var[ty * W + tx];
            const float tf = 1.0f / fmaxf(a.w, 1.0f);                  // texture_factor = 1 / max(uSampleCount + 1, 1)
            const float3 walkPx = f3(a.x * tf, a.y * tf, a.z * tf);
            const float3 stdPx = f3(sqrtf(q.x * tf), sqrtf(q.y * tf), sqrtf(q.z * tf));
            const float deltaFactor = expf(-dot3(stdPx, stdPx) * invThresholdSqx2) * invThresholdSqrt2PI * blurFactor;
            zBuff += deltaFactor;
            aBuff = aBuff + walkPx * deltaFactor; aW += deltaFactor * (a.w * tf);
        }
    }
    float3 c = aBuff; float alpha = aW;
    if (zBuff != 0.f) { c = c * (1.0f / zBuff); alpha = alpha / zBuff; }
    // the shader's diagnostics colours (:233-238), then the optional log scale (:240-241)
    if (c.x != c.x || c.y != c.y || c.z != c.z) c = f3(1.0f, 0.0f, 0.5f);
    else if (isinf(c.x) || isinf(c.y) || isinf(c.z)) c = f3(0.0f, 1.0f, 0.5f);
    else if (c.x < 0.f || c.y < 0.f || c.z < 0.f) c = f3(0.5f, 0.0f, 1.0f);
    if (color_log_scale > 0.f) c = f3(logf(c.x + 1.0f) / color_log_scale, logf(c.y + 1.0f) / color_log_scale, logf(c.z + 1.0f) / color_log_scale);
    out[y * W + x] = make_float4(c.x, c.y, c.z, alpha);
}
// float RGBA -> 8-bit like the canvas the shader draws to (clamp, round to nearest; alpha 255)
__global__ void quantize_kernel(const float4* __restrict__ in, uchar4* __restrict__ out, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 c = in[i];
    out[i] = make_uchar4((unsigned char)floorf(255.f * fminf(fmaxf(c.x, 0.f), 1.f) + 0.5f), (unsigned char)floorf(255.f * fminf(fmaxf(c.y, 0.f), 1.f) + 0.5f),
                         (unsigned char)floorf(255.f * fminf(fmaxf(c.z, 0.f), 1.f) + 0.5f), 255);
}

// optimistic queue sizing: a batch that did not overflow is folded from its scratch buffer into the pixel sums
__global__ void add_scratch_kernel(float4* __restrict__ accum, float4* __restrict__ scratch, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 b = scratch[i];
    if (b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f) { float4 a = accum[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; accum[i] = a; }
}

// parity probe: hits of the un-jittered pinhole primary rays -> (prim_id, t)
__global__ void hits_to_ids_kernel(const __grid_constant__ DeviceScene sc, const float4* __restrict__ hits, const float4* __restrict__ ro, int n,
                                   int* __restrict__ prim_id, float* __restrict__ tout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = hits[i];
    const int pixel = __float_as_int(ro[i].w), prim = __float_as_int(h.y);
    prim_id[pixel] = prim >= 0 ? sc.prims[prim].ext_id : -1;
    tout[pixel] = h.x;
}

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) throw std::runtime_error(std::string("jsrt: CUDA error: ") + cudaGetErrorString(e__) + " at " #call); } while (0)

int envInt(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

}  // namespace

// ---------------------------------------------------------------------------------
struct Renderer::Impl {
    const HostScene& hs;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    DeviceScene ds{};
    std::vector<void*> allocs;
    float4* accum = nullptr;
    float4* scratch = nullptr;          // W x H float4: radiance of the batch in flight (optimistic sizing) / peer sums (readAccum)
    RayQueue rq[2]{}; float4* hits = nullptr; float4* shadow_hits = nullptr; ShadowQueue sq{};
    float4* sdf_normals = nullptr;
    // JSRT_FLAG_AOV (allocated on first use): first-hit normal / distance sums, variance sums (w: first hits), and the
    // per-sample radiance of the batch in flight (`sample_span` frames)
    float4 *aov_nd = nullptr, *aov_var = nullptr, *sample_rad = nullptr;
    float4* denoised = nullptr;         // output of denoise_kernel (allocated on first use)
    int sample_span = 0;
    int4* tie_list = nullptr;           // FP32 near-ties between triangles reported by bvh_kernel<extend> (trace.cuh: tie_wave)
    static constexpr int kTieCap = 1 << 20;
    int2* work_list = nullptr;          // BVH work list (trace.cuh): extend wave; shadow wave too unless the shadow tests are fused into shade
    Counters* counters = nullptr;
    unsigned long long* stats_backup = nullptr;
    int* overflow = nullptr;
    int ray_cap = 0, shadow_cap = 0, batch = 0;
    bool optimistic = false;            // queues sized for an expected, not the worst, fan-out: every batch is checked and re-run smaller on overflow
    int passes = 0;
    size_t scene_bytes = 0, queue_bytes = 0;
    int grid_extend = 0, grid_extend_gen = 0, grid_shade = 0, grid_shadow = 0, grid_gen = 0, grid_bvh = 0, grid_sdf[2] = {0, 0};
    size_t bvh_smem = 0;
    unsigned long long launches = 0;
    bool profiling = false, has_sdf = false, has_rtu = false, sort_shade = false, fuse_shadow = false, fuse_gen = true;
    int lean = 0;                       // build of shade_kernel the scene can use (leanLevel)
    bool mesh_only = false;             // every BVHAggregate is a pure triangle mesh: bvh_kernel's build without the general-primitive leaves
    int sort_from = 0;                  // first level whose shading is sorted by material (camera rays are coherent as they come)
    double ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};     // 0-3: kernel classes; 4-9: prims / bvh / sdf kernels of extend, shadow
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uchar4* rgba = nullptr; int* hit_ids = nullptr; float* hit_t = nullptr;
    PeerAccum peers{};                  // other GPUs' accumulation buffers (summed by resolve / readAccum)
    std::vector<void*> ipc_mapped;

    explicit Impl(const HostScene& h) : hs(h) {}

    template <class T> T* dalloc(size_t n) { T* p = nullptr; CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); allocs.push_back(p); return p; }
    // First call allocates, later calls (jsrt_scene_upload) re-copy into the same buffers.
    void uploadScene() {
        scene_bytes = 0; up_items.clear();
        bvh_tops_host.clear();
        for (size_t i = 0; i < hs.tops.size(); ++i) if (hs.tops[i].kind == T_BVH && hs.tops[i].node_count > 0) bvh_tops_host.push_back((int)i);
        up(ds.tops, hs.tops); up(ds.prims, hs.prims); up(ds.xforms, hs.xforms); up(ds.xforms64, hs.xforms64); up(ds.nodes, hs.nodes);
        up(ds.tris, hs.tris); up(ds.tri_shade, hs.tri_shade); up(ds.boxes, hs.boxes); up(ds.materials, hs.materials);
        up(ds.lights, hs.lights); up(ds.sdfs, hs.sdfs); up(ds.sdf_code, hs.sdf_code);
        up(ds.textures, hs.textures); up(ds.texels, hs.texels);
        up(ds.bvh_tops, bvh_tops_host); ds.n_bvh = (int)bvh_tops_host.size();
        ds.n_staged = std::min<int>(hs.n_staged, (int)hs.nodes.size());
        // padded world-space box of every BVHAggregate (scene_flatten.cpp: computeWorldBoxes; trace.cuh: wbox_hit)
        {
            std::vector<float> wb;
            computeWorldBoxes(hs, wb);
            wboxes_host.clear();
            for (size_t i = 0; i + 8 <= wb.size(); i += 8) {
                wboxes_host.push_back(make_float4(wb[i], wb[i + 1], wb[i + 2], wb[i + 4]));        // centre | half x
                wboxes_host.push_back(make_float4(wb[i + 5], wb[i + 6], 0.f, 0.f));               // half y, half z
            }
        }
        up(ds.wboxes, wboxes_host);
        ds.tlas_root = (hs.tlas_root >= 0 && hs.tlas_root < (int)hs.nodes.size() && hs.sdfs.empty()) ? hs.tlas_root : -1;   // (SDF scenes keep the linear walk)
        ds.use_wbox = envInt("JSRT_WBOX", (ds.n_bvh >= 2 && ds.tlas_root < 0) ? 1 : 0);
        sdf_tops_host.clear();
        for (size_t i = 0; i < hs.tops.size(); ++i) if (hs.tops[i].kind == T_SDF) sdf_tops_host.push_back((int)i);
        up(ds.sdf_tops, sdf_tops_host); ds.n_sdf_tops = (int)sdf_tops_host.size();
        {
            std::vector<float> wb;
            computeSdfWorldBoxes(hs, wb);
            sdf_wboxes_host.clear();
            for (size_t i = 0; i + 8 <= wb.size(); i += 8) {
                sdf_wboxes_host.push_back(make_float4(wb[i], wb[i + 1], wb[i + 2], wb[i + 4]));
                sdf_wboxes_host.push_back(make_float4(wb[i + 5], wb[i + 6], 0.f, 0.f));
            }
        }
        up(ds.sdf_wboxes, sdf_wboxes_host);
        // analytic-primitive table of prims_wave: Primitives outside BVHAggregates, grouped by geometry kind; the shadow
        // copy leaves out primitives with does_cast_shadow = false (src/world.js:117-118) and geometries that never hit
        atab_host.clear();
        for (int tab = 0; tab < 2; ++tab)
            for (int grp = 0; grp < AG_COUNT; ++grp) {
                for (size_t ti = 0; ti < hs.tops.size(); ++ti) {
                    const Top& t = hs.tops[ti];
                    if (t.kind != T_PRIM && t.kind != T_LIST) continue;
                    for (int k = 0; k < t.prim_count; ++k) {
                        const Prim& p = hs.prims[t.first_prim + k];
                        const int g = p.geom_kind == G_PLANE ? AG_PLANE : p.geom_kind == G_SQUARE ? AG_SQUARE : p.geom_kind == G_BOX ? AG_BOX
                                    : p.geom_kind == G_SPHERE ? AG_SPHERE : AG_OTHER;
                        if (g != grp || p.geom_kind == G_NEVER) continue;
                        if (tab == 1 && !(p.flags & PF_CASTS_SHADOW)) continue;
                        atab_host.push_back(APrim{t.first_prim + k, (int)ti, (t.kind == T_LIST && t.xform != 0) ? t.xform : -1, p.xform, p.geom_index, p.flags, 0, 0});
                    }
                }
                ds.atab_end[tab][grp] = (int)atab_host.size();
            }
        up(ds.atab, atab_host);
        commitUpload();
        ds.n_top = (int)hs.tops.size(); ds.n_lights = (int)hs.lights.size(); ds.light_samples = hs.light_samples; ds.max_depth = hs.max_depth;
        for (int i = 0; i < 3; ++i) ds.bg[i] = hs.bg[i];
    }
    // The scene's arrays live in ONE device allocation and travel in ONE host->device copy from a pinned staging buffer
    // (jsrt_scene_upload re-copies into the same blob): seventeen pageable cudaMemcpyAsync calls were ~0.4 ms of host time
    // per upload, which the end-to-end path pays every step.
    struct UpItem { const void** dst; const void* src; size_t bytes, offset; };
    std::vector<UpItem> up_items;
    unsigned char* scene_blob = nullptr; unsigned char* scene_stage = nullptr; size_t scene_blob_bytes = 0;
    cudaEvent_t upload_done = nullptr;
    std::vector<int> bvh_tops_host, sdf_tops_host;
    std::vector<float4> wboxes_host, sdf_wboxes_host;
    std::vector<APrim> atab_host;
    template <class T> void up(const T*& dst, const std::vector<T>& vec) {
        up_items.push_back(UpItem{(const void**)(void*)&dst, vec.data(), vec.size() * sizeof(T), 0});
        scene_bytes += vec.size() * sizeof(T);
    }
    void commitUpload() {
        size_t total = 0;
        for (UpItem& it : up_items) { it.offset = total; total += (std::max<size_t>(it.bytes, 16) + 255) & ~(size_t)255; }
        if (!scene_blob || total > scene_blob_bytes) {
            if (scene_blob) { CK(cudaStreamSynchronize(stream)); cudaFree(scene_blob); cudaFreeHost(scene_stage); scene_blob = nullptr; scene_stage = nullptr; }
            CK(cudaMalloc(&scene_blob, total)); CK(cudaMallocHost(&scene_stage, total)); scene_blob_bytes = total;
            if (!upload_done) CK(cudaEventCreateWithFlags(&upload_done, cudaEventDisableTiming));
        } else CK(cudaEventSynchronize(upload_done));     // the previous copy out of the staging buffer must be over before it is refilled
        for (const UpItem& it : up_items) if (it.bytes) memcpy(scene_stage + it.offset, it.src, it.bytes);
        CK(cudaMemcpyAsync(scene_blob, scene_stage, total, cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(upload_done, stream));
        for (const UpItem& it : up_items) *it.dst = scene_blob + it.offset;
        up_items.clear();
    }

    // shade_kernel<..., LEAN> (shade.cuh): which build of the kernel the scene can use.  Level 2 — a ground plane + triangle
    // meshes lit by point lights: 7 576 -> ~5 000 SASS instructions, no spills, shade 6.49 -> 5.13 ms per 16 passes of
    // bunny_path (+6 % overall), dragon +3.7 %.  Level 1 — no textures / positional UVs / solid or transparent materials
    // (profiles/r2_ab.md §3).  JSRT_SHADE_LEAN caps the level (0 = the general build).
    int leanLevel() const {
        const int cap = envInt("JSRT_SHADE_LEAN", 2);
        if (cap <= 0 || !hs.sdfs.empty() || !hs.textures.empty()) return 0;
        for (const Material& m : hs.materials) {
            if (m.kind != M_PHONG && m.kind != M_FRESNEL && m.kind != M_PATH) return 0;
            if (m.uv_from_position) return 0;
            for (const Color* c : {&m.ambient, &m.diffusivity, &m.specularity, &m.reflectivity, &m.transmissivity}) if (c->checker == CK_TEXTURE) return 0;
        }
        if (cap < 2) return 1;
        for (const Light& l : hs.lights) if (l.kind != L_POINT) return 1;
        for (const Prim& p : hs.prims) if (p.geom_kind != G_PLANE && p.geom_kind != G_TRIANGLE) return 1;
        return 2;
    }
    void init(int dev, size_t queue_budget) {
        device = dev;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
        stream = own_stream;
        uploadScene();
        has_sdf = !hs.sdfs.empty();
        // Shadow tests fused into shade_kernel: scenes without SDFs (a march cannot run inside the shade loop) whose shadow
        // rays mostly die at a BVH root box and meet few analytic shadow casters.  Measured (profiles/r2_ab.md): bunny_path
        // (1 plane, 2 point lights) shade + shadow 19.5 -> 18.7 ms per 16 passes; cornell_box_path (12 analytic primitives,
        // 4 light samples, no BVH) 21.0 -> 23.7 ms — there the tests run better in prims_kernel<shadow>, one compacted ray
        // per lane at 4 CTAs / SM, than under shade's 80 registers with only the lit lanes of each warp at work.
        const int n_shadow_casters = ds.atab_end[1][AG_COUNT - 1] - ds.atab_end[0][AG_COUNT - 1];
        fuse_shadow = !has_sdf && envInt("JSRT_FUSE_SHADOW", (ds.n_bvh > 0 && n_shadow_casters <= 4) ? 1 : 0) != 0;
        fuse_gen = envInt("JSRT_FUSE_GEN", 1) != 0;
        coalesce = envInt("JSRT_COALESCE", 1) != 0;
        const size_t npix = (size_t)hs.width * hs.height;
        accum = dalloc<float4>(npix);
        CK(cudaMemsetAsync(accum, 0, npix * sizeof(float4), stream));
        rgba = dalloc<uchar4>(npix); hit_ids = dalloc<int>(npix); hit_t = dalloc<float>(npix);
        counters = dalloc<Counters>(1); overflow = dalloc<int>(1); stats_backup = dalloc<unsigned long long>(24);
        CK(cudaMemsetAsync(counters, 0, sizeof(Counters), stream));
        CK(cudaMemsetAsync(overflow, 0, sizeof(int), stream));

        // Queue sizing.  A camera sample owns at most fanout^L rays at level L and light_samples shadow rays per ray, so
        // the worst case over the levels is at the deepest one.  When a useful batch fits the budget at that worst case
        // (every BASELINE config), overflow is impossible by construction and nothing is checked while rendering.
        // Otherwise (fan-out 2 beyond depth ~13: 2^12 rays per sample) the queues are sized for an expected fan-out
        // (JSRT_QUEUE_GROWTH, default 4 rays per camera sample at the widest level — most paths die early, which is why the
        // reference can render such scenes at all) and every batch renders into a scratch buffer that is only folded into
        // the pixel sums if the overflow flag stayed clear; an overflowing batch is re-run with half the samples.
        const int ls = std::max(1, hs.light_samples);
        double worst = 1; for (int l = 1; l < hs.max_depth; ++l) { worst *= hs.fanout; if (worst > 1e12) break; }
        const double ray_bytes = 2.0 * 48 + 16 + (has_sdf ? 16 : 0) + 8;                           // two queues, hit record, (SDF normal), work-list entry
        const double shadow_bytes = fuse_shadow ? 48.0 : 48.0 + 16 + 8;                            // queue (+ partial hit, work-list entry)
        const double per_ray = ray_bytes + shadow_bytes * ls;
        // up to 16 passes per wave (JSRT_BATCH_PASSES overrides): the persistent trace kernels end with a tail of
        // long walks, so bigger waves are faster (bunny_path 1080p: 2.75 / 3.47 / 3.78 / 3.90 Grays/s at 1 / 3.2 / 8 / 16 passes)
        double batch_passes = 16;
        if (const char* e = getenv("JSRT_BATCH_PASSES")) { const double v = atof(e); if (v >= 0.01) batch_passes = v; }
        const double want = std::max(1.0, (double)npix * batch_passes);
        const double min_batch = std::min(want, 65536.0);
        double growth = worst;
        double b = (double)queue_budget / (per_ray * worst);
        if (const char* e = getenv("JSRT_QUEUE_GROWTH")) { const double v = atof(e); if (v >= 1.0 && v < worst) { growth = v; optimistic = true; b = (double)queue_budget / (per_ray * growth); } }
        if (!optimistic && b < min_batch) {
            optimistic = true; growth = std::min(worst, 4.0);
            b = (double)queue_budget / (per_ray * growth);
        }
        b = std::min(std::max(b, min_batch), want);
        // (the floor can exceed a tiny budget: the budget is a target, 65 536 samples in flight are the minimum that keeps 148 SMs busy)
        if (b * growth > 1.0e9) b = 1.0e9 / growth;
        batch = std::max(1, (int)b);
        ray_cap = (int)std::min<double>(std::ceil((double)batch * growth), 2.0e9);
        shadow_cap = (int)std::min<double>((double)ray_cap * ls, 2.0e9);
        for (int k = 0; k < 2; ++k) { rq[k].o = dalloc<float4>(ray_cap); rq[k].d = dalloc<float4>(ray_cap); rq[k].w = dalloc<float4>(ray_cap); }
        hits = dalloc<float4>(ray_cap);
        sq.o = dalloc<float4>(shadow_cap); sq.d = dalloc<float4>(shadow_cap); sq.c = dalloc<float4>(shadow_cap);
        if (!fuse_shadow) shadow_hits = dalloc<float4>(shadow_cap);
        work_list = dalloc<int2>(fuse_shadow ? ray_cap : std::max(ray_cap, shadow_cap));
        tie_list = dalloc<int4>(bvh_tops_host.empty() ? 1 : kTieCap);
        if (has_sdf) { sdf_normals = dalloc<float4>(ray_cap); CK(cudaMemsetAsync(sdf_normals, 0, (size_t)ray_cap * sizeof(float4), stream)); }
        if (optimistic) { scratch = dalloc<float4>(npix); CK(cudaMemsetAsync(scratch, 0, npix * sizeof(float4), stream)); }
        queue_bytes = (size_t)((double)ray_cap * ray_bytes + (double)shadow_cap * shadow_bytes);

        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
        auto grid_for = [&](const void* fn, int block = kBlock, size_t smem = 0) {
            int per = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, block, smem)); return prop.multiProcessorCount * std::max(1, per);
        };
        grid_extend = has_sdf ? grid_for((const void*)prims_kernel<TM_EXTEND, false, true, false>) : grid_for((const void*)prims_kernel<TM_EXTEND, false, false, false>);
        grid_extend_gen = has_sdf ? grid_for((const void*)prims_kernel<TM_EXTEND, false, true, true>) : grid_for((const void*)prims_kernel<TM_EXTEND, false, false, true>);
        // bvh_kernel: dynamic shared memory for the staged top levels (opt-in above 48 KB)
        bvh_smem = (size_t)ds.n_staged * sizeof(BvhNode);
        if ((int)bvh_smem > prop.sharedMemPerBlockOptin) throw std::runtime_error("jsrt: staged BVH block exceeds the shared memory of an SM (lower JSRT_STAGE_NODES)");
        #define JSRT_BVH_ATTR(...) CK(cudaFuncSetAttribute((const void*)bvh_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bvh_smem))
        if (has_sdf) { JSRT_BVH_ATTR(TM_EXTEND, false, true, false, false); JSRT_BVH_ATTR(TM_EXTEND, true, true, false, false); JSRT_BVH_ATTR(TM_SHADOW, false, true, false, false); JSRT_BVH_ATTR(TM_SHADOW, true, true, false, false); }
        else {
            JSRT_BVH_ATTR(TM_EXTEND, false, false, false, false); JSRT_BVH_ATTR(TM_EXTEND, true, false, false, false); JSRT_BVH_ATTR(TM_SHADOW, false, false, false, false); JSRT_BVH_ATTR(TM_SHADOW, true, false, false, false);
            JSRT_BVH_ATTR(TM_SHADOW, false, false, true, false); JSRT_BVH_ATTR(TM_SHADOW, true, false, true, false);
            JSRT_BVH_ATTR(TM_EXTEND, false, false, false, true); JSRT_BVH_ATTR(TM_EXTEND, true, false, false, true); JSRT_BVH_ATTR(TM_SHADOW, false, false, false, true); JSRT_BVH_ATTR(TM_SHADOW, true, false, false, true);
            JSRT_BVH_ATTR(TM_SHADOW, false, false, true, true); JSRT_BVH_ATTR(TM_SHADOW, true, false, true, true);
            // the mesh-only builds (untimed COUNT variants keep the general one)
            JSRT_BVH_ATTR(TM_EXTEND, false, false, false, false, true); JSRT_BVH_ATTR(TM_EXTEND, false, false, false, true, true);
            JSRT_BVH_ATTR(TM_SHADOW, false, false, false, false, true); JSRT_BVH_ATTR(TM_SHADOW, false, false, false, true, true);
            JSRT_BVH_ATTR(TM_SHADOW, false, false, true, false, true); JSRT_BVH_ATTR(TM_SHADOW, false, false, true, true, true);
        }
        #undef JSRT_BVH_ATTR
        grid_bvh = has_sdf ? grid_for((const void*)bvh_kernel<TM_EXTEND, false, true, false, false>, JSRT_BVH_BLOCK, bvh_smem)
                           : grid_for((const void*)bvh_kernel<TM_EXTEND, false, false, false, false>, JSRT_BVH_BLOCK, bvh_smem);
        // material-sorted shading: on for scenes made of analytic primitives only (see shade_kernel)
        sort_shade = bvh_tops_host.empty();
        if (const char* e = getenv("JSRT_SHADE_SORT")) sort_shade = atoi(e) != 0;
        sort_from = envInt("JSRT_SHADE_SORT_FROM", 0);
        lean = leanLevel();
        mesh_only = envInt("JSRT_BVH_MESH", 1) != 0;
        for (const Top& t : hs.tops) if (t.kind == T_BVH && t.node_count > 0 && t.tri_base < 0) mesh_only = false;
        grid_shade = has_sdf ? (sort_shade ? grid_for((const void*)shade_kernel<true, true, false, false>, kShadeBlock) : grid_for((const void*)shade_kernel<true, false, false, false>, kShadeBlock))
                   : fuse_shadow ? (sort_shade ? grid_for((const void*)shade_kernel<false, true, true, false>, kShadeBlock) : grid_for((const void*)shade_kernel<false, false, true, false>, kShadeBlock))
                                 : (sort_shade ? grid_for((const void*)shade_kernel<false, true, false, false>, kShadeBlock) : grid_for((const void*)shade_kernel<false, false, false, false>, kShadeBlock));
        grid_shadow = has_sdf ? grid_for((const void*)prims_kernel<TM_SHADOW, false, true, false>) : grid_for((const void*)prims_kernel<TM_SHADOW, false, false, false>);
        grid_gen = grid_for((const void*)generate_kernel);
        has_rtu = false;
        for (const SdfInstr& in : hs.sdf_code) if (in.op == S_RTU_CROSS) has_rtu = true;
        grid_sdf[TM_EXTEND] = has_rtu ? grid_for((const void*)sdf_kernel<TM_EXTEND, false, true>) : grid_for((const void*)sdf_kernel<TM_EXTEND, false, false>);
        grid_sdf[TM_SHADOW] = has_rtu ? grid_for((const void*)sdf_kernel<TM_SHADOW, false, true>) : grid_for((const void*)sdf_kernel<TM_SHADOW, false, false>);
        CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
        CK(cudaStreamSynchronize(stream));
        setupAccumPersistence(prop);
    }

    // Every radiance term is added to its pixel with an FP32 reduction (RED) in L2.  Between two touches of one
    // pixel the kernels stream hundreds of MB of queue data through L2, so without help the accumulation lines are
    // evicted and every RED turns into a DRAM read-modify-write: measured 2.4 ms of 8.6 (shade) and 1.5 ms of
    // 14.5 (shadow) per 16 passes of bunny_path at 1080p (profiles/r1_ab.md).  The buffer is therefore pinned in
    // the persisting part of L2 (cudaAccessPolicyWindow); the queues stay on the normal (streaming) policy.
    bool accum_persist = false;
    void setupAccumPersistence(const cudaDeviceProp& prop) {
        if (envInt("JSRT_ACCUM_PERSIST", 1) == 0) return;
        const size_t bytes = (size_t)hs.width * hs.height * sizeof(float4);
        if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
        const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
        l2_window.base_ptr = optimistic ? scratch : accum;
        l2_window.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
        l2_window.hitRatio = (float)std::min(1.0, (double)carve / (double)l2_window.num_bytes);
        l2_window.hitProp = cudaAccessPropertyPersisting;
        l2_window.missProp = cudaAccessPropertyStreaming;
        accum_persist = true;
        applyAccumPersistence();
    }
    cudaAccessPolicyWindow l2_window{};
    void applyAccumPersistence() {         // per stream: called again when the host moves the scene to another stream
        if (!accum_persist) return;
        cudaStreamAttrValue v{}; v.accessPolicyWindow = l2_window;
        if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
    }

    ~Impl() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        for (void* p : ipc_mapped) cudaIpcCloseMemHandle(p);
        for (void* p : allocs) cudaFree(p);
        if (scene_blob) cudaFree(scene_blob);
        if (scene_stage) cudaFreeHost(scene_stage);
        if (upload_done) cudaEventDestroy(upload_done);
        for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        for (auto e : free_events) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream) cudaStreamDestroy(own_stream);
    }

    // Per-kernel device time without perturbing the pipeline: event pairs are recorded
    // around each launch on the launching stream and only read back (after a stream
    // sync) when the statistics are fetched.
    struct EvPair { cudaEvent_t a, b; int which; };
    std::vector<EvPair> pending, pending_parts;
    std::vector<cudaEvent_t> free_events;
    unsigned long long kernel_launches[4] = {0, 0, 0, 0};
    cudaEvent_t getEvent() {
        if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
        cudaEvent_t e; CK(cudaEventCreate(&e)); return e;
    }
    void flushEvents() {
        if (pending.empty() && pending_parts.empty()) return;
        CK(cudaStreamSynchronize(stream));
        for (auto& p : pending) { float t = 0; CK(cudaEventElapsedTime(&t, p.a, p.b)); ms[p.which] += t; free_events.push_back(p.a); free_events.push_back(p.b); }
        pending.clear();
        // the per-kernel parts share their boundary events: each event is released once
        for (size_t i = 0; i < pending_parts.size(); ++i) {
            auto& p = pending_parts[i];
            float t = 0; CK(cudaEventElapsedTime(&t, p.a, p.b)); ms[p.which] += t;
            if (i % 3 == 0) free_events.push_back(p.a);
            free_events.push_back(p.b);
        }
        pending_parts.clear();
    }
    template <class F> void timed(int which, F&& launch) {
        EvPair p{nullptr, nullptr, which};
        if (profiling) { p.a = getEvent(); p.b = getEvent(); CK(cudaEventRecord(p.a, stream)); }
        launch();
        ++launches; ++kernel_launches[which];
        CK(cudaGetLastError());
        if (profiling) { CK(cudaEventRecord(p.b, stream)); pending.push_back(p); if (pending.size() >= 8192) flushEvents(); }
    }

    // One trace wave: prims_kernel (analytic primitives + root boxes; `gen`: level 0, camera rays computed in place),
    // bvh_kernel over the work list, tie_kernel, sdf_kernel.  `direct`: the fused shadow wave — the queue holds walkers
    // only, so just bvh_kernel<shadow, DIRECT>.
    template <int MODE> void launchTrace(const TraceIO& io0, bool count_work, int grid_prims, const GenParams* gen, bool direct) {
        TraceIO io = io0;
        const bool has_bvh = ds.n_bvh > 0, has_sdf_tops = ds.n_sdf_tops > 0;
        io.final_pass = has_sdf_tops ? 0 : 1;        // prims_wave finishes the rays that need no BVH walk unless an SDF march follows
        static const GenParams no_gen{};
        const GenParams& gp = gen ? *gen : no_gen;
        // profiling: the three kernels of a trace wave are also timed one by one (ms[4..9])
        const int part0 = 4 + 3 * MODE;
        cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr}; int np = 0;
        auto mark = [&] { if (profiling) { pe[np] = getEvent(); CK(cudaEventRecord(pe[np], stream)); ++np; } };
        mark();
        if (!direct) {
            #define JSRT_PRIMS(C, S, G) prims_kernel<MODE, C, S, G><<<grid_prims, kBlock, 0, stream>>>(ds, io, gp)
            if (gen && MODE == TM_EXTEND) {
                if (count_work) { if (has_sdf) JSRT_PRIMS(true, true, (MODE == TM_EXTEND)); else JSRT_PRIMS(true, false, (MODE == TM_EXTEND)); }
                else { if (has_sdf) JSRT_PRIMS(false, true, (MODE == TM_EXTEND)); else JSRT_PRIMS(false, false, (MODE == TM_EXTEND)); }
            } else {
                if (count_work) { if (has_sdf) JSRT_PRIMS(true, true, false); else JSRT_PRIMS(true, false, false); }
                else { if (has_sdf) JSRT_PRIMS(false, true, false); else JSRT_PRIMS(false, false, false); }
            }
            #undef JSRT_PRIMS
        }
        mark();
        if (has_bvh) {
            ++launches;
            const bool tlas = ds.tlas_root >= 0;
            #define JSRT_BVH(C, S, D, T) bvh_kernel<MODE, C, S, D, T><<<grid_bvh, JSRT_BVH_BLOCK, bvh_smem, stream>>>(ds, io)
            #define JSRT_BVH_MESH(D, T) bvh_kernel<MODE, false, false, D, T, true><<<grid_bvh, JSRT_BVH_BLOCK, bvh_smem, stream>>>(ds, io)
            const bool mesh = mesh_only && !count_work && !has_sdf;
            if (has_sdf) { if (count_work) JSRT_BVH(true, true, false, false); else JSRT_BVH(false, true, false, false); }
            else if (direct) {
                if (tlas) { if (mesh) JSRT_BVH_MESH((MODE == TM_SHADOW), true); else if (count_work) JSRT_BVH(true, false, (MODE == TM_SHADOW), true); else JSRT_BVH(false, false, (MODE == TM_SHADOW), true); }
                else { if (mesh) JSRT_BVH_MESH((MODE == TM_SHADOW), false); else if (count_work) JSRT_BVH(true, false, (MODE == TM_SHADOW), false); else JSRT_BVH(false, false, (MODE == TM_SHADOW), false); }
            } else if (tlas) { if (mesh) JSRT_BVH_MESH(false, true); else if (count_work) JSRT_BVH(true, false, false, true); else JSRT_BVH(false, false, false, true); }
            else { if (mesh) JSRT_BVH_MESH(false, false); else if (count_work) JSRT_BVH(true, false, false, false); else JSRT_BVH(false, false, false, false); }
            #undef JSRT_BVH_MESH
            #undef JSRT_BVH
            if (MODE == TM_EXTEND && JSRT_TRI_TIE) { ++launches; tie_kernel<<<4, kBlock, 0, stream>>>(ds, io); }      // a few dozen entries per frame
        }
        mark();
        if (has_sdf_tops) {
            ++launches;
            io.final_pass = 1;
            io.cursor = io0.cursor + 2 * kCntStride;          // cursor_extend_sdf / cursor_shadow_sdf (two counters further on)
            #define JSRT_SDF(C, R) sdf_kernel<MODE, C, R><<<grid_sdf[MODE], kBlock, 0, stream>>>(ds, io)
            if (has_rtu) { if (count_work) JSRT_SDF(true, true); else JSRT_SDF(false, true); }
            else { if (count_work) JSRT_SDF(true, false); else JSRT_SDF(false, false); }
            #undef JSRT_SDF
        }
        mark();
        if (profiling) {
            for (int k = 0; k < 3; ++k) pending_parts.push_back(EvPair{pe[k], pe[k + 1], part0 + k});
        }
    }
    void launchExtend(int cur, bool count_work, const GenParams* gen) {
        TraceIO io{}; io.o = rq[cur].o; io.d = rq[cur].d; io.hits = hits; io.count = &counters->ray[cur].v; io.cap = ray_cap;
        io.cursor = &counters->cursor_extend.v; io.stats = counters->stats; io.aux = sdf_normals;
        io.tie_list = tie_list; io.tie_count = &counters->ties.v; io.tie_cap = kTieCap;
        io.list = work_list; io.list_count = &counters->list_extend.v;
        timed(1, [&] { launchTrace<TM_EXTEND>(io, count_work, gen ? grid_extend_gen : grid_extend, gen, false); });
    }
    void launchShadow(bool count_work, float4* radiance, int accum_stride, int pass0) {
        TraceIO io{}; io.o = sq.o; io.d = sq.d; io.c = sq.c; io.hits = shadow_hits; io.accum = radiance; io.count = &counters->shadow.v; io.cap = shadow_cap;
        io.accum_stride = accum_stride; io.pass0 = pass0;
        io.cursor = &counters->cursor_shadow.v; io.stats = counters->stats;
        io.list = work_list; io.list_count = &counters->list_shadow.v;
        if (fuse_shadow && ds.n_bvh == 0) return;          // nothing can be queued: every shadow ray was settled in shade_kernel
        timed(3, [&] { launchTrace<TM_SHADOW>(io, count_work, grid_shadow, nullptr, fuse_shadow); });
    }

    void ensureAov(int npix_active) {
        const size_t npix = (size_t)hs.width * hs.height;
        if (!aov_nd) {
            aov_nd = dalloc<float4>(npix); aov_var = dalloc<float4>(npix);
            CK(cudaMemsetAsync(aov_nd, 0, npix * sizeof(float4), stream)); CK(cudaMemsetAsync(aov_var, 0, npix * sizeof(float4), stream));
        }
        const int need = batch / std::max(1, npix_active) + 2;        // frames a batch of camera samples can touch
        if (need > sample_span) {
            // (the old, smaller buffer stays in `allocs` until the scene is destroyed: this happens at most when the striping changes)
            sample_rad = dalloc<float4>(npix * (size_t)need); sample_span = need;
            CK(cudaMemsetAsync(sample_rad, 0, npix * (size_t)need * sizeof(float4), stream));
        }
    }

    GenParams genParams(int first_pass, uint64_t seed, int x_offset, int x_delt, int flags) const {
        GenParams g{};
        g.cam = hs.camera; g.width = hs.width; g.height = hs.height;
        g.x_offset = x_offset; g.x_delt = x_delt < 1 ? 1 : x_delt;
        g.ncols = x_offset < hs.width ? (hs.width - x_offset + g.x_delt - 1) / g.x_delt : 0;
        g.npix_active = g.ncols * hs.height;
        g.first_pass = first_pass; g.jitter = (flags & 1) ? 0 : 1; g.max_depth = hs.max_depth; g.seed = seed;
        g.use_lens = 1; g.count_samples = 1;
        g.qo = rq[0].o; g.qd = rq[0].d; g.qw = rq[0].w; g.accum = accum;
        return g;
    }

    // one batch of camera samples through all levels; radiance terms go to `radiance`
    void runBatch(GenParams& g, uint64_t seed, bool count_work, bool aov, float4* radiance, int rstride, int pass0) {
        set_count_kernel<<<1, 1, 0, stream>>>(counters, 0, g.n_samples, 1); ++launches;
        if (!fuse_gen) timed(0, [&] { generate_kernel<<<grid_gen, kBlock, 0, stream>>>(g); });
        int cur = 0;
        for (int level = 0; level < hs.max_depth; ++level) {
            launchExtend(cur, count_work, (level == 0 && fuse_gen) ? &g : nullptr);
            timed(2, [&] {
                ShadeIO io{};
                io.q = rq[cur]; io.count = &counters->ray[cur].v; io.hits = hits; io.next = rq[cur ^ 1]; io.next_count = &counters->ray[cur ^ 1].v; io.next_cap = ray_cap;
                io.sq = sq; io.shadow_count = &counters->shadow.v; io.shadow_cap = shadow_cap; io.accum = radiance; io.seed = seed; io.stats = counters->stats;
                io.overflow = overflow; io.sdf_normals = has_sdf ? sdf_normals : nullptr; io.accum_stride = rstride; io.pass0 = pass0;
                io.aov_nd = aov ? aov_nd : nullptr; io.aov_var = aov ? aov_var : nullptr;
                const bool sort_now = sort_shade && level >= sort_from;
                #define JSRT_SHADE(S, O, F, C) shade_kernel<S, O, F, C><<<grid_shade, kShadeBlock, 0, stream>>>(ds, io)
                #define JSRT_SHADE_L(O, F, L) shade_kernel<false, O, F, false, L><<<grid_shade, kShadeBlock, 0, stream>>>(ds, io)
                if (has_sdf) { if (count_work) JSRT_SHADE(true, false, false, true); else if (sort_now) JSRT_SHADE(true, true, false, false); else JSRT_SHADE(true, false, false, false); }
                else if (count_work) { if (fuse_shadow) JSRT_SHADE(false, false, true, true); else JSRT_SHADE(false, false, false, true); }
                else if (lean >= 2 && !sort_now) { if (fuse_shadow) JSRT_SHADE_L(false, true, 2); else JSRT_SHADE_L(false, false, 2); }
                else if (lean >= 1) {
                    if (fuse_shadow) { if (sort_now) JSRT_SHADE_L(true, true, 1); else JSRT_SHADE_L(false, true, 1); }
                    else { if (sort_now) JSRT_SHADE_L(true, false, 1); else JSRT_SHADE_L(false, false, 1); }
                } else if (fuse_shadow) { if (sort_now) JSRT_SHADE(false, true, true, false); else JSRT_SHADE(false, false, true, false); }
                else { if (sort_now) JSRT_SHADE(false, true, false, false); else JSRT_SHADE(false, false, false, false); }
                #undef JSRT_SHADE_L
                #undef JSRT_SHADE
            });
            if (hs.light_samples > 0) launchShadow(count_work, radiance, rstride, pass0);
            level_end_kernel<<<1, 1, 0, stream>>>(counters, cur, level, ray_cap, shadow_cap, fuse_shadow ? 1 : 0); ++launches;
            cur ^= 1;
        }
    }

    // Submission coalescing.  jsrt_render is asynchronous, and small waves are slow (the persistent trace kernels end with a
    // tail of long walks: 2.75 / 3.9 Grays/s at 1 / 16 passes per wave on bunny_path), so consecutive calls that continue
    // each other — same seed, striping and flags, first_pass = the previous call's end — are held back and rendered as one
    // wave once a full batch is pending or anything observes the scene (synchronize, resolve, read-back, statistics,
    // upload, reset).  A host that hands over one pass per call (the reference's per-pass loop, src/renderers.js:87; a
    // strong-scaling split that leaves each GPU two passes per step) gets big waves anyway.  JSRT_COALESCE=0 disables.
    struct Pending { bool any = false; int first = 0, n = 0; uint64_t seed = 0; int x_offset = 0, x_delt = 1, flags = 0; } pend;
    bool coalesce = true;
    void submit(int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags) {
        if (hs.max_depth <= 0 || hs.max_depth > 255) throw std::runtime_error("jsrt: maxRecursionDepth must be in 1..255");
        if (x_offset < 0) throw std::runtime_error("jsrt: x_offset must be >= 0");
        if (x_delt < 1) x_delt = 1;
        if (pend.any && !(seed == pend.seed && x_offset == pend.x_offset && x_delt == pend.x_delt && flags == pend.flags && first_pass == pend.first + pend.n)) flush();
        if (!pend.any) { pend.any = true; pend.first = first_pass; pend.n = 0; pend.seed = seed; pend.x_offset = x_offset; pend.x_delt = x_delt; pend.flags = flags; }
        pend.n += n_passes;
        if (x_offset == 0 && x_delt == 1) passes += n_passes; else passes = std::max(passes, first_pass + n_passes);
        const long long ncols = x_offset < hs.width ? (hs.width - x_offset + x_delt - 1) / x_delt : 0;
        if (!coalesce || (flags & 6) || ncols * hs.height * (long long)pend.n >= (long long)batch) flush();
    }
    void flush() {
        if (!pend.any) return;
        pend.any = false;
        renderNow(pend.first, pend.n, pend.seed, pend.x_offset, pend.x_delt, pend.flags);
    }
    void renderNow(int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags) {
        CK(cudaSetDevice(device));
        GenParams g = genParams(first_pass, seed, x_offset, x_delt, flags);
        const bool count_work = (flags & 2) != 0;
        const bool aov = (flags & 4) != 0;
        const int npix = hs.width * hs.height;
        if (aov && optimistic) throw std::runtime_error("jsrt: JSRT_FLAG_AOV needs worst-case queue sizing (this scene's depth x fan-out does not fit JSRT_QUEUE_BYTES)");
        if (aov) ensureAov(g.npix_active);
        const long long total = (long long)g.npix_active * n_passes;
        int cur_batch = batch;
        for (long long done = 0; done < total;) {
            g.first_sample = done;
            g.n_samples = (int)std::min<long long>(cur_batch, total - done);
            // AOV renders: radiance terms land in the per-sample buffer, folded into the pixel sums after the batch
            const int pass0 = first_pass + (int)(done / std::max(1, g.npix_active));
            const int span = aov ? first_pass + (int)((done + g.n_samples - 1) / std::max(1, g.npix_active)) - pass0 + 1 : 0;
            if (!optimistic) {
                runBatch(g, seed, count_work, aov, aov ? sample_rad : accum, aov ? npix : 0, pass0);
                if (aov) { fold_samples_kernel<<<grid_gen, kBlock, 0, stream>>>(g, pass0, span, sample_rad, accum, aov_var, npix); ++launches; }
                done += g.n_samples;
                continue;
            }
            // optimistic sizing: render into the scratch buffer, check, fold in or re-run smaller
            CK(cudaMemcpyAsync(stats_backup, counters->stats, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToDevice, stream));
            g.accum = scratch;
            runBatch(g, seed, count_work, false, scratch, 0, pass0);
            int ov = 0;
            CK(cudaMemcpyAsync(&ov, overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (!ov) {
                add_scratch_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accum, scratch, npix); ++launches;
                CK(cudaMemsetAsync(scratch, 0, (size_t)npix * sizeof(float4), stream));
                done += g.n_samples;
            } else {
                CK(cudaMemsetAsync(overflow, 0, sizeof(int), stream));
                CK(cudaMemsetAsync(scratch, 0, (size_t)npix * sizeof(float4), stream));
                CK(cudaMemcpyAsync(counters->stats, stats_backup, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToDevice, stream));
                if (g.n_samples <= 256) throw std::runtime_error("jsrt: wavefront queues overflow even with 256 camera samples in flight; raise JSRT_QUEUE_BYTES");
                cur_batch = std::max(256, g.n_samples / 2);
            }
        }
        CK(cudaGetLastError());
    }

    // peers' work must be finished before their buffers are read: the caller orders that (events / barriers)
    void sumPeersInto(float4* dst) {
        const int npix = hs.width * hs.height;
        sum_peers_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accum, peers, dst, npix); ++launches;
    }
};

Renderer::Renderer(const HostScene& hs, int device, size_t queue_budget) : impl_(new Impl(hs)) { impl_->init(device, queue_budget); }
Renderer::~Renderer() { delete impl_; }
void Renderer::render(int first_pass, int n_passes, uint64_t seed, int x_offset, int x_delt, int flags) { impl_->submit(first_pass, n_passes, seed, x_offset, x_delt, flags); }
void Renderer::flushPending() { impl_->flush(); }
void Renderer::upload() { impl_->flush(); CK(cudaSetDevice(impl_->device)); impl_->uploadScene(); CK(cudaStreamSynchronize(impl_->stream)); }
void Renderer::setStream(void* s) { impl_->flush(); impl_->stream = s ? (cudaStream_t)s : impl_->own_stream; impl_->applyAccumPersistence(); }
void* Renderer::stream() const { return (void*)impl_->stream; }
int Renderer::device() const { return impl_->device; }
void Renderer::synchronize() {
    impl_->flush();
    CK(cudaSetDevice(impl_->device)); CK(cudaStreamSynchronize(impl_->stream));
    int ov = 0; CK(cudaMemcpy(&ov, impl_->overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (ov) {
        // reported once: the flag is cleared so that the handle stays usable (jsrt_reset_accum + a smaller JSRT_BATCH_PASSES)
        CK(cudaMemset(impl_->overflow, 0, sizeof(int)));
        throw std::runtime_error("jsrt: wavefront queue overflow (internal sizing error): the image of this call is incomplete");
    }
}
void Renderer::resetAccum() {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    CK(cudaMemsetAsync(impl_->accum, 0, (size_t)impl_->hs.width * impl_->hs.height * sizeof(float4), impl_->stream));
    CK(cudaMemsetAsync(impl_->overflow, 0, sizeof(int), impl_->stream));
    impl_->passes = 0;
    if (impl_->aov_nd) {
        const size_t bytes = (size_t)impl_->hs.width * impl_->hs.height * sizeof(float4);
        CK(cudaMemsetAsync(impl_->aov_nd, 0, bytes, impl_->stream)); CK(cudaMemsetAsync(impl_->aov_var, 0, bytes, impl_->stream));
    }
}
void Renderer::readAov(float* normal_depth, float* variance) {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    const size_t bytes = (size_t)impl_->hs.width * impl_->hs.height * sizeof(float4);
    if (!impl_->aov_nd) { memset(normal_depth, 0, bytes); memset(variance, 0, bytes); return; }
    CK(cudaMemcpyAsync(normal_depth, impl_->aov_nd, bytes, cudaMemcpyDeviceToHost, impl_->stream));
    CK(cudaMemcpyAsync(variance, impl_->aov_var, bytes, cudaMemcpyDeviceToHost, impl_->stream));
    synchronize();
}
void Renderer::denoise(float sigma, float k_sigma, float threshold, float color_log_scale, float* out_rgba, uint8_t* out_rgba8) {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    Impl& m = *impl_;
    if (!m.aov_var) throw std::runtime_error("jsrt: the denoiser needs the variance buffer: render the passes with JSRT_FLAG_AOV");
    if (!(sigma > 0.f) || !(k_sigma >= 0.f) || !(threshold > 0.f) || k_sigma * sigma > 64.f) throw std::runtime_error("jsrt: denoise parameters out of range (sigma > 0, kSigma >= 0, kSigma * sigma <= 64, threshold > 0)");
    const int W = m.hs.width, H = m.hs.height, npix = W * H;
    if (!m.denoised) m.denoised = m.dalloc<float4>(npix);
    const dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8);
    denoise_kernel<<<grd, blk, 0, m.stream>>>(m.accum, m.aov_var, m.denoised, W, H, sigma, k_sigma, threshold, color_log_scale); ++m.launches;
    if (out_rgba) CK(cudaMemcpyAsync(out_rgba, m.denoised, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToHost, m.stream));
    if (out_rgba8) {
        quantize_kernel<<<(npix + 255) / 256, 256, 0, m.stream>>>(m.denoised, m.rgba, npix); ++m.launches;
        CK(cudaMemcpyAsync(out_rgba8, m.rgba, (size_t)npix * 4, cudaMemcpyDeviceToHost, m.stream));
    }
    synchronize();
}
void Renderer::resolve(uint8_t* out) {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    const int npix = impl_->hs.width * impl_->hs.height;
    resolve_kernel<<<(npix + 255) / 256, 256, 0, impl_->stream>>>(impl_->accum, impl_->peers, impl_->rgba, npix); ++impl_->launches;
    CK(cudaMemcpyAsync(out, impl_->rgba, (size_t)npix * 4, cudaMemcpyDeviceToHost, impl_->stream));
    synchronize();
}
void Renderer::readAccum(float* out, int* passes) {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    Impl& m = *impl_;
    const size_t npix = (size_t)m.hs.width * m.hs.height;
    const float4* src = m.accum;
    if (m.peers.n > 0) {
        if (!m.scratch) { m.scratch = m.dalloc<float4>(npix); }
        m.sumPeersInto(m.scratch); src = m.scratch;
    }
    CK(cudaMemcpyAsync(out, src, npix * sizeof(float4), cudaMemcpyDeviceToHost, m.stream));
    synchronize();
    if (m.peers.n > 0 && m.optimistic) CK(cudaMemsetAsync(m.scratch, 0, npix * sizeof(float4), m.stream));
    if (passes) *passes = m.passes;
}
void* Renderer::accumPtr() { return impl_->accum; }
void Renderer::addPasses(int n) { impl_->passes += n; }
void Renderer::setPeers(const void* const* ptrs, int n) {
    if (n < 0 || n > kMaxPeers) throw std::runtime_error("jsrt: too many peer accumulation buffers");
    impl_->peers.n = n;
    for (int i = 0; i < n; ++i) impl_->peers.p[i] = (const float4*)ptrs[i];
}
void Renderer::exportAccum(void* handle64) {
    CK(cudaSetDevice(impl_->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, impl_->accum));
    memcpy(handle64, &h, 64);
}
void Renderer::attachAccum(const void* handles64, int n) {
    CK(cudaSetDevice(impl_->device));
    Impl& m = *impl_;
    for (void* p : m.ipc_mapped) cudaIpcCloseMemHandle(p);
    m.ipc_mapped.clear(); m.peers.n = 0;
    if (n > kMaxPeers) throw std::runtime_error("jsrt: too many peer accumulation buffers");
    for (int i = 0; i < n; ++i) {
        cudaIpcMemHandle_t h; memcpy(&h, (const char*)handles64 + 64 * (size_t)i, 64);
        void* p = nullptr; CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m.ipc_mapped.push_back(p); m.peers.p[m.peers.n++] = (const float4*)p;
    }
}
void Renderer::primaryHits(int32_t* prim_id, float* t) {
    impl_->flush();
    CK(cudaSetDevice(impl_->device));
    Impl& m = *impl_;
    GenParams g = m.genParams(0, 0, 0, 1, 1);
    g.use_lens = 0; g.count_samples = 0;
    const int npix = m.hs.width * m.hs.height;
    for (int done = 0; done < npix; done += m.batch) {
        g.first_sample = done; g.n_samples = std::min(m.batch, npix - done);
        set_count_kernel<<<1, 1, 0, m.stream>>>(m.counters, 0, g.n_samples, 0); ++m.launches;
        generate_kernel<<<m.grid_gen, kBlock, 0, m.stream>>>(g); ++m.launches;
        m.launchExtend(0, false, nullptr);
        hits_to_ids_kernel<<<(g.n_samples + 255) / 256, 256, 0, m.stream>>>(m.ds, m.hits, m.rq[0].o, g.n_samples, m.hit_ids, m.hit_t); ++m.launches;
        set_count_kernel<<<1, 1, 0, m.stream>>>(m.counters, 0, 0, 0); ++m.launches;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(prim_id, m.hit_ids, (size_t)npix * 4, cudaMemcpyDeviceToHost, m.stream));
    CK(cudaMemcpyAsync(t, m.hit_t, (size_t)npix * 4, cudaMemcpyDeviceToHost, m.stream));
    synchronize();
}
void Renderer::getStats(RenderStats& s) {
    synchronize();
    impl_->flushEvents();
    Counters c; CK(cudaMemcpy(&c, impl_->counters, sizeof(Counters), cudaMemcpyDeviceToHost));
    s.rays_primary = c.stats[ST_PRIMARY]; s.rays_secondary = c.stats[ST_SECONDARY]; s.rays_shadow = c.stats[ST_SHADOW];
    s.shaded_hits = c.stats[ST_SHADED]; s.camera_samples = c.stats[ST_SAMPLES]; s.launches = impl_->launches;
    for (int k = 0; k < 3; ++k) { s.nodes[k] = c.stats[ST_NODES + k]; s.leaf_prims[k] = c.stats[ST_LEAF_PRIMS + k]; s.top_prims[k] = c.stats[ST_TOP_PRIMS + k]; s.sdf_evals[k] = c.stats[ST_SDF_EVALS + k]; }
    for (int i = 0; i < 4; ++i) { s.ms[i] = impl_->ms[i]; s.kernel_launches[i] = impl_->kernel_launches[i]; }
    for (int i = 0; i < 6; ++i) s.ms_part[i] = impl_->ms[4 + i];
}
void Renderer::resetStats() {
    synchronize();
    CK(cudaMemsetAsync(impl_->counters, 0, sizeof(Counters), impl_->stream));
    impl_->flushEvents();
    impl_->launches = 0; for (double& m : impl_->ms) m = 0; for (auto& k : impl_->kernel_launches) k = 0;
}
void Renderer::setProfiling(bool on) { impl_->flush(); impl_->profiling = on; }
int Renderer::passes() const { return impl_->passes; }
int Renderer::batchSamples() const { return impl_->batch; }
size_t Renderer::sceneBytes() const { return impl_->scene_bytes; }
size_t Renderer::queueBytes() const { return impl_->queue_bytes; }

__global__ void __launch_bounds__(256) read_bw_kernel(const float4* __restrict__ buf, size_t n, int iters, float4* sink) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const float4 v = buf[i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    if (acc.x == 1234567.f) *sink = acc;      // never true: keeps the loads alive
}
double measureReadBandwidth(int device, size_t bytes, int iters) {
    CK(cudaSetDevice(device));
    const size_t n = bytes / sizeof(float4);
    float4 *buf = nullptr, *sink = nullptr;
    CK(cudaMalloc(&buf, n * sizeof(float4))); CK(cudaMalloc(&sink, sizeof(float4)));
    CK(cudaMemset(buf, 0, n * sizeof(float4)));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    read_bw_kernel<<<grid, 256>>>(buf, n, 1, sink);                  // warm: brings the buffer into L2
    CK(cudaEventRecord(a));
    read_bw_kernel<<<grid, 256>>>(buf, n, iters, sink);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf); cudaFree(sink);
    CK(cudaGetLastError());
    return (double)n * sizeof(float4) * iters / (ms * 1e-3) / 1e9;
}

int deviceCount() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

void enablePeerAccess(int a, int b) {
    if (a == b) return;
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, a, b));
    if (!can) throw std::runtime_error("jsrt: CUDA devices " + std::to_string(a) + " and " + std::to_string(b) + " are not peers (no NVLink / P2P path)");
    CK(cudaSetDevice(a));
    const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CK(e);
}
void orderAfter(Renderer& consumer, Renderer& producer) {
    producer.flushPending();
    cudaEvent_t ev;
    CK(cudaSetDevice(producer.device()));
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, (cudaStream_t)producer.stream()));
    CK(cudaSetDevice(consumer.device()));
    CK(cudaStreamWaitEvent((cudaStream_t)consumer.stream(), ev, 0));
    CK(cudaEventDestroy(ev));          // released once the wait has consumed it
}

}  // namespace jsrt
