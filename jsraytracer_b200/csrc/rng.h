// Counter-based RNG standing in for Math.random() (reference:
// src/renderers.js:95-96, src/math.js:175-188, src/materials.js:399,408,
// src/geometry.js:296-298).  Keyed by (seed, pixel, pass, path-tree node,
// dimension) so that a wavefront — which cannot follow the reference's
// depth-first draw order — is still reproducible; the test oracle implements
// the same functions (oracle/oracle_math.h) for per-sample comparison.
// Path-tree node ids: 1 = camera ray, 2k = reflection child of k, 2k+1 =
// transmission / refraction child of k.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define JSRT_HD __host__ __device__ __forceinline__
#else
#define JSRT_HD inline
#endif

namespace jsrt {

JSRT_HD uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
JSRT_HD uint32_t rng_sample_key(uint64_t seed, uint32_t pixel, uint32_t pass) {
    uint32_t k = hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + 0x68bc21ebU));
    k = hash32(k + pixel);
    k = hash32(k ^ hash32(pass + 0x9e3779b9U));
    return k;
}
JSRT_HD uint32_t rng_node_key(uint32_t sample_key, uint32_t node) { return hash32(sample_key + 0x85ebca6bU * node); }
// path-tree node id of child `which` (0 reflection, 1 transmission / refraction): 2k + which while that fits
// (levels <= 30), a hash with the top bit set beyond (never 0 or 1, never back in the doubling range), so that
// maxRecursionDepth up to 255 keeps distinct RNG keys and never re-creates the camera ray's id
JSRT_HD uint32_t rng_child_node(uint32_t node, uint32_t which) {
    return (node < 0x40000000U) ? 2U * node + which : (hash32(node ^ (0x9e3779b9U + which)) | 0x80000000U);
}
// U[0,1) with 24 bits: exactly representable in f32 and f64
JSRT_HD float rng_u01(uint32_t node_key, uint32_t dim) {
    return (float)(hash32(node_key + 0xc2b2ae35U * (dim + 1)) >> 8) * (1.0f / 16777216.0f);
}

// dimension assignment within a node
enum : uint32_t {
    DIM_JITTER_X = 0, DIM_JITTER_Y = 1,   // src/renderers.js:95-96 (node 1 only)
    DIM_LENS_A = 2, DIM_LENS_R = 3,       // Vec.circlePick, src/math.js:176-177 (node 1 only)
    DIM_LIGHTS = 8                        // 2 per light sample, in world.lights order; then 4 for the
                                          // reflection scatter and 4 for the refraction scatter
};

}  // namespace jsrt
