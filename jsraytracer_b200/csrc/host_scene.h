// Host-side product of flattening the serializer graph (scene_flatten.cpp).
#pragma once
#include <string>
#include <vector>

#include "scene_types.h"
#include "wire.h"

namespace jsrt {

struct HostScene {
    int width = 0, height = 0, samples_per_pixel = 1, max_depth = 3;
    bool jitter = true;                 // false for SimpleRenderer (src/renderers.js:21-25)
    std::string renderer_type;
    Camera camera{};
    float bg[4] = {0, 0, 0, 0};

    std::vector<Top> tops;              // flattened top-level entries in the reference's visit order (scene_flatten.cpp)
    std::vector<int> top_world;         // parallel to tops: the entry of world.objects each one came from
    int world_object_count = 0;
    int tlas_root = -1;                 // root of the top-level BVH over the BVHAggregates (buildTlas), -1: linear walk
    int n_staged = 0;                   // nodes[0, n_staged): the top levels of every tree (staged in shared memory by bvh_kernel)
    std::vector<Prim> prims;
    std::vector<Xform> xforms;          // xforms[0] is the identity
    std::vector<Xform64> xforms64;      // parallel to xforms
    std::vector<BvhNode> nodes;
    std::vector<Tri> tris;
    std::vector<TriShade> tri_shade;    // parallel to tris when any triangle has vertex data, else empty
    std::vector<float> boxes;           // 8 floats per non-unit AABB geometry: centre xyz_, half xyz_
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<unsigned char> texels;  // RGBA8, all textures back to back (16-byte aligned starts)
    std::vector<Light> lights;
    std::vector<SdfProgram> sdfs;
    std::vector<SdfInstr> sdf_code;

    int ext_prim_count = 0;             // number of distinct reference Primitives (prim_id range)
    int light_samples = 0;              // sum of samples over lights = shadow rays per shaded hit
    int fanout = 1;                     // 2 if any material can spawn both a reflection and a transmission child
    int max_bvh_depth = 0;
    int tree_node_count = 0;            // BVH nodes over the distinct trees (one layout each)
};

void flattenScene(const WireDoc& doc, HostScene& out);

// Padded world-space box of every BVHAggregate with nodes, in world.objects order: 8 floats each (centre xyz, 0, half xyz, 0).
// The device tests a ray against it before paying for the aggregate's ray transform and local root-box test
// (trace.cuh: wbox_hit); it must contain the aggregate's root box mapped to world space.
void computeWorldBoxes(const HostScene& hs, std::vector<float>& out);
void computeSdfWorldBoxes(const HostScene& hs, std::vector<float>& out);     // the same for the top-level SDF primitives

// Appends a top-level BVH over the scene's BVHAggregates to hs.nodes (leaves = aggregate ordinals in world.objects order)
// and returns its root index, or -1 if it cannot be built (an unbounded aggregate).
int buildTlas(HostScene& hs);

// SDF tree -> bytecode (sdf_compile.cpp).  Returns the index of the new program.
int compileSdf(const WireDoc& doc, const Val* sdf_geometry, HostScene& out);

}  // namespace jsrt
