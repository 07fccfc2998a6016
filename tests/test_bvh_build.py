"""BVH build restatement (src/aggregates.js:65-185): the literal Python mirror and the native builder
must produce the same tree, and the trees must have the shape the reference's would."""
import math
import os

import numpy as np
import pytest

from jsraytracer_b200 import bvh_native, scenes
from jsraytracer_b200.jsmath import Vec, Mat4
from jsraytracer_b200.geometry import Triangle, Sphere
from jsraytracer_b200.materials import PhongMaterial
from jsraytracer_b200.serializer import Serializer
from jsraytracer_b200.world import BVHAggregate, BVHAggregateNode, Primitive


def _tree_signature(node, objects):
    """Canonical nested description: (depth, aabb, leaf object indices | (lesser, greater))."""
    idx = {id(o): i for i, o in enumerate(objects)}
    out, stack = [], [node]
    while stack:
        n = stack.pop()
        box = tuple(n.aabb.center.v[:3]) + tuple(n.aabb.half_size.v[:3])
        if n.isLeaf:
            out.append((n.depth, box, tuple(idx[id(o)] for o in n.objects)))
        else:
            out.append((n.depth, box, None))
            stack.append(n.lesser_node)
            stack.append(n.greater_node)
    return out


def _both(objects):
    py = BVHAggregateNode.build(list(objects), 0, math.inf, 1)
    for o in objects:       # drop cached boxes so the native path recomputes them its own way
        o.aabb = None
    nat = bvh_native.build_tree(list(objects), math.inf, 1)
    return py, nat


@pytest.mark.parametrize("mesh", ["teapot", "hollow_tetrahedron", "star", "cube"])
def test_native_builder_matches_literal_restatement(mesh):
    mat = PhongMaterial(Vec.of(1, 1, 1))
    tris = scenes.load_mesh(mesh, mat, Mat4.translation([0.3, -1, 2]).times(Mat4.rotation(0.4, Vec.of(0, 1, 0))).times(Mat4.scale(0.7)))
    py, nat = _both(tris)
    assert _tree_signature(py, tris) == _tree_signature(nat, tris)
    assert py.nodeCount() == 2 * len(tris) - 1


def test_degenerate_inputs_take_the_fallback_paths():
    """Identical boxes (median fallback fails -> halving) and pairs sharing a centre (two triangles of a
    quad: the length-2 median reads one past the end of the JS array)."""
    mat = PhongMaterial(Vec.of(1, 1, 1))
    rng = np.random.default_rng(3)
    objs = []
    for _ in range(7):                      # seven copies of one sphere: nothing can split them
        objs.append(Primitive(Sphere(), mat, Mat4.translation([1, 2, 3])))
    for k in range(12):                     # quads: two triangles with the same AABB
        x, y = float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))
        a, b, c, d = Vec.of(x, y, 0, 1), Vec.of(x + 1, y, 0, 1), Vec.of(x + 1, y + 1, 0, 1), Vec.of(x, y + 1, 0, 1)
        objs.append(Primitive(Triangle([a, b, c]), mat))
        objs.append(Primitive(Triangle([a, c, d]), mat))
    py, nat = _both(objs)
    assert _tree_signature(py, objs) == _tree_signature(nat, objs)
    assert py.nodeCount() == 2 * len(objs) - 1


def test_bunny_tree_shape_and_wire_identity():
    """bunny2.obj: 4 968 triangles -> 9 935 nodes, depth 15 (SURVEY.md §8d scratch estimate); the wire
    blob built from the native tree is byte-identical to the one from the literal Python build."""
    mat = PhongMaterial(Vec.of(1, 1, 1))
    tris = scenes.load_mesh("bunny2", mat)
    assert len(tris) == 4968
    nat = BVHAggregate.build(tris, Mat4.translation([-0.6, 1, -4]), native=True)
    assert nat.nodeCount() == 9935 and nat.maxDepth() == 15
    if os.environ.get("JSRT_SLOW_TESTS", "1") != "0":
        for o in tris:
            o.aabb = None
        py = BVHAggregate.build(tris, Mat4.translation([-0.6, 1, -4]), native=False)
        for o in tris:
            o.aabb = None       # the native path never caches per-object boxes on the primitives
        assert Serializer(py).to_json() == Serializer(nat).to_json()


def test_dragon_tree_shape():
    """dragon.obj: 100 000 faces, 99 968 pass minArea = 1e-5 (src/objloader.js:203,240) -> 199 935 nodes, depth 24."""
    test = scenes.configure("dragon", width=8, height=8)
    bvh = test["renderer"].world.objects[1]
    assert len(bvh.objects) == 99968
    assert bvh.nodeCount() == 199935 and bvh.maxDepth() == 24


@pytest.mark.parametrize("name,kw", [("starwars", {}), ("AMultipleBVH", {}), ("AHollowTetrahedron", {}), ("x_wing", {}), ("bottle", {})])
def test_world_boxes_contain_the_aggregates(name, kw):
    """The padded world-space box the device tests in front of each BVHAggregate (csrc/scene_flatten.cpp:
    computeWorldBoxes) must contain every vertex of the aggregate mapped to world space — otherwise the reject would
    drop hits the reference's walk over world.objects (src/world.js:7-15, src/aggregates.js:43-46) finds — and should
    not be much larger than the tight box of those vertices."""
    from jsraytracer_b200 import lib, scenes
    from jsraytracer_b200.serializer import Serializer
    from jsraytracer_b200.world import BVHAggregate
    test = scenes.configure(name, width=16, height=16, **kw)
    boxes = lib.Scene(Serializer(test).to_msgpack(), lib.FORMAT_MSGPACK, device=None).bvh_world_boxes()
    aggs = [o for o in test["renderer"].world.objects if isinstance(o, BVHAggregate)]
    assert len(boxes) == len(aggs) >= 1
    for (c, h), agg in zip(boxes.astype(np.float64), aggs):
        M = np.array(agg.transform.rows, dtype=np.float64)
        pts = np.array([[float(x) for x in p.v[:3]] + [1.0] for prim in agg.objects for p in prim.geometry.ps], dtype=np.float64)
        # triangles of an OBJ carry the loader's transform themselves (identity in these scenes); the aggregate's maps them to the world
        pts = pts @ np.array(agg.objects[0].transform.rows, dtype=np.float64).T
        w = (pts @ M.T)[:, :3]
        lo, hi = w.min(axis=0), w.max(axis=0)
        assert np.all(c - h <= lo) and np.all(c + h >= hi), (name, c, h, lo, hi)
        assert np.all(h <= 0.5 * (hi - lo) * 1.75 + 1e-2 * (1 + np.abs(c).max())), (name, h, hi - lo)     # an AABB of a rotated AABB: bounded slack
