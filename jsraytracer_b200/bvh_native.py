"""Bridge from the scene-graph mirror to the native BVH builder
(`jsrt_bvh_build`, csrc/bvh_build.cpp — the same restatement of
src/aggregates.js:65-185 as world.BVHAggregateNode.build, for large meshes)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .jsmath import Vec
from .geometry import AABB, Triangle


def object_boxes(objects):
    """center / half_size / min / max (n,3) f32 of `o.getBoundingBox()` for every
    object; Triangle primitives are vectorised (AABB.fromPoints on the
    transformed vertices, src/geometry.js:386-388,94-120), the rest go through
    the mirror."""
    n = len(objects)
    cen = np.zeros((n, 3), np.float32)
    half = np.zeros((n, 3), np.float32)
    mn = np.zeros((n, 3), np.float32)
    mx = np.zeros((n, 3), np.float32)
    tri_idx = [i for i, o in enumerate(objects) if isinstance(getattr(o, "geometry", None), Triangle)
               and getattr(o, "aabb", None) is None]
    groups = {}
    for i in tri_idx:
        groups.setdefault(id(objects[i].transform), []).append(i)
    for idxs in groups.values():
        T = np.array(objects[idxs[0]].transform.rows, dtype=np.float64)
        P = np.array([[p.v for p in objects[i].geometry.ps] for i in idxs], dtype=np.float64)   # (m,3,4) f32 values
        # transform.times(p): f64 dot, left-to-right, stored f32
        W = np.empty(P.shape[:2] + (3,), dtype=np.float64)
        for r in range(3):
            W[..., r] = ((P[..., 0] * T[r, 0] + P[..., 1] * T[r, 1]) + P[..., 2] * T[r, 2]) + P[..., 3] * T[r, 3]
        W = W.astype(np.float32)
        lo, hi = W.min(axis=1), W.max(axis=1)
        lo64, hi64 = lo.astype(np.float64), hi.astype(np.float64)
        ii = np.array(idxs)
        mn[ii], mx[ii] = lo, hi
        cen[ii] = ((1 - 0.5) * lo64 + 0.5 * hi64).astype(np.float32)
        half[ii] = ((hi64 - lo64).astype(np.float32).astype(np.float64) * 0.5).astype(np.float32)
    done = set(tri_idx)
    for i, o in enumerate(objects):
        if i in done:
            continue
        b = o.getBoundingBox()
        cen[i], half[i], mn[i], mx[i] = b.center.v[:3], b.half_size.v[:3], b.min.v[:3], b.max.v[:3]
    return cen, half, mn, mx


def build_arrays(cen, half, mn, mx, maxDepth=float("inf"), minNodeSize=1):
    L = lib.load()
    n = len(cen)
    cen, half, mn, mx = (np.ascontiguousarray(a, dtype=np.float32) for a in (cen, half, mn, mx))
    h = L.jsrt_bvh_build(n, cen.ctypes.data, half.ctypes.data, mn.ctypes.data, mx.ctypes.data, float(maxDepth),
                         int(minNodeSize))
    try:
        nn = L.jsrt_bvh_node_count(h)
        nodes = (lib.BvhNode * nn)()
        leaf = np.empty(L.jsrt_bvh_leaf_object_count(h), dtype=np.int32)
        L.jsrt_bvh_copy(h, nodes, leaf.ctypes.data)
    finally:
        L.jsrt_bvh_free(h)
    return nodes, leaf


def build_tree(objects, maxDepth=float("inf"), minNodeSize=1):
    """Returns the root `BVHAggregateNode` of the tree the reference would build."""
    from .world import BVHAggregateNode
    cen, half, mn, mx = object_boxes(objects)
    nodes, leaf = build_arrays(cen, half, mn, mx, maxDepth, minNodeSize)
    built = [None] * len(nodes)
    # children have larger indices than their parent (pre-order emission)
    for i in range(len(nodes) - 1, -1, -1):
        nd = nodes[i]
        box = AABB(Vec(list(nd.center)), Vec(list(nd.half_size)), Vec(list(nd.min)), Vec(list(nd.max)))
        if nd.is_leaf:
            objs = [objects[j] for j in leaf[nd.obj_first:nd.obj_first + nd.obj_count]]
            built[i] = BVHAggregateNode(nd.depth, True, objs, box, None, None)
        else:
            built[i] = BVHAggregateNode(nd.depth, False, [], box, built[nd.lesser], built[nd.greater])
    return built[0]
