"""Scene-graph mirror: World, WorldObject, Primitive, Aggregate, BVHAggregate
(reference: src/world.js, src/aggregates.js).

`BVHAggregateNode.build` / `split_objects` restate src/aggregates.js:65-185
literally (8-bin SAH on f32 centre/half-size boxes, median fallback, halving
fallback) so that the tree *topology* — which decides hit-ID ties and the
algorithmic node/triangle counts of the roofline — is the one the reference
would build.  The build is setup, not hot path; for large meshes the same
algorithm runs natively (`jsrt_bvh_build`, csrc/bvh_build.cpp) and
tests/test_bvh_build.py checks the two agree node for node.
"""
from __future__ import annotations

import math

from .jsmath import Vec, Mat, Mat4, median, _jsdiv
from .geometry import AABB, JSObject, INF


class World(JSObject):  # src/world.js:1-42
    JS_NAME = "World"

    def __init__(self, objects, lights=None, bg_color=None):
        self.bg_color = bg_color if bg_color is not None else Vec.of(0, 0, 0)
        self.objects = objects
        self.lights = lights if lights is not None else []


class WorldObject(JSObject):  # src/world.js:44-80
    def __init__(self, transform, inv_transform):
        self.transform = transform
        self.inv_transform = inv_transform

    def getBoundingBox(self):
        if getattr(self, "aabb", None) is None:
            self.aabb = self.buildBoundingBox()
        return self.aabb

    def getTransform(self):
        return self.transform

    def getInvTransform(self):
        return self.inv_transform

    def setTransform(self, transform, inv_transform=None):
        self.transform = transform
        self.inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)
        self.aabb = None


class Primitive(WorldObject):  # src/world.js:104-141
    JS_NAME = "Primitive"

    def __init__(self, geometry, material, transform=None, inv_transform=None, does_cast_shadow=True):
        transform = transform if transform is not None else Mat4.identity()
        inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)
        super().__init__(transform, inv_transform)
        self.geometry = geometry
        self.material = material
        self.does_cast_shadow = does_cast_shadow

    def buildBoundingBox(self):
        return self.geometry.getBoundingBox(self.transform, self.inv_transform)


class Aggregate(WorldObject):  # src/aggregates.js:1-19
    JS_NAME = "Aggregate"

    def __init__(self, objects=None, transform=None, inv_transform=None):
        transform = transform if transform is not None else Mat4.identity()
        inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)
        super().__init__(transform, inv_transform)
        self.objects = objects if objects is not None else []

    def buildBoundingBox(self):
        if not self.objects:
            return AABB.empty()
        return AABB.hull([o.getBoundingBox().getBoundingBox(self.transform, self.inv_transform) for o in self.objects])


class BVHAggregate(Aggregate):  # src/aggregates.js:26-61
    JS_NAME = "BVHAggregate"

    def __init__(self, objects, kdtree, transform=None, inv_transform=None):
        super().__init__(objects, transform, inv_transform)
        self.kdtree = kdtree

    @staticmethod
    def build(objects, transform=None, maxDepth=INF, minNodeSize=1, inv_transform=None, native=None):
        """src/aggregates.js:34-42.  `native=None` picks the native builder for
        large inputs when the library is available; both produce the same tree."""
        transform = transform if transform is not None else Mat4.identity()
        inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)
        objects = list(objects)
        for o in objects:
            if not o.getBoundingBox().isFinite():
                raise ValueError("Infinite objects not allowed in")
        if native is None:
            import os
            native = len(objects) > 2000 and not os.environ.get("JSRT_PY_BVH")     # JSRT_PY_BVH=1: never touch the native library
        if native:
            from . import bvh_native
            tree = bvh_native.build_tree(objects, maxDepth, minNodeSize)
        else:
            tree = BVHAggregateNode.build(objects, 0, maxDepth, minNodeSize)
        return BVHAggregate(objects, tree, transform, inv_transform)

    def buildBoundingBox(self):
        if not self.objects:
            return AABB.empty()
        return self.kdtree.aabb.getBoundingBox(self.transform, self.inv_transform)

    def maxDepth(self):
        return self.kdtree.maxDepth()

    def nodeCount(self):
        return self.kdtree.nodeCount()


class BVHAggregateNode(JSObject):  # src/aggregates.js:63-232
    JS_NAME = "BVHAggregateNode"

    def __init__(self, depth, isLeaf, objects, aabb, lesser_node, greater_node):
        self.depth = depth
        self.isLeaf = isLeaf
        if isLeaf:
            self.objects = objects
        self.aabb = aabb
        self.lesser_node = lesser_node
        self.greater_node = greater_node
        if self.aabb is None:
            raise ValueError("Empty aabb for BVH node")

    @staticmethod
    def _leaf(objects, depth):
        aabb = AABB.hull([o.getBoundingBox() for o in objects]) if objects else AABB.empty()
        return BVHAggregateNode(depth, True, objects, aabb, None, None)

    @staticmethod
    def build(objects, depth, maxDepth, minNodeSize):  # src/aggregates.js:65-86
        # Iterative form of the reference's recursion (Python's recursion limit
        # is far below the reference tree depths); children are created in the
        # same order (lesser first, then greater), which only matters for UIDs.
        root_slot = [None]
        stack = [(objects, depth, root_slot, 0)]
        while stack:
            objs, d, parent, which = stack.pop()
            if d >= maxDepth or len(objs) <= minNodeSize:
                node = BVHAggregateNode._leaf(objs, d)
            else:
                split = BVHAggregateNode.split_objects(objs)
                if split:
                    node = BVHAggregateNode(d, False, [], split["bounds"], _PENDING, _PENDING)
                    stack.append((split["greater_objs"], d + 1, node, 2))
                    stack.append((split["lesser_objs"], d + 1, node, 1))
                else:
                    node = BVHAggregateNode._leaf(objs, d)
            if which == 0:
                parent[0] = node
            elif which == 1:
                parent.lesser_node = node
            else:
                parent.greater_node = node
        return root_slot[0]

    @staticmethod
    def split_objects(objects, binsPerAxis=8):  # src/aggregates.js:87-185
        if len(objects) < 2:
            return None
        boxes = [o.getBoundingBox() for o in objects]
        bounds = AABB.hull(boxes)

        best_axis, best_sep_value, best_cost = -1, INF, INF
        bsa = bounds.surfaceArea()
        for axis in range(3):
            if bounds.half_size[axis] < 0.000001:
                continue
            counts = [0] * binsPerAxis
            bbounds = [AABB.empty() for _ in range(binsPerAxis)]
            bmin = bounds.min[axis]
            ext = 2 * bounds.half_size[axis]
            for b in boxes:
                bin_index = math.floor(binsPerAxis * _jsdiv(b.center[axis] - bmin, ext))
                if bin_index == binsPerAxis:
                    bin_index = binsPerAxis - 1
                # (an index outside [0, bins) would throw in the reference;
                # it cannot happen for finite boxes inside their own hull)
                counts[bin_index] += 1
                bbounds[bin_index] = AABB.hull([bbounds[bin_index], b])
            for i in range(binsPerAxis - 1):
                b0, b1 = AABB.empty(), AABB.empty()
                count0 = count1 = 0
                for j in range(i + 1):
                    if counts[j] > 0:
                        b0 = AABB.hull([b0, bbounds[j]])
                        count0 += counts[j]
                for j in range(i + 1, binsPerAxis):
                    if counts[j] > 0:
                        b1 = AABB.hull([b1, bbounds[j]])
                        count1 += counts[j]
                cost = .125 + _jsdiv(count0 * b0.surfaceArea() + count1 * b1.surfaceArea(), bsa)
                if cost < best_cost and count0 > 0 and count1 > 0:
                    best_axis = axis
                    best_sep_value = bmin + ((i + 1) / binsPerAxis) * ext
                    best_cost = cost

        if best_axis < 0:
            for axis in range(3):
                median_axis_val = median([a.center[axis] for a in boxes])
                aabbs0 = [a for a in boxes if a.center[axis] < median_axis_val]
                aabbs1 = [a for a in boxes if a.center[axis] >= median_axis_val]
                cost = .125 + _jsdiv(len(aabbs0) * AABB.hull(aabbs0).surfaceArea()
                                     + len(aabbs1) * AABB.hull(aabbs1).surfaceArea(), bsa)
                if cost < best_cost and len(aabbs0) > 0 and len(aabbs1) > 0:
                    best_axis = axis
                    best_sep_value = median_axis_val
                    best_cost = cost
            if best_axis < 0:
                split = len(objects) // 2
                return {"bounds": bounds, "sep_axis": 0, "sep_value": boxes[0].center[0],
                        "lesser_objs": objects[:split], "greater_objs": objects[split:]}

        return {"bounds": bounds, "sep_axis": best_axis, "sep_value": best_sep_value,
                "lesser_objs": [o for o, b in zip(objects, boxes) if b.center[best_axis] < best_sep_value],
                "greater_objs": [o for o, b in zip(objects, boxes) if b.center[best_axis] >= best_sep_value]}

    def maxDepth(self):
        best, stack = 0, [self]
        while stack:
            n = stack.pop()
            if n.isLeaf:
                best = max(best, n.depth)
            else:
                stack.extend((n.greater_node, n.lesser_node))
        return best

    def nodeCount(self):
        count, stack = 0, [self]
        while stack:
            n = stack.pop()
            count += 1
            if not n.isLeaf:
                stack.extend((n.greater_node, n.lesser_node))
        return count


_PENDING = object()
