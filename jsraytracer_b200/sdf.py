"""Scene-graph mirror of the SDF operator tree (reference: src/sdf.js).

Construction-time surface only: the tree shape, the constants every node
derives in its constructor, and `getBoundingBox` (needed for
`SDFGeometry.aabb`, src/sdf.js:6).  `distance()` is compiled to bytecode by
the C-ABI library and marched by the CUDA register-stack interpreter.

Every SDF / transformer node owns an *enumerable* `UID` (src/sdf.js:55,372),
so it travels on the wire; the reader ignores it.
"""
from __future__ import annotations

import math

from .jsmath import Vec, Mat, Mat4, _jsdiv
from .geometry import AABB, Geometry, Sphere, UnitBox, JSObject, INF

_SDF_UID_GEN = [0]


def _next_uid():
    _SDF_UID_GEN[0] += 1
    return _SDF_UID_GEN[0]


class SDFGeometry(Geometry):  # src/sdf.js:1-51
    JS_NAME = "SDFGeometry"

    def __init__(self, root_sdf, max_samples=1000, distance_epsilon=0.0001, max_trace_distance=1000,
                 normal_step_size=0.001):
        self.root_sdf = root_sdf
        self.aabb = root_sdf.getBoundingBox(Mat4.identity(), Mat4.identity())
        self.max_samples = max_samples
        self.distance_epsilon = distance_epsilon
        self.max_trace_distance = max_trace_distance
        self.normal_step_size = normal_step_size

    def getBoundingBox(self, transform, inv_transform):
        return self.root_sdf.getBoundingBox(transform, inv_transform)


class SDF(JSObject):
    def __init__(self):
        self.UID = _next_uid()


class UnionSDF(SDF):  # src/sdf.js:78-92
    JS_NAME = "UnionSDF"

    def __init__(self, *children):
        super().__init__()
        self.children = list(children)

    def getBoundingBox(self, t, it):
        return AABB.hull([c.getBoundingBox(t, it) for c in self.children])


class IntersectionSDF(SDF):  # src/sdf.js:94-108
    JS_NAME = "IntersectionSDF"

    def __init__(self, *children):
        super().__init__()
        self.children = list(children)

    def getBoundingBox(self, t, it):
        return AABB.intersection([c.getBoundingBox(t, it) for c in self.children])


class DifferenceSDF(SDF):  # src/sdf.js:110-125
    JS_NAME = "DifferenceSDF"

    def __init__(self, positive, negative):
        super().__init__()
        self.positive = positive
        self.negative = negative

    def getBoundingBox(self, t, it):
        return self.positive.getBoundingBox(t, it)


class SmoothUnionSDF(SDF):  # src/sdf.js:139-158
    JS_NAME = "SmoothUnionSDF"

    def __init__(self, childA, childB, k=1):
        super().__init__()
        self.k = k
        self.childA = childA
        self.childB = childB

    def getBoundingBox(self, t, it):
        return AABB.hull([self.childA.getBoundingBox(t, it), self.childB.getBoundingBox(t, it)]).expand(self.k / 6)


class SmoothIntersectionSDF(SDF):  # src/sdf.js:160-179
    JS_NAME = "SmoothIntersectionSDF"

    def __init__(self, childA, childB, k=1):
        super().__init__()
        self.k = k
        self.childA = childA
        self.childB = childB

    def getBoundingBox(self, t, it):
        return AABB.intersection([self.childA.getBoundingBox(t, it), self.childB.getBoundingBox(t, it)])


class SmoothDifferenceSDF(SDF):  # src/sdf.js:181-200
    JS_NAME = "SmoothDifferenceSDF"

    def __init__(self, positive, negative, k=1):
        super().__init__()
        self.k = k
        self.positive = positive
        self.negative = negative

    def getBoundingBox(self, t, it):
        return self.positive.getBoundingBox(t, it)


class RoundSDF(SDF):  # src/sdf.js:204-219
    JS_NAME = "RoundSDF"

    def __init__(self, child_sdf, rounding):
        super().__init__()
        self.child_sdf = child_sdf
        self.rounding = rounding

    def getBoundingBox(self, t, it):
        return self.child_sdf.getBoundingBox(t, it).expand(self.rounding)


class SphereSDF(SDF):  # src/sdf.js:226-244
    JS_NAME = "SphereSDF"

    def __init__(self, radius=1, basecolor=None):
        super().__init__()
        self.radius = radius
        self.basecolor = basecolor if basecolor is not None else Vec.of(1, 1, 1)

    def getBoundingBox(self, t, it):
        # src/sdf.js:242 passes ONE argument (the second product is an ignored
        # extra argument to Mat.times); computeBoundingBox only needs the first.
        return Sphere.computeBoundingBox(t.times(Mat4.scale(self.radius)))


class BoxSDF(SDF):  # src/sdf.js:266-293
    JS_NAME = "BoxSDF"

    def __init__(self, size=0.5, basecolor=None):
        super().__init__()
        if isinstance(size, Vec):
            self.size = size.to4(False)
        else:
            self.size = Vec.of(size, size, size, 0)
        self.basecolor = basecolor if basecolor is not None else Vec.of(1, 1, 1)

    def getBoundingBox(self, t, it):
        scale_vec = self.size
        return UnitBox().getBoundingBox(t.times(Mat4.scale(scale_vec.times(2))),
                                        Mat4.scale(scale_vec.inverse(2)).times(it))


class TetrahedronSDF(SDF):  # src/sdf.js:295-318
    JS_NAME = "TetrahedronSDF"
    vertices = [Vec.of(1, 1, 1, 0), Vec.of(-1, 1, -1, 0), Vec.of(-1, -1, 1, 0), Vec.of(1, -1, -1, 0)]

    def __init__(self, basecolor=None):
        super().__init__()
        self.basecolor = basecolor if basecolor is not None else Vec.of(1, 1, 1)

    def getBoundingBox(self, t, it):
        return AABB.fromPoints([t.times(v) for v in TetrahedronSDF.vertices])


class TransformSDF(SDF):  # src/sdf.js:324-340
    JS_NAME = "TransformSDF"

    def __init__(self, child_sdf, transformer):
        super().__init__()
        self.child_sdf = child_sdf
        self.transformer = transformer

    def getBoundingBox(self, t, it):
        return self.transformer.transformBoundingBox(self.child_sdf.getBoundingBox(t, it))


class RecursiveTransformUnionSDF(SDF):  # src/sdf.js:342-368
    JS_NAME = "RecursiveTransformUnionSDF"

    def __init__(self, sdf, transformer, iterations):
        super().__init__()
        self.sdf = sdf
        self.transformer = transformer
        self.iterations = iterations

    def getBoundingBox(self, t, it):
        aabb = self.sdf.getBoundingBox(t, it)
        for _ in range(self.iterations):
            aabb = AABB.hull([aabb, self.transformer.transformBoundingBox(aabb)])
        return aabb


class SDFTransformer(JSObject):
    def __init__(self):
        self.UID = _next_uid()


class SDFTransformerSequence(SDFTransformer):  # src/sdf.js:382-400
    JS_NAME = "SDFTransformerSequence"

    def __init__(self, *transformers):
        super().__init__()
        self.transformers = list(transformers)

    def transformBoundingBox(self, aabb):
        for t in self.transformers:
            aabb = t.transformBoundingBox(aabb)
        return aabb


class SDFRecursiveTransformer(SDFTransformer):  # src/sdf.js:402-421
    JS_NAME = "SDFRecursiveTransformer"

    def __init__(self, transformer, iterations):
        super().__init__()
        self.transformer = transformer
        self.iterations = iterations

    def transformBoundingBox(self, aabb):
        for _ in range(self.iterations):
            aabb = self.transformer.transformBoundingBox(aabb)
        return aabb


class SDFMatrixTransformer(SDFTransformer):  # src/sdf.js:423-439
    JS_NAME = "SDFMatrixTransformer"

    def __init__(self, transform, inv_transform=None):
        super().__init__()
        self._transform = transform
        self._inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)
        self._scale = min(transform.column(i).to3().norm() for i in range(3))

    def transformBoundingBox(self, aabb):
        return aabb.getBoundingBox(self._transform, self._inv_transform)


class SDFReflectionTransformer(SDFTransformer):  # src/sdf.js:441-464
    JS_NAME = "SDFReflectionTransformer"

    def __init__(self, normal, delta):
        super().__init__()
        normal = normal.to4(False)
        if isinstance(delta, Vec):
            delta = normal.dot(delta)
        self.delta = _jsdiv(delta, normal.norm())
        self.normal = normal.normalized()

    @staticmethod
    def transformComp(p, normal, delta):
        dot = normal.dot(p) - delta
        if dot < 0:
            return p.minus(normal.times(2 * dot))
        return p

    def transformBoundingBox(self, aabb):
        corners = aabb.getCorners()
        neg = self.normal.times(-1)
        corners = corners + [SDFReflectionTransformer.transformComp(c, neg, -self.delta) for c in corners]
        return AABB.fromPoints(corners)


class SDFInfiniteRepetitionTransformer(SDFTransformer):  # src/sdf.js:466-477
    JS_NAME = "SDFInfiniteRepetitionTransformer"

    def __init__(self, sizes=None):
        super().__init__()
        self.sizes = sizes if sizes is not None else Vec.of(1, 1, 1)

    def transformBoundingBox(self, aabb):
        return AABB(Vec.of(0, 0, 0, 1), Vec.of(INF, INF, INF, 0))
