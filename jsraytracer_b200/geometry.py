"""Scene-graph mirror of the reference geometry classes (reference: src/geometry.js).

Only the *construction-time* surface lives here: constructors, bounding boxes
(needed by `BVHAggregate.build`, src/aggregates.js:34-42, and by
`SDFGeometry`, src/sdf.js:6) and the serialisation shape each class has on the
wire (src/serializer.js:12-60).  `intersect` / `materialData` / `sampleSurface`
are evaluated by the CUDA kernels, never here.
"""
from __future__ import annotations

import math

from .jsmath import Vec, Mat, Mat4, _jsdiv

INF = math.inf


class JSObject:
    """Base for mirrored classes.  `JS_NAME` is `obj.constructor.name`
    (src/serializer.js:38-41); `js_items()` yields the own enumerable
    properties in constructor-assignment order (src/serializer.js:54-57)."""

    JS_NAME = "Object"

    def js_items(self):
        return list(self.__dict__.items())


class Geometry(JSObject):
    pass


class AABB(Geometry):  # src/geometry.js:77-228
    JS_NAME = "AABB"

    def __init__(self, center, half_size, min_=None, max_=None):
        self.center = center
        self.half_size = half_size
        self.min = min_ if min_ is not None else center.minus(half_size)
        self.max = max_ if max_ is not None else center.plus(half_size)

    def serialize(self, ser):  # src/geometry.js:85-87 (min/max are not emitted)
        return {"center": ser.serialize_step(self.center), "half_size": ser.serialize_step(self.half_size)}

    @staticmethod
    def empty():
        return AABB(Vec.of(0, 0, 0, 1), Vec.of(0, 0, 0, 0), Vec.of(INF, INF, INF, 1), Vec.of(-INF, -INF, -INF, 1))

    @staticmethod
    def fromMinMax(mn, mx):  # src/geometry.js:94-105
        center = mn.mix(mx, 0.5)
        half_size = mx.minus(mn).times(0.5)
        for i in range(3):
            if not math.isfinite(mn[i]) and not math.isfinite(mx[i]):
                if mn[i] == mx[i]:
                    half_size[i] = 0
                if mn[i] == -INF and mx[i] == INF:
                    center[i] = 0
        return AABB(center, half_size, mn, mx)

    @staticmethod
    def fromPoints(points):  # src/geometry.js:106-120
        if len(points) == 0:
            return AABB.empty()
        mn = [INF, INF, INF, 1.0]
        mx = [-INF, -INF, -INF, 1.0]
        for p in points:
            for i in range(3):
                if p[i] < mn[i]:
                    mn[i] = p[i]
                if p[i] > mx[i]:
                    mx[i] = p[i]
        return AABB.fromMinMax(Vec(mn), Vec(mx))

    @staticmethod
    def hull(boxes):  # src/geometry.js:121-135
        if len(boxes) == 0:
            return AABB.empty()
        mn = [INF, INF, INF, 1.0]
        mx = [-INF, -INF, -INF, 1.0]
        for b in boxes:
            bmin, bmax = b.min, b.max
            for i in range(3):
                if bmin[i] < mn[i]:
                    mn[i] = bmin[i]
                if bmax[i] > mx[i]:
                    mx[i] = bmax[i]
        return AABB.fromMinMax(Vec(mn), Vec(mx))

    @staticmethod
    def intersection(boxes):  # src/geometry.js:136-151
        mn = [-INF, -INF, -INF, 1.0]
        mx = [INF, INF, INF, 1.0]
        for b in boxes:
            for i in range(3):
                if b.min[i] > mn[i]:
                    mn[i] = b.min[i]
                if b.max[i] < mx[i]:
                    mx[i] = b.max[i]
        for i in range(3):
            if mn[i] > mx[i]:
                return None
        return AABB.fromMinMax(Vec(mn), Vec(mx))

    @staticmethod
    def infinite():
        return AABB.fromMinMax(Vec.of(-INF, -INF, -INF, 1), Vec.of(INF, INF, INF, 1))

    def surfaceArea(self):  # src/geometry.js:160-164
        h = self.half_size
        return 4 * (h[0] * h[1] + h[0] * h[2] + h[1] * h[2])

    def isFinite(self):
        return all(math.isfinite(self.min[i]) and math.isfinite(self.max[i]) for i in range(3))

    def expand(self, amount):  # src/geometry.js:186-188
        return AABB(self.center, self.half_size.plus(amount))

    def getCorners(self):  # src/geometry.js:165-172
        a, b = self.min, self.max
        return [Vec(c) for c in (
            [a[0], a[1], a[2], 1], [b[0], a[1], a[2], 1],
            [a[0], b[1], a[2], 1], [b[0], b[1], a[2], 1],
            [a[0], a[1], b[2], 1], [b[0], a[1], b[2], 1],
            [a[0], b[1], b[2], 1], [b[0], b[1], b[2], 1])]

    def getBoundingBox(self, transform, inv_transform=None):  # src/geometry.js:225-227
        return AABB.fromPoints([transform.times(c) for c in self.getCorners()])


class UnitBox(AABB):  # src/geometry.js:230-237
    JS_NAME = "UnitBox"

    def __init__(self):
        super().__init__(Vec.of(0, 0, 0, 1), Vec.of(0.5, 0.5, 0.5, 0))


class SimplePlane(Geometry):  # src/geometry.js:239-255
    JS_NAME = "SimplePlane"

    def js_items(self):
        return []


class Plane(SimplePlane):  # src/geometry.js:257-278
    JS_NAME = "Plane"

    def getBoundingBox(self, transform, inv_transform):
        normal = inv_transform.transposed().times(Vec.of(0, 0, 1, 0)).normalized()
        s = Vec.of(INF, INF, INF, 0)
        for i in range(3):
            found = False
            for j in range(3):
                if not found and i != j and normal[j] != 0:
                    found = True
            if not found:
                s[i] = 0
        p = transform.times(Vec.of(0, 0, 0, 1))
        return AABB(p, s, p.minus(s), p.plus(s.to4(True)))


class Square(SimplePlane):  # src/geometry.js:280-301
    JS_NAME = "Square"

    def getBoundingBox(self, transform, inv_transform):
        pts = [Vec.of(a, b, 0, 1) for a, b in ((-0.5, -0.5), (0.5, -0.5), (-0.5, 0.5), (0.5, 0.5))]
        return AABB.fromPoints([transform.times(p) for p in pts])


class Circle(SimplePlane):  # src/geometry.js:303-332
    JS_NAME = "Circle"

    @staticmethod
    def getTransformedEdgePoints(transform, inv_transform):
        world_axis = transform.times(Vec.axis(2, 4))
        world_center = transform.column(3)
        ps = []
        for i in range(3):
            edge = transform.times(inv_transform.times(world_axis.cross(Vec.axis(i, 4)).to4(False)).normalized())
            ps.append(world_center.plus(edge))
            ps.append(world_center.minus(edge))
        return ps

    def getBoundingBox(self, transform, inv_transform):
        return AABB.fromPoints(Circle.getTransformedEdgePoints(transform, inv_transform))


class Triangle(Geometry):  # src/geometry.js:334-410
    JS_NAME = "Triangle"

    def __init__(self, ps, psdata=None):
        self.ps = ps
        self.psdata = psdata if psdata is not None else {}
        self.v0 = ps[1].minus(ps[0]).to3()
        self.v1 = ps[2].minus(ps[0]).to3()
        heron = self.v0.cross(self.v1)
        self.area = heron.norm() / 2.0
        self.normal = heron.normalized().to4(False)
        self.delta = self.normal.dot(ps[0])
        self.d00 = self.v0.squarednorm()
        self.d11 = self.v1.squarednorm()
        self.d01 = self.v0.dot(self.v1)
        self.denom = self.d00 * self.d11 - self.d01 * self.d01

    def serialize(self, ser):
        """The reference writes `psdata: serializeStep(this.ps)`
        (src/geometry.js:355-357), which drops vertex normals / UVs that the
        worker path keeps.  The drop-in glue overrides it with the intended
        `{ps, psdata}` (SURVEY.md §8b hazard 1); set
        `Serializer(reference_bug_compat=True)` to get the lossy original."""
        if ser.reference_bug_compat:
            return {"ps": ser.serialize_step(self.ps), "psdata": ser.serialize_step(self.ps)}
        return {"ps": ser.serialize_step(self.ps), "psdata": ser.serialize_step(self.psdata)}

    def getBoundingBox(self, transform, inv_transform=None):  # src/geometry.js:386-388
        return AABB.fromPoints([transform.times(p) for p in self.ps])


class Sphere(Geometry):  # src/geometry.js:412-456
    JS_NAME = "Sphere"

    def js_items(self):
        return []

    def getBoundingBox(self, transform, inv_transform=None):
        return Sphere.computeBoundingBox(transform, inv_transform)

    @staticmethod
    def computeBoundingBox(transform, inv_transform=None):  # src/geometry.js:422-428
        c = transform.column(3)
        h = Vec.of(0, 0, 0, 0)
        tt = transform.transposed()
        for i in range(3):
            h[i] = transform.times(tt.times(Vec.axis(i, 4)).to4(False).normalized().to4(True))[i] - c[i]
        return AABB(c, h)


class Cylinder(Geometry):  # src/geometry.js:458-488
    JS_NAME = "Cylinder"

    def js_items(self):
        return []

    def getBoundingBox(self, transform, inv_transform):
        axis = transform.times(Vec.axis(2, 4))
        pts = Circle.getTransformedEdgePoints(transform, inv_transform)
        out = []
        for v in pts:
            out.append(v.plus(axis))
            out.append(v.minus(axis))
        return AABB.fromPoints(out)
