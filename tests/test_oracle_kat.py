"""Known-answer tests pinning the CPU restatement oracle to the reference's *formulas*
(SURVEY.md §8c iv): each expectation is derived by hand from the cited source lines, not from
running the oracle.  The reference has no fixtures of its own and cannot run here, so this is
the only pin available ("parity unpinned" otherwise)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle as O

INF = math.inf


def isect(kind, o, d, minD=0.0, maxD=INF, extra=None):
    L = O.lib()
    oo = np.array(list(o) + [1.0], dtype=np.float64)
    dd = np.array(list(d) + [0.0], dtype=np.float64)
    ex = np.array(extra if extra is not None else [0.0] * 12, dtype=np.float64)
    return L.orc_kat_intersect(kind, oo.ctypes.data, dd.ctypes.data, minD, maxD, ex.ctypes.data)


def mdata(kind, pos, extra=None):
    L = O.lib()
    pp = np.array(list(pos) + [1.0], dtype=np.float64)
    ex = np.array(extra if extra is not None else [0.0] * 12, dtype=np.float64)
    out = np.zeros(5)
    assert L.orc_kat_material_data(kind, pp.ctypes.data, ex.ctypes.data, out.ctypes.data) == 0
    return out


PLANE, SQUARE, CIRCLE, BOX, SPHERE, CYL, TRI = range(7)
TRI_PS = [0, 0, 0, 1, 1, 0, 0, 1, 0, 1, 0, 1]


def test_plane_square_circle():                      # src/geometry.js:246-248,287-291,310-314
    assert isect(PLANE, (0, 0, 2), (0, 0, -1)) == 2.0
    assert isect(PLANE, (0, 0, 2), (1, 0, 0)) == -INF          # direction.z == 0
    assert isect(PLANE, (0, 0, -3), (0, 0, -1)) == -3.0        # negative t is returned; callers filter
    assert isect(SQUARE, (0.5, -0.5, 1), (0, 0, -1)) == 1.0    # edges inclusive
    assert isect(SQUARE, (0.5001, 0, 1), (0, 0, -1)) == -INF
    assert isect(CIRCLE, (0.6, 0.8, 1), (0, 0, -1)) == -INF or isect(CIRCLE, (0.6, 0.8, 1), (0, 0, -1)) == 1.0
    assert isect(CIRCLE, (0.5, 0.5, 2), (0, 0, -1)) == 2.0
    assert isect(CIRCLE, (0.8, 0.8, 2), (0, 0, -1)) == -INF


def test_unit_box():                                 # AABB.intersect / get_intersects src/geometry.js:173-209
    assert isect(BOX, (0, 0, 5), (0, 0, -1)) == 4.5
    assert isect(BOX, (0, 0, 0), (0, 0, -1)) == 0.5             # from inside: t_min < minDistance -> t_max
    assert isect(BOX, (2, 0, 5), (0, 0, -1)) == -INF            # |p.x| > half with d.x == 0
    assert isect(BOX, (0, 0, 5), (0, 0, -2)) == 2.25            # parametric units of an unnormalised ray
    assert isect(BOX, (0, 0, 5), (0, 0, -1), 0.0, 4.0) == -INF  # t_min > maxDistance


def test_sphere():                                   # Sphere.staticIntersect src/geometry.js:429-442
    assert isect(SPHERE, (0, 0, 5), (0, 0, -1)) == 4.0
    assert isect(SPHERE, (0, 0, 0), (0, 0, -1)) == 1.0          # inside: the root behind is < minDistance
    assert isect(SPHERE, (0, 2, 5), (0, 0, -1)) == -INF
    assert isect(SPHERE, (0, 0, 5), (0, 0, -2)) == 2.0
    assert isect(SPHERE, (0, 0, 5), (0, 0, -1), 4.5) == 6.0     # near root below minDistance -> far root
    assert isect(SPHERE, (0, 0, 5), (0, 0, 0)) == -INF          # a == 0


def test_cylinder_has_no_caps():                     # src/geometry.js:473-478
    assert isect(CYL, (5, 0, 0.5), (-1, 0, 0)) == 4.0
    assert isect(CYL, (5, 0, 1.5), (-1, 0, 0)) == -INF          # |z| > 1 at the hit
    assert isect(CYL, (0, 0, 5), (0, 0, -1)) == -INF            # along the axis: a == 0, no caps
    # entering through where a cap would be: minDistance is clamped to the slab entry (t = 4) and the wall is hit later
    t = isect(CYL, (0, 0, 5), (0.2, 0, -1))
    assert abs(t - 5.0) < 1e-6


def test_triangle():                                 # src/geometry.js:341-353,368-375,389-396
    assert isect(TRI, (0.25, 0.25, 1), (0, 0, -1), extra=TRI_PS) == 1.0
    assert isect(TRI, (0.5, 0.5, 1), (0, 0, -1), extra=TRI_PS) == 1.0        # on the hypotenuse: bary 0 is inside
    assert isect(TRI, (0.6, 0.6, 1), (0, 0, -1), extra=TRI_PS) == -INF
    assert isect(TRI, (0.25, 0.25, 1), (1, 0, 0), extra=TRI_PS) == -INF      # parallel: denom == 0
    assert isect(TRI, (0.25, 0.25, -1), (0, 0, -1), extra=TRI_PS) == -1.0    # behind: returned as is (< 0)
    n = mdata(TRI, (0.25, 0.25, 0), TRI_PS)
    assert n[:3].tolist() == [0, 0, 1]


def test_material_data_quirks():
    # Sphere.materialData normalises the 4-vector position with w = 1 (src/geometry.js:449-455):
    # n = (0,1,0,1)/sqrt(2); v = 0.5 - asin(1/sqrt(2))/pi = 0.25 exactly up to f32
    m = mdata(SPHERE, (0, 1, 0))
    assert abs(m[1] - 1 / math.sqrt(2)) < 1e-7 and abs(m[3] - 0.5) < 1e-7 and abs(m[4] - 0.25) < 1e-6
    # AABB.materialData: first largest |component| wins (src/geometry.js:213-221)
    assert mdata(BOX, (0.5, 0.5, 0.1))[:3].tolist() == [1, 0, 0]
    assert mdata(BOX, (0.1, -0.5, 0.5))[:3].tolist() == [0, -1, 0]
    # plane: normal +z, UV = local xy (src/geometry.js:249-254)
    m = mdata(PLANE, (3.5, -2.25, 0))
    assert m.tolist() == [0, 0, 1, 3.5, -2.25]
    # cylinder: radial normal, UV = (0.5 + atan2(y,x)/2pi, 0.5 + z) (src/geometry.js:479-487)
    m = mdata(CYL, (0, 1, 0.25))
    assert m[:3].tolist() == [0, 1, 0] and abs(m[3] - 0.75) < 1e-7 and m[4] == 0.75


def test_fresnel():                                  # src/materials.js:358-386
    L = O.lib()
    k = C.c_double()
    kr = L.orc_kat_fresnel(1.3, 1.0, 0, C.byref(k))
    assert abs(kr - ((1.3 - 1) / (1.3 + 1)) ** 2) < 1e-12     # normal incidence: ((n-1)/(n+1))^2
    assert L.orc_kat_fresnel(1.3, 0.5, 1, C.byref(k)) == 1.0 and k.value < 0     # inside, beyond the critical angle: TIR
    assert L.orc_kat_fresnel(math.inf, 0.3, 0, C.byref(k)) == 1.0               # infinite IOR: mirror
    kr = L.orc_kat_fresnel(1.5, 0.0, 0, C.byref(k))
    assert abs(kr - 1.0) < 1e-12                                                 # grazing


def test_fmod_and_rng():
    L = O.lib()
    assert L.orc_kat_fmod(5.3, 2.0) == 1.3                     # toPrecision(8) cleans 1.2999999999999998
    assert L.orc_kat_fmod(-0.25, 2.0) == 1.75
    assert L.orc_kat_fmod(1 / 3 + 2, 2.0) == 0.33333333
    u = [L.orc_kat_rng(1, 7, 3, 1, d) for d in range(4)]
    assert all(0 <= x < 1 for x in u) and len(set(u)) == 4
    assert u[0] * 2 ** 24 == int(u[0] * 2 ** 24)               # 24-bit values
    assert L.orc_kat_rng(1, 7, 3, 1, 0) != L.orc_kat_rng(2, 7, 3, 1, 0) != L.orc_kat_rng(1, 8, 3, 1, 0)
    # statistical sanity of the generator that replaces Math.random()
    xs = np.array([L.orc_kat_rng(5, p, 0, 1, 0) for p in range(20000)])
    assert abs(xs.mean() - 0.5) < 0.01 and abs(xs.var() - 1 / 12) < 0.005


def test_rng_matches_device_header_constants():
    """oracle/oracle_math.h and csrc/rng.h must implement the same generator."""
    import re, os
    root = os.path.join(os.path.dirname(__file__), "..")
    a = open(os.path.join(root, "oracle", "oracle_math.h")).read()
    b = open(os.path.join(root, "jsraytracer_b200", "csrc", "rng.h")).read()
    for const in ("0x7feb352dU", "0x846ca68bU", "0x68bc21ebU", "0x9e3779b9U", "0x85ebca6bU", "0xc2b2ae35U"):
        assert const in a and const in b


def test_sdf_probes(blobs):
    """SDF distances from the formulas of src/sdf.js at hand-picked points."""
    js, _ = blobs("SDF_Menger", width=8, height=8)
    sc = O.OracleScene(js)
    d, _ = sc.sdf_probe(1, (0, 0, 2))      # outside the unit cube along z: the box term dominates: 2 - 1
    assert abs(d - 1.0) < 1e-6
    d, _ = sc.sdf_probe(1, (0, 0, 0))      # centre of the sponge lies in the removed cross: max(box, -cross) > 0
    assert d > 0
    js, _ = blobs("SDF_Simple", width=8, height=8)
    sc = O.OracleScene(js)
    d, m = sc.sdf_probe(1, (0, 3, 4))      # SphereSDF(1): |p| - 1 = 4
    assert abs(d - 4.0) < 1e-6
    assert abs(m[1] - 0.6) < 1e-3 and abs(m[2] - 0.8) < 1e-3       # forward-difference normal ~ p/|p|


def test_world_cast_semantics(blobs):
    """World.cast strict bounds (src/world.js:9-13) and shadow-ray window (src/materials.js:250-252)."""
    js, _ = blobs("BoxBall", width=8, height=8)
    sc = O.OracleScene(js)
    pid, t = sc.cast((-2, 5, -9), (0, -1, 0))                 # straight down onto the sphere at (-2,0.3,-9)
    assert pid == 2 and abs(t - 3.7) < 1e-6
    pid, t = sc.cast((-2, 5, -9), (0, -1, 0), 0.0, 3.0)       # maxDistance before the sphere and the floor
    assert pid == -1 and t == INF
    pid, t = sc.cast((-2, 5, -9), (0, -10, 0), 0.0001, 1.0, shadow=True)   # parametric units: hit at t = 0.37
    assert pid == 2 and abs(t - 0.37) < 1e-6
