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


# ---- whole-pipeline known answers on hand-built scenes: camera, Phong lighting, shadows, recursion, resolve ----
def _scene(objects, lights, depth, width=2, height=2, bg=None, fov=math.pi / 2):
    from jsraytracer_b200.cameras import PerspectiveCamera
    from jsraytracer_b200.jsmath import Mat4
    from jsraytracer_b200.renderers import SimpleRenderer
    from jsraytracer_b200.serializer import Serializer
    from jsraytracer_b200.world import World
    world = World(objects, lights) if bg is None else World(objects, lights, bg)
    test = {"renderer": SimpleRenderer(world, PerspectiveCamera(fov, 1, Mat4.identity()), depth), "width": width, "height": height}
    return O.OracleScene(Serializer(test).to_json())


def _wall(z, material):
    from jsraytracer_b200.geometry import Plane
    from jsraytracer_b200.jsmath import Mat4
    from jsraytracer_b200.world import Primitive
    return Primitive(Plane(), material, Mat4.translation([0, 0, z]))      # local plane z = 0, normal +z: faces the camera


def test_camera_rays_are_pixel_corner_and_unnormalised():
    """`x = 2 px / W - 1`, `y = -2 py / H + 1` (pixel corner, src/renderers.js:22,25) and `dir = T (x tan(fov/2) aspect,
    y tan(fov/2), -1, 0)`, not normalised (src/cameras.js:29-34): every camera ray reaches the wall z = -d at t = d exactly."""
    from jsraytracer_b200.materials import PhongMaterial
    from jsraytracer_b200.jsmath import Vec
    sc = _scene([_wall(-3.5, PhongMaterial(Vec.of(1, 1, 1), 1))], [], 1, width=4, height=4)
    ids, t, _ = sc.primary_hits()
    assert (ids == 0).all() and np.all(t == 3.5)
    # pixel (0, 0) looks along (-tan(45 deg), +tan(45 deg), -1): straight rays only at the centre corner px = W/2, py = H/2
    pid, tt = sc.cast((0, 0, 0), (-1, 1, -1))
    assert pid == 0 and tt == 3.5


def test_phong_point_light_falloff_shadow_and_ambient():
    """colorFromLights (src/materials.js:240-259) with SimplePointLight (src/lights.js:45-53) and Light.falloff (:21-23):
    radiance = ambient + colour * intensity / (4 pi |delta|^2) * (diffusivity max(L.N, 0) + specularity max(L.R, 0)^smoothness),
    nothing from a light whose shadow ray finds an occluder at 0 < t < 1 (:250-252)."""
    from jsraytracer_b200.geometry import Sphere
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.materials import PhongMaterial, SimplePointLight
    from jsraytracer_b200.world import Primitive
    I = 36 * math.pi * 0.8                                       # so that the falloff at |delta| = 3 is exactly 0.8
    light = SimplePointLight(Vec.of(0, 0, -2, 1), Vec.of(1, 0.5, 0.25), I)
    wall = _wall(-5, PhongMaterial(Vec.of(1, 1, 1), 0.1, 0.5, 0, 7))           # ambient 0.1, diffusivity 0.5, no specular, no mirror
    img, cnt = _scene([wall], [light], 1).render(1, seed=1, jitter=False)
    # pixel (1, 1) of a 2 x 2 frame is NDC (0, 0): the ray along -z hits (0, 0, -5); L = N = (0, 0, 1)
    assert np.allclose(img[1, 1], [0.1 + 0.5 * 0.8, 0.1 + 0.5 * 0.4, 0.1 + 0.5 * 0.2], rtol=0, atol=2e-7)
    assert cnt["rays_shadow"] == 4 and cnt["rays_secondary"] == 0
    # with a specular term: R = N at normal incidence, so max(L.R, 0)^smoothness = 1
    wall2 = _wall(-5, PhongMaterial(Vec.of(1, 1, 1), 0.1, 0.5, 0.25, 7))
    img2, _ = _scene([wall2], [light], 1).render(1, seed=1, jitter=False)
    assert np.allclose(img2[1, 1], [0.1 + 0.75 * 0.8, 0.1 + 0.75 * 0.4, 0.1 + 0.75 * 0.2], atol=3e-7)
    # an occluder between the wall and the light (the camera ray is blocked too, so look at the wall through a shadow-only
    # blocker: does_cast_shadow = true, placed off the camera ray but on the light ray of pixel (1,1)?  Simpler: move the
    # light behind a ball that the camera ray misses.)
    light_b = SimplePointLight(Vec.of(3, 0, -5, 1), Vec.of(1, 1, 1), 1000.0)   # grazing light from +x at the wall's depth: L.N = 0
    img3, _ = _scene([wall], [light_b], 1).render(1, seed=1, jitter=False)
    assert np.allclose(img3[1, 1], [0.1, 0.1, 0.1], atol=1e-7)                 # max(L.N, 0) = 0: ambient only
    ball = Primitive(Sphere(), PhongMaterial(Vec.of(1, 1, 1), 1), Mat4.translation([0, 1.5, -3.5]).times(Mat4.scale(0.5)))
    light_c = SimplePointLight(Vec.of(0, 3, -2, 1), Vec.of(1, 1, 1), I)        # the ball sits on the segment wall point -> light
    lit, _ = _scene([wall], [light_c], 1).render(1, seed=1, jitter=False)
    shadowed, _ = _scene([wall, ball], [light_c], 1).render(1, seed=1, jitter=False)
    assert lit[1, 1, 0] > 0.11 and np.allclose(shadowed[1, 1], [0.1, 0.1, 0.1], atol=1e-7)


def test_recursion_depth_exhaustion_is_black_not_background():
    """World.color returns (0,0,0) at depth 0 (src/world.js:32-33), the background only on a miss (:35): a perfect mirror
    seen with maxRecursionDepth = 1 shows its local term only; with depth 2 the reflected ray escapes to bg_color."""
    from jsraytracer_b200.jsmath import Vec
    from jsraytracer_b200.materials import PhongMaterial
    mirror = PhongMaterial(Vec.of(1, 1, 1), 0.2, 0, 0, 5, 1.0)                  # ambient 0.2, reflectivity 1
    bg = Vec.of(0.3, 0.6, 0.9)
    d1, c1 = _scene([_wall(-4, mirror)], [], 1, bg=bg).render(1, seed=1, jitter=False)
    d2, c2 = _scene([_wall(-4, mirror)], [], 2, bg=bg).render(1, seed=1, jitter=False)
    assert np.allclose(d1[1, 1], [0.2, 0.2, 0.2], atol=1e-7)
    assert np.allclose(d2[1, 1], [0.2 + 0.3, 0.2 + 0.6, 0.2 + 0.9], atol=2e-7)
    assert c1["rays_secondary"] == 0 and c2["rays_secondary"] == 4              # the (D+1)-th ray is never cast


def test_resolve_rounds_like_pixelbuffer_setcolor():
    """`round(255 * clamp(c, 0, 1))` per channel, alpha 255, no gamma (src/pixelbuffer.js:39-49) on sum / passes."""
    acc = np.array([[[0.0, 0.5, 1.0], [2.0, -1.0, 0.25], [0.501960784, 1.5, 3.0]]], dtype=np.float32)     # (1, 3, 3) sums
    out = O.resolve_rgba8(acc, 1)
    assert out[0, 0].tolist() == [0, 128, 255, 255]            # Math.round(127.5) = 128
    assert out[0, 1].tolist() == [255, 0, 64, 255]             # clamped; Math.round(63.75) = 64
    out2 = O.resolve_rgba8(acc, 2)                              # two passes: mean = sum / 2
    assert out2[0, 2].tolist() == [64, 191, 255, 255]          # 0.25098 -> 64, 0.75 -> Math.round(191.25) = 191, 1.5 -> clamped


def test_area_light_falloff_and_cosine_factor():
    """RandomSampleAreaLight.sampleIterator (src/lights.js:80-93): colour * falloff(delta) * |normalize(delta) . n_light| with
    n_light = normalize((T^-1)^T n_local).  A 1e-6-sized Square makes the sampled point independent of the random numbers."""
    from jsraytracer_b200.geometry import Square
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.materials import PhongMaterial, RandomSampleAreaLight
    I = 36 * math.pi * 0.8
    wall = _wall(-5, PhongMaterial(Vec.of(1, 1, 1), 0.1, 0.5, 0, 7))
    for tilt, cosine in ((0.0, 1.0), (math.pi / 3, 0.5), (math.pi / 2 - 1e-3, math.sin(1e-3))):
        T = Mat4.translation([0, 0, -2]).times(Mat4.rotation(tilt, Vec.of(1, 0, 0))).times(Mat4.scale(1e-6))
        light = RandomSampleAreaLight(Square(), T, Vec.of(1, 0.5, 0.25), I, 3)          # three samples, averaged
        img, cnt = _scene([wall], [light], 1).render(1, seed=5, jitter=False)
        expect = np.array([0.1, 0.1, 0.1]) + 0.5 * 0.8 * cosine * np.array([1, 0.5, 0.25])
        assert np.allclose(img[1, 1], expect, atol=2e-6), (tilt, img[1, 1], expect)
        assert cnt["rays_shadow"] == 4 * 3                                             # one shadow ray per sample per shaded hit


def test_path_tracing_diffuse_weight_has_no_pdf_division():
    """PhongPathTracingMaterial.scatter (src/materials.js:398-412): with mirrorProbability 0 and no specular lobe the
    diffuse branch is always taken and weighs the scattered radiance by diffusivity / pi — as written, without dividing by
    the selection probability or the cosine pdf.  Above an infinite wall every scattered ray escapes to bg_color, so the
    pixel is ambient + bg * diffusivity / pi whatever the random numbers are."""
    from jsraytracer_b200.jsmath import Vec
    from jsraytracer_b200.materials import PhongPathTracingMaterial
    bg = Vec.of(0.3, 0.6, 0.9)
    mat = PhongPathTracingMaterial(Vec.of(1, 0.5, 1), 0.1, 0.6, 0, 5)                   # IOR infinite: kr = 1, no refraction branch
    img, cnt = _scene([_wall(-4, mat)], [], 2, bg=bg).render(1, seed=9, jitter=False)
    base = np.array([1, 0.5, 1])
    expect = 0.1 * base + np.array([0.3, 0.6, 0.9]) * (0.6 * base) / math.pi
    assert np.allclose(img[1, 1], expect, atol=3e-7), (img[1, 1], expect)
    assert np.allclose(img[0, 0], expect, atol=3e-7)                                   # every pixel: the wall fills the frame
    assert cnt["rays_secondary"] == 4
    # depth 3: the escaped ray adds nothing more
    img3, _ = _scene([_wall(-4, mat)], [], 3, bg=bg).render(1, seed=9, jitter=False)
    assert np.allclose(img3, img, atol=1e-7)


def test_fresnel_branch_weights_at_normal_incidence():
    """FresnelPhongMaterial.color (src/materials.js:309-333): reflection weighted kr, refraction weighted 1 - kr, both traced;
    unpolarised Fresnel at normal incidence: kr = ((n - 1) / (n + 1))^2 (:366-386).  A Fresnel sheet in front of an ambient-only
    wall: the pixel is kr * bg (mirror ray escapes behind the camera) + (1 - kr) * wall colour."""
    from jsraytracer_b200.jsmath import Vec
    from jsraytracer_b200.materials import FresnelPhongMaterial, PhongMaterial
    bg = Vec.of(0.2, 0.4, 0.8)
    back = PhongMaterial(Vec.of(0.9, 0.6, 0.3), 1)                               # ambient = its colour, nothing else
    for n in (1.5, 2.4):
        sheet = FresnelPhongMaterial(Vec.of(1, 1, 1), 0, 0, 0, 5, n)
        img, cnt = _scene([_wall(-3, sheet), _wall(-6, back)], [], 2, bg=bg).render(1, seed=1, jitter=False)
        kr = ((n - 1) / (n + 1)) ** 2
        expect = kr * np.array([0.2, 0.4, 0.8]) + (1 - kr) * np.array([0.9, 0.6, 0.3])
        assert np.allclose(img[1, 1], expect, atol=3e-7), (n, img[1, 1], expect)
        assert cnt["rays_secondary"] == 8                                          # two children per camera ray


def test_vertex_normals_are_blended_by_barycentrics():
    """Triangle.materialData / blend (src/geometry.js:376-385,397-409): the shading normal is the barycentric mix of the
    per-vertex normals, normalised by Primitive.color (src/world.js:134).  Hit point (0,0,-5) has barycentrics
    (0.25, 0.25, 0.5) in the triangle below; with n2 tilted 45 degrees the blend is tilted 22.5 degrees: L.N = cos(22.5 deg)."""
    from jsraytracer_b200.geometry import Triangle
    from jsraytracer_b200.jsmath import Vec
    from jsraytracer_b200.materials import PhongMaterial, SimplePointLight
    from jsraytracer_b200.world import Primitive
    I = 36 * math.pi * 0.8
    ps = [Vec.of(-4, -4, -5, 1), Vec.of(4, -4, -5, 1), Vec.of(0, 4, -5, 1)]
    s = math.sqrt(0.5)
    mat = PhongMaterial(Vec.of(1, 1, 1), 0.1, 0.5, 0, 7)
    light = SimplePointLight(Vec.of(0, 0, -2, 1), Vec.of(1, 1, 1), I)
    flat = Primitive(Triangle(ps, {}), mat)
    smooth = Primitive(Triangle(ps, {"normal": [Vec.of(0, 0, 1, 0), Vec.of(0, 0, 1, 0), Vec.of(0, s, s, 0)]}), mat)
    a, _ = _scene([flat], [light], 1).render(1, seed=1, jitter=False)
    b, _ = _scene([smooth], [light], 1).render(1, seed=1, jitter=False)
    assert np.allclose(a[1, 1], 0.1 + 0.5 * 0.8, atol=2e-7)
    assert np.allclose(b[1, 1], 0.1 + 0.5 * 0.8 * math.cos(math.pi / 8), atol=3e-7), b[1, 1]


def test_depth_of_field_rays_meet_on_the_focal_plane():
    """DepthOfFieldPerspectiveCamera.getRayForPixel (src/cameras.js:46-52): origin += lens offset, direction =
    normalize(dir * focus_distance - offset) — every lens sample of a pixel passes through the pinhole ray's point at
    parameter focus_distance, i.e. on the plane z = -focus_distance for this camera.  A checkerboard wall exactly there
    renders identically for any sensor size; the same wall nearer or farther does not."""
    from jsraytracer_b200.cameras import DepthOfFieldPerspectiveCamera
    from jsraytracer_b200.geometry import Plane
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.materials import CheckerboardMaterialColor, PhongMaterial
    from jsraytracer_b200.renderers import SimpleRenderer
    from jsraytracer_b200.serializer import Serializer
    from jsraytracer_b200.world import Primitive, World

    def render(z, sensor):
        wall = Primitive(Plane(), PhongMaterial(CheckerboardMaterialColor(Vec.of(1, 1, 1), Vec.of(0, 0, 0)), 1), Mat4.translation([0.25, 0.25, z]))
        cam = DepthOfFieldPerspectiveCamera(math.pi / 2, 1, Mat4.identity(), 4.0, sensor)
        test = {"renderer": SimpleRenderer(World([wall], []), cam, 1), "width": 16, "height": 16}
        return O.OracleScene(Serializer(test).to_json()).render(1, seed=3, jitter=False)[0]

    sharp = render(-4.0, 0.0)
    assert 0.3 < sharp.mean() < 0.7                                      # a checkerboard, not a flat frame
    assert np.array_equal(render(-4.0, 0.6), sharp)                      # in focus: the lens does not matter
    assert not np.array_equal(render(-8.0, 0.6), render(-8.0, 0.0))      # out of focus: it does


def test_sphere_tracing_converges_from_outside_within_epsilon():
    """SDFGeometry.intersect (src/sdf.js:12-40): clip to the SDF's box, then t += distance / |d| until distance <= epsilon.
    On a unit SphereSDF the march approaches the analytic root from the near side and stops within epsilon of it (in units of
    distance; the returned t is parametric, so an unnormalised direction scales it); a ray through the centre lands on the
    surface in one step; a ray that misses the sphere inside its box reports no hit."""
    from jsraytracer_b200.jsmath import Vec
    from jsraytracer_b200.materials import PhongMaterial
    from jsraytracer_b200.sdf import SDFGeometry, SphereSDF
    from jsraytracer_b200.world import Primitive
    eps = 1e-5
    prim = Primitive(SDFGeometry(SphereSDF(1), 300, eps, 100), PhongMaterial(Vec.of(1, 1, 1), 1))
    sc = _scene([prim], [], 1)
    pid, t = sc.cast((0, 0, 5), (0, 0, -1))
    assert pid == 0 and abs(t - 4.0) <= eps
    root = 5 - math.sqrt(1 - 0.25)
    pid, t = sc.cast((0.5, 0, 5), (0, 0, -1))
    assert pid == 0 and 0 <= root - t <= 3 * eps
    pid, t2 = sc.cast((0.5, 0, 5), (0, 0, -2))                 # unnormalised direction: same point, half the parameter
    assert pid == 0 and 0 <= root / 2 - t2 <= 3 * eps
    pid, t = sc.cast((0.99, 0.99, 5), (0, 0, -1))              # inside the box's corner, outside the sphere
    assert pid == -1 and t == INF
    pid, t = sc.cast((0, 0, 5), (0, 0, -1), 0.0, 3.9)          # maxDistance in front of the surface
    assert pid == -1
