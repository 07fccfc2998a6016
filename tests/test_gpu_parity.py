"""Parity of the CUDA path (through the C ABI) with the CPU restatement oracle.

Bars (BASELINE.json north_star):
  * primary-ray hit primitive IDs equal on >= 99.99 % of pixels;
  * hit distance within 1e-4 relative (reference: f64 scalars on f32 vectors; device:
    FP32, except the SDF path which mirrors the reference's arithmetic) — asserted
    as: at most 0.01 % of the commonly-hit pixels exceed 1e-4;
  * deterministic Whitted / Phong images: PSNR >= 50 dB on the f32 image;
  * stochastic scenes, same counter-based RNG on both sides: PSNR >= 50 dB per
    sample set, and the ray counts agree to 1e-3 relative (rare discrete flips).
"""
import numpy as np
import pytest

from conftest import psnr, scene_blobs

pytestmark = pytest.mark.gpu


def _pair(name, **kw):
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    js, mp = scene_blobs(name, **kw)
    return lib.Scene(mp, lib.FORMAT_MSGPACK, device=0), OracleScene(js)


PRIMARY = [
    ("BoxBall", dict(width=512, height=512)),                      # BASELINE configs[0]
    ("ASimpleScene", dict(width=256, height=256)),                 # every analytic primitive incl. Cylinder, Circle
    ("cornell_box_path", dict(width=256, height=256)),
    ("bunny_path", dict(width=640, height=360, aspect=16 / 9)),
    ("AHollowTetrahedron", dict(width=256, height=256)),           # two BVHAggregates sharing one kdtree
    ("SDF_Sierpinski", dict(width=160, height=160)),
    ("SDF_Menger", dict(width=160, height=160)),
    ("SDF_CrossFolds", dict(width=192, height=192)),               # the fused axis-bar leaf in its push / fold-min / fold-max forms
    ("spheres010", dict(width=200, height=200)),
    ("Aggregates", dict(width=256, height=256)),                   # plain Aggregates, nested, with a shared primitive
    ("dragon_grid", dict(width=320, height=180, aspect=16 / 9, n=2)),   # 8 instances of one kdtree (Toledo-scale stand-in)
    ("starwars", dict(width=480, height=270, aspect=16 / 9)),      # 4 BVH instances (3 share a kdtree), MTL materials; stands in for Toledo
    ("spheres100", dict(width=200, height=200)),                   # 101 top-level primitives in the linear world list
    ("cornell_box", dict(width=256, height=256)),
    ("AMultipleBVH", dict(width=256, height=256)),                 # two different meshes in two BVHs, a sphere that casts no shadow
    ("cat", dict(width=256, height=256)),
    ("diamond", dict(width=256, height=256)),
    ("heart", dict(width=256, height=256)),
    ("utah_teapot", dict(width=256, height=256)),                  # high-poly-teapot.obj, 6 320 faces
    ("x_wing", dict(width=256, height=256)),                       # 18 849 triangles, many slivers
    ("SDF_SphereRepetition", dict(width=160, height=160)),         # Math.fmod with a period that is not a power of two
    ("bottle", dict(width=256, height=256)),                       # MTL map_Kd texture on a mesh
]


@pytest.mark.parametrize("name,kw", PRIMARY, ids=[p[0] for p in PRIMARY])
def test_primary_hit_ids_and_distances(name, kw):
    sc, orc = _pair(name, **kw)
    ids, t = sc.primary_hits()
    oids, ot, _ = orc.primary_hits()
    agree = float((ids == oids).mean())
    assert agree >= 0.9999, "hit-ID agreement %.6f" % agree
    both = (ids == oids) & (oids >= 0)
    rel = np.abs(t[both].astype(np.float64) - ot[both]) / np.abs(ot[both])
    assert float((rel > 1e-4).mean()) <= 1e-4, "t rel err: max %.3e, frac>1e-4 %.2e" % (rel.max(), (rel > 1e-4).mean())


def test_primary_hits_dragon_1080p():
    """BASELINE configs[2]: dragon at 1920x1080, 99 968 triangles / 199 935 nodes."""
    sc, orc = _pair("dragon", width=1920, height=1080, aspect=16 / 9)
    assert sc.info["n_nodes"] == 199935 and sc.info["n_tris"] == 99968
    ids, t = sc.primary_hits()
    oids, ot, _ = orc.primary_hits()
    assert float((ids == oids).mean()) >= 0.9999
    both = (ids == oids) & (oids >= 0)
    rel = np.abs(t[both].astype(np.float64) - ot[both]) / np.abs(ot[both])
    assert float((rel > 1e-4).mean()) <= 1e-4


WHITTED = [
    ("BoxBall", dict(width=512, height=512), 1),
    ("ASimpleScene", dict(width=256, height=256), 1),
    ("bunny", dict(width=480, height=270, aspect=16 / 9), 1),
    ("AHollowTetrahedron", dict(width=256, height=256), 1),
    # Fresnel spheres act as lenses that image the floor's horizon onto whole pixel rows; those pixels are
    # ill-conditioned in the reference itself (a 1e-7 rad camera roll flips ~50 of them at 256^2), so this
    # scene is compared at the resolution where such rows are a small share of the frame.
    ("refraction", dict(width=768, height=768), 1),
    ("Aggregates", dict(width=256, height=256), 1),
    ("SDF_Sierpinski", dict(width=160, height=160), 1),
    ("SDF_Menger", dict(width=160, height=160), 1),
    ("SDF_CrossFolds", dict(width=192, height=192), 1),
    ("SDF_BoxBall", dict(width=160, height=160), 1),            # per-leaf basecolor through UnionSDF.getMaterialData
    ("SDF_Combinations", dict(width=192, height=192), 1),       # all six combinators incl. the smooth blends
    ("SDF_Simple", dict(width=128, height=128), 1),
    # an infinite lattice of mirror balls (reflectivity 0.5): every ball-to-ball bounce multiplies a perturbation ~10x, so
    # the depth-4 image is ill-conditioned in the reference itself (test_ill_conditioned_mirror_lattice below); the
    # well-conditioned part — camera ray + first reflection — is compared here
    ("SDF_SphereRepetition", dict(width=160, height=160, depth=2), 1),
    ("spheres050", dict(width=256, height=256), 1),
    ("refraction_simple", dict(width=512, height=512), 1),      # one Fresnel sphere, three coloured point lights
    ("AMultipleBVH", dict(width=256, height=256), 1),
    ("cat", dict(width=256, height=256), 1),
    ("utah_teapot", dict(width=256, height=256), 1),
    ("diamond", dict(width=384, height=384), 1),                # Fresnel mesh, IOR 2.4, non-black background
    ("bottle", dict(width=256, height=256), 1),                 # TextureMaterialColor from the MTL, UVs blended per triangle hit
]


@pytest.mark.parametrize("name,kw,passes", WHITTED, ids=[p[0] for p in WHITTED])
def test_deterministic_images_psnr(name, kw, passes):
    from jsraytracer_b200 import lib
    sc, orc = _pair(name, **kw)
    sc.render(0, passes, seed=1, flags=lib.FLAG_NO_JITTER)
    acc, _ = sc.read_accum()
    oacc, _ = orc.render(passes, seed=1, jitter=False)
    g, o = np.clip(acc[..., :3] / passes, 0, 1), np.clip(oacc / passes, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
    assert np.all(acc[..., 3] == passes)


def test_ill_conditioned_mirror_lattice():
    """tests/SDF_SphereRepetition at its own depth 4.  Measured per recursion depth 1..4: 0 / 1 / 251 / 1 108 of 25 600
    pixels differ by more than 1e-3 between the CUDA path and the oracle — differences are born at one-ulp level and
    grow with every bounce between convex mirrors.  The bar is therefore the reference's own conditioning: the oracle
    re-rendered with the camera rolled by 1e-7 rad (less than one f32 ulp of a pixel's ray direction) moves away from
    itself by more (37.8 dB) than the CUDA path is away from the oracle (45.7 dB)."""
    from jsraytracer_b200 import lib, scenes
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.serializer import Serializer
    from oracle.oracle import OracleScene
    sc, orc = _pair("SDF_SphereRepetition", width=160, height=160)
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    g = np.clip(sc.read_accum()[0][..., :3], 0, 1)
    o = np.clip(orc.render(1, seed=1, jitter=False)[0], 0, 1)
    test = scenes.configure("SDF_SphereRepetition", width=160, height=160)
    cam = test["renderer"].camera
    cam.transform = cam.transform.times(Mat4.rotation(1e-7, Vec.of(0, 0, 1)))
    cam.inv_transform = Mat4.inverse(cam.transform)
    o_rolled = np.clip(OracleScene(Serializer(test).to_json()).render(1, seed=1, jitter=False)[0], 0, 1)
    own = psnr(o, o_rolled)
    assert own < 45.0, "the scene is not as ill-conditioned as documented: %.2f dB" % own
    assert psnr(g, o) >= own + 3.0, "CUDA vs oracle %.2f dB, oracle vs rolled oracle %.2f dB" % (psnr(g, o), own)


STOCHASTIC = [
    ("BoxBall", dict(width=256, height=256), 4),
    ("BoxBall_DOF", dict(width=256, height=256), 4),
    ("BoxBall_path", dict(width=256, height=256), 4),
    ("cornell_box_path", dict(width=256, height=256), 4),
    ("bunny_path", dict(width=480, height=270, aspect=16 / 9), 4),
    ("refraction_path", dict(width=192, height=192), 4),
    ("starwars", dict(width=320, height=180, aspect=16 / 9), 2),   # DOF + 3 square area lights x 4 samples + point light
    ("cornell_box", dict(width=192, height=192), 2),               # Whitted materials under the area light, depth 7
    ("cornell_box_emissive", dict(width=192, height=192), 4),      # world.lights = []: emissive ceiling only
    ("heart", dict(width=256, height=256), 2),                     # spherical area light (Sphere.sampleSurface)
    ("x_wing", dict(width=256, height=256), 2),
]


@pytest.mark.parametrize("name,kw,passes", STOCHASTIC, ids=[p[0] for p in STOCHASTIC])
def test_same_rng_sample_parity(name, kw, passes):
    sc, orc = _pair(name, **kw)
    sc.stats_reset()
    sc.render(0, passes, seed=7)
    acc, _ = sc.read_accum()
    st = sc.stats()
    oacc, cnt = orc.render(passes, seed=7)
    g, o = np.clip(acc[..., :3] / passes, 0, 1), np.clip(oacc / passes, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
    for k in ("rays_primary", "rays_secondary", "rays_shadow"):
        assert abs(st[k] - cnt[k]) <= 1e-3 * max(1, cnt[k]), (k, st[k], cnt[k])
    assert st["rays_primary"] == cnt["rays_primary"]


def test_pass_ranges_compose():
    """Rendering passes [0,2) then [2,5) equals [0,5) up to atomic summation order."""
    sc, _ = _pair("cornell_box_path", width=128, height=128)
    sc.render(0, 5, seed=3)
    a, pa = sc.read_accum()
    sc.reset_accum()
    sc.render(0, 2, seed=3)
    sc.render(2, 3, seed=3)
    b, pb = sc.read_accum()
    assert pa == pb == 5
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)


def test_column_striping_matches_full_frame():
    """x_offset / x_delt (src/renderers.js:21,88): 3 interleaved stripes = the full frame."""
    from jsraytracer_b200 import lib
    sc, _ = _pair("BoxBall", width=200, height=120)
    sc.render(0, 2, seed=5)
    full, _ = sc.read_accum()
    sc.reset_accum()
    sc.render(0, 2, seed=5, x_offset=1, x_delt=3)
    part = sc.resolve_rgba8().copy()
    assert np.all(part[:, 0::3, 3] == 0) and np.all(part[:, 1::3, 3] == 255) and np.all(part[:, 2::3, 3] == 0)
    sc.render(0, 2, seed=5, x_offset=0, x_delt=3)
    sc.render(0, 2, seed=5, x_offset=2, x_delt=3)
    tri, _ = sc.read_accum()
    assert np.allclose(full, tri, rtol=1e-5, atol=1e-5)


def test_resolve_matches_pixelbuffer_semantics():
    """round(255 * clamp(sum / passes)) and alpha 255 (src/pixelbuffer.js:39-49)."""
    from oracle.oracle import resolve_rgba8
    sc, _ = _pair("BoxBall", width=160, height=160)
    sc.render(0, 3, seed=2)
    acc, passes = sc.read_accum()
    img = sc.resolve_rgba8().copy()
    expect = resolve_rgba8(acc[..., :3].copy(), passes)
    assert np.array_equal(img, expect)
    assert np.array_equal(sc.resolve_rgba8(), img)           # idempotent


def test_cuda_renderer_drop_in():
    """CUDARenderer.render(img, timelimit, callback, x_offset, x_delt) fills img.imgdata.data like the
    reference renderers (src/renderers.js:70-117) and reports {pass, completion}."""
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.pixelbuffer import PixelBuffer
    from jsraytracer_b200.renderers import CUDARenderer
    from jsraytracer_b200.serializer import Serializer
    from oracle.oracle import OracleScene, resolve_rgba8
    test = scenes.configure("BoxBall", width=128, height=96, spp=4, renderer_cls=CUDARenderer)
    img = PixelBuffer(test["width"], test["height"])
    seen = []
    out = test["renderer"].render(img, 1e-6, lambda s: seen.append(s))
    assert out is img
    assert seen and all(0 < s["completion"] <= 1 and 0 <= s["pass"] < 4 for s in seen)
    orc = OracleScene(Serializer(test).to_json())
    oacc, _ = orc.render(4, seed=1, width=128, height=96)
    expect = resolve_rgba8(oacc, 4)
    got = img.as_array()
    assert np.all(got[..., 3] == 255)
    diff = np.abs(got[..., :3].astype(int) - expect[..., :3].astype(int))
    assert float((diff > 1).mean()) < 2e-3                    # 8-bit images agree except rare discrete flips


def test_scene_reupload_and_stats():
    sc, _ = _pair("bunny_path", width=320, height=180, aspect=16 / 9)
    sc.stats_reset()
    sc.render(0, 1, seed=1)
    a, _ = sc.read_accum()
    s1 = sc.stats()
    sc.upload()
    sc.reset_accum()
    sc.render(0, 1, seed=1)
    b, _ = sc.read_accum()
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert s1["rays_primary"] == 320 * 180 and s1["launches"] > 0 and s1["rays_shadow"] == 2 * s1["shaded_hits"]
