"""Pin against an output of the reference itself: its committed screenshot of tests/tie_fighter.

The reference has no machine-checked fixtures, but `tests/tie_fighter/download (10).png` /
`download (11).png` are 600x600 renders made by the reference's CPU renderer of a scene file that still matches
(`tests/tie_fighter/test.mjs`: Plane + Fresnel bolt sphere + Tie_Fighter.obj BVH, one point light, one green area
light, IncrementalMultisamplingRenderer 16 spp depth 4).  tools/make_reference_screenshot_fixture.py decodes them into
tests/golden/reference_screenshot_tie_fighter_600.npz.

What can be compared pixel-for-pixel, and why only that:
  * the R and B channels.  The area light's colour is (0,1,0) (`test.mjs:13-20`) and every light term is a
    per-channel product (`src/materials.js:261-268`), so R and B do not depend on the randomly sampled area light;
  * pixels whose camera ray hits the Plane and whose mirror ray leaves the scene, and sky pixels.  The mesh body in
    the screenshots carries a constant +0.1 ambient term that today's Tie_Fighter.mtl (`#Ka` commented out) no
    longer produces, so body pixels (and their reflections) are stale and excluded;
  * pixels at least 2 px away from any classification edge (silhouettes, shadow boundaries): the reference jitters
    with Math.random(), so edge pixels differ by anti-aliasing noise.
That leaves ~304 000 of 360 000 pixels, which exercise: camera ray generation, Plane intersection, the point-light
Phong terms with falloff, the shadow-ray walk through the 10 712-triangle BVH (lit/shadowed classification of every
plane pixel), reflection rays that miss (black, `src/world.js:32-36`), accumulation over 16 jittered passes and the
8-bit resolve (`src/pixelbuffer.js:39-49`).  Bar: every such pixel within 2 levels, >= 99.9 % within 1 level,
>= 94 % bit-identical.  The mesh silhouette against the sky is compared too (hit / no hit, away from edges).

The green channel was examined for a statistical pin of the RandomSampleAreaLight (G - R on plane pixels = the area
light's term: the plane's colour is grey and the light's is (0,1,0)) and cannot be used: the two screenshots disagree
with EACH OTHER by 2-3x in that term (mean G - R on the lit plane 2.92 vs 6.27 levels, on the shadowed plane 5.18 vs
15.17) and their ratio to a render of today's test.mjs varies from 0.57 to 4.9 across the image, i.e. the bolt light was
moved / resized between the two screenshots and again before test.mjs was committed.  test_area_light_term_is_stale
pins that finding, so that the exclusion of G is a checked fact rather than a convenience.
"""
import math
import os

import numpy as np
import pytest

from conftest import scene_blobs

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "reference_screenshot_tie_fighter_600.npz")
W = H = 600
LIGHT = np.array([-10.0, 10.0, -12.0])         # tests/tie_fighter/test.mjs:8


def _classes():
    """Per-pixel class from the oracle's un-jittered rays: 0 sky, 1 plane lit, 2 plane shadowed, 3 plane whose mirror
    ray hits something, 4 mesh / bolt; plus the 'interior' mask (5x5 neighbourhood of one class)."""
    from scipy.ndimage import maximum_filter, minimum_filter
    from jsraytracer_b200 import scenes
    from oracle.oracle import OracleScene
    js, _ = scene_blobs("tie_fighter", width=W, height=H)
    orc = OracleScene(js)
    ids, t, _ = orc.primary_hits()
    cam = scenes.configure("tie_fighter", width=W, height=H)["renderer"].camera
    T = np.array(cam.transform, dtype=np.float64).reshape(4, 4)
    tan = math.tan(math.pi / 8)
    X, Y = np.meshgrid(2 * np.arange(W) / W - 1, -2 * np.arange(H) / H + 1)      # src/renderers.js:89,92
    d = (np.stack([X * tan, Y * tan, -np.ones_like(X), np.zeros_like(X)], -1) @ T.T)[..., :3]
    cls = np.where(ids < 0, 0, 4)
    for y, x in zip(*np.nonzero(ids == 0)):
        p = T[:3, 3] + d[y, x] * t[y, x]
        r = d[y, x] * np.array([1.0, -1.0, 1.0])                                 # mirror about the plane normal +Y
        if orc.cast(list(p) + [1.0], list(r) + [0.0], 1e-4)[0] >= 0:
            cls[y, x] = 3
        else:
            shadowed = orc.cast(list(p) + [1.0], list(LIGHT - p) + [0.0], 1e-4, 1.0, True)[0] >= 0
            cls[y, x] = 2 if shadowed else 1
    interior = minimum_filter(cls, 5) == maximum_filter(cls, 5)
    return orc, ids, cls, interior


_CACHE = {}


def _setup():
    if not _CACHE:
        _CACHE["v"] = _classes()
    return _CACHE["v"]


def _check(img8, what):
    _, ids, cls, interior = _setup()
    shots = np.load(FIXTURE)
    img = img8[..., :3].astype(np.int32)
    for key in ("shot10", "shot11"):
        ref = shots[key].astype(np.int32)
        d = np.abs(img[..., [0, 2]] - ref[..., [0, 2]]).max(-1)
        sky = interior & (cls == 0)
        assert sky.sum() > 75000 and d[sky].max() == 0, "%s vs %s: sky" % (what, key)
        for c, name, exact in ((1, "plane lit", 0.94), (2, "plane shadowed", 0.999)):
            m = interior & (cls == c)
            assert m.sum() > (150000 if c == 1 else 25000)
            assert d[m].max() <= 2, "%s vs %s: %s max %d" % (what, key, name, d[m].max())
            assert (d[m] <= 1).mean() >= 0.999, "%s vs %s: %s within-1 %.5f" % (what, key, name, (d[m] <= 1).mean())
            assert (d[m] == 0).mean() >= exact, "%s vs %s: %s exact %.5f" % (what, key, name, (d[m] == 0).mean())
        # silhouette of the mesh + bolt against the sky: the stale ambient term makes every body pixel non-black
        body = interior & (cls == 4)
        body[150:] = False                         # rows above the horizon only (below it the plane is not black)
        assert body.sum() > 8000
        assert (ref[body].sum(-1) > 0).mean() >= 0.999, "%s: silhouette" % key
        assert (ref[sky].sum(-1) == 0).all()


def test_fixture_is_the_reference_png():
    ref = "/root/reference/tests/tie_fighter/download (10).png"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(ref).convert("RGB")), np.load(FIXTURE)["shot10"])


def test_area_light_term_is_stale():
    """G - R on plane pixels (the green area light's contribution) differs between the reference's own two screenshots by
    more than any sampling noise could explain (16 spp x 4 light samples over ~195 000 pixels): they were not rendered
    from one scene, so neither pins the area light of today's tests/tie_fighter/test.mjs:12-20."""
    _, _, cls, interior = _setup()
    shots = np.load(FIXTURE)
    g_minus_r = {k: shots[k][..., 1].astype(np.int32) - shots[k][..., 0] for k in ("shot10", "shot11")}
    for c, lo in ((1, 1.8), (2, 2.5)):
        m = interior & (cls == c)
        a, b = g_minus_r["shot10"][m].mean(), g_minus_r["shot11"][m].mean()
        assert b / a > lo, (c, a, b)
        # ... while R and B agree between them to within one grey level on the same pixels (the pinned part)
        d = np.abs(shots["shot10"][..., [0, 2]].astype(np.int32) - shots["shot11"][..., [0, 2]])[m]
        assert d.max() <= 4 and (d <= 1).mean() > 0.99


def test_oracle_matches_reference_screenshot():
    from oracle.oracle import resolve_rgba8
    orc = _setup()[0]
    acc, _ = orc.render(16, seed=1)
    _check(resolve_rgba8(acc, 16), "oracle")


@pytest.mark.gpu
def test_cuda_path_matches_reference_screenshot():
    from jsraytracer_b200 import lib
    _, mp = scene_blobs("tie_fighter", width=W, height=H)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    sc.render(0, 16, seed=1)
    _check(sc.resolve_rgba8(), "cuda")
