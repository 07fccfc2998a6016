"""Parity at the BASELINE.json configurations that round 1 only benchmarked ("What's missing" 2-3, "What's weak"):

  * configs[3] as stated: SDF_Menger / SDF_Sierpinski seen through the depth-of-field camera of tests/BoxBall_DOF
    (`DepthOfFieldPerspectiveCamera(pi/4, 16/9, BoxBall transform, 9.2, 0.2)`, /root/reference/tests/BoxBall_DOF/test.mjs:3-8);
  * configs[2] dragon: the shaded Whitted image and a jittered same-RNG sample set — secondary rays on a >= 32 768-node
    tree take the octant layouts that primary-hit parity never reaches;
  * configs[1] cornell_box_path rendered in several wavefront batches whose boundaries fall inside a frame (the way its
    1024^2 x 64 spp BASELINE size runs), compared with the single-batch render and with the oracle;
  * a per-pixel 3-sigma convergence test for cornell_box_path (area light, IOR-2 sphere, depth 8).

All through the C ABI, all against the CPU restatement oracle on the same counter-based RNG.
"""
import os

import numpy as np
import pytest

from conftest import psnr, scene_blobs

pytestmark = pytest.mark.gpu


def _pair(name, **kw):
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    js, mp = scene_blobs(name, **kw)
    return lib.Scene(mp, lib.FORMAT_MSGPACK, device=0), OracleScene(js)


@pytest.mark.parametrize("name", ["SDF_Menger", "SDF_Sierpinski"])
def test_sdf_with_depth_of_field_camera_same_rng(name):
    """BASELINE configs[3]: SDF march + lens sampling (Vec.circlePick, src/math.js:175-179) + jitter, same RNG both sides."""
    sc, orc = _pair(name, width=192, height=108, aspect=16 / 9, dof=(9.2, 0.2))
    passes = 2
    sc.stats_reset()
    sc.render(0, passes, seed=5)
    acc, _ = sc.read_accum()
    st = sc.stats()
    oacc, cnt = orc.render(passes, seed=5)
    g, o = np.clip(acc[..., :3] / passes, 0, 1), np.clip(oacc / passes, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
    assert st["rays_primary"] == cnt["rays_primary"] == passes * 192 * 108
    for k in ("rays_secondary", "rays_shadow"):
        assert abs(st[k] - cnt[k]) <= 2e-3 * max(1, cnt[k]), (k, st[k], cnt[k])


def test_dragon_whitted_image():
    """Deterministic (SimpleRenderer sampling) dragon: reflection + shadow rays through the 199 935-node tree."""
    from jsraytracer_b200 import lib
    sc, orc = _pair("dragon", width=480, height=270, aspect=16 / 9)
    assert sc.info["n_nodes"] == 199935
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    acc, _ = sc.read_accum()
    oacc, _ = orc.render(1, seed=1, jitter=False)
    g, o = np.clip(acc[..., :3], 0, 1), np.clip(oacc, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)


def test_dragon_same_rng_two_passes():
    sc, orc = _pair("dragon", width=480, height=270, aspect=16 / 9)      # same blob as the Whitted case (cached)
    sc.stats_reset()
    sc.render(0, 2, seed=9)
    acc, _ = sc.read_accum()
    st = sc.stats()
    oacc, cnt = orc.render(2, seed=9)
    g, o = np.clip(acc[..., :3] / 2, 0, 1), np.clip(oacc / 2, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
    assert st["rays_primary"] == cnt["rays_primary"]
    for k in ("rays_secondary", "rays_shadow"):
        assert abs(st[k] - cnt[k]) <= 1e-3 * max(1, cnt[k]), (k, st[k], cnt[k])


def test_cornell_multi_batch_waves_match_single_batch_and_oracle():
    """JSRT_QUEUE_BYTES forced to its minimum: 65 536-sample batches on a 320 x 240 frame, so one 3-pass call is split into
    four batches with every boundary inside a frame (render.cu: `for (done = 0; done < total; done += batch)`)."""
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    kw = dict(width=320, height=240, aspect=4 / 3)
    js, mp = scene_blobs("cornell_box_path", **kw)
    passes = 3
    big = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    old = os.environ.get("JSRT_QUEUE_BYTES")
    os.environ["JSRT_QUEUE_BYTES"] = "1000000"
    try:
        small = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    finally:
        if old is None:
            del os.environ["JSRT_QUEUE_BYTES"]
        else:
            os.environ["JSRT_QUEUE_BYTES"] = old
    assert small.info["batch_samples"] == 65536 and big.info["batch_samples"] >= 320 * 240 * passes
    assert (320 * 240) % small.info["batch_samples"] != 0          # boundaries fall inside frames
    out = []
    for sc in (big, small):
        sc.stats_reset()
        sc.render(0, passes, seed=4)
        acc, n = sc.read_accum()
        assert n == passes and np.all(acc[..., 3] == passes)
        out.append((acc, sc.stats()))
    (a, sa), (b, sb) = out
    for k in ("rays_primary", "rays_secondary", "rays_shadow", "shaded_hits"):
        assert sa[k] == sb[k], k                                   # same rays, whatever the batching
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)                 # FP32 summation order only
    oacc, cnt = OracleScene(js).render(passes, seed=4)
    g, o = np.clip(b[..., :3] / passes, 0, 1), np.clip(oacc / passes, 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)
    assert sb["rays_primary"] == cnt["rays_primary"]


def test_cornell_box_path_converges_to_oracle_mean():
    """north_star: "path-traced images converge to within a stated per-pixel 3 sigma tolerance of a high-spp reference
    render".  Reference R = 384-spp oracle render with its per-pixel sample variance; tolerance
    |mean_64 - R| <= 3 sqrt(var / 64 + var / 384) + 2e-3 per channel.  The GPU's 64-spp mean (its own seed) must satisfy
    it on no fewer pixels (-1 %) than an independent 64-spp *oracle* render does — this scene's samples are heavy-tailed (a
    small area light seen through depth-8 paths), so a Gaussian 3 sigma bound is missed by the reference itself on ~7 % of
    the pixels (measured: control 0.930, CUDA 0.930); the control run sets the achievable rate, 90 % is the floor."""
    W = H = 64
    sc, orc = _pair("cornell_box_path", width=W, height=H)
    n, nref = 64, 384
    s1 = np.zeros((H, W, 3)); s2 = np.zeros((H, W, 3))
    for p in range(nref):
        acc, _ = orc.render(1, first_pass=p, seed=77)
        s1 += acc; s2 += acc.astype(np.float64) ** 2
    ref = s1 / nref
    var = np.maximum(s2 / nref - ref ** 2, 0)
    tol = 3 * np.sqrt(var / n + var / nref) + 2e-3
    sc.render(0, n, seed=21)
    g = sc.read_accum()[0][..., :3] / n
    ctrl = orc.render(n, seed=4242)[0] / n
    frac_gpu = float((np.abs(g - ref) <= tol).all(axis=-1).mean())
    frac_ctrl = float((np.abs(ctrl - ref) <= tol).all(axis=-1).mean())
    assert frac_gpu >= 0.90 and frac_gpu >= frac_ctrl - 0.01, (frac_gpu, frac_ctrl)
    assert abs(float(g.mean()) - float(ref.mean())) < 0.01 * float(ref.mean()) + 1e-3
