"""BASELINE-size checks through size-independent properties (the oracle is too slow to
render these sizes many times): determinism of the ray census, linearity of the
accumulation in the pass range, statistical agreement with an oracle render at reduced
resolution."""
import numpy as np
import pytest

from conftest import scene_blobs

pytestmark = pytest.mark.gpu


def test_bunny_path_1080p_properties():
    from jsraytracer_b200 import lib
    js, mp = scene_blobs("bunny_path", width=1920, height=1080, aspect=16 / 9)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    assert sc.info["n_nodes"] == 9935 and sc.info["n_tris"] == 4968 and sc.info["max_bvh_depth"] == 15
    sc.stats_reset()
    sc.render(0, 4, seed=1)
    a, pa = sc.read_accum()
    s1 = sc.stats()
    sc.reset_accum(); sc.stats_reset()
    sc.render(0, 1, seed=1); sc.render(1, 3, seed=1)
    b, pb = sc.read_accum()
    s2 = sc.stats()
    assert pa == pb == 4 and np.all(a[..., 3] == 4)
    for k in ("rays_primary", "rays_secondary", "rays_shadow", "shaded_hits"):
        assert s1[k] == s2[k], k                               # the ray census is deterministic
    assert s1["rays_primary"] == 4 * 1920 * 1080
    assert np.allclose(a, b, rtol=1e-4, atol=1e-4)             # atomic order only
    assert np.isfinite(a).all()


def test_bunny_path_converges_to_oracle_mean():
    """Path-traced convergence (north_star: "within a stated per-pixel 3 sigma tolerance of a high-spp
    reference render").  Reference R = 512-spp oracle render with its per-pixel sample variance;
    stated tolerance: |mean_64 - R| <= 3 * sqrt(var/64 + var/512) + 2e-3 on every channel.  The GPU's
    64-spp mean (its own seed) must satisfy it on >= 97 % of pixels and on no fewer pixels (-1 %)
    than an independent 64-spp *oracle* render does — heavy-tailed pixels (specular paths) miss a
    Gaussian 3 sigma bound for the reference itself, so the control run sets the achievable rate."""
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    W, H = 160, 90
    kw = dict(width=W, height=H, aspect=16 / 9)
    js, mp = scene_blobs("bunny_path", **kw)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    orc = OracleScene(js)
    n, nref = 64, 512
    s1 = np.zeros((H, W, 3)); s2 = np.zeros((H, W, 3))
    for p in range(nref):
        acc, _ = orc.render(1, first_pass=p, seed=99)
        s1 += acc; s2 += acc.astype(np.float64) ** 2
    ref = s1 / nref
    var = np.maximum(s2 / nref - ref ** 2, 0)
    tol = 3 * np.sqrt(var / n + var / nref) + 2e-3
    sc.render(0, n, seed=11)
    g = sc.read_accum()[0][..., :3] / n
    ctrl = orc.render(n, seed=12345)[0] / n
    frac_gpu = float((np.abs(g - ref) <= tol).all(axis=-1).mean())
    frac_ctrl = float((np.abs(ctrl - ref) <= tol).all(axis=-1).mean())
    assert frac_gpu >= 0.97 and frac_gpu >= frac_ctrl - 0.01, (frac_gpu, frac_ctrl)
    assert abs(float(g.mean()) - float(ref.mean())) < 0.01 * float(ref.mean()) + 1e-3
