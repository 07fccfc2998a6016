"""Committed golden vectors (tests/golden/*.npz, made by tools/make_golden.py from the oracle)."""
import glob
import os

import numpy as np
import pytest

from conftest import psnr, scene_blobs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
from make_golden import CASES  # noqa: E402


@pytest.mark.parametrize("key", sorted(CASES))
def test_oracle_reproduces_golden(key):
    from oracle.oracle import OracleScene
    name, kw, wp, sp = CASES[key]
    z = np.load(os.path.join(GOLD, key + ".npz"))
    orc = OracleScene(scene_blobs(name, **kw)[0])
    ids, t, _ = orc.primary_hits()
    assert np.array_equal(ids, z["prim_id"])
    assert np.array_equal(t, z["t"])
    if wp:
        assert np.array_equal(orc.render(wp, seed=1, jitter=False)[0], z["whitted"])
    if sp:
        acc, cnt = orc.render(sp, seed=1)
        assert np.array_equal(acc, z["stochastic"])
        assert [cnt["rays_primary"], cnt["rays_secondary"], cnt["rays_shadow"]] == z["stochastic_rays"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(CASES))
def test_cuda_matches_golden(key):
    from jsraytracer_b200 import lib
    name, kw, wp, sp = CASES[key]
    z = np.load(os.path.join(GOLD, key + ".npz"))
    sc = lib.Scene(scene_blobs(name, **kw)[1], lib.FORMAT_MSGPACK, device=0)
    ids, t = sc.primary_hits()
    assert float((ids == z["prim_id"]).mean()) >= 0.9995          # small frames: one pixel is already 1e-4
    both = (ids == z["prim_id"]) & (ids >= 0)
    rel = np.abs(t[both] - z["t"][both]) / np.abs(z["t"][both])
    assert float((rel > 1e-4).mean()) <= 1e-3
    if wp:
        sc.render(0, wp, seed=1, flags=lib.FLAG_NO_JITTER)
        g = sc.read_accum()[0][..., :3] / wp
        assert psnr(np.clip(g, 0, 1), np.clip(z["whitted"] / wp, 0, 1)) >= 45.0    # tiny frames: one flipped pixel costs ~5 dB
        sc.reset_accum()
    if sp:
        sc.stats_reset()
        sc.render(0, sp, seed=1)
        g = sc.read_accum()[0][..., :3] / sp
        assert psnr(np.clip(g, 0, 1), np.clip(z["stochastic"] / sp, 0, 1)) >= 45.0
        st = sc.stats()
        got = [st["rays_primary"], st["rays_secondary"], st["rays_shadow"]]
        for a, b in zip(got, z["stochastic_rays"].tolist()):
            assert abs(a - b) <= 2e-3 * b + 2
