"""GPU tests of the round-2 kernel paths, all through the C ABI and against the oracle (or against the unfused path):

  * shadow tests fused into shade_kernel + walker-only shadow queue, camera rays generated inside prims_kernel<extend>,
    BVH top levels staged in shared memory: each switched off in turn must give the same rays and the same image;
  * optimistic queue sizing: a batch that overflows its queues is re-run smaller, the image is unchanged;
  * maxRecursionDepth beyond 32 (path-tree node ids no longer wrap);
  * aggregates nested the ways src/aggregates.js:14-18,43-49 allows;
  * one scene handle on two devices (jsrt_scene_create with ndev = 2): same image as one device;
  * kernel builds specialised on the scene's content (shade_kernel<LEAN>, bvh_kernel<MESH>) and the fused SDF instructions:
    same rays, same image as the general builds / the plain bytecode.
"""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from conftest import psnr, scene_blobs

pytestmark = pytest.mark.gpu


@contextmanager
def env(**kv):
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update({k: str(v) for k, v in kv.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def _scene(mp, **kv):
    from jsraytracer_b200 import lib
    with env(**kv):
        return lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)


def _render(sc, passes, seed=3):
    sc.stats_reset()
    sc.render(0, passes, seed=seed)
    acc, n = sc.read_accum()
    assert n == passes
    return acc, sc.stats()


CENSUS = ("rays_primary", "rays_secondary", "rays_shadow", "shaded_hits")


@pytest.mark.parametrize("name,kw", [
    ("bunny_path", dict(width=480, height=270, aspect=16 / 9)),      # plane + one mesh, 2 point lights
    ("cornell_box_path", dict(width=160, height=160)),               # analytic primitives only, 4 area-light samples
    ("starwars", dict(width=320, height=180, aspect=16 / 9)),        # 4 BVH instances, area + point lights, DOF
    ("AMultipleBVH", dict(width=192, height=192)),                   # a sphere that casts no shadow
], ids=["bunny_path", "cornell_box_path", "starwars", "AMultipleBVH"])
def test_fused_paths_equal_the_unfused_ones(name, kw):
    _, mp = scene_blobs(name, **kw)
    ref, sref = _render(_scene(mp, JSRT_FUSE_SHADOW=0, JSRT_FUSE_GEN=0, JSRT_STAGE_NODES=0), 3)
    for toggles in (dict(), dict(JSRT_FUSE_SHADOW=0), dict(JSRT_FUSE_GEN=0), dict(JSRT_STAGE_NODES=0), dict(JSRT_STAGE_NODES=37)):
        acc, st = _render(_scene(mp, **toggles), 3)
        for k in CENSUS:
            assert st[k] == sref[k], (toggles, k, st[k], sref[k])
        assert np.allclose(acc, ref, rtol=1e-5, atol=1e-5), toggles           # FP32 summation order only


@pytest.mark.parametrize("name,kw", [
    ("bunny_path", dict(width=480, height=270, aspect=16 / 9)),      # LEAN 2 + MESH: plane + mesh, point lights
    ("starwars", dict(width=320, height=180, aspect=16 / 9)),        # LEAN 1 + MESH: spheres + area light beside the meshes
    ("cornell_box_path", dict(width=160, height=160)),               # LEAN 1, sorted tiles, no BVH
    ("AHollowTetrahedron", dict(width=192, height=192)),             # two aggregates sharing one kdtree
], ids=["bunny_path", "starwars", "cornell_box_path", "AHollowTetrahedron"])
def test_specialised_builds_equal_the_general_ones(name, kw):
    """shade_kernel<LEAN> and bvh_kernel<MESH> only drop code the scene cannot reach: same ray census, same image as the general
    builds (JSRT_SHADE_LEAN=0 JSRT_BVH_MESH=0) up to FP32 summation order."""
    _, mp = scene_blobs(name, **kw)
    ref, sref = _render(_scene(mp, JSRT_SHADE_LEAN=0, JSRT_BVH_MESH=0), 3)
    for toggles in (dict(), dict(JSRT_SHADE_LEAN=1), dict(JSRT_BVH_MESH=0), dict(JSRT_SHADE_LEAN=0)):
        acc, st = _render(_scene(mp, **toggles), 3)
        for k in CENSUS:
            assert st[k] == sref[k], (toggles, k, st[k], sref[k])
        assert np.allclose(acc, ref, rtol=1e-5, atol=1e-5), toggles


def test_sdf_fused_instructions_equal_the_plain_program():
    """S_CROSS / S_RTU_CROSS and the world-box reject are bit-for-bit the plain instruction stream: SDF_Menger rendered with the
    peepholes off (JSRT_SDF_FUSE=0) gives the same hits, distances and image."""
    from jsraytracer_b200 import lib
    _, mp = scene_blobs("SDF_Menger", width=128, height=128)
    plain = _scene(mp, JSRT_SDF_FUSE=0)
    fused = _scene(mp)
    assert fused.info["n_sdf_instrs"] == 8 and plain.info["n_sdf_instrs"] > 45
    ids0, t0 = plain.primary_hits()
    ids1, t1 = fused.primary_hits()
    assert np.array_equal(ids0, ids1) and np.array_equal(t0, t1)
    plain.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    fused.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    a0, a1 = plain.read_accum()[0], fused.read_accum()[0]
    assert np.allclose(a0, a1, rtol=1e-6, atol=1e-7)                  # (the same arithmetic; only the order of the FP32 reductions may differ)


def test_staged_walk_hit_ids_dragon():
    """Octant layouts + staged prefixes on the 199 935-node tree: hit IDs do not depend on how much is staged."""
    _, mp = scene_blobs("dragon", width=480, height=270, aspect=16 / 9)
    ids0, t0 = _scene(mp, JSRT_STAGE_NODES=0).primary_hits()
    for n in (4096, 7168, 100):
        ids, t = _scene(mp, JSRT_STAGE_NODES=n).primary_hits()
        assert np.array_equal(ids, ids0) and np.array_equal(t, t0), n


def test_queue_overflow_is_rerun_smaller():
    """JSRT_QUEUE_GROWTH=1: queues as long as the batch itself, so every level that grows overflows; the library must
    notice, re-run the batch with fewer camera samples and deliver the same image and ray census."""
    from oracle.oracle import OracleScene
    js, mp = scene_blobs("refraction", width=160, height=160)
    ref, sref = _render(_scene(mp), 2)
    tight = _scene(mp, JSRT_QUEUE_GROWTH=1, JSRT_BATCH_PASSES=2)
    acc, st = _render(tight, 2)
    for k in CENSUS:
        assert st[k] == sref[k], (k, st[k], sref[k])
    assert np.allclose(acc, ref, rtol=1e-5, atol=1e-5)
    assert sref["rays_secondary"] > sref["rays_primary"]              # the scene does fan out (Fresnel spheres)
    oacc, _ = OracleScene(js).render(2, seed=3)
    assert psnr(np.clip(acc[..., :3] / 2, 0, 1), np.clip(oacc / 2, 0, 1)) >= 50.0


def test_recursion_depth_beyond_32():
    """maxRecursionDepth = 40 with a fan-out of 2 (Fresnel spheres): worst-case queues (2^39 rays per sample) cannot be
    allocated, so this runs on optimistic sizing, and the path-tree node ids go through the hashed range (rng.h)."""
    sc_kw = dict(width=96, height=96, depth=40)
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    js, mp = scene_blobs("refraction", **sc_kw)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    assert sc.info["max_depth"] == 40
    acc, st = _render(sc, 2, seed=8)
    oacc, cnt = OracleScene(js).render(2, seed=8)
    assert psnr(np.clip(acc[..., :3] / 2, 0, 1), np.clip(oacc / 2, 0, 1)) >= 50.0
    assert st["rays_primary"] == cnt["rays_primary"]
    for k in ("rays_secondary", "rays_shadow"):
        assert abs(st[k] - cnt[k]) <= 5e-3 * max(1, cnt[k]), (k, st[k], cnt[k])


def test_nested_aggregates():
    """A BVHAggregate inside a plain Aggregate and a BVHAggregate over BVHAggregate instances + a Primitive."""
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    js, mp = scene_blobs("nested_aggregates", width=256, height=256)
    sc, orc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0), OracleScene(js)
    ids, t = sc.primary_hits()
    oids, ot, _ = orc.primary_hits()
    assert float((ids == oids).mean()) >= 0.9999
    both = (ids == oids) & (oids >= 0)
    rel = np.abs(t[both].astype(np.float64) - ot[both]) / np.abs(ot[both])
    assert float((rel > 1e-4).mean()) <= 1e-4
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    g = np.clip(sc.read_accum()[0][..., :3], 0, 1)
    o = np.clip(orc.render(1, seed=1, jitter=False)[0], 0, 1)
    assert psnr(g, o) >= 50.0, "PSNR %.2f dB" % psnr(g, o)


def test_two_devices_one_handle():
    """jsrt_scene_create(devices = {0, 1}): passes dealt to both GPUs, summed over NVLink inside the resolve kernel."""
    from jsraytracer_b200 import lib
    if lib.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _, mp = scene_blobs("bunny_path", width=480, height=270, aspect=16 / 9)
    one = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    two = lib.Scene(mp, lib.FORMAT_MSGPACK, device=[0, 1])
    a, sa = _render(one, 5)
    b, sb = _render(two, 5)
    for k in CENSUS:
        assert sa[k] == sb[k], k
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert np.all(b[..., 3] == 5)
    ia, ib = one.resolve_rgba8().astype(int), two.resolve_rgba8().astype(int)
    assert np.abs(ia - ib).max() <= 1


def test_top_level_bvh_equals_linear_walk():
    """dragon_grid (8 instances of one kdtree): the top-level BVH over the aggregates (scene_flatten.cpp: buildTlas) only
    prunes — hit IDs, ray census and image equal those of the reference's linear walk over world.objects."""
    _, mp = scene_blobs("dragon_grid", width=320, height=180, aspect=16 / 9, n=2)
    lin, tl = _scene(mp, JSRT_TLAS_MIN=0), _scene(mp, JSRT_TLAS_MIN=2)
    ids0, t0 = lin.primary_hits()
    ids1, t1 = tl.primary_hits()
    assert np.array_equal(ids0, ids1) and np.array_equal(t0, t1)
    a, sa = _render(lin, 2)
    b, sb = _render(tl, 2)
    for k in CENSUS:
        assert sa[k] == sb[k], (k, sa[k], sb[k])
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
