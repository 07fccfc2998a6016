"""The oracle pinned to THE REFERENCE ITSELF.

tests/golden/refjs_<scene>.npz hold what the reference's own, unmodified sources computed for its own demo scenes
(/root/reference/src/*.js + tests/<scene>/test.mjs executed by oracle/jsvm; generator: oracle/refjs_golden.py): the
scene in the reference's wire format, the f32 colour its render loop handed to PixelBuffer.setColor, and the ImageData
it filled.  Math.random() was a per-sample splitmix64 tape; the oracle's tape mode draws from the same tape, so every
random decision (pixel jitter, lens sample, area-light point, Fresnel / path-tracing lobe choice, scatter direction)
must be taken in the reference's order with the reference's arithmetic for the colours to agree.

The bar is equality of every f32 and of every byte — not a tolerance.
"""
import glob
import json
import os
import zlib

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refjs_*.npz")))


def _load(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    js = zlib.decompress(z["json"].tobytes()).decode("utf8")
    return meta, js, z


def _same(a, b):
    """equal f32 arrays, NaN == NaN (a NaN pixel of the reference must be a NaN pixel of the oracle)"""
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return np.array_equal(a, b, equal_nan=True)


def _oracle_mean(js, meta, simple=False):
    from oracle.oracle import OracleScene
    sc = OracleScene(js)
    W, H = meta["width"], meta["height"]
    no_jitter = simple or meta["renderer"] == "SimpleRenderer"
    n = 1 if no_jitter else meta["passes"]
    acc, _ = sc.render(n, seed=meta["seed"], jitter=not no_jitter, width=W, height=H, threads=1, tape=True)
    # `buffer[px][py].times(1 / (iter + 1))` (src/renderers.js:98): Vec.times stores f32(x * s), s = 1 / n in f64
    mean = (acc.astype(np.float64) * (1.0 / n)).astype(np.float32)
    return acc, mean, n


def test_fixtures_cover_the_path():
    names = {os.path.basename(p)[6:-4] for p in FIXTURES}
    # every part of SURVEY §8(a): analytic primitives, BVH meshes with vertex normals, SDF march + operators, Phong / Fresnel /
    # path-tracing materials, point + area lights, pinhole + depth-of-field cameras
    need = {"BoxBall", "BoxBall_DOF", "ASimpleScene", "Aggregates", "AHollowTetrahedron", "refraction", "cornell_box_path",
            "bunny", "bunny_path", "SDF_Menger", "SDF_Sierpinski", "SDF_Combinations", "cornell_box"}
    assert need <= names, "missing fixtures: %s" % sorted(need - names)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[6:-4] for p in FIXTURES])
def test_oracle_equals_reference_output(path):
    from oracle.oracle import resolve_rgba8
    meta, js, z = _load(path)
    acc, mean, n = _oracle_mean(js, meta)
    ref = z["mean"]
    bad = ~((mean == ref) | (np.isnan(mean) & np.isnan(ref)))
    assert not bad.any(), "%d of %d pixels differ from the reference's own output (max |d| %.3e)" % (
        int(bad.any(-1).sum()), ref.shape[0] * ref.shape[1], float(np.nanmax(np.abs(mean.astype(np.float64) - ref))))
    assert np.array_equal(resolve_rgba8(acc, n), z["rgba8"]), "8-bit image differs from the reference's ImageData"
    if "random_mean" in z.files:       # the same world through the reference's RandomMultisamplingRenderer (src/renderers.js:47-63)
        from oracle.oracle import OracleScene
        spp = int(z["random_spp"])
        col, _ = OracleScene(js).render(spp, seed=meta["seed"], width=meta["width"], height=meta["height"], threads=1, tape=True,
                                        random_multisampling=True)
        assert _same(col, z["random_mean"])
        assert np.array_equal(resolve_rgba8(col, 1), z["random_rgba8"])
    if "simple_mean" in z.files:       # the same world through the reference's SimpleRenderer (no jitter)
        sacc, smean, _ = _oracle_mean(js, meta, simple=True)
        assert _same(smean, z["simple_mean"])
        assert np.array_equal(resolve_rgba8(sacc, 1), z["simple_rgba8"])


def test_fixtures_exercise_random_decisions():
    """the draw counts recorded from the reference show what the tape pinned: jitter only (2), jitter + lens (4),
    and long path-tracing chains"""
    draws = {}
    for p in FIXTURES:
        meta, _, z = _load(p)
        draws[meta["name"]] = (int(z["draws"].min()), int(z["draws"].max()))
    assert draws["BoxBall"] == (2, 2)
    assert draws["BoxBall_DOF"] == (4, 4)
    assert draws["AHollowTetrahedron"] == (0, 0)
    assert draws["cornell_box_path"][1] >= 50
    assert draws["bunny_path"][1] >= 10


def test_random_renderer_fixtures_exist():
    assert sum("random_mean" in np.load(p).files for p in FIXTURES) >= 2


def test_live_reference_run_matches_oracle():
    """where the reference tree is present (the build container): run it now, not from a fixture"""
    from oracle import refjs
    if not refjs.available():
        pytest.skip("no reference tree on this machine: the committed fixtures stand in")
    from oracle.oracle import OracleScene, resolve_rgba8
    r = refjs.RefJS()
    r.load_test("BoxBall_DOF")
    W, H, P = 9, 6, 2
    js = r.scene_json(W, H)
    mean, rgba, draws = r.render(W, H, P, seed=7)
    assert draws.min() == 4 and draws.max() == 4
    acc, _ = OracleScene(js).render(P, seed=7, width=W, height=H, threads=1, tape=True)
    assert _same((acc.astype(np.float64) * (1.0 / P)).astype(np.float32), mean)
    assert np.array_equal(resolve_rgba8(acc, P), rgba)
    # one worker of three: columns 1, 4, 7 (src/worker.js:30-32 -> render(img, ..., workerIndex, workerCount))
    mean2, rgba2, _ = r.render(W, H, P, x_offset=1, x_delt=3, seed=7)
    acc2, _ = OracleScene(js).render(P, seed=7, width=W, height=H, threads=1, tape=True, x_offset=1, x_delt=3)
    assert _same((acc2.astype(np.float64) * (1.0 / P)).astype(np.float32), mean2)
    assert np.array_equal(mean2[:, 1::3], mean[:, 1::3]) and not mean2[:, 0::3].any() and not rgba2[:, 0::3].any()


def test_fixtures_are_current():
    """a fixture is only evidence about the sources it was made from"""
    from oracle import refjs
    if not refjs.available():
        pytest.skip("no reference tree on this machine")
    from oracle.refjs_golden import sources_digest
    for p in FIXTURES:
        meta, _, _ = _load(p)
        assert meta["sources_sha256"] == sources_digest(refjs.REF_ROOT, meta["name"]), meta["name"]
