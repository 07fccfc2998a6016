"""Single calls of the reference's own functions (run in oracle/jsvm; generator oracle/refjs_probes.py, fixture
tests/golden/probes_refjs.npz) against the oracle's restatements, for equality: `World.cast` on random rays — closest hit
and shadow-ray windows, normalised and unnormalised directions — through analytic scenes, BVH meshes and SDF scenes;
`root_sdf.distance` and the SDF normal at random points; `Math.fmod` on arguments that include exact decimal ties."""
import json
import os
import zlib

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Z = np.load(os.path.join(GOLDEN, "probes_refjs.npz"))
META = json.loads(str(Z["meta"]))
CAST = sorted(k[5:-5] for k in Z.files if k.startswith("cast_") and k.endswith("_rays"))
SDF = sorted(k[4:-2] for k in Z.files if k.startswith("sdf_") and k.endswith("_p"))


def _scene(name):
    from oracle.oracle import OracleScene
    z = np.load(os.path.join(GOLDEN, "refjs_%s.npz" % name))
    return OracleScene(zlib.decompress(z["json"].tobytes()).decode("utf8"))


def _eq(a, b):
    return a == b or (a != a and b != b)


@pytest.mark.parametrize("name", CAST)
def test_world_cast_equals_reference(name):
    sc = _scene(name)
    rays, flags, want = Z["cast_%s_rays" % name], Z["cast_%s_flags" % name], Z["cast_%s_t" % name]
    assert np.isfinite(want).sum() > len(want) // 5, "the probe rays must hit something"
    bad = []
    for r, f, t in zip(rays, flags, want):
        _, got = sc.cast(r[0:3], r[3:6], min_d=r[6], max_d=r[7], shadow=not f)
        if not _eq(got, t):
            bad.append((r.tolist(), int(f), float(t), got))
    assert not bad, "%d of %d rays differ from the reference's World.cast: %s" % (len(bad), len(want), bad[:3])


@pytest.mark.parametrize("name", SDF)
def test_sdf_distance_and_normal_equal_reference(name):
    sc = _scene(name)
    prim = META["sdf_%s_prim" % name]
    pts, want = Z["sdf_%s_p" % name], Z["sdf_%s_out" % name]
    bad = []
    for p, w in zip(pts, want):
        d, out = sc.sdf_probe(prim, p)
        if not (_eq(d, w[0]) and all(_eq(float(out[i]), w[1 + i]) for i in range(3))):
            bad.append((p.tolist(), w.tolist(), d, out[:3].tolist()))
    assert not bad, "%d of %d points differ from the reference: %s" % (len(bad), len(want), bad[:3])


def test_math_fmod_equals_reference_including_exact_ties():
    from oracle import oracle as orc
    L = orc.lib()
    a, b, want = Z["fmod_a"], Z["fmod_b"], Z["fmod_out"]
    got = np.array([L.orc_kat_fmod(float(x), float(y)) for x, y in zip(a, b)])
    assert np.array_equal(got, want, equal_nan=True)
    # the fixture does contain the case the pin found: printf's round-half-even would give a different answer
    even = np.array([float("%.7e" % (x - np.floor(x / y) * y)) if np.isfinite(x - np.floor(x / y) * y) and (x - np.floor(x / y) * y) != 0 else 0.0
                     for x, y in zip(a, b)])
    assert (even != want).sum() >= 20
    from jsraytracer_b200 import jsmath
    assert all(_eq(jsmath.fmod(float(x), float(y)), float(w)) for x, y, w in zip(a, b, want))
