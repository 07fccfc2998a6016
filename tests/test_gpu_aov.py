"""JSRT_FLAG_AOV: the GL path's auxiliary buffers — first-hit normal / distance sums and the
running per-pixel variance (reference: gl/src/WebGLRendererAdapter.js:352-356,376-379) — checked against numpy
restatements of the shader's recurrences on per-pass sample images rendered by the same path."""
import numpy as np
import pytest

from conftest import scene_blobs

pytestmark = pytest.mark.gpu


def gl_variance(samples):
    """outSampleVariance of gl/src/WebGLRendererAdapter.js:352-356 over a list of per-pass images, in FP32 like the shader."""
    s = np.zeros_like(samples[0], dtype=np.float32)
    v = np.zeros_like(s)
    for n, c in enumerate(samples):
        c = c.astype(np.float32)
        count1 = np.float32(n + 1)
        mean = s / count1
        delta = c - mean
        delta2 = c - (delta / count1 + mean)
        v = v + delta * delta2
        s = s + c
    return s, v


@pytest.mark.parametrize("name,kw", [("BoxBall_path", dict(width=96, height=64)), ("bunny_path", dict(width=120, height=68, aspect=120 / 68))])
def test_variance_and_image_match_per_pass_samples(name, kw):
    from jsraytracer_b200 import lib
    _, mp = scene_blobs(name, **kw)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    N = 6
    samples = []
    for p in range(N):                                    # the radiance of sample p of every pixel, on the plain path
        sc.reset_accum()
        sc.render(p, 1, seed=5)
        a, _ = sc.read_accum()
        assert np.all(a[..., 3] == 1)
        samples.append(a[..., :3].copy())
    ref_sum, ref_var = gl_variance(samples)
    scale = max(1.0, float(np.abs(ref_sum).max()))
    for chunks in ([N], [1] * N, [2, 4]):                 # one batch, one call per pass, uneven calls: same buffers
        sc.reset_accum()
        first = 0
        for n in chunks:
            sc.render(first, n, seed=5, flags=lib.FLAG_AOV)
            first += n
        a, passes = sc.read_accum()
        nd, var = sc.read_aov()
        assert passes == N and np.all(a[..., 3] == N)
        assert np.allclose(a[..., :3], ref_sum, rtol=2e-5, atol=2e-5 * scale)                 # same image (summation order only)
        assert np.allclose(var[..., :3], ref_var, rtol=2e-3, atol=2e-4 * scale * scale), chunks
    # a plain render after AOV renders still accumulates into the pixel sums
    sc.reset_accum()
    sc.render(0, N, seed=5)
    b, _ = sc.read_accum()
    assert np.allclose(b[..., :3], ref_sum, rtol=2e-5, atol=2e-5 * scale)


def test_first_hit_normal_and_distance():
    from jsraytracer_b200 import lib, scenes
    W, H = 128, 96
    test = scenes.configure("BoxBall", width=W, height=H, aspect=W / H)
    _, mp = scene_blobs("BoxBall", width=W, height=H, aspect=W / H)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    ids, t = sc.primary_hits()
    P = 3
    sc.render(0, P, seed=1, flags=lib.FLAG_AOV | lib.FLAG_NO_JITTER)      # un-jittered: every pass repeats the primary_hits ray
    nd, var = sc.read_aov()
    hit = ids >= 0
    assert np.array_equal(var[..., 3], np.where(hit, P, 0).astype(np.float32))
    # |d| of the un-normalised pinhole ray (src/cameras.js:29-34) -> distance = |o + d t - o|
    cam = test["renderer"].camera
    T = np.array([[cam.transform[r][c] for c in range(4)] for r in range(4)], dtype=np.float64)
    xs = 2.0 * np.arange(W) / W - 1.0
    ys = -2.0 * np.arange(H) / H + 1.0
    X, Y = np.meshgrid(xs, ys)
    d_cam = np.stack([X * cam.tan_fov * cam.aspect, Y * cam.tan_fov, -np.ones_like(X), np.zeros_like(X)], axis=-1)
    d_world = (d_cam @ T.T)[..., :3]
    dist = np.linalg.norm(d_world, axis=-1) * t
    assert np.allclose(nd[..., 3][hit] / P, dist[hit], rtol=2e-4)
    assert np.all(nd[..., 3][~hit] == 0)
    # world normals: unit length everywhere, (0, 1, 0) on the floor plane (prim 0 of tests/BoxBall/test.mjs:16-19)
    n = nd[..., :3] / P
    assert np.allclose(np.linalg.norm(n[hit], axis=-1), 1.0, atol=1e-4)
    floor = ids == 0
    assert floor.any() and np.allclose(np.abs(n[floor]), np.array([0, 1, 0]), atol=1e-5)
    # reset clears the buffers
    sc.reset_accum()
    nd2, var2 = sc.read_aov()
    assert not nd2.any() and not var2.any()
