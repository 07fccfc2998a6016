"""Known answers for the numpy restatement of the GL path's denoiser (oracle/oracle.py: denoise_passthrough; reference
gl/src/WebGLRendererAdapter.js:183-246), derived by hand from the shader text."""
import numpy as np

from oracle.oracle import denoise_passthrough


def _const(H=9, W=11, n=4, rgb=(0.2, 0.4, 0.6)):
    a = np.zeros((H, W, 4), np.float32)
    a[..., :3] = np.array(rgb, np.float32) * n
    a[..., 3] = n
    return a, np.zeros((H, W, 4), np.float32)


def test_constant_image_is_a_fixed_point():
    a, v = _const()
    out = denoise_passthrough(a, v)
    assert np.allclose(out[..., :3], [0.2, 0.4, 0.6], atol=1e-6) and np.allclose(out[..., 3], 1.0, atol=1e-6)


def test_window_taps_follow_the_shader_loops():
    """radius = round(2 * 1) = 2; d.x = -2..2; d.y runs from -sqrt(4 - d.x^2) in unit steps, so the off-centre columns are
    sampled at NON-integer row offsets (d.x = +-1: -1.73, -0.73, 0.27, 1.27 -> texel rows y-2, y-1, y, y+1 under NEAREST).
    An impulse at the centre therefore spreads to exactly the pixels whose window contains it."""
    H = W = 9
    a, v = _const(H, W, n=1, rgb=(0, 0, 0))
    a[4, 4, :3] = 1.0
    out = denoise_passthrough(a, v)[..., 0]
    reached = {(int(y) - 4, int(x) - 4) for y, x in zip(*np.nonzero(out > 0))}
    expect = set()
    for dx in (-2, -1, 0, 1, 2):
        pt = np.sqrt(np.float32(4 - dx * dx))
        dy = -pt
        while dy <= pt:
            # pixel (y, x) reads texel (floor(y + 0.5 + dy), x + dx): the impulse reaches (4 - row offset, 4 - dx)
            expect.add((-int(np.floor(0.5 + dy)), -dx))
            dy += 1
    assert reached == expect


def test_noisy_taps_are_weighted_down_by_their_own_standard_deviation():
    """deltaFactor = exp(-|std(tap)|^2 / (2 threshold^2)) ... (the shader's active line :210): a tap with a large variance
    contributes less, independent of the centre pixel's variance."""
    a, v = _const(5, 5, n=4, rgb=(0.5, 0.5, 0.5))
    a[2, 3, :3] = 4 * 1.0                       # one bright neighbour ...
    hi = v.copy(); hi[2, 3, :3] = 4 * 9.0       # ... whose variance is 9 (std 3 per channel)
    calm = denoise_passthrough(a, v, threshold=1.0)[2, 2, 0]
    damp = denoise_passthrough(a, hi, threshold=1.0)[2, 2, 0]
    assert calm > damp > 0.5 - 1e-6
    w = np.exp(-27.0 / 2.0)                     # |std|^2 = 3 * 9, threshold 1
    assert damp - 0.5 < (calm - 0.5) * w * 1.5


def test_diagnostic_colours_and_log_scale():
    a, v = _const(4, 4, n=1, rgb=(0.5, 0.25, 1.0))
    out = denoise_passthrough(a, v, color_log_scale=2.0)
    assert np.allclose(out[..., :3], np.log(np.array([1.5, 1.25, 2.0])) / 2.0, atol=1e-6)
    a[...] = 0; a[..., 3] = 1; a[1, 1, 0] = -50.0
    out = denoise_passthrough(a, v, sigma=0.4, k_sigma=1.0)          # radius round(0.4) = 0: only the centre tap
    assert np.allclose(out[1, 1, :3], [0.5, 0.0, 1.0]) and np.allclose(out[0, 0, :3], 0)
