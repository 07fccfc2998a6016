"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU restatement oracle
(oracle/oracle.cpp).  Import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  PARITY PINNED to runs of the reference's own sources (see
oracle_math.h, tests/test_refjs_pin.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COUNTER_NAMES = ["rays_primary", "rays_secondary", "rays_shadow",
                 "bvh_nodes_primary", "bvh_nodes_secondary", "bvh_nodes_shadow",
                 "bvh_prims_primary", "bvh_prims_secondary", "bvh_prims_shadow",
                 "top_tests_primary", "top_tests_secondary", "top_tests_shadow",
                 "sdf_evals_primary", "sdf_evals_secondary", "sdf_evals_shadow", "shaded_hits",
                 # diagnostic: node records a 4-wide / 8-wide collapse of the reference's tree would fetch for the same rays
                 "wide4_primary", "wide4_secondary", "wide4_shadow", "wide8_primary", "wide8_secondary", "wide8_shadow"]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_primary_hits.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p]
        L.orc_resolve_rgba8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_kat_intersect.restype = C.c_double
        L.orc_kat_intersect.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orc_kat_material_data.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_kat_fresnel.restype = C.c_double
        L.orc_kat_fresnel.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.orc_kat_fmod.restype = C.c_double
        L.orc_kat_fmod.argtypes = [C.c_double, C.c_double]
        L.orc_kat_rng.restype = C.c_double
        L.orc_kat_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_sdf_probe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def default_threads():
    return max(1, len(os.sched_getaffinity(0)))


class OracleScene:
    """A scene loaded from the serializer wire format (JSON text)."""

    def __init__(self, json_text):
        if isinstance(json_text, str):
            json_text = json_text.encode("utf8")
        self._L = lib()
        self._h = self._L.orc_load(json_text, len(json_text))
        if not self._h:
            raise RuntimeError("oracle: " + self._L.orc_last_error().decode())
        info = (C.c_int * 8)()
        self._L.orc_info(self._h, info)
        (self.width, self.height, self.samplesPerPixel, self.maxRecursionDepth, self.nprims, self.jitter,
         self.nlights, self.ntop) = list(info)

    def close(self):
        if self._h:
            self._L.orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def primary_hits(self, width=None, height=None, threads=None):
        W, H = width or self.width, height or self.height
        ids = np.empty(W * H, dtype=np.int32)
        t = np.empty(W * H, dtype=np.float64)
        cnt = np.zeros(22, dtype=np.uint64)
        rc = self._L.orc_primary_hits(self._h, W, H, ids.ctypes.data, t.ctypes.data, threads or default_threads(),
                                      cnt.ctypes.data)
        if rc:
            raise RuntimeError("oracle: " + self._L.orc_last_error().decode())
        return ids.reshape(H, W), t.reshape(H, W), dict(zip(COUNTER_NAMES, cnt.tolist()))

    def render(self, n_passes, first_pass=0, seed=1, jitter=True, x_offset=0, x_delt=1, width=None, height=None,
               threads=None, accum=None, tape=False, random_multisampling=False):
        """Returns (sum buffer (H,W,3) f32, counters dict).  tape: the random numbers of one pixel sample come from one
        sequential stream in call order (what Math.random() is to the reference) instead of the counter-based generator
        the GPU shares — the mode the oracle is pinned to the reference's own output in (oracle/refjs.py).
        random_multisampling: RandomMultisamplingRenderer (src/renderers.js:47-63): all n_passes samples of a pixel inside
        one getPixelColor; the returned buffer then holds each pixel's colour itself."""
        W, H = width or self.width, height or self.height
        if accum is None:
            accum = np.zeros((H, W, 3), dtype=np.float32)
        cnt = np.zeros(22, dtype=np.uint64)
        rc = self._L.orc_render(self._h, W, H, first_pass, n_passes, seed, (0 if jitter else 1) | (2 if tape else 0) | (4 if random_multisampling else 0), x_offset, x_delt,
                                accum.ctypes.data, threads or default_threads(), cnt.ctypes.data)
        if rc:
            raise RuntimeError("oracle: " + self._L.orc_last_error().decode())
        return accum, dict(zip(COUNTER_NAMES, cnt.tolist()))

    def sdf_probe(self, prim_id, p):
        p = np.asarray(p, dtype=np.float64)
        dist = C.c_double()
        out = np.zeros(8, dtype=np.float64)
        rc = self._L.orc_sdf_probe(self._h, prim_id, p.ctypes.data, C.byref(dist), out.ctypes.data)
        if rc:
            raise RuntimeError("oracle: " + self._L.orc_last_error().decode())
        return dist.value, out

    def cast(self, o, d, min_d=0.0, max_d=float("inf"), shadow=False):
        o = np.asarray(o, dtype=np.float64)
        d = np.asarray(d, dtype=np.float64)
        t = C.c_double()
        pid = self._L.orc_cast(self._h, o.ctypes.data, d.ctypes.data, min_d, max_d, int(shadow), C.byref(t))
        return pid, t.value


def resolve_rgba8(accum, passes):
    H, W, _ = accum.shape
    out = np.zeros((H, W, 4), dtype=np.uint8)
    a = np.ascontiguousarray(accum, dtype=np.float32)
    lib().orc_resolve_rgba8(a.ctypes.data, H * W, passes, out.ctypes.data)
    return out


def denoise_passthrough(accum, variance, sigma=1.0, k_sigma=2.0, threshold=5.0, color_log_scale=0.0):
    """TEST INFRASTRUCTURE — numpy restatement (FP32) of the GL path's display pass with its variance-guided denoiser:
    `smartDeNoise` + `main` of the passthrough shader, gl/src/WebGLRendererAdapter.js:183-246.
    accum = (H, W, 4) colour sums with the per-pixel sample count in w (the shader's uSampleSumTexture and
    1 / texture_factor); variance = (H, W, >=3) running variance sums (uVarianceTexture, :352-356).  Textures are sampled
    NEAREST with coordinates clamped to [0, 1] (gl/src/WebGLUtilHelpers.js:452-453).  Returns (H, W, 4) f32: rgb + mean weight."""
    f = np.float32
    a = np.asarray(accum, dtype=f)
    q = np.asarray(variance, dtype=f)
    H, W = a.shape[:2]
    tf = (f(1.0) / np.maximum(a[..., 3], f(1.0))).astype(f)
    mean = (a[..., :3] * tf[..., None]).astype(f)
    std = np.sqrt((q[..., :3] * tf[..., None]).astype(f)).astype(f)
    wmean = (a[..., 3] * tf).astype(f)
    radius = f(np.rint(f(k_sigma) * f(sigma)))                      # GLSL round(); rintf on the device
    radQ = f(radius * radius)
    invSigmaQx2 = f(f(.5) / f(f(sigma) * f(sigma)))
    invSigmaQx2PI = f(f(0.31830988618379067153776752674503) * invSigmaQx2)
    invThresholdSqx2 = f(f(.5) / f(f(threshold) * f(threshold)))
    invThresholdSqrt2PI = f(f(0.39894228040143267793994605993439) / f(threshold))
    xs = ((np.arange(W, dtype=f) + f(0.5)) / f(W)).astype(f)
    ys = ((np.arange(H, dtype=f) + f(0.5)) / f(H)).astype(f)
    z = np.zeros((H, W), dtype=f)
    acc = np.zeros((H, W, 3), dtype=f)
    accw = np.zeros((H, W), dtype=f)
    dx = f(-radius)
    while dx <= radius:
        pt = f(np.sqrt(f(radQ - f(dx * dx))))
        dy = f(-pt)
        while dy <= pt:
            blur = f(f(np.exp(f(-f(f(dx * dx) + f(dy * dy)) * invSigmaQx2))) * invSigmaQx2PI)
            cu = np.maximum(np.minimum((xs + f(dx / f(W))).astype(f), f(1.0)), f(0.0))
            cv = np.maximum(np.minimum((ys + f(dy / f(H))).astype(f), f(1.0)), f(0.0))
            tx = np.clip(np.floor((cu * f(W)).astype(f)).astype(np.int64), 0, W - 1)
            ty = np.clip(np.floor((cv * f(H)).astype(f)).astype(np.int64), 0, H - 1)
            walk = mean[ty][:, tx]
            s2 = (std[ty][:, tx] ** 2).astype(f)
            dots = (s2[..., 0] + s2[..., 1] + s2[..., 2]).astype(f)
            delta = (np.exp((-dots * invThresholdSqx2).astype(f)).astype(f) * invThresholdSqrt2PI * blur).astype(f)
            z = (z + delta).astype(f)
            acc = (acc + delta[..., None] * walk).astype(f)
            accw = (accw + delta * wmean[ty][:, tx]).astype(f)
            dy = f(dy + f(1.0))
        dx = f(dx + f(1.0))
    nz = z != 0
    out = np.zeros((H, W, 4), dtype=f)
    out[..., :3] = np.where(nz[..., None], acc / np.where(nz, z, f(1.0))[..., None], acc)
    out[..., 3] = np.where(nz, accw / np.where(nz, z, f(1.0)), accw)
    c = out[..., :3]
    nan = np.isnan(c).any(-1)
    inf = np.isinf(c).any(-1) & ~nan
    neg = (c < 0).any(-1) & ~nan & ~inf
    c[nan] = (1.0, 0.0, 0.5); c[inf] = (0.0, 1.0, 0.5); c[neg] = (0.5, 0.0, 1.0)
    if color_log_scale > 0:
        c[...] = (np.log(c + f(1.0)) / f(color_log_scale)).astype(f)
    return out
