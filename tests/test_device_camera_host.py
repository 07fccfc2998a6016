"""The device's camera rays checked ON THE CPU against the reference's own render loop.

`camera_ray` (csrc/trace.cuh) restates `IncrementalMultisamplingRenderer.render`'s pixel / jitter arithmetic and
`PerspectiveCamera` / `DepthOfFieldPerspectiveCamera.getRayForPixel` (src/renderers.js:87-98, src/cameras.js:29-34,46-52) in
un-contracted f64.  This test cuts `GenParams` + `camera_ray` out of trace.cuh as they are, compiles them for the host behind
shims for the CUDA spellings, feeds them the camera the product's flattener extracts from the reference-written scene
document, and compares every ray with the one the reference's own render loop handed to `world.color` when `Math.random()`
returned the device generator's numbers for that pixel sample (tests/golden/probes_camera_refjs.npz, made by
`python -m oracle.refjs_probes camera`).  The bar is equality of origin and direction, component by component."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "jsraytracer_b200", "csrc")

SHIM = r"""
#define _GNU_SOURCE 1
#include <cmath>
#include <cstdint>
#include <cstring>
#include "host_scene.h"
#include "rng.h"
using namespace jsrt;
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
#define JSRT_DEV static inline
static inline float3 f3(float x, float y, float z) { return float3{x, y, z}; }
static inline double dmul(double a, double b) { return a * b; }      // __dmul_rn & co: never contracted (-ffp-contract=off)
static inline double dadd(double a, double b) { return a + b; }
static inline double dsub(double a, double b) { return a - b; }
static inline double ddot4(double ax, double ay, double az, double aw, double bx, double by, double bz, double bw) {
    return dadd(dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz)), dmul(aw, bw));
}
"""

DRIVER = r"""
extern "C" int dev_camera_rays(const char* blob, size_t len, int W, int H, int passes, unsigned long long seed, float* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        GenParams g{};
        g.cam = hs.camera; g.width = W; g.height = H;
        for (int pass = 0; pass < passes; ++pass)
            for (int py = 0; py < H; ++py)
                for (int px = 0; px < W; ++px) {
                    float3 o, d;
                    camera_ray(g, px, py, rng_sample_key(seed, (uint32_t)(py * W + px), (uint32_t)pass), hs.jitter, true, o, d);
                    float* r = out + (((size_t)pass * H + py) * W + px) * 6;
                    r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
                }
        return hs.camera.dof;
    } catch (const std::exception&) { return -2; }
}
"""


@pytest.fixture(scope="module")
def dev(tmp_path_factory):
    text = open(os.path.join(CSRC, "trace.cuh")).read()
    a = text.index("struct GenParams {")
    b = text.index("// sample s of the batch -> (pixel, pass) and the three queue words of its camera ray")
    block = text[a:b]
    assert "JSRT_DEV void camera_ray(" in block
    d = tmp_path_factory.mktemp("dev_cam")
    cpp = d / "dev_cam.cpp"
    cpp.write_text(SHIM + block + DRIVER)
    so = d / "dev_cam.so"
    srcs = [os.path.join(CSRC, f) for f in ("wire.cpp", "scene_flatten.cpp", "sdf_compile.cpp", "bvh_build.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I" + CSRC,
                           "-o", str(so), str(cpp)] + srcs)
    L = ctypes.CDLL(str(so))
    L.dev_camera_rays.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p]
    return L


Z = np.load(os.path.join(HERE, "golden", "probes_camera_refjs.npz"))
SCENES = sorted(k[4:-5] for k in Z.files if k.endswith("_rays"))


@pytest.mark.parametrize("name", SCENES)
def test_device_camera_rays_equal_the_reference_render_loop(dev, name):
    W, H, P, seed = (int(v) for v in Z["cam_%s_meta" % name])
    want = Z["cam_%s_rays" % name]                      # (P, H, W, 8): origin xyzw, direction xyzw, f32 values
    blob = zlib.decompress(np.load(os.path.join(HERE, "golden", "refjs_%s.npz" % name))["json"].tobytes())
    out = np.zeros((P, H, W, 6), dtype=np.float32)
    dof = dev.dev_camera_rays(blob, len(blob), W, H, P, seed, out.ctypes.data)
    assert dof >= 0
    assert dof == (1 if name in ("BoxBall_DOF",) else 0)
    ref = np.concatenate([want[..., 0:3], want[..., 4:7]], axis=-1).astype(np.float32)
    assert np.all(want[..., 3] == 1) and np.all(want[..., 7] == 0)           # points and directions
    bad = np.argwhere(~(out == ref).all(-1))
    assert bad.size == 0, "%d of %d rays differ, first at (pass, y, x) = %s: reference %s device %s" % (
        len(bad), P * H * W, bad[0].tolist(), ref[tuple(bad[0])].tolist(), out[tuple(bad[0])].tolist())
    assert len({tuple(r) for r in out.reshape(-1, 6)[:, 3:].tolist()}) == P * H * W      # every sample has its own direction (jitter)
