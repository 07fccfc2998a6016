"""The device's texture lookup checked ON THE CPU against the reference's `TextureMaterialColor.color`.

`texture_eval` (csrc/shade.cuh) restates src/materials.js:101-130 — coordinate clamp / wrap, bilinear and nearest modes,
f64 weights and sums, the result stored as an f32 RGBA vector.  The test cuts `tex_normalize_uv` + `texture_eval` out of
shade.cuh as they are, compiles them for the host behind shims for the CUDA spellings, takes the texture table and the texels
from the product's flattener (fed the reference-written document of the extra materials scene) and compares with what the
reference's own class returned in oracle/jsvm for 790 UVs per texture — random ones, far outside [0, 1], and the edges where
a texel boundary, a half-texel or a wrap falls exactly (tests/golden/probes_texture_refjs.npz, `python -m oracle.refjs_probes
texture`).  The bar is equality of all four f32 components."""
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
using namespace jsrt;
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
#define __device__
#define __noinline__
#define JSRT_DEV static inline
static inline double dmul(double a, double b) { return a * b; }      // __dmul_rn & co: never contracted (-ffp-contract=off)
static inline double dadd(double a, double b) { return a + b; }
static inline double dsub(double a, double b) { return a - b; }
"""

DRIVER = r"""
extern "C" int dev_texture(const char* blob, size_t len, int nearest, int n, const float* uv, float* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        int which = -1;
        for (size_t i = 0; i < hs.textures.size(); ++i) if (((hs.textures[i].flags & TF_NEAREST) != 0) == (nearest != 0)) which = (int)i;
        if (which < 0) return -1;
        for (int i = 0; i < n; ++i) {
            const float4 c = texture_eval(hs.textures.data(), hs.texels.data(), which, uv[2 * i], uv[2 * i + 1], 1.f, 1.f, 1.f, 1.f);
            out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = c.w;
        }
        return hs.textures[which].flags;
    } catch (const std::exception&) { return -2; }
}
"""


@pytest.fixture(scope="module")
def dev(tmp_path_factory):
    text = open(os.path.join(CSRC, "shade.cuh")).read()
    a = text.index("JSRT_DEV double tex_normalize_uv(")
    b = text.index("// LEAN (here and below): builds of shade_kernel")
    block = text[a:b]
    assert "float4 texture_eval(" in block
    d = tmp_path_factory.mktemp("dev_tex")
    cpp = d / "dev_tex.cpp"
    cpp.write_text(SHIM + block + DRIVER)
    so = d / "dev_tex.so"
    srcs = [os.path.join(CSRC, f) for f in ("wire.cpp", "scene_flatten.cpp", "sdf_compile.cpp", "bvh_build.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I" + CSRC,
                           "-o", str(so), str(cpp)] + srcs)
    L = ctypes.CDLL(str(so))
    L.dev_texture.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


@pytest.mark.parametrize("which,nearest,flags", [("bilinear_wrap", 0, 0), ("nearest_clamp", 1, 1 | 2 | 4)])
def test_device_texture_equals_reference(dev, which, nearest, flags):
    z = np.load(os.path.join(HERE, "golden", "probes_texture_refjs.npz"))
    blob = zlib.decompress(np.load(os.path.join(HERE, "golden", "refjs_extra_materials_whitted.npz"))["json"].tobytes())
    uv = np.ascontiguousarray(z["uv"], dtype=np.float32)
    want = z["tex_" + which].astype(np.float32)
    out = np.zeros((len(uv), 4), dtype=np.float32)
    assert dev.dev_texture(blob, len(blob), nearest, len(uv), uv.ctypes.data, out.ctypes.data) == flags
    bad = np.nonzero(~(out == want).all(-1))[0]
    assert bad.size == 0, "%d of %d lookups differ, e.g. uv=%s reference %s device %s" % (
        bad.size, len(uv), uv[bad[0]].tolist(), want[bad[0]].tolist(), out[bad[0]].tolist())
    assert len({tuple(r) for r in want.tolist()}) > (25 if nearest else 300)       # the probes do see the texture
