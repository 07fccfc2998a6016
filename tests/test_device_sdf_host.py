"""The device's SDF distance function — bytecode compiler (csrc/sdf_compile.cpp) + interpreter (`sdf_eval`,
csrc/device_math.cuh) — checked ON THE CPU against the reference's own `root_sdf.distance`.

`sdf_eval` is written in the reference's arithmetic (f32 vectors, f64 scalars, un-contracted operations) and is plain
C++ apart from a dozen CUDA spellings.  This test cuts the block from `xf64_apply` to the end of `sdf_eval` out of
device_math.cuh as it is, compiles it for the host behind shims for those spellings (`__dmul_rn` -> `*` under
-ffp-contract=off, `__ldg(p)` -> `*p`, `float3`, `__hiloint2double`, ...), links it with the product's real wire reader,
flattener and SDF compiler, and evaluates the compiled program of every SDF fixture scene at the probe points of
tests/golden/probes_refjs.npz, where the reference's own sources (run in oracle/jsvm) returned `root_sdf.distance(p)`.
The bar is equality of every double."""
import ctypes
import json
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
#include <limits>
#include <vector>
#include "host_scene.h"
using namespace jsrt;
using std::isfinite; using std::isinf;
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct DeviceScene { const SdfInstr* sdf_code; const Xform64* xforms64; };       // the two members the SDF material program reads
#define __device__
#define __noinline__
#define JSRT_DEV static inline
#define CUDART_NAN (std::numeric_limits<double>::quiet_NaN())
#define CUDART_INF (std::numeric_limits<double>::infinity())
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float3 f3(float x, float y, float z) { return float3{x, y, z}; }
// IEEE round-to-nearest single operations that the compiler must not contract: this file is built with -ffp-contract=off
static inline double dmul(double a, double b) { return a * b; }
static inline double dadd(double a, double b) { return a + b; }
static inline double dsub(double a, double b) { return a - b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline double ddot3(double ax, double ay, double az, double bx, double by, double bz) { return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz)); }
static inline double ddot4(double ax, double ay, double az, double aw, double bx, double by, double bz, double bw) {
    return dadd(dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz)), dmul(aw, bw));
}
static inline double jsd_min(double a, double b) { return (a != a || b != b) ? CUDART_NAN : fmin(a, b); }
static inline double jsd_max(double a, double b) { return (a != a || b != b) ? CUDART_NAN : fmax(a, b); }
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
"""

DRIVER = r"""
@MATERIAL_BLOCK@
// root_sdf.getMaterialData(p): the material program of the scene's SDF (uniform basecolor: the program's `base`)
extern "C" int dev_sdf_materials(const char* blob, size_t len, int n, const double* pts, double* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        if (hs.sdfs.empty()) return -1;
        const SdfProgram& pr = hs.sdfs[0];
        const DeviceScene sc{hs.sdf_code.data(), hs.xforms64.data()};
        for (int i = 0; i < n; ++i) {
            float3 base = f3(pr.base[0], pr.base[1], pr.base[2]); float2 uv = make_float2(0.f, 0.f); bool has_uv = false;
            if (pr.mat_first >= 0) sdf_material(sc, pr.mat_first, f3((float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]), base, uv, has_uv);
            double* o = out + 5 * i;
            o[0] = base.x; o[1] = base.y; o[2] = base.z; o[3] = has_uv ? uv.x : CUDART_NAN; o[4] = has_uv ? uv.y : CUDART_NAN;
        }
        return pr.mat_first >= 0 ? 2 : 1;
    } catch (const std::exception&) { return -2; }
}
// the forward-difference normal of SDFGeometry.materialData (src/sdf.js:41-47): the statements of shade.cuh's G_SDF case
struct ScShim { const SdfInstr* sdf_code; const Xform64* xforms64; };
static float3 dev_sdf_normal(const ScShim& sc, const SdfProgram& pr, float3 lp) {
    float3 n;
@NORMAL_BLOCK@
    return n;
}
extern "C" int dev_sdf_normals(const char* blob, size_t len, int n, const double* pts, double* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        if (hs.sdfs.empty()) return -1;
        const ScShim sc{hs.sdf_code.data(), hs.xforms64.data()};
        for (int i = 0; i < n; ++i) {
            const float3 v = dev_sdf_normal(sc, hs.sdfs[0], f3((float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]));
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        }
        return 1;
    } catch (const std::exception&) { return -2; }
}
extern "C" int dev_sdf_probe(const char* blob, size_t len, int n, const double* pts, double* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        if (hs.sdfs.empty()) return -1;
        const SdfProgram& pr = hs.sdfs[0];
        const SdfInstr* prog = hs.sdf_code.data() + pr.first_instr;
        for (int i = 0; i < n; ++i) {
            const float3 p = f3((float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]);
            out[2 * i] = sdf_eval<1>(prog, hs.xforms64.data(), p);        // the build with the fused Menger step inlined
            out[2 * i + 1] = sdf_eval<-1>(prog, hs.xforms64.data(), p);   // ... and the one that calls it
        }
        return pr.instr_count;
    } catch (const std::exception&) { return -2; }
}

// SDFGeometry.intersect: the sphere-tracing march (sdf_intersect) over the same compiled program
extern "C" int dev_sdf_hits(const char* blob, size_t len, int n, const double* rays, double* out) {
    try {
        WireDoc doc((const uint8_t*)blob, len, 0);
        HostScene hs;
        flattenScene(doc, hs);
        if (hs.sdfs.empty()) return -1;
        const SdfProgram& pr = hs.sdfs[0];
        for (int i = 0; i < n; ++i) {
            const double* r = rays + 8 * i;
            unsigned long long evals = 0;
            out[2 * i] = sdf_intersect(pr, hs.sdf_code.data(), hs.xforms64.data(), f3((float)r[0], (float)r[1], (float)r[2]),
                                       f3((float)r[3], (float)r[4], (float)r[5]), r[6], r[7], &evals);
            out[2 * i + 1] = (double)evals;
        }
        return pr.max_samples;
    } catch (const std::exception&) { return -2; }
}
"""


@pytest.fixture(scope="module")
def dev(tmp_path_factory):
    text = open(os.path.join(CSRC, "device_math.cuh")).read()
    a = text.index("JSRT_DEV float3 xf64_apply(")
    b = text.rindex("}  // namespace jsrt")
    block = text[a:b]
    assert "double sdf_eval(" in block and "sdf_rtu_cross" in block and "js_fmod_pow2" in block and "double sdf_intersect(" in block
    shade = open(os.path.join(CSRC, "shade.cuh")).read()
    na = shade.index("            const SdfInstr* prog = sc.sdf_code + pr.first_instr;")
    nb = shade.index("\n", shade.index("n = (nn > 0.00001) ?", na)) + 1
    normal_block = shade[na:nb]
    assert normal_block.count("sdf_eval(") == 4
    ma = shade.index("struct SdfMat {")
    mb = shade.index("// geometry.materialData in the primitive's local space")
    material_block = shade[ma:mb]
    assert "void sdf_material(" in material_block
    d = tmp_path_factory.mktemp("dev_sdf")
    cpp = d / "dev_sdf.cpp"
    cpp.write_text(SHIM + block + DRIVER.replace("@NORMAL_BLOCK@", normal_block).replace("@MATERIAL_BLOCK@", material_block))
    so = d / "dev_sdf.so"
    srcs = [os.path.join(CSRC, f) for f in ("wire.cpp", "scene_flatten.cpp", "sdf_compile.cpp", "bvh_build.cpp")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I" + CSRC,
                           "-o", str(so), str(cpp)] + srcs)
    L = ctypes.CDLL(str(so))
    L.dev_sdf_probe.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.dev_sdf_materials.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.dev_sdf_normals.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.dev_sdf_hits.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


Z = np.load(os.path.join(HERE, "golden", "probes_refjs.npz"))
SCENES = sorted(k[4:-2] for k in Z.files if k.startswith("sdf_") and k.endswith("_p"))


@pytest.mark.parametrize("name", SCENES)
def test_device_sdf_distance_equals_reference(dev, name):
    z = np.load(os.path.join(HERE, "golden", "refjs_%s.npz" % name))
    blob = zlib.decompress(z["json"].tobytes())
    pts = np.ascontiguousarray(Z["sdf_%s_p" % name], dtype=np.float64)
    want = Z["sdf_%s_out" % name][:, 0]
    out = np.zeros((len(pts), 2))
    n_instr = dev.dev_sdf_probe(blob, len(blob), len(pts), pts.ctypes.data, out.ctypes.data)
    assert n_instr > 0, "flatten / compile failed (%d)" % n_instr
    for col, build in ((0, "RTU inlined"), (1, "RTU called")):
        got = out[:, col]
        bad = np.nonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))[0]
        assert bad.size == 0, "%s, %s: %d of %d points differ, e.g. p=%s reference %r device %r" % (
            name, build, bad.size, len(pts), pts[bad[0]].tolist(), float(want[bad[0]]), float(got[bad[0]]))


ZH = np.load(os.path.join(HERE, "golden", "probes_sdfhit_refjs.npz"))


@pytest.mark.parametrize("name", SCENES)
def test_device_sdf_march_equals_reference_intersect(dev, name):
    """`SDFGeometry.intersect` (src/sdf.js:12-40): bounding-box clip, march, epsilon stop, trace-distance and window exits —
    the hit parameter t of the device's `sdf_intersect` equals the reference's on every probe ray, misses are misses"""
    z = np.load(os.path.join(HERE, "golden", "refjs_%s.npz" % name))
    blob = zlib.decompress(z["json"].tobytes())
    rays = np.ascontiguousarray(ZH["hit_%s_rays" % name], dtype=np.float64)
    want = ZH["hit_%s_t" % name]
    out = np.zeros((len(rays), 2))
    assert dev.dev_sdf_hits(blob, len(blob), len(rays), rays.ctypes.data, out.ctypes.data) > 0
    got = out[:, 0]
    miss_w, miss_g = ~np.isfinite(want), ~np.isfinite(got)
    assert np.array_equal(miss_w, miss_g), "hit / miss differs on rays %s" % np.nonzero(miss_w != miss_g)[0][:5].tolist()
    bad = np.nonzero(~miss_w & (got != want))[0]
    assert bad.size == 0, "%d of %d hits differ, e.g. ray %s reference t %r device t %r" % (
        bad.size, int((~miss_w).sum()), rays[bad[0]].tolist(), float(want[bad[0]]), float(got[bad[0]]))
    if int((~miss_w).sum()) > 5:
        assert out[:, 1].max() > 5      # the march really marched


@pytest.mark.parametrize("name", SCENES)
def test_device_sdf_normal_equals_reference_material_data(dev, name):
    """the normal of `SDFGeometry.materialData` (src/sdf.js:41-47: three forward differences over `normal_step_size`, then
    `normalized()`): the statements of shade.cuh's G_SDF case, cut out as they are, give the reference's f32 normal"""
    z = np.load(os.path.join(HERE, "golden", "refjs_%s.npz" % name))
    blob = zlib.decompress(z["json"].tobytes())
    pts = np.ascontiguousarray(Z["sdf_%s_p" % name], dtype=np.float64)
    want = Z["sdf_%s_out" % name][:, 1:4]
    out = np.zeros((len(pts), 3))
    assert dev.dev_sdf_normals(blob, len(blob), len(pts), pts.ctypes.data, out.ctypes.data) == 1
    bad = np.nonzero(~((out == want) | (np.isnan(out) & np.isnan(want))).all(-1))[0]
    assert bad.size == 0, "%d of %d normals differ, e.g. p=%s reference %s device %s" % (
        bad.size, len(pts), pts[bad[0]].tolist(), want[bad[0]].tolist(), out[bad[0]].tolist())


ZM = np.load(os.path.join(HERE, "golden", "probes_sdfmat_refjs.npz"))


@pytest.mark.parametrize("name", SCENES)
def test_device_sdf_material_equals_reference_get_material_data(dev, name):
    """`root_sdf.getMaterialData(p)`: the leaf whose distance wins a Union / Intersection / Difference, the blends of the smooth
    variants, SphereSDF's spherical UVs — sdf_compile.cpp's material program run by shade.cuh's `sdf_material` gives the
    reference's basecolor and UV (f32) at every probe point"""
    z = np.load(os.path.join(HERE, "golden", "refjs_%s.npz" % name))
    blob = zlib.decompress(z["json"].tobytes())
    pts = np.ascontiguousarray(Z["sdf_%s_p" % name], dtype=np.float64)
    want = ZM["mat_%s" % name]
    out = np.zeros((len(pts), 5))
    kind = dev.dev_sdf_materials(blob, len(blob), len(pts), pts.ctypes.data, out.ctypes.data)
    assert kind in (1, 2)
    want = want.copy()
    want[np.isnan(want[:, 0]), 0:3] = 1.0                 # no basecolor in the data: PhongMaterial falls back to (1, 1, 1), src/materials.js:223
    bad = np.nonzero(~((out == want) | (np.isnan(out) & np.isnan(want))).all(-1))[0]
    assert bad.size == 0, "%d of %d points differ, e.g. p=%s reference %s device %s" % (
        bad.size, len(pts), pts[bad[0]].tolist(), want[bad[0]].tolist(), out[bad[0]].tolist())
    if name == "SDF_Combinations":
        assert kind == 2 and len({tuple(r) for r in want[:, :3].tolist()}) >= 2        # several leaf colours are really selected
