"""js/napi/jsrt_addon.cc (the N-API shim a maintainer adds under Node) compiled against a test double
of <node_api.h> and driven through a minimal N-API runtime (tests/napi_mock/): the image has no Node, so this is how
the addon's C code is compiled and executed at all.  The call sequence is js/cuda_renderer.js's:
createScene(blob, format, device) -> resetAccum -> render(scene, first, n, seed, x_offset, x_delt, flags) ->
resolveRGBA8(scene, Uint8ClampedArray) -> destroyScene."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "napi_mock")
SO = os.path.join(MOCK, "libnapi_mock.so")
NAPI_UINT8_ARRAY, NAPI_UINT8_CLAMPED, NAPI_INT32_ARRAY, NAPI_FLOAT32_ARRAY = 1, 2, 5, 7


def build_mock():
    srcs = [os.path.join(MOCK, "mock_runtime.cc"), os.path.join(ROOT, "js", "napi", "jsrt_addon.cc")]
    deps = srcs + [os.path.join(MOCK, "node_api.h"), os.path.join(ROOT, "include", "jsrt.h"), os.path.join(ROOT, "jsraytracer_b200", "libjsrt.so")]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I" + MOCK, "-o", SO] + srcs +
                          ["-L" + os.path.join(ROOT, "jsraytracer_b200"), "-ljsrt", "-Wl,-rpath," + os.path.join(ROOT, "jsraytracer_b200")])
    return SO


class Value:
    """An opaque napi_value handed back by the addon (the scene external)."""

    def __init__(self, ptr):
        self.ptr = ptr


class Addon:
    """`require('./napi/build/Release/jsrt_addon.node')` under the mock runtime."""

    def __init__(self):
        from jsraytracer_b200 import lib
        lib.load()                                    # libjsrt.so must exist (no fallback)
        self.m = C.CDLL(build_mock())
        m = self.m
        m.mock_env_create.restype = C.c_void_p
        m.mock_export_name.restype = C.c_char_p
        m.mock_export_name.argtypes = [C.c_void_p, C.c_int]
        m.mock_export_count.argtypes = [C.c_void_p]
        m.mock_number.restype = C.c_void_p
        m.mock_number.argtypes = [C.c_void_p, C.c_double]
        m.mock_buffer.restype = C.c_void_p
        m.mock_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        m.mock_typedarray.restype = C.c_void_p
        m.mock_typedarray.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        m.mock_call.restype = C.c_void_p
        m.mock_call.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        m.mock_take_exception.restype = C.c_char_p
        m.mock_take_exception.argtypes = [C.c_void_p]
        m.mock_is_number.argtypes = [C.c_void_p]
        m.mock_is_external.argtypes = [C.c_void_p]
        m.mock_number_value.restype = C.c_double
        m.mock_number_value.argtypes = [C.c_void_p]
        m.mock_env_destroy.argtypes = [C.c_void_p]
        m.mock_run_finalizer.argtypes = [C.c_void_p, C.c_void_p]
        self.env = m.mock_env_create()

    def exports(self):
        return [self.m.mock_export_name(self.env, i).decode() for i in range(self.m.mock_export_count(self.env))]

    def wrap(self, a):
        if isinstance(a, (int, float)):
            return self.m.mock_number(self.env, float(a))
        if isinstance(a, (bytes, bytearray)):
            self._keep = C.create_string_buffer(bytes(a), len(a))
            return self.m.mock_buffer(self.env, C.cast(self._keep, C.c_void_p), len(a))
        if isinstance(a, np.ndarray):
            ty = {np.dtype(np.uint8): NAPI_UINT8_CLAMPED, np.dtype(np.int32): NAPI_INT32_ARRAY, np.dtype(np.float32): NAPI_FLOAT32_ARRAY,
                  np.dtype(np.int8): 0}[a.dtype]
            return self.m.mock_typedarray(self.env, ty, a.ctypes.data, a.size)
        assert isinstance(a, Value), type(a)
        return a.ptr

    def call(self, name, *args):
        argv = (C.c_void_p * max(1, len(args)))(*[self.wrap(a) for a in args])
        r = self.m.mock_call(self.env, name.encode(), len(args), argv)
        exc = self.m.mock_take_exception(self.env)
        if exc is not None:
            raise RuntimeError(exc.decode())           # what Node would throw into JS
        if r and self.m.mock_is_number(r):
            return int(self.m.mock_number_value(r))
        return Value(r) if r else None


@pytest.fixture(scope="module")
def addon():
    return Addon()


def test_addon_registers_the_functions_cuda_renderer_js_calls(addon):
    src = open(os.path.join(ROOT, "js", "cuda_renderer.js")).read()
    ex = addon.exports()
    assert set(ex) == {"createScene", "destroyScene", "render", "resetAccum", "synchronize", "resolveRGBA8", "primaryHits", "deviceCount",
                       "sceneHeader", "readAccum", "readAov", "denoise"}
    import re
    used = set(re.findall(r"(?:addon|_addon\(\))\.(\w+)\(", src))
    assert used and used <= set(ex), used - set(ex)


def test_addon_argument_errors_become_js_exceptions(addon):
    from jsraytracer_b200 import lib
    assert addon.call("deviceCount") == lib.device_count()
    with pytest.raises(RuntimeError, match="TypeError: scene handle expected"):
        addon.call("render", 5, 0, 1, 1, 0, 1, 0)             # a number where the external is expected
    with pytest.raises(RuntimeError, match="Error: N-API call failed"):
        addon.call("createScene", 1, 0, 0)                    # not a Buffer
    if lib.device_count() == 0:
        # no GPU here: jsrt_scene_create fails, and its message must surface as a thrown Error (no CPU fallback)
        with pytest.raises(RuntimeError, match="Error: "):
            addon.call("createScene", b'{"renderer": null}', 0, 0)


def test_addon_scene_header_reads_a_wire_blob_without_a_gpu(addon):
    """CUDARenderer.fromWire (the *_json scenes): width / height / spp / depth / jitter straight from the blob."""
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.serializer import Serializer
    ser = Serializer(scenes.configure("BoxBall", width=40, height=30, spp=3, depth=5))
    for blob, fmt in ((ser.to_json().encode(), 0), (ser.to_msgpack(), 1)):
        head = np.zeros(5, np.int32)
        addon.call("sceneHeader", blob, fmt, head)
        assert head.tolist() == [40, 30, 3, 5, 1]
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("sceneHeader", ser.to_msgpack(), 1, np.zeros(4, np.int32))
    with pytest.raises(RuntimeError, match="Error: "):
        addon.call("sceneHeader", b"{}", 0, np.zeros(5, np.int32))


@pytest.mark.gpu
def test_addon_renders_the_same_bytes_as_the_ctypes_binding(addon):
    from jsraytracer_b200 import lib, scenes
    from jsraytracer_b200.serializer import Serializer
    W = H = 96
    ser = Serializer(scenes.configure("BoxBall", width=W, height=H))
    blob = ser.to_msgpack()
    # reference-facing sequence of js/cuda_renderer.js
    scene = addon.call("createScene", blob, 1, 0)
    assert addon.m.mock_is_external(scene.ptr)
    addon.call("resetAccum", scene)
    for done in range(0, 4, 2):
        addon.call("render", scene, done, 2, 1, 0, 1, 0)
    addon.call("synchronize", scene)
    img = np.zeros(W * H * 4, np.uint8)
    addon.call("resolveRGBA8", scene, img)
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("resolveRGBA8", scene, np.zeros(16, np.uint8))
    acc = np.zeros(W * H * 4, np.float32)
    assert addon.call("readAccum", scene, acc) == 4                 # passes accumulated
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("readAccum", scene, np.zeros(8, np.float32))
    ids, t = np.zeros(W * H, np.int32), np.zeros(W * H, np.float32)
    addon.call("primaryHits", scene, ids, t)
    # wrong element type / short arrays must be refused, not written through
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("primaryHits", scene, np.zeros(W * H, np.float32), t)
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("primaryHits", scene, np.zeros(W * H - 1, np.int32), t)
    # the GL path's variance buffer + variance-guided display filter through the same shim (flag 4 = JSRT_FLAG_AOV)
    addon.call("resetAccum", scene)
    addon.call("render", scene, 0, 4, 1, 0, 1, 4)
    nd, var = np.zeros(W * H * 4, np.float32), np.zeros(W * H * 4, np.float32)
    addon.call("readAov", scene, nd, var)
    assert var.reshape(H, W, 4)[..., 3].max() == 4 and np.isfinite(nd).all()           # first hits per pixel, normal / distance sums
    den8, denf = np.zeros(W * H * 4, np.uint8), np.zeros(W * H * 4, np.float32)
    addon.call("denoise", scene, 1.0, 2.0, 5.0, 0.0, den8)
    addon.call("denoise", scene, 1.0, 2.0, 5.0, 0.0, denf)
    assert den8.reshape(H, W, 4)[..., 3].min() == 255 and np.isfinite(denf).all()
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("readAov", scene, nd, np.zeros(8, np.float32))
    with pytest.raises(RuntimeError, match="RangeError"):
        addon.call("denoise", scene, 1.0, 2.0, 5.0, 0.0, np.zeros(W * H, np.int32))
    with pytest.raises(RuntimeError, match="TypeError"):
        addon.call("denoise", scene, nd, 2.0, 5.0, 0.0, den8)
    addon.call("destroyScene", scene)
    addon.call("destroyScene", scene)                                # a second destroy is a no-op, not a double free
    with pytest.raises(RuntimeError, match="after destroyScene"):
        addon.call("synchronize", scene)
    # a scene that is never destroyed explicitly is released by the external's finalizer
    leaked = addon.call("createScene", blob, 1, 0)
    assert addon.m.mock_run_finalizer(addon.env, leaked.ptr) == 1
    # the same through ctypes
    sc = lib.Scene(blob, lib.FORMAT_MSGPACK, device=0)
    sc.render(0, 4, seed=1)
    ref = np.zeros(W * H * 4, np.uint8)
    sc.resolve_rgba8(ref)
    racc, rpasses = sc.read_accum()
    rid, rt = sc.primary_hits()
    assert np.array_equal(img, ref)
    assert rpasses == 4 and np.allclose(acc.reshape(racc.shape), racc, rtol=1e-5, atol=1e-6)      # (atomic float sums: order may differ)
    assert np.array_equal(ids, rid.ravel()) and np.array_equal(t, rt.ravel())
