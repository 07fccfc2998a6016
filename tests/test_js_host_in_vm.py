"""`js/cuda_renderer.js` — the class a maintainer of the reference drops next to its renderers — EXECUTED, next to the
reference's unmodified sources, by oracle/jsvm (no Node in the image; until this test that file had never run).

The scene objects are the reference's own (`tests/<name>/test.mjs` -> configureTest), the renderer is the product's
JavaScript, the N-API addon is replaced by a Python object with the addon's entry points (js/napi/jsrt_addon.cc:196-200):
scene parsing goes through the product's C ABI (host-only handle: the blob the JavaScript wrote must be one the wire
reader accepts), rendering through the oracle (this is a CPU test; the CUDA side of the same entry points is covered by
tests/test_napi_addon.py and the -m gpu tests).  Runs only where /root/reference exists."""
import json

import numpy as np
import pytest

from oracle import refjs
from oracle.jsvm import UNDEF, JSObject, JSTypedArray

pytestmark = pytest.mark.skipif(not refjs.available(), reason="needs the reference tree (build container only)")


class Blob(JSObject):
    """what `Buffer.from(text, "utf8")` returns"""
    __slots__ = ("data",)


class MockAddon:
    def __init__(self, r):
        self.r, self.vm = r, r.vm
        self.log = []
        self.scenes = {}
        vm = self.vm
        self.obj = JSObject(vm.ObjectProto)
        for name in ("createScene", "destroyScene", "sceneHeader", "resetAccum", "render", "synchronize", "resolveRGBA8"):
            self.obj.props[name] = vm.native(name, getattr(self, "_" + name))
        buf = JSObject(vm.ObjectProto)
        buf.props["from"] = vm.native("from", self._buffer_from)
        vm.root.vars["Buffer"] = buf
        vm.root.vars["__addon"] = self.obj

    def _buffer_from(self, this, args):
        assert args[1] == "utf8"
        b = Blob(self.vm.ObjectProto)
        b.data = args[0].encode("utf8")
        return b

    def _createScene(self, this, args):
        from jsraytracer_b200 import lib
        from oracle.oracle import OracleScene
        blob, fmt, dev = args[0], int(args[1]), args[2]
        assert fmt == 0 and isinstance(blob, Blob)
        host = lib.Scene(blob.data, lib.FORMAT_JSON, device=None)      # the product's wire reader + flattener accept it
        h = JSObject(self.vm.ObjectProto)
        self.scenes[id(h)] = dict(blob=blob.data, info=host.info, oracle=OracleScene(blob.data.decode("utf8")), acc=None, passes=0)
        self.log.append(("createScene", len(blob.data), dev))
        return h

    def _destroyScene(self, this, args):
        self.log.append(("destroyScene",))
        del self.scenes[id(args[0])]
        return UNDEF

    def _sceneHeader(self, this, args):
        from jsraytracer_b200 import lib
        i = lib.Scene(args[0].data, lib.FORMAT_JSON, device=None).info
        out = args[2]
        assert isinstance(out, JSTypedArray) and out.kind == "Int32Array" and len(out.items) == 5
        for k, v in enumerate((i["width"], i["height"], i["samples_per_pixel"], i["max_depth"], i["jitter"])):
            out.items[k] = v
        self.log.append(("sceneHeader",))
        return UNDEF

    def _resetAccum(self, this, args):
        s = self.scenes[id(args[0])]
        W, H = s["info"]["width"], s["info"]["height"]
        s["acc"], s["passes"] = np.zeros((H, W, 3), dtype=np.float32), 0
        self.log.append(("resetAccum",))
        return UNDEF

    def _render(self, this, args):
        s = self.scenes[id(args[0])]
        first, n, seed, xo, xd, flags = (int(a) for a in args[1:7])
        s["oracle"].render(n, first_pass=first, seed=seed, jitter=not (flags & 1), x_offset=xo, x_delt=xd, accum=s["acc"], threads=2)
        s["passes"] += n
        self.log.append(("render", first, n, seed, xo, xd, flags))
        return UNDEF

    def _synchronize(self, this, args):
        self.log.append(("synchronize",))
        return UNDEF

    def _resolveRGBA8(self, this, args):
        from oracle.oracle import resolve_rgba8
        s = self.scenes[id(args[0])]
        out = args[1]
        assert isinstance(out, JSTypedArray) and out.kind == "Uint8ClampedArray"
        img = resolve_rgba8(s["acc"], max(s["passes"], 1)).reshape(-1)
        assert len(out.items) == img.size
        for k, v in enumerate(img.tolist()):
            out.items[k] = v
        self.log.append(("resolveRGBA8",))
        return UNDEF


@pytest.fixture(scope="module")
def env():
    r = refjs.RefJS()                       # loads the reference's sources, then js/cuda_renderer.js
    r.load_test("BoxBall_DOF")
    return r, MockAddon(r)


def test_cuda_renderer_renders_the_reference_scene(env):
    from oracle.oracle import OracleScene, resolve_rgba8
    r, addon = env
    vm = r.vm
    addon.log.clear()
    vm.run("""
        var __cr = new CUDARenderer(__test.renderer.world, __test.renderer.camera, 3, __test.renderer.maxRecursionDepth, {addon: __addon, seed: 5});
        var __img = __cr.render(new PixelBuffer(12, 8));
    """)
    names = [c[0] for c in addon.log]
    assert names == ["createScene", "resetAccum", "render", "render", "render", "resolveRGBA8"]
    assert [c[1:] for c in addon.log if c[0] == "render"] == [(0, 1, 5, 0, 1, 0), (1, 1, 5, 0, 1, 0), (2, 1, 5, 0, 1, 0)]
    assert addon.log[0][2] == 0.0                                       # options.device default
    assert vm.eval_expr("__img instanceof PixelBuffer && __cr instanceof IncrementalMultisamplingRenderer") is True
    # the blob: the reference's serializer output for {renderer: <CUDARenderer>, width, height}, private fields left out
    scene = next(iter(addon.scenes.values()))
    doc = json.loads(scene["blob"])
    assert doc["_v"]["width"] == 12 and doc["_v"]["height"] == 8
    rend = doc["_v"]["renderer"]
    assert rend["_t"][0] == "CUDARenderer" and sorted(rend["_v"]) == ["camera", "maxRecursionDepth", "samplesPerPixel", "world"]
    assert scene["info"]["jitter"] == 1 and scene["info"]["samples_per_pixel"] == 3
    # the image in the reference's ImageData is the 3-pass image of that blob
    acc, _ = OracleScene(scene["blob"].decode()).render(3, seed=5, threads=2)
    got = np.array(vm.eval_expr("__img.imgdata.data").items, dtype=np.uint8).reshape(8, 12, 4)
    assert np.array_equal(got, resolve_rgba8(acc, 3))
    assert got[..., :3].any()


def test_scene_is_kept_until_invalidated_and_column_stripes_pass_through(env):
    r, addon = env
    vm = r.vm
    addon.log.clear()
    vm.run("__cr.samplesPerPixel = 1; __cr.render(new PixelBuffer(12, 8), 0, false, 1, 4);")
    assert [c[0] for c in addon.log] == ["resetAccum", "render", "resolveRGBA8"]          # no second createScene
    assert addon.log[1][1:] == (0, 1, 5, 1, 4, 0)                                        # src/worker.js:30-32 striping
    addon.log.clear()
    vm.run("__cr.render(new PixelBuffer(6, 4));")                                        # another size: re-serialised
    assert [c[0] for c in addon.log][:2] == ["destroyScene", "createScene"]
    addon.log.clear()
    vm.run("__cr.camera.setTransform(Mat4.translation([0, 1, 0])); __cr.invalidate(); __cr.render(new PixelBuffer(6, 4)); __cr.close(); __cr.close();")
    assert [c[0] for c in addon.log] == ["destroyScene", "createScene", "resetAccum", "render", "resolveRGBA8", "destroyScene"]
    assert not addon.scenes


def test_progress_callback_is_grouped_and_rate_limited(env):
    r, addon = env
    vm = r.vm
    addon.log.clear()
    vm.run("""
        var __calls = [];
        var __cr2 = new CUDARenderer(__test.renderer.world, __test.renderer.camera, 20, 2, {addon: __addon, passesPerGroup: 8});
        __cr2.render(new PixelBuffer(4, 3), 1e9, s => __calls.push(s));     // time limit never reached: no callback
        var __n0 = __calls.length;
        __cr2.render(new PixelBuffer(4, 3), -1, s => __calls.push([s.pass, s.completion]));   // always due
        __cr2.close();
    """)
    assert vm.eval_expr("__n0") == 0.0
    assert vm.to_py(vm.eval_expr("__calls")) == [[7, 0.4], [15, 0.8], [19, 1]]
    assert sum(1 for c in addon.log if c[0] == "synchronize") == 6          # one per group of 8 passes, not per pass


def test_from_wire_checks_the_image_size(env):
    r, addon = env
    vm = r.vm
    vm.run("""
        var __blob = Buffer.from(JSON.stringify(new Serializer({renderer: new SimpleRenderer(__test.renderer.world, __test.renderer.camera, 2), width: 5, height: 4}).plain()), "utf8");
        var __w = CUDARenderer.fromWire(__blob, 0, {addon: __addon});
        var __err = null;
        try { __w.renderer.render(new PixelBuffer(6, 4)); } catch (e) { __err = e.message; }
    """)
    assert vm.to_py(vm.eval_expr("[__w.width, __w.height, __w.renderer.samplesPerPixel, __w.renderer.maxRecursionDepth]")) == [5, 4, 1, 2]
    assert "5x4" in vm.eval_expr("__err") and "6x4" in vm.eval_expr("__err")
    addon.log.clear()
    vm.run("__w.renderer.render(new PixelBuffer(5, 4)); __w.renderer.close();")
    assert [c for c in addon.log if c[0] == "render"] == [("render", 0, 1, 1, 0, 1, 1)]      # a serialised SimpleRenderer: no jitter


def test_triangle_serialize_fix_is_what_keeps_vertex_normals():
    """src/geometry.js:355-357 writes `psdata: serializeStep(this.ps)`; the fix writes the real psdata"""
    r = refjs.RefJS()
    vm = r.vm
    vm.run("""
        var __t = new Triangle([Vec.of(0,0,0,1), Vec.of(1,0,0,1), Vec.of(0,1,0,1)], {normal: [Vec.of(0,0,1,0), Vec.of(0,1,0,0), Vec.of(1,0,0,0)]});
        var __before = JSON.stringify(new Serializer(__t).plain());
        installTriangleSerializeFix();
        var __after = JSON.stringify(new Serializer(__t).plain());
    """)
    before, after = json.loads(vm.eval_expr("__before")), json.loads(vm.eval_expr("__after"))
    assert before["_v"]["psdata"] == {"_r": before["_v"]["ps"]["_r"]}              # the reference: psdata IS ps
    assert sorted(after["_v"]["psdata"]["_v"]) == ["normal"]
