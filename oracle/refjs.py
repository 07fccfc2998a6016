"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (its `src/*.js` and `tests/<name>/test.mjs`, read where they
lie under /root/reference) inside `oracle.jsvm`, the way `src/worker.js:3-38` runs it in a browser worker:

    importScripts(math.js, world.js, ...)   ->  RefJS.__init__   (same files, same order, one global scope)
    import(test).configureTest(callback)    ->  RefJS.load_test
    test.renderer.render(buffer, ..., workerIndex, workerCount)   ->  RefJS.render

The only things supplied from outside are what a browser supplies: `ImageData`, `fetch`, and `Math.random`.
`Math.random` is a splitmix64 stream re-seeded for every pixel sample (seed, pixel, pass) — the oracle's "tape
mode" (`orc_render` flags bit1, oracle_math.h `rng_tape_*`) draws from the identical stream, so the restatement and
the reference agree sample by sample when (and only when) they consume random numbers in the same order.

Nothing here runs on the GPU box: tests use the committed fixtures (tests/golden/refjs_*.npz, written by
oracle/refjs_golden.py) when /root/reference is absent.
"""
from __future__ import annotations

import os

import numpy as np

from .jsvm import VM, UNDEF, JSObject, JSTypedArray

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("JSRT_REFERENCE_ROOT", "/root/reference")

# src/worker.js:3-14, plus the serializer (tests/test_to_json.js:8-19 loads it the same way)
SOURCES = ["math.js", "world.js", "pixelbuffer.js", "geometry.js", "materials.js", "cameras.js", "renderers.js",
           "lights.js", "objloader.js", "sdf.js", "aggregates.js", "serializer.js"]

MASK = (1 << 64) - 1


def tape_seed(seed, pixel, pass_):
    return ((seed << 48) ^ (pass_ << 32) ^ pixel) & MASK


class Tape:
    """splitmix64 -> [0, 1) with 53 bits; mirrors rng_tape_next (oracle_math.h)"""

    def __init__(self):
        self.state = 0
        self.draws = 0

    def seed(self, s):
        self.state = s & MASK

    def __call__(self):
        self.state = (self.state + 0x9E3779B97F4A7C15) & MASK
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        z ^= z >> 31
        self.draws += 1
        return float(z >> 11) * (1.0 / 9007199254740992.0)


# the browser objects the reference touches on this path, and the hook that lets the host see every finished sample
_PRELUDE = r"""
class ImageData {
    constructor(width, height) { this.width = width; this.height = height; this.data = new Uint8ClampedArray(4 * width * height); }
}
class OffscreenCanvas {          // src/materials.js:91-96 draws a bitmap and reads the pixels back
    constructor(width, height) { this.width = width; this.height = height; }
    getContext(kind) {
        const canvas = this;
        return {
            drawImage(bitmap, x, y) { canvas._bitmap = bitmap; },
            getImageData(x, y, w, h) { const img = new ImageData(w, h); img.data.set(__decodeImage(canvas._bitmap)); return img; }
        };
    }
}
class __TapBuffer extends PixelBuffer {
    setColor(x, y, color) {
        const r = super.setColor(x, y, color);
        __sampleDone(x, y, color);
        return r;
    }
}
"""


def test_dir(root, name):
    """tests/<name> of the reference — or, for `extra_<x>`, a scene this repository wrote out of the reference's classes to
    reach what none of the reference's own scenes uses (tests/golden/extra_scenes/<x>/test.mjs)"""
    if name.startswith("extra_"):
        return os.path.join(_REPO, "tests", "golden", "extra_scenes", name[6:])
    return os.path.join(root, "tests", name)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src")) and os.path.isdir(os.path.join(REF_ROOT, "tests"))


class RefJS:
    def __init__(self, root=REF_ROOT):
        self.root = root
        self.vm = vm = VM()
        self.tape = Tape()
        vm.random = self.tape
        self.cwd = os.path.join(root, "tests", "_")
        G = vm.root.vars
        G["fetch"] = vm.native("fetch", self._fetch)
        G["__sampleDone"] = vm.native("__sampleDone", self._sample_done)
        G["createImageBitmap"] = vm.native("createImageBitmap", self._create_image_bitmap)
        G["__decodeImage"] = vm.native("__decodeImage", self._decode_image)
        for f in SOURCES:
            with open(os.path.join(root, "src", f)) as fh:
                vm.run(fh.read(), f)
        vm.run(_PRELUDE, "<prelude>")
        # The product's own host-side JavaScript, loaded the way its header says it is (as a classic script next to the
        # reference's sources).  It carries the fix for the reference serializer's lossy Triangle.serialize
        # (src/geometry.js:355-357 writes `psdata: serializeStep(this.ps)`: vertex normals never reach the wire).
        G["require"] = vm.native("require", self._require)
        with open(os.path.join(_REPO, "js", "cuda_renderer.js")) as fh:
            vm.run(fh.read(), "js/cuda_renderer.js")
        self.test = None
        self._on_sample = None

    # -- browser shims -------------------------------------------------------------------------
    def _fetch(self, this, args):
        vm = self.vm
        url = vm.tostr(args[0])
        # a worker resolves relative URLs against its own script, src/worker.js ("../assets/x.obj" -> <root>/assets/x.obj)
        path = os.path.normpath(os.path.join(self.root, "src", url))
        resp = JSObject(vm.ObjectProto)
        if not os.path.isfile(path):
            resp.props["ok"] = False
            return vm.promise_resolve(resp)
        resp.props["ok"] = True

        def text(this_, a):
            with open(path, encoding="utf8", errors="replace") as fh:
                return vm.promise_resolve(fh.read())
        resp.props["text"] = vm.native("text", text)

        def blob(this_, a):
            b = JSObject(vm.ObjectProto)
            b.props["__path"] = path
            return vm.promise_resolve(b)
        resp.props["blob"] = vm.native("blob", blob)
        return vm.promise_resolve(resp)

    def _create_image_bitmap(self, this, args):
        """the browser's image decoder: PIL here (a JPEG decoded by another library may differ by a grey level; both sides
        of every comparison get the pixels from this one decode, through the reference's ImageData)"""
        from PIL import Image
        vm = self.vm
        path = vm.get(args[0], "__path")
        with Image.open(path) as im:
            w, h = im.size
        bm = JSObject(vm.ObjectProto)
        bm.props.update({"width": float(w), "height": float(h), "__path": path})
        return vm.promise_resolve(bm)

    def _decode_image(self, this, args):
        from PIL import Image
        vm = self.vm
        with Image.open(vm.get(args[0], "__path")) as im:
            rgba = np.asarray(im.convert("RGBA"), dtype=np.uint8).reshape(-1)
        t = JSTypedArray(vm.root.vars["Uint8ClampedArray"].props["prototype"], "Uint8ClampedArray", 0)
        t.items.frombytes(rgba.tobytes())
        return t

    def _require(self, this, args):
        self.vm.throw("Error", "Cannot find module '%s'" % self.vm.tostr(args[0]))

    def _sample_done(self, this, args):
        if self._on_sample is not None:
            self._on_sample(int(args[0]), int(args[1]), args[2])
        return UNDEF

    # -- the worker's steps --------------------------------------------------------------------
    def load_test(self, name, config_seed=12345):
        """`import(testName).then(module => module.configureTest(test => ...))`, src/worker.js:23-24"""
        vm = self.vm
        self.cwd = test_dir(self.root, name)
        self.tape.seed(config_seed)          # scenes that place objects with Math.random() (tests/spheres*) stay reproducible
        with open(os.path.join(self.cwd, "test.mjs")) as fh:
            vm.run(fh.read(), name + "/test.mjs")
        vm.run("var __test = null; configureTest(function(t) { __test = t; });", "<configure>")
        vm.run_jobs()
        t = vm.root.vars.get("__test")
        if t is None or t is UNDEF:
            for p in vm.rejections:
                raise RuntimeError("configureTest failed: unhandled rejection: %s" % vm.inspect(p.value))
            raise RuntimeError("configureTest never called back (an asset failed to load?)")
        self.test = t
        return self.info()

    def load_wire(self, json_text):
        """`Serializer.deserializeJSON(text)` — how the reference's `*_json` scenes come to life (tests/dragon_json/test.mjs:1-9)"""
        vm = self.vm
        vm.root.vars["__wire_text"] = json_text
        vm.run("var __test = Serializer.deserializeJSON(__wire_text);", "<deserialize>")
        self.test = vm.root.vars["__test"]
        return self.info()

    def info(self):
        ev = self.vm.eval_expr
        return {
            "width": int(ev("__test.width")), "height": int(ev("__test.height")),
            "renderer": ev("__test.renderer.constructor.name"),
            "samplesPerPixel": int(ev("__test.renderer.samplesPerPixel || 1")),
            "maxRecursionDepth": int(ev("__test.renderer.maxRecursionDepth")),
        }

    def scene_json(self, width=None, height=None, triangle_fix=True):
        """the reference's own wire format of the configured test: `new Serializer(test)`, tests/test_to_json.js:30.
        triangle_fix: install js/cuda_renderer.js's `installTriangleSerializeFix()` and `installTextureSerializeFix()`
        first, as CUDARenderer does — without them the reference's serializer drops the vertex normals its own worker
        path shades with, and writes a texture's pixels as nothing (a browser ImageData has no own enumerable keys)."""
        if width is not None:
            self.vm.run("__test.width = %d; __test.height = %d;" % (width, height))
        if triangle_fix:
            self.vm.run("installTriangleSerializeFix(); installTextureSerializeFix();")
        return self.vm.eval_expr("JSON.stringify(new Serializer(__test).plain())")

    def camera_rays(self, width, height, passes, sample_random):
        """the camera rays of the reference's own render loop (src/renderers.js:87-98 -> camera.getRayForPixel): the world is
        replaced by a stub whose color() records the ray it is handed; returns (passes, H, W, 8): origin xyzw, direction xyzw"""
        vm = self.vm
        vm.run("var __rays = []; var __savedWorld = __test.renderer.world;"
               "__test.renderer.world = { color: function (ray, depth) { __rays.push(Array.from(ray.origin).concat(Array.from(ray.direction))); return Vec.of(0, 0, 0); } };")
        try:
            self.render(width, height, passes, sample_random=sample_random)
        finally:
            vm.run("__test.renderer.world = __savedWorld;")
        rays = np.array([r.items for r in vm.eval_expr("__rays").items], dtype=np.float64)
        # render order: pass, column, row
        return rays.reshape(passes, width, height, 8).transpose(0, 2, 1, 3)

    def render_simple(self, width, height):
        """the same world and camera through the reference's un-jittered `SimpleRenderer` (src/renderers.js:1-45) — for a
        scene without random decisions this is a deterministic image that any implementation can be compared with"""
        self.vm.run("var __saved = __test.renderer;"
                    "__test.renderer = new SimpleRenderer(__saved.world, __saved.camera, __saved.maxRecursionDepth);")
        try:
            return self.render(width, height, 1)
        finally:
            self.vm.run("__test.renderer = __saved;")

    def render_random(self, width, height, spp, seed=1):
        """the same world and camera through the reference's `RandomMultisamplingRenderer` (src/renderers.js:47-63): all
        samples of a pixel inside one getPixelColor, one setColor per pixel"""
        self.vm.run("var __saved = __test.renderer;"
                    "__test.renderer = new RandomMultisamplingRenderer(__saved.world, __saved.camera, %d, __saved.maxRecursionDepth);" % spp)
        try:
            return self.render(width, height, 1, seed=seed)
        finally:
            self.vm.run("__test.renderer = __saved;")

    def render(self, width, height, passes=1, x_offset=0, x_delt=1, seed=1, sample_random=None):
        """`test.renderer.render(new PixelBuffer(w, h), 1000, callback, workerIndex, workerCount)`, src/worker.js:26-32.
        sample_random(pixel, pass) -> iterable of floats: what Math.random() returns during that sample, instead of the tape.
        Returns (mean (H,W,3) f32 — the colour handed to the last setColor of each pixel, rgba8 (H,W,4) u8, draws per
        sample (H,W) of the last pass)."""
        vm = self.vm
        kind = self.info()["renderer"]
        incremental = kind == "IncrementalMultisamplingRenderer"
        if incremental:
            vm.run("__test.renderer.samplesPerPixel = %d;" % passes)
            n_iter = passes
        else:
            n_iter = 1
        cols = list(range(x_offset, width, x_delt))
        order = [(it, px, py) for it in range(n_iter) for px in cols for py in range(height)]
        mean = np.zeros((height, width, 3), dtype=np.float32)
        draws = np.zeros((height, width), dtype=np.int32)
        state = {"i": 0}
        tape = self.tape

        def arm(i):
            it, px, py = order[i]
            tape.draws = 0
            if sample_random is not None:
                vals = iter(sample_random(py * width + px, it))

                def rnd():
                    tape.draws += 1
                    return float(next(vals))
                vm.random = rnd
                return
            tape.seed(tape_seed(seed, py * width + px, it))

        def on_sample(x, y, color):
            i = state["i"]
            it, px, py = order[i]
            if (px, py) != (x, y):
                raise RuntimeError("sample order: expected pixel %r, renderer finished %r" % ((px, py), (x, y)))
            c = color.items
            mean[y, x, 0], mean[y, x, 1], mean[y, x, 2] = c[0], c[1], c[2]
            draws[y, x] = tape.draws
            state["i"] = i + 1
            if i + 1 < len(order):
                arm(i + 1)

        self._on_sample = on_sample
        arm(0)
        try:
            vm.run("var __buf = new __TapBuffer(%d, %d); __test.renderer.render(__buf, 0, false, %d, %d);"
                   % (width, height, x_offset, x_delt), "<render>")
        finally:
            self._on_sample = None
            vm.random = tape
        if state["i"] != len(order):
            raise RuntimeError("renderer finished %d samples, expected %d" % (state["i"], len(order)))
        data = vm.eval_expr("__buf.imgdata.data")
        rgba = np.array(data.items, dtype=np.uint8).reshape(height, width, 4)
        return mean, rgba, draws
