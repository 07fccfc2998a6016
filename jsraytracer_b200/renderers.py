"""Renderer classes (reference: src/renderers.js).

`SimpleRenderer`, `RandomMultisamplingRenderer` and
`IncrementalMultisamplingRenderer` are configuration mirrors: they carry
`{world, camera, maxRecursionDepth[, samplesPerPixel]}` exactly as the
reference's constructors do (src/renderers.js:2-6,48-51,66-69) and serialise
under the reference's class names.  Their `render()` is the CUDA path —
there is no CPU renderer in this package.

`CUDARenderer` is the new class that sits next to the web-worker renderer: the
same constructor shape and the same
`render(img, timelimit=0, callback=False, x_offset=0, x_delt=1)` signature as
`IncrementalMultisamplingRenderer.render` (src/renderers.js:70), driving the
C-ABI library (include/jsrt.h) instead of the per-pixel JS loops.
"""
from __future__ import annotations

import time

from .geometry import JSObject


class SimpleRenderer(JSObject):  # src/renderers.js:1-45
    JS_NAME = "SimpleRenderer"

    def __init__(self, world, camera, maxRecursionDepth=3):
        self.world = world
        self.camera = camera
        self.maxRecursionDepth = maxRecursionDepth

    @staticmethod
    def computePixelCount(img, x_offset, x_delt):  # src/renderers.js:7-9
        # Math.round rounds half up
        import math
        return ((img.width() / x_delt) + math.floor(1 - x_offset / x_delt + 0.5) * (img.width() % x_delt)) * img.height()

    def js_items(self):
        return [(k, v) for k, v in self.__dict__.items() if not k.startswith("_")]

    def render(self, img, timelimit=0, callback=False, x_offset=0, x_delt=1):
        return _cuda_render(self, img, timelimit, callback, x_offset, x_delt)


class RandomMultisamplingRenderer(SimpleRenderer):  # src/renderers.js:47-63
    JS_NAME = "RandomMultisamplingRenderer"

    def __init__(self, world, camera, samplesPerPixel, maxRecursionDepth=3):
        super().__init__(world, camera, maxRecursionDepth)
        self.samplesPerPixel = samplesPerPixel


class IncrementalMultisamplingRenderer(SimpleRenderer):  # src/renderers.js:65-117
    JS_NAME = "IncrementalMultisamplingRenderer"

    def __init__(self, world, camera, samplesPerPixel, maxRecursionDepth=3):
        super().__init__(world, camera, maxRecursionDepth)
        self.samplesPerPixel = samplesPerPixel


class CUDARenderer(IncrementalMultisamplingRenderer):
    """Drop-in for `IncrementalMultisamplingRenderer`: per-pass wavefront
    rendering on the GPU with the accumulation buffer resident in HBM."""
    JS_NAME = "CUDARenderer"

    def __init__(self, world, camera, samplesPerPixel, maxRecursionDepth=3, seed=1, device=0,
                 passes_per_call=None):
        super().__init__(world, camera, samplesPerPixel, maxRecursionDepth)
        self._seed = seed
        self._device = device
        self._passes_per_call = passes_per_call
        self._scene = None

    def close(self):
        if self._scene is not None:
            self._scene.close()
            self._scene = None

    def invalidate(self):
        """The uploaded scene is kept between render() calls; the reference's renderers re-read `world` / `camera` on
        every call, so after editing either (camera.setTransform, a material) call this: the next render() serialises again."""
        self.close()


class WireRenderer:
    """The renderer of a scene that only exists as a wire blob (`Serializer.deserializeJSON`, the reference's
    tests/dragon_json and tests/toledo_json): same `render(img, timelimit, callback, x_offset, x_delt)`; class,
    samplesPerPixel and maxRecursionDepth are read from the blob by the library's own parser (host-only, no CUDA)."""

    def __init__(self, blob, fmt, info, seed=1, device=0):
        self._wire = (blob, fmt)
        self.samplesPerPixel = info["samples_per_pixel"]
        self.maxRecursionDepth = info["max_depth"]
        self._jitter = bool(info["jitter"])          # False for a serialised SimpleRenderer (src/renderers.js:21-25)
        self._seed, self._device, self._passes_per_call, self._scene = seed, device, None, None

    @staticmethod
    def from_blob(blob, fmt):
        from . import lib
        host = lib.Scene(blob, fmt, device=None)     # parse + flatten only: validates the blob, yields the header fields
        info = host.info
        host.close()
        return {"renderer": WireRenderer(blob, fmt, info), "width": info["width"], "height": info["height"]}

    def render(self, img, timelimit=0, callback=False, x_offset=0, x_delt=1):
        return _cuda_render(self, img, timelimit, callback, x_offset, x_delt)

    def close(self):
        if self._scene is not None:
            self._scene.close()
            self._scene = None


def _cuda_render(renderer, img, timelimit, callback, x_offset, x_delt):
    """Shared body of every renderer's `render()` (src/renderers.js:10-41,70-117):
    serialise `{renderer, width, height}`, hand it to the C-ABI library, run the
    passes on the GPU, resolve into `img.imgdata.data` (PixelBuffer.setColor
    semantics, src/pixelbuffer.js:39-49), and call `callback({pass, completion})`
    at most every `timelimit` ms (src/renderers.js:103-112)."""
    from . import lib
    from .serializer import Serializer

    scene = getattr(renderer, "_scene", None)
    if scene is None or scene.size != (img.width(), img.height()):
        wire = getattr(renderer, "_wire", None)
        if wire is not None:                    # a scene that only exists in wire form: handed over as is
            blob, fmt = wire
        else:
            blob, fmt = Serializer({"renderer": renderer, "width": img.width(), "height": img.height()}).to_msgpack(), lib.FORMAT_MSGPACK
        scene = lib.Scene(blob, fmt, device=getattr(renderer, "_device", 0))
        if scene.size != (img.width(), img.height()):
            scene.close()
            raise ValueError("render: the serialised scene is %dx%d, the image %dx%d" % (scene.size + (img.width(), img.height())))
        if hasattr(renderer, "_scene"):
            renderer._scene = scene
    spp = getattr(renderer, "samplesPerPixel", 1)
    jitter = getattr(renderer, "_jitter", not type(renderer) is SimpleRenderer)
    if not jitter:
        spp = 1
    flags = 0 if jitter else lib.FLAG_NO_JITTER
    seed = getattr(renderer, "_seed", 1)
    # One library call per pass, like the reference's loop (src/renderers.js:87): the library coalesces consecutive
    # calls into full waves.  With a progress callback the passes go out in groups of 8 and the host synchronises once
    # per group — not per pass — to look at the clock (src/renderers.js:103-112: at most one callback per `timelimit` ms).
    group = getattr(renderer, "_passes_per_call", None) or 8
    scene.reset_accum()
    last = time.monotonic()
    done = 0
    while done < spp:
        scene.render(done, 1, seed, x_offset, x_delt, flags)
        done += 1
        if timelimit and callback and (done % group == 0 or done == spp):
            scene.synchronize()
            now = time.monotonic()
            if (now - last) * 1000.0 >= timelimit:
                last = now
                scene.resolve_rgba8(img.imgdata.data)
                callback({"pass": done - 1, "completion": done / spp})
    scene.resolve_rgba8(img.imgdata.data)
    return img
