"""Scene-graph mirror of materials and lights (reference: src/materials.js,
src/lights.js).  Constructors and coercion rules only; `color()`,
`scatter()`, `sampleIterator()` run in the CUDA shade / shadow kernels.
"""
from __future__ import annotations

import math

from .jsmath import Vec, Mat4
from .geometry import JSObject, INF


class MaterialColor(JSObject):  # src/materials.js:2-26
    @staticmethod
    def coerce(baseColor, mc=None, _two=False):
        """src/materials.js:4-16.  `coerce(x)` and `coerce(base, x)`."""
        if not _two and mc is None:
            mc, baseColor = baseColor, None
        if isinstance(mc, MaterialColor):
            return mc
        if isinstance(mc, Vec):
            return SolidMaterialColor(mc)
        if isinstance(mc, (int, float, list, tuple)):
            return ScaledMaterialColor(MaterialColor.coerce(baseColor), mc)
        raise TypeError("Type provided that cannot be coerced to MaterialColor")

    @staticmethod
    def coerce2(baseColor, mc):
        return MaterialColor.coerce(baseColor, mc, _two=True)


class SolidMaterialColor(MaterialColor):  # src/materials.js:27-42
    JS_NAME = "SolidMaterialColor"
    White = None

    def __init__(self, color):
        self._color = color


SolidMaterialColor.White = SolidMaterialColor(Vec.of(1, 1, 1))


class ScaledMaterialColor(MaterialColor):  # src/materials.js:43-61
    JS_NAME = "ScaledMaterialColor"

    def __init__(self, mc, scale):
        self._mc = MaterialColor.coerce(mc)
        self._scale = list(scale) if isinstance(scale, (list, tuple)) else scale


class CheckerboardMaterialColor(MaterialColor):  # src/materials.js:63-76
    JS_NAME = "CheckerboardMaterialColor"

    def __init__(self, color1, color2):
        self.color1 = MaterialColor.coerce(color1)
        self.color2 = MaterialColor.coerce(color2)


class ImageData(JSObject):
    """Stand-in for the browser `ImageData` a TextureMaterialColor holds (`context.getImageData`,
    src/materials.js:91-96): width, height and RGBA8 `data`.  The reference's serializer writes a browser
    ImageData as an empty object (its fields are prototype getters — SURVEY.md §8b hazard 2), so the wire form is
    defined by this glue: `{width, height, data}` with `data` a msgpack bin / JSON array of w*h*4 bytes."""
    JS_NAME = "ImageData"

    def __init__(self, width, height, data):
        data = bytes(data)
        if len(data) != width * height * 4:
            raise ValueError("ImageData: data must hold width*height*4 bytes")
        self.width, self.height, self.data = int(width), int(height), data

    @staticmethod
    def from_array(rgba):
        """(H, W, 4) uint8 array -> ImageData."""
        h, w, c = rgba.shape
        if c != 4:
            raise ValueError("ImageData.from_array: need RGBA")
        return ImageData(w, h, rgba.astype("uint8").tobytes())

    def serialize(self, ser):
        return {"width": self.width, "height": self.height, "data": self.data}


class TextureMaterialColor(MaterialColor):  # src/materials.js:77-131
    JS_NAME = "TextureMaterialColor"

    def __init__(self, imgdata, mode="bilinear", clampU=True, clampV=True):
        if mode not in ("bilinear", "nearest"):
            raise ValueError("Unsupported texture mode " + str(mode))          # src/materials.js:119
        self._imgdata = imgdata
        self.width = imgdata.width
        self.height = imgdata.height
        self.mode = mode
        self.clampU = clampU
        self.clampV = clampV


class Material(JSObject):
    pass


class PositionalUVMaterial(Material):  # src/materials.js:178-193
    JS_NAME = "PositionalUVMaterial"

    def __init__(self, baseMaterial, origin=None, u_axis=None, v_axis=None):
        self.baseMaterial = baseMaterial
        self.origin = origin if origin is not None else Vec.of(0, 0, 0)
        self.u_axis = u_axis if u_axis is not None else Vec.of(1, 0, 0)
        self.v_axis = v_axis if v_axis is not None else Vec.of(0, 0, 1)


class SolidColorMaterial(Material):  # src/materials.js:145-156
    JS_NAME = "SolidColorMaterial"

    def __init__(self, color):
        self._color = MaterialColor.coerce(color)


class TransparentMaterial(Material):  # src/materials.js:160-174
    JS_NAME = "TransparentMaterial"

    def __init__(self, color, opacity):
        self._color = MaterialColor.coerce(color)
        self._opacity = opacity


class PhongMaterial(Material):  # src/materials.js:195-208
    JS_NAME = "PhongMaterial"

    def __init__(self, baseColor, ambient=1, diffusivity=0, specularity=0, smoothness=5, reflectivity=0,
                 transmissivity=0):
        self.baseColor = MaterialColor.coerce(baseColor)
        self.ambient = MaterialColor.coerce2(self.baseColor, ambient)
        self.diffusivity = MaterialColor.coerce2(self.baseColor, diffusivity)
        self.specularity = MaterialColor.coerce2(SolidMaterialColor.White, specularity)
        self.reflectivity = MaterialColor.coerce2(SolidMaterialColor.White, reflectivity)
        self.transmissivity = MaterialColor.coerce2(SolidMaterialColor.White, transmissivity)
        self.smoothness = smoothness


class FresnelPhongMaterial(PhongMaterial):  # src/materials.js:294-301
    JS_NAME = "FresnelPhongMaterial"

    def __init__(self, baseColor, ambient=1, diffusivity=0, specularity=0, smoothness=0, refractiveIndexRatio=1,
                 reflectivity=1, transmissivity=1):
        super().__init__(baseColor, ambient, diffusivity, specularity, smoothness, reflectivity, transmissivity)
        self.refractiveIndexRatio = refractiveIndexRatio


class PhongPathTracingMaterial(FresnelPhongMaterial):  # src/materials.js:389-396
    JS_NAME = "PhongPathTracingMaterial"

    def __init__(self, baseColor, ambient=1, diffusivity=0, specularity=0, smoothness=0,
                 refractiveIndexRatio=INF, mirrorProbability=0):
        super().__init__(baseColor, ambient, diffusivity, specularity, smoothness, refractiveIndexRatio,
                         Vec.of(1, 1, 1),
                         Vec.of(1, 1, 1) if math.isfinite(refractiveIndexRatio) else Vec.of(0, 0, 0))
        self.mirrorProbability = mirrorProbability


# -- lights (src/lights.js) -------------------------------------------------
class Light(JSObject):
    pass


class SimplePointLight(Light):  # src/lights.js:27-54
    JS_NAME = "SimplePointLight"

    def __init__(self, position, color_mc, intensity=1):
        self.position = position
        self.color_mc = MaterialColor.coerce2(color_mc, intensity)


class RandomSampleAreaLight(Light):  # src/lights.js:56-94
    JS_NAME = "RandomSampleAreaLight"

    def __init__(self, surface_geometry, transform, color_mc, intensity=1, samples=1):
        self.surface_geometry = surface_geometry
        self.transform = transform
        self.inv_transform = Mat4.inverse(transform)
        self.color_mc = MaterialColor.coerce2(color_mc, intensity)
        self.samples = samples
        self.aabb = surface_geometry.getBoundingBox(transform, Mat4.inverse(transform))
