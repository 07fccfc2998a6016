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


class Material(JSObject):
    pass


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
