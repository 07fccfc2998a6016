"""Scene-graph mirror of the cameras (reference: src/cameras.js).
`getRayForPixel` runs in the CUDA generate kernel."""
from __future__ import annotations

import math

from .jsmath import Mat4
from .geometry import JSObject


class Camera(JSObject):  # src/cameras.js:1-16
    def __init__(self, transform, inv_transform=None):
        self.transform = transform
        self.inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)

    def setTransform(self, transform, inv_transform=None):
        self.transform = transform
        self.inv_transform = inv_transform if inv_transform is not None else Mat4.inverse(transform)


class PerspectiveCamera(Camera):  # src/cameras.js:18-39
    JS_NAME = "PerspectiveCamera"

    def __init__(self, fov, aspect, transform):
        super().__init__(transform)
        self.FOV = fov
        self.tan_fov = math.tan(fov / 2)
        self.aspect = aspect

    def changeFOV(self, fov):
        self.FOV = fov
        self.tan_fov = math.tan(fov / 2)


class DepthOfFieldPerspectiveCamera(PerspectiveCamera):  # src/cameras.js:41-53
    JS_NAME = "DepthOfFieldPerspectiveCamera"

    def __init__(self, fov, aspect, transform, focus_distance, sensor_size):
        super().__init__(fov, aspect, transform)
        self.focus_distance = focus_distance
        self.sensor_size = sensor_size
