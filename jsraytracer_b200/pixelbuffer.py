"""Mirror of `PixelBuffer` / browser `ImageData` (reference: src/pixelbuffer.js).
RGBA8, row-major, `coord(x, y) = y * (width * 4) + x * 4` (src/pixelbuffer.js:32-34)."""
from __future__ import annotations

import numpy as np


class ImageData:
    def __init__(self, width, height):
        self.width = int(width)
        self.height = int(height)
        self.data = np.zeros(self.width * self.height * 4, dtype=np.uint8)   # Uint8ClampedArray


class PixelBuffer:
    def __init__(self, width, height=None):
        if isinstance(width, ImageData):
            self.imgdata = width
        else:
            self.imgdata = ImageData(width, height)

    def width(self):
        return self.imgdata.width

    def height(self):
        return self.imgdata.height

    def coord(self, x, y):
        return y * (self.imgdata.width * 4) + x * 4

    def getColor(self, x, y):
        r = self.coord(x, y)
        return [float(self.imgdata.data[r + i]) / 255 for i in range(4)]

    def as_array(self):
        """(H, W, 4) uint8 view of the image."""
        return self.imgdata.data.reshape(self.imgdata.height, self.imgdata.width, 4)
