"""Host-side mirror of the reference's numeric model (reference: src/math.js).

`Vec` follows `class Vec extends Float32Array` (src/math.js:160-286): every
vector-returning operation is evaluated in f64 and rounded to f32 on store.
`Mat` follows `class Mat extends Array` (src/math.js:303-450): rows of f64.

This module is *scene construction* code (cameras, transforms, bounding boxes,
BVH build inputs).  Nothing here runs per ray; the per-ray path is CUDA only.
"""
from __future__ import annotations

import math
import struct

import numpy as np

_f32 = np.float32


def f32(x: float) -> float:
    """Round a Python float (f64) to the nearest f32, returned as f64."""
    if x != x or x in (math.inf, -math.inf):
        return x
    try:
        return struct.unpack("f", struct.pack("f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


class Vec:
    """f32-stored vector of any small length (src/math.js:160)."""

    __slots__ = ("v",)

    def __init__(self, comps):
        self.v = [f32(float(c)) for c in comps]

    # -- construction -----------------------------------------------------
    @staticmethod
    def of(*comps) -> "Vec":
        return Vec(comps)

    @staticmethod
    def from_(seq) -> "Vec":
        return Vec(list(seq))

    @staticmethod
    def axis(axis, dim, amt=1, default_val=0) -> "Vec":  # src/math.js:170-174
        c = [default_val] * dim
        c[axis] = amt
        return Vec(c)

    def copy(self) -> "Vec":
        return Vec(self.v)

    # -- container protocol -----------------------------------------------
    def __len__(self):
        return len(self.v)

    def __getitem__(self, i):
        return self.v[i]

    def __setitem__(self, i, x):
        self.v[i] = f32(float(x))

    def __iter__(self):
        return iter(self.v)

    def __repr__(self):
        return "Vec(" + ", ".join(repr(c) for c in self.v) + ")"

    def tolist(self):
        return list(self.v)

    # -- arithmetic (src/math.js:197-235) ---------------------------------
    @staticmethod
    def _is_seq(b):
        return isinstance(b, (Vec, list, tuple))

    def _zip(self, b, fn):
        # Float32Array.map iterates over *this*; b[i] beyond b's length is
        # undefined -> NaN in JS.
        if Vec._is_seq(b) and len(b):
            out = []
            for i, x in enumerate(self.v):
                out.append(fn(x, b[i]) if i < len(b) else math.nan)
            return Vec(out)
        return Vec([fn(x, b) for x in self.v])

    def plus(self, b):
        return self._zip(b, lambda x, y: x + y)

    def minus(self, b):
        return self._zip(b, lambda x, y: x - y)

    def times(self, s):
        return self._zip(s, lambda x, y: x * y)

    def mult_pairs(self, b):
        return self._zip(b, lambda x, y: x * y)

    def divide(self, s):
        return self._zip(s, _jsdiv)

    def inverse(self, nom=1):  # src/math.js:215-217
        return Vec([_jsdiv(nom, x) for x in self.v])

    def abs(self):
        return Vec([abs(x) for x in self.v])

    def mix(self, b, s):  # src/math.js:233-235
        return Vec([_jsmul(1 - s, x) + _jsmul(s, b[i]) for i, x in enumerate(self.v)])

    def scale(self, s):
        self.v = [f32(x * s) for x in self.v]

    @staticmethod
    def min(a, b):
        return Vec([min(x, b[i] if Vec._is_seq(b) else b) for i, x in enumerate(a.v)])

    @staticmethod
    def max(a, b):
        return Vec([max(x, b[i] if Vec._is_seq(b) else b) for i, x in enumerate(a.v)])

    def dot(self, b):  # src/math.js:252-260 (f64 accumulation, left to right)
        n = len(self.v)
        if n == 3:
            return self.v[0] * b[0] + self.v[1] * b[1] + self.v[2] * b[2]
        if n == 4:
            return self.v[0] * b[0] + self.v[1] * b[1] + self.v[2] * b[2] + self.v[3] * b[3]
        if n > 4:
            acc = 0.0
            for i, x in enumerate(self.v):
                acc = acc + x * b[i]
            return acc
        return self.v[0] * b[0] + self.v[1] * b[1]

    def squarednorm(self):
        return self.dot(self)

    def norm(self):
        return math.sqrt(self.dot(self))

    def normalized(self):  # src/math.js:242-245
        n = self.norm()
        return self.times(1 / n) if n > 0.00001 else self

    def sum(self):
        acc = 0.0
        for x in self.v:
            acc += x
        return acc

    def average(self):
        return self.sum() / len(self.v) if len(self.v) else 0

    def to3(self):  # src/math.js:271-273
        g = lambda i: (self.v[i] if i < len(self.v) else 0) or 0
        return Vec([self.v[0], g(1), g(2)])

    def to4(self, is_point=False):  # src/math.js:274-276
        g = lambda i: (self.v[i] if i < len(self.v) else 0) or 0
        return Vec([self.v[0], g(1), g(2), 1 if is_point else 0])

    def cross(self, b):
        a = self.v
        return Vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def equals(self, b):
        return all(x == b[i] for i, x in enumerate(self.v))

    def every(self, pred):
        return all(pred(x) for x in self.v)


def _jsdiv(a, b):
    """IEEE division with JS semantics for division by zero."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def _jsmul(a, b):
    """IEEE multiply (0 * inf = NaN) — Python already does this for floats."""
    return a * b


class Ray:  # src/math.js:289-300
    __slots__ = ("origin", "direction")

    def __init__(self, origin, direction):
        self.origin = origin
        self.direction = direction

    def getTransformed(self, m):
        return Ray(m.times(self.origin), m.times(self.direction))

    def getPoint(self, t):
        return self.origin.plus(self.direction.times(t))


class Mat:
    """Row-major list of f64 rows (src/math.js:303-450)."""

    __slots__ = ("rows",)

    def __init__(self, *rows):
        self.rows = [[float(x) for x in r] for r in rows]

    @staticmethod
    def of(*rows):
        return Mat(*rows)

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return self.rows[i]

    def __iter__(self):
        return iter(self.rows)

    def __repr__(self):
        return "Mat(" + ", ".join(repr(r) for r in self.rows) + ")"

    def copy(self):
        return Mat(*self.rows)

    def column(self, index):  # src/math.js:362-364
        return Vec([r[index] for r in self.rows])

    def row(self, index):
        return Vec(self.rows[index])

    def transposed(self):
        n, m = len(self.rows), len(self.rows[0])
        return Mat(*[[self.rows[j][i] for j in range(n)] for i in range(m)])

    def times(self, b):  # src/math.js:385-410
        if isinstance(b, (int, float)):
            return Mat(*[[b * x for x in r] for r in self.rows])
        if isinstance(b, Vec) or (isinstance(b, (list, tuple)) and not isinstance(b[0], (list, tuple))):
            bv = b if isinstance(b, Vec) else Vec(b)
            ln = len(bv)
            res = [0.0] * len(self.rows)
            for r in range(min(ln, len(self.rows))):
                res[r] = bv.dot(self.rows[r])
            return Vec(res)
        brow = b.rows if isinstance(b, Mat) else b
        ln, len2 = len(brow), len(brow[0])
        out = []
        for r in range(len(self.rows)):
            row = []
            for c in range(len2):
                acc = 0.0
                for r2 in range(ln):
                    acc += self.rows[r][r2] * brow[r2][c]
                row.append(acc)
            out.append(row)
        return Mat(*out)

    def equals(self, b):
        return all(x == b[i][j] for i, r in enumerate(self.rows) for j, x in enumerate(r))


class Mat4:
    """Static 4x4 factories (src/math.js:487-584)."""

    @staticmethod
    def identity():
        return Mat.of([1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1])

    @staticmethod
    def rotation(angle, axis):  # src/math.js:492-499
        ax = axis if isinstance(axis, Vec) else Vec(axis)
        x, y, z = list(ax.normalized())[:3]
        c, s = math.cos(angle), math.sin(angle)
        omc = 1.0 - c
        return Mat.of(
            [x * x * omc + c, x * y * omc - z * s, x * z * omc + y * s, 0],
            [x * y * omc + z * s, y * y * omc + c, y * z * omc - x * s, 0],
            [x * z * omc - y * s, y * z * omc + x * s, z * z * omc + c, 0],
            [0, 0, 0, 1])

    @staticmethod
    def rotationX(angle):
        c, s = math.cos(angle), math.sin(angle)
        return Mat.of([1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1])

    @staticmethod
    def rotationY(angle):
        c, s = math.cos(angle), math.sin(angle)
        return Mat.of([c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1])

    @staticmethod
    def rotationZ(angle):
        c, s = math.cos(angle), math.sin(angle)
        return Mat.of([c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1])

    @staticmethod
    def scale(s):  # src/math.js:526-530
        if isinstance(s, (int, float)):
            s = [s, s, s]
        return Mat.of([s[0], 0, 0, 0], [0, s[1], 0, 0], [0, 0, s[2], 0], [0, 0, 0, 1])

    @staticmethod
    def translation(t):
        return Mat.of([1, 0, 0, t[0]], [0, 1, 0, t[1]], [0, 0, 1, t[2]], [0, 0, 0, 1])

    @staticmethod
    def eulerRotation(angles, order="YXZ"):  # src/math.js:586-597
        ax, ay, az = angles
        mats = {"X": Mat4.rotationX(ax) if ax else None,
                "Y": Mat4.rotationY(ay) if ay else None,
                "Z": Mat4.rotationZ(az) if az else None}
        ret = Mat4.identity()
        for k in reversed(order):
            if mats[k] is not None:
                ret = mats[k].times(ret)
        return ret

    @staticmethod
    def inverse(m):  # src/math.js:560-584 (cofactor form, same operation order)
        (m00, m01, m02, m03), (m10, m11, m12, m13), (m20, m21, m22, m23), (m30, m31, m32, m33) = \
            [tuple(r) for r in (m[0], m[1], m[2], m[3])]
        r = [[0.0] * 4 for _ in range(4)]
        r[0][0] = m12 * m23 * m31 - m13 * m22 * m31 + m13 * m21 * m32 - m11 * m23 * m32 - m12 * m21 * m33 + m11 * m22 * m33
        r[0][1] = m03 * m22 * m31 - m02 * m23 * m31 - m03 * m21 * m32 + m01 * m23 * m32 + m02 * m21 * m33 - m01 * m22 * m33
        r[0][2] = m02 * m13 * m31 - m03 * m12 * m31 + m03 * m11 * m32 - m01 * m13 * m32 - m02 * m11 * m33 + m01 * m12 * m33
        r[0][3] = m03 * m12 * m21 - m02 * m13 * m21 - m03 * m11 * m22 + m01 * m13 * m22 + m02 * m11 * m23 - m01 * m12 * m23
        r[1][0] = m13 * m22 * m30 - m12 * m23 * m30 - m13 * m20 * m32 + m10 * m23 * m32 + m12 * m20 * m33 - m10 * m22 * m33
        r[1][1] = m02 * m23 * m30 - m03 * m22 * m30 + m03 * m20 * m32 - m00 * m23 * m32 - m02 * m20 * m33 + m00 * m22 * m33
        r[1][2] = m03 * m12 * m30 - m02 * m13 * m30 - m03 * m10 * m32 + m00 * m13 * m32 + m02 * m10 * m33 - m00 * m12 * m33
        r[1][3] = m02 * m13 * m20 - m03 * m12 * m20 + m03 * m10 * m22 - m00 * m13 * m22 - m02 * m10 * m23 + m00 * m12 * m23
        r[2][0] = m11 * m23 * m30 - m13 * m21 * m30 + m13 * m20 * m31 - m10 * m23 * m31 - m11 * m20 * m33 + m10 * m21 * m33
        r[2][1] = m03 * m21 * m30 - m01 * m23 * m30 - m03 * m20 * m31 + m00 * m23 * m31 + m01 * m20 * m33 - m00 * m21 * m33
        r[2][2] = m01 * m13 * m30 - m03 * m11 * m30 + m03 * m10 * m31 - m00 * m13 * m31 - m01 * m10 * m33 + m00 * m11 * m33
        r[2][3] = m03 * m11 * m20 - m01 * m13 * m20 - m03 * m10 * m21 + m00 * m13 * m21 + m01 * m10 * m23 - m00 * m11 * m23
        r[3][0] = m12 * m21 * m30 - m11 * m22 * m30 - m12 * m20 * m31 + m10 * m22 * m31 + m11 * m20 * m32 - m10 * m21 * m32
        r[3][1] = m01 * m22 * m30 - m02 * m21 * m30 + m02 * m20 * m31 - m00 * m22 * m31 - m01 * m20 * m32 + m00 * m21 * m32
        r[3][2] = m02 * m11 * m30 - m01 * m12 * m30 - m02 * m10 * m31 + m00 * m12 * m31 + m01 * m10 * m32 - m00 * m11 * m32
        r[3][3] = m01 * m12 * m20 - m02 * m11 * m20 + m02 * m10 * m21 - m00 * m12 * m21 - m01 * m10 * m22 + m00 * m11 * m22
        det = m00 * r[0][0] + m10 * r[0][1] + m20 * r[0][2] + m30 * r[0][3]
        inv = _jsdiv(1, det)
        return Mat(*[[inv * x for x in row] for row in r])


# -- median / quickselect (src/math.js:95-157); used by the BVH build ------
# JS arrays return `undefined` for out-of-range reads and grow on out-of-range
# writes; `undefined` compares like NaN (every relational test is false) and
# `number + undefined` is NaN, so NaN stands in for `undefined` here.
def _cmp(a, b):
    return -1 if a < b else (1 if a > b else 0)


def _get(arr, i):
    return arr[i] if 0 <= i < len(arr) else math.nan


def _set(arr, i, x):
    while len(arr) <= i:
        arr.append(math.nan)
    arr[i] = x


def _swap(arr, i, j):
    tmp = _get(arr, i)
    _set(arr, i, _get(arr, j))
    _set(arr, j, tmp)


def _quick_select_step(arr, k, left=0, right=None):
    if right is None:
        right = len(arr) - 1
    while right > left:
        if right - left > 600:
            n = right - left + 1
            m = k - left + 1
            z = math.log(n)
            s = 0.5 * math.exp(2 * z / 3)
            sd = 0.5 * math.sqrt(z * s * (n - s) / n) * (-1 if m - n / 2 < 0 else 1)
            new_left = max(left, math.floor(k - m * s / n + sd))
            new_right = min(right, math.floor(k + (n - m) * s / n + sd))
            _quick_select_step(arr, k, new_left, new_right)
        t = _get(arr, k)
        i, j = left, right
        _swap(arr, left, k)
        if _cmp(_get(arr, right), t) > 0:
            _swap(arr, left, right)
        while i < j:
            _swap(arr, i, j)
            i += 1
            j -= 1
            while _cmp(_get(arr, i), t) < 0:
                i += 1
            while _cmp(_get(arr, j), t) > 0:
                j -= 1
        if _cmp(_get(arr, left), t) == 0:
            _swap(arr, left, j)
        else:
            j += 1
            _swap(arr, j, right)
        if j <= k:
            left = j + 1
        if k <= j:
            right = j - 1


def quick_select(arr, k):
    _quick_select_step(arr, k)
    return _get(arr, k)


def median(arr):
    """src/math.js:150-157.  Mutates `arr` like the reference.  For even
    lengths the reference averages elements len/2 and len/2+1 (sic), which
    reads one past the end when len == 2; `_get/_set` reproduce what a JS
    array does then."""
    if len(arr) == 0:
        return math.nan
    n = len(arr)
    len2 = n // 2
    if n % 2 == 1:
        return quick_select(arr, len2)
    return (quick_select(arr, len2) + quick_select(arr, len2 + 1)) / 2


def js_to_precision8(x: float) -> float:
    """Number(x.toPrecision(8)) — used by Math.fmod (src/math.js:27).  toPrecision rounds the exact decimal expansion
    and breaks exact ties upwards in magnitude (ECMA-262 21.1.3.5: "pick the larger n"), where "%.7e" goes to even."""
    if x != x or x in (math.inf, -math.inf) or x == 0:
        return x
    from decimal import Decimal, ROUND_HALF_UP
    d = Decimal(abs(x))
    r = d.quantize(Decimal(1).scaleb(d.adjusted() - 7), rounding=ROUND_HALF_UP)
    return math.copysign(float(r), x)


def fmod(a, b):  # src/math.js:27
    return js_to_precision8(a - (math.floor(_jsdiv(a, b)) * b))
