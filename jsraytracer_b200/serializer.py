"""Wire format of the drop-in boundary (reference: src/serializer.js:1-63).

`Serializer(obj).plain()` reproduces the reference's object-graph encoding:
every object becomes `{_t, _v[, _r]}`; `_t` is `[className, typeIndex]` on the
type's first appearance and the bare index afterwards (src/serializer.js:38-41);
`_r` is a numeric id added retroactively when the object is referenced again,
later occurrences are `{_r: id}` (src/serializer.js:20-24); `_v` comes from the
object's own `serialize()` if it has one (Vec, Mat, AABB, Triangle), else the
array elements, else the own enumerable keys in assignment order.

The blob handed to `jsrt_scene_create` is this structure as JSON text
(`to_json`) or msgpack (`to_msgpack`, like tests/test_to_json.js:38).
"""
from __future__ import annotations

import json
import math

from .jsmath import Vec, Mat
from .geometry import JSObject


class Serializer:
    def __init__(self, data, reference_bug_compat=False):
        self.reference_bug_compat = reference_bug_compat
        self._ids = {}          # id(obj) -> ref id
        self._keepalive = []
        self._refs = {}         # ref id -> emitted dict
        self._ref_counts = {}
        self._type_map = {}
        self._ref_uid_gen = 0
        self.data = self.serialize_step(data)

    # src/serializer.js:12-60
    def serialize_step(self, obj):
        if obj is None or isinstance(obj, (bool, int, float, str, bytes)):      # bytes: RGBA texels (ImageData.data)
            return obj
        key = id(obj)
        rid = self._ids.get(key)
        if rid is not None:
            if self._ref_counts[rid] == 1:
                self._refs[rid]["_r"] = rid
            self._ref_counts[rid] += 1
            return {"_r": rid}

        self._ref_uid_gen += 1
        rid = self._ref_uid_gen
        self._ids[key] = rid
        self._keepalive.append(obj)
        self._ref_counts[rid] = 1
        ref = self._refs[rid] = {}

        name = _js_type_name(obj)
        if name in self._type_map:
            ref["_t"] = self._type_map[name]
        else:
            idx = len(self._type_map)
            self._type_map[name] = idx
            ref["_t"] = [name, idx]

        if isinstance(obj, Vec):
            ref["_v"] = [_num(x) for x in obj.v]           # src/math.js:161-163
        elif isinstance(obj, Mat):
            ref["_v"] = [[_num(x) for x in r] for r in obj.rows]   # src/math.js:308-310
        elif hasattr(obj, "serialize"):
            ref["_v"] = obj.serialize(self)
        elif isinstance(obj, (list, tuple)):
            ref["_v"] = [self.serialize_step(v) for v in obj]
        elif isinstance(obj, dict):
            ref["_v"] = {k: self.serialize_step(v) for k, v in obj.items()}
        elif isinstance(obj, JSObject):
            ref["_v"] = {k: self.serialize_step(v) for k, v in obj.js_items()}
        else:
            raise TypeError("Cannot serialize %r" % type(obj))
        return ref

    # -- the *_json scenes (tests/dragon_json/test.mjs, tests/toledo_json/test.mjs) --------------------
    @staticmethod
    def deserializeJSON(text):
        """`Serializer.deserializeJSON(text)` (src/serializer.js:69-75) as the *_json scenes use it: the
        `{renderer, width, height}` of a scene that exists only in wire form.  The reference rebuilds the JS
        object graph; here the blob itself is what the C ABI consumes, so the returned renderer keeps it and
        `render()` hands it over untouched (`renderers.WireRenderer`)."""
        from .renderers import WireRenderer
        return WireRenderer.from_blob(text.encode("utf8") if isinstance(text, str) else bytes(text), 0)

    @staticmethod
    def deserializeMsgpack(blob):
        """The msgpack form of the same (tests/toledo_json/test.mjs:4-12, tests/test_to_json.js:38)."""
        from .renderers import WireRenderer
        return WireRenderer.from_blob(bytes(blob), 1)

    def plain(self):
        return self.data

    def to_json(self, js_compatible=False) -> str:
        """`JSON.stringify(plain)`.  JS writes non-finite numbers as `null`
        (`js_compatible=True`); the default keeps them as the `Infinity` /
        `-Infinity` / `NaN` tokens, which both readers in this repo accept and
        which lose nothing (IOR and SDF box sizes are legitimately infinite)."""
        if js_compatible:
            return json.dumps(_nonfinite_to_null(self.data), separators=(",", ":"), default=list)
        return json.dumps(self.data, separators=(",", ":"), default=list)      # default: bytes -> array of numbers

    def to_msgpack(self) -> bytes:
        import msgpack
        return msgpack.packb(self.data, use_single_float=False)


def _js_type_name(obj):
    if isinstance(obj, Vec):
        return "Vec"
    if isinstance(obj, Mat):
        return "Mat"
    if isinstance(obj, (list, tuple)):
        return "Array"
    if isinstance(obj, dict):
        return "Object"
    return obj.JS_NAME


def _num(x):
    """JS numbers print integer values without a fraction."""
    if isinstance(x, float) and math.isfinite(x) and x == int(x) and abs(x) < 2 ** 53:
        return int(x)
    return x


def _nonfinite_to_null(o):
    if isinstance(o, float) and not math.isfinite(o):
        return None
    if isinstance(o, list):
        return [_nonfinite_to_null(v) for v in o]
    if isinstance(o, dict):
        return {k: _nonfinite_to_null(v) for k, v in o.items()}
    return o
