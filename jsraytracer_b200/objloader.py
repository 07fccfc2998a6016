"""OBJ / MTL loading (reference: src/objloader.js:9-20,58-116,144-268).

Restates `parseObjFile`'s semantics: fan triangulation of every `f` line,
`v/vt/vn` index triples, the `minArea` filter on `Triangle.area`, one
`Primitive` per triangle sharing `transform` / `inv_transform`, and the
`makeMaterial` quirk (src/objloader.js:16-19: the Fresnel / path-tracing
branch lacks `return`, so MTL materials are always `PhongMaterial`).

Parsing is split in two so that meshes can be committed as compact derived
fixtures (`scenes/data/*.npz`; `/root/reference` is absent on the GPU box):
`parse_obj_text` -> `ParsedObj` (arrays), `triangles_from_parsed` -> Primitives.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

from .jsmath import Vec, Mat4
from .geometry import Triangle
from .materials import (MaterialColor, ScaledMaterialColor, PhongMaterial, TextureMaterialColor, ImageData)
from .world import Primitive

_SKIP = re.compile(r"^\s*($|#)")
_IDX = re.compile(r"(\d+)(?:/(\d*)(?:/(\d+))?)?")


def _parse_float(tok: str) -> float:
    """Number.parseFloat: longest numeric prefix, NaN if none."""
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|Infinity)", tok)
    if not m:
        return math.nan
    s = m.group(0).strip()
    return float(s.replace("Infinity", "inf"))


@dataclass
class ParsedObj:
    positions: np.ndarray            # (P,4) f32
    texcoords: np.ndarray            # (T,3) f32
    normals: np.ndarray              # (N,4) f32
    faces: np.ndarray                # (F,3,3) int32: per corner (v, vt, vn); -1 = absent
    face_material: np.ndarray        # (F,) int32 index into material_names; -1 = default
    material_names: list = field(default_factory=list)
    mtllibs: list = field(default_factory=list)
    mtl_texts: list = field(default_factory=list)   # contents of the mtllibs (fixtures only; parse_obj_text leaves it empty)
    textures: dict = field(default_factory=dict)    # map_K* file name -> decoded (H,W,4) uint8 RGBA (fixtures only)

    def save(self, path):
        tex = {"tex_%d" % i: np.ascontiguousarray(a, dtype=np.uint8) for i, a in enumerate(self.textures.values())}
        np.savez_compressed(path, positions=self.positions, texcoords=self.texcoords, normals=self.normals,
                            faces=self.faces, face_material=self.face_material,
                            material_names=np.array(self.material_names, dtype="U"),
                            mtllibs=np.array(self.mtllibs, dtype="U"), mtl_texts=np.array(self.mtl_texts, dtype="U"),
                            texture_names=np.array(list(self.textures.keys()), dtype="U"), **tex)

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        return ParsedObj(z["positions"], z["texcoords"], z["normals"], z["faces"], z["face_material"],
                         [str(s) for s in z["material_names"]], [str(s) for s in z["mtllibs"]],
                         [str(s) for s in z["mtl_texts"]] if "mtl_texts" in z.files else [],
                         {str(n): z["tex_%d" % i] for i, n in enumerate(z["texture_names"])}
                         if "texture_names" in z.files else {})


def parse_obj_text(data: str) -> ParsedObj:
    """The geometry half of parseObjFile (src/objloader.js:149-238): already
    fan-triangulated (`for i = 2..`: [0, i-1, i], :190-191)."""
    positions, textures, normals, faces, fmat = [], [], [], [], []
    names, mtllibs = [], []
    cur = -1
    for l in data.split("\n"):
        if _SKIP.match(l):
            continue
        t = re.findall(r"\S+", l)
        if t[0] == "mtllib":
            mtllibs.append(t[1])
            continue
        if t[0] == "usemtl":
            if t[1] not in names:
                names.append(t[1])
            cur = names.index(t[1])
            continue
        if t[0] == "f":
            idx = []
            for tok in t[1:]:
                m = _IDX.search(tok)
                tri = []
                for g in m.groups():
                    tri.append(int(g) - 1 if g not in (None, "") else -1)
                idx.append(tri)
            for i in range(2, len(idx)):
                faces.append([idx[0], idx[i - 1], idx[i]])
                fmat.append(cur)
            continue
        nums = [_parse_float(x) for x in t[1:]]
        if t[0] == "v":
            positions.append([nums[0], nums[1], nums[2], 1 if len(t) < 5 else nums[3]])
        elif t[0] == "vt":
            g = lambda i: (nums[i] if i < len(nums) and nums[i] == nums[i] else 0) or 0
            textures.append([nums[0], g(1), g(2)])
        elif t[0] == "vn":
            normals.append([nums[0], nums[1], nums[2], 0])
        elif t[0] in ("s", "o", "g", "vp"):
            continue
        else:
            raise ValueError('Error while attempting to parse obj file on line "%s"' % l)
    return ParsedObj(np.asarray(positions, dtype=np.float32).reshape(-1, 4),
                     np.asarray(textures, dtype=np.float32).reshape(-1, 3),
                     np.asarray(normals, dtype=np.float32).reshape(-1, 4),
                     np.asarray(faces, dtype=np.int32).reshape(-1, 3, 3),
                     np.asarray(fmat, dtype=np.int32), names, mtllibs)


def parse_obj_text_native(data: str) -> ParsedObj:
    """The same through the native reader (`jsrt_obj_parse`, csrc/obj_parse.cpp): milliseconds where the
    line-by-line mirror above takes seconds on a 100 k-face mesh.  Raises ValueError with the reference's message."""
    import ctypes as C
    from . import lib
    L = lib.load()
    raw = data.encode("utf8") if isinstance(data, str) else bytes(data)
    h = L.jsrt_obj_parse(raw, len(raw))
    try:
        err = L.jsrt_obj_error(h)
        if err is not None:
            raise ValueError(err.decode("utf8", "replace"))
        cnt = (C.c_int32 * 6)()
        L.jsrt_obj_counts(h, cnt)
        npos, ntex, nnrm, ntri, nmat, nlib = list(cnt)
        pos = np.zeros((npos, 4), np.float32); tex = np.zeros((ntex, 3), np.float32); nrm = np.zeros((nnrm, 4), np.float32)
        faces = np.zeros((ntri, 3, 3), np.int32); fmat = np.zeros((ntri,), np.int32)
        L.jsrt_obj_copy(h, pos.ctypes.data, tex.ctypes.data, nrm.ctypes.data, faces.ctypes.data, fmat.ctypes.data)
        names = [L.jsrt_obj_material_name(h, i).decode("utf8") for i in range(nmat)]
        libs = [L.jsrt_obj_mtllib(h, i).decode("utf8") for i in range(nlib)]
        return ParsedObj(pos, tex, nrm, faces, fmat, names, libs)
    finally:
        L.jsrt_obj_free(h)


def triangles_from_parsed(parsed: ParsedObj, defaultMaterial=None, transform=None, minArea=0.0, materials=None):
    """The triangle-building half of parseObjFile (src/objloader.js:188-206)."""
    transform = transform if transform is not None else Mat4.identity()
    inv_transform = Mat4.inverse(transform)
    pos = [Vec(p) for p in parsed.positions.tolist()]
    tex = [Vec(p) for p in parsed.texcoords.tolist()]
    nrm = [Vec(p) for p in parsed.normals.tolist()]
    faces = parsed.faces.tolist()
    fmat = parsed.face_material.tolist()
    triangles = []
    for f, mi in zip(faces, fmat):
        data = {}
        if all(c[1] >= 0 for c in f):
            data["UV"] = [tex[c[1]] for c in f]
        if all(c[2] >= 0 for c in f):
            data["normal"] = [nrm[c[2]] for c in f]
        tri = Triangle([pos[c[0]] for c in f], data)
        if tri.area >= minArea:
            mat = defaultMaterial
            if mi >= 0:
                name = parsed.material_names[mi]
                if not materials or name not in materials:
                    raise ValueError("No material defined with name: " + name)
                mat = materials[name]
            triangles.append(Primitive(tri, mat, transform, inv_transform))
    return triangles


# -- MTL (src/objloader.js:1-20,58-116) --------------------------------------
def _make_material_color(a, b, default=None):
    default = default if default is not None else Vec.of(0, 0, 0)
    if isinstance(a, MaterialColor):
        # `new ScaledMaterialColor(a, b)` with b a Vec: Vec.times(Vec) is component-wise (src/math.js:209-211)
        return ScaledMaterialColor(a, [float(x) for x in b.v] if isinstance(b, Vec) else b) if b else a
    if a or b:
        # MaterialColor.coerce(a, b) with a undefined and b a Vec -> Solid(b)
        return MaterialColor.coerce2(a, b) if b is not None else MaterialColor.coerce(a)
    return MaterialColor.coerce(default)


def _make_material(d):
    ambient = _make_material_color(d.get("map_Ka"), d.get("Ka"))
    diffuse = _make_material_color(d.get("map_Kd"), d.get("Kd"))
    specular = _make_material_color(d.get("map_Ks"), d.get("Ks"))
    smoothness = d.get("Ns") or 0
    # src/objloader.js:16-19 constructs the Fresnel/path material without
    # returning it; every MTL material is a PhongMaterial.
    return PhongMaterial(Vec.of(1, 1, 1), ambient, diffuse, specular, smoothness)


def mtl_texture_names(text: str):
    """The pre-scan of parseMtlFile (src/objloader.js:61-70): file names the MTL's texture maps refer to."""
    names = []
    for l in text.split("\n"):
        if _SKIP.match(l):
            continue
        t = re.findall(r"\S+", l)
        if t[0] in ("map_Ka", "map_Kd", "map_Ks", "map_Ns"):
            names.append(t[-1])
    return names


def parse_mtl_text(text: str, textures=None):
    """parseMtlFile (src/objloader.js:58-116).  `textures`: file name -> TextureMaterialColor, the table
    `loadTextures` (:43-56) fills from decoded bitmaps; the browser's image decode has no counterpart here, so
    the caller supplies decoded RGBA (`TextureMaterialColor(ImageData.from_array(rgba))` = `fromBitmap`, :91-96)."""
    textures = textures or {}
    ret, curr, name = {}, None, None
    for l in text.split("\n"):
        if _SKIP.match(l):
            continue
        t = re.findall(r"\S+", l)
        if t[0] == "newmtl":
            if curr is not None:        # `if (curr)`: {} is truthy in JS
                ret[name] = _make_material(curr)
            name, curr = t[1], {}
            continue
        vals = [t[0]] + [(_parse_float(x) if _parse_float(x) == _parse_float(x) else x) for x in t[1:]]
        if vals[0] in ("Ka", "Kd", "Ks", "Ke", "Tf"):
            curr[vals[0]] = Vec.of(vals[1], vals[2], vals[3])
        elif vals[0] in ("Ns", "Ni", "illum", "d", "Tr"):
            curr[vals[0]] = vals[1]
        elif vals[0] in ("map_Ka", "map_Kd", "map_Ks"):
            if not textures.get(vals[-1]):
                raise ValueError("Unknown texture: " + str(vals[-1]))             # src/objloader.js:101-102
            curr[vals[0]] = textures[vals[-1]]
        else:
            raise ValueError("Unsupported material parameter: " + vals[0])
    if curr is not None:
        ret[name] = _make_material(curr)
    return ret


def loadObjFile(filename, defaultMaterial=None, transform=None, minArea=0.00001, decode_image=None):
    """src/objloader.js:240-247 (synchronous; returns the Primitive list).  `decode_image(path) -> (H,W,4) uint8`
    stands in for the browser's `createImageBitmap` (:34-41) when the MTL names texture maps."""
    with open(filename, "r", encoding="utf8") as fh:
        text = fh.read()
    parsed = parse_obj_text_native(text)
    prefix = filename[: filename.rfind("/") + 1] if filename.rfind("/") > 0 else ""
    materials = {}
    for lib in parsed.mtllibs:
        with open(prefix + lib, "r", encoding="utf8") as fh:
            mtl = fh.read()
        textures = {}
        for tname in mtl_texture_names(mtl):
            if decode_image is None:
                raise ValueError("MTL texture map %s: loadObjFile needs a decode_image callback" % tname)
            textures[tname] = TextureMaterialColor(ImageData.from_array(decode_image(prefix + tname)))
        materials.update(parse_mtl_text(mtl, textures))
    return triangles_from_parsed(parsed, defaultMaterial, transform, minArea, materials)
