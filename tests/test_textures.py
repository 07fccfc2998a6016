"""TextureMaterialColor (src/materials.js:77-131) and PositionalUVMaterial (:178-193) — SURVEY.md §8f items 2-3.

CPU: the wire form of ImageData (JSON array / msgpack bin), the flattener accepting both, and the oracle's texel
arithmetic against a numpy restatement of the reference formula on a Square seen head-on.
GPU: the CUDA path against the oracle on the `textured` scene (bilinear + nearest, clamp + wrap, scaled textures,
positional UVs, a texture on a geometry without UVs).
"""
import json
import math

import numpy as np
import pytest

from conftest import psnr, scene_blobs


def _js_mod(a, b):
    return math.fmod(a, b)


def _texture_color(img, u, v, mode, clamp_u, clamp_v):
    """src/materials.js:97-130, written out with Python floats (f64)."""
    h, w, _ = img.shape

    def norm(c, clamp):
        return min(max(c, 0.0), 1.0) if clamp else _js_mod(_js_mod(c, 1) + 1, 1)
    U, V = norm(u, clamp_u), norm(v, clamp_v)
    fx, fy = U * w - 0.5, (1.0 - V) * h - 0.5
    if mode == "bilinear":
        mx, my = math.floor(fx), math.floor(fy)
        xs = [(mx, 1 - _js_mod(fx, 1)), (mx + 1, _js_mod(fx, 1))]
        ys = [(my, 1 - _js_mod(fy, 1)), (my + 1, _js_mod(fy, 1))]
    else:
        xs, ys = [(math.floor(fx + 0.5), 1)], [(math.floor(fy + 0.5), 1)]
    ret = [0.0] * 4
    for px, bx in xs:
        for py, by in ys:
            cy, cx = int(min(max(py, 0), h - 1)), int(min(max(px, 0), w - 1))
            for i in range(4):
                ret[i] += bx * by * (img[cy, cx, i] / 255)
    return np.array(ret, dtype=np.float32)


def _square_scene(mode, clamp):
    from jsraytracer_b200.cameras import PerspectiveCamera
    from jsraytracer_b200.geometry import Square
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.materials import SolidColorMaterial, TextureMaterialColor
    from jsraytracer_b200.renderers import SimpleRenderer
    from jsraytracer_b200.scenes import procedural_texture
    from jsraytracer_b200.world import Primitive, World
    img = procedural_texture(7, 5)
    tex = TextureMaterialColor(img, mode, clamp, clamp)
    # a 3x3 square at z = -2 facing the camera: local (x, y) = world (x, y) / 3, UV = local (x, y) in [-0.5, 0.5]
    objs = [Primitive(Square(), SolidColorMaterial(tex), Mat4.translation([0, 0, -2]).times(Mat4.scale(3)))]
    cam = PerspectiveCamera(math.pi / 4, 1, Mat4.identity())
    return {"renderer": SimpleRenderer(World(objs, []), cam, 2), "width": 48, "height": 48}, img


@pytest.mark.parametrize("mode,clamp", [("bilinear", True), ("bilinear", False), ("nearest", True), ("nearest", False)])
def test_oracle_texture_lookup_matches_formula(mode, clamp):
    from jsraytracer_b200.serializer import Serializer
    from oracle.oracle import OracleScene
    test, imgdata = _square_scene(mode, clamp)
    img = np.frombuffer(imgdata.data, dtype=np.uint8).reshape(imgdata.height, imgdata.width, 4)
    orc = OracleScene(Serializer(test).to_json())
    acc, _ = orc.render(1, jitter=False)
    W = H = 48
    tan = math.tan(math.pi / 8)
    worst, seen = 0.0, 0
    for py in range(0, H, 2):
        for px in range(0, W, 2):
            x, y = 2 * px / W - 1, -2 * py / H + 1                    # src/renderers.js:22,25
            dx, dy = np.float32(x * tan), np.float32(y * tan)          # Vec.of stores f32
            t = 2.0                                                    # plane z = -2, direction z = -1
            lx, ly = np.float32(np.float32(dx * t) / 3), np.float32(np.float32(dy * t) / 3)
            if abs(lx) > 0.5 or abs(ly) > 0.5:
                continue
            want = _texture_color(img, float(lx), float(ly), mode, clamp, clamp)[:3]
            worst = max(worst, float(np.abs(acc[py, px] - want).max()))
            seen += 1
    assert seen > 200 and worst < 1e-6       # measured: bit-identical on all pixels, all four modes


def test_imagedata_wire_forms_and_flattener():
    import msgpack
    from jsraytracer_b200 import lib, scenes
    from jsraytracer_b200.serializer import Serializer
    ser = Serializer(scenes.configure("textured", width=32, height=32))
    js, mp = ser.to_json(), ser.to_msgpack()

    def find_imgdata(o):
        if isinstance(o, dict):
            v = o.get("_v")
            if isinstance(v, dict) and set(v) == {"width", "height", "data"}:
                return v
            for x in o.values():
                r = find_imgdata(x)
                if r:
                    return r
        elif isinstance(o, list):
            for x in o:
                r = find_imgdata(x)
                if r:
                    return r
        return None
    ij, im = find_imgdata(json.loads(js)), find_imgdata(msgpack.unpackb(mp))
    assert isinstance(ij["data"], list) and len(ij["data"]) == ij["width"] * ij["height"] * 4
    assert isinstance(im["data"], bytes) and list(im["data"]) == ij["data"]
    for blob, fmt in ((js, lib.FORMAT_JSON), (mp, lib.FORMAT_MSGPACK)):
        info = lib.Scene(blob, fmt, device=None).info       # parse + flatten only
        assert info["n_prims"] == 4 and info["n_materials"] == 5


def test_flattener_rejects_empty_imagedata():
    """A browser ImageData serialises empty in the reference (SURVEY.md §8b hazard 2): refuse it loudly."""
    from jsraytracer_b200 import lib, scenes
    from jsraytracer_b200.serializer import Serializer
    test = scenes.configure("textured", width=16, height=16)
    blob = json.loads(Serializer(test).to_json())

    def strip(o):
        if isinstance(o, dict):
            v = o.get("_v")
            if isinstance(v, dict) and set(v) == {"width", "height", "data"}:
                o["_v"] = {}
            for x in o.values():
                strip(x)
        elif isinstance(o, list):
            for x in o:
                strip(x)
    strip(blob)
    with pytest.raises(lib.JsrtError):
        lib.Scene(json.dumps(blob), lib.FORMAT_JSON, device=None)


@pytest.mark.gpu
def test_textured_scene_matches_oracle():
    from jsraytracer_b200 import lib
    from oracle.oracle import OracleScene
    js, mp = scene_blobs("textured", width=320, height=320)
    sc, orc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=0), OracleScene(js)
    ids, _ = sc.primary_hits()
    oids, _, _ = orc.primary_hits()
    assert float((ids == oids).mean()) >= 0.9999
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    acc, _ = sc.read_accum()
    oacc, _ = orc.render(1, jitter=False)
    assert psnr(acc[..., :3], oacc) >= 50, "whitted %.1f dB" % psnr(acc[..., :3], oacc)
    sc.reset_accum()
    sc.render(0, 4, seed=3)
    acc, _ = sc.read_accum()
    oacc, _ = orc.render(4, seed=3)
    assert psnr(acc[..., :3] / 4, oacc / 4) >= 50
