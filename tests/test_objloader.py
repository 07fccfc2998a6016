"""OBJ loading restatement (src/objloader.js:144-247) and the committed mesh fixtures."""
import os
import sys

import numpy as np
import pytest

from jsraytracer_b200 import scenes
from jsraytracer_b200.jsmath import Vec
from jsraytracer_b200.objloader import ParsedObj, parse_obj_text, triangles_from_parsed, parse_mtl_text
from jsraytracer_b200.materials import PhongMaterial

REF_ASSETS = "/root/reference/assets"


def test_fan_triangulation_indices_and_min_area():
    text = "\n".join([
        "# comment", "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 0 0 1e-9",
        "vn 0 0 1", "vt 0.5 0.25",
        "f 1//1 2//1 3//1 4//1",        # quad -> two triangles [0,1,2], [0,2,3]
        "f 1/1/1 2/1/1 3/1/1",
        "f 1 2 5",                      # needle: area ~ 5e-10 < minArea
        "s off", "g grp", ""])
    p = parse_obj_text(text)
    assert p.faces.shape == (4, 3, 3)
    assert p.faces[0, :, 0].tolist() == [0, 1, 2] and p.faces[1, :, 0].tolist() == [0, 2, 3]
    assert p.faces[0, 0].tolist() == [0, -1, 0] and p.faces[2, 0].tolist() == [0, 0, 0]
    tris = triangles_from_parsed(p, PhongMaterial(Vec.of(1, 1, 1)), minArea=0.00001)
    assert len(tris) == 3
    assert "normal" in tris[0].geometry.psdata and "UV" not in tris[0].geometry.psdata
    assert list(tris[2].geometry.psdata.keys()) == ["UV", "normal"]          # insertion order of src/objloader.js:193-196
    assert tris[0].transform is tris[1].transform                             # one shared Mat per file
    with pytest.raises(ValueError):
        parse_obj_text("bogus 1 2 3")


def test_mtl_materials_are_always_phong():
    mats = parse_mtl_text("newmtl a\nKa 0 0 0\nKd 0.3 0.8 0.7\nKs 0.35 0.35 0.35\nNs 32\nNi 1.5\nillum 2\nd 1\n")
    m = mats["a"]
    assert m.JS_NAME == "PhongMaterial" and m.smoothness == 32      # makeMaterial lacks `return` (src/objloader.js:16-19)
    assert m.diffusivity._color.tolist() == Vec.of(0.3, 0.8, 0.7).tolist()


def test_mtl_texture_maps():
    """parseMtlFile's map_K* handling (src/objloader.js:61-70,100-105) and makeMaterialColor (:1-7)."""
    from jsraytracer_b200.materials import ImageData, TextureMaterialColor
    from jsraytracer_b200.objloader import mtl_texture_names
    text = "newmtl a\nmap_Kd a.jpg\nKs 0.1 0.1 0.1\nNs 1000\nnewmtl b\nmap_Kd -o 1 1 b.png\nKd 0.5 1 0.25\n"
    assert mtl_texture_names(text) == ["a.jpg", "b.png"]                       # the last token is the file name
    tex = {n: TextureMaterialColor(ImageData.from_array(np.full((2, 2, 4), 255, np.uint8))) for n in ("a.jpg", "b.png")}
    mats = parse_mtl_text(text, tex)
    assert mats["a"].diffusivity is tex["a.jpg"]                                # map without K: the texture itself
    assert mats["a"].ambient._color.tolist() == [0, 0, 0]                       # neither Ka nor map_Ka: default black
    b = mats["b"].diffusivity
    assert b.JS_NAME == "ScaledMaterialColor" and b._mc is tex["b.png"] and b._scale == [0.5, 1.0, 0.25]
    with pytest.raises(ValueError, match="Unknown texture"):
        parse_mtl_text(text, {"a.jpg": tex["a.jpg"]})


def test_bottle_fixture_has_decoded_textures():
    """tests/bottle: potion_bottle.mtl names two JPEGs; the fixture carries them decoded to RGBA."""
    fix = ParsedObj.load(os.path.join(scenes.DATA_DIR, "Potion_bottle.npz"))
    assert sorted(fix.textures) == ["bottle_health.jpg", "bottle_mana.jpg"]
    assert all(a.shape == (256, 256, 4) and a.dtype == np.uint8 and (a[..., 3] == 255).all() for a in fix.textures.values())
    test = scenes.configure("bottle", width=8, height=8)
    bvh = test["renderer"].world.objects[1]
    mats = {id(p.material): p.material for p in bvh.objects}
    assert len(bvh.objects) == 198 and len(mats) == 1                             # one usemtl: bottle_blue
    m = next(iter(mats.values()))
    assert m.diffusivity.JS_NAME == "TextureMaterialColor" and m.diffusivity.width == 256 and m.smoothness == 1000
    ref = os.path.join(REF_ASSETS, "potion_bottle", "bottle_mana.jpg")
    if os.path.exists(ref):
        sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
        from import_reference_assets import decode_rgba
        assert np.array_equal(decode_rgba(ref), fix.textures["bottle_mana.jpg"])


@pytest.mark.parametrize("name", ["bunny2", "dragon", "hollow_tetrahedron", "star", "teapot"])
def test_fixtures_match_reference_assets(name):
    path = os.path.join(REF_ASSETS, name + ".obj")
    if not os.path.exists(path):
        pytest.skip("reference assets not mounted")
    fresh = parse_obj_text(open(path, encoding="utf8").read())
    fix = ParsedObj.load(os.path.join(scenes.DATA_DIR, name + ".npz"))
    for f in ("positions", "texcoords", "normals", "faces", "face_material"):
        assert np.array_equal(getattr(fresh, f), getattr(fix, f)), f


def test_fixture_counts():
    b = ParsedObj.load(os.path.join(scenes.DATA_DIR, "bunny2.npz"))
    assert len(b.positions) == 2503 and len(b.normals) == 2503 and len(b.faces) == 4968
    d = ParsedObj.load(os.path.join(scenes.DATA_DIR, "dragon.npz"))
    assert len(d.positions) == 50000 and len(d.faces) == 100000 and len(d.normals) == 0


EDGE_OBJ = """# comment
   \t# indented comment
mtllib a.mtl
o thing
v 1 2 3
v 1.5e0 -2.25 +.5 0.5
v 1e-3x 2abc 3. 7
v Infinity -Infinity 0
vt 0.25
vt 0.5 0.75
vt 0.1 nan 0
vt 1 2 3
vn 0 0 1
vn 0.6 0.8 0 junk
usemtl red
f 1 2 3
f 1/1 2/2 3/3 4/4
usemtl blue
f 1//1 2//2 3//1
f 1/2/1 2/1/2 3/3/1 4/2/2 1/1/1
usemtl red
f -1 -2 -3
f x1/ y2// z3/4/
s off
g grp
vp 1 2
\r
"""


def _same_parse(a, b):
    for f in ("positions", "texcoords", "normals", "faces", "face_material"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape and x.dtype == y.dtype, f
        assert np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y), f
    assert a.material_names == b.material_names and a.mtllibs == b.mtllibs


def test_native_obj_reader_matches_the_mirror_on_edge_cases():
    """jsrt_obj_parse (csrc/obj_parse.cpp) against the line-by-line restatement of parseObjFile: Number.parseFloat prefixes,
    the corner regex's first-digit-run rule (negative indices lose their sign, as in the reference), `t[2] || 0`, fans."""
    from jsraytracer_b200.objloader import parse_obj_text_native
    a, b = parse_obj_text(EDGE_OBJ), parse_obj_text_native(EDGE_OBJ)
    _same_parse(a, b)
    assert b.material_names == ["red", "blue"] and b.mtllibs == ["a.mtl"]
    assert b.positions[2].tolist() == [np.float32(1e-3), 2.0, 3.0, 7.0] and np.isinf(b.positions[3][:2]).all()
    assert b.texcoords[2].tolist() == [np.float32(0.1), 0.0, 0.0]
    assert len(b.faces) == 1 + 2 + 1 + 3 + 1 + 1 and b.face_material.tolist() == [0, 0, 0, 1, 1, 1, 1, 0, 0]
    assert b.faces[-2].tolist() == [[0, -1, -1], [1, -1, -1], [2, -1, -1]]         # "-1" -> 1
    assert b.faces[-1].tolist() == [[0, -1, -1], [1, -1, -1], [2, 3, -1]]          # "z3/4/": the dangling slash matches nothing
    with pytest.raises(ValueError, match="Error while attempting to parse obj file on line"):
        parse_obj_text_native("v 0 0 0\nbogus 1 2 3\n")
    e = parse_obj_text_native("")
    assert len(e.positions) == 0 and len(e.faces) == 0 and e.faces.shape == (0, 3, 3)


@pytest.mark.parametrize("name", ["bunny2", "dragon", "Tie_Fighter", "x_wing_fighter", "cat", "high-poly-teapot", "heart",
                                  "potion_bottle/Potion_bottle", "hammer/neuro_hammer_obj", "Aztec_Templ_2", "hollow_dodecahedron"])
def test_native_obj_reader_matches_the_mirror_on_reference_assets(name):
    from jsraytracer_b200.objloader import parse_obj_text_native
    path = os.path.join(REF_ASSETS, name + ".obj")
    if not os.path.exists(path):
        pytest.skip("reference assets not mounted")
    text = open(path, encoding="utf8", errors="replace").read()
    try:
        a = parse_obj_text(text)
    except Exception as exc:                      # a file the reference's parser rejects: the native reader must reject it too
        with pytest.raises(ValueError):
            parse_obj_text_native(text)
        return
    _same_parse(a, parse_obj_text_native(text))


def test_native_obj_reader_on_the_committed_fixtures():
    """Round trip without the reference tree: OBJ text regenerated from a fixture parses back to the fixture."""
    from jsraytracer_b200.objloader import parse_obj_text_native
    fix = ParsedObj.load(os.path.join(scenes.DATA_DIR, "teapot.npz"))
    lines = ["v %r %r %r" % tuple(float(x) for x in p[:3]) for p in fix.positions]
    lines += ["vt %r %r %r" % tuple(float(x) for x in p) for p in fix.texcoords]
    lines += ["vn %r %r %r" % tuple(float(x) for x in p[:3]) for p in fix.normals]
    for f in fix.faces:
        lines.append("f " + " ".join("/".join("" if c < 0 else str(c + 1) for c in corner).rstrip("/") for corner in f.tolist()))
    got = parse_obj_text_native("\n".join(lines))
    for k in ("positions", "texcoords", "normals", "faces"):
        assert np.array_equal(getattr(got, k), getattr(fix, k)), k
