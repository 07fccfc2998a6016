"""OBJ loading restatement (src/objloader.js:144-247) and the committed mesh fixtures."""
import os

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
