#!/usr/bin/env python
"""Derive the mesh fixtures under jsraytracer_b200/scenes/data/ from the
reference's OBJ assets (run in the build container, where /root/reference is
mounted; the GPU box has only the derived arrays).

    python tools/import_reference_assets.py [/root/reference/assets]

Each .npz holds what `parseObjFile` (src/objloader.js:149-238) extracts from
the text: f32 positions / texcoords / normals and fan-triangulated index
triples.  tests/test_objloader.py re-parses the originals when they are
present and checks the fixtures still match.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from jsraytracer_b200.objloader import parse_obj_text, mtl_texture_names  # noqa: E402


def decode_rgba(path):
    """Stand-in for the browser's createImageBitmap + getImageData (src/objloader.js:34-41, src/materials.js:91-96):
    the image decoded to (H, W, 4) uint8 RGBA, alpha 255 where the file has none."""
    import numpy as np
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8)

ASSETS = ["bunny2", "dragon", "hollow_tetrahedron", "star", "tetrahedron", "cube", "teapot", "Tie_Fighter",
          "x_wing_fighter", "cat", "diamond", "heart", "high-poly-teapot", "potion_bottle/Potion_bottle"]


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets"
    out = os.path.join(os.path.dirname(__file__), "..", "jsraytracer_b200", "scenes", "data")
    os.makedirs(out, exist_ok=True)
    for name in ASSETS:
        with open(os.path.join(src, name + ".obj"), encoding="utf8") as fh:
            parsed = parse_obj_text(fh.read())
        prefix = os.path.dirname(os.path.join(src, name))
        for lib in parsed.mtllibs:          # MTL files are a few hundred bytes: keep their text with the mesh
            with open(os.path.join(prefix, lib), encoding="utf8") as fh:
                parsed.mtl_texts.append(fh.read())
            for tname in mtl_texture_names(parsed.mtl_texts[-1]):   # and the decoded texture maps they name
                parsed.textures[tname] = decode_rgba(os.path.join(prefix, tname))
        path = os.path.join(out, os.path.basename(name) + ".npz")
        parsed.save(path)
        print("%-20s verts=%d normals=%d uvs=%d tris=%d mats=%s -> %d bytes" % (
            name, len(parsed.positions), len(parsed.normals), len(parsed.texcoords), len(parsed.faces),
            parsed.material_names, os.path.getsize(path)))


if __name__ == "__main__":
    main()
