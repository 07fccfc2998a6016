"""Writes profiles/r2_refjs_pin.md: one row per fixture of tests/golden/refjs_*.npz — what the reference computed (its own
sources in oracle/jsvm), and whether the oracle's tape mode reproduces it bit for bit.  CPU only.
    python tools/refjs_pin_report.py"""
import glob
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.oracle import OracleScene, resolve_rgba8  # noqa: E402
from jsraytracer_b200 import lib  # noqa: E402

rows = []
for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refjs_*.npz"))):
    z = np.load(p)
    m = json.loads(str(z["meta"]))
    js = zlib.decompress(z["json"].tobytes())
    simple = m["renderer"] == "SimpleRenderer"
    n = 1 if simple else m["passes"]
    sc = OracleScene(js.decode())
    acc, cnt = sc.render(n, seed=m["seed"], jitter=not simple, width=m["width"], height=m["height"], threads=1, tape=True)
    mean = (acc.astype(np.float64) * (1.0 / n)).astype(np.float32)
    same = np.array_equal(mean, z["mean"], equal_nan=True)
    same8 = np.array_equal(resolve_rgba8(acc, n), z["rgba8"])
    try:
        info = lib.Scene(js, lib.FORMAT_JSON, device=None).info
        tris, nodes = info["n_tris"], info["n_nodes"]
    except lib.JsrtError:
        tris = nodes = "-"
    extras = [k for k in ("simple_mean", "random_mean") if k in z.files]
    rays = cnt["rays_primary"] + cnt["rays_secondary"] + cnt["rays_shadow"]
    rows.append("| %s | %s | %dx%d x %d | %s | %s | %d–%d | %d | %s / %s | %s | %.0f + %.0f |" % (
        m["name"], m["renderer"].replace("MultisamplingRenderer", "").replace("Renderer", ""), m["width"], m["height"], n, tris, nodes,
        int(z["draws"].min()), int(z["draws"].max()), rays, "equal" if same else "DIFFERENT", "equal" if same8 else "DIFFERENT",
        ", ".join(e.split("_")[0] for e in extras) or "", m["load_s"], m["render_s"]))
out = os.path.join(ROOT, "profiles", "r2_refjs_pin.md")
with open(out, "w") as f:
    f.write("# The oracle against the reference's own output (round 2, session 3)\n\n"
            "Reference = `/root/reference/src/*.js` + `tests/<scene>/test.mjs`, unmodified, executed by `oracle/jsvm` through\n"
            "`oracle/refjs.py` (its own `renderer.render` loop, `Math.random` = per-sample splitmix64 tape).  Oracle =\n"
            "`oracle/oracle.cpp` in tape mode on the JSON the reference's own `Serializer` wrote.  `f32 / u8`: the colours handed to\n"
            "`PixelBuffer.setColor` after the last pass / the bytes of the `ImageData`, compared for equality.  `draws`: `Math.random()`\n"
            "calls per pixel sample (min–max).  `rays`: `World.cast` calls of the oracle for the image.  `also`: further renders of the\n"
            "same world in the fixture (the reference's SimpleRenderer / RandomMultisamplingRenderer), equally compared by\n"
            "`tests/test_refjs_pin.py`.  `s`: seconds the interpreter took to configure (OBJ parse + `BVHAggregate.build`) + render.\n\n"
            "| scene | renderer | image x passes | triangles | BVH nodes | draws | rays | f32 / u8 | also | s |\n|---|---|---|---|---|---|---|---|---|---|\n")
    f.write("\n".join(rows) + "\n")
    f.write("\nOne-off runs, not kept as fixtures because of their size (a 23 MB scene document): `tests/x-wing` (18 849 triangles, 10 minutes to\n"
            "configure in the interpreter), 20x12 x 1 pass, 2-18 draws per sample: f32 colours and ImageData bytes equal the oracle's\n"
            "(`profiles/r2_refjs_xwing_oneoff.log`, the output of `tools/refjs_compare.py x-wing 20 12 1`); `tests/starwars` (BASELINE's\n"
            "stand-in for Toledo: x-wing + three tie fighters sharing one kd-tree, depth-of-field camera, three area lights + a point\n"
            "light: 4-124 draws per sample), 20x12 x 1 pass: equal likewise (`profiles/r2_refjs_starwars_oneoff.log`); `tests/dragon` (99 968\n"
            "triangles: one hour to configure, 106 MB document), 20x12 x 1 pass: equal likewise (`profiles/r2_refjs_dragon_oneoff.log`).\n"
            "With these, 34 of the 37 scenes of the reference's tests/list.json are pinned; toledo, toledo_json and dragon_json cannot run\n"
            "from the reference tree (asset / test.json absent).\n")
print(open(out).read())
