"""The C-ABI library loads, exports every symbol include/jsrt.h declares, and fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

from jsraytracer_b200 import lib

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jsrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jsrt_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(L, s), "libjsrt.so does not export " + s
    assert set(syms) == set(lib.EXPORTS)


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "jsrt.h")).read()
    assert "torch" not in text and "at::" not in text and "std::" not in text


def test_bad_blob_reports_error():
    with pytest.raises(lib.JsrtError) as e:
        lib.Scene(b"{not json", lib.FORMAT_JSON, device=None)
    assert "jsrt" in str(e.value)
    with pytest.raises(lib.JsrtError):
        lib.Scene(b'{"_t":["Object",0],"_v":{"width":4,"height":4}}', lib.FORMAT_JSON, device=None)   # no renderer
    with pytest.raises(lib.JsrtError):
        lib.Scene(b"\xc1", lib.FORMAT_MSGPACK, device=None)


def test_host_handle_refuses_to_render(blobs):
    js, mp = blobs("BoxBall", width=16, height=16)
    sc = lib.Scene(mp, lib.FORMAT_MSGPACK, device=None)
    assert sc.info["n_prims"] == 3 and sc.info["n_lights"] == 1 and sc.info["max_depth"] == 4
    with pytest.raises(lib.JsrtError) as e:
        sc.render(0, 1)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(lib.JsrtError):
        sc.primary_hits()


def test_scene_create_without_gpu_fails_loudly(blobs):
    if lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    js, mp = blobs("BoxBall", width=16, height=16)
    with pytest.raises(lib.JsrtError) as e:
        lib.Scene(mp, lib.FORMAT_MSGPACK, device=0)
    assert "no CUDA device" in str(e.value)


def test_unsupported_feature_is_an_error_not_a_guess():
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.serializer import Serializer
    from jsraytracer_b200.geometry import JSObject

    class TextureMaterialColor(JSObject):
        JS_NAME = "TextureMaterialColor"

        def __init__(self):
            self.width = 1
    test = scenes.configure("BoxBall", width=8, height=8)
    test["renderer"].world.objects[1].material.ambient = TextureMaterialColor()
    with pytest.raises(lib.JsrtError) as e:
        lib.Scene(Serializer(test).to_msgpack(), lib.FORMAT_MSGPACK, device=None)
    assert "TextureMaterialColor" in str(e.value)


def test_degenerate_scenes_flatten_and_render_in_the_oracle():
    """Edge cases of the boundary: an empty world (every ray returns bg_color, src/world.js:35), a world without lights
    (ambient term only, src/materials.js:240-259), a 1x1 image, header values out of range."""
    import json
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.cameras import PerspectiveCamera
    from jsraytracer_b200.jsmath import Mat4, Vec
    from jsraytracer_b200.renderers import SimpleRenderer
    from jsraytracer_b200.serializer import Serializer
    from jsraytracer_b200.world import World
    from oracle.oracle import OracleScene
    cam = PerspectiveCamera(0.7, 1, Mat4.identity())
    empty = {"renderer": SimpleRenderer(World([], [], Vec.of(0.25, 0.5, 0.75)), cam, 3), "width": 5, "height": 3}
    ser = Serializer(empty)
    info = lib.Scene(ser.to_msgpack(), lib.FORMAT_MSGPACK, device=None).info
    assert (info["n_prims"], info["n_lights"], info["n_nodes"], info["light_samples"]) == (0, 0, 0, 0)
    img, cnt = OracleScene(ser.to_json()).render(1, seed=1, jitter=False)
    assert img.shape == (3, 5, 3) and np.allclose(img, [0.25, 0.5, 0.75]) and cnt["rays_primary"] == 15 and cnt["rays_shadow"] == 0
    ids, t, _ = OracleScene(ser.to_json()).primary_hits()
    assert (ids == -1).all()

    one = scenes.configure("BoxBall", width=1, height=1)
    one["renderer"].world.lights = []                                   # no lights: ambient + reflections only
    ser = Serializer(one)
    info = lib.Scene(ser.to_json(), lib.FORMAT_JSON, device=None).info
    assert (info["width"], info["height"], info["n_lights"], info["light_samples"]) == (1, 1, 0, 0)
    img, cnt = OracleScene(ser.to_json()).render(2, seed=1)
    assert img.shape == (1, 1, 3) and np.isfinite(img).all() and cnt["rays_shadow"] == 0

    doc = json.loads(Serializer(scenes.configure("BoxBall", width=8, height=8)).to_json())
    for key, bad in (("width", 0), ("width", 1e12), ("height", -3), ("height", 70000)):
        d = json.loads(json.dumps(doc))
        d["_v"][key] = bad
        with pytest.raises(lib.JsrtError, match="width|height"):
            lib.Scene(json.dumps(d), lib.FORMAT_JSON, device=None)


def test_product_never_touches_the_oracle():
    """oracle/ (the C++ restatement, the JavaScript interpreter, the reference runner) is test infrastructure: nothing under
    jsraytracer_b200/ or js/ may import, load or name it; bench.py and __graft_entry__ use it only in the legs the contract
    allows (cpu_baseline / --impl reference / smoke's checker)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    offenders = []
    for base in ("jsraytracer_b200", "js", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if not f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".js", ".cc")):
                    continue
                text = open(os.path.join(dirpath, f), encoding="utf8", errors="replace").read()
                if re.search(r"(from|import)\s+oracle\b|liboracle|oracle/|oracle\.(oracle|refjs|jsvm)", text):
                    # comments that cite the oracle as the checker are fine; code that reaches it is not
                    code = "\n".join(l for l in text.splitlines() if not l.strip().startswith(("#", "//", "*", '"""')))
                    if re.search(r"(from|import)\s+oracle\b|liboracle|dlopen\([^)]*oracle|CDLL\([^)]*oracle", code):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    # bench.py: the oracle is reached from exactly two functions — the cpu_baseline leg and the --impl reference arm
    import ast
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    where = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for n in ast.walk(fn):
            if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle":
                where.add(fn.name)
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]
    assert where == {"cpu_reference_run", "run_reference"} and not top
