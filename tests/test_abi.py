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
