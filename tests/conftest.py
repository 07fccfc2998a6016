import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the C-ABI library and the oracle once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


_SCENE_CACHE = {}


def scene_blobs(name, **kw):
    """(json_text, msgpack_bytes) of `Serializer(configure(name, **kw))`, cached per session."""
    key = (name, tuple(sorted(kw.items())))
    if key not in _SCENE_CACHE:
        from jsraytracer_b200 import scenes
        from jsraytracer_b200.serializer import Serializer
        ser = Serializer(scenes.configure(name, **kw))
        _SCENE_CACHE[key] = (ser.to_json(), ser.to_msgpack())
    return _SCENE_CACHE[key]


@pytest.fixture(scope="session")
def blobs():
    return scene_blobs


def psnr(a, b):
    import numpy as np
    mse = float(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))
    return 10 * np.log10(1.0 / max(mse, 1e-30))
