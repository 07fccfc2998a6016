"""On the GPU box: the CUDA path against the reference's own SimpleRenderer images (tests/golden/refjs_*.npz), one line
per deterministic fixture — the numbers behind the allowances of tests/test_gpu_refjs.py."""
import glob
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import lib  # noqa: E402

for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refjs_*.npz"))):
    z = np.load(p)
    meta = json.loads(str(z["meta"]))
    if not ("simple_mean" in z.files or (meta["renderer"] == "SimpleRenderer" and int(z["draws"].max()) == 0)):
        continue
    if meta["name"] == "SDF_RecursiveUnionTest":
        continue
    ref = z["simple_mean"] if "simple_mean" in z.files else z["mean"]
    ref8 = z["simple_rgba8"] if "simple_rgba8" in z.files else z["rgba8"]
    sc = lib.Scene(zlib.decompress(z["json"].tobytes()), lib.FORMAT_JSON, device=0)
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    acc, _ = sc.read_accum()
    g, r = np.clip(acc[..., :3], 0, 1), np.clip(ref, 0, 1)
    d = np.abs(g.astype(np.float64) - r).max(-1)
    d8 = np.abs(sc.resolve_rgba8().astype(int) - ref8.astype(int)).max(-1)
    mse = float(np.mean((g.astype(np.float64) - r) ** 2))
    bad = np.argwhere(d > 2e-3)
    print(json.dumps(dict(name=meta["name"], pixels=int(d.size), over_2e3=int((d > 2e-3).sum()), over_1e4=int((d > 1e-4).sum()),
                          max=float(d.max()), median=float(np.median(d)), psnr=round(10 * np.log10(1 / max(mse, 1e-30)), 1),
                          grey_over_1=int((d8 > 1).sum()), where=[(int(y), int(x), round(float(d[y, x]), 4)) for y, x in bad[:8]])))
