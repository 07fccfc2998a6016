#!/usr/bin/env python
"""Generate tests/golden/*.npz from the CPU restatement oracle.

The reference ships no machine-checked fixtures for this path and cannot be
executed in this image, so these vectors come from the restatement, not from the reference
running: they pin the oracle against regressions and give the GPU tests a committed target, but
they do not pin the oracle to the reference ("parity unpinned").

    python tools/make_golden.py [key ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import scenes  # noqa: E402
from jsraytracer_b200.serializer import Serializer  # noqa: E402
from oracle.oracle import OracleScene  # noqa: E402

CASES = {
    # name: (scene, kwargs, whitted passes (no jitter), stochastic passes (seed 1))
    "BoxBall_96": ("BoxBall", dict(width=96, height=96), 1, 2),
    "ASimpleScene_96": ("ASimpleScene", dict(width=96, height=96), 1, 0),
    "cornell_box_path_64": ("cornell_box_path", dict(width=64, height=64), 0, 2),
    "bunny_path_128x72": ("bunny_path", dict(width=128, height=72, aspect=16 / 9), 1, 2),
    "AHollowTetrahedron_96": ("AHollowTetrahedron", dict(width=96, height=96), 1, 0),
    "SDF_Menger_64": ("SDF_Menger", dict(width=64, height=64), 1, 0),
    "SDF_Sierpinski_64": ("SDF_Sierpinski", dict(width=64, height=64), 1, 0),
    "BoxBall_DOF_64": ("BoxBall_DOF", dict(width=64, height=64), 0, 2),
    # session 4: textured mesh through the MTL loader, coincident faces (f64 tie-break), instanced / multiple aggregates
    "bottle_96": ("bottle", dict(width=96, height=96), 1, 0),
    "x_wing_128": ("x_wing", dict(width=128, height=128), 0, 1),
    "starwars_128x72": ("starwars", dict(width=128, height=72, aspect=16 / 9), 0, 1),
    "dragon_grid_96x54": ("dragon_grid", dict(width=96, height=54, aspect=16 / 9, n=2), 1, 0),
    "textured_96": ("textured", dict(width=96, height=96), 1, 0),
}


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    only = sys.argv[1:]                 # keys to (re)generate; default: all
    for key, (name, kw, wp, sp) in CASES.items():
        if only and key not in only:
            continue
        orc = OracleScene(Serializer(scenes.configure(name, **kw)).to_json())
        ids, t, _ = orc.primary_hits()
        data = {"prim_id": ids.astype(np.int32), "t": t.astype(np.float64)}
        if wp:
            data["whitted"] = orc.render(wp, seed=1, jitter=False)[0]
        if sp:
            acc, cnt = orc.render(sp, seed=1)
            data["stochastic"] = acc
            data["stochastic_rays"] = np.array([cnt["rays_primary"], cnt["rays_secondary"], cnt["rays_shadow"]], dtype=np.int64)
        np.savez_compressed(os.path.join(out, key + ".npz"), **data)
        print(key, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
