#!/usr/bin/env python
"""Parity debugging helper: render a scene at recursion depths 1..D on GPU and oracle (no jitter, 1 pass)
and dump both images per depth under gpurun_out/ so differing pixels can be localised to a bounce."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import lib, scenes  # noqa: E402
from jsraytracer_b200.serializer import Serializer  # noqa: E402
from oracle.oracle import OracleScene  # noqa: E402

name = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
maxd = int(sys.argv[3]) if len(sys.argv) > 3 else 4
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
for depth in range(1, maxd + 1):
    ser = Serializer(scenes.configure(name, width=size, height=size, depth=depth))
    sc = lib.Scene(ser.to_msgpack(), lib.FORMAT_MSGPACK, device=0)
    orc = OracleScene(ser.to_json())
    sc.render(0, 1, seed=1, flags=lib.FLAG_NO_JITTER)
    g = sc.read_accum()[0][..., :3]
    o = orc.render(1, seed=1, jitter=False)[0]
    d = np.abs(g - o).max(axis=-1)
    print(name, "depth", depth, "pixels>1e-3:", int((d > 1e-3).sum()), "max", float(d.max()))
    np.save(os.path.join(out, "%s_d%d_gpu.npy" % (name, depth)), g.astype(np.float32))
    np.save(os.path.join(out, "%s_d%d_oracle.npy" % (name, depth)), o.astype(np.float32))
