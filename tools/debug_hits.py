#!/usr/bin/env python
"""Parity debugging helper: primary-hit mismatches between GPU and oracle for one scene."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import lib, scenes  # noqa: E402
from jsraytracer_b200.serializer import Serializer  # noqa: E402
from oracle.oracle import OracleScene  # noqa: E402

name = sys.argv[1]
W, H = int(sys.argv[2]), int(sys.argv[3])
ser = Serializer(scenes.configure(name, width=W, height=H, aspect=W / H))
sc = lib.Scene(ser.to_msgpack(), lib.FORMAT_MSGPACK, device=0)
orc = OracleScene(ser.to_json())
ids, t = sc.primary_hits()
oids, ot, _ = orc.primary_hits()
bad = ids != oids
print("mismatch", int(bad.sum()), "of", ids.size)
ys, xs = np.where(bad)
kinds = {"gpu_miss": 0, "orc_miss": 0, "both_hit_same_t": 0, "both_hit_diff_t": 0}
for y, x in zip(ys, xs):
    if ids[y, x] < 0:
        kinds["gpu_miss"] += 1
    elif oids[y, x] < 0:
        kinds["orc_miss"] += 1
    elif abs(t[y, x] - ot[y, x]) <= 1e-5 * abs(ot[y, x]):
        kinds["both_hit_same_t"] += 1
    else:
        kinds["both_hit_diff_t"] += 1
print(kinds)
for y, x in list(zip(ys, xs))[:25]:
    print(y, x, "gpu", ids[y, x], float(t[y, x]), "orc", oids[y, x], float(ot[y, x]))
np.savez(os.path.join(ROOT, "gpurun_out", "hits_%s.npz" % name), ids=ids, t=t, oids=oids, ot=ot)
