#!/usr/bin/env python
"""Per-kernel device time and work counters for one scene (perf iteration helper)."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import lib, scenes  # noqa: E402
from jsraytracer_b200.serializer import Serializer  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "bunny_path"
    W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
    passes = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    kw = {}
    for a in sys.argv[5:]:                      # extra scene keywords, e.g. n=3
        k, v = a.split("=")
        kw[k] = (9.2, 0.2) if k == "dof" else int(v)     # dof=1: the BoxBall_DOF camera (BASELINE configs[3])
    ser = Serializer(scenes.configure(name, width=W, height=H, aspect=W / H, **kw))
    sc = lib.Scene(ser.to_msgpack(), lib.FORMAT_MSGPACK, device=0)
    sc.render(0, passes, seed=1)
    sc.synchronize()
    sc.stats_reset()
    sc.set_profiling(True)
    t0 = time.perf_counter()
    sc.render(passes, passes, seed=1)
    sc.synchronize()
    dt = time.perf_counter() - t0
    st = sc.stats()
    sc.set_profiling(False)
    sc.stats_reset()
    sc.render(1 << 20, 1, seed=1, flags=lib.FLAG_COUNT_WORK)
    cw = sc.stats()
    rays = st["rays"]
    out = {"scene": name, "size": [W, H], "passes": passes, "wall_ms": dt * 1e3, "Mrays_s": rays / dt / 1e6,
           "ms": {k: round(st["ms_" + k], 3) for k in ("generate", "extend", "shade", "shadow")},
           "ms_parts": {k: round(st["ms_" + k], 3) for k in ("extend_prims", "extend_bvh", "extend_sdf", "shadow_prims", "shadow_bvh", "shadow_sdf")},
           "rays": {k: st["rays_" + k] for k in ("primary", "secondary", "shadow")},
           "per_ray": {cls: {"nodes": cw["bvh_nodes"][i] / max(1, cw["rays_" + cls]), "leaf_prims": cw["bvh_prims"][i] / max(1, cw["rays_" + cls]),
                             "top_prims": cw["top_prims"][i] / max(1, cw["rays_" + cls]), "sdf_evals": cw["sdf_evals"][i] / max(1, cw["rays_" + cls])}
                       for i, cls in enumerate(("primary", "secondary", "shadow"))},
           "info": {k: sc.info[k] for k in ("n_nodes", "n_tris", "batch_samples", "scene_bytes", "queue_bytes")}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
