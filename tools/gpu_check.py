#!/usr/bin/env python
"""First-contact GPU check: parity numbers against the oracle for a few scenes,
plus timings.  Writes images and a log under gpurun_out/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from jsraytracer_b200 import lib, scenes  # noqa: E402
from jsraytracer_b200.serializer import Serializer  # noqa: E402
from oracle.oracle import OracleScene  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10 * np.log10(1.0 / max(mse, 1e-30))


def check(name, passes, jitter, **kw):
    t0 = time.time()
    test = scenes.configure(name, **kw)
    ser = Serializer(test)
    blob, js = ser.to_msgpack(), ser.to_json()
    t1 = time.time()
    sc = lib.Scene(blob, lib.FORMAT_MSGPACK, device=0)
    orc = OracleScene(js)
    t2 = time.time()
    res = {"scene": name, "kw": {k: v for k, v in kw.items()}, "build_s": round(t1 - t0, 2), "load_s": round(t2 - t1, 2), "info": sc.info}
    ids, t = sc.primary_hits()
    oids, ot, cnt = orc.primary_hits()
    same = ids == oids
    res["hit_id_agreement"] = float(same.mean())
    both = same & (oids >= 0)
    rel = np.abs(t[both].astype(np.float64) - ot[both]) / np.maximum(np.abs(ot[both]), 1e-12)
    res["t_rel_max"] = float(rel.max()) if rel.size else 0.0
    res["t_rel_p9999"] = float(np.quantile(rel, 0.9999)) if rel.size else 0.0
    sc.stats_reset()
    tg = time.time()
    sc.render(0, passes, seed=1, flags=0 if jitter else lib.FLAG_NO_JITTER)
    sc.synchronize()
    res["gpu_render_s"] = round(time.time() - tg, 4)
    st = sc.stats()
    res["gpu_stats"] = st
    acc, np_ = sc.read_accum()
    tg = time.time()
    oacc, ocnt = orc.render(passes, seed=1, jitter=jitter)
    res["oracle_render_s"] = round(time.time() - tg, 3)
    res["oracle_counts"] = {k: v for k, v in ocnt.items() if v}
    gimg, oimg = acc[..., :3] / passes, oacc / passes
    res["psnr_f32"] = psnr(np.clip(gimg, 0, 1), np.clip(oimg, 0, 1))
    res["max_abs_diff"] = float(np.abs(gimg - oimg).max())
    res["frac_pixels_diff_gt_1e-3"] = float((np.abs(gimg - oimg).max(axis=-1) > 1e-3).mean())
    res["ray_count_match"] = [st["rays_primary"] == ocnt["rays_primary"], st["rays_secondary"] == ocnt["rays_secondary"], st["rays_shadow"] == ocnt["rays_shadow"]]
    if os.environ.get("JSRT_DUMP"):
        np.save(os.path.join(OUT, "%s_gpu.npy" % name), gimg.astype(np.float32))
        np.save(os.path.join(OUT, "%s_oracle.npy" % name), oimg.astype(np.float32))
    try:
        from PIL import Image
        Image.fromarray(sc.resolve_rgba8().copy()).save(os.path.join(OUT, "%s_gpu.png" % name))
        Image.fromarray((np.clip(oimg, 0, 1) * 255 + 0.5).astype(np.uint8)).save(os.path.join(OUT, "%s_oracle.png" % name))
    except Exception as e:  # pragma: no cover
        res["png_error"] = str(e)
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    print("devices", lib.device_count())
    out = []
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            out.append(check(name, 1, False, width=256, height=256))
        sys.exit(0)
    out.append(check("BoxBall", 1, False, width=512, height=512))
    out.append(check("BoxBall", 4, True, width=256, height=256))
    out.append(check("bunny", 1, False, width=480, height=270, aspect=16 / 9))
    out.append(check("bunny_path", 2, True, width=480, height=270, aspect=16 / 9))
    out.append(check("cornell_box_path", 2, True, width=256, height=256))
    out.append(check("ASimpleScene", 1, False, width=256, height=256))
    out.append(check("SDF_Sierpinski", 1, False, width=192, height=192))
    out.append(check("SDF_Menger", 1, False, width=192, height=192))
    json.dump(out, open(os.path.join(OUT, "gpu_check.json"), "w"), indent=1)
