"""TEST INFRASTRUCTURE — writes tests/golden/refjs_<scene>.npz: outputs of THE REFERENCE ITSELF (its unmodified
src/*.js and tests/<scene>/test.mjs executed by oracle/jsvm through oracle/refjs.py), which the oracle is pinned to.

    python -m oracle.refjs_golden                # every scene of the table below, in parallel
    python -m oracle.refjs_golden BoxBall bunny  # some of them

Runs only where /root/reference exists (this container).  Each fixture holds
    json    the scene in the reference's own wire format (`new Serializer(test)`, zlib-compressed UTF-8)
    mean    (H, W, 3) f32: the colour the reference's render loop hands to PixelBuffer.setColor after the last pass
    rgba8   (H, W, 4) u8:  the ImageData the reference filled
    draws   (H, W) i32:    Math.random() calls of each pixel's last sample
    random_mean / random_rgba8 / random_spp   (two scenes) the same world through the reference's RandomMultisamplingRenderer
    simple_mean / simple_rgba8   (scenes whose only random numbers are the pixel jitter) the same world and camera through
            the reference's un-jittered SimpleRenderer: a deterministic image the CUDA path is compared with directly
    meta    name, width, height, passes, seed, renderer class, depth, seconds, sha256 of the sources that ran
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (scene of /root/reference/tests, width, height, passes).  Sizes are small because the interpreter is ~1000x slower than
# V8; widths differ from heights so that a transposed image cannot pass; heights are odd where possible so that no image
# row has camera rays that are exactly horizontal (they meet the ground planes at the horizon: ill-conditioned for any
# implementation that is compared with a tolerance, like the FP32 device path).  Not covered: dragon / dragon_json / x-wing /
# starwars (19 000 - 100 000 triangles: an hour to many hours in the interpreter), toledo* (asset missing from the
# reference tree).
TABLE = [
    ("BoxBall", 48, 31, 3), ("BoxBall_DOF", 48, 31, 3), ("BoxBall_path", 32, 21, 2), ("ASimpleScene", 40, 27, 2),
    ("Aggregates", 40, 27, 2), ("AHollowTetrahedron", 48, 31, 1), ("AMultipleBVH", 48, 31, 1),
    ("refraction", 32, 21, 2), ("refraction_simple", 40, 27, 2), ("refraction_path", 24, 15, 2),
    ("cornell_box", 28, 19, 2), ("cornell_box_emissive", 28, 19, 2), ("cornell_box_path", 24, 16, 2),
    ("spheres010", 40, 27, 1), ("spheres050", 24, 15, 1), ("spheres100", 20, 13, 1),
    ("SDF_Simple", 40, 27, 2), ("SDF_BoxBall", 32, 21, 2), ("SDF_Combinations", 32, 21, 1), ("SDF_Menger", 20, 12, 2),
    ("SDF_Sierpinski", 24, 15, 2), ("SDF_SphereRepetition", 24, 15, 1), ("SDF_RecursiveUnionTest", 16, 10, 1),
    ("diamond", 40, 27, 1), ("heart", 48, 31, 1), ("cat", 48, 31, 1), ("utah_teapot", 32, 21, 1),
    ("bunny", 48, 31, 1), ("bunny_path", 32, 21, 3), ("tie_fighter", 32, 21, 2), ("bottle", 48, 31, 1),
    # not the reference's: tests/golden/extra_scenes/materials/test.mjs, the reference's classes its own scenes leave unused
    ("extra_materials", 42, 27, 3), ("extra_materials_whitted", 48, 31, 2),
    # ... and aggregates nested in ways its API allows and its scenes never do (the twin of scenes.nested_aggregates)
    ("extra_nested_aggregates", 48, 31, 2),
]


# scenes also rendered through the reference's RandomMultisamplingRenderer (no test.mjs uses it; same world and camera), spp
RANDOM_RENDERER = {"BoxBall_DOF": 3, "cornell_box_emissive": 2}


def sources_digest(root, name):
    from .refjs import SOURCES
    h = hashlib.sha256()
    for f in SOURCES:
        h.update(open(os.path.join(root, "src", f), "rb").read())
    from .refjs import test_dir
    h.update(open(os.path.join(test_dir(root, name), "test.mjs"), "rb").read())
    return h.hexdigest()


def make(name, W, H, passes, seed=1):
    from .refjs import RefJS, REF_ROOT
    t0 = time.time()
    r = RefJS()
    info = r.load_test(name)
    js = r.scene_json(W, H)
    t1 = time.time()
    mean, rgba, draws = r.render(W, H, passes, seed=seed)
    extra = {}
    if info["renderer"] != "SimpleRenderer" and draws.max() <= 2:
        smean, srgba, sdraws = r.render_simple(W, H)
        assert sdraws.max() == 0
        extra = dict(simple_mean=smean, simple_rgba8=srgba)
    if name in RANDOM_RENDERER:
        rmean, rrgba, _ = r.render_random(W, H, RANDOM_RENDERER[name], seed=seed)
        extra.update(random_mean=rmean, random_rgba8=rrgba, random_spp=np.array(RANDOM_RENDERER[name]))
    meta = dict(name=name, width=W, height=H, passes=passes if info["renderer"] != "SimpleRenderer" else 1, seed=seed,
                renderer=info["renderer"], depth=info["maxRecursionDepth"], ref_width=info["width"], ref_height=info["height"],
                ref_spp=info["samplesPerPixel"], load_s=round(t1 - t0, 1), render_s=round(time.time() - t1, 1),
                sources_sha256=sources_digest(REF_ROOT, name), json_bytes=len(js))
    out = os.path.join(GOLDEN, "refjs_%s.npz" % name)
    np.savez_compressed(out, json=np.frombuffer(zlib.compress(js.encode("utf8"), 9), dtype=np.uint8), mean=mean, rgba8=rgba,
                        draws=draws, meta=np.array(json.dumps(meta)), **extra)
    return meta


def _worker(job):
    name, W, H, passes = job
    try:
        return make(name, W, H, passes)
    except Exception as e:      # noqa: BLE001 — reported per scene
        return dict(name=name, error="%s: %s" % (type(e).__name__, e))


def main(argv):
    import multiprocessing as mp
    jobs = [j for j in TABLE if not argv or j[0] in argv]
    # longest first
    slow = {"bunny": 9, "bunny_path": 9, "tie_fighter": 10, "utah_teapot": 8, "cat": 7, "cornell_box_path": 5, "SDF_Menger": 5}
    jobs.sort(key=lambda j: -slow.get(j[0], 0))
    with mp.Pool(min(len(jobs), max(1, (os.cpu_count() or 2) - 1))) as pool:
        for meta in pool.imap_unordered(_worker, jobs):
            print(json.dumps(meta), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
