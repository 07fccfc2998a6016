"""TEST INFRASTRUCTURE — writes tests/golden/probes_refjs.npz: single calls of the reference's own functions (run in
oracle/jsvm) on random arguments, denser than what a small image reaches:

    cast_<scene>      rays (N, 8: origin xyz, direction xyz, minDistance, maxDistance) + intersectTransparent flag ->
                      `world.cast(ray, minD, maxD, flag).distance`                                  (src/world.js:29-31)
    sdf_<scene>       points (N, 3) -> `root_sdf.distance(p)` and the normal of `SDFGeometry.materialData`   (src/sdf.js:12-47)
    fmod              (a, b) -> `Math.fmod(a, b)`, including arguments whose 9th significant digit is an exact 5   (src/math.js:27)

The scenes are those of tests/golden/refjs_<scene>.npz (the oracle loads the JSON stored there).
    python -m oracle.refjs_probes
"""
from __future__ import annotations

import json
import os

import numpy as np

from .refjs import RefJS
from .refjs_golden import GOLDEN

CAST_SCENES = ["ASimpleScene", "Aggregates", "AMultipleBVH", "refraction", "cat", "SDF_Combinations", "SDF_Sierpinski", "spheres050"]
SDF_SCENES = ["SDF_Simple", "SDF_BoxBall", "SDF_Combinations", "SDF_Menger", "SDF_Sierpinski", "SDF_SphereRepetition"]


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def cast_probes(r, rng, n):
    vm = r.vm
    vm.run("var __cam = __test.renderer.camera.transform.column(3);")
    cam = np.array(vm.eval_expr("__cam").items[:3])
    rays = np.zeros((n, 8))
    flags = np.zeros(n, dtype=np.int32)
    out = np.zeros(n)
    fn = vm.eval_expr("(function(ox, oy, oz, dx, dy, dz, lo, hi, f) { return __test.renderer.world.cast(new Ray(Vec.of(ox, oy, oz, 1), Vec.of(dx, dy, dz, 0)), lo, hi, f).distance; })")
    for i in range(n):
        o = f32(cam + rng.normal(0, 1.5, 3) * (i % 3 > 0))
        d = f32(rng.normal(0, 1, 3) + np.array([0, -0.2, -1.5]))
        if i % 4 == 3:                       # a shadow ray: unnormalised towards a "light", window (1e-4, 1), opaque casters only
            d = f32(d * rng.uniform(3, 30))
            lo, hi, fl = 0.0001, 1.0, False
        else:
            lo, hi, fl = (0.0001 if i % 2 else 0.0), float("inf"), True
        rays[i] = [*o, *d, lo, hi]
        flags[i] = fl
        out[i] = vm.call(fn, None, [*map(float, o), *map(float, d), lo, hi, fl])
    return rays, flags, out


def sdf_probes(r, rng, n):
    vm = r.vm
    idx = int(vm.eval_expr("__test.renderer.world.objects.findIndex(o => o.geometry instanceof SDFGeometry)"))
    fn = vm.eval_expr("(function(i, x, y, z) { const g = __test.renderer.world.objects[i].geometry, p = Vec.of(x, y, z, 1);"
                      " const m = g.materialData({position: p}, Vec.of(0, 0, -1, 0)); return [g.root_sdf.distance(p), m.normal[0], m.normal[1], m.normal[2]]; })")
    pts = f32(rng.uniform(-2.5, 2.5, (n, 3)))
    pts[: n // 4] = f32(rng.normal(0, 0.6, (n // 4, 3)))
    out = np.zeros((n, 4))
    for i in range(n):
        out[i] = vm.call(fn, None, [float(idx), *map(float, pts[i])]).items
    return idx, pts, out


def fmod_probes(r, rng, n):
    vm = r.vm
    fn = vm.eval_expr("Math.fmod")
    a = rng.uniform(-60, 60, n)
    a[::3] = f32(a[::3])
    # exact decimal ties at the 9th significant digit
    k = n // 5
    a[:k] = rng.integers(10 ** 7, 10 ** 8, k) / 10.0 ** rng.integers(1, 7, k) + 0.5 / 10.0 ** 8
    a[k:2 * k] = (rng.integers(0, 2 ** 20, k) + 0.5) / 2.0 ** rng.integers(1, 16, k)
    b = rng.choice([2.0, 3.0, 1.5, 4.0, 2.5, 0.7, 6.0, 1.0], n)
    out = np.array([vm.call(fn, None, [float(x), float(y)]) for x, y in zip(a, b)])
    return a, b, out


def camera_probes(out_path):
    """camera rays of the reference's render loop when Math.random() returns what the device's counter-based generator
    returns for (seed, pixel, pass, node 1, dims 0..3) — so that the device's camera_ray() can be compared ray by ray (the
    scene documents are those of tests/golden/refjs_<scene>.npz: a camera does not depend on the image size)"""
    from . import oracle as orc
    L = orc.lib()
    arrays = {}
    for name, W, H, P, seed in (("BoxBall", 13, 9, 2, 1), ("BoxBall_DOF", 13, 9, 2, 5), ("bunny_path", 11, 7, 1, 9), ("cornell_box_path", 8, 8, 1, 2)):
        r = RefJS()
        r.load_test(name)

        def rnd(pixel, pass_, seed=seed):
            return [L.orc_kat_rng(seed, pixel, pass_, 1, d) for d in range(4)]
        rays = r.camera_rays(W, H, P, rnd)
        arrays["cam_%s_rays" % name] = rays
        arrays["cam_%s_meta" % name] = np.array([W, H, P, seed])
        print(name, rays.shape, flush=True)
    np.savez_compressed(out_path, **arrays)


def sdf_hit_probes(out_path):
    """`SDFGeometry.intersect(ray, minDistance, maxDistance)` (src/sdf.js:12-40, the sphere-tracing march) on rays in the
    geometry's own space, aimed at its bounding box from outside and from inside"""
    rng = np.random.default_rng(77)
    arrays = {}
    for name in SDF_SCENES:
        r = RefJS()
        r.load_test(name)
        vm = r.vm
        idx = int(vm.eval_expr("__test.renderer.world.objects.findIndex(o => o.geometry instanceof SDFGeometry)"))
        fn = vm.eval_expr("(function(i, ox, oy, oz, dx, dy, dz, lo, hi) { return __test.renderer.world.objects[i].geometry.intersect("
                          "new Ray(Vec.of(ox, oy, oz, 1), Vec.of(dx, dy, dz, 0)), lo, hi); })")
        n = 60 if name == "SDF_Menger" else 160
        rays = np.zeros((n, 8))
        out = np.zeros(n)
        for i in range(n):
            o = f32(rng.normal(0, 1, 3) * (3.0 if i % 3 else 0.4))
            target = rng.normal(0, 0.5, 3)
            d = f32((target - o) * rng.uniform(0.2, 3.0))               # unnormalised on purpose
            lo, hi = (0.0001, float("inf")) if i % 4 else (0.0001, 1.0)
            rays[i] = [*o, *d, lo, hi]
            out[i] = vm.call(fn, None, [float(idx), *map(float, o), *map(float, d), lo, hi])
        arrays["hit_%s_rays" % name], arrays["hit_%s_t" % name] = rays, out
        print(name, int(np.isfinite(out).sum()), "hits of", n, flush=True)
    np.savez_compressed(out_path, **arrays)


def texture_probes(out_path):
    """`TextureMaterialColor.color({UV})` (src/materials.js:101-130) of the two textures of the extra materials scene — bilinear
    with wrapped coordinates, nearest with clamped ones — at random and at edge-case UVs"""
    rng = np.random.default_rng(5)
    r = RefJS()
    r.load_test("extra_materials_whitted")
    vm = r.vm
    fn = vm.eval_expr("(function(tex, u, v) { return Array.from(tex.color({UV: Vec.of(u, v)})); })")
    texs = {"bilinear_wrap": vm.eval_expr("__test.renderer.world.objects[0].material.baseMaterial.diffusivity"),
            "nearest_clamp": vm.eval_expr("__test.renderer.world.objects[2].material._color")}
    edge = [0.0, 1.0, 0.5, 0.0625, 0.125, 0.9375, 0.25, 0.75, -0.0, 1e-9, 1 - 1e-9, -1.0, 2.0, 0.0624999, 0.1875]
    uv = [(a, b) for a in edge for b in edge[:6]] + [tuple(x) for x in rng.uniform(-2.5, 2.5, (400, 2))] + [tuple(x) for x in rng.uniform(0, 1, (300, 2))]
    uv = f32(np.array(uv))
    arrays = {"uv": uv}
    for k, t in texs.items():
        assert vm.eval_expr("(function(t) { return t instanceof TextureMaterialColor; })").__class__ is not None
        arrays["tex_" + k] = np.array([vm.call(fn, None, [t, float(a), float(b)]).items for a, b in uv])
    np.savez_compressed(out_path, **arrays)
    print({k: v.shape for k, v in arrays.items()})


def sdf_material_probes(out_path):
    """`root_sdf.getMaterialData(p)` (src/sdf.js: per-leaf basecolor, SphereSDF UVs, the selection rules of Union /
    Intersection / Difference and the blends of their smooth variants) at the probe points of probes_refjs.npz"""
    pts_all = np.load(os.path.join(GOLDEN, "probes_refjs.npz"))
    arrays = {}
    for name in SDF_SCENES:
        r = RefJS()
        r.load_test(name)
        vm = r.vm
        idx = int(vm.eval_expr("__test.renderer.world.objects.findIndex(o => o.geometry instanceof SDFGeometry)"))
        fn = vm.eval_expr("(function(i, x, y, z) { const m = __test.renderer.world.objects[i].geometry.root_sdf.getMaterialData(Vec.of(x, y, z, 1));"
                          " const b = m.basecolor, u = m.UV; return [b ? b[0] : NaN, b ? b[1] : NaN, b ? b[2] : NaN, u ? u[0] : NaN, u ? u[1] : NaN]; })")
        pts = pts_all["sdf_%s_p" % name]
        arrays["mat_%s" % name] = np.array([vm.call(fn, None, [float(idx), *map(float, p)]).items for p in pts])
        print(name, arrays["mat_%s" % name][:2].tolist(), flush=True)
    np.savez_compressed(out_path, **arrays)


def main():
    rng = np.random.default_rng(20241019)
    arrays = {}
    meta = {}
    for name in sorted(set(CAST_SCENES + SDF_SCENES)):
        r = RefJS()
        r.load_test(name)
        if name in CAST_SCENES:
            n = 150 if name in ("SDF_Combinations", "SDF_Sierpinski") else 600
            rays, flags, out = cast_probes(r, rng, n)
            arrays["cast_%s_rays" % name], arrays["cast_%s_flags" % name], arrays["cast_%s_t" % name] = rays, flags, out
        if name in SDF_SCENES:
            idx, pts, out = sdf_probes(r, rng, 200 if name == "SDF_Menger" else 500)
            arrays["sdf_%s_p" % name], arrays["sdf_%s_out" % name] = pts, out
            meta["sdf_%s_prim" % name] = idx
        print(name, "done", flush=True)
    if "fmod" not in arrays:
        a, b, out = fmod_probes(RefJS(), rng, 4000)
        arrays["fmod_a"], arrays["fmod_b"], arrays["fmod_out"] = a, b, out
    np.savez_compressed(os.path.join(GOLDEN, "probes_refjs.npz"), meta=np.array(json.dumps(meta)), **arrays)
    print({k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    import sys
    if "sdfmat" in sys.argv[1:]:
        sdf_material_probes(os.path.join(GOLDEN, "probes_sdfmat_refjs.npz"))
    elif "texture" in sys.argv[1:]:
        texture_probes(os.path.join(GOLDEN, "probes_texture_refjs.npz"))
    elif "sdfhit" in sys.argv[1:]:
        sdf_hit_probes(os.path.join(GOLDEN, "probes_sdfhit_refjs.npz"))
    elif "camera" in sys.argv[1:]:
        camera_probes(os.path.join(GOLDEN, "probes_camera_refjs.npz"))
    else:
        main()
