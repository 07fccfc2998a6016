#!/usr/bin/env python
"""Benchmark of the hot path: Mrays/s (and spp/s) at 1920x1080 on tests/bunny_path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: BASELINE.json configs[2] `tests/bunny_path` at 1920x1080, camera aspect 16/9, depth 4.  A step = 16 full-frame
sample passes of the reference's `IncrementalMultisamplingRenderer.render` loop (src/renderers.js:87-98); 16 steps are
the config's 256 spp.  Rays are counted as the reference would: one per `World.cast` call (primary, secondary and
shadow rays; src/world.js:28-30).

N = 1: `value` = rays / device time of the K steps + the final 8-bit resolve and its read-back (time to image).
N > 1: one process per GPU (torchrun), scene replicated, and the SAME frame: the 16 passes of every step are dealt to the
ranks (strong scaling; rank r renders a contiguous block of pass indices, so the library coalesces its calls into full
waves), and the image is composed once at the end — every rank's accumulation buffer is mapped on rank 0 through CUDA
IPC and summed over NVLink inside rank 0's resolve kernel (jsrt_accum_export / jsrt_accum_attach), the replacement of the
reference's compositing (src/raytrace_launcher.js:92-97).  The weak-scaling figure (16 passes per rank per step) is
reported beside it under "weak".

`--impl reference`: the reference's own CPU algorithm (the restatement oracle, oracle/oracle.cpp — no JavaScript engine
exists in this image) on all host threads, same scene / metric, one pass per step; builds its scene without loading
the CUDA library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "tests/bunny_path 1920x1080 aspect 16/9 depth 4 (BASELINE configs[2]; 256 spp = 16 steps x 16 passes)"
METRIC = "Mrays/s at 1080p on bunny_path"


# what makes the "port" a fair stand-in for the reference's own CPU implementation
PORT_PINNED = ("bit-identical (every f32 colour, every ImageData byte) to the reference's unmodified sources run in oracle/jsvm on 34 of its 37 "
               "demo scenes incl. bunny_path: tests/test_refjs_pin.py, profiles/r2_refjs_pin.md")

def is_headline(args):
    return args.scene == "bunny_path" and args.width == 1920 and args.height == 1080


def common_config(args):
    """The part of `config` both arms print identically (the driver compares the two lines' configs)."""
    return {"workload": WORKLOAD if is_headline(args) else "%s %dx%d" % (args.scene, args.width, args.height),
            "scene": args.scene, "width": args.width, "height": args.height, "camera_aspect": "%d/%d" % (args.width, args.height),
            "rng": "counter-based, seed 1", "rays_counted": "every World.cast: primary + secondary + shadow"}


def metric_name(args):
    return METRIC if is_headline(args) else "Mrays/s on %s %dx%d" % (args.scene, args.width, args.height)


def build_scene(args, scene=None):
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.serializer import Serializer
    test = scenes.configure(scene or args.scene, width=args.width, height=args.height, aspect=args.width / args.height)
    return Serializer(test)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 3 + i and s[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_reference_run(ser, args, passes, threads=None):
    """Times the CPU restatement of the reference algorithm on `passes` full-frame passes."""
    from oracle.oracle import OracleScene, default_threads
    threads = threads or default_threads()
    orc = OracleScene(ser.to_json())
    t0 = time.perf_counter()
    _, cnt = orc.render(passes, seed=1, jitter=True, width=args.width, height=args.height, threads=threads)
    dt = time.perf_counter() - t0
    rays = cnt["rays_primary"] + cnt["rays_secondary"] + cnt["rays_shadow"]
    return rays, dt, threads, cnt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ["JSRT_PY_BVH"] = "1"          # the scene is built by the pure-Python mirror: this arm maps no library of the repo but the oracle
    ser = build_scene(args)
    from oracle.oracle import OracleScene, default_threads
    threads = default_threads()
    orc = OracleScene(ser.to_json())
    ppass = 1
    done = 0
    for _ in range(args.warmup if args.warmup < 2 else 1):   # the CPU needs no warm-up waves; one pass pages everything in
        orc.render(ppass, first_pass=done, seed=1, width=args.width, height=args.height, threads=threads)
        done += ppass
    rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, cnt = orc.render(ppass, first_pass=done, seed=1, width=args.width, height=args.height, threads=threads)
        done += ppass
        rays += cnt["rays_primary"] + cnt["rays_secondary"] + cnt["rays_shadow"]
    dt = time.perf_counter() - t0
    val = rays / dt / 1e6
    from jsraytracer_b200 import lib as _lib
    line = {
        "impl": "reference", "metric": metric_name(args),
        "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "spp_per_s": args.steps * ppass / dt,
        "config": common_config(args),
        "run": {"step": "1 full-frame pass per step (bounded sample of the workload)", "cuda_library_loaded": _lib._LIB is not None},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": "%d full-frame passes; C++ restatement of the reference's JS algorithm (the image has no JS engine; oracle/jsvm runs the reference ~1000x slower than V8 and is a checker, not a baseline)" % args.steps,
                         "pinned": PORT_PINNED},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def ncu_profile_block():
    """Per-kernel ncu figures of the committed capture (profiles/r2_ncu_kernels.json, written by tools/ncu_kernels_json.py from
    the `ncu --set full` report of tools/gpu_ncu.sh on the same workload)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")))
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from jsraytracer_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the render path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ser = build_scene(args)
    blob = ser.to_msgpack()
    t0 = time.perf_counter()
    scene = lib.Scene(blob, lib.FORMAT_MSGPACK, device=local)
    create_s = time.perf_counter() - t0
    # all work (kernels, timing events) goes on one non-default torch stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    scene.set_stream(stream.cuda_stream)
    info = scene.info
    W, H = scene.size
    P = args.passes_per_step
    # the host image the resolve writes into: pinned, so that the device->host copy of every step is one DMA
    img = torch.empty(W * H * 4, dtype=torch.uint8, pin_memory=True).numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- cross-GPU composition: every rank's accumulation buffer mapped on rank 0 (CUDA IPC), summed inside its resolve kernel
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, scene.accum_export())
        if rank == 0:
            scene.accum_attach([h for r, h in enumerate(handles) if r != 0])

    from jsraytracer_b200.parallel import strong_share
    share = strong_share(P, rank, world)                 # this rank's passes of every step's P (strong scaling)
    next_pass = [rank * 10_000_000]                      # a pass-index range of its own per rank: contiguous, so calls coalesce

    def step(n_passes):
        if n_passes > 0:
            scene.render(next_pass[0], n_passes, seed=1)
            next_pass[0] += n_passes

    def compose():
        """Time to image: all ranks finish, rank 0 resolves (summing the peers' buffers over NVLink) and reads the bytes back."""
        scene.synchronize()
        if world > 1:
            dist.barrier()
        if rank == 0:
            scene.resolve_rgba8(img)
        if world > 1:
            dist.barrier()                               # the peers' buffers stay untouched until rank 0 has read them

    def reset():
        scene.reset_accum()
        scene.synchronize()
        barrier()

    def timed(n_steps, passes_each, sampler=None):
        reset()
        scene.stats_reset()
        barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(n_steps):
            step(passes_each)
        compose()
        ev1.record(stream)
        barrier()
        if sampler:
            sampler.stop_flag.set()
        ms = ev0.elapsed_time(ev1)
        st = scene.stats()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        r = torch.tensor([float(st["rays"]), float(st["launches"])], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
        return float(t.item()), float(r[0].item()), int(r[1].item()), st

    # ---- warm-up ------------------------------------------------------------------
    for _ in range(args.warmup):
        step(share)
    compose()

    # ---- timed region (device-resident inputs): the fixed frame ---------------------------
    scene.set_profiling(True)
    sampler = ClockSampler(local)
    ms, total_rays, total_launches, st = timed(args.steps, share, sampler)
    scene.set_profiling(False)
    value = total_rays / (ms * 1e-3) / 1e6
    spp_per_s = args.steps * P / (ms * 1e-3)

    # ---- weak scaling beside it (N > 1): every rank renders P passes per step ------------------
    weak = None
    if world > 1:
        wsteps = max(1, min(args.steps, 8))
        wms, wrays, _, _ = timed(wsteps, P)
        weak = {"value": wrays / (wms * 1e-3) / 1e6, "unit": "Mrays/s", "steps": wsteps, "ms_per_step": wms / wsteps,
                "passes_per_step": P * world, "note": "every rank renders %d passes per step, one composition at the end" % P}

    # ---- end-to-end leg: host buffers in, host image out, every step ---------------------
    # scene arrays host->device (jsrt_scene_upload), this rank's passes, composition across the ranks, 8-bit resolve
    # device->host (jsrt_resolve_rgba8 = what CUDARenderer.render hands back in img.imgdata.data).
    e2e_steps = max(1, min(args.steps, 8))
    reset()
    scene.stats_reset()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.upload()
        step(share)
        compose()
    barrier()
    e2e_s = time.perf_counter() - t0
    st2 = scene.stats()
    e2e_t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    e2e_r = torch.tensor([float(st2["rays"])], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_r, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_r.item()) / float(e2e_t.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline beside it (bounded sample, rank 0, N = 1 only) --------------------------
    # The same oracle run supplies the ALGORITHMIC work per ray for the roofline ("defined by
    # the reference algorithm on the reference-built BVH, counted by the oracle").
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpasses = args.cpu_passes
        crays, cdt, threads, ocnt = cpu_reference_run(ser, args, cpasses)
        cpu = {"value": crays / cdt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
               "sample": "%d full-frame passes of the same workload on the C++ restatement oracle (%.1f s)" % (cpasses, cdt),
               "pinned": PORT_PINNED}
    else:
        _, _, _, ocnt = cpu_reference_run(ser, args, 1)       # counts only (one pass, untimed leg)

    # ---- roofline of the dominant kernel --------------------------------------------------
    # Algorithmic bytes per ray = 32 B per BVH node the REFERENCE algorithm visits + 36 B per triangle it tests
    # (oracle counters, per ray class).  The device's own walk (collapsed leaves, octant order) visits fewer nodes
    # and tests more triangles; its counters (an instrumented, untimed pass) are reported beside it.
    scene.stats_reset()
    scene.render(1 << 20, 1, seed=1, flags=lib.FLAG_COUNT_WORK)
    cw = scene.stats()
    kern_ms = {"generate": st["ms_generate"], "extend": st["ms_extend"], "shade": st["ms_shade"], "shadow": st["ms_shadow"]}
    kern_n = {"generate": st["n_generate"], "extend": st["n_extend"], "shade": st["n_shade"], "shadow": st["n_shadow"]}
    part_ms = {k: st["ms_" + k] for k in ("extend_prims", "extend_bvh", "extend_sdf", "shadow_prims", "shadow_bvh", "shadow_sdf")}
    has_walk = (part_ms["extend_bvh"] + part_ms["shadow_bvh"]) > 0.05 * (kern_ms["extend"] + kern_ms["shadow"])
    dominant = max(("extend", "shadow"), key=lambda k: part_ms[k + "_bvh"] if has_walk else kern_ms[k])     # the wave whose walk kernel is the longest
    if dominant == "extend":
        o_rays = ocnt["rays_primary"] + ocnt["rays_secondary"]
        o_nodes = ocnt["bvh_nodes_primary"] + ocnt["bvh_nodes_secondary"]
        o_prims = ocnt["bvh_prims_primary"] + ocnt["bvh_prims_secondary"]
        d_rays = cw["rays_primary"] + cw["rays_secondary"]
        d_nodes, d_prims = cw["bvh_nodes"][0] + cw["bvh_nodes"][1], cw["bvh_prims"][0] + cw["bvh_prims"][1]
        rays_t = st["rays_primary"] + st["rays_secondary"]
    else:
        o_rays, o_nodes, o_prims = ocnt["rays_shadow"], ocnt["bvh_nodes_shadow"], ocnt["bvh_prims_shadow"]
        d_rays, d_nodes, d_prims = cw["rays_shadow"], cw["bvh_nodes"][2], cw["bvh_prims"][2]
        rays_t = st["rays_shadow"]
    nodes_per_ray, prims_per_ray = o_nodes / max(1, o_rays), o_prims / max(1, o_rays)
    bytes_per_ray = 32 * nodes_per_ray + 36 * prims_per_ray
    no_bvh = o_nodes == 0           # a scene of analytic primitives only (cornell_box_path): prims_kernel streams the queue
    if no_bvh:
        bytes_per_ray = 48.0        # 32 B ray record read + 16 B hit / contribution written
    flops_per_ray = 27 * nodes_per_ray + 40 * prims_per_ray
    # the dominant KERNEL is the walk of that wave (bvh_kernel); its own event time, not the wave's
    dom_ms_total = part_ms[dominant + "_bvh"] if not no_bvh else part_ms[dominant + "_prims"]
    avg_ms = dom_ms_total / max(1, kern_n[dominant])
    rays_per_launch = rays_t / max(1, kern_n[dominant])
    achieved = bytes_per_ray * rays_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12            # SMs x FP32 lanes x 2 (FMA) x max SM clock, TFLOP/s
    # the scene is L2-resident, so the memory-side ceiling that can bind the walk is L2, measured here on this box
    # (SURVEY.md §8d): 128-bit reads of a 32 MB buffer from every SM
    try:
        l2_gbs = lib.measure_read_bandwidth(local, 32 << 20, 200)
    except Exception:
        l2_gbs = None
    l2_resident = info["scene_bytes"] < 100e6
    kernel_name = "prims_kernel<%s>" % dominant if no_bvh else "bvh_kernel<%s>" % dominant
    ncu = ncu_profile_block()
    ncu_dom = (ncu or {}).get("kernels", {}).get(kernel_name) if ncu and ncu.get("scene") == args.scene and ncu.get("width") == args.width else None
    traffic = ncu_dom.get("dram_bytes") if ncu_dom else None
    if l2_resident and not no_bvh and l2_gbs:
        bound, peak, peak_src = "l2", l2_gbs, "measured on this box: 128-bit reads of an L2-resident 32 MB buffer from every SM (jsrt_measure_read_bandwidth)"
    else:
        bound, peak, peak_src = "hbm", hbm_peak, hbm_src
    roofline = {
        "bound": bound, "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_ms,
        "algorithmic_work": {"source": "oracle counters (reference algorithm, reference-built BVH)",
                             "nodes_per_ray": nodes_per_ray, "tris_per_ray": prims_per_ray,
                             "device_nodes_per_ray": d_nodes / max(1, d_rays), "device_tris_per_ray": d_prims / max(1, d_rays)},
        "hbm": {"peak": hbm_peak, "frac": achieved / hbm_peak, "peak_source": hbm_src,
                "note": "not a utilisation: the %.1f MB scene never leaves L1 / L2 / shared memory, so its algorithmic bytes do not reach HBM (see traffic)" % (info["scene_bytes"] / 1e6)},
        "fp32": {"flops_per_ray": flops_per_ray, "achieved_tflops": flops_per_ray * rays_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0,
                 "peak_tflops": fp32_peak},
        "what_binds": "issue slots and SIMT utilisation of the walk, not a memory ceiling: see the ncu block (issue %, threads per instruction, lts bytes); DESIGN.md §4",
        "ncu": ncu_dom, "ncu_source": (ncu or {}).get("source"),
        "kernel_ms": kern_ms, "kernel_part_ms": part_ms, "kernel_launches": kern_n,
    }

    # ---- every kernel against its own ceiling (north_star: "each kernel reported as achieved fraction of its roofline") ----
    # Queue-streaming kernels (prims, shade) are HBM-side: bytes = the queue records they must read and write
    # (ray 48 B, hit 16 B, shadow walker 48 B, work-list entry 8 B, one 16 B reduction per radiance term);
    # the BVH walks are L2-side: the reference algorithm's node + triangle bytes (oracle counters) over the measured L2
    # read bandwidth.  Times are the live CUDA-event sums of the timed region; the ncu columns come from the committed capture.
    def _k(name, ms_total, nbytes, peak_gbs, bound, per_unit, units):
        ach = nbytes / (ms_total * 1e-3) / 1e9 if ms_total > 0 else 0.0
        d = {"kernel": name, "ms_per_step": ms_total / args.steps, "bound": bound, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
             "frac": (ach / peak_gbs) if peak_gbs else None, "bytes_per_unit": per_unit, "units": units}
        if ncu and name in ncu.get("kernels", {}):
            d["ncu"] = ncu["kernels"][name]
        return d
    n_cam, n_sec, n_sh = st["rays_primary"], st["rays_secondary"], st["rays_shadow"]
    n_ext = n_cam + n_sec
    o_ext_rays = ocnt["rays_primary"] + ocnt["rays_secondary"]
    ext_bpr = (32 * (ocnt["bvh_nodes_primary"] + ocnt["bvh_nodes_secondary"]) + 36 * (ocnt["bvh_prims_primary"] + ocnt["bvh_prims_secondary"])) / max(1, o_ext_rays)
    sh_bpr = (32 * ocnt["bvh_nodes_shadow"] + 36 * ocnt["bvh_prims_shadow"]) / max(1, ocnt["rays_shadow"])
    l2_peak = l2_gbs or 0.0
    kernels = [
        _k("prims_kernel<extend>", st["ms_extend_prims"] + st["ms_generate"], n_cam * 64.0 + n_sec * 48.0, hbm_peak, "hbm", None,
           "camera rays: 48 B written + 16 B hit; other rays: 32 B read + 16 B hit written"),
        _k("bvh_kernel<extend>", st["ms_extend_bvh"], n_ext * ext_bpr, l2_peak, "l2", ext_bpr, "rays (reference nodes x 32 B + triangles x 36 B)"),
        _k("shade_kernel", st["ms_shade"], n_ext * 64.0 + n_sec * 48.0 + st["shaded_hits"] * 16.0, hbm_peak, "hbm", None,
           "64 B per ray read, 48 B per child (and per shadow walker, not counted) written, 16 B reduction per shaded hit"),
        _k("prims_kernel<shadow>", st["ms_shadow_prims"], 0.0 if st["ms_shadow_prims"] < 0.02 * max(1e-9, st["ms_shadow"]) else n_sh * 48.0, hbm_peak, "hbm", 48,
           "shadow rays (only scenes whose shadow tests are not fused into shade_kernel)"),
        _k("bvh_kernel<shadow>", st["ms_shadow_bvh"], n_sh * sh_bpr, l2_peak, "l2", sh_bpr, "shadow rays (reference nodes x 32 B + triangles x 36 B)"),
    ]
    roofline["kernels"] = kernels

    # ---- second workload north_star names: tests/cornell_box_path at 1920x1080 (N = 1, short) ----------
    extra = None
    if world == 1 and is_headline(args) and not args.no_extra:
        try:
            scene.close()
            ser2 = build_scene(args, "cornell_box_path")
            sc2 = lib.Scene(ser2.to_msgpack(), lib.FORMAT_MSGPACK, device=local)
            sc2.set_stream(stream.cuda_stream)
            sc2.render(0, 8, seed=1); sc2.synchronize()
            sc2.reset_accum(); sc2.stats_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(4):
                sc2.render(8 + 8 * k, 8, seed=1)
            sc2.resolve_rgba8(img)
            e1.record(stream)
            torch.cuda.synchronize()
            s2 = sc2.stats()
            ms2 = e0.elapsed_time(e1)
            extra = {"cornell_box_path_1080p": {"value": s2["rays"] / (ms2 * 1e-3) / 1e6, "unit": "Mrays/s", "passes": 32, "ms": ms2,
                                                "spp_per_s": 32 / (ms2 * 1e-3), "config": "tests/cornell_box_path 1920x1080 aspect 16/9 depth 8, 4 area-light samples per hit"}}
            sc2.close()
        except Exception as e:          # the headline line must not die on the side measurement
            extra = {"cornell_box_path_1080p": {"error": str(e)[:200]}}

    line = {
        "metric": metric_name(args),
        "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "spp_per_s": spp_per_s,
        "config": common_config(args),
        "run": {"passes_per_step": P, "passes_per_step_this_rank": share,
                "parallelism": "single GPU" if world == 1 else "fixed frame: the %d passes of a step dealt to %d ranks, scene replicated, accumulation buffers summed over NVLink (CUDA IPC peer reads) inside rank 0's resolve kernel" % (P, world),
                "timed_region": "K steps + composition + 8-bit resolve + read-back of the image",
                "l2": "wavefront queues (%.2f GB) exceed L2; the scene itself is L2-resident by nature" % (info["queue_bytes"] / 1e9),
                "scene_create_s": create_s},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(info["scene_bytes"]), "d2h_bytes_per_step": W * H * 4, "steps": e2e_steps},
        "gpu_launches": total_launches,
        "rays": {"primary": st["rays_primary"], "secondary": st["rays_secondary"], "shadow": st["rays_shadow"]},
        "roofline": roofline,
    }
    if weak:
        line["weak"] = weak
    if cpu:
        line["cpu_baseline"] = cpu
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="bunny_path")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--passes-per-step", type=int, default=16)
    ap.add_argument("--cpu-passes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
