#!/usr/bin/env python
"""Benchmark of the hot path: Mrays/s (and spp/s) at 1920x1080 on tests/bunny_path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = `--passes-per-step` (default 16) full-frame sample passes of the
reference's `IncrementalMultisamplingRenderer.render` loop (src/renderers.js:87-98)
on the BASELINE workload `tests/bunny_path` at 1920x1080, camera aspect 16/9,
depth 4 (BASELINE.json configs[2]; 16 steps = the config's 256 spp).  Rays are
counted as the reference would: one per `World.cast` call (primary, secondary
and shadow rays; src/world.js:28-30).

N > 1: one process per GPU (torchrun), scene replicated, sample passes sharded
across ranks (rank r renders its own pass indices; weak scaling: every rank
renders `passes-per-step` passes per step), accumulation buffers combined with
one NCCL reduce to rank 0 inside the timed region.

`--impl reference`: the reference's own CPU algorithm (the restatement oracle,
oracle/oracle.cpp — no JavaScript engine exists in this image) on all host
threads, same scene / metric, one pass per step.
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


def build_scene(args):
    from jsraytracer_b200 import scenes
    from jsraytracer_b200.serializer import Serializer
    test = scenes.configure(args.scene, width=args.width, height=args.height, aspect=args.width / args.height)
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
    line = {
        "impl": "reference", "metric": METRIC if args.scene == "bunny_path" and args.width == 1920 else "Mrays/s on %s %dx%d" % (args.scene, args.width, args.height),
        "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "spp_per_s": args.steps * ppass / dt,
        "config": {"workload": WORKLOAD if args.scene == "bunny_path" and args.width == 1920 else "%s %dx%d" % (args.scene, args.width, args.height),
                   "step": "1 full-frame pass per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": "%d full-frame passes; C++ restatement of the reference's JS algorithm (no JS engine in this image)" % args.steps},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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
    # all work (kernels, NCCL reduce, timing events) goes on one non-default torch stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    scene.set_stream(stream.cuda_stream)
    info = scene.info
    W, H = scene.size
    P = args.passes_per_step

    class _Accum:   # CUDA array interface view of the HBM-resident accumulation buffer
        __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "<f4", "data": (scene.accum_device_ptr(), False), "version": 2}
    accum_t = torch.as_tensor(_Accum(), device="cuda:%d" % local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from jsraytracer_b200.parallel import pass_block, reduce_accum, column_stripe, gather_columns
    step_no = [0]
    columns = args.shard == "columns" and world > 1

    def step():
        if columns:     # the reference's own split (src/worker.js:30-32): every rank renders its interleaved columns of the same passes
            x_offset, x_delt = column_stripe(rank, world)
            scene.render(step_no[0] * P, P, seed=1, x_offset=x_offset, x_delt=x_delt)
        else:
            first, n = pass_block(step_no[0], rank, world, P)
            scene.render(first, n, seed=1)
        step_no[0] += 1

    def combine():
        if columns:
            gather_columns(accum_t, dst=0)
        else:
            reduce_accum(accum_t, dst=0)

    # ---- warm-up ------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    if world > 1:
        combine()       # NCCL sets up its channels (and the send/recv connections of a gather) on first use
    scene.synchronize()

    # ---- timed region (device-resident inputs) ------------------------------------------
    scene.reset_accum()
    scene.stats_reset()
    scene.set_profiling(True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    combine()
    ev1.record(stream)
    barrier()
    sampler.stop_flag.set()
    ms = ev0.elapsed_time(ev1)
    st = scene.stats()
    scene.set_profiling(False)
    tmax = torch.tensor([ms], device="cuda", dtype=torch.float64)
    rays = torch.tensor([float(st["rays"]), float(st["launches"])], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    ms = float(tmax.item())
    total_rays, total_launches = float(rays[0].item()), int(rays[1].item())
    value = total_rays / (ms * 1e-3) / 1e6
    spp_per_s = args.steps * P * (1 if columns else world) / (ms * 1e-3)

    # ---- end-to-end leg: host buffers in, host image out, every step ---------------------
    # scene arrays host->device (jsrt_scene_upload), P passes, 8-bit resolve device->host
    # (jsrt_resolve_rgba8 = what CUDARenderer.render hands back in img.imgdata.data).
    img = np.empty(W * H * 4, dtype=np.uint8)
    e2e_steps = max(1, min(args.steps, 8))
    scene.reset_accum()
    scene.stats_reset()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.upload()
        step()
        scene.resolve_rgba8(img)
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
               "sample": "%d full-frame passes of the same workload on the C++ restatement oracle (%.1f s)" % (cpasses, cdt)}
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
    dominant = max(("extend", "shadow"), key=lambda k: kern_ms[k])
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
    avg_ms = kern_ms[dominant] / max(1, kern_n[dominant])
    rays_per_launch = rays_t / max(1, kern_n[dominant])
    achieved = bytes_per_ray * rays_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12            # SMs x FP32 lanes x 2 (FMA) x max SM clock, TFLOP/s
    # the scene is L2-resident, so the memory-side ceiling that can bind the walk is L2, measured here on this box
    # (SURVEY.md §8d): 128-bit reads of a 32 MB buffer from every SM
    try:
        l2_gbs = lib.measure_read_bandwidth(local, 32 << 20, 200)
    except Exception:
        l2_gbs = None
    kernel_name = {"extend": "bvh_kernel<extend> (+ prims_kernel<extend>)", "shadow": "bvh_kernel<shadow> (+ prims_kernel<shadow>)"}[dominant]
    if no_bvh:
        kernel_name = "prims_kernel<%s>" % dominant
    traffic = None
    try:   # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.scene == tj.get("scene") and args.width == tj.get("width"):
            traffic = tj["traffic_bytes_per_launch"].get(dominant)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch, "avg_launch_ms": avg_ms,
                "algorithmic_work": {"source": "oracle counters (reference algorithm, reference-built BVH)",
                                     "nodes_per_ray": nodes_per_ray, "tris_per_ray": prims_per_ray,
                                     "device_nodes_per_ray": d_nodes / max(1, d_rays), "device_tris_per_ray": d_prims / max(1, d_rays)},
                "l2": {"measured_read_gbs": l2_gbs, "frac": (achieved / l2_gbs) if l2_gbs else None,
                       "note": "algorithmic bytes / measured L2 read bandwidth (32 MB buffer, all SMs): the ceiling that applies to an L2-resident scene"},
                "fp32": {"flops_per_ray": flops_per_ray, "achieved_tflops": flops_per_ray * rays_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0,
                         "peak_tflops": fp32_peak, "note": "the HBM-side ceiling (peak / bytes_per_ray) is the lower of the two, so bound = hbm"},
                "kernel_ms": kern_ms, "kernel_launches": kern_n,
                "note": "scene (%.1f MB) is L1/L2-resident, so its algorithmic bytes never reach HBM and frac (of the HBM peak) can exceed 1; the l2 fraction is the meaningful one; the kernel is latency/issue bound, see DESIGN.md" % (info["scene_bytes"] / 1e6)}

    # ---- every kernel against its own ceiling (north_star: "each kernel reported as achieved fraction of its roofline") ----
    # Queue-streaming kernels (generate, prims, shade) are HBM-side: bytes = the queue records they must read and write
    # (ray 48 B, hit 16 B, shadow ray 48 B, work-list entry 8 B, one 16 B reduction per radiance term);
    # the BVH walks are L2-side: the reference algorithm's node + triangle bytes (oracle counters) over the measured L2
    # read bandwidth.  Times are the live CUDA-event sums of the timed region.
    def _k(name, ms_total, nbytes, peak_gbs, bound, per_unit, units):
        ach = nbytes / (ms_total * 1e-3) / 1e9 if ms_total > 0 else 0.0
        return {"kernel": name, "ms_per_step": ms_total / args.steps, "bound": bound, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                "frac": (ach / peak_gbs) if peak_gbs else None, "bytes_per_unit": per_unit, "units": units}
    n_cam, n_sec, n_sh = st["rays_primary"], st["rays_secondary"], st["rays_shadow"]
    n_ext = n_cam + n_sec
    o_ext_rays = ocnt["rays_primary"] + ocnt["rays_secondary"]
    ext_bpr = (32 * (ocnt["bvh_nodes_primary"] + ocnt["bvh_nodes_secondary"]) + 36 * (ocnt["bvh_prims_primary"] + ocnt["bvh_prims_secondary"])) / max(1, o_ext_rays)
    sh_bpr = (32 * ocnt["bvh_nodes_shadow"] + 36 * ocnt["bvh_prims_shadow"]) / max(1, ocnt["rays_shadow"])
    l2_peak = l2_gbs or 0.0
    kernels = [
        _k("generate_kernel", st["ms_generate"], n_cam * 64.0, peak, "hbm", 64, "camera samples"),
        _k("prims_kernel<extend>", st["ms_extend_prims"], n_ext * 48.0, peak, "hbm", 48, "rays (32 B read, 16 B hit written)"),
        _k("bvh_kernel<extend>", st["ms_extend_bvh"], n_ext * ext_bpr, l2_peak, "l2", ext_bpr, "rays (reference nodes x 32 B + triangles x 36 B)"),
        _k("shade_kernel", st["ms_shade"], n_ext * 64.0 + n_sec * 48.0 + n_sh * 48.0 + st["shaded_hits"] * 16.0, peak, "hbm", None,
           "64 B per ray read, 48 B per child and per shadow ray written, 16 B reduction per shaded hit"),
        _k("prims_kernel<shadow>", st["ms_shadow_prims"], n_sh * 48.0, peak, "hbm", 48, "shadow rays (32 B read + 16 B contribution or partial hit)"),
        _k("bvh_kernel<shadow>", st["ms_shadow_bvh"], n_sh * sh_bpr, l2_peak, "l2", sh_bpr, "shadow rays (reference nodes x 32 B + triangles x 36 B)"),
    ]
    roofline["kernels"] = kernels

    line = {
        "metric": METRIC if args.scene == "bunny_path" and args.width == 1920 else "Mrays/s on %s %dx%d" % (args.scene, args.width, args.height),
        "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if columns else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "spp_per_s": spp_per_s,
        "config": {"workload": WORKLOAD if args.scene == "bunny_path" and args.width == 1920 else "%s %dx%d" % (args.scene, W, H),
                   "passes_per_step": P, "parallelism": ("single GPU" if world == 1 else "column-striped x%d (x_offset = rank, x_delt = %d), scene replicated, 1 NCCL gather" % (world, world) if columns
                                   else "pass-sharded x%d, scene replicated, 1 NCCL reduce" % world),
                   "rng": "counter-based, seed 1", "l2": "wavefront queues (%.2f GB) exceed L2; the scene itself is L2-resident by nature" % (info["queue_bytes"] / 1e9),
                   "scene_create_s": create_s},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(info["scene_bytes"]), "d2h_bytes_per_step": W * H * 4, "steps": e2e_steps},
        "gpu_launches": total_launches,
        "rays": {"primary": st["rays_primary"], "secondary": st["rays_secondary"], "shadow": st["rays_shadow"]},
        "roofline": roofline,
    }
    if cpu:
        line["cpu_baseline"] = cpu
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
    ap.add_argument("--shard", default="passes", choices=["passes", "columns"],
                    help="N > 1: shard sample passes (weak scaling, one reduce) or interleaved columns like the reference's workers (strong scaling, one gather)")
    ap.add_argument("--cpu-passes", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
