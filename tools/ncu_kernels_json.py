#!/usr/bin/env python
"""`ncu --page raw --csv` export of a tools/gpu_ncu.sh capture -> profiles/r2_ncu_kernels.json, the per-kernel ncu block that
bench.py attaches to its roofline (the roofline must reproduce from files under profiles/ of the same
commit, not from a hand-scaled number).

    ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_kernels_json.py /tmp/raw.csv <scene> <width> <height> <passes per wave> <tag> > profiles/r2_ncu_kernels.json

Per kernel (all captured launches of it): launches, mean duration, and per launch the DRAM bytes (dram__bytes_read + write),
the L2 bytes (lts__t_bytes), then duration-weighted means of the utilisation figures."""
import csv
import json
import re
import subprocess
import sys

BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
PCT = {"issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
       "ipc": "sm__inst_executed.avg.per_cycle_active",
       "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
       "l1tex_wavefront_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
       "lts_throughput_pct": "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
       "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "registers": "launch__registers_per_thread"}


def short_name(full):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", full)
    if not m:
        return full[:40]
    name, targs = m.group(1), m.group(2) or ""
    if name in ("bvh_kernel", "prims_kernel", "sdf_kernel"):
        mode = re.search(r"<\s*(?:\(int\))?(\d)", targs)
        return "%s<%s>" % (name, "shadow" if mode and mode.group(1) == "1" else "extend")
    return name


def main():
    path, scene, W, H, passes, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, key):
        if key not in col:
            return None
        try:
            return float(r[col[key]].replace(",", ""))
        except ValueError:
            return None

    def scaled(r, key, table):
        v = num(r, key)
        return None if v is None else v * table.get(units[col[key]], 1.0)

    acc = {}
    for r in rows[2:]:
        k = short_name(r[col["Kernel Name"]])
        a = acc.setdefault(k, {"n": 0, "us": 0.0, "dram": 0.0, "lts": 0.0, "winst": 0.0, "w": {}})
        us = scaled(r, "gpu__time_duration.sum", TIME_US) or 0.0
        a["n"] += 1
        a["us"] += us
        a["dram"] += (scaled(r, "dram__bytes_read.sum", BYTES) or 0.0) + (scaled(r, "dram__bytes_write.sum", BYTES) or 0.0)
        a["lts"] += 32.0 * (num(r, "lts__t_sectors.sum") or 0.0)              # 32-byte sectors through the L2 tag stage
        a["smem_wf"] = a.get("smem_wf", 0.0) + (num(r, "memory_l1_wavefronts_shared") or 0.0)
        a["smem_wf_ideal"] = a.get("smem_wf_ideal", 0.0) + (num(r, "memory_l1_wavefronts_shared_ideal") or 0.0)
        a["winst"] += num(r, "smsp__inst_executed.sum") or 0.0
        for label, key in PCT.items():
            v = num(r, key)
            if v is not None:
                a["w"][label] = a["w"].get(label, 0.0) + v * us
    kernels = {}
    for k, a in acc.items():
        d = {"launches": a["n"], "duration_us": a["us"] / a["n"], "dram_bytes": a["dram"] / a["n"], "lts_bytes": a["lts"] / a["n"],
             "warp_instructions": a["winst"] / a["n"],
             "shared_wavefronts": a.get("smem_wf", 0.0) / a["n"], "shared_wavefronts_ideal": a.get("smem_wf_ideal", 0.0) / a["n"]}
        for label, s in a["w"].items():
            d[label] = round(s / a["us"], 3) if a["us"] > 0 else None
        kernels[k] = d
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"source": "profiles/r2_ncu_raw_%s.csv (ncu --set full --clock-control none, tools/gpu_ncu.sh %s); per-launch means over the captured launches; cold-cache, serialised replays: shares, not absolute times" % (tag, tag),
           "commit": head, "scene": scene, "width": W, "height": H, "passes_per_wave": passes, "kernels": kernels}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
