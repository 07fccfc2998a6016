#!/usr/bin/env python
"""Digest of an `ncu --page raw --csv` export: one line of key metrics per captured kernel launch."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "ms"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("sm__inst_executed.avg.per_cycle_active", "IPC"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wave%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("smsp__inst_executed.sum", "winst"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].replace("void jsrt::<unnamed>::", "").split("(jsrt")[0]
        out = [name[:34].ljust(34)]
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    if units[i] in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                        v = "%.0fMB" % (f / 1e6)
                    elif label == "winst":
                        v = "%.1fM" % (f / 1e6)
                    else:
                        v = "%.2f" % f
                except ValueError:
                    pass
                out.append("%s=%s" % (label, v))
        print(" ".join(out))


if __name__ == "__main__":
    main()
