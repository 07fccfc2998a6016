#!/usr/bin/env python
"""Hot spots of one kernel from an `ncu --page source --csv` export (SASS view):
   tools/ncu_hot.py src.csv <kernel substring> [occurrence] [top]"""
import collections
import csv
import sys


def kernels(path):
    out, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    return out


def main():
    path, pat = sys.argv[1], sys.argv[2]
    occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ks = [k for k in kernels(path) if pat in k["name"]]
    # ncu prints each kernel twice (cuda view / sass view): keep those with an Address column
    ks = [k for k in ks if "Address" in k["hdr"]]
    k = ks[occ]
    h = k["hdr"]
    src, ie, smp, tie = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") or c.startswith("Stall")]
    rows = k["rows"]
    tot_i = sum(int(r[ie]) for r in rows)
    tot_s = sum(int(r[smp]) for r in rows)
    print(k["name"][:80], "| static", len(rows), "| warp-inst", tot_i, "| samples", tot_s, "| thr/inst %.1f" % (sum(int(r[tie]) for r in rows) / max(1, tot_i)))
    ops = collections.Counter()
    for r in rows:
        t = r[src].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[ie])
    print("opcode mix %:", " ".join("%s=%.1f" % (o, 100 * c / tot_i) for o, c in ops.most_common(22)))
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][smp]))[:top]
    print("top instructions by stall samples (line, %samples, %exec, thr, sass):")
    for i in sorted(idx):
        r = rows[i]
        e = int(r[ie])
        print("%5d %5.2f%% %5.2f%% %4.1f  %s" % (i, 100 * int(r[smp]) / tot_s, 100 * e / tot_i, int(r[tie]) / max(1, e), r[src].strip()[:90]))


if __name__ == "__main__":
    main()
