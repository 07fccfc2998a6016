#!/bin/bash
# On the GPU box: per-scene throughput for every BASELINE config scene (tools/perf_probe.py), one line per scene.
mkdir -p gpurun_out
out="gpurun_out/scene_table_${1:-x}.jsonl"; : > "$out"
run() { python tools/perf_probe.py "$@" 2>&1 | tail -1 >> "$out"; }
run BoxBall 512 512 16
run cornell_box_path 1024 1024 8
run bunny_path 1920 1080 16
run dragon 1920 1080 16
run SDF_Menger 1920 1080 16 dof=1
run SDF_Sierpinski 1920 1080 16 dof=1
run starwars 1920 1080 8
run dragon_grid 1920 1080 8 n=3
python - "$out" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try:
        r=json.loads(l); print("%-18s %5dx%-5d p%-3d %9.1f Mrays/s  ms %s" % (r["scene"], r["size"][0], r["size"][1], r["passes"], r["Mrays_s"], r["ms"]))
    except Exception as e: print("ERR", l[:300])
PY
