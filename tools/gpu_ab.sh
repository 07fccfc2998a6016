#!/bin/bash
# On the GPU box: run tools/perf_probe.py for every variant library under jsraytracer_b200/variants/ (and the main one).
#   tools/gpu_ab.sh <tag> [scene W H passes]
tag="$1"; shift
scene="${1:-bunny_path}"; W="${2:-1920}"; H="${3:-1080}"; P="${4:-16}"
mkdir -p gpurun_out
out="gpurun_out/ab_${tag}.jsonl"; : > "$out"
for lib in jsraytracer_b200/libjsrt.so jsraytracer_b200/variants/libjsrt_*.so; do
  [ -f "$lib" ] || continue
  for rep in 1 2; do
    r=$(JSRT_LIB="$PWD/$lib" python tools/perf_probe.py "$scene" "$W" "$H" "$P" 2>&1 | tail -1)
    echo "{\"lib\": \"$(basename $lib)\", \"rep\": $rep, \"r\": $r}" >> "$out"
  done
done
python - "$out" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try:
        d=json.loads(l); r=d["r"]; print("%-28s %8.1f Mrays/s  ms %s" % (d["lib"], r["Mrays_s"], r["ms"]))
    except Exception as e: print("ERR", l[:300])
PY
