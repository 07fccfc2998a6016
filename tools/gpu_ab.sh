#!/bin/bash
# On the GPU box: run tools/perf_probe.py for every variant library under jsraytracer_b200/variants/ (and the main one),
# once per environment setting in $SWEEP (space-separated NAME=VALUE items; default: one run with the plain environment).
#   SWEEP="JSRT_LEAF_TRIS=1 JSRT_LEAF_TRIS=4" LIBS="libjsrt.so variants/libjsrt_x.so" tools/gpu_ab.sh <tag> [scene W H passes]
tag="$1"; shift
scene="${1:-bunny_path}"; W="${2:-1920}"; H="${3:-1080}"; P="${4:-16}"; EXTRA="${5:-}"     # EXTRA: one more perf_probe keyword, e.g. dof=1
mkdir -p gpurun_out
out="gpurun_out/ab_${tag}.jsonl"; : > "$out"
libs="${LIBS:-$(cd jsraytracer_b200 && ls libjsrt.so variants/libjsrt_*.so 2>/dev/null)}"
for lib in $libs; do
  for kv in ${SWEEP:-_=_}; do
    r=$(env "$kv" JSRT_LIB="$PWD/jsraytracer_b200/$lib" python tools/perf_probe.py "$scene" "$W" "$H" "$P" $EXTRA 2>&1 | tail -1)
    echo "{\"lib\": \"$(basename $lib) $kv\", \"r\": $r}" >> "$out"
  done
done
python - "$out" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try:
        d=json.loads(l); r=d["r"]; print("%-44s %8.1f Mrays/s  ms %s" % (d["lib"], r["Mrays_s"], r["ms"]))
    except Exception as e: print("ERR", l[:300])
PY
