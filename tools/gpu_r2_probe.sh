#!/bin/bash
# On the GPU box: environment-toggle A/B of one scene through tools/perf_probe.py.
#   tools/gpu_r2_probe.sh <tag> <scene> <W> <H> <passes> "ENV1=a ENV2=b" "ENV1=c" ...   (each quoted item = one run; "_" = plain)
tag="$1"; scene="$2"; W="$3"; H="$4"; P="$5"; shift 5
mkdir -p gpurun_out
out="gpurun_out/ab_${tag}.jsonl"; : > "$out"
for kv in "$@"; do
  if [ "$kv" = "_" ]; then envs=""; else envs="$kv"; fi
  r=$(env $envs timeout 600 python tools/perf_probe.py "$scene" "$W" "$H" "$P" 2>&1 | tail -1)
  echo "{\"cfg\": \"$kv\", \"r\": $r}" >> "$out"
done
python - "$out" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try:
        d=json.loads(l); r=d["r"]; print("%-44s %8.1f Mrays/s  ms %s parts %s" % (d["cfg"], r["Mrays_s"], r["ms"], r["ms_parts"]))
    except Exception as e: print("ERR", l[:400])
PY
