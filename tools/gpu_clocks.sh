#!/bin/bash
# sample clocks/power during a long perf_probe run:  tools/gpu_clocks.sh <tag> [env...]
tag="$1"; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 100 > gpurun_out/clocks_$tag.csv &
SMI=$!
env "$@" python tools/perf_probe.py bunny_path 1920 1080 96 | tail -1 | cut -c1-220
kill $SMI
python - gpurun_out/clocks_$tag.csv <<'PY'
import sys,collections
rows=[l.strip().split(', ') for l in open(sys.argv[1]) if l.strip()]
busy=[r for r in rows if float(r[2].split()[0])>300]
print("samples",len(rows),"busy",len(busy))
if busy:
    sm=[int(r[0].split()[0]) for r in busy]; pw=[float(r[2].split()[0]) for r in busy]
    print("sm MHz min/med/max",min(sm),sorted(sm)[len(sm)//2],max(sm),"power W med/max",sorted(pw)[len(pw)//2],max(pw),"temp",busy[-1][3],"reasons",collections.Counter(r[4] for r in busy).most_common(3),"pcap",collections.Counter(r[5] for r in busy).most_common(2))
PY
