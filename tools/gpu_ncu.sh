#!/bin/bash
# On the GPU box: launch list + one full ncu capture of the trace and shade kernels (B200_PROFILING.md recipe).
#   tools/gpu_ncu.sh <tag> [kernel regex] [skip] [count] [passes]
# Defaults capture the second (timed) 16-pass wave of tools/perf_probe.py: 4 levels x {prims, bvh, shade, bvh<shadow>}.
tag="$1"; rx="${2:-bvh_kernel|shade_kernel|prims_kernel}"; skip="${3:-16}"; cnt="${4:-16}"; passes="${5:-16}"
mkdir -p gpurun_out
cmd="python tools/perf_probe.py ${SCENE:-bunny_path} ${W:-1920} ${H:-1080} $passes"
$cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o gpurun_out/prof_$tag $cmd > gpurun_out/ncu_f_$tag.log 2>&1
tail -3 gpurun_out/ncu_f_$tag.log
ls -la gpurun_out/prof_$tag.ncu-rep
