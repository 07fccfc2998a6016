#!/bin/bash
# On the GPU box: launch list + one full ncu capture of the trace and shade kernels (B200_PROFILING.md recipe).
#   tools/gpu_ncu.sh <tag> [kernel regex] [skip] [count]
tag="$1"; rx="${2:-bvh_kernel|shade_kernel}"; skip="${3:-3}"; cnt="${4:-6}"
mkdir -p gpurun_out
# SCENE / W / H select the workload (default bunny_path 1920 1080); 2 passes per wave keep the capture short
export JSRT_BATCH_PASSES=2
cmd="python tools/perf_probe.py ${SCENE:-bunny_path} ${W:-1920} ${H:-1080} 2"
$cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv $cmd > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o gpurun_out/prof_$tag $cmd > gpurun_out/ncu_f_$tag.log 2>&1
tail -3 gpurun_out/ncu_f_$tag.log
ls -la gpurun_out/prof_$tag.ncu-rep
