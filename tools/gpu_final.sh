#!/bin/bash
# On the GPU box: the end-of-session validation in one gpurun call (~5 GPU-minutes):
#   gpurun --timeout 900 -- 'tools/gpu_final.sh <tag>'
# GPU test suite, the bench line, the reference arm, the ncu launch list of the bench command, one `ncu --set full` capture of the
# trace / shade kernels (tools/gpu_ncu.sh) and the per-scene table; everything lands in gpurun_out/<tag>_*.
tag="${1:-final}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py 2> gpurun_out/${tag}_bench_n1.err | grep "^{" > gpurun_out/${tag}_bench_n1.json
python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/${tag}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
tools/gpu_ncu.sh "$tag" > gpurun_out/${tag}_gpu_ncu.log 2>&1
tools/gpu_scene_table.sh "$tag" 2>&1 | tail -9
python tools/gpu_refjs_report.py > gpurun_out/${tag}_refjs_cuda_vs_reference.jsonl 2>&1    # CUDA path vs the reference's own images
cut -c1-400 gpurun_out/${tag}_bench_n1.json
# afterwards, here:  ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv > x.csv && python tools/ncu_digest.py x.csv
