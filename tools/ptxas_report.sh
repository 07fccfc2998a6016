#!/bin/bash
# Registers / stack / spills of every kernel in render.cu (extra -D flags may be passed).
cd "$(dirname "$0")/../jsraytracer_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c render.cu -o /tmp/render_report.o 2>&1 | c++filt > /tmp/ptxas.txt
python - <<'PY'
import re
t=open('/tmp/ptxas.txt').read()
for m in re.finditer(r"Compiling entry function '([^']*)' for 'sm_100a'\s*\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n.*?Used (\d+) registers", t):
    name=m.group(1).replace('jsrt::(anonymous namespace)::','')[:60]
    print("%-62s regs=%s stack=%s spill_st=%s spill_ld=%s"%(name,m.group(5),m.group(2),m.group(3),m.group(4)))
if 'error' in t: print(t[-3000:])
PY
