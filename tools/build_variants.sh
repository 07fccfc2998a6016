#!/bin/bash
# Build A/B variants of libjsrt.so with different -D tuning flags (selected at run time with JSRT_LIB=...).
#   tools/build_variants.sh name1:"-DJSRT_X=1 -DJSRT_Y=2" name2:"..."
# Output: jsraytracer_b200/variants/libjsrt_<name>.so (git-ignored; travels to the GPU box with gpurun).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CSRC="$ROOT/jsraytracer_b200/csrc"
OUT="$ROOT/jsraytracer_b200/variants"
OBJ=/tmp/jsrt_variant_obj
mkdir -p "$OUT" "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for f in api wire scene_flatten sdf_compile bvh_build obj_parse; do
  nvcc $FLAGS -c "$CSRC/$f.cpp" -o "$OBJ/$f.o" &
done
wait
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( nvcc $FLAGS $defs -c "$CSRC/render.cu" -o "$OBJ/render_$name.o" && \
    nvcc -shared -o "$OUT/libjsrt_$name.so" "$OBJ/render_$name.o" "$OBJ"/{api,wire,scene_flatten,sdf_compile,bvh_build,obj_parse}.o && echo "built $name ($defs)" ) &
done
wait
ls -la "$OUT"
