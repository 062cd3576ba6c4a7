#!/bin/bash
# Variant builds of the library for A/B timing: tools/ab_build.sh <name> [-D... flags for clip_thread.cu / clip.cu]
# -> ab_build/libtess_<name>.so (objects of the other sources are compiled once and reused)
set -e
cd "$(dirname "$0")/../the-tessellator_b200/csrc"
OUT=../../ab_build
mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-O2,-Wall"
name=$1; shift
for f in capi grid outputs query clip clip_thread; do
  if [ ! -f $OUT/obj/$f.o ] || [ $f.cu -nt $OUT/obj/$f.o ] || [ common.cuh -nt $OUT/obj/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o $OUT/obj/$f.o &
  fi
done
wait
# AB_SRC=clip builds the variant of clip.cu instead of clip_thread.cu
src=${AB_SRC:-clip_thread}
nvcc $FLAGS "$@" -Xptxas -v -c $src.cu -o $OUT/obj/${src}_$name.o 2>&1 | grep -E "Compiling entry|registers|spill" | grep -A2 -E "${AB_GREP:-.}" | cut -c1-160 | head -${AB_LINES:-4}
objs=""
for f in capi grid outputs query clip clip_thread; do
  if [ $f = $src ]; then objs="$objs $OUT/obj/${f}_$name.o"; else objs="$objs $OUT/obj/$f.o"; fi
done
nvcc $FLAGS -shared -o $OUT/libtess_$name.so $objs -lcudart
echo built $OUT/libtess_$name.so
