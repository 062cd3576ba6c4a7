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
nvcc $FLAGS "$@" -Xptxas -v -c clip_thread.cu -o $OUT/obj/clip_thread_$name.o 2>&1 | grep -E "registers|spill" | head -4
nvcc $FLAGS -shared -o $OUT/libtess_$name.so $OUT/obj/capi.o $OUT/obj/grid.o $OUT/obj/outputs.o $OUT/obj/query.o $OUT/obj/clip.o $OUT/obj/clip_thread_$name.o -lcudart
echo built $OUT/libtess_$name.so
