#!/bin/bash
# tools/build_variant.sh <name> [extra nvcc flags...]  ->  ab/lib_<name>.so
# Rebuilds gemm_tcgen05.cu / attention_tcp.cu with the extra flags and links them with the regular objects
# (run `python -m aihab_clip_b200.build` first).  For same-box A/B runs with tools/ab_lib.sh.
set -e
cd "$(dirname "$0")/../aihab_clip_b200/csrc"
NAME=$1; shift
mkdir -p ../../ab build/var_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden"
OBJS=""
for f in gemm_tcgen05 attention_tcp attention_tcd; do
  nvcc $FLAGS "$@" -c $f.cu -o build/var_$NAME/$f.o &
done
wait
for f in elementwise attention attention_tcf score preprocess api; do OBJS="$OBJS build/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/lib_$NAME.so build/var_$NAME/*.o $OBJS -cudart static
echo ab/lib_$NAME.so
