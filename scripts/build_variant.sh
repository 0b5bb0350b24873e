#!/bin/bash
# Tuning experiments: build a variant of libsegp.so with extra nvcc flags for ONE source file into build_variants/.
#   scripts/build_variant.sh <name> <source.cu> <flags...>      e.g.  scripts/build_variant.sh ks8 tri_i8.cu -DSEGP_KS_MINB=8
# On the GPU box a run picks a variant by copying it over safe_exploration_b200/libsegp.so (the snapshot is scratch).
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p build_variants
obj=build_variants/${src%.cu}_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" \
  -I include -I safe_exploration_b200/csrc -c safe_exploration_b200/csrc/$src -o $obj 2> build_variants/${src%.cu}_$name.log
objs=$(ls safe_exploration_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libsegp_$name.so $objs $obj -lcudart
echo built build_variants/libsegp_$name.so
