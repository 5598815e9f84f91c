#!/bin/bash
# tools/build_variant.sh NAME -DFLAG... : libjps variant with paint_sorted.cu recompiled under extra -D flags
# (A/B runs on the GPU box: cp tools/variants/NAME.so jax_powspec_b200/libjps.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  "$@" -c jax_powspec_b200/csrc/paint_sorted.cu -o tools/variants/$name.o
objs=$(ls jax_powspec_b200/build/*.o | grep -v paint_sorted.o)
nvcc -shared -o tools/variants/$name.so tools/variants/$name.o $objs -gencode arch=compute_100a,code=sm_100a -lcufft -Xlinker -rpath=/usr/local/cuda/lib64
rm tools/variants/$name.o
echo tools/variants/$name.so
