#!/bin/bash
# build_variant.sh <name> "<extra nvcc -D flags>"  ->  rvpt_b200/variants/lib<name>.so (A/B timing builds; RVPT_B200_LIB selects one)
set -e
cd "$(dirname "$0")/.."
mkdir -p rvpt_b200/variants/obj_$1
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden -cudart static"
for s in kernels.cu engine.cu bvh_gpu.cu bvh_build.cpp camera.cpp; do
  nvcc $F $2 -c rvpt_b200/csrc/$s -o rvpt_b200/variants/obj_$1/$s.o &
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o rvpt_b200/variants/lib$1.so rvpt_b200/variants/obj_$1/*.o
rm -rf rvpt_b200/variants/obj_$1
echo rvpt_b200/variants/lib$1.so
