#!/bin/bash
# variant of the library that differs from the main build in ONE translation unit:
#   profiles/build_variant_one.sh <suffix> <unit.cu> <flags...>  ->  caracal_b200/libcaracal_gpu_<suffix>.so (select with CRCL_LIB_PATH)
set -e
cd "$(dirname "$0")/../caracal_b200"
suf=$1; unit=$2; shift 2
mkdir -p build_$suf
cp build/*.o build_$suf/
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o build_$suf/${unit%.cu}.o csrc/$unit
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libcaracal_gpu_$suf.so build_$suf/*.o -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo built libcaracal_gpu_$suf.so
