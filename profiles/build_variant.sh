#!/bin/bash
# build a variant of the library with extra -D flags: profiles/build_variant.sh <suffix> <flags...>
# -> caracal_b200/libcaracal_gpu_<suffix>.so (select with CRCL_LIB_PATH)
set -e
cd "$(dirname "$0")/../caracal_b200"
suf=$1; shift
mkdir -p build_$suf
for f in csrc/*.cu; do
  o=build_$suf/$(basename ${f%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o $o $f &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libcaracal_gpu_$suf.so build_$suf/*.o -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo built libcaracal_gpu_$suf.so
