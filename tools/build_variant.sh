#!/bin/bash
# Build a variant of libcfem_b200.so with extra -D flags, for kernel A/B runs through CFEM_LIB.
#   tools/build_variant.sh minb8 -DCFEM_T16_MINB=8   ->  conservation-fem_b200/cfem_b200/libcfem_b200_minb8.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/conservation-fem_b200/csrc
tmp=$(mktemp -d)
for f in setup.cpp partition.cpp assembly.cu linalg.cu persist.cu rv.cu comm.cu euler.cu smooth.cu api.cu; do
  x=""; case "$f" in *.cpp) x="-x cu";; esac
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fopenmp,-O3 \
    --expt-relaxed-constexpr "$@" $x -c $src/$f -o $tmp/${f%.*}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fopenmp \
  -o $root/conservation-fem_b200/cfem_b200/libcfem_b200_$name.so $tmp/*.o /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -lcudart -lgomp -ldl
rm -rf $tmp
echo built libcfem_b200_$name.so
