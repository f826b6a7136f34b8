#!/bin/bash
# Development: libskeres variants that differ in ba_kernels.cu only (-D flags), into gpurun_variants/ (git-ignored, travels to the GPU box).
# usage: tools/gpu/build_variants.sh name1 "-DFLAG=1" name2 "-DOTHER=2" ...
set -e
cd "$(dirname "$0")/../../skeres_b200/csrc"
make -j8 > /dev/null
mkdir -p ../../gpurun_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177,550 $flags -c -o /tmp/ba_kernels_$name.o ba_kernels.cu
  objs=$(ls build/*.o | grep -v ba_kernels.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libskeres_$name.so $objs /tmp/ba_kernels_$name.o -lcudart_static -ldl -lpthread -lrt -Xcompiler -fPIC
  echo built gpurun_variants/libskeres_$name.so
done
