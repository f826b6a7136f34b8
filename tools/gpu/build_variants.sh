#!/bin/bash
# Development: libskeres variants that differ in ONE source file only (-D flags), into gpurun_variants/ (git-ignored, travels to the GPU box).
# usage: [VARIANT_SRC=pcg_fused] tools/gpu/build_variants.sh name1 "-DFLAG=1" name2 "-DOTHER=2" ...   (default source: ba_kernels)
set -e
cd "$(dirname "$0")/../../skeres_b200/csrc"
make -j8 > /dev/null
SRC=${VARIANT_SRC:-ba_kernels}
mkdir -p ../../gpurun_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177,550 $flags -c -o /tmp/${SRC}_$name.o $SRC.cu
  objs=$(ls build/*.o | grep -v "build/$SRC.o")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libskeres_$name.so $objs /tmp/${SRC}_$name.o -lcudart -ldl -lpthread -lrt -Xcompiler -fPIC
  echo built gpurun_variants/libskeres_$name.so
done
