#!/bin/bash
# usage: tools/build_variant.sh "<extra nvcc flags>"  -> rebuilds k_kernels.cu with the flags and relinks
set -e
cd "$(dirname "$0")/../metalquicha_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p build
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC $1 -c k_kernels.cu -o build/k_kernels.o
nvcc $ARCH -shared -o ../libmqcb200.so build/*.o -lcudart -ldl
