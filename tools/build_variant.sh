#!/bin/bash
# usage: tools/build_variant.sh "<extra nvcc flags>"  -> rebuilds metalquicha_b200/libmqcb200.so
set -e
cd "$(dirname "$0")/../metalquicha_b200/csrc"
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p build
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC $1 -c k_kernels.cu -o build/k_kernels.o
nvcc $ARCH -shared -o ../libmqcb200.so build/engine.o build/prep_kernels.o build/j_kernels.o build/k_kernels.o -lcudart -ldl
