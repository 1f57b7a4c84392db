#!/bin/bash
# Development aid: build tuning variants of the library into out_of_nothing_b200/lib/variants/ (not shipped)
set -e
cd "$(dirname "$0")/../out_of_nothing_b200/csrc"
rm -rf ../lib/variants; mkdir -p ../lib/variants
build() { name=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -shared -o ../lib/variants/$name.so oon_kernels.cu oon_capi.cu -ldl & }
build default
wait
