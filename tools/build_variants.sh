#!/bin/bash
# Development aid: build tuning variants of the library into out_of_nothing_b200/lib/variants/ (not shipped, git-ignored *.so)
set -e
cd "$(dirname "$0")/../out_of_nothing_b200/csrc"
mkdir -p ../lib/variants
build() { name=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -shared -o ../lib/variants/$name.so oon_kernels.cu oon_capi.cu -ldl & }
build base
build minb4 -DOON_K1_MINB=4
build minb2 -DOON_K1_MINB=2
build t4 -DOON_K1_T=4 -DOON_K1_MINB=2
build t1 -DOON_K1_T=1 -DOON_K1_MINB=4
build unroll1 -DOON_K1_UNROLL_CHECKED=1
build unroll4 -DOON_K1_UNROLL_CHECKED=4
wait
ls -la ../lib/variants
