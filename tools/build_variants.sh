#!/bin/bash
# Development aid: build tuning variants of the library into out_of_nothing_b200/lib/variants/ (not shipped)
set -e
cd "$(dirname "$0")/../out_of_nothing_b200/csrc"
rm -rf ../lib/variants; mkdir -p ../lib/variants
build() { name=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -shared -o ../lib/variants/$name.so oon_kernels.cu oon_capi.cu -ldl & }
build t1_sync
build t1_ring3 -DOON_K1_SYNC=0
build t1_ring4 -DOON_K1_SYNC=0 -DOON_K1_STAGES=4
build t1_ring6 -DOON_K1_SYNC=0 -DOON_K1_STAGES=6
build t1_ring8 -DOON_K1_SYNC=0 -DOON_K1_STAGES=8
build t1_ring8_u2 -DOON_K1_SYNC=0 -DOON_K1_STAGES=8 -DOON_K1_UNROLL_CHECKED=2
build t2_ring6 -DOON_K1_T=2 -DOON_K1_MINB=3 -DOON_K1_SYNC=0 -DOON_K1_STAGES=6
build t2_ring8 -DOON_K1_T=2 -DOON_K1_MINB=3 -DOON_K1_SYNC=0 -DOON_K1_STAGES=8
wait
ls ../lib/variants | tr '\n' ' '
