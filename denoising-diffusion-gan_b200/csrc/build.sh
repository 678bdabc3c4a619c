#!/bin/bash
# Builds libddgan_b200.so for sm_100a (cross-compiles without a GPU).  Output: ../lib/libddgan_b200.so
set -e
cd "$(dirname "$0")"
OUT=../lib
mkdir -p $OUT build
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include"
pids=()
for f in api_common conv_tc wgrad_tc attn_tc elementwise upfirdn2d lowprec_ops groupnorm train_ops train_pack optim; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ ../../include/ddgan_b200.h -nt build/$f.o ]; then
    nvcc $FLAGS -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libddgan_b200.so build/*.o -lcudart
echo built $OUT/libddgan_b200.so
