#!/bin/bash
# Builds libpdn_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ ../../include/pdn_b200.h -nt "$o" ] || [ gemm_args.h -nt "$o" ] || [ gemm_tc.h -nt "$o" ] || [ tc_ptx.cuh -nt "$o" ] || [ conv_gather.cuh -nt "$o" ]; then
    $NVCC $FLAGS $EXTRA -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { echo "BUILD FAILED"; exit 1; }; done
$NVCC -shared -o ../libpdn_b200.so build/*.o -lcudart -ldl
echo "built $(pwd)/../libpdn_b200.so"
