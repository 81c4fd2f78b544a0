#!/bin/bash
# Builds libabcdls.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
SRCS="api.cu select.cu dist.cu loclinear.cu tail.cu summary.cu smc.cu peer.cu small.cu single.cu"
[ -f bracket.cu ] && SRCS="$SRCS bracket.cu"
OBJS=""
for s in $SRCS; do
  o="${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ fused.cuh -nt "$o" ] || [ fused_wide.cuh -nt "$o" ] || [ compact.cuh -nt "$o" ] || [ peer.cuh -nt "$o" ] || [ columns.cuh -nt "$o" ] || [ ../../include/abcdls.h -nt "$o" ]; then
    $NVCC $FLAGS ${EXTRA:-} -c "$s" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libabcdls.so $OBJS
echo "built $(cd .. && pwd)/libabcdls.so"
