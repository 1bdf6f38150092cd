#!/bin/bash
# Builds libnwb.so in-tree for sm_100a (called by __graft_entry__.build()).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libnwb.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -pthread -Xcompiler -fvisibility=hidden
       --expt-relaxed-constexpr -cudart static -Xptxas -v -DNWB_FAST_SINCOS ${NWB_EXTRA_FLAGS:-})
SRCS=(validity.cu nwb_abi.cu tree.cu nn_tensor.cu planner_kernels.cu planner_abi.cu rrt_driver.cu arm_kernels.cu service_abi.cu)
OBJS=()
PIDS=()
NAMES=()
mkdir -p build
for s in "${SRCS[@]}"; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o")" ] || [ ../../include/nwb.h -nt "$o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" 2> "build/${s%.cu}.ptxas.log" &     # translation units compile side by side
    PIDS+=($!)
    NAMES+=("${s%.cu}")
  fi
  OBJS+=("$o")
done
for k in "${!PIDS[@]}"; do
  if ! wait "${PIDS[$k]}"; then
    cat "build/${NAMES[$k]}.ptxas.log"
    rm -f "build/${NAMES[$k]}.o"
    exit 1
  fi
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -pthread -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
