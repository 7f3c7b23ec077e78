#!/usr/bin/env bash
# Build libppo_b200.so in-tree for sm_100a (B200).  Usage: csrc/build.sh [-j N]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libppo_b200.so
NVCC=${NVCC:-nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       -Xptxas -v --expt-relaxed-constexpr)
mkdir -p build
pids=()
for src in abi.cu scan.cu shuffle.cu gather.cu loss.cu gemm_simt.cu gemm_tc.cu gemm_f16.cu adam.cu dp_p2p.cu bench_hooks.cu; do
  obj=build/${src%.cu}.o
  if [[ ! -f $obj || $src -nt $obj || common.cuh -nt $obj || gemm_tc.cuh -nt $obj || gemm_f16.cuh -nt $obj || tc_ptx.cuh -nt $obj || ../../include/ppo_b200.h -nt $obj ]]; then
    ( $NVCC "${FLAGS[@]}" -c "$src" -o "$obj" > "build/${src%.cu}.log" 2>&1 || { cat "build/${src%.cu}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for src in nccl_dl.cpp disk_replay.cpp; do
  obj=build/${src%.cpp}.o
  if [[ ! -f $obj || $src -nt $obj || common.cuh -nt $obj || ../../include/ppo_b200.h -nt $obj ]]; then
    ( $NVCC "${FLAGS[@]}" -x cu -c $src -o $obj > build/${src%.cpp}.log 2>&1 || { cat build/${src%.cpp}.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" build/*.o -lcudart -ldl -lpthread
echo "built $(realpath $OUT)"
