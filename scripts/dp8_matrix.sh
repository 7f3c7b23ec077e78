#!/usr/bin/env bash
# 8-GPU diagnostic: how much of the data-parallel step is the gradient all-reduce, and how does NCCL's SM footprint interact
# with the persistent GEMM kernels?  Each line: variant, ms/step.
run() {
  echo "=== $1"
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --steps 2 --warmup 2 --profile 2>&1 | grep -E "rank 0\] resident|profile_only" | head -2
}
run "default (overlapped per-layer all-reduce)" "X=1"
run "no overlap (one all-reduce after backward)" "PPO_B200_NO_DP_OVERLAP=1"
run "overlap, NCCL_MAX_NCHANNELS=2" "NCCL_MAX_NCHANNELS=2"
run "no all-reduce at all (wrong results; timing floor)" "PPO_B200_SKIP_ALLREDUCE=1"
