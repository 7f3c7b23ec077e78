#!/usr/bin/env bash
# 8-GPU A/B on one box: gradient exchange over peer memory (fused into Adam) vs NCCL all-reduce
run() {
  echo "=== $1"
  env $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --steps 2 --warmup 2 --profile 2>&1 | grep -E "rank 0\] resident|profile_only|rror" | head -3
}
run "peer memory, fused into Adam" "X=1"
run "NCCL (overlapped per-layer all-reduce)" "PPO_B200_NO_P2P=1"
