#!/usr/bin/env bash
# tuning matrix of the fp16-split kk kernel: tile width, chain length, RZ compensation
for cfg in "0 2 0" "1 2 0" "0 4 0" "0 4 2.1e-8" "0 2 2.1e-8" "1 2 2.1e-8" "0 3 2.1e-8"; do
  set -- $cfg
  echo "=== BN128=$1 chunk_kb=$2 rz_comp=$3"
  PPO_F16_BN128=$1 PPO_F16_KK_CHUNK=$2 PPO_F16_RZ_COMP=$3 python tests/tool_engine_accuracy.py 3 2>&1 | grep -v "^$"
  PPO_F16_BN128=$1 PPO_F16_KK_CHUNK=$2 PPO_F16_RZ_COMP=$3 python scripts/f16_probe.py 2>&1 | grep "tc3_fwd\|tc3_dgrad"
done
