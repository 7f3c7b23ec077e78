"""one launch group per fp16-engine kernel, for ncu (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppo_b200 as P
ctx = P.Context(0)
for which, M, K, N in (("tc3_fwd", 1 << 20, 512, 512), ("tc3_fwd", 1 << 20, 64, 512), ("tc3_dgrad", 1 << 20, 512, 512),
                       ("tc3_wgrad", 1 << 20, 512, 512), ("head16_bwd", 1 << 20, 512, 4), ("head16_fwd", 1 << 20, 512, 4)):
    ms, fl = ctx.bench_kernel(which, M, K, N, 0, 1, True)
    print(which, M, K, N, ms, flush=True)
