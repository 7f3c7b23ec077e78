"""fp16-split engine: per-kernel timings through ppo_bench_kernel (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppo_b200 as P
ctx = P.Context(0)
for eng in ("tc1", "tc3"):
    for which in ("fwd", "dgrad", "wgrad"):
        for (M, K, N) in ((1 << 20, 512, 512), (1 << 20, 64, 512)):
            if which == "dgrad" and K == 64:
                continue
            ms, fl = ctx.bench_kernel(f"{eng}_{which}", M, K, N, 0, 3, True)
            print(f"{eng}_{which:6s} M={M} K={K} N={N} {ms:9.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s fp32-equivalent", flush=True)
for which, nm in (("head_fwd", "fp32 head_fwd"), ("head16_fwd", "f16 head_fwd"), ("head_bwd", "fp32 head_bwd"), ("head16_bwd", "f16 head_bwd")):
    ms, work = ctx.bench_kernel(which, 1 << 20, 512, 4, 0, 5, True)
    print(f"{nm:16s} {ms*1e3:9.1f} us  {work/ms/1e6:8.1f} GB/s (fp32-algorithmic bytes)", flush=True)
