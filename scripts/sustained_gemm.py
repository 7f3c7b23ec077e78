"""Sustained (power-capped) per-kernel GEMM timings: many back-to-back launches without L2 flush, vs the isolated,
flushed timing the bench hooks report.  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppo_b200 as P
ctx = P.Context(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for which in ("tc3_fwd", "tc3_dgrad", "tc3_wgrad", "head16_fwd", "head16_bwd"):
    for iters, flush in ((3, True), (300, False)):
        if which.startswith("head"):
            ms, work = ctx.bench_kernel(which, M, 512, 4, 0, iters, flush)
            print(f"{which:12s} M={M} iters={iters:4d} flush={flush}: {ms*1e3:8.1f} us  {work/ms/1e6:8.1f} GB/s", flush=True)
        else:
            ms, fl = ctx.bench_kernel(which, M, 512, 512, 0, iters, flush)
            print(f"{which:12s} M={M} iters={iters:4d} flush={flush}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s fp32-equiv", flush=True)
