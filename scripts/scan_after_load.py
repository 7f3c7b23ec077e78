"""does the scan's timing depend on what ran before it?  (a) cold, (b) right after a sustained GEMM load,
(c) after allocating and freeing several GB (allocator state like after bench.py's main legs)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppo_b200 as P
ctx = P.Context(0)
def scans(tag, n=8):
    ts = [ctx.bench_kernel("scan", 64 << 20, 15, 0, 0, 1, True)[0] * 1e3 for _ in range(n)]
    print(f"{tag:34s}", " ".join(f"{t:.0f}" for t in ts), flush=True)
scans("cold")
ctx.bench_kernel("tc3_fwd", 1 << 20, 512, 512, 0, 600, False)
scans("after 1 s of back-to-back GEMMs")
bufs = [P.DeviceRollouts(64, 16, 4, 1 << 20, ctx) for _ in range(3)]
for b in bufs: b.close()
scans("after 3 x 4.6 GB alloc/free")
ctx.bench_kernel("tc3_fwd", 1 << 20, 512, 512, 0, 2000, False)
scans("after 3 s of back-to-back GEMMs", 12)
time.sleep(2.0)
scans("2 s later")
ms, _ = ctx.bench_kernel("scan", 64 << 20, 15, 0, 0, 20, True); print("mean of 20:", f"{ms*1e3:.1f}")
