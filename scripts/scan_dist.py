"""distribution of single-launch timings of the returns scan (64 M transitions, L2 flushed before each)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_b200 as P
ctx = P.Context(0)
for b in (0, 1):
    ts = []
    for i in range(30):
        ms, work = ctx.bench_kernel("scan", 64 << 20, 15, b, 0, 1, True)
        ts.append(ms * 1e3)
    print("gamma", "0.99" if b else "1", " ".join(f"{t:.0f}" for t in ts), flush=True)
for b in (0, 1):
    ms, work = ctx.bench_kernel("scan", 64 << 20, 15, b, 0, 20, True)
    print("mean of 20 in one call:", f"{ms*1e3:.1f} us")
    ms, work = ctx.bench_kernel("scan", 64 << 20, 15, b, 0, 5, True)
    print("mean of 5 in one call:", f"{ms*1e3:.1f} us")
