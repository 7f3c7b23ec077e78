"""the returns scan alone, for ncu (run on the GPU box): gamma = 1, gamma = 0.99, episodes of ~4096, one episode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_b200 as P
ctx = P.Context(0)
for a, b in ((15, 0), (15, 1), (4096, 1), (1 << 30, 1)):
    ms, work = ctx.bench_kernel("scan", 64 << 20, a, b, 0, 1, True)
    print(f"mean episode {a} gamma {'0.99' if b else '1'}: {ms*1e3:.1f} us {work/ms/1e6:.1f} GB/s", flush=True)
