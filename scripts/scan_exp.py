import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_b200 as P
ctx = P.Context(0)
for n in (1 << 20, 64 << 20):
    for dbg in (0, 8, 4, 13):
        for ep in (15, 1 << 30):
            ms, work = ctx.bench_kernel("scan", n, ep, 0, dbg, 5, True)
            print(f"n={n>>20}M dbg={dbg} ep_len~{ep}: {ms*1e3:8.1f} us {work/ms/1e6:8.1f} GB/s", flush=True)
