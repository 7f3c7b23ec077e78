import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppo_b200 as P
ctx = P.Context(0)
ms, work = ctx.bench_kernel("scan", 64 << 20, 15, 0, 0, 2, True)
print(f"{ms*1e3:.1f} us {work/ms/1e6:.1f} GB/s")
