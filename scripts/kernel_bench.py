"""Per-kernel timings through ppo_bench_kernel (run on the GPU box)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppo_b200 as P
ctx = P.Context(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def hbm(name, which, n, a=0, b=0, c=0, iters=10):
    ms, work = ctx.bench_kernel(which, n, a, b, c, iters, True)
    print(f"{name:28s} {ms*1e3:9.1f} us  {work/ms/1e6:8.1f} GB/s  {100*work/ms/1e6/PEAK:5.1f} %", flush=True)
hbm("scan 1M", "scan", 1 << 20, 15)
hbm("scan 64M", "scan", 64 << 20, 15, iters=5)
hbm("scan 64M gamma.99", "scan", 64 << 20, 15, 1, iters=5)
hbm("scan 64M +K2 stats", "scan_norm", 64 << 20, 15, iters=5)
hbm("scan 64M no-terminals", "scan", 64 << 20, 1 << 30, 1, iters=5)
hbm("loss 64k A=64", "loss", 65536, 64)
hbm("loss 1M A=64", "loss", 1 << 20, 64, iters=5)
hbm("loss 256k A=256", "loss", 1 << 18, 256, iters=5)
hbm("gather ldg c3", "gather0", 1 << 20, 1024, 64, 65536)
hbm("gather ldg c2", "gather0", 65536, 72 * 64, 256, 4096)
hbm("head_fwd 1M x512", "head_fwd", 1 << 20, 512, 4, iters=5)
hbm("head_bwd 1M x512", "head_bwd", 1 << 20, 512, 4, iters=5)
hbm("adam 560k", "adam", 560644)
hbm("shuffle 1M", "shuffle", 1 << 20)
for which in ("tc1_fwd", "tc1_dgrad", "tc1_wgrad"):
    ms, fl = ctx.bench_kernel(which, 1 << 20, 512, 512, 3, 3, True)
    print(f"{which:28s} {ms:9.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s fp32-equivalent", flush=True)
