"""Bring-up probe for the tcgen05 engine: one Dense op per subprocess (a hang cannot block the rest)."""
import ctypes as C
import subprocess
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(mode, op, M, K, N, seed=0, ints=False):
    import ppo_b200 as P
    from ppo_b200 import _lib
    ctx = P.Context(0)
    rng = np.random.default_rng(seed)
    X = (rng.integers(-3, 9, (M, K)) if ints else rng.normal(size=(M, K))).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    dY = rng.normal(size=(M, N)).astype(np.float32)
    X64, W64, dY64 = X.astype(np.float64), W.astype(np.float64), dY.astype(np.float64)
    if op == 0:
        z = X64 @ W64 + b
        want = np.where(z > 0, z, 0.01 * z)
        out = np.empty((M, N), np.float32)
    elif op == 1:
        g = dY64 @ W64.T
        want = np.where(X64 > 0, g, 0.01 * g)
        out = np.empty((M, K), np.float32)
    else:
        want = X64.T @ dY64
        out = np.empty((K, N), np.float32)
    out2 = np.zeros(max(K, N), np.float32)
    lib = _lib.load()
    _lib.check(lib.ppo_dense_op(ctx.handle, mode, op, M, K, N, _lib.ptr(X, C.c_float), _lib.ptr(W, C.c_float),
                                _lib.ptr(b, C.c_float), _lib.ptr(dY, C.c_float), 0.01, _lib.ptr(out, C.c_float),
                                _lib.ptr(out2, C.c_float)))
    err = np.abs(out - want)
    scale = np.abs(want).max()
    msg = f"mode={mode} op={op} M={M} K={K} N={N}: max|err|={err.max():.3e} rel-to-max={err.max() / scale:.3e} mean|err|={err.mean():.3e}"
    if op == 2:
        e2 = np.abs(out2[:N] - dY64.sum(0)).max() / np.abs(dY64.sum(0)).max()
        msg += f" colsum rel={e2:.2e}"
    if err.max() / scale > 1e-3:
        bad = np.argwhere(err > 1e-3 * scale)
        msg += f"\n   BAD entries: {len(bad)} of {err.size}; first {bad[:6].tolist()}; rows bad {np.unique(bad[:,0])[:12].tolist()} cols bad {np.unique(bad[:,1])[:12].tolist()}"
        msg += f"\n   got[0,:6]={out[0,:6]} want[0,:6]={want[0,:6]}"
    print(msg, flush=True)
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "time":
        args = [int(x) for x in sys.argv[1:6]]
        run_case(*args, ints=len(sys.argv) > 6)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        import ppo_b200 as P
        ctx = P.Context(0)
        for which, c in (("tc1_fwd", 3), ("tc1_dgrad", 3), ("tc1_wgrad", 3)):
            ms, fl = ctx.bench_kernel(which, 1 << 20, 512, 512, c, 3, True)
            print(f"{which} passes={c}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s (fp32-equivalent)", flush=True)
        for which, c in (("tc1_fwd", 3), ("tc1_wgrad", 3)):
            ms, fl = ctx.bench_kernel(which, 1 << 20, 64, 512, c, 3, True)
            print(f"{which} K=64 passes={c}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        sys.exit(0)
    cases = [
        (1, 0, 1000, 512, 512), (1, 0, 65536, 512, 512), (1, 0, 300, 72, 128), (1, 1, 1000, 512, 512),
        (1, 2, 4096, 512, 512), (1, 2, 100000, 512, 512), (1, 2, 1000000, 64, 512), (1, 2, 1000, 72, 128),
    ]
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__] + [str(x) for x in c], timeout=90, capture_output=True, text=True)
            print((r.stdout + r.stderr[-600:]).strip() if r.returncode else r.stdout.strip(), flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT case {c}", flush=True)
