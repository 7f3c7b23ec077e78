"""Launch-bound regime: config C2 with the reference-like minibatch of 32 (and 4096), graph replay on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppo_b200 as P
from ppo_b200 import synthetic as S
cfg = S.CONFIGS["c2"]
ctx = P.Context(0)
data = S.make_buffer(cfg)
W, b = S.make_weights(cfg)
old = np.full(cfg.N, 1.0 / cfg.A, np.float32)
for gemm in (P.GEMM_FP32_SIMT, P.GEMM_TF32X3_TC):
    for B in (32, 4096):
        for ng in ("1", "0"):
            os.environ["PPO_B200_NO_GRAPH"] = ng
            buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
            buf.append(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
            P.compute_state_value_(buf, 1.0)
            pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
            try:
                pol.set_gemm_mode(gemm)
            except P.PPOError as e:
                print("gemm mode", gemm, "unsupported:", str(e)[:80]); pol.close(); buf.close(); continue
            opt = P.Adam(1e-4)
            ds = P.construct_dataset(buf)
            P.step_epoch_(pol, opt, ds, 0.05, B, 0.01, seed=1)
            ctx.sync(); t0 = time.perf_counter()
            res = P.step_epoch_(pol, opt, ds, 0.05, B, 0.01, seed=2)
            ctx.sync(); dt = time.perf_counter() - t0
            print(f"gemm={gemm} B={B:5d} graph={'off' if ng == '1' else 'on '}: {dt*1e3:9.2f} ms/epoch  {cfg.N/dt/1e3:9.1f} k samples/s  loss={res[0]:.6f}", flush=True)
            pol.close(); buf.close()
