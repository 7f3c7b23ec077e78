"""diagnostic: per-tensor gradient differences between engines (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ppo_b200 as P
from ppo_b200 import synthetic as S
import test_gpu_tc as T
ctx = P.Context(0)
for slope in (1.0, 0.01):
    cfg, rng, feat, mask, act, W, b, adv = T._c3_case(512, 77)
    nb = feat.shape[0]
    pol0 = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, leaky_slope=slope)
    probs = P.batch_action_probabilities(pol0, P.StateData(feat, mask))
    pol0.close()
    old = (probs[np.arange(nb), act - 1].astype(np.float64) * np.exp(rng.normal(0, 0.1, nb))).clip(1e-6, 1).astype(np.float32)
    print("old min/median", old.min(), np.median(old), " adv/old max", np.max(np.abs(adv) / old))
    lin = P.get_linear_action_index(act, cfg.A)
    res = {}
    for mode in (0, 1, 3):
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, leaky_slope=slope)
        pol.set_gemm_mode(mode)
        gp, ge, grads = P.step_batch_(pol, None, P.StateData(feat, mask), lin, old, adv, 0.05, 0.01, return_grads=True)
        res[mode] = grads.astype(np.float64)
        pol.close()
    d = cfg.dims
    off = 0
    print("slope", slope, "global max", np.max(np.abs(res[0])))
    for li, (i, o) in enumerate(zip(d[:-1], d[1:])):
        for nm, size in (("W", i * o), ("b", o)):
            ref = res[0][off:off + size]
            line = f"  {nm}{li} max|ref|={np.max(np.abs(ref)):.3e}"
            for mode in (1, 3):
                g = res[mode][off:off + size]
                line += f" | mode{mode}: maxerr={np.max(np.abs(g - ref)):.3e} frac>1e-5*tmax={np.mean(np.abs(g - ref) > 1e-5 * np.max(np.abs(ref))):.4f}"
            print(line)
            off += size
