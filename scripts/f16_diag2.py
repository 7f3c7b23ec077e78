"""diagnostic: (a) dense-op error of both tensor-core engines vs fp64, (b) how often one leakyrelu' branch flip
(visible as a rank-1 change of a gradient tensor) separates each engine from the fp32 FFMA engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ppo_b200 as P
import test_gpu_tc as T
ctx = P.Context(0)
rng = np.random.default_rng(1)
M, K, N = 8192, 512, 512
X = rng.normal(size=(M, K)).astype(np.float32); W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
b = rng.normal(size=N).astype(np.float32); dY = rng.normal(size=(M, N)).astype(np.float32)
for op in (0, 1, 2):
    want = T.truth(op, X, W, b, dY)
    line = f"op{op}:"
    for mode in (0, 1, 3):
        got, _ = T.dense(ctx, mode, op, X, W, b, dY)
        line += f" mode{mode} {np.max(np.abs(got - want)) / np.max(np.abs(want)):.2e}"
    print(line, flush=True)
for seed in range(70, 82):
    cfg, rng, feat, mask, act, Wt, bt, adv = T._c3_case(512, seed)
    nb = feat.shape[0]
    pol0 = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=Wt, biases=bt, leaky_slope=0.01)
    probs = P.batch_action_probabilities(pol0, P.StateData(feat, mask)); pol0.close()
    old = (probs[np.arange(nb), act - 1].astype(np.float64) * np.exp(rng.normal(0, 0.1, nb))).clip(1e-6, 1).astype(np.float32)
    lin = P.get_linear_action_index(act, cfg.A)
    res = {}
    for mode in (0, 1, 3):
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=Wt, biases=bt, leaky_slope=0.01)
        pol.set_gemm_mode(mode)
        _, _, g = P.step_batch_(pol, None, P.StateData(feat, mask), lin, old, adv, 0.05, 0.01, return_grads=True)
        res[mode] = g.astype(np.float64); pol.close()
    line = f"seed {seed}:"
    for mode in (1, 3):
        errs = T._tensor_errors(cfg, res[mode], res[0])
        line += f" mode{mode} max tensor err {max(errs):.2e}"
    print(line, flush=True)
