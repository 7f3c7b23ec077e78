"""Diagnostic (not a test): gradient error of each GEMM engine against the fp64 oracle for the smooth (slope = 1) C3-width
policy, plus dense-op errors.  Usage on the GPU box: python tests/tool_engine_accuracy.py [modes...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ppo_b200 as P
from oracle import ppo_oracle as O
import test_gpu_tc as T

modes = [int(a) for a in sys.argv[1:]] or [0, 1, 3]
ctx = P.Context(0)
rng = np.random.default_rng(1)
M, K, N = 8192, 512, 512
X = rng.normal(size=(M, K)).astype(np.float32); W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
b = rng.normal(size=N).astype(np.float32); dY = rng.normal(size=(M, N)).astype(np.float32)
for op in (0, 1, 2):
    want = T.truth(op, X, W, b, dY)
    print(f"dense op{op}:", " ".join(f"mode{m} {np.max(np.abs(T.dense(ctx, m, op, X, W, b, dY)[0] - want)) / np.max(np.abs(want)):.2e}" for m in modes), flush=True)
for seed in (73, 77, 5):
    cfg, rng, feat, mask, act, Wt, bt, adv = T._c3_case(512, seed)
    nb = feat.shape[0]
    o64 = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa); o64.slope = 1.0
    o64.W, o64.b = [w.astype(np.float64) for w in Wt], [x.astype(np.float64) for x in bt]
    probs = O.batch_action_probabilities(o64, feat.astype(np.float64), mask.astype(np.float64))
    old = (probs[np.arange(nb), act - 1] * np.exp(rng.normal(0, 0.1, nb))).clip(1e-6, 1).astype(np.float32)
    pl, ew, dW, db = O.policy_gradient(o64, feat.astype(np.float64), mask.astype(np.float64), act, old.astype(np.float64),
                                       adv.astype(np.float64), 0.05, 0.01)
    want = T._flat(dW, db)
    lin = P.get_linear_action_index(act, cfg.A)
    for mode in modes:
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=Wt, biases=bt, leaky_slope=1.0)
        pol.set_gemm_mode(mode)
        _, _, g = P.step_batch_(pol, None, P.StateData(feat, mask), lin, old, adv, 0.05, 0.01, return_grads=True)
        pol.close()
        errs = T._tensor_errors(cfg, g, want)
        sgn = np.sum((g - want) * want) / np.sum(want * want)
        print(f"seed {seed} mode{mode}: max tensor err {max(errs):.2e}  [{' '.join(f'{e:.1e}' for e in errs)}]  scale bias {sgn:+.2e}", flush=True)
