"""Whole-network gradient parity of every GEMM engine against the Float64 oracle at the REFERENCE's leakyrelu slope
(0.01), many seeds, C3 widths — and of the default engine (fp16-split tcgen05, what bench.py times) on one full
C3 minibatch (65 536 samples, 2^20 tokens).

leakyrelu' is discontinuous at 0 (ASSUMED NNlib.leakyrelu, reference test/policy.jl:11-15), so an fp32 evaluation and
the fp64 oracle can take different branches for a pre-activation within rounding of zero, and ONE such flip is a
rank-1 change of every gradient tensor below it.  The comparison is therefore made well posed instead of being
seed-picked: the device exports the branches its backward pass actually applied (ppo_policy_read_gates), the oracle
evaluates its Float64 pullback with exactly those branches (oracle.policy_gradient(gates=...)), and a separate
assertion bounds every disagreement between the device's branches and the oracle's own ones to pre-activations below
1e-5 of that layer's largest one (where either branch is a correct Float32 answer).

Tolerances (north_star): loss scalars 1e-5 relative; every parameter tensor of the gradient within 1e-5 of
that tensor's max-abs."""
import numpy as np
import pytest

import ppo_b200 as P
from ppo_b200 import synthetic as S
from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
TC, SIMT, F16 = P.GEMM_TF32X3_TC, P.GEMM_FP32_SIMT, P.GEMM_F16X3_TC
EPS, W_ENT = 0.05, 0.01


def _flat(W, b):
    return np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])


def _tensor_errors(dims, got, want):
    """max-abs error of every parameter tensor relative to that tensor's max-abs (Flux.params order)."""
    out, off = [], 0
    for i, o in zip(dims[:-1], dims[1:]):
        for size in (i * o, o):
            w = want[off:off + size]
            out.append(float(np.max(np.abs(got[off:off + size] - w)) / (np.max(np.abs(w)) + 1e-30)))
            off += size
    return out


def _case(cfg, nb, seed, trained_like=True):
    """a seeded minibatch at cfg's shapes: small-int features, quad-group masks, Glorot weights drawn from the seed,
    biases ~ N(0, 0.05) (a policy that has taken a few steps), integer advantages, old probabilities = the policy's
    own x exp(N(0, 0.1)) so that both clip branches occur."""
    rng = np.random.default_rng(1000 + seed)
    feat = rng.integers(-3, 9, (nb, cfg.nhe, cfg.nf)).astype(np.float32)
    mask = S.make_masks(rng, nb, cfg.nhe, cfg.apa)
    act = S.make_actions(rng, mask)
    W, b = [], []
    d = cfg.dims
    for i, o in zip(d[:-1], d[1:]):
        lim = np.sqrt(6.0 / (i + o))
        W.append(rng.uniform(-lim, lim, size=(i, o)).astype(np.float32))
        b.append((rng.normal(0, 0.05, o) if trained_like else np.zeros(o)).astype(np.float32))
    adv = rng.integers(-4, 5, nb).astype(np.float32)
    return rng, feat, mask, act, W, b, adv


def _oracle64(cfg, W, b, slope=0.01):
    o = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    o.slope = slope
    o.W, o.b = [w.astype(np.float64) for w in W], [x.astype(np.float64) for x in b]
    return o


def _old_probs(o64, rng, feat, mask, act, chunk=8192):
    nb = feat.shape[0]
    sel = np.empty(nb)
    for s in range(0, nb, chunk):
        e = min(nb, s + chunk)
        pr = O.batch_action_probabilities(o64, feat[s:e].astype(np.float64), mask[s:e].astype(np.float64))
        sel[s:e] = pr[np.arange(e - s), act[s:e] - 1]
    return (sel * np.exp(rng.normal(0, 0.1, nb))).clip(1e-6, 1).astype(np.float32)


SKIPPED = 255      # PPO_GATE_SKIPPED: a token the compacted MLP did not run


def _check_gate_disagreements(acts, gates, slope, tag, mask=None, apa=None):
    """every device branch that differs from the oracle's own sits at a ~0 pre-activation; the tokens a compacted pass
    skipped are exactly the tokens all of whose actions are masked (their incoming gradient is an exact zero)"""
    flips = []
    for l, g in gates.items():
        a = acts[l]
        z = np.where(a > 0, a, a / slope)               # Float64 pre-activation of hidden layer l
        skipped = g == SKIPPED
        if mask is not None:
            dead = np.all(np.isneginf(mask.reshape(-1, apa)), axis=1)
            assert np.array_equal(skipped.all(axis=1), skipped.any(axis=1)), (tag, l)
            assert np.array_equal(skipped.all(axis=1), dead) or not skipped.any(), (tag, l)
        diff = (g.astype(bool) != (a > 0)) & ~skipped
        n = int(diff.sum())
        flips.append(n)
        if n:
            worst = float(np.max(np.abs(z[diff])) / np.max(np.abs(z)))
            assert worst <= 1e-5, (tag, l, n, worst)
    return flips


def _device_vs_oracle(ctx, cfg, mode, seed, nb, compact=True):
    rng, feat, mask, act, W, b, adv = _case(cfg, nb, seed)
    o64 = _oracle64(cfg, W, b)
    old = _old_probs(o64, rng, feat, mask, act)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, gemm_mode=mode)
    assert pol.gemm_mode == mode
    pol.set_token_compaction(compact)
    lin = P.get_linear_action_index(act, cfg.A)
    gp, ge, grads = P.step_batch_(pol, None, P.StateData(feat, mask), lin, old, adv, EPS, W_ENT, return_grads=True)
    M = nb * cfg.nhe
    gates = {l: pol.read_gates(l, M) for l in range(1, cfg.L + 1)}
    live = int((~np.all(np.isneginf(mask.reshape(-1, cfg.apa)), axis=1)).sum())
    assert pol.active_tokens() == (live if (compact and mode == F16) else -1), (pol.active_tokens(), live, M)
    assert (gates[1] == SKIPPED).any() == (compact and mode == F16 and live < M)
    pol.close()
    pl, ew, dW, db, acts = O.policy_gradient(o64, feat.astype(np.float64), mask.astype(np.float64), act,
                                             old.astype(np.float64), adv.astype(np.float64), EPS, W_ENT,
                                             gates=gates, return_acts=True)
    flips = _check_gate_disagreements(acts, gates, 0.01, (mode, seed), mask, cfg.apa)
    errs = _tensor_errors(cfg.dims, grads, _flat(dW, db))
    return (gp, ge), (pl, ew), errs, flips


@pytest.mark.parametrize("compact", [True, False], ids=["compact", "dense"])
@pytest.mark.parametrize("seed", range(24))
def test_f16_engine_gradient_vs_fp64_oracle(ctx, seed, compact):
    """the engine every bench number rides on, 24 seeds, MLP 3x512 on 64 features x 16 tokens, 512 samples; with token
    compaction (the default: ~19 % of these tokens are fully masked and skipped) and with every token run"""
    if not compact and seed >= 6:
        pytest.skip("dense evaluation: 6 seeds")
    cfg = S.CONFIGS["c3"]
    (gp, ge), (pl, ew), errs, flips = _device_vs_oracle(ctx, cfg, F16, seed, 512, compact)
    assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7, (seed, gp, pl)
    assert abs(ge - ew) <= 1e-5 * abs(ew) + 1e-8, (seed, ge, ew)
    assert max(errs) <= 1e-5, (seed, errs, flips)


@pytest.mark.parametrize("mode", [SIMT, TC], ids=["ffma", "tf32x3"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_other_engines_gradient_vs_fp64_oracle(ctx, mode, seed):
    cfg = S.CONFIGS["c3"]
    (gp, ge), (pl, ew), errs, flips = _device_vs_oracle(ctx, cfg, mode, seed, 256)
    assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7, (mode, seed, gp, pl)
    assert abs(ge - ew) <= 1e-5 * abs(ew) + 1e-8
    assert max(errs) <= 1e-5, (mode, seed, errs, flips)


@pytest.mark.parametrize("mode", [SIMT, TC, F16], ids=["ffma", "tf32x3", "f16x3"])
def test_smooth_network_needs_no_gates(ctx, mode):
    """slope = 1: leakyrelu is the identity, the loss is smooth in the weights, and every engine must match the plain
    Float64 oracle (no gate hand-over) to 1e-5 of every parameter tensor's max-abs"""
    cfg = S.CONFIGS["c3"]
    rng, feat, mask, act, W, b, adv = _case(cfg, 512, 77)
    o64 = _oracle64(cfg, W, b, slope=1.0)
    old = _old_probs(o64, rng, feat, mask, act)
    pl, ew, dW, db = O.policy_gradient(o64, feat.astype(np.float64), mask.astype(np.float64), act,
                                       old.astype(np.float64), adv.astype(np.float64), EPS, W_ENT)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, leaky_slope=1.0, gemm_mode=mode)
    gp, ge, grads = P.step_batch_(pol, None, P.StateData(feat, mask), P.get_linear_action_index(act, cfg.A), old, adv,
                                  EPS, W_ENT, return_grads=True)
    pol.close()
    assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7
    errs = _tensor_errors(cfg.dims, grads, _flat(dW, db))
    assert max(errs) <= 1e-5, (mode, errs)


@pytest.mark.parametrize("key,nb", [("c2", 64), ("t2", 256)])
def test_f16_engine_other_shapes_vs_fp64_oracle(ctx, key, nb):
    """Policy(72,128,2,4) on 64 tokens (config C2's shape) and the small golden shape"""
    cfg = S.CONFIGS[key]
    for seed in (0, 1, 2):
        (gp, ge), (pl, ew), errs, flips = _device_vs_oracle(ctx, cfg, F16, seed, nb)
        assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7
        assert max(errs) <= 1e-5, (key, seed, errs, flips)


def test_f16_engine_full_c3_minibatch_vs_fp64_oracle(ctx):
    """ONE FULL C3 minibatch — 65 536 samples = 2^20 token rows, the size bench.py runs (the wgrad contraction runs over
    all 2^20 rows, split over the SMs) — against the Float64 oracle evaluated in sample chunks with the device's gates"""
    cfg = S.CONFIGS["c3"]
    nb = cfg.B
    rng, feat, mask, act, W, b, adv = _case(cfg, nb, 4242)
    o64 = _oracle64(cfg, W, b)
    old = _old_probs(o64, rng, feat, mask, act)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    assert pol.gemm_mode == F16, "PPO_GEMM_AUTO must pick the fp16-split engine for C3"
    gp, ge, grads = P.step_batch_(pol, None, P.StateData(feat, mask), P.get_linear_action_index(act, cfg.A), old, adv,
                                  EPS, W_ENT, return_grads=True)
    M = nb * cfg.nhe
    gates = {l: pol.read_gates(l, M) for l in range(1, cfg.L + 1)}
    live = int((~np.all(np.isneginf(mask.reshape(-1, cfg.apa)), axis=1)).sum())
    assert pol.active_tokens() == live and live < M          # token compaction is on by default
    pol.close()
    chunk = 4096
    want = np.zeros(cfg.num_params)
    pl = ew = 0.0
    flips = np.zeros(cfg.L, np.int64)
    for s in range(0, nb, chunk):
        e = min(nb, s + chunk)
        tok = slice(s * cfg.nhe, e * cfg.nhe)
        g = {l: gates[l][tok] for l in gates}
        p_, e_, dW, db, acts = O.policy_gradient(o64, feat[s:e].astype(np.float64), mask[s:e].astype(np.float64), act[s:e],
                                                 old[s:e].astype(np.float64), adv[s:e].astype(np.float64), EPS, W_ENT,
                                                 gates=g, nb_total=nb, return_acts=True)
        flips += np.array(_check_gate_disagreements(acts, g, 0.01, ("full", s), mask[s:e], cfg.apa))
        want += _flat(dW, db)
        pl += p_ * (e - s) / nb
        ew += e_ * (e - s) / nb
    assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7, (gp, pl)
    assert abs(ge - ew) <= 1e-5 * abs(ew) + 1e-8, (ge, ew)
    errs = _tensor_errors(cfg.dims, grads, want)
    assert max(errs) <= 1e-5, (errs, flips.tolist())


def test_deep_policies_leave_the_f16_contract(ctx):
    """The fp16-split engine's per-tensor exponents come from an a-priori bound that loosens by ~2^4.8 per 512-wide layer;
    beyond 4 hidden layers the scaled activations would sink into fp16's subnormal range and the 1e-5 bound would be
    lost silently, so deeper policies are outside its contract: explicit selection fails loudly, PPO_GEMM_AUTO falls
    back to the tf32 engine; the deepest policy inside the contract is checked against the oracle."""
    deep = S.Config("deep6", 95, 256, 64, 4, 4, 256, 6, 64)
    W, b = S.make_weights(deep)
    pol = P.Policy(deep.nf, deep.H, deep.L, deep.apa, ctx, weights=W, biases=b)
    assert pol.gemm_mode == TC
    with pytest.raises(P.PPOError):
        pol.set_gemm_mode(F16)
    pol.close()
    ok = S.Config("deep4", 96, 256, 64, 4, 4, 512, 4, 64)
    for seed in (0, 1):
        (gp, ge), (pl, ew), errs, flips = _device_vs_oracle(ctx, ok, F16, seed, 128)
        assert abs(gp - pl) <= 1e-5 * abs(pl) + 1e-7
        assert max(errs) <= 1e-5, (seed, errs, flips)
