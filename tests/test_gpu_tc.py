"""tcgen05 engines (PPO_GEMM_TF32X3_TC: 3xTF32 split; PPO_GEMM_F16X3_TC: scaled fp16 hi/lo split): the
error-compensated GEMMs against fp64 numpy and the whole update against the oracle, at the stated fp32
tolerance (1e-5 of the tensor's max-abs).  Whole-network gradients against the Float64 oracle (many seeds, the reference's
leakyrelu slope, full C3 minibatch): tests/test_gpu_engine_parity.py."""
import ctypes as C
import os

import numpy as np
import pytest

import ppo_b200 as P
from ppo_b200 import _lib
from ppo_b200 import synthetic as S
from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TC, SIMT, F16 = P.GEMM_TF32X3_TC, P.GEMM_FP32_SIMT, P.GEMM_F16X3_TC


def dense(ctx, mode, op, X, W, b, dY, slope=0.01):
    M, K = X.shape
    N = W.shape[1]
    out = np.empty({0: (M, N), 1: (M, K), 2: (K, N)}[op], np.float32)
    out2 = np.zeros(max(K, N), np.float32)
    _lib.check(_lib.load().ppo_dense_op(ctx.handle, mode, op, M, K, N, _lib.ptr(X, C.c_float), _lib.ptr(W, C.c_float),
                                        _lib.ptr(b, C.c_float), _lib.ptr(dY, C.c_float), slope,
                                        _lib.ptr(out, C.c_float), _lib.ptr(out2, C.c_float)))
    return out, out2


def truth(op, X, W, b, dY):
    X, W, dY = X.astype(np.float64), W.astype(np.float64), dY.astype(np.float64)
    if op == 0:
        z = X @ W + b
        return np.where(z > 0, z, 0.01 * z)
    if op == 1:
        g = dY @ W.T
        return np.where(X > 0, g, 0.01 * g)
    return X.T @ dY


def _outside_contract(mode, op, K, N):
    """shapes the tensor-core engines refuse (set_gemm_mode refuses such policies loudly)"""
    if mode == TC:
        return (op in (0, 2) and (K % 4 or N % 32)) or (op == 1 and (K % 16 or N % 4))
    if mode == F16:
        return (op in (0, 2) and (K % 8 or N % 32)) or (op == 1 and (K % 32 or N % 8))
    return False


@pytest.mark.parametrize("mode", [SIMT, TC, F16])
@pytest.mark.parametrize("M,K,N", [(128, 32, 128), (300, 72, 128), (1000, 64, 512), (4096, 512, 512), (257, 128, 48),
                                   (5000, 512, 256), (40000, 512, 512)])
def test_dense_ops_vs_fp64(ctx, mode, M, K, N):
    rng = np.random.default_rng(M + K + N)
    X = rng.normal(size=(M, K)).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    dY = rng.normal(size=(M, N)).astype(np.float32)
    for op in (0, 1, 2):
        if _outside_contract(mode, op, K, N):
            continue
        got, got2 = dense(ctx, mode, op, X, W, b, dY)
        want = truth(op, X, W, b, dY)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        assert err <= 1e-5, (mode, op, M, K, N, err)
        if op == 2 and mode != F16:     # (the fp16 engine takes the bias gradient from the kernel that produced dY)
            cs = dY.astype(np.float64).sum(0)
            assert np.max(np.abs(got2[:N] - cs)) <= 1e-5 * np.max(np.abs(cs))
        if op == 1 and mode in (TC, F16):      # column sums fused into the dgrad epilogue (bias gradient of the layer below)
            cs = want.sum(0)
            assert np.max(np.abs(got2[:K] - cs)) <= 1e-5 * np.max(np.abs(cs)) + 1e-6


@pytest.mark.parametrize("M,K,N", [(4096, 512, 512), (2048, 1024, 256)])
def test_f16_engine_all_positive_operands(ctx, M, K, N):
    """Operands of one sign are the worst case for the tensor core's round-toward-zero accumulation (every add of a chain
    loses in the same direction: up to ~n/2 * 2^-24 of the sum over an n-MMA chain).  The compensated fold
    (KK16Params::rz_comp) is a statistical correction calibrated on mixed-sign data; the 1e-5 bound must hold with a wide
    margin here too, so the constant is a refinement, not what parity rests on."""
    rng = np.random.default_rng(M + K)
    X = np.abs(rng.normal(size=(M, K))).astype(np.float32)
    W = np.abs(rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = np.abs(rng.normal(size=N)).astype(np.float32)
    dY = np.abs(rng.normal(size=(M, N))).astype(np.float32)
    for op in (0, 1, 2):
        got, _ = dense(ctx, F16, op, X, W, b, dY)
        want = truth(op, X, W, b, dY)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        assert err <= 3e-6, (op, err)


def _flat(W, b):
    return np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])


def test_tc_mode_refuses_unsupported_shapes_loudly(ctx):
    cfg = S.CONFIGS["t0"]          # hidden width 16: outside the tensor-core engine's contract
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    with pytest.raises(P.PPOError):
        pol.set_gemm_mode(TC)
    with pytest.raises(P.PPOError):
        pol.set_gemm_mode(P.GEMM_BF16_TC)
    with pytest.raises(P.PPOError):
        pol.set_gemm_mode(F16)
    pol.close()


@pytest.mark.parametrize("scale_x,scale_g", [(1.0, 1.0), (3e4, 1e-6), (1e-5, 1e5)])
def test_f16_engine_scales_follow_the_data(ctx, scale_x, scale_g):
    """fp16 has a 5-bit exponent: the engine's per-tensor power-of-two scales must keep the 1e-5 bound for
    operands far outside fp16's own range (features ~3e4, gradients ~1e-6 and the reverse)."""
    M, K, N = 2048, 128, 256
    rng = np.random.default_rng(5)
    X = (rng.normal(size=(M, K)) * scale_x).astype(np.float32)
    W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = (rng.normal(size=N) * scale_x).astype(np.float32)
    dY = (rng.normal(size=(M, N)) * scale_g).astype(np.float32)
    for op in (0, 1, 2):
        got, _ = dense(ctx, F16, op, X, W, b, dY)
        want = truth(op, X, W, b, dY)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        assert err <= 1e-5, (op, err)


def test_f16_engine_epoch_matches_fp32_engine(ctx):
    """one PPO epoch at the C2 policy shape (Policy(72,128,2,4)): weights after Adam, fp16-split vs fp32 FFMA engine"""
    cfg = S.Config("c2-mini", 92, 4096, 72, 16, 4, 128, 2, 512)
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    out = {}
    for mode in (SIMT, F16):
        buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
        old = np.full(cfg.N, 0.05, np.float32)
        buf.append(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
        P.compute_state_value_(buf, 1.0)
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
        pol.set_gemm_mode(mode)
        ds = P.construct_dataset(buf)
        perm = np.random.default_rng(3).permutation(cfg.N) + 1
        mp_, me_ = P.step_epoch_(pol, P.Adam(1e-4), ds, 0.05, cfg.B, 0.01, perm=perm)
        Wd, bd = pol.weights()
        out[mode] = (mp_, me_, _flat(Wd, bd))
        pol.close(); buf.close()
    assert abs(out[F16][0] - out[SIMT][0]) <= 1e-5 * abs(out[SIMT][0]) + 1e-6
    assert abs(out[F16][1] - out[SIMT][1]) <= 1e-5 * abs(out[SIMT][1]) + 1e-7
    # Adam normalises the step (|dw| ~ eta): a sign-level disagreement on a ~0 gradient moves a weight by 2 eta at most
    assert np.max(np.abs(out[F16][2] - out[SIMT][2])) <= 2e-5


@pytest.mark.parametrize("name,key", [("oracle_t1_g1", "t1")])
def test_golden_epoch_tc_mode(ctx, name, key):
    z = np.load(os.path.join(G, name + ".npz"))
    cfg = S.CONFIGS[key]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    buf.append(data["feat"], data["mask"], z["old"], data["action"], data["reward"], data["terminal"])
    P.compute_state_value_(buf, float(z["gamma"]))
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    pol.set_gemm_mode(TC)
    ds = P.construct_dataset(buf)
    buf.set_permutation(z["perm0"] + 1)
    pl, ew, grads = None, None, None
    import ctypes
    p_, e_ = ctypes.c_double(), ctypes.c_double()
    grads = np.empty(pol.num_params, np.float32)
    _lib.check(_lib.load().ppo_step_batch(pol.handle, None, buf.handle, 0, cfg.B, float(z["eps"]), float(z["w_ent"]),
                                          ctypes.byref(p_), ctypes.byref(e_), _lib.ptr(grads, C.c_float)))
    assert abs(p_.value - float(z["ppoloss"])) <= 1e-5 * abs(float(z["ppoloss"])) + 1e-7
    assert np.max(np.abs(grads - z["grads"])) <= 1e-5 * np.max(np.abs(z["grads"])) + 1e-7
    mp_, me_ = P.step_epoch_(pol, P.Adam(float(z["eta"])), ds, float(z["eps"]), cfg.B, float(z["w_ent"]), perm=z["perm0"] + 1)
    assert abs(mp_ - float(z["mean_ppo"])) <= 1e-5 * abs(float(z["mean_ppo"])) + 1e-6
    Wd, bd = pol.weights()
    assert np.max(np.abs(_flat(Wd, bd) - z["flat_after"])) <= 2e-5
    pol.close(); buf.close()


def test_gemm_auto_picks_the_fastest_engine_inside_its_contract(ctx):
    for key, want in (("t0", SIMT), ("t1", TC), ("t2", F16), ("c3", F16), ("c2", F16)):
        cfg = S.CONFIGS[key]
        W, b = S.make_weights(cfg)
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
        assert pol.gemm_mode == want, key          # PPO_GEMM_AUTO is the default of a new policy
        assert pol.set_gemm_mode(SIMT) == SIMT
        assert pol.set_gemm_mode(P.GEMM_AUTO) == want, key
        pol.close()


def _compaction_case(cfg, nb, seed, live_fraction):
    """a minibatch whose tokens are fully masked (all apa actions -Inf) with probability 1 - live_fraction, token 0 of
    every state alive (a state needs one legal action); live tokens carry per-action masks too"""
    rng = np.random.default_rng(seed)
    feat = rng.integers(-3, 9, (nb, cfg.nhe, cfg.nf)).astype(np.float32)
    live = rng.random((nb, cfg.nhe)) < live_fraction
    live[:, 0] = True
    am = np.where(rng.random((nb, cfg.nhe, cfg.apa)) < 0.2, -np.inf, 0.0)
    am[:, 0, 0] = 0.0
    mask = np.where(live[:, :, None], am, -np.inf).astype(np.float32).reshape(nb, cfg.A)
    act = S.make_actions(rng, mask)
    W, b = S.make_weights(cfg)
    b = [(x + rng.normal(0, 0.05, x.shape)).astype(np.float32) for x in b]
    adv = rng.integers(-4, 5, nb).astype(np.float32)
    old = rng.uniform(0.02, 0.9, nb).astype(np.float32)
    return feat, mask, act, W, b, adv, old


@pytest.mark.parametrize("key,nb,live", [("c3", 512, 0.8), ("c3", 512, 0.05), ("c3", 300, 1.0), ("c3", 7, 0.1),
                                         ("c2", 64, 0.3), ("t2", 1000, 0.5)])
def test_token_compaction_matches_the_dense_evaluation(ctx, key, nb, live):
    """Token compaction (fp16-split engine, default on) runs the MLP only on tokens with an unmasked action.  Fully
    masked tokens have probability exactly 0 and dlogits exactly 0, so against the same engine run on every token:
    probabilities and losses are BIT-IDENTICAL (a row's logits do not depend on which tile it sits in), gradients agree
    to summation order (1e-5 of every tensor's max-abs is the parity bound; here they are ~1e-6), and the run count
    equals the number of live tokens -- from a mostly padded minibatch (5 % live, fewer rows than one 128-row tile) to
    one with nothing to skip."""
    cfg = S.CONFIGS[key]
    feat, mask, act, W, b, adv, old = _compaction_case(cfg, nb, 11 + nb, live)
    lin = P.get_linear_action_index(act, cfg.A)
    n_live = int((~np.all(np.isneginf(mask.reshape(-1, cfg.apa)), axis=1)).sum())
    out = {}
    for compact in (False, True):
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, gemm_mode=F16)
        pol.set_token_compaction(compact)
        probs = P.batch_action_probabilities(pol, P.StateData(feat, mask))
        assert pol.active_tokens() == (n_live if compact else -1)
        gp, ge, grads = P.step_batch_(pol, None, P.StateData(feat, mask), lin, old, adv, 0.05, 0.01, return_grads=True)
        out[compact] = (probs, gp, ge, grads)
        pol.close()
    np.testing.assert_array_equal(out[True][0], out[False][0])
    assert np.all(out[True][0][np.isneginf(mask)] == 0.0)
    assert out[True][1] == out[False][1] and out[True][2] == out[False][2]
    off = 0
    for i, o in zip(cfg.dims[:-1], cfg.dims[1:]):
        for size in (i * o, o):
            d, c = out[False][3][off:off + size], out[True][3][off:off + size]
            assert np.max(np.abs(d - c)) <= 2e-6 * np.max(np.abs(d)) + 1e-12, (key, nb, live, off)
            off += size


def test_token_compaction_epoch_with_cuda_graph(ctx):
    """a whole epoch (first minibatch eager, the rest replayed from the CUDA graph, ragged last minibatch): the number
    of live tokens differs from minibatch to minibatch and is only known on the device; compacted and dense epochs end
    in the same weights to Adam's step granularity and the same loss history to 1e-6"""
    cfg = S.Config("c3-mini", 93, 9000, 64, 16, 4, 512, 3, 1024)
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    out = {}
    for compact in (False, True):
        buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
        old = np.full(cfg.N, 0.05, np.float32)
        buf.append(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
        P.compute_state_value_(buf, 1.0)
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, gemm_mode=F16)
        pol.set_token_compaction(compact)
        perm = np.random.default_rng(3).permutation(cfg.N) + 1
        mp_, me_ = P.step_epoch_(pol, P.Adam(1e-4), P.construct_dataset(buf), 0.05, cfg.B, 0.01, perm=perm)
        Wd, bd = pol.weights()
        out[compact] = (mp_, me_, _flat(Wd, bd))
        pol.close(); buf.close()
    assert abs(out[True][0] - out[False][0]) <= 1e-6 * abs(out[False][0]) + 1e-9
    assert abs(out[True][1] - out[False][1]) <= 1e-6 * abs(out[False][1]) + 1e-9
    # Adam normalises every step to ~eta: entries whose gradient is at summation-noise level may step either way
    dw = np.abs(out[True][2] - out[False][2])
    assert np.mean(dw > 2e-6) <= 1e-3 and np.max(dw) <= 2 * 1e-4 * 9, (float(np.max(dw)), float(np.mean(dw > 2e-6)))
