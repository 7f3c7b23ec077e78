"""Parity of the CUDA hot path against the oracle, through the C ABI (pytest -m gpu).

Tolerances:
  permutation, gather, permute!                 bit-exact
  returns, gamma == 1 and integer rewards        bit-exact
  returns otherwise                              |d| <= 1e-6 + 1e-5 |x|  (Float64 carry on both sides;
                                                 2e-5 + 1e-5 |x| for a Float32 discount/carry)
  loss scalars                                   1e-5 relative
  dlogits given identical logits                 1e-5 relative + 1e-8 absolute
  weight gradients / post-Adam weights (fp32)    1e-5 relative of the tensor's max-abs (+1e-7)
"""
import os

import numpy as np
import pytest

import ppo_b200 as P
from ppo_b200 import synthetic as S
from ppo_b200.rollout_buffer import gather_minibatch
from oracle import c_oracle as CO
from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _scan(ctx, r, t, g):
    return P.compute_returns(r, t, g, ctx)


# ---------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 255, 4095, 4096, 4097, 8192, 12289, 100003])
def test_returns_bit_exact_gamma1(ctx, n):
    rng = np.random.default_rng(n)
    r = rng.integers(-4, 5, n).astype(np.float32)
    t = rng.random(n) < 0.08
    assert np.array_equal(_scan(ctx, r, t, 1.0), CO.compute_returns(r, t, 1.0))


def test_returns_reference_known_answers(ctx):
    # output/trajectory.csv:1-7 and the TestEnv run of test/test_rollout_buffer.jl
    assert _scan(ctx, np.ones(6, np.float32), np.array([0, 0, 0, 0, 0, 1], bool), 1.0).tolist() == [6, 5, 4, 3, 2, 1]
    t = np.zeros(100, bool); t[9::10] = True
    assert np.array_equal(_scan(ctx, np.ones(100, np.float32), t, 1.0), np.tile(np.arange(10, 0, -1, dtype=np.float32), 10))


@pytest.mark.parametrize("gamma", [0.99, 0.5, np.float32(0.99)])
@pytest.mark.parametrize("p_term", [0.0, 0.0005, 0.05, 1.0])
def test_returns_discounted(ctx, gamma, p_term):
    rng = np.random.default_rng(11)
    n = 50001
    r = rng.normal(size=n).astype(np.float32)
    t = rng.random(n) < p_term
    got = _scan(ctx, r, t, gamma)
    f32 = isinstance(gamma, np.float32)
    want = CO.compute_returns(r, t, float(gamma), f32)
    # Float64 carry (every in-tree call): 1e-6 abs + 1e-5 rel.  Float32 carry: the reference's own serial
    # Float32 chain accumulates ~eps32*sqrt(horizon)*|v| of rounding, so the absolute part is 2e-5.
    atol = 2e-5 if f32 else 1e-6
    assert np.all(np.abs(got - want) <= atol + 1e-5 * np.abs(want))


def test_returns_long_unterminated_and_single_step(ctx):
    rng = np.random.default_rng(12)
    n = 3 * 4096 + 5
    r = rng.integers(-4, 5, n).astype(np.float32)
    for t in (np.zeros(n, bool), np.ones(n, bool)):          # one 12k-step episode / all single-step
        assert np.array_equal(_scan(ctx, r, t, 1.0), CO.compute_returns(r, t, 1.0))
    t = np.zeros(n, bool); t[4095] = True; t[4096] = True; t[8191] = True     # ends on tile edges
    assert np.array_equal(_scan(ctx, r, t, 1.0), CO.compute_returns(r, t, 1.0))


def test_returns_full_size_1m(ctx):
    cfg = S.CONFIGS["c3"]
    rng = S.rng_for(cfg, 5)
    r = rng.integers(-4, 5, cfg.N).astype(np.float32)
    t = S.make_episode_terminals(rng, cfg.N, 30)
    got = _scan(ctx, r, t, 1.0)
    assert np.array_equal(got, CO.compute_returns(r, t, 1.0))
    # size-independent property: returns at episode ends equal the reward; linearity in the rewards
    assert np.array_equal(got[t], r[t])
    got2 = _scan(ctx, 2 * r, t, 1.0)
    assert np.array_equal(got2, 2 * got)


# ---------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n", [1, 2, 3, 5, 64, 1000, 4097, 65536, 1048576])
def test_device_permutation_bit_exact(ctx, n):
    buf = P.DeviceRollouts(1, 1, 1, n, ctx)
    buf.append(np.zeros((n, 1, 1), np.float32), np.zeros((n, 1), np.float32), np.ones(n, np.float32),
               np.ones(n, np.int64), np.zeros(n, np.float32), np.zeros(n, np.uint8))
    for seed in (0, 7, 2 ** 64 - 1):
        p1 = buf.generate_permutation(seed, want=True)
        assert np.array_equal(p1 - 1, CO.feistel_permutation(n, seed))
    assert np.array_equal(np.sort(p1), np.arange(1, n + 1))
    buf.close()


# ---------------------------------------------------------------------------------------- K4
def _filled(ctx, cfg, normalize=False):
    data = S.make_buffer(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    half = cfg.N // 2
    buf.append(data["feat"][:half], data["mask"][:half], old[:half], data["action"][:half], data["reward"][:half],
               data["terminal"][:half])
    buf.append(data["feat"][half:].astype(np.int64), data["mask"][half:], old[half:], data["action"][half:],
               data["reward"][half:], data["terminal"][half:])         # Int64 features path
    obuf = O.BufferRollouts(cfg.nf, cfg.nhe, cfg.apa)
    obuf.update(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
    return data, old, buf, obuf


@pytest.mark.parametrize("key", ["t0", "t1"])
def test_buffer_roundtrip_and_gather_bit_exact(ctx, key):
    cfg = S.CONFIGS[key]
    data, old, buf, obuf = _filled(ctx, cfg)
    assert len(buf) == cfg.N == len(obuf)
    back = buf.read()
    assert np.array_equal(back["feat"], data["feat"]) and np.array_equal(back["mask"], data["mask"])
    assert np.array_equal(back["selected_actions"], data["action"])
    assert np.array_equal(back["selected_action_probabilities"], old)
    assert np.array_equal(back["rewards"], data["reward"]) and np.array_equal(back["terminal"], data["terminal"])
    ds = P.construct_dataset(buf)
    rng = np.random.default_rng(3)
    # dataset[idx] with an arbitrary 1-based index vector (duplicates allowed), get_batch :117-133
    for nb in (1, 7, cfg.B, cfg.N):
        idx = rng.integers(1, cfg.N + 1, nb)
        got, want = ds[idx], O.get_batch(obuf, idx)
        assert np.array_equal(got["state"].vertex_score, want["state"][0])
        assert np.array_equal(got["state"].action_mask, want["state"][1])
        for k in ("selected_action", "selected_action_probability", "returns"):
            assert np.array_equal(got[k], want[k]), k
        cw = CO.get_batch(obuf.feat, obuf.mask, obuf.selected_actions, obuf.selected_action_probabilities,
                          obuf.rewards, idx)
        assert np.array_equal(got["state"].vertex_score, cw["state"][0])
    s = ds[5]
    assert s["selected_action"] == data["action"][4] and np.array_equal(s["state"].vertex_score, data["feat"][4])
    # minibatches of a host-supplied permutation (the reference's randperm), incl. the ragged last one
    perm1 = rng.permutation(cfg.N) + 1
    buf.set_permutation(perm1)
    start = 0
    while start < cfg.N:
        cnt = min(cfg.B, cfg.N - start)
        got = gather_minibatch(ds, start, cnt)
        want = O.get_batch(obuf, perm1[start:start + cnt])
        assert np.array_equal(got["state"].vertex_score, want["state"][0])
        assert np.array_equal(got["state"].action_mask, want["state"][1])
        assert np.array_equal(got["selected_action"], want["selected_action"])
        assert np.array_equal(got["returns"], want["returns"])
        start += cnt
    buf.close()


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int64])
def test_integer_feature_append_is_exact(ctx, dtype):
    """StateData.vertex_score is a Matrix{Int64} of small integers (test/quad_game_utilities.jl:50-56): the Int64 entry
    point and the narrowed Int8 / Int16 ones (4x / 2x fewer host->device bytes) must land the same Float32 features as the
    Float32 append, also across a growth of the buffer and for row counts that are not multiples of 16."""
    nf, nhe, apa = 9, 3, 2
    rng = np.random.default_rng(3)
    lim = {np.int8: 127, np.int16: 32767, np.int64: 10 ** 6}[dtype]
    buf = P.DeviceRollouts(nf, nhe, apa, 5, ctx)
    ref = P.DeviceRollouts(nf, nhe, apa, 64, ctx)
    for n in (1, 7, 33):
        f = rng.integers(-lim - 1, lim + 1, (n, nhe, nf)).astype(dtype)
        m = np.zeros((n, nhe * apa), np.float32)
        args = (m, np.full(n, 0.5, np.float32), np.ones(n, np.int64), np.arange(n, dtype=np.float32), np.zeros(n, bool))
        buf.append(f, *args)
        ref.append(f.astype(np.float32), *args)
    a, b = buf.read(), ref.read()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    buf.close(); ref.close()


def test_packed_append_equals_the_float_mask_append(ctx):
    """ppo_buffer_append_packed (Int8 features + one bit per action) fills the buffer with exactly the bytes of the
    Float32-mask append, also when a second append starts at an offset that is not a multiple of 64 actions"""
    rng = np.random.default_rng(8)
    nf, nhe, apa = 8, 3, 4                          # A = 12
    out = {}
    for packed in (False, True):
        buf = P.DeviceRollouts(nf, nhe, apa, 64, ctx)
        r = np.random.default_rng(8)
        for n in (37, 5, 100):
            f = r.integers(-3, 9, (n, nhe, nf)).astype(np.int8)
            m = np.where(r.random((n, nhe * apa)) < 0.4, -np.inf, 0.0).astype(np.float32)
            m[:, 0] = 0.0
            a = S.make_actions(r, m)
            args = (r.random(n).astype(np.float32), a, r.integers(-4, 5, n).astype(np.float32), r.random(n) < 0.1)
            buf.append(f, P.pack_action_mask(m) if packed else m, *args)
        out[packed] = buf.read()
        buf.close()
    for k in out[False]:
        assert np.array_equal(out[False][k], out[True][k]), k
    assert np.isneginf(out[True]["mask"]).sum() > 0 and set(np.unique(out[True]["mask"])) == {-np.inf, 0.0}


def test_gather_variants_agree_bit_exact(ctx):
    # LDG/STG path (variant 0) and TMA bulk-copy ring (variant 1, cp.async.bulk) against the oracle
    import ctypes as C
    from ppo_b200 import _lib
    lib = _lib.load()
    for key, start, cnt in (("t1", 3, 777), ("t1", 0, 1000), ("t0", 250, 7), ("t1", 999, 1)):
        cfg = S.CONFIGS[key]
        data, old, buf, obuf = _filled(ctx, cfg)
        perm1 = np.random.default_rng(5).permutation(cfg.N) + 1
        buf.set_permutation(perm1)
        want = O.get_batch(obuf, perm1[start:start + cnt])
        for variant in (0, 1):
            _lib.check(lib.ppo_gather_device(buf.handle, start, cnt, variant))
            feat = np.full((cnt, cfg.nhe, cfg.nf), np.nan, np.float32)
            mask = np.full((cnt, cfg.A), np.nan, np.float32)
            act = np.empty(cnt, np.int64); prob = np.empty(cnt, np.float32); ret = np.empty(cnt, np.float32)
            _lib.check(lib.ppo_batch_read(buf.handle, cnt, _lib.ptr(feat, C.c_float), _lib.ptr(mask, C.c_float),
                                          _lib.ptr(act, C.c_int64), _lib.ptr(prob, C.c_float), _lib.ptr(ret, C.c_float)))
            assert np.array_equal(feat, want["state"][0]), variant
            assert np.array_equal(mask, want["state"][1]), variant
            assert np.array_equal(act, want["selected_action"]) and np.array_equal(prob, want["selected_action_probability"])
            assert np.array_equal(ret, want["returns"])
        buf.close()


def test_permute_and_shuffle(ctx):
    cfg = S.CONFIGS["t1"]
    data, old, buf, obuf = _filled(ctx, cfg)
    idx1 = np.random.default_rng(9).permutation(cfg.N) + 1
    P.permute_(buf, idx1)
    obuf.permute(idx1)
    back = buf.read()
    assert np.array_equal(back["feat"], obuf.feat) and np.array_equal(back["mask"], obuf.mask)
    assert np.array_equal(back["selected_actions"], obuf.selected_actions)
    assert np.array_equal(back["rewards"], obuf.rewards) and np.array_equal(back["terminal"], obuf.terminal)
    P.shuffle_(buf, seed=21)
    obuf.permute(O.feistel_permutation(cfg.N, 21) + 1)
    assert np.array_equal(buf.read()["feat"], obuf.feat)
    with pytest.raises(P.PPOError):
        P.permute_(buf, idx1[:-1])          # @assert length(idx) == length(rollouts)
    buf.close()


# ---------------------------------------------------------------------------------------- K6
@pytest.mark.parametrize("A", [4, 12, 16, 64, 100, 256, 512, 1000])
@pytest.mark.parametrize("nb", [1, 37, 1024])
def test_loss_and_dlogits(ctx, A, nb):
    rng = np.random.default_rng(A * 1000 + nb)
    logits = rng.normal(0, 2, (nb, A)).astype(np.float32)
    mask = np.where(rng.random((nb, A)) < 0.3, -np.inf, 0.0).astype(np.float32)
    mask[:, 0] = 0
    act = np.array([rng.choice(np.flatnonzero(mask[b] == 0)) + 1 for b in range(nb)], np.int64)
    old = rng.uniform(0.01, 1.0, nb).astype(np.float32)
    adv = rng.integers(-4, 5, nb).astype(np.float32)
    for eps, w in ((0.05, 0.01), (0.2, 0.0)):
        pl, el, dl = P.ppo_loss_with_entropy_from_logits(ctx, logits, mask, act, old, adv, eps, w, want_grad=True)
        wp, we, p, wdz = O.loss_grad_logits(logits, mask, act, old, adv, eps, w)
        # fp64 truth for the tolerance reference
        tp, te, _, tdz = O.loss_grad_logits(logits.astype(np.float64), mask.astype(np.float64), act,
                                            old.astype(np.float64), adv.astype(np.float64), eps, w)
        assert abs(pl - wp) <= 1e-5 * abs(wp) + 1e-7
        assert abs(el - float(we)) <= 1e-5 * abs(float(we)) + 1e-7
        assert abs(pl - tp) <= 1e-5 * abs(tp) + 1e-6
        scale = np.max(np.abs(tdz)) + 1e-30
        assert np.max(np.abs(dl - tdz)) <= 1e-5 * scale + 1e-8
        assert np.all(dl[np.isneginf(mask)] == 0.0)            # masked actions: exactly zero gradient


def test_batch_action_probabilities_masked_exact_zero(ctx):
    cfg = S.CONFIGS["t1"]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    opol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa); opol.W, opol.b = W, b
    st = P.StateData(data["feat"][:300], data["mask"][:300])
    got = P.batch_action_probabilities(pol, st)
    want = O.batch_action_probabilities(opol, st.vertex_score, st.action_mask)
    assert np.all(got[np.isneginf(st.action_mask)] == 0.0)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-8)
    assert np.allclose(got.sum(1), 1.0, atol=1e-5)
    one = P.action_probabilities(pol, P.StateData(data["feat"][0], data["mask"][0]))
    assert np.allclose(one, want[0], rtol=1e-5, atol=1e-8)
    pol.close()


# ------------------------------------------------------------------------------- K5/K7/K8 + loop
def _grad_tol(got, want, rel=1e-5):
    return np.max(np.abs(got - want)) <= rel * np.max(np.abs(want)) + 1e-7


# engines PPO_GEMM_AUTO (the default of a new policy) resolves to: t0 is outside both tensor-core contracts, t1 inside
# the tf32 engine's, t2 inside the fp16-split engine's (the one the benchmark configs run on)
_AUTO_ENGINE = {"t0": P.GEMM_FP32_SIMT, "t1": P.GEMM_TF32X3_TC, "t2": P.GEMM_F16X3_TC}


@pytest.mark.parametrize("name,key", [("oracle_t0_g1", "t0"), ("oracle_t0_g099", "t0"), ("oracle_t1_g1", "t1"),
                                      ("oracle_t2_g1", "t2")])
def test_golden_minibatch_and_epoch(ctx, name, key):
    z = np.load(os.path.join(G, name + ".npz"))
    cfg = S.CONFIGS[key]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    gamma, eps, w_ent, eta = float(z["gamma"]), float(z["eps"]), float(z["w_ent"]), float(z["eta"])
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    buf.append(data["feat"], data["mask"], z["old"], data["action"], data["reward"], data["terminal"])
    P.compute_state_value_(buf, gamma)
    got_ret = buf.rewards
    if gamma == 1.0:
        assert np.array_equal(got_ret, z["returns"])
    else:
        assert np.all(np.abs(got_ret - z["returns"]) <= 1e-6 + 1e-5 * np.abs(z["returns"]))
    assert np.array_equal(buf.generate_permutation(int(z["seed"]), want=True) - 1, z["perm0"])
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    assert pol.gemm_mode == _AUTO_ENGINE[key]
    ds = P.construct_dataset(buf)
    # first minibatch through step_batch! on host arrays (gradient only)
    batch = gather_minibatch(ds, 0, cfg.B)
    lin = P.get_linear_action_index(batch["selected_action"], P.number_of_actions_per_state(batch["state"]))
    adv = P.batch_advantage(batch["state"], batch["returns"])
    pl, ew, grads = P.step_batch_(pol, None, batch["state"], lin, batch["selected_action_probability"], adv, eps,
                                  w_ent, return_grads=True)
    assert abs(pl - float(z["ppoloss"])) <= 1e-5 * abs(float(z["ppoloss"])) + 1e-7
    assert abs(ew - float(z["entw"])) <= 1e-5 * abs(float(z["entw"])) + 1e-8
    assert _grad_tol(grads, z["grads"])
    # the whole epoch with Adam
    opt = P.Optimiser(P.Adam(eta))
    mp_, me_ = P.step_epoch_(pol, opt, ds, eps, cfg.B, w_ent, perm=z["perm0"] + 1)
    assert abs(mp_ - float(z["mean_ppo"])) <= 1e-5 * abs(float(z["mean_ppo"])) + 1e-6
    assert abs(me_ - float(z["mean_ent"])) <= 1e-5 * abs(float(z["mean_ent"])) + 1e-8
    Wd, bd = pol.weights()
    flat = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(Wd, bd)])
    # Adam normalises the step to ~eta per element: compare the parameter DISPLACEMENT
    flat0 = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])
    d_got, d_want = flat - flat0, z["flat_after"] - flat0
    assert np.max(np.abs(d_got - d_want)) <= 0.02 * np.max(np.abs(d_want))
    assert np.max(np.abs(flat - z["flat_after"])) <= 2e-5
    pol.close(); buf.close()


def test_adam_bit_exact_vs_c_oracle(ctx):
    cfg = S.CONFIGS["t0"]
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    opt = P.Adam(1e-3)
    n = pol.num_params
    x = np.concatenate([np.concatenate([w.ravel(), v.ravel()]) for w, v in zip(W, b)]).copy()
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    rng = np.random.default_rng(17)
    b1p, b2p = 0.9, 0.999
    for _ in range(4):
        g = rng.normal(size=n).astype(np.float32) * 0.01
        opt.update_(pol, g)
        CO.adam(x, m, v, g, 1e-3, 0.9, 0.999, 1e-8, b1p, b2p)
        b1p *= 0.9; b2p *= 0.999
        Wd, bd = pol.weights()
        flat = np.concatenate([np.concatenate([w.ravel(), q.ravel()]) for w, q in zip(Wd, bd)])
        assert np.array_equal(flat, x)
    pol.close()


def test_ppo_train_history_and_print(ctx):
    import io
    cfg = S.CONFIGS["t0"]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    buf.append(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
    P.compute_state_value_(buf, 1.0)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    opt = P.Optimiser(P.Adam(1e-4))
    perms = [O.feistel_permutation(cfg.N, 100 + e) + 1 for e in range(3)]
    out = io.StringIO()
    ph, eh, lh = P.ppo_train_(pol, opt, P.construct_dataset(buf), 0.05, cfg.B, 3, 0.01, perms=perms, out=out)
    opol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa); opol.W, opol.b = [w.copy() for w in W], [x.copy() for x in b]
    obuf = O.BufferRollouts(cfg.nf, cfg.nhe, cfg.apa)
    obuf.update(data["feat"], data["mask"], old, data["action"], O.compute_returns(data["reward"], data["terminal"], 1.0),
                data["terminal"])
    lines = []
    wph, weh, wlh = O.ppo_train(opol, O.Adam(1e-4), obuf, 0.05, cfg.B, 3, 0.01, perms, printer=lines.append)
    assert np.allclose(ph, wph, rtol=1e-4, atol=1e-6) and np.allclose(eh, weh, rtol=1e-4, atol=1e-8)
    assert lh == wlh == [1e-4] * 3
    assert out.getvalue().count("EPOCH : ") == 3
    assert out.getvalue().splitlines()[0].startswith("EPOCH : 1 \t PPO LOSS : ")
    with pytest.raises(AssertionError):
        P.step_epoch_(pol, opt, P.construct_dataset(buf), 0.05, cfg.N + 1, 0.01)   # @assert batch_size <= num_data
    pol.close(); buf.close()


def test_c1_testenv_collect_rollouts(ctx):
    # config C1: the reference's TestEnv (test/test_rollout_buffer.jl:4-50) through collect_rollouts!
    class TestEnv:
        def __init__(self, max_steps): self.num_steps, self.max_steps = 0, max_steps
        def state(self): return P.StateData(np.random.rand(1, 9).astype(np.float32), np.zeros(3, np.float32))
        def step_(self, a): self.num_steps += 1
        def reward(self): return 1.0
        def is_terminal(self): return self.num_steps >= self.max_steps
        def reset_(self): self.num_steps = 0

    class FixedPolicy:
        def action_probabilities(self, state): return np.array([1.0, 0.0, 0.0])

    env = TestEnv(10)
    rollouts = P.DeviceRollouts(9, 1, 3, 128, ctx)
    P.collect_rollouts_(rollouts, env, FixedPolicy(), 10, 1.0)
    assert len(rollouts) == 100
    back = rollouts.read()
    assert np.array_equal(back["rewards"], np.tile(np.arange(10, 0, -1, dtype=np.float32), 10))
    assert back["selected_actions"].tolist() == [1] * 100
    assert np.all(back["selected_action_probabilities"] == 1.0)
    assert np.flatnonzero(back["terminal"]).tolist() == list(range(9, 100, 10))
    # ... then one PPO epoch with B = 10 on a Policy(9 -> 3)-style stand-in, against the oracle
    cfg = S.CONFIGS["c1"]
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    opol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa); opol.W, opol.b = [w.copy() for w in W], [x.copy() for x in b]
    perm1 = np.random.default_rng(2).permutation(100) + 1
    got = P.step_epoch_(pol, P.Adam(1e-4), P.construct_dataset(rollouts), 0.05, 10, 0.01, perm=perm1)
    obuf = O.BufferRollouts(9, 1, 3)
    obuf.update(back["feat"], back["mask"], back["selected_action_probabilities"], back["selected_actions"],
                back["rewards"], back["terminal"])
    want = O.step_epoch(opol, O.Adam(1e-4), obuf, 0.05, 10, 0.01, perm1)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7)
    m, s = P.average_returns(FixedPolicy(), env, 5)
    assert m == 10.0 and s == 0.0
    pol.close(); rollouts.close()


def test_ppo_iterate_outer_loop_on_the_testenv(ctx):
    """``ppo_iterate!`` (src/train.jl:210-249) with the DEVICE policy in the loop: every iteration collects rollouts with
    the current device weights (weight round trip through ppo_batch_action_probabilities), trains, and hands the loss
    history to the evaluator's save_loss hook.  The TestEnv rewards action 1 only, so the policy's probability of action 1
    must grow over the iterations; history lengths and the printed protocol follow the reference."""
    import io

    class Env:
        def __init__(self): self.k, self.last = 0, 1
        def state(self): return P.StateData(np.ones((1, 9), np.float32), np.zeros(3, np.float32))
        def step_(self, a): self.k += 1; self.last = a
        def reward(self): return 1.0 if self.last == 1 else 0.0
        def is_terminal(self): return self.k >= 5
        def reset_(self): self.k = 0

    class Evaluator:
        def __init__(self): self.calls, self.saved = 0, None
        def __call__(self, policy, env, optimizer): self.calls += 1
        def save_loss(self, loss): self.saved = {k: list(v) for k, v in loss.items()}

    cfg = S.CONFIGS["c1"]
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    env, ev, out = Env(), Evaluator(), io.StringIO()
    from ppo_b200 import collect_rollouts as CR
    CR.seed_sampling(5)
    p_before = P.action_probabilities(pol, env.state())[0]
    loss = P.ppo_iterate_(pol, env, P.Adam(3e-3), 40, 50, 3, ev, 4, 1.0, 0.2, 0.0,
                          rollouts_factory=lambda: P.DeviceRollouts(9, 1, 3, 256, ctx), out=out)
    p_after = P.action_probabilities(pol, env.state())[0]
    assert ev.calls == 3 and ev.saved == loss
    assert len(loss["ppo"]) == len(loss["entropy"]) == len(loss["lr"]) == 12 and all(lr == 3e-3 for lr in loss["lr"])
    text = out.getvalue()
    assert text.count("PPO ITERATION : ") == 3 and text.count("EPOCH : ") == 12
    assert np.all(np.isfinite(loss["ppo"])) and p_after > p_before + 0.02, (p_before, p_after)
    pol.close()


def test_advantage_normalisation_extension(ctx):
    cfg = S.CONFIGS["t1"]
    data, old, buf, obuf = _filled(ctx, cfg)
    P.compute_state_value_(buf, 1.0)
    ret = buf.rewards
    buf.normalize_advantage(True, 1e-8)
    buf.set_permutation(np.arange(1, cfg.N + 1))
    got = gather_minibatch(P.construct_dataset(buf), 0, cfg.N)["returns"]
    assert np.allclose(got, O.normalize_advantage(ret), rtol=1e-5, atol=1e-6)
    buf.normalize_advantage(False)
    assert np.array_equal(gather_minibatch(P.construct_dataset(buf), 0, cfg.N)["returns"], ret)
    buf.close()


def test_error_behaviour(ctx):
    buf = P.DeviceRollouts(4, 2, 2, 8, ctx)
    f = np.zeros((3, 2, 4), np.float32); m = np.zeros((3, 4), np.float32)
    with pytest.raises(P.PPOError):      # action outside 1..A
        buf.append(f, m, np.ones(3), np.array([1, 5, 2]), np.zeros(3), np.zeros(3))
    buf.append(f, m, np.ones(3), np.array([1, 4, 2]), np.zeros(3), np.zeros(3))
    # the capacity is an initial reservation: like the reference's push!-grown vectors the buffer grows on demand
    buf.append(np.ones((6, 2, 4), np.float32), np.zeros((6, 4), np.float32), np.ones(6), np.ones(6), np.zeros(6), np.zeros(6))
    assert len(buf) == 9 and np.array_equal(buf.read()["feat"][:3], f) and np.all(buf.read()["feat"][3:] == 1.0)
    ds = P.construct_dataset(buf)
    with pytest.raises(P.PPOError):      # index outside 1..length
        ds[np.array([0, 1])]
    assert ds[4]["selected_action"] == 1
    for bad in (0, 10):                      # @assert 1 <= idx <= length, src/rollout_buffer.jl:105-106
        with pytest.raises(AssertionError):
            ds[bad]
    with pytest.raises(TypeError):
        ds["x"]
    buf.close()


def test_cuda_graph_epoch_is_bit_identical(ctx, monkeypatch):
    # the epoch loop replayed from a CUDA graph (device-side minibatch counter) vs plain launches
    cfg = S.CONFIGS["t1"]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    perm1 = O.feistel_permutation(cfg.N, 5) + 1
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PPO_B200_NO_GRAPH", mode)
        buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
        buf.append(data["feat"], data["mask"], old, data["action"], data["reward"], data["terminal"])
        P.compute_state_value_(buf, 1.0)
        pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
        opt = P.Adam(1e-4)
        losses = [P.step_epoch_(pol, opt, P.construct_dataset(buf), 0.05, 48, 0.01, perm=perm1) for _ in range(2)]
        Wd, bd = pol.weights()
        out[mode] = (losses, np.concatenate([w.ravel() for w in Wd] + [x.ravel() for x in bd]))
        pol.close(); buf.close()
    assert out["0"][0] == out["1"][0]
    assert np.array_equal(out["0"][1], out["1"][1])


# ---- batched rollout inference ----------------------------------------------
def test_sample_actions_matches_oracle_sampler(ctx):
    """device categorical sampling == the oracle's restatement of rand(Categorical(ap)) on the SAME probabilities and
    draws (bit-exact actions); the returned probability is the sampled entry; masked actions are never drawn."""
    cfg = S.CONFIGS["t1"]
    rng = np.random.default_rng(11)
    nb = 3000
    feat = rng.integers(-3, 9, (nb, cfg.nhe, cfg.nf)).astype(np.float32)
    mask = S.make_masks(rng, nb, cfg.nhe, cfg.apa)
    W, b = S.make_weights(cfg)
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    state = P.StateData(feat, mask)
    act, prob, probs = P.batch_sample_actions(pol, state, 12345, return_probabilities=True)
    np.testing.assert_array_equal(probs, P.batch_action_probabilities(pol, state))
    want_act, want_prob = O.sample_actions_from_probs(probs, 12345)
    np.testing.assert_array_equal(act, want_act)
    np.testing.assert_array_equal(prob, want_prob)
    assert np.all(np.isfinite(mask[np.arange(nb), act - 1])), "a masked action was drawn"
    act2, _ = P.batch_sample_actions(pol, state, 12346)
    assert np.mean(act2 != act) > 0.3          # another seed, another stream
    pol.close()


def test_sample_actions_follow_the_distribution(ctx):
    """200 000 draws for one state: empirical frequencies within 5 sigma of the probabilities"""
    cfg = S.CONFIGS["t1"]
    rng = np.random.default_rng(12)
    nb = 200_000
    f1 = rng.integers(-3, 9, (1, cfg.nhe, cfg.nf)).astype(np.float32)
    m1 = S.make_masks(rng, 1, cfg.nhe, cfg.apa)
    W, b = S.make_weights(cfg)
    W = [w * 3.0 for w in W]                   # a less uniform distribution
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)
    state = P.StateData(np.repeat(f1, nb, 0), np.repeat(m1, nb, 0))
    act, prob, probs = P.batch_sample_actions(pol, state, 7, return_probabilities=True)
    p = probs[0].astype(np.float64)
    freq = np.bincount(act - 1, minlength=cfg.A) / nb
    sigma = np.sqrt(p * (1 - p) / nb)
    assert np.all(np.abs(freq - p) <= 5 * sigma + 1e-9), np.max(np.abs(freq - p) / (sigma + 1e-12))
    assert np.all(freq[p == 0] == 0)
    pol.close()


# ---- variable-size states: the reference pads (triangle_utilities.jl:31-55) ---------------------
def test_variable_size_states_padded_like_the_reference(ctx):
    """states with 3..8 tokens go into a buffer of 8: the update matches the oracle run on the reference's padding
    (zeros / -Inf32 to the largest state), and padded actions get probability exactly 0"""
    nf, nhe, apa, H, L = 8, 8, 4, 32, 2
    rng = np.random.default_rng(21)
    n = 96
    states, feat, mask = [], np.zeros((n, nhe, nf), np.float32), np.full((n, nhe * apa), -np.inf, np.float32)
    for i in range(n):
        k = int(rng.integers(3, nhe + 1))
        vs = rng.integers(-3, 9, (k, nf)).astype(np.float32)
        am = np.where(rng.random(k * apa) < 0.2, -np.inf, 0.0).astype(np.float32)
        am[0] = 0.0
        states.append(P.StateData(vs, am))
        feat[i, :k], mask[i, :k * apa] = vs, am
    padded = P.prepare_state_data_for_batching_(list(states), nhe, apa)
    np.testing.assert_array_equal(np.stack([s.vertex_score for s in padded]), feat)
    np.testing.assert_array_equal(np.stack([s.action_mask for s in padded]), mask)
    cfg = S.Config("ragged", 94, n, nf, nhe, apa, H, L, 32)
    W, b = S.make_weights(cfg)
    opol = O.Policy(nf, H, L, apa)
    opol.W, opol.b = [w.copy() for w in W], [x.copy() for x in b]
    probs = O.batch_action_probabilities(opol, feat, mask)
    act = np.array([rng.choice(np.flatnonzero(np.isfinite(m))) + 1 for m in mask], np.int64)
    old = probs[np.arange(n), act - 1].astype(np.float32)
    rew = rng.integers(-4, 5, n).astype(np.float32)
    term = (np.arange(n) % 12 == 11)
    buf = P.DeviceRollouts(nf, nhe, apa, n, ctx)
    for i in range(n):
        P.update_(buf, states[i], old[i], act[i], rew[i], term[i])          # un-padded states
    P.compute_state_value_(buf, 1.0)
    pol = P.Policy(nf, H, L, apa, ctx, weights=W, biases=b)
    dprobs = P.batch_action_probabilities(pol, P.StateData(feat, mask))
    assert np.all(dprobs[~np.isfinite(mask)] == 0.0)
    # the padded tokens are stored (the buffer keeps the reference's padded layout) but the MLP never
    # runs them: the fp16-split engine compacts every minibatch to the tokens that have an unmasked action
    if pol.gemm_mode == P.GEMM_F16X3_TC:
        live = int((~np.all(np.isneginf(mask.reshape(-1, apa)), axis=1)).sum())
        assert pol.active_tokens() == live < n * nhe
    perm = np.random.default_rng(2).permutation(n) + 1
    got = P.step_epoch_(pol, P.Adam(1e-4), P.construct_dataset(buf), 0.05, cfg.B, 0.01, perm=perm)
    obuf = O.BufferRollouts(nf, nhe, apa)
    obuf.update(feat, mask, old, act, O.compute_returns(rew, term, 1.0), term)
    want = O.step_epoch(opol, O.Adam(1e-4), obuf, 0.05, cfg.B, 0.01, perm)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7), (got, want)
    Wd, bd = pol.weights()
    for l in range(len(Wd)):
        assert np.allclose(Wd[l], opol.W[l], rtol=1e-4, atol=2e-6)
    pol.close(); buf.close()
