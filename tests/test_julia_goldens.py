"""Consumers of the goldens that baseline/julia_ref.jl writes when run by someone who has Julia + Flux
(tests/golden/julia_<case>.bson: the REAL reference's returns / loss / gradient / post-Adam weights on this repository's
seeded inputs, see baseline/julia_ref.jl).

* present  -> the oracle (CPU, always) and the CUDA path (-m gpu) are compared with them at the stated tolerances: this is
              what turns "parity unpinned"  into "pinned" for the MLP / softmax / loss / gradient / Adam rows;
* absent   -> the tests SKIP with the reason "parity unpinned": no Julia exists in the build image, so no such file is
              committed yet.

The reader and the comparison logic themselves are exercised on every run by `test_consumer_on_a_synthetic_golden`, which
fabricates a golden from the oracle in a temporary directory (it proves the plumbing, not parity)."""
import importlib.util
import os

import numpy as np
import pytest

from ppo_b200 import bson_io as B
from oracle import ppo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
_spec = importlib.util.spec_from_file_location("make_julia_inputs", os.path.join(ROOT, "baseline", "make_julia_inputs.py"))
MJ = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MJ)

UNPINNED = ("parity unpinned for this case: tests/golden/julia_{}.bson is absent (produce it with `python "
            "baseline/make_julia_inputs.py && julia --project=<reference checkout> baseline/julia_ref.jl`)")


def load_golden(path):
    doc, _ = B._parse_doc(open(path, "rb").read())
    return {k: B.raise_value(v) for k, v in doc.items()}


def oracle_outputs(name, perm1):
    """what julia_ref.jl computes, by the oracle, for the permutation the Julia run drew"""
    cfg, gamma, data, W, b = MJ.case_arrays(name)
    returns = O.compute_returns(data["reward"], data["terminal"], gamma)
    buf = O.BufferRollouts(cfg.nf, cfg.nhe, cfg.apa)
    buf.update(data["feat"], data["mask"], data["old"], data["action"], returns, data["terminal"])
    pol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    pol.W, pol.b = [w.copy() for w in W], [x.copy() for x in b]
    batch = O.get_batch(buf, perm1[:cfg.B])
    feat, mask = batch["state"]
    pl, ew, dW, db = O.policy_gradient(pol, feat, mask, batch["selected_action"], batch["selected_action_probability"],
                                       batch["returns"], MJ.EPS, MJ.W_ENT)
    grads = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(dW, db)])
    pol2 = pol.copy()
    mp, me = O.step_epoch(pol2, O.Adam(MJ.ETA), buf, MJ.EPS, cfg.B, MJ.W_ENT, perm1)
    return dict(returns=returns, ppoloss=pl, entw=ew, grads=grads, mean_ppo=mp, mean_ent=me, flat_after=pol2.flat(),
                flat_before=pol.flat())


def _steps(cfg):
    return -(-cfg.N // cfg.B)      # minibatches per epoch


def compare(got, z, gamma, what, steps):
    """tolerances of SURVEY 8(c): returns bit-exact for gamma = 1 (else 1e-6 + 1e-5 |x|), loss scalars 1e-5 relative,
    gradient 1e-5 of its max-abs.  Post-Adam weights: Adam normalises every step to ~eta = 1e-4 whatever the size of the
    gradient, so an entry whose gradient is at the level of Float32 summation noise (a dead unit's bias) takes steps of
    either sign in any two Float32 evaluations; the check is therefore 2e-5 absolute for 99.9 % of the entries and the
    worst case the optimiser allows (2 eta per minibatch) for the rest"""
    if gamma == 1.0:
        assert np.array_equal(got["returns"], z["returns"]), what
    else:
        assert np.all(np.abs(got["returns"] - z["returns"]) <= 1e-6 + 1e-5 * np.abs(z["returns"])), what
    assert abs(got["ppoloss"] - z["ppoloss"]) <= 1e-5 * abs(z["ppoloss"]) + 1e-7, (what, got["ppoloss"], z["ppoloss"])
    assert abs(got["entw"] - z["entw"]) <= 1e-5 * abs(z["entw"]) + 1e-8, what
    assert np.max(np.abs(got["grads"] - z["grads"])) <= 1e-5 * np.max(np.abs(z["grads"])) + 1e-7, what
    assert abs(got["mean_ppo"] - z["mean_ppo"]) <= 1e-5 * abs(z["mean_ppo"]) + 1e-6, what
    assert abs(got["mean_ent"] - z["mean_ent"]) <= 1e-5 * abs(z["mean_ent"]) + 1e-8, what
    dw = np.abs(got["flat_after"] - z["flat_after"])
    assert np.mean(dw > 2e-5) <= 1e-3 and np.max(dw) <= 2 * MJ.ETA * steps + 1e-6, (what, float(np.max(dw)), float(np.mean(dw > 2e-5)))


@pytest.mark.parametrize("name", MJ.CASES)
def test_oracle_against_the_julia_reference(name):
    path = os.path.join(G, f"julia_{name}.bson")
    if not os.path.exists(path):
        pytest.skip(UNPINNED.format(name))
    z = load_golden(path)
    cfg, gamma, *_ = MJ.case_arrays(name)
    compare(oracle_outputs(name, z["perm"]), z, gamma, f"oracle vs Julia ({z.get('flux_version')})", _steps(cfg))


def device_outputs(ctx, name, perm1):
    import ppo_b200 as P
    from ppo_b200.rollout_buffer import gather_minibatch
    cfg, gamma, data, W, b = MJ.case_arrays(name)
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, cfg.N, ctx)
    buf.append(data["feat"].astype(np.int64), data["mask"], data["old"], data["action"], data["reward"], data["terminal"])
    P.compute_state_value_(buf, gamma)
    out = {"returns": buf.rewards}
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b)      # PPO_GEMM_AUTO
    ds = P.construct_dataset(buf)
    buf.set_permutation(perm1)
    batch = gather_minibatch(ds, 0, cfg.B)
    lin = P.get_linear_action_index(batch["selected_action"], P.number_of_actions_per_state(batch["state"]))
    out["ppoloss"], out["entw"], out["grads"] = P.step_batch_(pol, None, batch["state"], lin, batch["selected_action_probability"],
                                                              P.batch_advantage(batch["state"], batch["returns"]), MJ.EPS,
                                                              MJ.W_ENT, return_grads=True)
    out["mean_ppo"], out["mean_ent"] = P.step_epoch_(pol, P.Adam(MJ.ETA), ds, MJ.EPS, cfg.B, MJ.W_ENT, perm=perm1)
    Wd, bd = pol.weights()
    out["flat_after"] = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(Wd, bd)])
    pol.close(); buf.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", MJ.CASES)
def test_device_against_the_julia_reference(ctx, name):
    path = os.path.join(G, f"julia_{name}.bson")
    if not os.path.exists(path):
        pytest.skip(UNPINNED.format(name))
    z = load_golden(path)
    cfg, gamma, *_ = MJ.case_arrays(name)
    compare(device_outputs(ctx, name, z["perm"]), z, gamma, f"device vs Julia ({z.get('flux_version')})", _steps(cfg))


def _fabricate(tmp_path, name):
    """a stand-in for julia_ref.jl's output, written with the same BSON encoding, from the ORACLE (plumbing test only)"""
    import struct
    cfg, *_ = MJ.case_arrays(name)
    perm1 = np.random.default_rng(7).permutation(cfg.N).astype(np.int64) + 1
    o = oracle_outputs(name, perm1)
    items = [MJ.arr("returns", o["returns"], o["returns"].shape), MJ.arr("perm", perm1, perm1.shape),
             MJ.e_f64("ppoloss", o["ppoloss"]), MJ.e_f64("entw", o["entw"]), MJ.arr("grads", o["grads"].astype(np.float32), o["grads"].shape),
             MJ.e_f64("mean_ppo", o["mean_ppo"]), MJ.e_f64("mean_ent", o["mean_ent"]),
             MJ.arr("flat_after", o["flat_after"].astype(np.float32), o["flat_after"].shape),
             B._e_str("flux_version", "fabricated-from-oracle"), B._e_str("julia_version", "none")]
    path = os.path.join(tmp_path, f"julia_{name}.bson")
    with open(path, "wb") as f:
        f.write(B._doc(items))
    return path, o


def test_consumer_on_a_synthetic_golden(tmp_path):
    path, o = _fabricate(str(tmp_path), "t0_g099")
    z = load_golden(path)
    assert z["flux_version"] == "fabricated-from-oracle" and z["perm"].dtype == np.int64
    steps = _steps(MJ.case_arrays("t0_g099")[0])
    compare(oracle_outputs("t0_g099", z["perm"]), z, 0.99, "self-check", steps)
    bad = dict(z)
    bad["grads"] = z["grads"] * (1 + 1e-4)
    with pytest.raises(AssertionError):
        compare(oracle_outputs("t0_g099", z["perm"]), bad, 0.99, "self-check must notice a 1e-4 gradient error", steps)


@pytest.mark.gpu
def test_device_consumer_on_a_synthetic_golden(ctx, tmp_path):
    """the same plumbing through the CUDA path, on the reference's own trained weights (catmull-clark policy) and on the
    shape PPO_GEMM_AUTO runs on the fp16-split engine"""
    for name in ("c2mini_trained_g1", "t2_g1"):
        path, _ = _fabricate(str(tmp_path), name)
        z = load_golden(path)
        compare(device_outputs(ctx, name, z["perm"]), z, 1.0, f"device vs fabricated golden {name}", _steps(MJ.case_arrays(name)[0]))
