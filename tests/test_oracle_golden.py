"""Pin the oracle against every known answer the reference holds for the hot path."""
import csv
import os

import numpy as np

from oracle import c_oracle as CO
from oracle import ppo_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_trajectory_csv_known_answer():
    # reference output/trajectory.csv:1-7 — six transitions, reward 1, gamma 1 => 6,5,4,3,2,1
    rows = list(csv.DictReader(open(os.path.join(G, "reference_trajectory.csv"))))
    assert [r["sample_names"] for r in rows] == [f"sample_{i}.bson" for i in range(1, 7)]
    want = np.array([float(r["returns"]) for r in rows], np.float32)
    assert want.tolist() == [6, 5, 4, 3, 2, 1]
    rewards = np.ones(6, np.float32)
    terminal = np.array([0, 0, 0, 0, 0, 1], bool)
    for fn in (O.compute_returns, CO.compute_returns):
        assert np.array_equal(fn(rewards, terminal, 1.0), want)
        # the reference leaves an un-terminated tail starting from 0 as well
        assert np.array_equal(fn(rewards, np.zeros(6, bool), 1.0), want)
    assert all(int(r["selected_actions"]) == 4 and float(r["selected_action_probabilities"]) == 0.5 for r in rows)


def test_testenv_run_known_answer():
    # reference test/test_rollout_buffer.jl:4-50 — 10 episodes x 10 steps, reward 1.0, discount 1.0
    buf = O.BufferRollouts(3, 3, 1)
    for _ in range(10):
        for step in range(10):
            buf.update(np.zeros((1, 3, 3)), np.zeros((1, 3)), 1.0, 1, 1.0, step == 9)
    buf.compute_state_value(1.0)
    assert len(buf) == 100
    assert np.array_equal(buf.rewards, np.tile(np.arange(10, 0, -1, dtype=np.float32), 10))
    assert buf.selected_actions.tolist() == [1] * 100
    assert np.flatnonzero(buf.terminal).tolist() == list(range(9, 100, 10))


def _parse_bson_doc(b, pos=0):
    import struct
    size = struct.unpack_from("<i", b, pos)[0]
    end = pos + size
    p = pos + 4
    out = {}
    while b[p] != 0:
        t = b[p]; p += 1
        e = b.index(b"\x00", p); key = b[p:e].decode(); p = e + 1
        if t in (3, 4):
            out[key], p = _parse_bson_doc(b, p)
            if t == 4:
                out[key] = [out[key][k] for k in sorted(out[key], key=int)]
        elif t == 2:
            n = struct.unpack_from("<i", b, p)[0]; out[key] = b[p + 4:p + 4 + n - 1].decode(); p += 4 + n
        elif t == 5:
            n = struct.unpack_from("<i", b, p)[0]; out[key] = b[p + 5:p + 5 + n]; p += 5 + n
        elif t == 18:
            out[key] = struct.unpack_from("<q", b, p)[0]; p += 8
        else:
            raise ValueError(f"bson type {t}")
    assert p + 1 == end
    return out, end


def test_bson_state_fixture():
    # reference test/test_rollout_to_disk.jl:17-23: update! writes states/sample_1.bson whose :state == [1,2,3,4,5]
    for name in ("reference_output_sample_1.bson.hex", "reference_rollout_to_disk_sample_1.bson.hex"):
        raw = bytes.fromhex(open(os.path.join(G, name)).read().strip())
        assert len(raw) == 183
        doc, end = _parse_bson_doc(raw)
        st = doc["state"]
        assert st["tag"] == "array" and st["type"]["name"] == ["Core", "Int64"] and st["size"] == [5]
        assert np.frombuffer(st["data"], "<i8").tolist() == [1, 2, 3, 4, 5]


def test_numpy_and_c_restatements_agree():
    rng = np.random.default_rng(7)
    for n in (1, 2, 17, 4097, 20000):
        r = rng.integers(-4, 5, n).astype(np.float32)
        t = rng.random(n) < 0.07
        for g, f32 in ((1.0, False), (0.99, False), (0.99, True), (0.5, False)):
            assert np.array_equal(O.compute_returns(r, t, g, f32), CO.compute_returns(r, t, g, f32))
    for n in (1, 2, 3, 4, 5, 31, 100, 4096, 4097, 65536, 100003):
        for seed in (0, 1, 2**63 + 5):
            p = O.feistel_permutation(n, seed)
            assert np.array_equal(p, CO.feistel_permutation(n, seed))
            assert np.array_equal(np.sort(p), np.arange(n))


def test_golden_npz_match_oracle():
    # the committed oracle-derived vectors are reproduced by the current oracle (drift guard)
    import ppo_b200  # noqa: F401
    from ppo_b200 import synthetic as S
    for name, key in (("oracle_t0_g1", "t0"), ("oracle_t0_g099", "t0"), ("oracle_t1_g1", "t1")):
        z = np.load(os.path.join(G, name + ".npz"))
        cfg = S.CONFIGS[key]
        data = S.make_buffer(cfg)
        assert np.array_equal(O.compute_returns(data["reward"], data["terminal"], float(z["gamma"])), z["returns"])
        assert np.array_equal(O.feistel_permutation(cfg.N, int(z["seed"])), z["perm0"])


def test_epoch_line_format():
    # src/train.jl:146
    assert O.format_epoch_line(3, 0.12345, -0.5, 1e-4) == "EPOCH : 3 \t PPO LOSS : 0.1235\t ENTROPY LOSS : -0.5000 \t LR : 1.0e-04\n"


def test_oracle_categorical_sampler_properties():
    """restated rand(Categorical(ap)): masked (p = 0) actions are never drawn, a one-hot row always returns its action
    (the TestEnv policy of test/test_rollout_buffer.jl:24-29 has probabilities [1, 0, 0] -> action 1), and the
    frequencies follow the probabilities"""
    from oracle import ppo_oracle as O
    probs = np.tile(np.array([[1.0, 0.0, 0.0]], np.float32), (50, 1))
    act, pr = O.sample_actions_from_probs(probs, 3)
    assert np.all(act == 1) and np.all(pr == 1.0)
    p = np.array([0.0, 0.25, 0.0, 0.5, 0.25, 0.0], np.float32)
    act, pr = O.sample_actions_from_probs(np.tile(p, (20000, 1)), 9)
    freq = np.bincount(act - 1, minlength=6) / 20000
    assert freq[0] == freq[2] == freq[5] == 0
    assert np.all(np.abs(freq - p) < 0.02)
    u = O.sample_uniforms(1, 1000)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
