#!/usr/bin/env python
"""Write the seeded inputs of the golden cases (t0, t1, t2 and a C2-shaped buffer on the reference's own trained weights,
test/output/catmull-clark-policy.bson, copied as data into tests/golden/reference_catmull_clark_policy.npz) as BSON documents that baseline/julia_ref.jl loads with BSON.load, so that the
real reference (Julia + Flux) and this repository's oracle / CUDA path see identical bytes.

    python baseline/make_julia_inputs.py [outdir = baseline/julia_inputs]
    python baseline/make_julia_inputs.py --fixture      # rebuild the weight fixture from /root/reference (build container only)

Arrays are written in the encoding BSON.jl itself uses for bits-type arrays (pinned against the reference's own
output/states/sample_1.bson by tests/test_disk_replay.py), with JULIA shapes: vertex_score [nf, nhe, N], action_mask [A, N],
Dense weights [out, in].  Nothing here is timed or shipped; it is the hand-over to a maintainer who has Julia."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppo_b200  # noqa: E402,F401
from ppo_b200 import bson_io as B  # noqa: E402
from ppo_b200 import synthetic as S  # noqa: E402
from oracle import ppo_oracle as O  # noqa: E402


def e_f64(k, v):
    return b"\x01" + B._cstr(k) + struct.pack("<d", float(v))


def arr(k, a, julia_shape):
    return B._e_doc(k, B.lower_array(julia_shape, np.ascontiguousarray(a)))


C2MINI = S.Config("c2-mini", 92, 512, 72, 64, 4, 128, 2, 128)
FIXTURE = os.path.join(ROOT, "tests", "golden", "reference_catmull_clark_policy.npz")


def case_arrays(name):
    """the seeded inputs of golden case `name` (what the .bson handed to Julia holds), as numpy arrays in the oracle's
    layout: (cfg, gamma, data dict incl. "old", W, b)"""
    key, g = name.rsplit("_", 1)
    gamma = {"g1": 1.0, "g099": 0.99}[g]
    if key == "c2mini_trained":
        cfg = C2MINI
        z = np.load(FIXTURE)
        W, b = [z[f"W{l}"] for l in (1, 2, 3)], [z[f"b{l}"] for l in (1, 2, 3)]
    else:
        cfg = S.CONFIGS[key]
        W, b = S.make_weights(cfg)
    data = S.make_buffer(cfg)
    pol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    pol.W, pol.b = [w.copy() for w in W], [x.copy() for x in b]
    probs = O.batch_action_probabilities(pol, data["feat"], data["mask"])
    data["old"] = S.make_old_probs(cfg, probs[np.arange(cfg.N), data["action"] - 1])
    return cfg, gamma, data, W, b


CASES = ["t0_g1", "t0_g099", "t1_g1", "t2_g1", "c2mini_trained_g1"]
EPS, W_ENT, ETA, SEED = 0.05, 0.01, 1e-4, 4242


def write_case(outdir, name):
    cfg, gamma, data, W, b = case_arrays(name)
    old, eps, w_ent, eta, seed = data["old"], EPS, W_ENT, ETA, SEED
    items = [
        B._e_i64("N", cfg.N), B._e_i64("nf", cfg.nf), B._e_i64("nhe", cfg.nhe), B._e_i64("apa", cfg.apa),
        B._e_i64("H", cfg.H), B._e_i64("L", cfg.L), B._e_i64("B", cfg.B), B._e_i64("seed", seed),
        e_f64("gamma", gamma), e_f64("eps", eps), e_f64("w_ent", w_ent), e_f64("eta", eta),
        # StateData.vertex_score is a Matrix{Int64} in the reference (test/quad_game_utilities.jl:50-56)
        arr("vertex_score", data["feat"].astype(np.int64), (cfg.nf, cfg.nhe, cfg.N)),
        arr("action_mask", data["mask"], (cfg.A, cfg.N)),
        arr("selected_actions", data["action"].astype(np.int64), (cfg.N,)),
        arr("selected_action_probabilities", old, (cfg.N,)),
        arr("rewards", data["reward"], (cfg.N,)),
        arr("terminal", data["terminal"].astype(np.bool_), (cfg.N,)),
    ]
    for l, (w, x) in enumerate(zip(W, b)):
        items.append(arr(f"W{l + 1}", w, (w.shape[1], w.shape[0])))     # C [in][out] == Julia [out, in]
        items.append(arr(f"b{l + 1}", x, (x.shape[0],)))
    path = os.path.join(outdir, f"{name}.bson")
    with open(path, "wb") as f:
        f.write(B._doc(items))
    print("wrote", path, os.path.getsize(path), "bytes")


def reference_trained_weights():
    """Dense weights of the reference's own trained policy (test/output/catmull-clark-policy.bson:
    SimplePolicy.Policy(72, 128, 2, 4)), when the reference tree is present: realistic, non-Glorot weights."""
    path = "/root/reference/test/output/catmull-clark-policy.bson"
    if not os.path.exists(path):
        return None
    try:
        doc, _ = B._parse_doc(open(path, "rb").read())
    except Exception as e:   # noqa: BLE001  (BSON.jl back-references etc.: outside the parser's subset)
        print("reference policy file not parsed:", e)
        return None
    found = []

    def walk(v):
        if isinstance(v, dict):
            if v.get("tag") == "array" and isinstance(v.get("data"), (bytes, bytearray)):
                try:
                    found.append(B.raise_value(v))
                except Exception:   # noqa: BLE001
                    pass
            else:
                for x in v.values():
                    walk(x)
        elif isinstance(v, list):
            for x in v:
                walk(x)

    walk(doc)
    mats = [a for a in found if a.dtype == np.float32 and a.ndim == 2]
    vecs = [a for a in found if a.dtype == np.float32 and a.ndim == 1]
    if len(mats) == 3 and len(vecs) == 3:
        # raise_value returns Julia [out, in] as C [in][out]: exactly the oracle's layout
        return [np.ascontiguousarray(m) for m in mats], [np.ascontiguousarray(v) for v in vecs]
    print(f"reference policy file: found {len(mats)} matrices / {len(vecs)} vectors, expected 3 / 3")
    return None


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--fixture":      # (re)build the committed weight fixture from the reference tree
        wb = reference_trained_weights()
        assert wb is not None and wb[0][0].shape == (72, 128)
        np.savez_compressed(FIXTURE, **{f"W{l + 1}": w for l, w in enumerate(wb[0])}, **{f"b{l + 1}": x for l, x in enumerate(wb[1])})
        print("wrote", FIXTURE)
        sys.exit(0)
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "baseline", "julia_inputs")
    os.makedirs(outdir, exist_ok=True)
    for name in CASES:
        write_case(outdir, name)
