"""Generate the committed golden fixtures (run in the build container; the GPU box never reads
/root/reference).

1. reference-derived known answers copied as DATA (no source): output/trajectory.csv and the two
   183-byte BSON state files -> tests/golden/reference_*.{csv,bson.hex}
2. oracle-derived vectors ("derived, not reference-pinned") for small seeded cases -> *.npz:
   inputs and the oracle's returns / permutation / loss / gradient / post-Adam weights, so a drift
   of either the oracle or the CUDA path is caught against a frozen file.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ppo_oracle as O  # noqa: E402
import ppo_b200  # noqa: E402,F401
from ppo_b200 import synthetic as S  # noqa: E402

REF = "/root/reference"


def reference_fixtures():
    if not os.path.isdir(REF):
        print("reference not present; skipping reference-derived fixtures")
        return
    with open(os.path.join(REF, "output/trajectory.csv")) as f:
        open(os.path.join(HERE, "reference_trajectory.csv"), "w").write(f.read())
    for src, dst in [("output/states/sample_1.bson", "reference_output_sample_1.bson.hex"),
                     ("examples/rollout_to_disk/states/sample_1.bson", "reference_rollout_to_disk_sample_1.bson.hex")]:
        data = open(os.path.join(REF, src), "rb").read()
        open(os.path.join(HERE, dst), "w").write(data.hex() + "\n")


def oracle_case(name, cfg_key, gamma, eps=0.05, w_ent=0.01, eta=1e-4, seed=4242):
    cfg = S.CONFIGS[cfg_key]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    pol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    pol.W, pol.b = [w.copy() for w in W], [x.copy() for x in b]
    probs = O.batch_action_probabilities(pol, data["feat"], data["mask"])
    old = S.make_old_probs(cfg, probs[np.arange(cfg.N), data["action"] - 1])
    returns = O.compute_returns(data["reward"], data["terminal"], gamma)
    perm0 = O.feistel_permutation(cfg.N, seed)
    # first minibatch: loss + gradient at the initial weights
    idx1 = perm0[:cfg.B] + 1
    buf = O.BufferRollouts(cfg.nf, cfg.nhe, cfg.apa)
    buf.update(data["feat"], data["mask"], old, data["action"], returns, data["terminal"])
    batch = O.get_batch(buf, idx1)
    feat, mask = batch["state"]
    ppoloss, entw, dW, db = O.policy_gradient(pol, feat, mask, batch["selected_action"],
                                              batch["selected_action_probability"], batch["returns"], eps, w_ent)
    grads = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(dW, db)])
    # one full epoch with Adam
    pol2 = pol.copy()
    mean_ppo, mean_ent = O.step_epoch(pol2, O.Adam(eta), buf, eps, cfg.B, w_ent, perm0 + 1)
    out = dict(gamma=gamma, eps=eps, w_ent=w_ent, eta=eta, seed=seed, old=old, returns=returns, perm0=perm0,
               ppoloss=ppoloss, entw=entw, grads=grads, mean_ppo=mean_ppo, mean_ent=mean_ent,
               flat_after=pol2.flat())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ppoloss", ppoloss, "entw", entw, "epoch", mean_ppo, mean_ent)


if __name__ == "__main__":
    reference_fixtures()
    oracle_case("oracle_t0_g1", "t0", 1.0)
    oracle_case("oracle_t0_g099", "t0", 0.99)
    oracle_case("oracle_t1_g1", "t1", 1.0)
    oracle_case("oracle_t2_g1", "t2", 1.0)
