#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): rollouts written to disk in the reference's format -- trajectory.csv plus one
states/sample_i.bson per transition, exactly what src/rollouts_to_disk.jl writes -- then replayed through the device PPO
update.  One process per GPU (run under torchrun for N > 1: every rank writes and replays its own shard and the minibatch
gradients are exchanged over NVLink peer memory).

The mesh-game environments (random_quad / rand_poly_env) need un-vendored packages, so the episodes are
synthetic with random_quad's shapes: nf = 216 features per half-edge, 4 actions per half-edge, <= 30 steps per episode,
Policy(216, 128, 2, 4) (test/random_quad.jl:43-49,61).  The on-disk set (--episodes) is replicated to --transitions rows
in the device buffer, as SURVEY 8(d) describes for C5.

    python scripts/c5_disk_replay.py [--episodes 20] [--transitions 131072] [--batch 8192] [--nhe 16]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=20)
    ap.add_argument("--transitions", type=int, default=131072)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--nhe", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=2)
    args = ap.parse_args()
    import torch
    import ppo_b200 as P
    from ppo_b200 import distributed as D
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    torch.cuda.set_device(local_rank)
    ctx = P.Context(local_rank)
    if world > 1:
        D.init_comm(ctx)
    nf, nhe, apa, H, L = 216, args.nhe, 4, 128, 2
    A = nhe * apa
    rng = np.random.default_rng(20260118 + 5 + 1000 * rank)

    # ---- 1. write the rollouts to disk in the reference's format (src/rollouts_to_disk.jl:23-132) ----
    root = tempfile.mkdtemp(prefix=f"c5_rank{rank}_")
    t0 = time.perf_counter()
    disk = P.DiskRollouts(root)
    host = {"feat": [], "mask": [], "act": [], "prob": []}
    for _ in range(args.episodes):
        steps = int(rng.integers(1, 31))
        for s in range(steps):
            vs = rng.integers(-3, 9, (nhe, nf)).astype(np.int64)               # Matrix{Int64}[nf, nhe]
            am = np.repeat(np.where(rng.random(nhe // 4 if nhe >= 4 else 1) < 0.25, -np.inf, 0.0), 4 * apa)[:A].astype(np.float32)
            am[:4 * apa] = 0.0                                                  # first quad always active
            a = int(rng.choice(np.flatnonzero(np.isfinite(am)))) + 1
            p = float(np.float32(rng.uniform(0.05, 1.0)))
            P.update_(disk, P.StateData(vs, am), p, a, float(rng.integers(-4, 5)), s == steps - 1)
            host["feat"].append(vs.astype(np.float32)); host["mask"].append(am); host["act"].append(a); host["prob"].append(p)
    P.write_returns_to_disk(disk, 1.0, ctx)
    t_write = time.perf_counter() - t0
    n_disk = len(disk)

    # ---- 2. replay: C++ bulk loader (CSV + BSON) into the device buffer, then replicate to the requested size ----
    t0 = time.perf_counter()
    ds = P.DiskDataset(root)
    buf, has_returns = ds.to_device(nf, nhe, apa, ctx, capacity=max(args.transitions, n_disk))
    ctx.sync()
    t_load = time.perf_counter() - t0
    assert has_returns and len(buf) == n_disk
    got = buf.read(0, n_disk)
    assert np.array_equal(got["feat"], np.stack(host["feat"])) and np.array_equal(got["mask"], np.stack(host["mask"]))
    assert np.array_equal(got["selected_actions"], np.array(host["act"])), "replayed actions differ from what was written"
    assert np.allclose(got["selected_action_probabilities"], np.array(host["prob"], np.float32), rtol=1e-6)
    feat, mask = np.stack(host["feat"]), np.stack(host["mask"])
    act, prob = np.array(host["act"], np.int64), np.array(host["prob"], np.float32)
    ret = buf.rewards.copy()
    while len(buf) + n_disk <= args.transitions:                                # replicate (returns are already final)
        buf.append(feat, mask, prob, act, ret, np.zeros(n_disk, bool))
    n = len(buf)

    # ---- 3. the device PPO update on the replayed buffer ----
    pol = P.Policy(nf, H, L, apa, ctx)
    engine = pol.set_gemm_mode(P.GEMM_AUTO)
    if world > 1:
        D.enable_p2p_gradients(pol)
    opt = P.Optimiser(P.Adam(1e-4))
    B = max(1, min(args.batch // world, n))
    P.ppo_train_(pol, opt, P.construct_dataset(buf), 0.05, B, 1, 0.01, seed=1, out=None)           # warm-up epoch
    ctx.sync()
    t0 = time.perf_counter()
    hist = P.ppo_train_(pol, opt, P.construct_dataset(buf), 0.05, B, args.epochs, 0.01, seed=2, out=None)
    ctx.sync()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"config": "c5-disk-replay", "n_gpus": world, "on_disk_transitions_per_rank": n_disk,
                          "replayed_transitions_per_rank": n, "nf": nf, "nhe": nhe, "mlp": f"{L}x{H}", "gemm_engine": engine,
                          "write_s": round(t_write, 3), "load_s": round(t_load, 4),
                          "load_transitions_per_s": round(n_disk / t_load, 1),
                          "update_samples_per_s": round(n * world * args.epochs / dt, 1),
                          "ppo_loss": hist[0], "minibatch_rows_per_rank": B}))
    pol.close(); buf.close(); ctx.close()
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
