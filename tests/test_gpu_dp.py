"""Minibatch-sharded data parallelism on 2 GPUs: one process per GPU, NCCL gradient all-reduce
inside ppo_step_epoch.  The oracle restates the sharded scheme on the CPU: global minibatch k is the union of
every rank's local minibatch k, normalised by the global row count.  Skipped when fewer than 2 GPUs are visible
(the driver's 1-GPU run); executed under `gpurun --gpus 2`."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the fp32 FFMA engine on t1; the fp16-split tensor-core engine (what bench.py / SCALE time, with its per-layer overlapped
# NCCL all-reduce or the peer-memory exchange, under CUDA-graph replay) on t2, a shape inside its contract
_CFG = {"ffma": "t1", "f16x3": "t2"}


def _worker(rank, world, port, q, p2p=False, engine="ffma"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ppo_b200 as P
    from ppo_b200 import distributed as D
    from ppo_b200 import synthetic as S
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = P.Context(rank)
    D.init_comm(ctx)
    cfg = S.CONFIGS[_CFG[engine]]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    bounds = D.shard_bounds_at_episode_ends(data["terminal"], world)
    n_use = D.equalize_counts(bounds)
    a = bounds[rank][0]
    sl = slice(a, a + n_use)
    # returns are computed on the whole local shard (cut at an episode end), then the first n_use rows are used
    full = slice(bounds[rank][0], bounds[rank][1])
    ret = P.compute_returns(data["reward"][full], data["terminal"][full], 1.0, ctx)[:n_use]
    buf = P.DeviceRollouts(cfg.nf, cfg.nhe, cfg.apa, n_use, ctx)
    buf.append(data["feat"][sl], data["mask"][sl], old[sl], data["action"][sl], ret, data["terminal"][sl])
    mode = {"ffma": P.GEMM_FP32_SIMT, "f16x3": P.GEMM_F16X3_TC}[engine]
    pol = P.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa, ctx, weights=W, biases=b, gemm_mode=mode)
    assert pol.gemm_mode == mode
    if p2p:
        assert D.enable_p2p_gradients(pol)
    opt = P.Adam(1e-4)
    losses = P.step_epoch_(pol, opt, P.construct_dataset(buf), 0.05, 64, 0.01, seed=D.local_seed(99, rank))
    Wd, bd = pol.weights()
    flat = np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(Wd, bd)])
    q.put((rank, losses, flat, n_use, bounds))
    dist.barrier()
    pol.close(); buf.close(); ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("engine", ["ffma", "f16x3"])
@pytest.mark.parametrize("p2p", [False, True], ids=["nccl", "peer-memory"])
def test_two_rank_epoch_matches_sharded_oracle(p2p, engine):
    """p2p = True: the gradient exchange over CUDA-IPC peer memory fused into Adam (csrc/dp_p2p.cu) instead of NCCL
    (fp16-split engine + NCCL = the per-layer all-reduce overlapped with the backward pass)"""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import ppo_b200 as P  # noqa: F401
    from ppo_b200 import distributed as D
    from ppo_b200 import synthetic as S
    from oracle import ppo_oracle as O
    world = 2
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    port = 29600 + (os.getpid() % 1000) + (7 if p2p else 0) + (13 if engine == "f16x3" else 0)
    procs = [mctx.Process(target=_worker, args=(r, world, port, q, p2p, engine)) for r in range(world)]
    [p.start() for p in procs]
    res = {}
    for _ in range(world):
        r = q.get(timeout=300)
        res[r[0]] = r
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # ---- CPU restatement of the sharded scheme ----
    cfg = S.CONFIGS[_CFG[engine]]
    data = S.make_buffer(cfg)
    W, b = S.make_weights(cfg)
    old = S.rng_for(cfg, 7).uniform(0.05, 1.0, cfg.N).astype(np.float32)
    bounds = res[0][4]
    n_use = res[0][3]
    returns = O.compute_returns(data["reward"], data["terminal"], 1.0)      # shards end on episode ends
    opol = O.Policy(cfg.nf, cfg.H, cfg.L, cfg.apa)
    opol.W, opol.b = [w.copy() for w in W], [x.copy() for x in b]
    oopt = O.Adam(1e-4)
    perms = [O.feistel_permutation(n_use, D.local_seed(99, r)) + bounds[r][0] for r in range(world)]
    B = 64
    ph, eh = [], []
    for start in range(0, n_use, B):
        idx = np.concatenate([pm[start:start + B] for pm in perms])
        pl, ew = O.step_batch(opol, oopt, data["feat"][idx], data["mask"][idx], data["action"][idx], old[idx],
                              returns[idx], 0.05, 0.01)
        ph.append(pl); eh.append(ew)
    want = (float(np.mean(ph)), float(np.mean(eh)))
    for r in range(world):
        assert np.allclose(res[r][1], want, rtol=1e-5, atol=1e-7), (r, res[r][1], want)
        assert np.max(np.abs(res[r][2] - opol.flat())) <= 2e-5
    assert np.array_equal(res[0][2], res[1][2])          # replicas stay bit-identical
