"""Host-side logic that needs no GPU: sharding, index helpers, formatting, gloo world_size=2."""
import os
import sys

import numpy as np
import pytest

import ppo_b200 as P
from ppo_b200 import distributed as D
from ppo_b200 import synthetic as S
from oracle import ppo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_linear_index_and_clip_match_oracle():
    a = np.array([3, 1, 4, 2], np.int64)
    assert np.array_equal(P.get_linear_action_index(a, 5), O.get_linear_action_index(a, 5))
    assert P.get_linear_action_index(a, 5).tolist() == [3, 6, 14, 17]
    for adv in (-2.0, 0.0, 1.5):
        assert P.simplified_ppo_clip(adv, 0.05) == float(O.simplified_ppo_clip(np.float32(adv), 0.05))


def test_epoch_line_and_lr():
    assert P.format_epoch_line(7, 1.0, 2.0, 1e-4) == O.format_epoch_line(7, 1.0, 2.0, 1e-4)
    opt = P.Optimiser(P.Adam(1e-4))
    assert P.get_optimizer_learning_rate(opt) == 1e-4
    assert P.get_optimizer_learning_rate(P.Adam(3e-3)) == 3e-3


def test_shard_bounds_cut_at_episode_ends():
    rng = np.random.default_rng(1)
    for n, g in ((1000, 2), (1000, 8), (37, 4), (8, 8)):
        term = S.make_episode_terminals(rng, n, 30)
        bounds = D.shard_bounds_at_episode_ends(term, g)
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        for (a, b), (c, d) in zip(bounds[:-1], bounds[1:]):
            assert b == c
        for a, b in bounds:
            assert a <= b
            if b > a:
                assert term[b - 1]      # every shard ends on an episode end
        # the scan of the shards concatenated == the scan of the whole buffer
        r = rng.integers(-4, 5, n).astype(np.float32)
        whole = O.compute_returns(r, term, 0.99)
        parts = [O.compute_returns(r[a:b], term[a:b], 0.99) for a, b in bounds]
        assert np.array_equal(np.concatenate(parts), whole)


def test_synthetic_configs():
    c3 = S.CONFIGS["c3"]
    assert c3.num_params == 560644 and c3.A == 64 and c3.record_bytes() == 4364
    assert c3.flops_per_sample() == 2 * 16 * (3 * 559104 - 32768)
    c2 = S.CONFIGS["c2"]
    assert c2.num_params == 26372 and c2.record_bytes() == 19468
    d = S.make_buffer(S.CONFIGS["t1"])
    assert d["terminal"][-1] and d["feat"].min() >= -3 and d["feat"].max() <= 8
    m = d["mask"]
    assert np.all((m == 0) | np.isneginf(m)) and np.all(m[:, :16] == 0)
    assert np.all(m[np.arange(len(m)), d["action"] - 1] == 0)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import ppo_b200  # noqa: F401
    from ppo_b200 import distributed as DD
    from ppo_b200 import synthetic as SS
    from oracle import ppo_oracle as OO
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the unique-id broadcast path, with a fake id (no NCCL / GPU here)
    payload = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    assert payload[0] == bytes(range(128))
    cfg = SS.CONFIGS["t1"]
    data = SS.make_buffer(cfg)
    bounds = DD.shard_bounds_at_episode_ends(data["terminal"], world)
    a, b = bounds[rank]
    ret = OO.compute_returns(data["reward"][a:b], data["terminal"][a:b], 1.0)
    n_use = DD.equalize_counts(bounds)
    perm = OO.feistel_permutation(n_use, DD.local_seed(99, rank))
    out = [None] * world
    dist.all_gather_object(out, (a, b, float(ret.sum()), int(perm[0])))
    dist.destroy_process_group()
    q.put((rank, out))


def test_gloo_world2_sharding_roundtrip():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0] == res[1]
    cfg = S.CONFIGS["t1"]
    data = S.make_buffer(cfg)
    whole = O.compute_returns(data["reward"], data["terminal"], 1.0)
    (a0, b0, s0, _), (a1, b1, s1, _) = res[0]
    assert a0 == 0 and b0 == a1 and b1 == cfg.N
    assert abs((s0 + s1) - float(whole.sum())) < 1e-3


def test_padding_hooks_like_the_reference():
    """pad_vertex_scores / pad_action_mask / prepare_state_data_for_batching! of
    examples/triangle/distance_weighted/triangle_utilities.jl:31-55: zero columns, -Inf32 actions, to the largest state"""
    import ppo_b200 as P
    states = [P.StateData(np.ones((3, 2), np.float32), np.zeros(6, np.float32)),
              P.StateData(2 * np.ones((5, 2), np.float32), np.zeros(10, np.float32))]
    P.prepare_state_data_for_batching_(states)
    assert [s.vertex_score.shape for s in states] == [(5, 2), (5, 2)]
    assert np.all(states[0].vertex_score[3:] == 0) and np.all(states[0].vertex_score[:3] == 1)
    assert np.all(np.isneginf(states[0].action_mask[6:])) and np.all(states[0].action_mask[:6] == 0)
    assert states[0].action_mask.dtype == np.float32
    batched = P.batch_state(states)
    assert batched.vertex_score.shape == (2, 5, 2) and batched.action_mask.shape == (2, 10)
    # padding to a fixed buffer capacity (what DeviceRollouts.update_ does)
    vs = P.pad_vertex_scores([np.ones((3, 2))], 8)
    am = P.pad_action_mask([np.zeros(12)], 32)
    assert vs[0].shape == (8, 2) and am[0].shape == (32,) and np.isneginf(am[0][12:]).all()


def test_p2p_gradient_exchange_is_a_noop_for_one_rank():
    """enable_p2p_gradients leaves the NCCL path in place in a single-rank job (no GPU needed)"""
    import torch.distributed as dist
    from ppo_b200 import distributed as D
    import tempfile
    with tempfile.NamedTemporaryFile() as f:
        dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
        try:
            assert D.enable_p2p_gradients(policy=None) is False
        finally:
            dist.destroy_process_group()


def test_pack_action_mask_is_the_julia_bitmatrix_layout():
    """bit (a + A*s) of the stream = bit ((a + A*s) & 63) of word (a + A*s) >> 6, 1 = allowed: what
    ``BitMatrix(isfinite.(action_mask)).chunks`` holds for an (A, n) mask in Julia (column-major, little-endian bits)"""
    import ppo_b200 as P
    rng = np.random.default_rng(4)
    n, A = 37, 12                                  # n*A = 444: not a multiple of 64
    mask = np.where(rng.random((n, A)) < 0.4, -np.inf, 0.0).astype(np.float32)
    bits = P.pack_action_mask(mask)
    assert bits.dtype == np.uint64 and bits.size == 7
    flat = np.isfinite(mask).reshape(-1)
    for k in range(n * A):
        assert ((int(bits[k >> 6]) >> (k & 63)) & 1) == int(flat[k])
    assert int(bits[6]) >> (444 - 384) == 0        # padding bits are zero
