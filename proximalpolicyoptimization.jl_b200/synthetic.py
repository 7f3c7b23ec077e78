"""Synthetic rollout buffers of the benchmark shapes (data generation only).

The mesh-game environments of the reference need packages that are not vendored
(test/quad_game_utilities.jl:1-6), so configs reproduce the data *shapes and value ranges*:
features = small integers (vertex score / degree, 0 for missing: test/quad_game_utilities.jl:35-37,
50-52), mask = 0 / -Inf per "quad" group of 4*apa actions (:39-44), rewards = small integers
(no_action_reward = -4, :151), episodes of <= 30 steps (test/random_quad.jl:49).
All draws come from numpy ``default_rng(PCG64(20260118 + config))`` so the oracle and the device
see identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Config:
    name: str
    cid: int
    N: int
    nf: int
    nhe: int
    apa: int
    H: int
    L: int          # hidden layers
    B: int
    epochs: int = 1
    max_episode: int = 30

    @property
    def A(self):
        return self.nhe * self.apa

    @property
    def dims(self):
        return [self.nf] + [self.H] * self.L + [self.apa]

    @property
    def num_params(self):
        d = self.dims
        return sum(i * o + o for i, o in zip(d[:-1], d[1:]))

    def flops_per_sample(self):
        """K5+K7 FLOPs per sample: fwd + wgrad for every layer, dgrad except layer 1."""
        nf, H, L, apa, nhe = self.nf, self.H, self.L, self.apa, self.nhe
        return 2 * nhe * (3 * (nf * H + (L - 1) * H * H + H * apa) - nf * H)

    def record_bytes(self):
        return 4 * self.nf * self.nhe + 4 * self.A + 12


CONFIGS = {
    # C1: TestEnv plumbing case (test/test_rollout_buffer.jl:4-50): 10 episodes x 10 steps
    "c1": Config("c1-testenv", 1, 100, 9, 1, 3, 16, 1, 10),
    # C2: Policy(72,128,2,4) of test/test_square_mesh.jl:29
    "c2": Config("c2-65k-mlp2x128", 2, 65536, 72, 64, 4, 128, 2, 4096),
    # C3: the headline config: 1M transitions, MLP 3x512, 64k minibatches
    "c3": Config("c3-1m-mlp3x512", 3, 1048576, 64, 16, 4, 512, 3, 65536),
    # C4: 8M transitions sharded over G GPUs (model / B as C3)
    "c4": Config("c4-8m-mlp3x512-dp", 4, 8388608, 64, 16, 4, 512, 3, 65536),
    # tiny shapes for tests
    "t0": Config("t0-tiny", 90, 257, 8, 3, 4, 16, 2, 50),
    "t1": Config("t1-small", 91, 1000, 12, 16, 4, 32, 3, 128),
    # small shape INSIDE the fp16-split tensor-core engine's contract (in % 8, hidden % 32): what PPO_GEMM_AUTO runs
    # the benchmark configs on, at test size
    "t2": Config("t2-f16", 93, 1024, 16, 8, 4, 64, 2, 256),
}


def rng_for(cfg: Config, stream: int = 0):
    return np.random.default_rng(np.random.PCG64(20260118 + cfg.cid + 1000 * stream))


def make_episode_terminals(rng, n, max_len):
    """terminal flags for episodes of uniform length in [1, max_len]; last transition terminal."""
    term = np.zeros(n, dtype=bool)
    # draw enough episode lengths, cumulative sum gives the end positions
    est = max(16, int(2.5 * n / (max_len + 1)) + 16)
    pos = np.cumsum(rng.integers(1, max_len + 1, size=est))
    while pos[-1] < n:
        pos = np.concatenate([pos, pos[-1] + np.cumsum(rng.integers(1, max_len + 1, size=est))])
    ends = pos[pos <= n] - 1
    term[ends] = True
    term[n - 1] = True
    return term


def make_masks(rng, n, nhe, apa):
    """0 / -Inf per group of 4*apa consecutive actions (test/quad_game_utilities.jl:39-44);
    inactive with p = 0.25, group 1 forced active.  When nhe is not a multiple of 4 the group is
    one half-edge (apa actions)."""
    A = nhe * apa
    g = 4 * apa if nhe % 4 == 0 else apa
    ng = A // g
    inactive = rng.random((n, ng)) < 0.25
    inactive[:, 0] = False
    mask = np.where(np.repeat(inactive, g, axis=1), -np.inf, 0.0).astype(np.float32)
    return mask


def make_actions(rng, mask):
    """uniform over the unmasked actions, Int64 1-based."""
    score = rng.random(mask.shape, dtype=np.float32)
    score[np.isneginf(mask)] = -1.0
    return (np.argmax(score, axis=1) + 1).astype(np.int64)


def make_buffer(cfg: Config, n=None, chunk=65536, alloc=None, feat_dtype=np.float32):
    """All rollout arrays except the old action probabilities (they need the policy).  ``alloc(shape, dtype)``
    optionally provides the output arrays (e.g. pinned host memory) so that large configs are generated in place.
    ``feat_dtype``: the features are small integers (vertex scores / degrees); np.int8 keeps them as the narrow integers
    ``ppo_buffer_append_i8`` takes (same values, same random stream)."""
    n = cfg.N if n is None else n
    rng = rng_for(cfg, 0)
    alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))
    feat = alloc((n, cfg.nhe, cfg.nf), feat_dtype)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        feat[s:e] = rng.integers(-3, 9, size=(e - s, cfg.nhe, cfg.nf), dtype=np.int8)
    mask = alloc((n, cfg.A), np.float32)
    act = alloc((n,), np.int64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        mask[s:e] = make_masks(rng, e - s, cfg.nhe, cfg.apa)
        act[s:e] = make_actions(rng, mask[s:e])
    reward = alloc((n,), np.float32)
    reward[:] = rng.integers(-4, 5, size=n)
    terminal = make_episode_terminals(rng, n, cfg.max_episode)
    return {"feat": feat, "mask": mask, "action": act, "reward": reward, "terminal": terminal}


def make_old_probs(cfg: Config, probs_of_actions):
    """old probability = the policy's own probability of the selected action at the initial weights
    x exp(N(0, 0.1)), clamped to (1e-6, 1] (exercises both clip branches)."""
    rng = rng_for(cfg, 1)
    p = np.asarray(probs_of_actions, np.float64) * np.exp(rng.normal(0.0, 0.1, size=len(probs_of_actions)))
    return np.clip(p, 1e-6, 1.0).astype(np.float32)


def make_weights(cfg: Config):
    """Glorot-uniform weights (Flux Dense default), zero biases; W[l] is [in, out]."""
    rng = rng_for(cfg, 2)
    W, b = [], []
    d = cfg.dims
    for i, o in zip(d[:-1], d[1:]):
        lim = np.sqrt(6.0 / (i + o))
        W.append(rng.uniform(-lim, lim, size=(i, o)).astype(np.float32))
        b.append(np.zeros(o, np.float32))
    return W, b
