"""Rollout collection — mirrors reference src/collect_rollouts.jl.

The environment is opaque host user code (sequential by construction), so stepping stays on the
host exactly as in the reference; only the storage (``update_``) and ``compute_returns`` (the
first stage of the hot path) run on the device.

Environment / policy protocol (reference src/ProximalPolicyOptimization.jl:16-30): an env object
provides ``state()``, ``reward()``, ``is_terminal()``, ``reset_()``, ``step_(action)``; the policy
hook is ``action_probabilities(policy, state)`` (policy.py provides it for the MLP policy, any
object with an ``action_probabilities(state)`` method overrides it).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import default_context
from .policy import action_probabilities
from .rollout_buffer import compute_state_value_

_rng = np.random.default_rng()


def seed_sampling(seed):
    """Seed the host RNG used for ``rand(Categorical(ap))``."""
    global _rng
    _rng = np.random.default_rng(seed)


def sample_categorical(ap):
    """``rand(Categorical(ap))`` — 1-based action."""
    ap = np.asarray(ap, dtype=np.float64)
    return int(_rng.choice(ap.size, p=ap / ap.sum())) + 1


def collect_step_data_(buffer, env, policy):
    """``collect_step_data!`` — src/collect_rollouts.jl:1-15."""
    cpu_state = env.state()
    ap = np.asarray(action_probabilities(policy, cpu_state))
    a = sample_categorical(ap)
    assert ap[a - 1] > 0.0
    env.step_(a)
    r = env.reward()
    t = env.is_terminal()
    buffer.update_(cpu_state, ap[a - 1], a, r, t)


def collect_episode_data_(episode_data, env, policy):
    """``collect_episode_data!`` — src/collect_rollouts.jl:17-24."""
    terminal = env.is_terminal()
    while not terminal:
        collect_step_data_(episode_data, env, policy)
        terminal = env.is_terminal()


def collect_rollouts_(rollouts, env, policy, num_episodes, discount):
    """``collect_rollouts!(rollouts::BufferRollouts, ...)`` — src/rollout_buffer.jl:66-79."""
    for _ in range(num_episodes):
        env.reset_()
        collect_episode_data_(rollouts, env, policy)
    compute_state_value_(rollouts, discount)


def compute_returns(rewards, terminal, discount, ctx=None):
    """``compute_returns(rewards, terminal, discount)`` — src/collect_rollouts.jl:26-42 — on host
    vectors: uploaded to a scratch device buffer, scanned by K1, returned as Float32."""
    from .rollout_buffer import DeviceRollouts
    r = np.ascontiguousarray(rewards, np.float32)
    n = r.size
    if n == 0:
        return r.copy()
    buf = DeviceRollouts(1, 1, 1, n, ctx or default_context())
    try:
        buf.append(np.zeros((n, 1, 1), np.float32), np.zeros((n, 1), np.float32), np.ones(n, np.float32),
                   np.ones(n, np.int64), r, np.asarray(terminal).astype(np.uint8))
        compute_state_value_(buf, discount)
        return buf.rewards
    finally:
        buf.close()
