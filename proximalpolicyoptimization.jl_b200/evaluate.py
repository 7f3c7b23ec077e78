"""Evaluation — mirrors reference src/evaluate.jl (host loop over the env; not on the hot path)."""
from __future__ import annotations

import numpy as np

from .collect_rollouts import sample_categorical
from .policy import action_probabilities


def single_trajectory_return(policy, env):
    """``single_trajectory_return`` — src/evaluate.jl:1-16."""
    ret = 0
    done = env.is_terminal()
    while not done:
        probs = np.asarray(action_probabilities(policy, env.state()))
        action = sample_categorical(probs)
        assert probs[action - 1] > 0.0
        env.step_(action)
        done = env.is_terminal()
        ret += env.reward()
    return ret


def average_returns(policy, env, num_trajectories):
    """``average_returns`` — src/evaluate.jl:18-25 -> (mean, sample std with n-1)."""
    ret = np.zeros(num_trajectories)
    for idx in range(num_trajectories):
        env.reset_()
        ret[idx] = single_trajectory_return(policy, env)
    return float(ret.mean()), float(ret.std(ddof=1)) if num_trajectories > 1 else float("nan")
