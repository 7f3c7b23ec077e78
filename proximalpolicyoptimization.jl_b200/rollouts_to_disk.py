"""Spill-to-disk rollouts — mirrors reference src/rollouts_to_disk.jl (host file I/O in the reference's format).

Only used to produce / replay the reference's on-disk data (config C5); the PPO update itself never touches the disk:
``construct_dataset(DiskRollouts)`` yields a ``DiskDataset`` whose ``to_device`` bulk-loads everything into a
``DeviceRollouts``.
"""
from __future__ import annotations

import csv
import os
import shutil

import numpy as np

from . import bson_io

HEADER = ["sample_names", "selected_actions", "selected_action_probabilities", "rewards", "terminal"]
RETURNS_HEADER = ["sample_names", "selected_actions", "selected_action_probabilities", "returns"]


def _julia_float(x):
    """Julia's shortest round-trip printing of a Float32/Float64 (what CSV.jl writes): 6.0, 0.2, 1.0e-5, 123456.7,
    1.0e6, 1.234567e6 -- plain notation for 1e-5 < |x| < 1e6, exponent notation outside."""
    r = str(x) if isinstance(x, np.float32) else repr(float(x))
    if r in ("inf", "-inf", "nan"):
        return {"inf": "Inf", "-inf": "-Inf", "nan": "NaN"}[r]
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    if "." not in r:
        r += ".0"
    if abs(float(x)) >= 1e6:                     # Python keeps plain notation up to 1e16, Julia only below 1e6
        sign = "-" if r.startswith("-") else ""
        ip, fp = r.lstrip("-").split(".")
        digits = (ip + fp).rstrip("0") or "0"
        return f"{sign}{digits[0]}.{digits[1:] or '0'}e{len(ip) - 1}"
    return r


def _fmt(x):
    """One CSV.jl cell."""
    if isinstance(x, (bool, np.bool_)):
        return "true" if x else "false"
    if isinstance(x, (float, np.floating)):
        return _julia_float(x)
    return str(x)


class DiskRollouts:
    """``DiskRollouts(state_data_dir)`` — src/rollouts_to_disk.jl:1-45: wipes the directory, creates ``states/`` and a
    ``trajectory.csv`` holding only the header."""

    def __init__(self, state_data_directory):
        self.state_data_directory = state_data_directory
        self.num_samples = 0
        if os.path.isdir(state_data_directory):            # prepare_state_data_directory, :7-13
            shutil.rmtree(state_data_directory)
        os.makedirs(os.path.join(state_data_directory, "states"))
        self.trajectory_filename = os.path.join(state_data_directory, "trajectory.csv")
        with open(self.trajectory_filename, "w", newline="") as f:
            f.write(",".join(HEADER) + "\n")

    def update_(self, state, action_probability, action, reward, terminal):
        """``update!`` — :73-95 (the sample is named after the incremented counter: 1-based)."""
        assert 0 <= action_probability <= 1
        assert isinstance(terminal, (bool, np.bool_))
        self.num_samples += 1
        sample_name = f"sample_{self.num_samples}.bson"
        bson_io.save_state(os.path.join(self.state_data_directory, "states", sample_name), state)
        with open(self.trajectory_filename, "a", newline="") as f:
            f.write(",".join([sample_name, _fmt(int(action)), _fmt(action_probability), _fmt(reward),
                              _fmt(bool(terminal))]) + "\n")

    def __len__(self):
        return self.num_samples

    def __repr__(self):
        return f"Rollouts\n\t{len(self)} data points\n"


def write_returns_to_disk(buffer: DiskRollouts, discount, ctx=None):
    """``write_returns_to_disk`` — :106-132: re-read the CSV, run compute_returns (on the device), rewrite it with the
    4-column returns schema."""
    from .collect_rollouts import compute_returns
    rows = list(csv.DictReader(open(buffer.trajectory_filename)))
    rewards = np.array([float(r["rewards"]) for r in rows], np.float32)
    terminal = np.array([r["terminal"] == "true" for r in rows], bool)
    returns = compute_returns(rewards, terminal, discount, ctx)
    with open(buffer.trajectory_filename, "w", newline="") as f:
        f.write(",".join(RETURNS_HEADER) + "\n")
        for r, ret in zip(rows, returns):
            f.write(",".join([r["sample_names"], r["selected_actions"], r["selected_action_probabilities"],
                              _fmt(np.float32(ret))]) + "\n")


def collect_rollouts_disk_(buffer: DiskRollouts, env, policy, num_episodes, discount, out=None):
    """``collect_rollouts!(buffer::DiskRollouts, ...)`` — :134-147."""
    from .collect_rollouts import collect_episode_data_
    if out is not None:
        out.write("\n\nCOLLECTING ROLLOUTS :\n")
    for _ in range(num_episodes):
        env.reset_()
        collect_episode_data_(buffer, env, policy)
    write_returns_to_disk(buffer, discount)
