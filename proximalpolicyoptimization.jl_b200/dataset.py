"""Disk dataset — mirrors reference src/dataset.jl, plus the bulk replay loader into the device buffer."""
from __future__ import annotations

import csv
import ctypes as C
import os

import numpy as np

from . import _lib, bson_io
from .rollout_buffer import DeviceRollouts, StateData, batch_state


class DiskDataset:
    """``DiskDataset(root_directory, trajectory_filename, states_dirname)`` — src/dataset.jl:1-20."""

    def __init__(self, root_directory, trajectory_filename="trajectory.csv", states_dirname="states"):
        self.root_directory = root_directory
        self.trajectory_filename, self.states_dirname = trajectory_filename, states_dirname
        trajectory_filepath = os.path.join(root_directory, trajectory_filename)
        assert os.path.isfile(trajectory_filepath)
        self.trajectory_df = list(csv.DictReader(open(trajectory_filepath)))
        self.states_directory = os.path.join(root_directory, states_dirname)
        assert os.path.isdir(self.states_directory)

    def __len__(self):
        return len(self.trajectory_df)

    def __repr__(self):
        return f"Dataset\n\t{len(self)} data points\n"

    def __getitem__(self, idx):
        """``Base.getindex`` — :74-82."""
        if isinstance(idx, (int, np.integer)):
            return load_sample(self, int(idx))
        if isinstance(idx, (list, tuple, np.ndarray)):
            return load_batch(self, idx)
        raise TypeError(f"Dataset index should be Int or Array, got {type(idx)}")

    def to_device(self, nf, nhe, apa, ctx=None, capacity=None, n_threads=None):
        """Bulk replay: every row + state file into a DeviceRollouts (C++ loader, threaded file reads).
        Returns (rollouts, has_returns)."""
        n = len(self)
        rollouts = DeviceRollouts(nf, nhe, apa, capacity or max(n, 1), ctx)
        loaded, has_ret = C.c_int64(), C.c_int()
        _lib.check(_lib.load().ppo_disk_dataset_load(rollouts.handle, self.root_directory.encode(),
                                                     self.trajectory_filename.encode(), self.states_dirname.encode(),
                                                     int(n_threads or min(32, os.cpu_count() or 1)),
                                                     C.byref(loaded), C.byref(has_ret)))
        assert loaded.value == n
        return rollouts, bool(has_ret.value)


def _to_state(raw):
    if isinstance(raw, list):          # lowered struct: [vertex_score, action_mask]
        return StateData(raw[0], raw[1])
    return raw


def load_sample(dataset: DiskDataset, idx):
    """``load_sample`` — :31-52 (1-based)."""
    assert isinstance(idx, (int, np.integer))
    assert 1 <= idx <= len(dataset)
    row = dataset.trajectory_df[idx - 1]
    state_filepath = os.path.join(dataset.states_directory, row["sample_names"])
    assert os.path.isfile(state_filepath)
    value_key = "returns" if "returns" in row else "rewards"
    return {"state": _to_state(bson_io.load_state(state_filepath)), "selected_action": int(row["selected_actions"]),
            "selected_action_probability": np.float32(row["selected_action_probabilities"]),
            "returns": np.float32(row[value_key])}


def load_batch(dataset: DiskDataset, indices):
    """``load_batch`` — :54-72."""
    samples = [dataset[int(i)] for i in indices]
    return {"state": batch_state([s["state"] for s in samples]),
            "selected_action": np.array([s["selected_action"] for s in samples], np.int64),
            "selected_action_probability": np.array([s["selected_action_probability"] for s in samples], np.float32),
            "returns": np.array([s["returns"] for s in samples], np.float32)}


def construct_disk_dataset(rollouts):
    """``construct_dataset(rollouts::DiskRollouts)`` — src/rollouts_to_disk.jl:169-171."""
    return DiskDataset(rollouts.state_data_directory)
