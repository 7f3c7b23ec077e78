"""Disk rollouts / replay loader: the reference's wire format, pinned against its fixtures."""
import ctypes as C
import csv
import os

import numpy as np
import pytest

import ppo_b200 as P
from ppo_b200 import _lib, bson_io
from ppo_b200 import synthetic as S

G = os.path.join(os.path.dirname(__file__), "golden")


def test_writer_reproduces_reference_bson_bytes(tmp_path):
    # reference test/test_rollout_to_disk.jl:17-23: update!(..., [1,2,3,4,5], ...) -> states/sample_1.bson
    want = bytes.fromhex(open(os.path.join(G, "reference_output_sample_1.bson.hex")).read().strip())
    path = tmp_path / "s.bson"
    bson_io.save_state(str(path), np.array([1, 2, 3, 4, 5], np.int64))
    assert path.read_bytes() == want
    assert np.array_equal(bson_io.load_state(str(path)), [1, 2, 3, 4, 5])


def test_cpp_parser_reads_reference_fixture(tmp_path):
    raw = bytes.fromhex(open(os.path.join(G, "reference_rollout_to_disk_sample_1.bson.hex")).read().strip())
    path = tmp_path / "ref.bson"
    path.write_bytes(raw)
    el = C.create_string_buffer(16 * 4)
    counts = (C.c_int64 * 4)(); ndims = (C.c_int * 4)(); dims = (C.c_int64 * 16)(); n = C.c_int()
    _lib.check(_lib.load().ppo_bson_state_arrays(str(path).encode(), 4, el, counts, ndims, dims, C.byref(n)))
    assert n.value == 1 and el.raw[:5] == b"Int64" and counts[0] == 5 and ndims[0] == 1 and dims[0] == 5
    # a lowered StateData: two arrays, in field order
    st = P.StateData(np.arange(12, dtype=np.int64).reshape(3, 4), np.array([0, -np.inf, 0, 0], np.float32))
    p2 = tmp_path / "sd.bson"
    bson_io.save_state(str(p2), st)
    _lib.check(_lib.load().ppo_bson_state_arrays(str(p2).encode(), 4, el, counts, ndims, dims, C.byref(n)))
    assert n.value == 2 and el.raw[:5] == b"Int64" and el.raw[16:23] == b"Float32"
    assert (dims[0], dims[1], counts[0], counts[1]) == (4, 3, 12, 4)      # Julia size [nf, nhe]
    back = bson_io.load_state(str(p2))
    assert np.array_equal(back[0], st.vertex_score) and np.array_equal(back[1], st.action_mask)


def test_disk_rollouts_layout_and_csv_schema(tmp_path):
    # reference test/test_rollout_to_disk.jl:8-23: constructing wipes the directory, creates states/, update! writes
    d = tmp_path / "rollouts"
    d.mkdir()
    (d / "test.txt").write_text("hello")
    traj = P.DiskRollouts(str(d))
    assert not (d / "test.txt").exists() and (d / "states").is_dir()
    traj.update_(np.array([1, 2, 3, 4, 5], np.int64), 0.2, 1, 0.5, False)
    assert (d / "states" / "sample_1.bson").is_file()
    assert np.array_equal(bson_io.load_state(str(d / "states" / "sample_1.bson")), [1, 2, 3, 4, 5])
    lines = (d / "trajectory.csv").read_text().splitlines()
    assert lines[0] == "sample_names,selected_actions,selected_action_probabilities,rewards,terminal"
    assert lines[1] == "sample_1.bson,1,0.2,0.5,false"


def test_reference_trajectory_csv_is_readable(tmp_path):
    # the reference's own output/trajectory.csv (returns schema) + states written in its format
    root = tmp_path / "out"
    (root / "states").mkdir(parents=True)
    (root / "trajectory.csv").write_text(open(os.path.join(G, "reference_trajectory.csv")).read())
    for i in range(1, 7):
        bson_io.save_state(str(root / "states" / f"sample_{i}.bson"), np.array([i, 0, 0], np.int64))
    ds = P.DiskDataset(str(root))
    assert len(ds) == 6
    s = ds[2]
    assert s["selected_action"] == 4 and s["selected_action_probability"] == np.float32(0.5) and s["returns"] == 5.0
    b = ds[[1, 6]]
    assert b["returns"].tolist() == [6.0, 1.0]


def _write_replay(root, cfg, n, seed=3):
    data = S.make_buffer(cfg, n)
    old = np.random.default_rng(seed).uniform(0.05, 1.0, n).astype(np.float32)
    traj = P.DiskRollouts(root)
    for i in range(n):
        st = P.StateData(data["feat"][i].astype(np.int64), data["mask"][i])
        traj.update_(st, float(old[i]), int(data["action"][i]), float(data["reward"][i]), bool(data["terminal"][i]))
    return data, old, traj


def test_disk_dataset_host_getindex(tmp_path):
    cfg = S.CONFIGS["t0"]
    data, old, traj = _write_replay(str(tmp_path / "r"), cfg, 40)
    ds = P.DiskDataset(str(tmp_path / "r"))
    assert len(ds) == 40
    s = ds[3]
    assert s["selected_action"] == data["action"][2] and np.array_equal(s["state"].vertex_score, data["feat"][2])
    b = ds[[1, 40, 7]]
    assert np.array_equal(b["state"].vertex_score, data["feat"][[0, 39, 6]])
    assert np.array_equal(b["state"].action_mask, data["mask"][[0, 39, 6]])
    assert np.array_equal(b["selected_action_probability"], old[[0, 39, 6]])


@pytest.mark.gpu
def test_replay_into_device_buffer_and_returns(ctx, tmp_path):
    from oracle import ppo_oracle as O
    cfg = S.CONFIGS["t0"]
    n = 120
    data, old, traj = _write_replay(str(tmp_path / "r"), cfg, n)
    # schema 1 (rewards, terminal): load, then compute_state_value! on the device
    ds = P.DiskDataset(str(tmp_path / "r"))
    buf, has_ret = ds.to_device(cfg.nf, cfg.nhe, cfg.apa, ctx, n_threads=4)
    assert not has_ret and len(buf) == n
    back = buf.read()
    assert np.array_equal(back["feat"], data["feat"][:n]) and np.array_equal(back["mask"], data["mask"][:n])
    assert np.array_equal(back["selected_actions"], data["action"][:n])
    assert np.array_equal(back["selected_action_probabilities"], old)
    assert np.array_equal(back["terminal"], data["terminal"][:n])
    P.compute_state_value_(buf, 1.0)
    want = O.compute_returns(data["reward"][:n], data["terminal"][:n], 1.0)
    assert np.array_equal(buf.rewards, want)
    buf.close()
    # schema 2 (returns): write_returns_to_disk (device scan), reload: returns column goes straight into the buffer
    P.write_returns_to_disk(traj, 1.0, ctx)
    rows = list(csv.DictReader(open(traj.trajectory_filename)))
    assert list(rows[0].keys()) == ["sample_names", "selected_actions", "selected_action_probabilities", "returns"]
    buf2, has_ret2 = P.DiskDataset(str(tmp_path / "r")).to_device(cfg.nf, cfg.nhe, cfg.apa, ctx, n_threads=2)
    assert has_ret2 and np.array_equal(buf2.rewards, want)
    buf2.close()
