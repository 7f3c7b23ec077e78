"""Device-resident rollout buffer and dataset — mirrors reference src/rollout_buffer.jl.

``DeviceRollouts`` is the third container type next to the reference's ``BufferRollouts``
(src/rollout_buffer.jl:1-22) and ``DiskRollouts`` (src/rollouts_to_disk.jl:1-5): same generic
functions (``update!``, ``length``, ``compute_state_value!``, ``collect_rollouts!``, ``permute!``,
``shuffle!``, ``construct_dataset``), storage as SoA arrays in HBM.  Shapes follow the quad-game
``StateData`` (test/quad_game_utilities.jl:17-20): vertex_score [nf, nhe], action_mask [A].
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import Context, default_context


class StateData:
    """``StateData(vertex_score, action_mask)`` — test/quad_game_utilities.jl:17-20.

    Arrays are C-ordered with Julia's trailing dimension first: a single state holds
    vertex_score [nhe, nf] (= Julia [nf, nhe]) and action_mask [A]; a batched state holds
    [nb, nhe, nf] and [nb, A]."""

    def __init__(self, vertex_score, action_mask):
        self.vertex_score = vertex_score
        self.action_mask = action_mask

    def __repr__(self):
        return "StateData"


def batch_state(state_data_vector):
    """``PPO.batch_state`` — test/quad_game_utilities.jl:26-33 (cat dims=3 / dims=2).  Bare-array states (as in
    the reference's disk-rollout test) are stacked on a new leading (= Julia trailing) dimension."""
    if not hasattr(state_data_vector[0], "vertex_score"):
        return np.stack([np.asarray(s) for s in state_data_vector])
    vs = np.stack([np.asarray(s.vertex_score) for s in state_data_vector])
    am = np.stack([np.asarray(s.action_mask) for s in state_data_vector])
    return StateData(vs, am)


def pad_vertex_scores(vertex_scores_vector, num_tokens=None):
    """``pad_vertex_scores`` — examples/triangle/distance_weighted/triangle_utilities.jl:31-37: states of meshes of
    different size are zero-padded to the largest token (half-edge) count (or to ``num_tokens``).  Arrays are
    [nhe_i, nf] here (= Julia [nf, nhe_i])."""
    vs = [np.asarray(v) for v in vertex_scores_vector]
    top = max(v.shape[0] for v in vs) if num_tokens is None else int(num_tokens)
    assert all(v.shape[0] <= top for v in vs)
    return [np.concatenate([v, np.zeros((top - v.shape[0],) + v.shape[1:], v.dtype)]) for v in vs]


def pad_action_mask(action_mask_vector, num_actions=None):
    """``pad_action_mask`` — triangle_utilities.jl:39-45: masks are padded with ``-Inf32`` (probability exactly 0)."""
    am = [np.asarray(m, np.float32).reshape(-1) for m in action_mask_vector]
    top = max(m.size for m in am) if num_actions is None else int(num_actions)
    assert all(m.size <= top for m in am)
    return [np.concatenate([m, np.full(top - m.size, -np.inf, np.float32)]) for m in am]


def prepare_state_data_for_batching_(state_data_vector, num_tokens=None, actions_per_token=None):
    """``PPO.prepare_state_data_for_batching!`` — triangle_utilities.jl:47-55 (in place)."""
    vs = pad_vertex_scores([s.vertex_score for s in state_data_vector], num_tokens)
    na = None if num_tokens is None or actions_per_token is None else int(num_tokens) * int(actions_per_token)
    am = pad_action_mask([s.action_mask for s in state_data_vector], na)
    state_data_vector[:] = [StateData(v, m) for v, m in zip(vs, am)]
    return state_data_vector


class DeviceRollouts:
    """``BufferRollouts()`` — src/rollout_buffer.jl:9-22 — on the device.

    Variable-size states: the buffer has a fixed token capacity ``nhe``; ``update_`` pads smaller
    states exactly like the reference's ``pad_vertex_scores`` / ``pad_action_mask`` (zeros / -Inf32), so the padded
    actions have probability exactly 0 and contribute nothing to the loss or the gradients.  (Padding is to the
    buffer's capacity instead of the largest state of each minibatch; the only visible difference is the ``s / A`` term
    of ``smoothed_entropy`` with s = 1f-8, below Float32 resolution.)

    ``update_`` stages transitions in a host chunk and appends them in batches (one H2D copy per
    chunk instead of five ``push!`` per transition).  ``capacity`` is an initial reservation: like the reference's
    ``push!``-grown vectors the device arrays grow (geometrically) when an append exceeds it."""

    def __init__(self, nf, nhe, apa, capacity, ctx: Context | None = None, chunk=4096):
        self.ctx = ctx or default_context()
        self.nf, self.nhe, self.apa, self.A = int(nf), int(nhe), int(apa), int(nhe) * int(apa)
        self.capacity = int(capacity)
        h = C.c_void_p()
        _lib.check(_lib.load().ppo_buffer_create(self.ctx.handle, self.capacity, self.nf, self.nhe, self.apa,
                                                 C.byref(h)))
        self._h = h
        self.ctx.adopt(self)
        self._chunk = int(chunk)
        self._pending = 0
        self._s_feat = np.empty((self._chunk, self.nhe, self.nf), np.float32)
        self._s_mask = np.empty((self._chunk, self.A), np.float32)
        self._s_act = np.empty(self._chunk, np.int64)
        self._s_prob = np.empty(self._chunk, np.float32)
        self._s_rew = np.empty(self._chunk, np.float32)
        self._s_term = np.empty(self._chunk, np.uint8)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("rollout buffer destroyed")
        return self._h

    # -- update! ------------------------------------------------------------------------------
    def update_(self, state, action_probability, action, reward, terminal):
        """``update!(episode, state, action_probability, action, reward, terminal)`` — :24-38."""
        if self._pending == self._chunk:      # a previous flush raised: try again before staging more
            self.flush()
        i = self._pending
        vs = np.asarray(state.vertex_score)
        am = np.asarray(state.action_mask).reshape(-1)
        if vs.size == self.nhe * self.nf and am.size == self.A:
            self._s_feat[i] = vs.reshape(self.nhe, self.nf)
            self._s_mask[i] = am
        else:       # a smaller state: pad like pad_vertex_scores / pad_action_mask
            vs = vs.reshape(-1, self.nf)
            assert vs.shape[0] <= self.nhe and am.size == vs.shape[0] * self.apa, \
                f"state with {vs.shape[0]} tokens / {am.size} actions does not fit a buffer of {self.nhe} x {self.apa}"
            self._s_feat[i] = 0
            self._s_feat[i, :vs.shape[0]] = vs
            self._s_mask[i] = -np.inf
            self._s_mask[i, :am.size] = am
        self._s_prob[i] = action_probability
        self._s_act[i] = action
        self._s_rew[i] = reward
        self._s_term[i] = 1 if terminal else 0
        self._pending += 1
        if self._pending == self._chunk:
            self.flush()

    def flush(self):
        n = self._pending
        if n:
            # (the staged transitions stay staged when the append raises, e.g. on a full buffer)
            self.append(self._s_feat[:n], self._s_mask[:n], self._s_prob[:n], self._s_act[:n], self._s_rew[:n],
                        self._s_term[:n])
            self._pending = 0

    def append(self, feat, mask, action_probability, action, reward, terminal):
        """Batched ``update!``: n transitions at once.  feat [n, nhe, nf] float32, int64 (the reference's
        ``Matrix{Int64}`` scores) or int8 / int16 (narrowed by the caller: exact, 4x / 2x fewer host->device bytes).
        mask: float32 [n, A] of 0 / -Inf, or packed bits (``pack_action_mask``: uint64 words, one bit per action,
        1 = allowed; 32x fewer bytes)."""
        feat = np.asarray(feat)
        n = feat.shape[0] if feat.ndim == 3 else feat.size // (self.nhe * self.nf)
        if isinstance(mask, np.ndarray) and mask.dtype == np.uint64:
            return self._append_packed(feat, n, mask, action_probability, action, reward, terminal)
        mask = np.ascontiguousarray(mask, np.float32).reshape(n, self.A)
        act = np.ascontiguousarray(action, np.int64).reshape(n)
        prob = np.ascontiguousarray(action_probability, np.float32).reshape(n)
        rew = np.ascontiguousarray(reward, np.float32).reshape(n)
        term = np.ascontiguousarray(np.asarray(terminal).astype(np.uint8)).reshape(n)
        lib = _lib.load()
        ints = {np.dtype(np.int64): (lib.ppo_buffer_append_i64, C.c_int64), np.dtype(np.int8): (lib.ppo_buffer_append_i8, C.c_int8),
                np.dtype(np.int16): (lib.ppo_buffer_append_i16, C.c_int16)}
        if feat.dtype in ints:
            fn, ct = ints[feat.dtype]
            f = np.ascontiguousarray(feat).reshape(n, self.nhe, self.nf)
            _lib.check(fn(self.handle, n, _lib.ptr(f, ct), _lib.ptr(mask, C.c_float), _lib.ptr(act, C.c_int64),
                          _lib.ptr(prob, C.c_float), _lib.ptr(rew, C.c_float), _lib.ptr(term, C.c_uint8)))
        else:
            f = np.ascontiguousarray(feat, np.float32).reshape(n, self.nhe, self.nf)
            _lib.check(lib.ppo_buffer_append(self.handle, n, _lib.ptr(f, C.c_float), _lib.ptr(mask, C.c_float),
                                             _lib.ptr(act, C.c_int64), _lib.ptr(prob, C.c_float),
                                             _lib.ptr(rew, C.c_float), _lib.ptr(term, C.c_uint8)))

    def _append_packed(self, feat, n, mask_bits, action_probability, action, reward, terminal):
        words = -(-n * self.A // 64)
        bits = np.ascontiguousarray(mask_bits, np.uint64).reshape(-1)
        assert bits.size == words, (bits.size, words)
        if feat.dtype not in (np.dtype(np.int8), np.dtype(np.int16), np.dtype(np.int64)):
            feat = feat.astype(np.float32, copy=False)
        f = np.ascontiguousarray(feat).reshape(n, self.nhe, self.nf)
        act = np.ascontiguousarray(action, np.int64).reshape(n)
        prob = np.ascontiguousarray(action_probability, np.float32).reshape(n)
        rew = np.ascontiguousarray(reward, np.float32).reshape(n)
        term = np.ascontiguousarray(np.asarray(terminal).astype(np.uint8)).reshape(n)
        _lib.check(_lib.load().ppo_buffer_append_packed(
            self.handle, n, f.ctypes.data_as(C.c_void_p), f.dtype.itemsize, _lib.ptr(bits, C.c_uint64),
            _lib.ptr(act, C.c_int64), _lib.ptr(prob, C.c_float), _lib.ptr(rew, C.c_float), _lib.ptr(term, C.c_uint8)))

    def __len__(self):
        """``Base.length`` — :40-48."""
        return int(_lib.load().ppo_buffer_length(self.handle)) + self._pending

    def clear(self):
        self._pending = 0
        _lib.check(_lib.load().ppo_buffer_clear(self.handle))

    # -- reads (tests, evaluators) --------------------------------------------------------------
    def read(self, start=0, count=None):
        self.flush()
        n = len(self)
        count = n - start if count is None else count
        feat = np.empty((count, self.nhe, self.nf), np.float32)
        mask = np.empty((count, self.A), np.float32)
        act = np.empty(count, np.int64)
        prob = np.empty(count, np.float32)
        rew = np.empty(count, np.float32)
        term = np.empty(count, np.uint8)
        _lib.check(_lib.load().ppo_buffer_read(self.handle, start, count, _lib.ptr(feat, C.c_float),
                                               _lib.ptr(mask, C.c_float), _lib.ptr(act, C.c_int64),
                                               _lib.ptr(prob, C.c_float), _lib.ptr(rew, C.c_float),
                                               _lib.ptr(term, C.c_uint8)))
        return {"feat": feat, "mask": mask, "selected_actions": act, "selected_action_probabilities": prob,
                "rewards": rew, "terminal": term.astype(bool)}

    @property
    def rewards(self):
        """``rollouts.rewards`` (returns after ``compute_state_value_``)."""
        self.flush()
        n = len(self)
        rew = np.empty(n, np.float32)
        _lib.check(_lib.load().ppo_buffer_read(self.handle, 0, n, None, None, None, None, _lib.ptr(rew, C.c_float),
                                               None))
        return rew

    def save_rewards(self):
        self.flush()
        _lib.check(_lib.load().ppo_buffer_save_rewards(self.handle))

    def restore_rewards(self):
        _lib.check(_lib.load().ppo_buffer_restore_rewards(self.handle))

    # -- extension -----------------------------------------------------------------------------
    def normalize_advantage(self, enable=True, eps=1e-8):
        _lib.check(_lib.load().ppo_normalize_advantage(self.handle, int(bool(enable)), float(eps)))

    # -- permutations for the dataset -------------------------------------------------------------
    def set_permutation(self, perm1):
        p = np.ascontiguousarray(perm1, np.int64)
        _lib.check(_lib.load().ppo_permutation_set(self.handle, _lib.ptr(p, C.c_int64), p.size))

    def generate_permutation(self, seed, want=False):
        self.flush()
        out = np.empty(len(self), np.int64) if want else None
        _lib.check(_lib.load().ppo_permutation_generate(self.handle, int(seed) & (2 ** 64 - 1),
                                                        _lib.ptr(out, C.c_int64)))
        return out

    def close(self):
        if self._h is not None:
            _lib.load().ppo_buffer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        return f"EpisodeData\n\t{len(self)} data points\n"


def update_(rollouts, state, action_probability, action, reward, terminal):
    return rollouts.update_(state, action_probability, action, reward, terminal)


def length(x):
    return len(x)


def compute_state_value_(rollouts: DeviceRollouts, discount):
    """``compute_state_value!(rollouts, discount)`` — src/rollout_buffer.jl:55-64: the rewards are
    overwritten in place by ``compute_returns`` (src/collect_rollouts.jl:26-42), on the device.
    A Python float is a Float64 discount (Float64 carry); pass ``numpy.float32`` for a Float32 one."""
    rollouts.flush()
    is_f32 = isinstance(discount, np.float32)
    _lib.check(_lib.load().ppo_compute_returns(rollouts.handle, float(discount), int(is_f32)))


def permute_(rollouts: DeviceRollouts, idx):
    """``permute!(rollouts, idx)`` — :81-88; idx 1-based, ``length(idx) == length(rollouts)``."""
    rollouts.flush()
    p = np.ascontiguousarray(idx, np.int64)
    _lib.check(_lib.load().ppo_buffer_permute(rollouts.handle, _lib.ptr(p, C.c_int64), p.size))


def shuffle_(rollouts: DeviceRollouts, seed=0):
    """``shuffle!(rollouts)`` — :90-93, with the device permutation of ``seed``."""
    rollouts.flush()
    _lib.check(_lib.load().ppo_buffer_shuffle(rollouts.handle, int(seed) & (2 ** 64 - 1)))


class DeviceDataset:
    """``BufferDataset`` — src/rollout_buffer.jl:95-101: a zero-cost view of the buffer."""

    def __init__(self, rollouts: DeviceRollouts):
        rollouts.flush()
        self.rollouts = rollouts

    def __len__(self):
        return len(self.rollouts)

    def __getitem__(self, idx):
        """``Base.getindex`` — :135-143."""
        if isinstance(idx, (int, np.integer)):
            return get_sample(self, int(idx))
        if isinstance(idx, (list, tuple, np.ndarray)):
            return get_batch(self, idx)
        raise TypeError(f"Dataset index should be Int or Array, got {type(idx)}")


def pack_action_mask(mask, out=None):
    """0 / -Inf action masks [n, A] -> one bit per action (1 = allowed), little-endian bits in uint64 words: the layout
    of ``BitMatrix(isfinite.(action_mask)).chunks`` in Julia, accepted by ``DeviceRollouts.append`` /
    ``ppo_buffer_append_packed``."""
    allowed = np.isfinite(np.asarray(mask)).reshape(-1)
    words = -(-allowed.size // 64)
    by = np.packbits(allowed, bitorder="little")
    if out is None:
        out = np.zeros(words, np.uint64)
    out.view(np.uint8)[:by.size] = by
    out.view(np.uint8)[by.size:] = 0
    return out


def get_sample(dataset: DeviceDataset, idx):
    """``get_sample`` — :103-115 (1-based index)."""
    n = len(dataset)
    assert isinstance(idx, (int, np.integer))
    assert 1 <= idx <= n
    b = get_batch(dataset, [idx])
    return {"state": StateData(b["state"].vertex_score[0], b["state"].action_mask[0]),
            "selected_action": int(b["selected_action"][0]),
            "selected_action_probability": np.float32(b["selected_action_probability"][0]),
            "returns": np.float32(b["returns"][0])}


def get_batch(dataset: DeviceDataset, indices):
    """``get_batch`` — :117-133: the gather runs on the device (K4), the 4-key Dict comes back to
    the host.  ``indices`` are 1-based."""
    r = dataset.rollouts
    idx = np.ascontiguousarray(indices, np.int64)
    assert idx.ndim == 1
    nb = idx.size
    feat = np.empty((nb, r.nhe, r.nf), np.float32)
    mask = np.empty((nb, r.A), np.float32)
    act = np.empty(nb, np.int64)
    prob = np.empty(nb, np.float32)
    ret = np.empty(nb, np.float32)
    _lib.check(_lib.load().ppo_gather_indices(r.handle, _lib.ptr(idx, C.c_int64), nb, _lib.ptr(feat, C.c_float),
                                              _lib.ptr(mask, C.c_float), _lib.ptr(act, C.c_int64),
                                              _lib.ptr(prob, C.c_float), _lib.ptr(ret, C.c_float)))
    return {"state": StateData(feat, mask), "selected_action": act, "selected_action_probability": prob,
            "returns": ret}


def gather_minibatch(dataset: DeviceDataset, start, count):
    """rows perm[start : start+count] of the current permutation (0-based start) -> host Dict."""
    r = dataset.rollouts
    feat = np.empty((count, r.nhe, r.nf), np.float32)
    mask = np.empty((count, r.A), np.float32)
    act = np.empty(count, np.int64)
    prob = np.empty(count, np.float32)
    ret = np.empty(count, np.float32)
    _lib.check(_lib.load().ppo_gather(r.handle, int(start), int(count), _lib.ptr(feat, C.c_float),
                                      _lib.ptr(mask, C.c_float), _lib.ptr(act, C.c_int64),
                                      _lib.ptr(prob, C.c_float), _lib.ptr(ret, C.c_float)))
    return {"state": StateData(feat, mask), "selected_action": act, "selected_action_probability": prob,
            "returns": ret}


def construct_dataset(rollouts: DeviceRollouts):
    """``construct_dataset`` — :145-147."""
    return DeviceDataset(rollouts)
