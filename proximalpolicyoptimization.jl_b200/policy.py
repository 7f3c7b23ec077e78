"""Policy MLP and optimiser handles — mirrors reference test/policy.jl and the Flux pieces the
package touches (``Flux.params``, ``Flux.Optimise.Adam``, ``Flux.Optimise.Optimiser``).

``Policy(in, hidden, num_hidden_layers, num_output)`` = ``SimplePolicy.Policy`` (test/policy.jl:9-21):
``Chain(Dense(in,h,leakyrelu), (L-1) x Dense(h,h,leakyrelu), Dense(h,out))``; weights live on the
device; ``weights()`` / ``load_weights()`` round-trip them as numpy arrays laid out like Julia's
(W[l] has shape [in, out] in C order = the bytes of Dense.weight [out, in]).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .context import Context, default_context

LEAKY_SLOPE = 0.01  # NNlib.leakyrelu default


def _pp(arrays):
    """float** from a list of float32 arrays."""
    arr = (C.POINTER(C.c_float) * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data_as(C.POINTER(C.c_float))
    return arr


class Policy:
    def __init__(self, in_channels, hidden_channels, num_hidden_layers, num_output, ctx: Context | None = None,
                 rng=None, weights=None, biases=None, leaky_slope=LEAKY_SLOPE, gemm_mode=None):
        self.ctx = ctx or default_context()
        self.hidden_channels, self.num_hidden_layers = hidden_channels, num_hidden_layers
        self.dims = [int(in_channels)] + [int(hidden_channels)] * int(num_hidden_layers) + [int(num_output)]
        L = len(self.dims) - 1
        if weights is None:
            # Flux Dense default: glorot_uniform weights, zero bias
            rng = rng if rng is not None else np.random.default_rng(0)
            weights, biases = [], []
            for i, o in zip(self.dims[:-1], self.dims[1:]):
                lim = np.sqrt(6.0 / (i + o))
                weights.append(rng.uniform(-lim, lim, size=(i, o)).astype(np.float32))
                biases.append(np.zeros(o, np.float32))
        W = [np.ascontiguousarray(w, np.float32) for w in weights]
        b = [np.ascontiguousarray(x, np.float32) for x in biases]
        for l in range(L):
            assert W[l].shape == (self.dims[l], self.dims[l + 1]), (l, W[l].shape)
            assert b[l].shape == (self.dims[l + 1],)
        dims = (C.c_int * (L + 1))(*self.dims)
        h = C.c_void_p()
        _lib.check(_lib.load().ppo_policy_create(self.ctx.handle, L, dims, _pp(W), _pp(b), float(leaky_slope),
                                                 C.byref(h)))
        self._h = h
        self._optimisers = weakref.WeakSet()
        self.ctx.adopt(self)
        # the library's default engine is GEMM_AUTO (fastest fp32-parity engine whose shape contract the policy meets)
        if gemm_mode is not None:
            self.set_gemm_mode(gemm_mode)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("policy destroyed")
        return self._h

    @property
    def num_params(self):
        return int(_lib.load().ppo_policy_num_params(self.handle))

    def weights(self):
        """(W list, b list) copied back from the device (``Flux.params`` order)."""
        W = [np.empty((i, o), np.float32) for i, o in zip(self.dims[:-1], self.dims[1:])]
        b = [np.empty(o, np.float32) for o in self.dims[1:]]
        _lib.check(_lib.load().ppo_policy_read(self.handle, _pp(W), _pp(b)))
        return W, b

    def load_weights(self, W, b):
        W = [np.ascontiguousarray(w, np.float32) for w in W]
        b = [np.ascontiguousarray(x, np.float32) for x in b]
        _lib.check(_lib.load().ppo_policy_write(self.handle, _pp(W), _pp(b)))

    def set_gemm_mode(self, mode: int):
        """GEMM engine of the MLP: GEMM_AUTO (-1) picks the fastest fp32-parity engine whose shape contract the policy
        meets (fp16-split tcgen05, else 3xTF32 tcgen05, else fp32 FFMA); returns the engine in use."""
        _lib.check(_lib.load().ppo_policy_set_gemm_mode(self.handle, int(mode)))
        return self.gemm_mode

    def read_gates(self, layer: int, rows: int):
        """leakyrelu' branches (1 = positive) the last backward pass applied to hidden activation ``layer``
        (1..L-1): uint8 [rows, dims[layer]] (parity instrumentation, see include/ppo_b200.h)."""
        out = np.empty((int(rows), self.dims[layer]), np.uint8)
        _lib.check(_lib.load().ppo_policy_read_gates(self.handle, int(layer), int(rows), _lib.ptr(out, C.c_uint8)))
        return out

    def set_token_compaction(self, enable: bool):
        """fp16-split engine: run the MLP only on tokens with at least one unmasked action (default on).  Fully masked
        tokens have probability exactly 0 and gradient exactly 0, so losses / probabilities / gradients are those of
        the dense evaluation (up to the weight gradient's summation order)."""
        _lib.check(_lib.load().ppo_policy_set_token_compaction(self.handle, 1 if enable else 0))

    def active_tokens(self) -> int:
        """tokens the last forward pass ran; -1 when it ran every token of the minibatch"""
        out = C.c_int64(-1)
        _lib.check(_lib.load().ppo_policy_active_tokens(self.handle, C.byref(out)))
        return int(out.value)

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        _lib.check(_lib.load().ppo_policy_p2p_export(self.handle, buf))
        return buf.raw

    def p2p_connect(self, nranks: int, rank: int, handles: bytes):
        assert len(handles) == 64 * nranks
        _lib.check(_lib.load().ppo_policy_p2p_connect(self.handle, int(nranks), int(rank), C.c_char_p(handles)))

    def p2p_wait(self, reset: bool = False):
        """(total ns, waits): time this rank's optimiser kernel spent waiting for its peers' gradients"""
        ns, n = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.load().ppo_policy_p2p_wait(self.handle, C.byref(ns), C.byref(n), 1 if reset else 0))
        return int(ns.value), int(n.value)

    @property
    def gemm_mode(self) -> int:
        return int(_lib.load().ppo_policy_get_gemm_mode(self.handle))

    def __call__(self, state):
        raise NotImplementedError("use batch_action_probabilities(policy, state)")

    def close(self):
        if self._h is not None:
            for o in list(self._optimisers):     # optimiser state is keyed by this policy's parameters
                o.close()
            _lib.load().ppo_policy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        return f"Policy\n\t{self.hidden_channels} channels\n\t{self.num_hidden_layers} layers\n"


def batch_action_probabilities(policy, state):
    """``PPO.batch_action_probabilities(policy, state)`` — test/quad_game_utilities.jl:73-79.
    state.vertex_score [nb, nhe, nf], state.action_mask [nb, A] -> probs [nb, A] (= Julia [A, nb])."""
    if hasattr(policy, "batch_action_probabilities"):
        return policy.batch_action_probabilities(state)
    feat = np.ascontiguousarray(state.vertex_score, np.float32)
    mask = np.ascontiguousarray(state.action_mask, np.float32)
    nb, nhe = feat.shape[0], feat.shape[1]
    probs = np.empty_like(mask)
    _lib.check(_lib.load().ppo_batch_action_probabilities(policy.handle, nb, nhe, _lib.ptr(feat, C.c_float),
                                                          _lib.ptr(mask, C.c_float), _lib.ptr(probs, C.c_float)))
    return probs


def batch_sample_actions(policy, state, seed, return_probabilities=False):
    """Batched rollout inference (extension): what ``collect_step_data!``
    (src/collect_rollouts.jl:1-15) does for one state -- ``action_probabilities`` then ``rand(Categorical(ap))`` --
    for a whole batch of states on the device.  Returns (selected_actions Int64 1-based [nb],
    selected_action_probabilities Float32 [nb]) and, optionally, all probabilities [nb, A]."""
    feat = np.ascontiguousarray(state.vertex_score, np.float32)
    mask = np.ascontiguousarray(state.action_mask, np.float32)
    nb, nhe = feat.shape[0], feat.shape[1]
    actions = np.empty(nb, np.int64)
    prob = np.empty(nb, np.float32)
    probs = np.empty_like(mask) if return_probabilities else None
    _lib.check(_lib.load().ppo_sample_actions(policy.handle, nb, nhe, _lib.ptr(feat, C.c_float), _lib.ptr(mask, C.c_float),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(actions, C.c_int64),
                                              _lib.ptr(prob, C.c_float),
                                              _lib.ptr(probs, C.c_float) if probs is not None else None))
    return (actions, prob, probs) if return_probabilities else (actions, prob)


def action_probabilities(policy, state):
    """``PPO.action_probabilities(policy, state)`` — test/quad_game_utilities.jl:65-71 (one state)."""
    if hasattr(policy, "action_probabilities"):
        return policy.action_probabilities(state)
    from .rollout_buffer import StateData
    s = StateData(np.asarray(state.vertex_score)[None], np.asarray(state.action_mask)[None])
    return batch_action_probabilities(policy, s)[0]


def number_of_actions_per_state(state):
    """hook ``number_of_actions_per_state`` (src/ProximalPolicyOptimization.jl:28): size(action_mask, 1)."""
    return int(np.asarray(state.action_mask).shape[-1])


class Adam:
    """``Flux.Optimise.Adam(eta, beta, epsilon)``; state (mt, vt, beta^t) lives on the device and is
    created when first bound to a policy (Flux's IdDict keyed by parameter array)."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
        self.beta, self.epsilon = (float(beta[0]), float(beta[1])), float(epsilon)
        self._eta = float(eta)
        self._h = None
        self._policy = None

    @property
    def eta(self):
        return self._eta

    @eta.setter
    def eta(self, v):
        self._eta = float(v)
        if self._h is not None:
            _lib.check(_lib.load().ppo_adam_set_eta(self._h, self._eta))

    def bind(self, policy: Policy):
        if self._h is None:
            h = C.c_void_p()
            _lib.check(_lib.load().ppo_adam_create(policy.handle, self._eta, self.beta[0], self.beta[1], self.epsilon,
                                                   C.byref(h)))
            self._h, self._policy = h, policy
            policy.ctx.adopt(self)
            policy._optimisers.add(self)
        elif self._policy is not policy:
            raise ValueError("this Adam instance already holds state for another policy")
        return self._h

    def update_(self, policy: Policy, grad_flat):
        """``Flux.update!(optimizer, weights, grad)`` with a host gradient in Flux.params order."""
        h = self.bind(policy)
        g = np.ascontiguousarray(grad_flat, np.float32)
        assert g.size == policy.num_params
        _lib.check(_lib.load().ppo_adam_update(h, _lib.ptr(g, C.c_float)))

    def close(self):
        if self._h is not None:
            _lib.load().ppo_adam_destroy(self._h)
            self._h = None
            self._policy = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Optimiser(list):
    """``Flux.Optimise.Optimiser(opts...)`` — an iterable chain; the reference iterates it in
    ``get_optimizer_learning_rate`` (src/train.jl:155-158).  Only chains whose single stateful
    member is an Adam are supported on the device."""

    def __init__(self, *opts):
        super().__init__(opts)

    def adam(self) -> Adam:
        adams = [o for o in self if isinstance(o, Adam)]
        if len(adams) != 1 or len(self) != 1:
            raise NotImplementedError("device optimiser chain must be exactly Optimiser(Adam(...))")
        return adams[0]


def as_adam(optimizer) -> Adam:
    if isinstance(optimizer, Adam):
        return optimizer
    if isinstance(optimizer, Optimiser):
        return optimizer.adam()
    raise TypeError(f"unsupported optimiser {type(optimizer)}")
