"""CPU oracle for the PPO-update hot path of ProximalPolicyOptimization.jl.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference's
algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product path
(``proximalpolicyoptimization.jl_b200``) never does and fails loudly when the CUDA
library is missing.

PARITY STATUS: **partially pinned**.  Julia is not installed in the build image and
the reference's arithmetic below ``src/train.jl`` lives in un-vendored, un-pinned
packages (Flux / Zygote / NNlib / Random; ``Project.toml:6-17`` has no versions, the
Manifest is git-ignored).  What IS pinned by the reference's own artefacts:

* ``compute_returns`` against ``output/trajectory.csv:1-7`` (returns 6,5,4,3,2,1) and
  the ``TestEnv`` run implied by ``test/test_rollout_buffer.jl:4-50``
  (see ``tests/test_oracle_golden.py``).
* the BSON/CSV wire format against ``output/states/sample_1.bson`` and
  ``examples/rollout_to_disk/states/sample_1.bson``.

Everything else (MLP forward, softmax, loss value, gradient, Adam step) is a
restatement of the reference call sites plus the ASSUMED third-party formulas marked
"ASSUMPTION" below, cross-checked against torch fp64 autograd as an independent second
opinion (``tests/test_oracle_autograd.py``): "parity unpinned" for those rows.

Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).  Julia is column-major and 1-based; here arrays are C-ordered
with the Julia trailing (batch) dimension first, and indices are 0-based unless a
function says otherwise:

    Julia vertex_score [nf, nhe, nb]   <->  feat  [nb, nhe, nf]
    Julia action_mask  [A, nb]         <->  mask  [nb, A]
    Julia Dense weight W [out, in]     <->  W     [in, out]  (same bytes)
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64

# --------------------------------------------------------------------------------------
# returns scan
# --------------------------------------------------------------------------------------


def compute_returns(rewards, terminal, discount, discount_is_f32=False):
    """``compute_returns`` — src/collect_rollouts.jl:26-42.

    Serial reverse loop ``v = rewards[idx] + discount * v`` with ``v = 0`` at every
    ``terminal[idx]`` (terminal marks the LAST transition of an episode).  Julia
    promotion rule: ``v`` starts as ``zero(Float32)`` but becomes Float64 after the first
    iteration when ``discount`` is a Float64 (1.0 everywhere in-tree) and stays Float64;
    the value is rounded to Float32 only when stored into ``values``.  With a Float32
    discount (``discount_is_f32``) the carry stays Float32.
    """
    r = np.asarray(rewards, dtype=F32)
    t = np.asarray(terminal).astype(bool)
    n = r.shape[0]
    out = np.zeros(n, dtype=F32)
    if discount_is_f32:
        g = F32(discount)
        v = F32(0)
        for i in range(n - 1, -1, -1):
            if t[i]:
                v = F32(0)
            v = F32(r[i] + F32(g * v))
            out[i] = v
    else:
        g = float(discount)
        v = 0.0
        rl = r.astype(F64)
        for i in range(n - 1, -1, -1):
            if t[i]:
                v = 0.0
            v = rl[i] + g * v
            out[i] = v  # rounds to Float32 on store
    return out


def compute_returns_fast(rewards, terminal, discount, discount_is_f32=False):
    """Same serial recurrence as :func:`compute_returns`, run by the plain-C restatement
    (oracle/ppo_oracle_c.c:oracle_compute_returns) so large N finishes in milliseconds."""
    from . import c_oracle
    return c_oracle.compute_returns(rewards, terminal, discount, discount_is_f32)


# --------------------------------------------------------------------------------------
# rollout buffer / dataset
# --------------------------------------------------------------------------------------


class BufferRollouts:
    """``BufferRollouts`` — src/rollout_buffer.jl:1-22 (struct + ctor), ``update!`` :24-38,
    ``length`` :40-48, ``compute_state_value!`` :55-64, ``permute!`` / ``shuffle!`` :81-93.

    ``state_data`` is held as two arrays (vertex_score, action_mask of the quad-game
    ``StateData`` — test/quad_game_utilities.jl:17-20); actions are kept 1-based Int64 as
    in the reference.
    """

    def __init__(self, nf, nhe, apa):
        self.nf, self.nhe, self.apa = nf, nhe, apa
        self.A = nhe * apa
        self.feat = np.zeros((0, nhe, nf), dtype=F32)
        self.mask = np.zeros((0, self.A), dtype=F32)
        self.selected_action_probabilities = np.zeros(0, dtype=F32)
        self.selected_actions = np.zeros(0, dtype=np.int64)
        self.rewards = np.zeros(0, dtype=F32)
        self.terminal = np.zeros(0, dtype=bool)

    def update(self, feat, mask, action_probability, action, reward, terminal):
        """Batched ``update!`` (5 ``push!`` per transition) — src/rollout_buffer.jl:24-38."""
        feat = np.asarray(feat, dtype=F32).reshape(-1, self.nhe, self.nf)
        n = feat.shape[0]
        self.feat = np.concatenate([self.feat, feat])
        self.mask = np.concatenate([self.mask, np.asarray(mask, dtype=F32).reshape(n, self.A)])
        self.selected_action_probabilities = np.concatenate(
            [self.selected_action_probabilities, np.asarray(action_probability, dtype=F32).reshape(n)])
        self.selected_actions = np.concatenate(
            [self.selected_actions, np.asarray(action, dtype=np.int64).reshape(n)])
        self.rewards = np.concatenate([self.rewards, np.asarray(reward, dtype=F32).reshape(n)])
        self.terminal = np.concatenate([self.terminal, np.asarray(terminal).astype(bool).reshape(n)])

    def __len__(self):
        n = self.selected_action_probabilities.shape[0]
        assert n == self.selected_actions.shape[0] == self.rewards.shape[0] == self.feat.shape[0] \
            == self.terminal.shape[0]
        return n

    def compute_state_value(self, discount, discount_is_f32=False):
        """``rollouts.rewards .= compute_returns(...)`` in place — src/rollout_buffer.jl:55-64."""
        self.rewards[:] = compute_returns(self.rewards, self.terminal, discount, discount_is_f32)

    def permute(self, idx1):
        """``permute!(rollouts, idx)`` with a 1-based index vector — src/rollout_buffer.jl:81-88."""
        idx = np.asarray(idx1, dtype=np.int64) - 1
        assert idx.shape[0] == len(self)
        self.feat = self.feat[idx]
        self.mask = self.mask[idx]
        self.selected_action_probabilities = self.selected_action_probabilities[idx]
        self.selected_actions = self.selected_actions[idx]
        self.rewards = self.rewards[idx]
        self.terminal = self.terminal[idx]


def get_batch(buf: BufferRollouts, indices1):
    """``get_batch`` / ``getindex`` — src/rollout_buffer.jl:117-143 with ``batch_state`` of
    test/quad_game_utilities.jl:26-33 (``cat(vs..., dims=3)``, ``cat(am..., dims=2)``).

    ``indices1`` are 1-based as in Julia.  Returns the reference's 4-key Dict; ``state`` is
    the tuple (feat [nb, nhe, nf], mask [nb, A]).
    """
    idx = np.asarray(indices1, dtype=np.int64) - 1
    assert idx.ndim == 1
    assert idx.size == 0 or (idx.min() >= 0 and idx.max() < len(buf))
    return {
        "state": (buf.feat[idx], buf.mask[idx]),
        "selected_action": buf.selected_actions[idx],
        "selected_action_probability": buf.selected_action_probabilities[idx],
        "returns": buf.rewards[idx],
    }


def get_linear_action_index(selected_actions1, num_actions_per_state):
    """``get_linear_action_index`` — src/train.jl:48-52: ``a_b + (b-1)*A`` (1-based,
    column-major linear index into probs[A, nb])."""
    a = np.asarray(selected_actions1, dtype=np.int64)
    return a + np.arange(a.shape[0], dtype=np.int64) * int(num_actions_per_state)


# --------------------------------------------------------------------------------------
# device-shuffle contract (no reference counterpart for the *sequence*: Julia's randperm
# is stdlib code outside the tree; the reference only requires "a uniformly random
# permutation", src/train.jl:93).  The device generates perm[i] = walk(i) with this exact
# bijection; the CUDA kernel must match it bit for bit.
# --------------------------------------------------------------------------------------

FEISTEL_ROUNDS = 10
_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return x, (z ^ (z >> 31)) & _M64


def feistel_keys(seed):
    """Round keys: the low 32 bits of FEISTEL_ROUNDS successive splitmix64 outputs."""
    s = int(seed) & _M64
    keys = []
    for _ in range(FEISTEL_ROUNDS):
        s, z = _splitmix64(s)
        keys.append(z & 0xFFFFFFFF)
    return keys


def _fmix32(h):
    # murmur3 finaliser, vectorised over uint32 (wraps mod 2^32)
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h


def feistel_bits(n):
    bits = max(2, int(n - 1).bit_length()) if n > 1 else 2
    abits = bits // 2          # high half
    bbits = bits - abits       # low half
    return abits, bbits


def feistel_permutation(n, seed):
    """0-based permutation of range(n): cycle-walking generalised Feistel network over
    2^(abits+bbits) >= n (K3 contract).  Round r (even) updates the high half from the low
    half, round r (odd) the low half from the high half:
        a ^= fmix32(b ^ key[r]) & amask      /     b ^= fmix32(a ^ key[r]) & bmask
    """
    n = int(n)
    abits, bbits = feistel_bits(n)
    amask = np.uint64((1 << abits) - 1)
    bmask = np.uint64((1 << bbits) - 1)
    keys = feistel_keys(seed)
    x = np.arange(n, dtype=np.uint64)
    out = np.empty(n, dtype=np.int64)
    todo = np.arange(n, dtype=np.int64)
    cur = x.copy()
    while todo.size:
        a = cur >> np.uint64(bbits)
        b = cur & bmask
        for r in range(FEISTEL_ROUNDS):
            k = np.uint64(keys[r])
            if r % 2 == 0:
                a ^= _fmix32(b ^ k) & amask
            else:
                b ^= _fmix32(a ^ k) & bmask
        y = (a << np.uint64(bbits)) | b
        ok = y < np.uint64(n)
        out[todo[ok]] = y[ok].astype(np.int64)
        todo = todo[~ok]
        cur = y[~ok]
    return out


# --------------------------------------------------------------------------------------
# policy (test/policy.jl:5-31) and batch_action_probabilities
# --------------------------------------------------------------------------------------

LEAKY_SLOPE = 0.01  # ASSUMPTION: NNlib.leakyrelu default a = 0.01 (not in tree)


class Policy:
    """``SimplePolicy.Policy(in, hidden, num_hidden_layers, num_output)`` — test/policy.jl:9-21:
    ``Chain(Dense(in,h,leakyrelu), (L-1) x Dense(h,h,leakyrelu), Dense(h,out))``.

    ``W[l]`` has shape [in_l, out_l] (C order) = the bytes of Julia's ``Dense.weight``
    [out, in] (column-major).  Initialisation: Flux ``Dense`` default glorot_uniform,
    zero bias (ASSUMPTION on the Flux default; any weights can be passed in).
    """

    def __init__(self, in_channels, hidden_channels, num_hidden_layers, num_output, rng=None, dtype=F32):
        dims = [in_channels] + [hidden_channels] * num_hidden_layers + [num_output]
        self.dims = dims
        self.slope = LEAKY_SLOPE
        self.W, self.b = [], []
        rng = rng if rng is not None else np.random.default_rng(0)
        for i, o in zip(dims[:-1], dims[1:]):
            lim = np.sqrt(6.0 / (i + o))
            self.W.append(rng.uniform(-lim, lim, size=(i, o)).astype(dtype))
            self.b.append(np.zeros(o, dtype=dtype))

    @property
    def num_params(self):
        return sum(w.size + b.size for w, b in zip(self.W, self.b))

    def copy(self):
        p = Policy.__new__(Policy)
        p.dims = list(self.dims)
        p.slope = getattr(self, "slope", LEAKY_SLOPE)
        p.W = [w.copy() for w in self.W]
        p.b = [b.copy() for b in self.b]
        return p

    def flat(self):
        """Parameters in Flux.params order: W1, b1, W2, b2, ... (bytes as Julia stores them)."""
        return np.concatenate([np.concatenate([w.ravel(), b.ravel()]) for w, b in zip(self.W, self.b)])


def leakyrelu(x, a=LEAKY_SLOPE):
    """ASSUMPTION (NNlib >= 0.7): ``leakyrelu(x, a) = ifelse(x > 0, x, a*x)``."""
    return np.where(x > 0, x, x.dtype.type(a) * x)


def mlp_forward(policy: Policy, x, keep=False):
    """``p.model(state)`` — test/policy.jl:29-31.  Dense acts on dim 1 of [nf, nhe, nb], i.e. on
    the last axis of ``x`` [.., nf] here.  Returns logits [.., apa] (and the activations)."""
    h = x
    acts = [h]
    L = len(policy.W)
    for l, (W, b) in enumerate(zip(policy.W, policy.b)):
        z = h @ W + b
        h = leakyrelu(z, getattr(policy, "slope", LEAKY_SLOPE)) if l < L - 1 else z
        acts.append(h)
    return (h, acts) if keep else h


def softmax_cols(z):
    """ASSUMPTION (NNlib.softmax, dims=1): ``exp.(x .- maximum(x))`` then ``./ sum``; in the
    working dtype.  ``z`` is [nb, A]; softmax over the last axis (= Julia dim 1)."""
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=-1, keepdims=True)


def batch_action_probabilities(policy: Policy, feat, mask):
    """``PPO.batch_action_probabilities`` — test/quad_game_utilities.jl:73-79:
    ``logits = reshape(policy(vertex_score), :, nb) + action_mask; softmax(logits, dims=1)``.
    feat [nb, nhe, nf], mask [nb, A] -> probs [nb, A] (A = nhe*apa, action index
    a = apa_idx + apa*he, the column-major order of [apa, nhe])."""
    nb = feat.shape[0]
    logits = mlp_forward(policy, feat).reshape(nb, -1) + mask
    return softmax_cols(logits)


# --------------------------------------------------------------------------------------
# loss (src/train.jl:1-46)
# --------------------------------------------------------------------------------------


def simplified_ppo_clip(advantage, epsilon):
    """``simplified_ppo_clip`` — src/train.jl:1-7.  ``(1f0 + eps) * adv`` promotes to Float64
    when eps is a Float64 (it is, at every in-tree call)."""
    adv = np.asarray(advantage)
    e = F64(epsilon)
    adv64 = adv.astype(F64)
    return np.where(adv >= 0, (1.0 + e) * adv64, (1.0 - e) * adv64)


def smoothed_entropy(probs, smooth=F32(1e-8)):
    """``smoothed_entropy`` — src/train.jl:21-26.  probs [nb, A].
    ``(1f0 - 1f-8) == 1f0`` in Float32, so p~ = fl32(p + fl32(1f-8 / A));  returns
    ``mean_b( -sum_a p~ log p~ )`` in the dtype of ``probs``."""
    dt = probs.dtype.type
    A = probs.shape[-1]
    sp = dt(dt(1.0) - dt(smooth)) * probs + dt(dt(smooth) / dt(A))
    h = sp * np.log(sp)
    h = -h.sum(axis=-1)
    return h.mean(dtype=probs.dtype)


def ppo_loss_terms(probs, actions1, old_probs, advantage, epsilon):
    """Body of ``ppo_loss_with_entropy`` — src/train.jl:35-46 — on given probabilities.
    actions1: 1-based action per sample (the linear index of :48-52 modulo the column).
    Returns (ppoloss Float64, entropyloss in probs dtype, sel, gain, clip)."""
    nb = probs.shape[0]
    sel = probs[np.arange(nb), np.asarray(actions1, dtype=np.int64) - 1]
    gain = sel / old_probs * advantage                    # fp32: (sel/old)*adv
    clip = simplified_ppo_clip(advantage, epsilon)        # Float64
    ppoloss = -np.mean(np.minimum(gain.astype(F64), clip))
    entropyloss = -smoothed_entropy(probs)
    return ppoloss, entropyloss, sel, gain, clip


def ppo_loss_with_entropy(policy, feat, mask, actions1, old_probs, advantage, epsilon):
    """``ppo_loss_with_entropy`` — src/train.jl:35-46."""
    probs = batch_action_probabilities(policy, feat, mask)
    ppoloss, entropyloss, *_ = ppo_loss_terms(probs, actions1, old_probs, advantage, epsilon)
    return ppoloss, entropyloss


def loss_grad_logits(logits, mask, actions1, old_probs, advantage, epsilon, entropy_weight, smooth=F32(1e-8),
                     nb_total=None):
    """Analytic d(ppoloss + entropy_weight*entropyloss)/dlogits — what Zygote differentiates at
    src/train.jl:67-79 restricted to the part after the MLP.  Returns
    (ppoloss, entropyloss_unweighted, probs, dlogits[nb, A]).

        g[b,a]  = w/nb * (1-s) * (log p~[b,a] + 1)  -  [a == a_b][gain_b <= clip_b] adv_b / (nb old_b)
        dz[b,a] = p[b,a] * (g[b,a] - sum_a' p[b,a'] g[b,a'])         (masked => exactly 0)

    ASSUMPTION: on a tie gain == clip the derivative follows ``gain`` (Base.min returns its
    first argument on ties; measure-zero in practice).

    ``nb_total``: the rows are a chunk of a minibatch of nb_total samples (test helper for minibatches too large to
    hold in Float64 at once): the 1/nb factors of the gradient use nb_total, the returned losses are the chunk's means.
    """
    dt = logits.dtype.type
    nb, A = logits.shape
    nbt = nb if nb_total is None else int(nb_total)
    z = logits + mask
    p = softmax_cols(z)
    ppoloss, entropyloss, sel, gain, clip = ppo_loss_terms(p, actions1, old_probs, advantage, epsilon)
    sp = dt(dt(1.0) - dt(smooth)) * p + dt(dt(smooth) / dt(A))
    w = dt(entropy_weight)
    g = (w / dt(nbt)) * dt(dt(1.0) - dt(smooth)) * (np.log(sp) + dt(1.0))
    active = gain.astype(F64) <= clip
    coef = np.where(active, advantage / old_probs / dt(nbt), dt(0)).astype(logits.dtype)
    g[np.arange(nb), np.asarray(actions1, dtype=np.int64) - 1] -= coef
    dz = p * (g - (p * g).sum(axis=-1, keepdims=True))
    return ppoloss, entropyloss, p, dz


def policy_gradient(policy: Policy, feat, mask, actions1, old_probs, advantage, epsilon, entropy_weight,
                    gates=None, nb_total=None, return_acts=False):
    """Gradient of ``ppoloss + entropyloss*entropy_weight`` w.r.t. ``Flux.params(policy)`` —
    src/train.jl:65-79 — by the Dense pullbacks (dW = x' * delta, db = sum delta, dx = delta * W').
    ASSUMPTION: leakyrelu'(z) = 1 for z > 0 else a (also at z == 0).
    Returns (ppoloss, entropyloss*entropy_weight, dW list, db list[, acts]).

    ``gates`` (test instrumentation, no reference counterpart): ``gates[l]`` (l = 1..L-1, boolean [tokens, dims[l]])
    replaces this evaluation's own ``acts[l] > 0`` as the leakyrelu' branch of hidden activation l.  leakyrelu' is
    discontinuous at 0, so two evaluations that round differently can take different branches for a pre-activation
    within rounding of zero; handing the device's branches (ppo_policy_read_gates) to the oracle makes the gradient
    comparison well posed, and the tests check separately that every disagreement sits at such a pre-activation.
    A gate value of 255 (PPO_GATE_SKIPPED: a token the device's compacted MLP never ran because all of its actions are
    masked) keeps this evaluation's own branch; the tests assert that the incoming gradient of such a row is exactly 0.
    ``nb_total``: see :func:`loss_grad_logits`."""
    nb = feat.shape[0]
    dt = policy.W[0].dtype
    x = feat.reshape(-1, feat.shape[-1]).astype(dt)        # tokens [nb*nhe, nf]
    logits_tok, acts = mlp_forward(policy, x, keep=True)
    logits = logits_tok.reshape(nb, -1)
    ppoloss, entropyloss, _, dz = loss_grad_logits(
        logits, mask.astype(dt), actions1, old_probs.astype(dt), advantage.astype(dt), epsilon, entropy_weight,
        nb_total=nb_total)
    delta = dz.reshape(logits_tok.shape).astype(dt)
    L = len(policy.W)
    dW, db = [None] * L, [None] * L
    for l in range(L - 1, -1, -1):
        dW[l] = acts[l].T @ delta
        db[l] = delta.sum(axis=0)
        if l > 0:
            dx = delta @ policy.W[l].T
            pos = acts[l] > 0
            if gates is not None:
                gl = np.asarray(gates[l])
                pos = np.where(gl == 255, pos, gl.astype(bool)) if gl.dtype == np.uint8 else gl.astype(bool)
            delta = np.where(pos, dx, dt.type(getattr(policy, "slope", LEAKY_SLOPE)) * dx)
    out = (ppoloss, F64(entropyloss) * F64(entropy_weight), dW, db)
    return out + (acts,) if return_acts else out


# --------------------------------------------------------------------------------------
# optimiser: Flux.Optimise.Adam (call site src/train.jl:81)
# --------------------------------------------------------------------------------------


class Adam:
    """ASSUMPTION (Flux 0.13/0.14 ``Optimise.Adam``; source not in tree):
        mt = b1*mt + (1-b1)*g ; vt = b2*vt + (1-b2)*g^2
        delta = mt / (1 - b1^t) / (sqrt(vt / (1 - b2^t)) + eps) * eta ;  x -= delta
    eta, betas, eps are Float64 scalars broadcast against Float32 arrays: every
    element-wise expression is evaluated in Float64 and rounded to Float32 when stored
    into mt / vt / delta; ``x .-= delta`` is Float32."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), epsilon=1e-8):
        self.eta, self.beta, self.epsilon = float(eta), (float(beta[0]), float(beta[1])), float(epsilon)
        self.state = {}

    def apply(self, key, x, g):
        if key not in self.state:
            self.state[key] = [np.zeros_like(x), np.zeros_like(x), [self.beta[0], self.beta[1]]]
        mt, vt, bp = self.state[key]
        b1, b2 = self.beta
        g64 = g.astype(F64)
        mt[...] = (b1 * mt.astype(F64) + (1 - b1) * g64).astype(x.dtype)
        vt[...] = (b2 * vt.astype(F64) + (1 - b2) * g64 * g64).astype(x.dtype)
        delta = (mt.astype(F64) / (1 - bp[0]) / (np.sqrt(vt.astype(F64) / (1 - bp[1])) + self.epsilon)
                 * self.eta).astype(x.dtype)
        bp[0] *= b1
        bp[1] *= b2
        x -= delta

    def update(self, policy: Policy, dW, db):
        """``Flux.update!(optimizer, weights, grad)`` — src/train.jl:81."""
        for l in range(len(policy.W)):
            self.apply(("W", l), policy.W[l], dW[l])
            self.apply(("b", l), policy.b[l], db[l])


def get_optimizer_learning_rate(optimizer):
    """``get_optimizer_learning_rate`` — src/train.jl:155-158 (product of etas of the chain)."""
    opts = optimizer if isinstance(optimizer, (list, tuple)) else [optimizer]
    lr = 1.0
    for o in opts:
        lr *= o.eta
    return lr


# --------------------------------------------------------------------------------------
# step_batch! / step_epoch! / ppo_train!
# --------------------------------------------------------------------------------------


def step_batch(policy, optimizer, feat, mask, actions1, old_probs, advantage, epsilon, entropy_weight):
    """``step_batch!`` — src/train.jl:54-84.  Returns (ppoloss, entropyloss*entropy_weight)."""
    ppoloss, entw, dW, db = policy_gradient(policy, feat, mask, actions1, old_probs, advantage,
                                            epsilon, entropy_weight)
    optimizer.update(policy, dW, db)
    return ppoloss, entw


def step_epoch(policy, optimizer, buf: BufferRollouts, epsilon, batch_size, entropy_weight, perm1,
               batch_advantage=None):
    """``step_epoch!`` — src/train.jl:86-128.  ``perm1`` is the 1-based permutation the
    reference draws with ``randperm(num_data)`` (:93); it is an input here (Julia's RNG stream
    is not reproducible outside Julia).  ``batch_advantage`` defaults to identity on the
    returns (no in-tree implementation of the hook, src/ProximalPolicyOptimization.jl:29).
    Returns the unweighted means of the per-minibatch losses (:127)."""
    num_data = len(buf)
    assert 1 <= batch_size <= num_data
    perm1 = np.asarray(perm1, dtype=np.int64)
    ppo_hist, ent_hist = [], []
    start = 0
    while start < num_data:
        stop = min(start + batch_size, num_data)
        batch = get_batch(buf, perm1[start:stop])
        feat, mask = batch["state"]
        returns = batch["returns"]
        adv = returns if batch_advantage is None else batch_advantage((feat, mask), returns)
        p, e = step_batch(policy, optimizer, feat, mask, batch["selected_action"],
                          batch["selected_action_probability"], adv, epsilon, entropy_weight)
        ppo_hist.append(p)
        ent_hist.append(e)
        start = stop
    return float(np.mean(ppo_hist)), float(np.mean(ent_hist))


def ppo_train(policy, optimizer, buf, epsilon, batch_size, num_epochs, entropy_weight, perms1,
              batch_advantage=None, printer=None):
    """``ppo_train!`` — src/train.jl:130-153.  ``perms1[e]`` is epoch e's permutation."""
    ph, eh, lh = [], [], []
    for epoch in range(1, num_epochs + 1):
        p, e = step_epoch(policy, optimizer, buf, epsilon, batch_size, entropy_weight, perms1[epoch - 1],
                          batch_advantage)
        lr = get_optimizer_learning_rate(optimizer)
        if printer is not None:
            printer(format_epoch_line(epoch, p, e, lr))
        ph.append(p), eh.append(e), lh.append(lr)
    return ph, eh, lh


def format_epoch_line(epoch, ppoloss, entropyloss, lr):
    """The ``@printf`` line of src/train.jl:146, byte for byte."""
    m, ex = ("%1.1e" % lr).split("e")
    # Julia's %e prints at least two exponent digits like C; identical to Python here.
    return "EPOCH : %d \t PPO LOSS : %1.4f\t ENTROPY LOSS : %1.4f \t LR : %se%s\n" % (
        epoch, ppoloss, entropyloss, m, ex)


def normalize_advantage(returns, eps=1e-8):
    """EXTENSION (no reference counterpart; default OFF in the product): (x - mean)/(std + eps)
    with the population std, statistics in Float64, result Float32."""
    r = np.asarray(returns, dtype=F64)
    mu = r.mean()
    sd = np.sqrt(np.maximum((r * r).mean() - mu * mu, 0.0))
    return ((r - mu) / (sd + eps)).astype(F32)


# ---------------------------------------------------------------------------------------------
# batched rollout inference (extension; the reference samples one state at a time on the host)
# ---------------------------------------------------------------------------------------------
def sample_uniforms(seed, n):
    """Float32 uniforms of ppo_sample_actions: draw i = top 24 bits of output i of the splitmix64 stream of `seed`."""
    s = int(seed) & _M64
    out = np.empty(n, np.float32)
    for i in range(n):
        s, z = _splitmix64(s)
        out[i] = np.float32(z >> 40) * np.float32(1.0 / 16777216.0)
    return out


def sample_actions_from_probs(probs, seed):
    """src/collect_rollouts.jl:5-7 ``rand(Categorical(ap))`` for every row of probs [nb, A] (Float32).

    ASSUMPTION (Distributions.jl is un-vendored and un-pinned): inverse CDF with a sequential cumulative sum in the
    element type of the probabilities, ``while cp <= draw && i < n``.  The RNG stream cannot be Julia's; the draws
    are sample_uniforms(seed, nb).  Extension: if rounding leaves the total below the draw, the last action with
    non-zero probability is returned (never a masked one).  Returns 1-based actions and their probabilities."""
    probs = np.asarray(probs, np.float32)
    nb, A = probs.shape
    u = sample_uniforms(seed, nb)
    act = np.empty(nb, np.int64)
    for b in range(nb):
        p = probs[b]
        c = np.float32(p[0])
        a = 0
        last = 0 if p[0] > 0 else -1
        while c <= u[b] and a < A - 1:
            a += 1
            c = np.float32(c + p[a])
            if p[a] > 0:
                last = a
        if not p[a] > 0 and last >= 0:
            a = last
        act[b] = a + 1
    return act, probs[np.arange(nb), act - 1]
