"""PPO update — mirrors reference src/train.jl on the device.

``step_epoch_`` is one C call that enqueues, per minibatch, gather (K4) -> MLP forward (K5) ->
fused masked-softmax/ratio/clip/entropy loss + dlogits (K6) -> MLP backward (K7) ->
[NCCL gradient all-reduce] -> Adam (K8), and reads back two doubles per minibatch at the end.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib
from .policy import Policy, as_adam
from .rollout_buffer import DeviceDataset, StateData


def simplified_ppo_clip(advantage, epsilon):
    """``simplified_ppo_clip`` — src/train.jl:1-7 (scalar; Float64 through the Float64 epsilon).
    Host helper kept for API parity; the update computes it inside the fused loss kernel."""
    adv = float(advantage)
    return (1.0 + float(epsilon)) * adv if adv >= 0 else (1.0 - float(epsilon)) * adv


def get_linear_action_index(selected_actions, num_actions_per_state):
    """``get_linear_action_index`` — src/train.jl:48-52 (1-based, column-major into probs[A, nb])."""
    a = np.asarray(selected_actions, dtype=np.int64)
    return a + np.arange(a.size, dtype=np.int64) * int(num_actions_per_state)


def batch_advantage(state, returns):
    """hook ``batch_advantage(state, returns)`` (src/ProximalPolicyOptimization.jl:29; called at
    src/train.jl:105).  The reference ships no implementation; the default is the identity
    (advantage = returns, as the older API did — examples/.../profile.jl:64)."""
    return returns


def ppo_loss_with_entropy_from_logits(ctx, logits, mask, actions, old_action_probabilities, advantage, epsilon,
                                      entropy_weight=0.0, want_grad=False):
    """``ppo_loss_with_entropy`` (src/train.jl:35-46) evaluated by the fused loss kernel on given
    logits [nb, A] (+ mask).  ``actions`` 1-based within each column.  Returns
    (ppoloss, entropyloss[, dlogits])."""
    lg = np.ascontiguousarray(logits, np.float32)
    mk = np.ascontiguousarray(mask, np.float32)
    nb, A = lg.shape
    act = np.ascontiguousarray(actions, np.int64)
    old = np.ascontiguousarray(old_action_probabilities, np.float32)
    adv = np.ascontiguousarray(advantage, np.float32)
    p, e = C.c_double(), C.c_double()
    dl = np.empty_like(lg) if want_grad else None
    _lib.check(_lib.load().ppo_loss_from_logits(ctx.handle, nb, A, _lib.ptr(lg, C.c_float), _lib.ptr(mk, C.c_float),
                                                _lib.ptr(act, C.c_int64), _lib.ptr(old, C.c_float),
                                                _lib.ptr(adv, C.c_float), float(epsilon), float(entropy_weight),
                                                C.byref(p), C.byref(e), _lib.ptr(dl, C.c_float)))
    return (p.value, e.value, dl) if want_grad else (p.value, e.value)


def step_batch_(policy: Policy, optimizer, state: StateData, linear_action_index, old_action_probabilities,
                advantage, epsilon, entropy_weight, return_grads=False):
    """``step_batch!`` — src/train.jl:54-84, on host arrays (uploaded, then the same device path as
    ``step_epoch_``).  ``optimizer=None`` computes loss and gradient without updating.
    Returns (ppoloss, entropyloss*entropy_weight[, grads in Flux.params order])."""
    feat = np.ascontiguousarray(state.vertex_score, np.float32)
    mask = np.ascontiguousarray(state.action_mask, np.float32)
    nb, nhe = feat.shape[0], feat.shape[1]
    lin = np.ascontiguousarray(linear_action_index, np.int64)
    old = np.ascontiguousarray(old_action_probabilities, np.float32)
    adv = np.ascontiguousarray(advantage, np.float32)
    assert lin.size == nb and old.size == nb and adv.size == nb
    oh = as_adam(optimizer).bind(policy) if optimizer is not None else None
    p, e = C.c_double(), C.c_double()
    grads = np.empty(policy.num_params, np.float32) if return_grads else None
    _lib.check(_lib.load().ppo_step_batch_host(policy.handle, oh, nb, nhe, _lib.ptr(feat, C.c_float),
                                               _lib.ptr(mask, C.c_float), _lib.ptr(lin, C.c_int64),
                                               _lib.ptr(old, C.c_float), _lib.ptr(adv, C.c_float), float(epsilon),
                                               float(entropy_weight), C.byref(p), C.byref(e),
                                               _lib.ptr(grads, C.c_float)))
    return (p.value, e.value, grads) if return_grads else (p.value, e.value)


_epoch_counter = [0]


def step_epoch_(policy: Policy, optimizer, dataset: DeviceDataset, epsilon, batch_size, entropy_weight,
                perm=None, seed=None):
    """``step_epoch!`` — src/train.jl:86-128.

    ``file_indices = randperm(num_data)`` (:93) is either supplied (``perm``, 1-based, e.g. drawn
    by the caller's own RNG) or drawn on the device from ``seed`` (a process-wide counter when
    neither is given).  Returns the unweighted means of the per-minibatch (ppoloss,
    entropyloss*entropy_weight), :127."""
    num_data = len(dataset)
    assert 1 <= batch_size <= num_data
    r = dataset.rollouts
    if perm is not None:
        assert len(perm) == num_data
        r.set_permutation(perm)
    else:
        if seed is None:
            _epoch_counter[0] += 1
            seed = 0x5EED0000 + _epoch_counter[0]
        r.generate_permutation(seed)
    oh = as_adam(optimizer).bind(policy)
    p, e = C.c_double(), C.c_double()
    _lib.check(_lib.load().ppo_step_epoch(policy.handle, oh, r.handle, float(epsilon), int(batch_size),
                                          float(entropy_weight), C.byref(p), C.byref(e)))
    return p.value, e.value


def get_optimizer_learning_rate(optimizer):
    """``get_optimizer_learning_rate`` — src/train.jl:155-158: ``prod(opt.eta for opt in optimizer)``."""
    try:
        opts = list(optimizer)
    except TypeError:
        opts = [optimizer]
    lr = 1.0
    for o in opts:
        lr *= o.eta
    return lr


def format_epoch_line(epoch, ppoloss, entropyloss, lr):
    """The ``@printf`` line of src/train.jl:146, byte for byte."""
    return "EPOCH : %d \t PPO LOSS : %1.4f\t ENTROPY LOSS : %1.4f \t LR : %1.1e\n" % (epoch, ppoloss, entropyloss, lr)


def ppo_train_(policy, optimizer, dataset, epsilon, batch_size, num_epochs, entropy_weight, perms=None, seed=None,
               out=sys.stdout):
    """``ppo_train!`` — src/train.jl:130-153.  ``perms[e]`` optionally fixes epoch e's permutation."""
    ppo_loss_history, entropy_loss_history, lr_history = [], [], []
    for epoch in range(1, num_epochs + 1):
        ppoloss, entropyloss = step_epoch_(policy, optimizer, dataset, epsilon, batch_size, entropy_weight,
                                           perm=None if perms is None else perms[epoch - 1],
                                           seed=None if seed is None else seed + epoch)
        lr = get_optimizer_learning_rate(optimizer)
        if out is not None:
            out.write(format_epoch_line(epoch, ppoloss, entropyloss, lr))
        ppo_loss_history.append(ppoloss)
        entropy_loss_history.append(entropyloss)
        lr_history.append(lr)
    return ppo_loss_history, entropy_loss_history, lr_history


def ppo_iterate_(policy, env, optimizer, episodes_per_iteration, minibatch_size, num_ppo_iterations, evaluator,
                 epochs_per_iteration, discount, epsilon, entropy_weight, rollouts_factory=None, out=sys.stdout):
    """``ppo_iterate!`` (buffer variant) — src/train.jl:210-249.  ``rollouts_factory()`` builds the
    device container (the reference's ``BufferRollouts()`` needs no shapes; the device one does)."""
    from .collect_rollouts import collect_rollouts_
    from .rollout_buffer import construct_dataset
    if rollouts_factory is None:
        raise ValueError("rollouts_factory is required: e.g. lambda: DeviceRollouts(nf, nhe, apa, capacity)")
    loss = {"ppo": [], "entropy": [], "lr": []}
    for it in range(1, num_ppo_iterations + 1):
        evaluator(policy, env, optimizer)
        if out is not None:
            out.write(f"\nPPO ITERATION : {it}\n")
        rollouts = rollouts_factory()
        collect_rollouts_(rollouts, env, policy, episodes_per_iteration, discount)
        dataset = construct_dataset(rollouts)
        ppoloss, entropyloss, lr_history = ppo_train_(policy, optimizer, dataset, epsilon, minibatch_size,
                                                      epochs_per_iteration, entropy_weight, out=out)
        loss["ppo"] += ppoloss
        loss["entropy"] += entropyloss
        loss["lr"] += lr_history
        if hasattr(evaluator, "save_loss"):
            evaluator.save_loss(loss)        # hook save_loss(evaluator, loss), :247
        rollouts.close()
    return loss
