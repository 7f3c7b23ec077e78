"""ctypes binding of libppo_b200.so (include/ppo_b200.h).

There is NO fallback: if the CUDA library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``csrc/build.sh``) importing the package works but the
first call raises; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPO_B200_LIB: an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("PPO_B200_LIB") or os.path.join(_HERE, "libppo_b200.so")

c_i64, c_u64, c_int, c_dbl, c_flt = C.c_int64, C.c_uint64, C.c_int, C.c_double, C.c_float
vp = C.c_void_p
PF = C.POINTER(C.c_float)
PD = C.POINTER(C.c_double)
PI64 = C.POINTER(C.c_int64)
PU8 = C.POINTER(C.c_uint8)
PPF = C.POINTER(PF)

# name -> (restype, argtypes); mirrors include/ppo_b200.h one to one
SIGNATURES = {
    "ppo_last_error": (C.c_char_p, []),
    "ppo_version": (C.c_char_p, []),
    "ppo_ctx_create": (c_int, [c_int, C.POINTER(vp)]),
    "ppo_ctx_destroy": (c_int, [vp]),
    "ppo_sync": (c_int, [vp]),
    "ppo_ctx_launch_count": (c_i64, [vp]),
    "ppo_ctx_stream": (vp, [vp]),
    "ppo_comm_unique_id": (c_int, [vp]),
    "ppo_comm_init": (c_int, [vp, c_int, c_int, vp]),
    "ppo_comm_destroy": (c_int, [vp]),
    "ppo_comm_allreduce_f64": (c_int, [vp, PD, c_int]),
    "ppo_buffer_create": (c_int, [vp, c_i64, c_int, c_int, c_int, C.POINTER(vp)]),
    "ppo_buffer_destroy": (c_int, [vp]),
    "ppo_buffer_append": (c_int, [vp, c_i64, PF, PF, PI64, PF, PF, PU8]),
    "ppo_buffer_append_i64": (c_int, [vp, c_i64, PI64, PF, PI64, PF, PF, PU8]),
    "ppo_buffer_append_i8": (c_int, [vp, c_i64, C.POINTER(C.c_int8), PF, PI64, PF, PF, PU8]),
    "ppo_buffer_append_i16": (c_int, [vp, c_i64, C.POINTER(C.c_int16), PF, PI64, PF, PF, PU8]),
    "ppo_buffer_append_packed": (c_int, [vp, c_i64, vp, c_int, C.POINTER(C.c_uint64), PI64, PF, PF, PU8]),
    "ppo_buffer_length": (c_i64, [vp]),
    "ppo_buffer_clear": (c_int, [vp]),
    "ppo_compute_returns": (c_int, [vp, c_dbl, c_int]),
    "ppo_normalize_advantage": (c_int, [vp, c_int, c_dbl]),
    "ppo_buffer_save_rewards": (c_int, [vp]),
    "ppo_buffer_restore_rewards": (c_int, [vp]),
    "ppo_buffer_read": (c_int, [vp, c_i64, c_i64, PF, PF, PI64, PF, PF, PU8]),
    "ppo_buffer_permute": (c_int, [vp, PI64, c_i64]),
    "ppo_buffer_shuffle": (c_int, [vp, c_u64]),
    "ppo_permutation_set": (c_int, [vp, PI64, c_i64]),
    "ppo_permutation_generate": (c_int, [vp, c_u64, PI64]),
    "ppo_gather": (c_int, [vp, c_i64, c_i64, PF, PF, PI64, PF, PF]),
    "ppo_gather_indices": (c_int, [vp, PI64, c_i64, PF, PF, PI64, PF, PF]),
    "ppo_gather_device": (c_int, [vp, c_i64, c_i64, c_int]),
    "ppo_batch_read": (c_int, [vp, c_i64, PF, PF, PI64, PF, PF]),
    "ppo_disk_dataset_load": (c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, c_int, PI64, C.POINTER(c_int)]),
    "ppo_bson_state_arrays": (c_int, [C.c_char_p, c_int, C.c_char_p, PI64, C.POINTER(c_int), PI64, C.POINTER(c_int)]),
    "ppo_policy_create": (c_int, [vp, c_int, C.POINTER(c_int), PPF, PPF, c_flt, C.POINTER(vp)]),
    "ppo_policy_destroy": (c_int, [vp]),
    "ppo_policy_read": (c_int, [vp, PPF, PPF]),
    "ppo_policy_write": (c_int, [vp, PPF, PPF]),
    "ppo_policy_set_gemm_mode": (c_int, [vp, c_int]),
    "ppo_policy_get_gemm_mode": (c_int, [vp]),
    "ppo_policy_read_gates": (c_int, [vp, c_int, c_i64, PU8]),
    "ppo_policy_set_token_compaction": (c_int, [vp, c_int]),
    "ppo_policy_active_tokens": (c_int, [vp, PI64]),
    "ppo_policy_p2p_export": (c_int, [vp, vp]),
    "ppo_policy_p2p_connect": (c_int, [vp, c_int, c_int, vp]),
    "ppo_policy_p2p_wait": (c_int, [vp, PI64, PI64, c_int]),
    "ppo_policy_num_params": (c_i64, [vp]),
    "ppo_batch_action_probabilities": (c_int, [vp, c_i64, c_int, PF, PF, PF]),
    "ppo_sample_actions": (c_int, [vp, c_i64, c_int, PF, PF, c_u64, PI64, PF, PF]),
    "ppo_adam_create": (c_int, [vp, c_dbl, c_dbl, c_dbl, c_dbl, C.POINTER(vp)]),
    "ppo_adam_destroy": (c_int, [vp]),
    "ppo_adam_set_eta": (c_int, [vp, c_dbl]),
    "ppo_adam_get_eta": (c_dbl, [vp]),
    "ppo_adam_update": (c_int, [vp, PF]),
    "ppo_loss_from_logits": (c_int, [vp, c_i64, c_int, PF, PF, PI64, PF, PF, c_dbl, c_dbl, PD, PD, PF]),
    "ppo_step_batch_host": (c_int, [vp, vp, c_i64, c_int, PF, PF, PI64, PF, PF, c_dbl, c_dbl, PD, PD, PF]),
    "ppo_step_batch": (c_int, [vp, vp, vp, c_i64, c_i64, c_dbl, c_dbl, PD, PD, PF]),
    "ppo_step_epoch": (c_int, [vp, vp, vp, c_dbl, c_i64, c_dbl, PD, PD]),
    "ppo_train": (c_int, [vp, vp, vp, c_dbl, c_i64, c_int, c_dbl, c_u64, PD, PD, PD]),
    "ppo_dense_op": (c_int, [vp, c_int, c_int, c_i64, c_int, c_int, PF, PF, PF, PF, c_flt, PF, PF]),
    "ppo_bench_kernel": (c_int, [vp, C.c_char_p, c_i64, c_int, c_int, c_int, c_int, c_int, PD, PD]),
}

_lib = None


class PPOError(RuntimeError):
    """Raised for any non-zero ppo_status; ``code`` is the ppo_status value."""

    def __init__(self, code, msg):
        super().__init__(f"ppo_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen libppo_b200.so and attach the prototypes; loud failure when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or proximalpolicyoptimization.jl_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise PPOError(status, load().ppo_last_error().decode("utf-8", "replace"))


def ptr(arr, ctype):
    """Pointer to a C-contiguous numpy array (or None)."""
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(ctype))
