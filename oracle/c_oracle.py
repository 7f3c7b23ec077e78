"""ctypes front for oracle/ppo_oracle_c.c (TEST INFRASTRUCTURE ONLY — see ppo_oracle.py header).

``build()`` compiles the C restatement with gcc into ``oracle/_build/libppo_oracle.so``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ppo_oracle_c.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libppo_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(_OUT_DIR, exist_ok=True)
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_OUT):
            build()
        _lib = ctypes.CDLL(_OUT)
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def compute_returns(rewards, terminal, discount, discount_is_f32=False):
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    t = np.ascontiguousarray(np.asarray(terminal).astype(np.uint8))
    out = np.empty_like(r)
    lib().oracle_compute_returns(_p(r, ctypes.c_float), _p(t, ctypes.c_uint8), ctypes.c_int64(r.size),
                                 ctypes.c_double(discount), ctypes.c_int(int(discount_is_f32)),
                                 _p(out, ctypes.c_float))
    return out


def get_batch(feat, mask, actions, probs, returns, indices1):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    actions = np.ascontiguousarray(actions, dtype=np.int64)
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    returns = np.ascontiguousarray(returns, dtype=np.float32)
    idx = np.ascontiguousarray(indices1, dtype=np.int64)
    nb = idx.size
    fe = int(np.prod(feat.shape[1:]))
    me = int(np.prod(mask.shape[1:]))
    fo = np.empty((nb,) + feat.shape[1:], dtype=np.float32)
    mo = np.empty((nb,) + mask.shape[1:], dtype=np.float32)
    ao = np.empty(nb, dtype=np.int64)
    po = np.empty(nb, dtype=np.float32)
    ro = np.empty(nb, dtype=np.float32)
    lib().oracle_get_batch(_p(feat, ctypes.c_float), _p(mask, ctypes.c_float), _p(actions, ctypes.c_int64),
                           _p(probs, ctypes.c_float), _p(returns, ctypes.c_float),
                           ctypes.c_int64(fe), ctypes.c_int64(me), _p(idx, ctypes.c_int64), ctypes.c_int64(nb),
                           _p(fo, ctypes.c_float), _p(mo, ctypes.c_float), _p(ao, ctypes.c_int64),
                           _p(po, ctypes.c_float), _p(ro, ctypes.c_float))
    return {"state": (fo, mo), "selected_action": ao, "selected_action_probability": po, "returns": ro}


def feistel_permutation(n, seed):
    out = np.empty(int(n), dtype=np.int64)
    lib().oracle_feistel_permutation(ctypes.c_int64(int(n)), ctypes.c_uint64(int(seed) & (2**64 - 1)),
                                     _p(out, ctypes.c_int64))
    return out


def adam(x, mt, vt, g, eta, b1, b2, eps, b1p, b2p):
    for a in (x, mt, vt, g):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oracle_adam(_p(x, ctypes.c_float), _p(mt, ctypes.c_float), _p(vt, ctypes.c_float),
                      _p(g, ctypes.c_float), ctypes.c_int64(x.size), ctypes.c_double(eta), ctypes.c_double(b1),
                      ctypes.c_double(b2), ctypes.c_double(eps), ctypes.c_double(b1p), ctypes.c_double(b2p))


if __name__ == "__main__":
    print(build(force=True))
