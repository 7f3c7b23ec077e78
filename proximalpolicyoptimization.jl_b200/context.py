"""ppo_ctx wrapper: one context per GPU per process."""
from __future__ import annotations

import ctypes as C
import weakref

from . import _lib


class Context:
    def __init__(self, device: int = 0):
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.ppo_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.nranks, self.rank = 1, 0
        self._children = weakref.WeakSet()   # buffers / policies / optimisers living on this ctx

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("context destroyed")
        return self._h

    def sync(self):
        _lib.check(_lib.load().ppo_sync(self.handle))

    def launch_count(self) -> int:
        return int(_lib.load().ppo_ctx_launch_count(self.handle))

    def stream(self) -> int:
        return int(_lib.load().ppo_ctx_stream(self.handle) or 0)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(_lib.load().ppo_comm_init(self.handle, int(nranks), int(rank), buf))
        self.nranks, self.rank = int(nranks), int(rank)

    def bench_kernel(self, which: str, n: int, a: int = 0, b: int = 0, c: int = 0, iters: int = 10, flush_l2=True):
        ms, work = C.c_double(), C.c_double()
        _lib.check(_lib.load().ppo_bench_kernel(self.handle, which.encode(), int(n), int(a), int(b), int(c),
                                                int(iters), int(bool(flush_l2)), C.byref(ms), C.byref(work)))
        return ms.value, work.value

    def adopt(self, child):
        self._children.add(child)

    def close(self):
        for ch in list(self._children):
            ch.close()
        if self._h is not None:
            _lib.load().ppo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_context(device: int = 0) -> Context:
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().ppo_comm_unique_id(buf))
    return buf.raw
