"""Context: one per GPU.  Owns the device-side state behind the C ABI (streams, scratch, resident shard)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = L.vp()
        L.check(self._lib.srgp_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = device
        self.n = 0
        self.d = 0

    def close(self):
        if getattr(self, "handle", None):
            self._lib.srgp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- instrumentation -------------------------------------------------------------------
    def sync(self):
        L.check(self._lib.srgp_ctx_sync(self.handle))

    def timer_start(self):
        L.check(self._lib.srgp_timer_start(self.handle))

    def timer_stop_ms(self) -> float:
        ms = L.cd()
        L.check(self._lib.srgp_timer_stop_ms(self.handle, C.byref(ms)))
        return ms.value

    def prof_enable(self, on=True):
        L.check(self._lib.srgp_prof_enable(self.handle, int(bool(on))))

    def prof_reset(self):
        L.check(self._lib.srgp_prof_reset(self.handle))

    def prof_get(self, name):
        n, ms = L.i64(), L.cd()
        L.check(self._lib.srgp_prof_get(self.handle, L.PROF[name], C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self) -> int:
        return int(self._lib.srgp_launch_count(self.handle))

    def flush_l2(self):
        L.check(self._lib.srgp_flush_l2(self.handle))

    # ---- raw device memory (bench / tests) --------------------------------------------------
    def dev_alloc(self, nbytes: int):
        p = L.vp()
        L.check(self._lib.srgp_dev_alloc(self.handle, int(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        L.check(self._lib.srgp_dev_free(self.handle, p))

    def h2d(self, dev, arr: np.ndarray):
        L.check(self._lib.srgp_memcpy_h2d(self.handle, dev, arr.ctypes.data_as(L.vp), arr.nbytes))

    def d2h(self, arr: np.ndarray, dev):
        L.check(self._lib.srgp_memcpy_d2h(self.handle, arr.ctypes.data_as(L.vp), dev, arr.nbytes))

    def fill_normal(self, dev, n, seed, mean=0.0, sd=1.0):
        L.check(self._lib.srgp_fill_normal_dev(self.handle, dev, int(n), int(seed), float(mean), float(sd)))


_default = None


def default_context() -> Context:
    """Process-wide context on cuda:0 used by the reference-shaped functional API."""
    global _default
    if _default is None:
        _default = Context(0)
    return _default
