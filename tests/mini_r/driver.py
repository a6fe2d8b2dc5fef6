"""Python driver for the miniature R runtime (tests/mini_r/) -- TEST INFRASTRUCTURE.

build() compiles r/shim.c (the `.Call` glue of INTEGRATION.md) UNCHANGED against tests/mini_r/include and links it
with the product library sparsergps_b200/libsrgp.so into tests/mini_r/_build/sparseRGPs.so -- the shared object name
R's useDynLib(sparseRGPs) would load (NAMESPACE:89). dot_call() is `.Call(name, ...)`: registered routines only,
arity checked, R errors raised as RError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
SHIM = os.path.join(ROOT, "r", "shim.c")
RUNTIME = os.path.join(_HERE, "mini_r.c")
LIBDIR = os.path.join(ROOT, "sparsergps_b200")
OUT = os.path.join(_HERE, "_build", "sparseRGPs.so")

NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 0, 10, 13, 14, 16, 19


class RError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    deps = [SHIM, RUNTIME] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(_HERE, "include")) for f in fs]
    deps.append(os.path.join(ROOT, "include", "srgp.h"))
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-g", "-Wall", "-Werror",
                               "-fPIC", "-shared", "-I", os.path.join(_HERE, "include"), "-I", os.path.join(ROOT, "include"),
                               SHIM, RUNTIME,
                               "-L", LIBDIR, "-lsrgp", "-Wl,-rpath," + LIBDIR, "-lm", "-o", OUT])
    return OUT


_lib = None
_P = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        from sparsergps_b200 import _lib as product
        product.load()                        # fails loudly when libsrgp.so is not built
        _lib = C.CDLL(build())
        for name, res, args in [
            ("mr_dll", _P, []), ("mr_n_routines", C.c_int, []), ("mr_routine_name", C.c_char_p, [C.c_int]),
            ("mr_routine_nargs", C.c_int, [C.c_int]), ("mr_dynamic_symbols", C.c_int, []),
            ("mr_last_error", C.c_char_p, []), ("mr_call", _P, [C.c_char_p, C.c_int, C.POINTER(_P)]),
            ("mr_real", _P, [C.POINTER(C.c_double), C.c_ssize_t, C.c_int, C.c_int]),
            ("mr_int", _P, [C.POINTER(C.c_int), C.c_ssize_t, C.c_int, C.c_int, C.c_int]),
            ("mr_na_matrix", _P, []), ("mr_strings", _P, [C.POINTER(C.c_char_p), C.c_int]),
            ("mr_list", _P, [C.POINTER(C.c_char_p), C.POINTER(_P), C.c_int]), ("mr_nil", _P, []),
            ("mr_type", C.c_int, [_P]), ("mr_len", C.c_ssize_t, [_P]), ("mr_has_dim", C.c_int, [_P]),
            ("mr_dim", C.c_int, [_P, C.c_int]), ("mr_has_names", C.c_int, [_P]), ("mr_name", C.c_char_p, [_P, C.c_int]),
            ("mr_data", _P, [_P]), ("mr_elt", _P, [_P, C.c_int]), ("mr_string", C.c_char_p, [_P, C.c_int]),
            ("mr_reset", None, []), ("R_init_sparseRGPs", None, [_P]), ("R_unload_sparseRGPs", None, [_P]),
        ]:
            f = getattr(_lib, name)
            f.restype, f.argtypes = res, args
        _lib.R_init_sparseRGPs(_lib.mr_dll())          # what library(sparseRGPs) triggers
    return _lib


def routines():
    L = lib()
    return {L.mr_routine_name(i).decode(): L.mr_routine_nargs(i) for i in range(L.mr_n_routines())}


# ---- Python value -> SEXP ------------------------------------------------------------------------------------
class NAMatrix:
    """R's `matrix()`."""


def to_sexp(v):
    L = lib()
    if v is None:
        return L.mr_nil()
    if isinstance(v, NAMatrix) or v is NAMatrix:
        return L.mr_na_matrix()
    if isinstance(v, str):
        v = [v]
    if isinstance(v, dict):
        names = (C.c_char_p * len(v))(*[k.encode() for k in v])
        vals = (_P * len(v))(*[to_sexp(x) for x in v.values()])
        return L.mr_list(names, vals, len(v))
    if isinstance(v, (list, tuple)) and len(v) and all(isinstance(s, str) for s in v):
        arr = (C.c_char_p * len(v))(*[s.encode() for s in v])
        return L.mr_strings(arr, len(v))
    if isinstance(v, bool):
        a = np.array([int(v)], dtype=np.int32)
        return L.mr_int(a.ctypes.data_as(C.POINTER(C.c_int)), 1, -1, -1, 1)
    a = np.asarray(v)
    if a.dtype.kind in "iub":
        a = np.asfortranarray(a.astype(np.int32))
        nr, nc = (a.shape if a.ndim == 2 else (-1, -1))
        return L.mr_int(a.ctypes.data_as(C.POINTER(C.c_int)), a.size, nr, nc, int(np.asarray(v).dtype.kind == "b"))
    a = np.asfortranarray(a.astype(np.float64))
    nr, nc = (a.shape if a.ndim == 2 else (-1, -1))
    return L.mr_real(a.ctypes.data_as(C.POINTER(C.c_double)), a.size, nr, nc)


class NamedArray(np.ndarray):
    """An atomic vector that came back with a names attribute (e.g. the gradients of the fused entry points)."""
    names = None


def from_sexp(s):
    L = lib()
    t, n = L.mr_type(s), L.mr_len(s)
    if t == NILSXP:
        return None
    if t == VECSXP:
        vals = [from_sexp(L.mr_elt(s, i)) for i in range(n)]
        if L.mr_has_names(s):
            return {L.mr_name(s, i).decode(): vals[i] for i in range(n)}
        return vals
    if t == STRSXP:
        return [L.mr_string(s, i).decode() for i in range(n)]
    ctype, dt = (C.c_double, np.float64) if t == REALSXP else (C.c_int, np.int32)
    a = np.ctypeslib.as_array(C.cast(L.mr_data(s), C.POINTER(ctype)), shape=(max(n, 1),))[:n].astype(dt).copy() if n else np.zeros(0, dt)
    if L.mr_has_dim(s):
        return a.reshape((L.mr_dim(s, 0), L.mr_dim(s, 1)), order="F")
    if L.mr_has_names(s):
        a = a.view(NamedArray)
        a.names = [L.mr_name(s, i).decode() for i in range(n)]
    return a


def dot_call(name, *args):
    """.Call(name, ...) on the registered table; the arena is reclaimed after the result is copied out."""
    L = lib()
    try:
        sx = (_P * max(len(args), 1))(*[to_sexp(a) for a in args])
        r = L.mr_call(name.encode(), len(args), sx)
        if not r:
            raise RError(L.mr_last_error().decode())
        return from_sexp(r)
    finally:
        L.mr_reset()
