"""The reference's OWN kernel sources, compiled here -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`build()` compiles /root/reference/src/covariance_functionsC.cpp and covariance_function_derivativesC.cpp,
unmodified and where they lie, together with oracle/ref_glue.cpp against the Rcpp stand-in header
oracle/rcpp_shim/Rcpp.h into oracle/_ref/libsparseRGPs_ref.so (git-ignored; it travels to the GPU box with the
gpurun snapshot, /root/reference does not). The functions below carry the R-level signatures of
R/RcppExports.R:7-127 on NumPy arrays and call the compiled reference code.

What this pins: every per-element kernel and the four matrix builders of SURVEY.md 8(a) rows a1-a13 (the whole
Rcpp layer). What it cannot pin: the R-level model algebra (rows a14-a25), for which no R interpreter exists here.

Only tests/, tests/tools/make_golden.py, __graft_entry__.build() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC_DIR = "/root/reference/src"
_REF_SOURCES = ["covariance_functionsC.cpp", "covariance_function_derivativesC.cpp"]
_GLUE = os.path.join(_HERE, "ref_glue.cpp")
_SHIM = os.path.join(_HERE, "rcpp_shim")
LIB = os.path.join(_HERE, "_ref", "libsparseRGPs_ref.so")


def sources_present() -> bool:
    return all(os.path.exists(os.path.join(REF_SRC_DIR, s)) for s in _REF_SOURCES)


def available() -> bool:
    return os.path.exists(LIB) or sources_present()


def build(force: bool = False) -> str | None:
    """g++ -O2 (R's default CXXFLAGS are `-g -O2`) on the reference sources in place. Returns the library path,
    or None when /root/reference is absent and no prebuilt library travelled with the snapshot."""
    if not sources_present():
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(REF_SRC_DIR, s) for s in _REF_SOURCES] + [_GLUE, os.path.join(_SHIM, "Rcpp.h")]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-w", "-fPIC", "-shared", "-I", _SHIM, _GLUE]
                              + deps[:2] + ["-o", LIB])
    return LIB


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_sp = C.POINTER(C.c_char_p)


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = C.CDLL(path)
        _lib.ref_last_error.restype = C.c_char_p
        for name in ("ref_transform", "ref_cov_fun", "ref_pair_derivative", "ref_make_cov_mat", "ref_dsig_dtheta"):
            getattr(_lib, name).restype = C.c_int
        _lib.ref_transform.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, _dp]
        _lib.ref_cov_fun.argtypes = [C.c_int, _dp, _dp, C.c_int, _sp, _dp, C.c_int, _sp, C.c_int, _dp]
        _lib.ref_pair_derivative.argtypes = [C.c_int, _dp, _dp, C.c_int, _sp, _dp, C.c_int, _sp, C.c_int, C.c_double,
                                             _dp, _dp, _dp, _dp, _dp, _ip]
        _lib.ref_make_cov_mat.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, _sp, _dp, C.c_int, C.c_char_p,
                                          C.c_double, _sp, C.c_int, _dp, C.c_longlong, _ip, _ip]
        _lib.ref_dsig_dtheta.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, _sp, _dp, C.c_int, C.c_char_p,
                                         C.c_char_p, _sp, C.c_int, _dp, C.c_longlong, _ip, _ip]
    return _lib


class ReferenceError_(RuntimeError):
    """A C++ exception of the reference (an R error behind BEGIN_RCPP/END_RCPP)."""


def _check(rc):
    if rc != 0:
        raise ReferenceError_(lib().ref_last_error().decode())


def _f(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def _v(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def _p(a):
    return a.ctypes.data_as(_dp)


def _strs(names):
    arr = (C.c_char_p * max(len(names), 1))()
    for i, s in enumerate(names):
        arr[i] = str(s).encode()
    return arr


def _list(cov_par):
    names = list(cov_par)
    vals = np.array([float(cov_par[k]) for k in names], dtype=np.float64)
    return _strs(names), vals, len(names)


def _is_empty(x_pred) -> bool:
    if x_pred is None:
        return True
    xp = np.asarray(x_pred, dtype=np.float64)
    return xp.size >= 1 and bool(np.isnan(xp.reshape(-1)[0]))


# ------------------------------------------------------------------------------------------- transforms (a13)
def real_to_pos(x):
    x = _v(x)
    out = np.empty_like(x)
    _check(lib().ref_transform(0, _p(x), None, None, x.size, _p(out)))
    return out


def pos_to_real(x):
    x = _v(x)
    out = np.empty_like(x)
    _check(lib().ref_transform(1, _p(x), None, None, x.size, _p(out)))
    return out


def real_to_bounded(x, ub, lb):
    x, ub, lb = _v(x), _v(ub), _v(lb)
    out = np.empty_like(x)
    _check(lib().ref_transform(2, _p(x), _p(ub), _p(lb), x.size, _p(out)))
    return out


# ------------------------------------------------------------------------------------------- per-pair (a1-a3, a6-a10)
def _cov(which, x1, x2, cov_par, lnames=()):
    x1, x2 = _v(x1), _v(x2)
    nm, vals, npar = _list(cov_par)
    out = C.c_double()
    out_arr = np.empty(1)
    _check(lib().ref_cov_fun(which, _p(x1), _p(x2), x1.size, nm, _p(vals), npar, _strs(lnames), len(lnames),
                             _p(out_arr)))
    return float(out_arr[0])


def cov_fun_sqrd_expC(x1, x2, cov_par):
    return _cov(0, x1, x2, cov_par)


def cov_fun_sqrd_exp_ardC(x1, x2, cov_par, lnames):
    return _cov(1, x1, x2, cov_par, list(lnames))


def cov_fun_expC(x1, x2, cov_par):
    return _cov(2, x1, x2, cov_par)


def _pair(which, x1, x2, cov_par, lnames=(), comp=0.0, lb=None, ub=None):
    x1, x2 = _v(x1), _v(x2)
    d = x1.size
    nm, vals, npar = _list(cov_par)
    dv, tp, ip = np.full(d, np.nan), np.full(d, np.nan), np.full(d, np.nan)
    ln = C.c_int()
    lbp = _p(_v(lb)) if lb is not None else None
    ubp = _p(_v(ub)) if ub is not None else None
    _check(lib().ref_pair_derivative(which, _p(x1), _p(x2), d, nm, _p(vals), npar, _strs(lnames), len(lnames),
                                     float(comp), lbp, ubp, _p(dv), _p(tp), _p(ip), C.byref(ln)))
    k = ln.value
    if k == 1:
        return {"derivative": float(dv[0]), "trans_par": float(tp[0]), "inv_trans_par": float(ip[0])}
    return {"derivative": dv[:k].copy(), "trans_par": tp[:k].copy(), "inv_trans_par": ip[:k].copy()}


def dsqexp_dsigmaC(x1, x2, cov_par):
    return _pair(0, x1, x2, cov_par)


def dsqexp_dsigma_ardC(x1, x2, cov_par, lnames):
    return _pair(1, x1, x2, cov_par, list(lnames))


def dsqexp_dlC(x1, x2, cov_par):
    return _pair(2, x1, x2, cov_par)


def dsqexp_dl_ardC(x1, x2, cov_par, lnames, comp):
    return _pair(3, x1, x2, cov_par, list(lnames), comp=comp)


def dsqexp_dtauC(x1, x2, cov_par):
    return _pair(4, x1, x2, cov_par)


def dsqexp_dx2C(x1, x2, cov_par, lb, ub):
    return _pair(5, x1, x2, cov_par, lb=lb, ub=ub)


def dsqexp_dx2_ardC(x1, x2, cov_par, lb, ub, lnames):
    return _pair(6, x1, x2, cov_par, list(lnames), lb=lb, ub=ub)


def dexp_dsigmaC(x1, x2, cov_par):
    return _pair(7, x1, x2, cov_par)


def dexp_dlC(x1, x2, cov_par):
    return _pair(8, x1, x2, cov_par)


def dexp_dtauC(x1, x2, cov_par):
    return _pair(9, x1, x2, cov_par)


# ------------------------------------------------------------------------------------------- matrices (a4, a5, a11, a12)
def _matrix(fn, x, x_pred, cov_par, cov_fun, extra, lnames):
    x = _f(x)
    n1, d = x.shape
    if _is_empty(x_pred):
        xp, n2, xpp = None, n1, None
    else:
        xp = _f(x_pred)
        n2, xpp = xp.shape[0], _p(xp)
    nm, vals, npar = _list(cov_par)
    out = np.empty(n1 * max(n2, 1), dtype=np.float64)
    nr, nc = C.c_int(), C.c_int()
    ln = None if lnames is None else _strs(list(lnames))
    nl = 0 if lnames is None else len(lnames)
    _check(fn(_p(x), n1, d, xpp, 0 if xp is None else n2, nm, _p(vals), npar, str(cov_fun).encode(), *extra,
              ln, nl, _p(out), out.size, C.byref(nr), C.byref(nc)))
    return out[: nr.value * nc.value].reshape((nr.value, nc.value), order="F").copy(order="F")


def make_cov_matC(x, x_pred, cov_par, cov_fun, delta):
    return _matrix(lib().ref_make_cov_mat, x, x_pred, cov_par, cov_fun, (float(delta),), None)


def make_cov_mat_ardC(x, x_pred, cov_par, cov_fun, delta, lnames):
    return _matrix(lib().ref_make_cov_mat, x, x_pred, cov_par, cov_fun, (float(delta),), list(lnames))


def dsig_dthetaC(x, x_pred, cov_par, cov_fun, par_name):
    return _matrix(lib().ref_dsig_dtheta, x, x_pred, cov_par, cov_fun, (str(par_name).encode(),), None)


def dsig_dtheta_ardC(x, x_pred, cov_par, cov_fun, par_name, lnames):
    return _matrix(lib().ref_dsig_dtheta, x, x_pred, cov_par, cov_fun, (str(par_name).encode(),), list(lnames))
