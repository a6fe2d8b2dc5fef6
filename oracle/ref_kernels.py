"""ctypes loader for the C oracle (oracle/ref_kernels.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Mirrors the R-level signatures of the reference's Rcpp exports
(R/RcppExports.R:7-127 -> src/covariance_functionsC.cpp, src/covariance_function_derivativesC.cpp)
on NumPy arrays, evaluated by the plain-C restatement in ref_kernels.c on the CPU.

PARITY PINNED against the reference's own compiled sources (oracle/ref_native.py, tests/test_reference_pin.py):
see the header of ref_kernels.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ref_kernels.c")
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

SQEXP, EXP, ARD = 0, 1, 2
SIGMA, L, TAU, LC = 0, 1, 2, 3
_KERNELS = {"sqexp": SQEXP, "exp": EXP, "ard": ARD}


def build(force: bool = False) -> str:
    """Compile ref_kernels.c -> oracle/_build/liboracle.so (gcc -O2, single-threaded like Rcpp)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"])
    return _LIB


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ora_make_cov_mat.restype = C.c_int
        _lib.ora_make_cov_mat.argtypes = [C.c_int, _dp, C.c_int64, _dp, C.c_int64, C.c_int,
                                          C.c_double, _dp, C.c_double, C.c_double, _dp]
        _lib.ora_dsig_dtheta.restype = C.c_int
        _lib.ora_dsig_dtheta.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int64, _dp, C.c_int64,
                                         C.c_int, C.c_double, _dp, C.c_double, _dp]
        for name in ("ora_cov_fun_sqrd_exp", "ora_cov_fun_exp", "ora_dsqexp_dsigma", "ora_dsqexp_dl",
                     "ora_dexp_dsigma", "ora_dexp_dl"):
            f = getattr(_lib, name)
            f.restype = C.c_double
            f.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double]
        for name in ("ora_cov_fun_sqrd_exp_ard", "ora_dsqexp_dsigma_ard"):
            f = getattr(_lib, name)
            f.restype = C.c_double
            f.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp]
        _lib.ora_dsqexp_dl_ard.restype = C.c_double
        _lib.ora_dsqexp_dl_ard.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp, C.c_int]
        _lib.ora_dk_dtau.restype = C.c_double
        _lib.ora_dk_dtau.argtypes = [_dp, _dp, C.c_int, C.c_double]
    return _lib


def _f(a):
    """R matrix -> column-major float64 buffer."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def _as_matrix(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return _f(x)


def _is_empty(x_pred) -> bool:
    """The reference's `matrix()` sentinel: a 1x1 NA (src/covariance_functionsC.cpp:81)."""
    if x_pred is None:
        return True
    xp = np.asarray(x_pred, dtype=np.float64)
    return xp.size >= 1 and bool(np.isnan(xp.reshape(-1)[0]))


def _lvec(cov_par, cov_fun, lnames, d):
    if cov_fun == "ard":
        return np.array([float(cov_par[str(nm)]) for nm in lnames], dtype=np.float64)
    return np.array([float(cov_par["l"])], dtype=np.float64)


def _assemble(x, x_pred, cov_par, cov_fun, delta, lnames):
    if cov_fun not in _KERNELS:
        sys.stderr.write("Error: invalid covariance function")
        return np.zeros((0, 0))
    x = _as_matrix(x)
    n1, d = x.shape
    l = _lvec(cov_par, cov_fun, lnames, d)
    sigma = float(cov_par["sigma"])
    if _is_empty(x_pred):
        out = np.empty((n1, n1), order="F")
        rc = lib().ora_make_cov_mat(_KERNELS[cov_fun], _p(x), n1, None, 0, d, sigma, _p(l),
                                    float(cov_par["tau"]), float(delta), _p(out))
    else:
        xp = _as_matrix(x_pred)
        n2 = xp.shape[0]
        out = np.empty((n1, n2), order="F")
        # tau is only read on the self-covariance diagonal (src/covariance_functionsC.cpp:91)
        rc = lib().ora_make_cov_mat(_KERNELS[cov_fun], _p(x), n1, _p(xp), n2, d, sigma, _p(l),
                                    float(cov_par.get("tau", 0.0)), float(delta), _p(out))
    assert rc == 0
    return out


def make_cov_matC(x, x_pred, cov_par, cov_fun, delta):
    """src/covariance_functionsC.cpp:72-169."""
    if cov_fun not in ("sqexp", "exp"):
        sys.stderr.write("Error: invalid covariance function")
        return np.zeros((0, 0))
    return _assemble(x, x_pred, cov_par, cov_fun, delta, None)


def make_cov_mat_ardC(x, x_pred, cov_par, cov_fun, delta, lnames):
    """src/covariance_functionsC.cpp:191-252."""
    if cov_fun != "ard":
        sys.stderr.write("Error: invalid covariance function")
        return np.zeros((0, 0))
    return _assemble(x, x_pred, cov_par, cov_fun, delta, lnames)


def _dsig(x, x_pred, cov_par, cov_fun, par_name, lnames):
    x = _as_matrix(x)
    n1, d = x.shape
    l = _lvec(cov_par, cov_fun, lnames, d)
    comp0 = -1
    if par_name == "sigma":
        par = SIGMA
    elif par_name == "tau":
        par = TAU
    elif cov_fun == "ard" and par_name in [str(s) for s in lnames]:
        par = LC
        # the reference keeps the LAST matching name (covariance_function_derivativesC.cpp:596-605)
        comp0 = max(i for i, s in enumerate(lnames) if str(s) == par_name)
    elif cov_fun != "ard" and par_name == "l":
        par = L
    else:
        par = 99
    self_ = _is_empty(x_pred)
    if self_:
        xp, n2, xpp = None, n1, None
    else:
        xp = _as_matrix(x_pred)
        n2, xpp = xp.shape[0], _p(xp)
    out = np.empty((n1, n2), order="F")
    rc = lib().ora_dsig_dtheta(_KERNELS[cov_fun], par, comp0, _p(x), n1, xpp, 0 if self_ else n2, d,
                               float(cov_par["sigma"]), _p(l), float(cov_par.get("tau", 0.0)), _p(out))
    if rc != 0:
        sys.stderr.write("Error: invalid parameter name for chosen covariance function")
        return np.zeros((0, 0))
    return out


def dsig_dthetaC(x, x_pred, cov_par, cov_fun, par_name):
    """src/covariance_function_derivativesC.cpp:307-552."""
    if cov_fun not in ("sqexp", "exp"):
        sys.stderr.write("Error: invalid covariance function")
        return np.zeros((0, 0))
    return _dsig(x, x_pred, cov_par, cov_fun, par_name, None)


def dsig_dtheta_ardC(x, x_pred, cov_par, cov_fun, par_name, lnames):
    """src/covariance_function_derivativesC.cpp:555-722."""
    if cov_fun != "ard":
        sys.stderr.write("Error: invalid covariance function")
        return np.zeros((0, 0))
    return _dsig(x, x_pred, cov_par, cov_fun, par_name, lnames)


def real_to_pos(x):
    """src/covariance_function_derivativesC.cpp:11-13."""
    return np.exp(np.asarray(x, dtype=np.float64))


def pos_to_real(x):
    """src/covariance_function_derivativesC.cpp:19-21."""
    return np.log(np.asarray(x, dtype=np.float64))


def real_to_bounded(x, ub, lb):
    """src/covariance_function_derivativesC.cpp:27-29."""
    x = np.asarray(x, dtype=np.float64)
    return (np.asarray(ub) * np.exp(x) + np.asarray(lb)) / (np.exp(x) + 1)
