"""Host-side mirror of the reference's R wrappers around `.Call` (R/RcppExports.R:7-127).

Same names, argument order and error behaviour as the R functions; every matrix-valued routine runs on the
GPU through the C ABI (include/srgp.h).  cov_par is a dict keyed like the R named list ("sigma", "l" or
"l1".."ld", "tau"); x_pred = None (or an array whose first entry is NaN) is R's `matrix()` sentinel.
"""
from __future__ import annotations

import sys

import numpy as np

from . import _lib as L
from .context import default_context


def _is_empty(x_pred) -> bool:
    if x_pred is None:
        return True
    xp = np.asarray(x_pred, dtype=np.float64)
    return xp.size >= 1 and bool(np.isnan(xp.reshape(-1)[0]))


def _invalid(msg):
    # reference: message on Rcerr and a 0 x 0 matrix, no error (src/covariance_functionsC.cpp:161-168)
    sys.stderr.write(msg)
    return np.zeros((0, 0))


def _lvec(cov_par, cov_fun, lnames):
    if cov_fun == "ard":
        return np.array([float(cov_par[str(nm)]) for nm in lnames], dtype=np.float64)
    return np.array([float(cov_par["l"])], dtype=np.float64)


def _assemble(x, x_pred, cov_par, cov_fun, delta, lnames, ctx):
    ctx = ctx or default_context()
    x = L.fmat(x)
    n1, d = x.shape
    l = _lvec(cov_par, cov_fun, lnames)
    if _is_empty(x_pred):
        out = np.empty((n1, n1), order="F")
        st = ctx._lib.srgp_make_cov_mat(ctx.handle, L.KERNELS[cov_fun], L.ptr(x), n1, None, 0, d,
                                        float(cov_par["sigma"]), L.ptr(l), float(cov_par["tau"]), float(delta),
                                        L.ptr(out))
    else:
        xp = L.fmat(x_pred)
        out = np.empty((n1, xp.shape[0]), order="F")
        st = ctx._lib.srgp_make_cov_mat(ctx.handle, L.KERNELS[cov_fun], L.ptr(x), n1, L.ptr(xp), xp.shape[0], d,
                                        float(cov_par["sigma"]), L.ptr(l), float(cov_par.get("tau", 0.0)),
                                        float(delta), L.ptr(out))
    L.check(st)
    return out


def make_cov_matC(x, x_pred, cov_par, cov_fun, delta, ctx=None):
    """R/RcppExports.R `make_cov_matC` -> src/covariance_functionsC.cpp:72-169."""
    if cov_fun not in ("sqexp", "exp"):
        return _invalid("Error: invalid covariance function")
    return _assemble(x, x_pred, cov_par, cov_fun, delta, None, ctx)


def make_cov_mat_ardC(x, x_pred, cov_par, cov_fun, delta, lnames, ctx=None):
    """R/RcppExports.R `make_cov_mat_ardC` -> src/covariance_functionsC.cpp:191-252."""
    if cov_fun != "ard":
        return _invalid("Error: invalid covariance function")
    return _assemble(x, x_pred, cov_par, cov_fun, delta, lnames, ctx)


def _dsig(x, x_pred, cov_par, cov_fun, par_name, lnames, ctx):
    ctx = ctx or default_context()
    x = L.fmat(x)
    n1, d = x.shape
    l = _lvec(cov_par, cov_fun, lnames)
    comp0 = -1
    if par_name == "sigma":
        par = L.PAR_SIGMA
    elif par_name == "tau":
        par = L.PAR_TAU
    elif cov_fun == "ard" and par_name in [str(s) for s in lnames]:
        par = L.PAR_LC
        comp0 = max(i for i, s in enumerate(lnames) if str(s) == par_name)
    elif cov_fun != "ard" and par_name == "l":
        par = L.PAR_L
    else:
        par = 99
    if _is_empty(x_pred):
        xpp, n2, ncol = None, 0, n1
    else:
        xp = L.fmat(x_pred)
        xpp, n2, ncol = L.ptr(xp), xp.shape[0], xp.shape[0]
    out = np.empty((n1, ncol), order="F")
    st = ctx._lib.srgp_dsig_dtheta(ctx.handle, L.KERNELS[cov_fun], par, comp0, L.ptr(x), n1, xpp, n2, d,
                                   float(cov_par["sigma"]), L.ptr(l), float(cov_par.get("tau", 0.0)), L.ptr(out))
    if st == L.ERR_UNKNOWN_PAR:
        return _invalid("Error: invalid parameter name for chosen covariance function")
    L.check(st)
    return out


def dsig_dthetaC(x, x_pred, cov_par, cov_fun, par_name, ctx=None):
    """R/RcppExports.R `dsig_dthetaC` -> src/covariance_function_derivativesC.cpp:307-552."""
    if cov_fun not in ("sqexp", "exp"):
        return _invalid("Error: invalid covariance function")
    return _dsig(x, x_pred, cov_par, cov_fun, par_name, None, ctx)


def dsig_dtheta_ardC(x, x_pred, cov_par, cov_fun, par_name, lnames, ctx=None):
    """R/RcppExports.R `dsig_dtheta_ardC` -> src/covariance_function_derivativesC.cpp:555-722."""
    if cov_fun != "ard":
        return _invalid("Error: invalid covariance function")
    return _dsig(x, x_pred, cov_par, cov_fun, par_name, lnames, ctx)


# ---- transforms and per-pair scalars (host code in the library) -----------------------------------
def _vec_call(fn, x):
    x = L.fvec(x)
    out = np.empty_like(x)
    fn(L.ptr(x), x.size, L.ptr(out))
    return out


def real_to_pos(x):
    """src/covariance_function_derivativesC.cpp:11-13."""
    return _vec_call(L.load().srgp_real_to_pos, x)


def pos_to_real(x):
    """src/covariance_function_derivativesC.cpp:19-21."""
    return _vec_call(L.load().srgp_pos_to_real, x)


def real_to_bounded(x, ub, lb):
    """src/covariance_function_derivativesC.cpp:27-29."""
    x = L.fvec(x)
    ub = L.fvec(np.broadcast_to(np.asarray(ub, dtype=np.float64), x.shape))
    lb = L.fvec(np.broadcast_to(np.asarray(lb, dtype=np.float64), x.shape))
    out = np.empty_like(x)
    L.load().srgp_real_to_bounded(L.ptr(x), L.ptr(ub), L.ptr(lb), x.size, L.ptr(out))
    return out


def _pair(x1, x2):
    a, b = L.fvec(x1), L.fvec(x2)
    assert a.size == b.size
    return a, b, a.size


def cov_fun_sqrd_expC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    return L.load().srgp_cov_fun_sqrd_exp(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))


def cov_fun_sqrd_exp_ardC(x1, x2, cov_par, lnames):
    a, b, d = _pair(x1, x2)
    l = _lvec(cov_par, "ard", lnames)
    return L.load().srgp_cov_fun_sqrd_exp_ard(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), L.ptr(l))


def cov_fun_expC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    return L.load().srgp_cov_fun_exp(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))


def _deriv_list(value, par):
    # the Rcpp helpers return list(derivative, trans_par = log(par), inv_trans_par = real_to_pos(par)): the R
    # closure is applied to the UNtransformed value, i.e. exp(par), not par
    # (src/covariance_function_derivativesC.cpp:49,80,101,136,168 -- pinned by tests/golden/rcpp_layer.json)
    return {"derivative": value, "trans_par": float(np.log(par)), "inv_trans_par": float(real_to_pos([par])[0])}


def dsqexp_dsigmaC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dsqexp_dsigma(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))
    return _deriv_list(v, cov_par["sigma"])


def dsqexp_dsigma_ardC(x1, x2, cov_par, lnames):
    a, b, d = _pair(x1, x2)
    l = _lvec(cov_par, "ard", lnames)
    v = L.load().srgp_dsqexp_dsigma_ard(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), L.ptr(l))
    return _deriv_list(v, cov_par["sigma"])


def dsqexp_dlC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dsqexp_dl(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))
    return _deriv_list(v, cov_par["l"])


def dsqexp_dl_ardC(x1, x2, cov_par, lnames, comp):
    """`comp` is 1-based, as in the reference (decremented at covariance_function_derivativesC.cpp:121)."""
    a, b, d = _pair(x1, x2)
    l = _lvec(cov_par, "ard", lnames)
    c0 = int(comp) - 1
    v = L.load().srgp_dsqexp_dl_ard(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), L.ptr(l), c0)
    return _deriv_list(v, l[c0])


def dsqexp_dtauC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dsqexp_dtau(L.ptr(a), L.ptr(b), d, float(cov_par["tau"]))
    return _deriv_list(v, cov_par["tau"])


def dexp_dsigmaC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dexp_dsigma(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))
    return _deriv_list(v, cov_par["sigma"])


def dexp_dlC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dexp_dl(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), float(cov_par["l"]))
    return _deriv_list(v, cov_par["l"])


def dexp_dtauC(x1, x2, cov_par):
    a, b, d = _pair(x1, x2)
    v = L.load().srgp_dexp_dtau(L.ptr(a), L.ptr(b), d, float(cov_par["tau"]))
    return _deriv_list(v, cov_par["tau"])


def _dx2(fn, x1, x2, cov_par, lb, ub, larg):
    a, b, d = _pair(x1, x2)
    lb = L.fvec(np.broadcast_to(np.asarray(lb, dtype=np.float64), a.shape))
    ub = L.fvec(np.broadcast_to(np.asarray(ub, dtype=np.float64), a.shape))
    deriv, tp = np.empty(d), np.empty(d)
    fn(L.ptr(a), L.ptr(b), d, float(cov_par["sigma"]), larg, L.ptr(lb), L.ptr(ub), L.ptr(deriv), L.ptr(tp))
    return {"derivative": deriv, "trans_par": tp, "inv_trans_par": real_to_bounded(b, ub, lb)}


def dsqexp_dx2C(x1, x2, cov_par, lb, ub):
    return _dx2(L.load().srgp_dsqexp_dx2, x1, x2, cov_par, lb, ub, float(cov_par["l"]))


def dsqexp_dx2_ardC(x1, x2, cov_par, lb, ub, lnames):
    l = _lvec(cov_par, "ard", lnames)
    return _dx2(L.load().srgp_dsqexp_dx2_ard, x1, x2, cov_par, lb, ub, L.ptr(l))
