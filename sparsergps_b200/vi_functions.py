"""Host-side mirror of the reference's R-level functions on the hot path (R/vi_functions.R,
R/laplace_approx_obj_funs.R, R/laplace_approx_gradient.R): same names and argument meaning, evaluated on the GPU
through the C ABI.  The three trace-term functions keep the reference signatures exactly (north_star); the
objective / gradient functions take the reference's arguments and run the fused two-pass path, so no n x m
matrix is ever built.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .context import default_context


def _lnames(d):
    return ["l%d" % (i + 1) for i in range(d)]


def _theta(cov_par, cov_fun, d):
    if cov_fun == "ard":
        l = [float(cov_par[nm]) for nm in _lnames(d)]
        names = ["sigma"] + _lnames(d) + ["tau"]
    elif cov_fun == "sqexp":
        l = [float(cov_par["l"])]
        names = ["sigma", "l", "tau"]
    else:
        raise ValueError("Error: invalid covariance function")
    return float(cov_par["sigma"]), l, float(cov_par["tau"]), names


# ---- trace term (signatures fixed by the reference) --------------------------------------------------
def trace_term_fun(cov_par, Sigma12, Sigma22, delta, ctx=None):
    """R/vi_functions.R:14-27: -(1/(2 tau^2)) * sum(sigma^2 + delta - rowSums(Sigma12 * t(solve(Sigma22, t(Sigma12)))))."""
    ctx = ctx or default_context()
    S12, S22 = L.fmat(Sigma12), L.fmat(Sigma22)
    n, m = S12.shape
    out = L.cd()
    L.check(ctx._lib.srgp_trace_term(ctx.handle, float(cov_par["sigma"]), float(cov_par["tau"]), float(delta),
                                     L.ptr(S12), n, m, L.ptr(S22), C.byref(out)))
    return out.value


def dtrace_term_dtau(cov_par, trace_term):
    """R/vi_functions.R:38-44 (derivative wrt log tau)."""
    return L.load().srgp_dtrace_term_dtau(float(trace_term))


def dtrace_term_dcov_par(cov_par, A_trace, ctx=None):
    """R/vi_functions.R:54-60."""
    ctx = ctx or default_context()
    a = L.fvec(A_trace)
    out = L.cd()
    L.check(ctx._lib.srgp_dtrace_term_dcov_par(ctx.handle, float(cov_par["tau"]), L.ptr(a), a.size, C.byref(out)))
    return out.value


def omega_dk_reduce(cov_par, cov_fun, xy, xu, Omega, ctx=None):
    """sum_ij Omega_ij * dSigma12_ij/dlog(theta) for every theta, dSigma12 never materialised -- what each
    `dSigma12_dtheta` use in delbo_dcov_par (R/vi_functions.R:344-398) reduces to.  Returns a dict by name."""
    ctx = ctx or default_context()
    xy, xu, Om = L.fmat(xy), L.fmat(xu), L.fmat(Omega)
    n, d = xy.shape
    m = xu.shape[0]
    sigma, l, tau, names = _theta(cov_par, cov_fun, d)
    lv = L.fvec(l)
    out = np.zeros(len(names))
    L.check(ctx._lib.srgp_omega_dk_reduce(ctx.handle, L.KERNELS[cov_fun], L.ptr(xy), n, L.ptr(xu), m, d, sigma,
                                          L.ptr(lv), tau, L.ptr(Om), L.ptr(out)))
    return dict(zip(names, out))


# ---- fused objective + gradient -----------------------------------------------------------------------
def _fused(model, cov_par, cov_fun, xu, xy, y, mu, delta, want_grad, ctx):
    ctx = ctx or default_context()
    xy = L.fmat(xy)
    sigma, l, tau, names = _theta(cov_par, cov_fun, xy.shape[1])
    obj, grad = ctx.gauss_obj_grad_host(model, cov_fun, xy, y, mu, xu, sigma, l, tau, delta, want_grad=want_grad)
    out = {"objective": obj, "trans_par": {k: float(np.log(cov_par[k])) for k in names}}
    if want_grad:
        out["gradient"] = dict(zip(names, grad))
    return out


def delbo_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None, **_ignored):
    """R/vi_functions.R:126-420 with dcov_fun_dknot = NA, transform = TRUE: list(gradient, trans_par); the
    ELBO of the same theta (elbo_fun, :64-121) comes back as "objective" from the same two passes."""
    return _fused("vi", cov_par, cov_fun, xu, xy, y, mu, delta, True, ctx)


def elbo_xy(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None):
    """elbo_fun (R/vi_functions.R:64-121) evaluated from (xy, xu) as norm_grad_ascent_vi obtains its arguments
    (:733-760): objective only (one pass)."""
    return _fused("vi", cov_par, cov_fun, xu, xy, y, mu, delta, False, ctx)["objective"]


def dlogp_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None, **_ignored):
    """R/laplace_approx_gradient.R:720-968 (FIC Gaussian) + obj_fun_norm (R/laplace_approx_obj_funs.R:6-52)."""
    return _fused("fic", cov_par, cov_fun, xu, xy, y, mu, delta, True, ctx)


def obj_norm_xy(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None):
    """obj_fun_norm with Z built as norm_grad_ascent does (R/laplace_gradient_ascent.R:1241-1265)."""
    return _fused("fic", cov_par, cov_fun, xu, xy, y, mu, delta, False, ctx)["objective"]
