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
def knot_bounds(xy):
    """[min - range/10, max + range/10] per input dimension: R/vi_functions.R:175-178 (host side, like the R code)."""
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    return np.stack([lo - (hi - lo) / 10, hi + (hi - lo) / 10], axis=1)


def _fused(model, cov_par, cov_fun, xu, xy, y, mu, delta, want_grad, ctx, dcov_fun_dknot=None, knot_opt=None,
           transform=True):
    ctx = ctx or default_context()
    xy = L.fmat(xy)
    sigma, l, tau, names = _theta(cov_par, cov_fun, xy.shape[1])
    out = {"trans_par": {k: float(np.log(cov_par[k])) for k in names}}
    if dcov_fun_dknot is None or dcov_fun_dknot is False:
        obj, grad = ctx.gauss_obj_grad_host(model, cov_fun, xy, y, mu, xu, sigma, l, tau, delta, want_grad=want_grad)
    else:
        ctx.set_data(xy, y, mu)
        obj, grad, kgrad, tk = ctx.gauss_obj_grad_knots(model, cov_fun, xu, sigma, l, tau, delta,
                                                        knot_bounds(xy) if transform else None, knot_opt)
        out["knot_gradient"], out["trans_knot"] = kgrad, tk
    out["objective"] = obj
    if want_grad:
        out["gradient"] = dict(zip(names, grad))
    return out


def delbo_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None, dcov_fun_dknot=None, knot_opt=None,
                   transform=True, **_ignored):
    """R/vi_functions.R:126-592, transform = TRUE for the covariance parameters: list(gradient, trans_par); the
    ELBO of the same theta (elbo_fun, :64-121) comes back as "objective" from the same two passes.
    dcov_fun_dknot: None / False = R's NA; anything else (R passes dsqexp_dx2 / dsqexp_dx2_ard, chosen by cov_fun)
    adds "knot_gradient" (m*d, knot-major) and "trans_knot" (:425-592).  knot_opt: 0-based indices (R: 1-based),
    None = all knots; `transform` is the knot transform flag of the reference."""
    return _fused("vi", cov_par, cov_fun, xu, xy, y, mu, delta, True, ctx, dcov_fun_dknot, knot_opt, transform)


def elbo_xy(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None):
    """elbo_fun (R/vi_functions.R:64-121) evaluated from (xy, xu) as norm_grad_ascent_vi obtains its arguments
    (:733-760): objective only (one pass)."""
    return _fused("vi", cov_par, cov_fun, xu, xy, y, mu, delta, False, ctx)["objective"]


def obj_fun_norm(mu, Z, Sigma12, Sigma22, y, ff=None, ctx=None, **_ignored):
    """obj_fun_norm(ff = NA, mu, Z, Sigma12, Sigma22, y, ...) (R/laplace_approx_obj_funs.R:6-52) from the matrices the R
    callers materialise; Z and mu recycle like R vectors (length 1 or n)."""
    ctx = ctx or default_context()
    S12, S22 = L.fmat(Sigma12), L.fmat(Sigma22)
    n, m = S12.shape
    Zv, yv = L.fvec(np.atleast_1d(Z)), L.fvec(y)
    muv = L.fvec(np.atleast_1d(mu)) if mu is not None else None
    obj = L.cd()
    L.check(ctx._lib.srgp_gauss_obj_mats(ctx.handle, L.ptr(S12), n, m, L.ptr(S22), L.ptr(Zv), Zv.size, L.ptr(yv),
                                         L.ptr(muv) if muv is not None else None, muv.size if muv is not None else 0,
                                         C.byref(obj)))
    return obj.value


def elbo_fun(mu, Z, Sigma12, Sigma22, y, cov_par, delta, trace_term_fun=None, ff=None, ctx=None, **_ignored):
    """elbo_fun(ff = NA, mu, Z, Sigma12, Sigma22, y, trace_term_fun, cov_par, ...) (R/vi_functions.R:64-121): the
    obj_fun_norm expression plus trace_term_fun(cov_par, Sigma12, Sigma22, delta) (delta arrives through `...`)."""
    tt = (trace_term_fun or globals()["trace_term_fun"])(cov_par, Sigma12, Sigma22, delta, ctx=ctx)
    return obj_fun_norm(mu, Z, Sigma12, Sigma22, y, ctx=ctx) + tt


def dlogp_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None, dcov_fun_dknot=None, knot_opt=None,
                   transform=True, **_ignored):
    """R/laplace_approx_gradient.R:720-1126 (FIC Gaussian) + obj_fun_norm (R/laplace_approx_obj_funs.R:6-52); knot
    arguments as in delbo_dcov_par."""
    return _fused("fic", cov_par, cov_fun, xu, xy, y, mu, delta, True, ctx, dcov_fun_dknot, knot_opt, transform)


def obj_norm_xy(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, ctx=None):
    """obj_fun_norm with Z built as norm_grad_ascent does (R/laplace_gradient_ascent.R:1241-1265)."""
    return _fused("fic", cov_par, cov_fun, xu, xy, y, mu, delta, False, ctx)["objective"]


# ---- optimiser loops ----------------------------------------------------------
def _grad_ascent(model, cov_par_start, cov_fun, xu, xy, y, mu, muu, opt, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, ctx):
    ctx = ctx or default_context()
    xy = L.fmat(xy)
    opt = dict(opt or {})
    delta = opt.get("delta", 1e-6)
    sigma, l, tau, names = _theta(cov_par_start, cov_fun, xy.shape[1])
    ctx.set_data(xy, y, mu)
    knots = not (dcov_fun_dknot is None or dcov_fun_dknot is False)
    res = ctx.gauss_fit(model, cov_fun, xu, sigma, l, tau, delta, opt, opt_theta=dcov_fun_dtheta is not False,
                        opt_knots=knots, knot_bounds=knot_bounds(xy) if knots else None, knot_opt=knot_opt)
    cov_par = dict(zip(names, [res["sigma"], *res["l"], res["tau"]]))
    m = res["xu"].shape[0]
    muu = np.zeros(m) if muu is None else muu
    um, uv = gauss_posterior_u(cov_par, cov_fun, res["xu"], xy, y, mu, muu, delta, vi=(model == "vi"), ctx=ctx)
    return {"cov_par": cov_par, "cov_fun": cov_fun, "xu": res["xu"], "xy": xy, "mu": mu, "muu": muu, "u_mean": um,
            "u_var": uv, "iter": res["iter"], "obj_fun": res["obj_fun"], "grad": res["grad"],
            "cov_par_history": res["cov_par_history"]}


def norm_grad_ascent_vi(cov_par_start, cov_fun, xu, xy, y, mu=0.0, muu=None, opt=None, dcov_fun_dtheta=True,
                        dcov_fun_dknot=None, knot_opt=None, ctx=None, **_ignored):
    """R/vi_functions.R:596-1218: the whole fit in one call -- ADADELTA / gradient ascent on log(theta) (and the
    knots when dcov_fun_dknot is given), one fused GPU evaluation per iteration, then u_mean / u_var.  Returns the
    reference's list (cov_par, xu, u_mean, u_var, iter, obj_fun, grad, cov_par_history ...).  mu / muu must be
    numeric here (the reference's mean(y) default for non-numeric mu is the caller's one-liner)."""
    return _grad_ascent("vi", cov_par_start, cov_fun, xu, xy, y, mu, muu, opt, dcov_fun_dtheta, dcov_fun_dknot,
                        knot_opt, ctx)


def norm_grad_ascent(cov_par_start, cov_fun, xu, xy, y, mu=0.0, muu=None, opt=None, dcov_fun_dtheta=True,
                     dcov_fun_dknot=None, knot_opt=None, ctx=None, **_ignored):
    """R/laplace_gradient_ascent.R:1111-1696 (FIC marginal likelihood), same conventions."""
    return _grad_ascent("fic", cov_par_start, cov_fun, xu, xy, y, mu, muu, opt, dcov_fun_dtheta, dcov_fun_dknot,
                        knot_opt, ctx)


# ---- OAT candidate scoring -----------------------------------------------------
def oat_candidate_scores(cov_par, cov_fun, xu, xy, y, mu, pseudo_prop, delta=1e-6, vi=True, ctx=None):
    """The candidate loop of knot_prop_random_norm_vi (R/vi_functions.R:2211-2298, vi = True) / knot_prop_random_norm
    (R/knot_proposal_functions.R:1283-1353, vi = False): (objective with xu, objective with each row of pseudo_prop
    appended to xu).  NaN marks a candidate the reference would resample (its chol() error)."""
    ctx = ctx or default_context()
    xy = L.fmat(xy)
    sigma, l, tau, _ = _theta(cov_par, cov_fun, xy.shape[1])
    ctx.set_data(xy, y, mu)
    return ctx.oat_scores("vi" if vi else "fic", cov_fun, xu, pseudo_prop, sigma, l, tau, delta)


def _knot_prop_random(norm_opt, pseudo_prop, vi, opt, ctx, y):
    delta = (opt or {}).get("delta", 1e-6)
    xu = np.asarray(norm_opt["xu"], dtype=np.float64).reshape(len(norm_opt["xu"]), -1)
    pseudo_prop = np.asarray(pseudo_prop, dtype=np.float64).reshape(-1, xu.shape[1])
    _, scores = oat_candidate_scores(norm_opt["cov_par"], norm_opt["cov_fun"], xu, norm_opt["xy"], y, norm_opt["mu"],
                                     pseudo_prop, delta, vi, ctx)
    vals = np.concatenate([np.repeat(norm_opt["obj_fun"][-1], len(xu)), scores])
    vals = np.where(np.isnan(vals), -np.inf, vals)
    return np.vstack([xu, pseudo_prop])[int(np.argmax(vals))].reshape(1, -1)


def knot_prop_random_norm_vi(norm_opt, pseudo_prop, opt=None, ctx=None, y=None, **_ignored):
    """R/vi_functions.R:2108-2304 with the sampled rows `pseudo_prop` (:2209) passed in instead of drawn by R's RNG:
    returns obj_fun_x[which.max(c(rep(last objective, nrow(xu)), scores)), ] as a 1 x d matrix.  norm_opt is the
    optimiser's list (xu, cov_par, xy, mu, cov_fun, obj_fun); y arrives through `...` in the reference."""
    return _knot_prop_random(norm_opt, pseudo_prop, True, opt, ctx, y)


def knot_prop_random_norm(norm_opt, pseudo_prop, opt=None, ctx=None, y=None, **_ignored):
    """R/knot_proposal_functions.R:1176-1357 (FIC objective), same conventions."""
    return _knot_prop_random(norm_opt, pseudo_prop, False, opt, ctx, y)


# ---- posterior at the knots and prediction ------------------------------
def gauss_posterior_u(cov_par, cov_fun, xu, xy, y, mu, muu, delta=1e-6, vi=True, ctx=None):
    """u_mean, u_var as the tails of norm_grad_ascent_vi (R/vi_functions.R:1160-1180, vi = True) and
    norm_grad_ascent (R/laplace_gradient_ascent.R:1637-1656, vi = False) compute them."""
    ctx = ctx or default_context()
    xy, xu = L.fmat(xy), L.fmat(xu)
    m = xu.shape[0]
    sigma, l, tau, _ = _theta(cov_par, cov_fun, xy.shape[1])
    ctx.set_data(xy, y, mu)
    muu = L.fvec(np.broadcast_to(np.asarray(muu, dtype=np.float64).reshape(-1), (m,)))
    lv = L.fvec(l)
    um, uv = np.zeros(m), np.zeros((m, m), order="F")
    L.check(ctx._lib.srgp_gauss_posterior_u(ctx.handle, L.VI if vi else L.FIC, L.KERNELS[cov_fun], L.ptr(xu), m,
                                            L.ptr(muu), sigma, L.ptr(lv), tau, float(delta), L.ptr(um), L.ptr(uv)))
    return um, uv


def _predict(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, s22_nugget, var_const, ctx):
    ctx = ctx or default_context()
    xp, xu = L.fmat(x_pred), L.fmat(xu)
    n, d = xp.shape
    m = xu.shape[0]
    sigma, l, tau, _ = _theta(cov_par, cov_fun, d)
    lv = L.fvec(l)
    mu = L.fvec(np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (n,)))
    muu = L.fvec(np.broadcast_to(np.asarray(muu, dtype=np.float64).reshape(-1), (m,)))
    um, uv = L.fvec(u_mean), L.fmat(u_var)
    pm, pv = np.zeros(n), np.zeros(n)
    L.check(ctx._lib.srgp_predict(ctx.handle, L.KERNELS[cov_fun], L.ptr(xp), n, d, L.ptr(mu), L.ptr(xu), m,
                                  L.ptr(muu), L.ptr(um), L.ptr(uv), sigma, L.ptr(lv), float(s22_nugget),
                                  float(var_const), L.ptr(pm), L.ptr(pv)))
    return {"pred_mean": pm, "pred_var": pv}


def predict_vi(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, full_cov=False, family="gaussian", delta=1e-6,
               ctx=None):
    """R/vi_functions.R:1222-1336 (same argument list).  full_cov = TRUE is the dense n_pred x n_pred branch of the
    reference and is not on the GPU path."""
    if family != "gaussian":
        return "Error: Only Gaussian data currently supported."       # the reference returns this string
    if full_cov:
        raise NotImplementedError("full_cov = TRUE builds a dense n_pred x n_pred matrix: out of the hot path")
    tau, sigma = float(cov_par["tau"]), float(cov_par["sigma"])
    return _predict(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, delta, tau ** 2 + sigma ** 2 + delta, ctx)
