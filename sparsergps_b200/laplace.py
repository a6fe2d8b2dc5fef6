"""Host-side mirror of the reference's sparse Laplace functions (R/newtrap_sparseGP.R,
R/laplace_approx_gradient.R:25-339) on top of the C ABI: same argument meaning, GPU evaluation."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .context import default_context
from .vi_functions import _theta

FAMILIES = {"bernoulli": L.BERNOULLI, "poisson": L.POISSON}


def newtrap_sparseGP(start_vals, family, cov_par, cov_fun, xy, xu, y, mu, muu, maxit=1000, tol=1e-6, delta=1e-6,
                     m=1.0, ctx=None):
    """R/newtrap_sparseGP.R:6-186.  Returns the reference's list as a dict: gp, objective_function_values,
    gradient, u_posterior_mean, u_posterior_variance.  `m` is the Poisson offset of the reference's `...`."""
    ctx = ctx or default_context()
    xy, xu = L.fmat(xy), L.fmat(xu)
    n, d = xy.shape
    mk = xu.shape[0]
    sigma, l, tau, _ = _theta(cov_par, cov_fun, d)
    ctx.set_data(xy, y, mu)
    ff = L.fvec(start_vals).copy()
    assert ff.size == n
    hist = np.zeros(int(maxit) + 1)
    nit = L.ci()
    gpsi, um, uv = np.zeros(n), np.zeros(mk), np.zeros((mk, mk), order="F")
    muu = L.fvec(np.broadcast_to(np.asarray(muu, dtype=np.float64).reshape(-1), (mk,)))
    lv = L.fvec(l)
    L.check(ctx._lib.srgp_laplace_newton(ctx.handle, FAMILIES[family], L.KERNELS[cov_fun], L.ptr(xu), mk, L.ptr(muu),
                                         sigma, L.ptr(lv), tau, float(delta), float(m), int(maxit), float(tol),
                                         L.ptr(ff), L.ptr(hist), C.byref(nit), L.ptr(gpsi), L.ptr(um), L.ptr(uv)))
    return {"gp": ff, "objective_function_values": hist[:nit.value].copy(), "gradient": gpsi,
            "u_posterior_mean": um, "u_posterior_variance": uv}


def dlogq_dcov_par(cov_par, cov_fun, xu, xy, y, ff, family, mu, delta=1e-6, m=1.0, ctx=None, dcov_fun_dknot=None,
                   knot_opt=None, transform=True, **_ignored):
    """R/laplace_approx_gradient.R:25-715, transform = TRUE for the covariance parameters: list(gradient, trans_par)
    and, with a dcov_fun_dknot (None / False = R's NA), knot_gradient (m*d, knot-major) and trans_knot (:345-705).
    knot_opt: 0-based indices (R: 1-based), None = all; `transform` is the knot transform flag."""
    from .vi_functions import knot_bounds
    ctx = ctx or default_context()
    xy, xu = L.fmat(xy), L.fmat(xu)
    n, d = xy.shape
    mk = xu.shape[0]
    sigma, l, tau, names = _theta(cov_par, cov_fun, d)
    ctx.set_data(xy, y, mu)
    ffv, lv = L.fvec(ff), L.fvec(l)
    grad = np.zeros(len(names))
    out = {"trans_par": {k: float(np.log(cov_par[k])) for k in names}}
    if dcov_fun_dknot is None or dcov_fun_dknot is False:
        L.check(ctx._lib.srgp_laplace_grad(ctx.handle, FAMILIES[family], L.KERNELS[cov_fun], L.ptr(xu), mk, sigma,
                                           L.ptr(lv), tau, float(delta), float(m), L.ptr(ffv), L.ptr(grad)))
    else:
        kgrad, tk = np.zeros(mk * d), np.zeros((mk, d), order="F")
        lb = ub = None
        if transform:
            kb = knot_bounds(xy)
            lb, ub = L.fvec(kb[:, 0]), L.fvec(kb[:, 1])
        opt, n_opt = None, 0
        if knot_opt is not None:
            opt = np.ascontiguousarray(np.asarray(list(knot_opt), dtype=np.int32))
            n_opt = len(opt)
        L.check(ctx._lib.srgp_laplace_grad_knots(
            ctx.handle, FAMILIES[family], L.KERNELS[cov_fun], L.ptr(xu), mk, sigma, L.ptr(lv), tau, float(delta),
            float(m), L.ptr(ffv), L.ptr(lb) if lb is not None else None, L.ptr(ub) if ub is not None else None,
            opt.ctypes.data_as(C.POINTER(C.c_int)) if n_opt else None, n_opt, L.ptr(grad), L.ptr(kgrad), L.ptr(tk)))
        out["knot_gradient"], out["trans_knot"] = kgrad, tk
    out["gradient"] = dict(zip(names, grad))
    return out


def laplace_grad_ascent(cov_par_start, cov_fun, xu, xy, y, ff, family, mu, muu, opt=None, m=1.0, dcov_fun_dtheta=True,
                        dcov_fun_dknot=None, knot_opt=None, ctx=None, **_ignored):
    """R/laplace_gradient_ascent.R:10-628 in one call: ADADELTA / gradient ascent on log(theta) (and the knots when
    dcov_fun_dknot is given); every iteration is a warm-started Newton mode search + dlogq_dcov_par on the GPU.
    opt: the reference's option names (optim_method, decay, epsilon, eta, learn_rate, maxit, obj_tol, grad_tol,
    maxit_nr, tol_nr, delta).  Returns the reference's list: cov_par, xu, fmax, iter, obj_fun, u_mean, u_var, grad,
    cov_par_history (+ nr_iter)."""
    from .vi_functions import knot_bounds
    ctx = ctx or default_context()
    xy = L.fmat(xy)
    n, d = xy.shape
    o = {"optim_method": "adadelta", "decay": 0.95, "epsilon": 1e-6, "learn_rate": 1e-2, "eta": 1e3, "maxit": 1000,
         "obj_tol": 1e-3, "grad_tol": float("inf"), "maxit_nr": 1000, "delta": 1e-6, "tol_nr": 1e-6}
    o.update({k: v for k, v in (opt or {}).items() if k in o})
    sigma, l, tau, names = _theta(cov_par_start, cov_fun, d)
    ctx.set_data(xy, y, mu)
    xu = np.array(L.fmat(xu), order="F", copy=True)
    mk = xu.shape[0]
    nl = d if cov_fun == "ard" else 1
    lv = np.array(np.broadcast_to(np.asarray(l, dtype=np.float64).reshape(-1), (nl,)), copy=True)
    p = nl + 2
    knots = not (dcov_fun_dknot is None or dcov_fun_dknot is False)
    fo = L.FitOpt({"adadelta": L.OPT_ADADELTA, "ga": L.OPT_GA}[o["optim_method"]], o["decay"], o["epsilon"], o["eta"],
                  o["learn_rate"], int(o["maxit"]), o["obj_tol"], o["grad_tol"], int(dcov_fun_dtheta is not False),
                  int(knots))
    lb = ub = None
    if knots:
        kb = knot_bounds(xy)
        lb, ub = L.fvec(kb[:, 0]), L.fvec(kb[:, 1])
    ko, n_opt = None, 0
    if knot_opt is not None:
        ko = np.ascontiguousarray(np.asarray(list(knot_opt), dtype=np.int32))
        n_opt = len(ko)
    muu_v = L.fvec(np.broadcast_to(np.asarray(muu, dtype=np.float64).reshape(-1), (mk,)))
    ffv = L.fvec(ff).copy()
    assert ffv.size == n
    maxit = int(o["maxit"])
    sg, ta, it = L.cd(float(sigma)), L.cd(float(tau)), C.c_int(0)
    obj_hist, par_hist, grad_hist = np.full(maxit, np.nan), np.full((maxit, p), np.nan), np.full((maxit, p), np.nan)
    nr_iter = np.zeros(maxit, dtype=np.int32)
    um, uv = np.zeros(mk), np.zeros((mk, mk), order="F")
    L.check(ctx._lib.srgp_laplace_fit(
        ctx.handle, FAMILIES[family], L.KERNELS[cov_fun], L.ptr(xu), mk, L.ptr(muu_v), C.byref(sg), L.ptr(lv),
        C.byref(ta), float(o["delta"]), float(m), int(o["maxit_nr"]), float(o["tol_nr"]), C.byref(fo),
        L.ptr(lb) if lb is not None else None, L.ptr(ub) if ub is not None else None,
        ko.ctypes.data_as(C.POINTER(C.c_int)) if n_opt else None, n_opt, L.ptr(ffv), C.byref(it), L.ptr(obj_hist),
        L.ptr(par_hist), L.ptr(grad_hist), nr_iter.ctypes.data_as(C.POINTER(C.c_int)), L.ptr(um), L.ptr(uv)))
    k = it.value
    cov_par = dict(zip(names, [sg.value, *lv, ta.value]))
    return {"cov_par": cov_par, "cov_fun": cov_fun, "xu": xu, "xy": xy, "mu": mu, "muu": muu, "fmax": ffv, "iter": k,
            "obj_fun": obj_hist[:k], "u_mean": um, "u_var": uv, "grad": grad_hist[:k], "cov_par_history": par_hist[:k],
            "nr_iter": nr_iter[:k]}


def knot_prop_random(laplace_opt, pseudo_prop, family, y, opt=None, m=1.0, maxit=1000, tol=1e-6, ctx=None,
                     return_scores=False, **_ignored):
    """R/knot_proposal_functions.R:1001-1175 with the sampled rows `pseudo_prop` (:1093) passed in: one warm-started
    Newton search per candidate on the GPU, then obj_fun_x[which.max(c(rep(last objective, nrow(xu)), scores)), ].
    laplace_opt: the optimiser's list (xu, cov_par, xy, mu, muu, cov_fun, fmax, obj_fun); maxit / tol are
    newtrap_sparseGP's (they reach it through `...` in the reference)."""
    ctx = ctx or default_context()
    delta = (opt or {}).get("delta", 1e-6)
    xy = L.fmat(laplace_opt["xy"])
    xu = L.fmat(laplace_opt["xu"])
    cand = L.fmat(np.asarray(pseudo_prop, dtype=np.float64).reshape(-1, xu.shape[1]))
    sigma, l, tau, _ = _theta(laplace_opt["cov_par"], laplace_opt["cov_fun"], xy.shape[1])
    ctx.set_data(xy, y, laplace_opt["mu"])
    lv, fm = L.fvec(l), L.fvec(laplace_opt["fmax"])
    scores = np.zeros(cand.shape[0])
    L.check(ctx._lib.srgp_laplace_oat_scores(ctx.handle, FAMILIES[family], L.KERNELS[laplace_opt["cov_fun"]], L.ptr(xu),
                                             xu.shape[0], L.ptr(cand), cand.shape[0], sigma, L.ptr(lv), tau, float(delta),
                                             float(m), int(maxit), float(tol), L.ptr(fm), L.ptr(scores)))
    vals = np.concatenate([np.repeat(laplace_opt["obj_fun"][-1], xu.shape[0]), scores])
    vals = np.where(np.isnan(vals), -np.inf, vals)
    pick = np.vstack([xu, cand])[int(np.argmax(vals))].reshape(1, -1)
    return (pick, scores) if return_scores else pick


def predict_laplace(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, full_cov=False, family="gaussian",
                    delta=1e-6, ctx=None):
    """R/laplace_approx_prediction.R:3-123 (same argument list), full_cov = FALSE."""
    from .vi_functions import _predict
    if full_cov:
        raise NotImplementedError("full_cov = TRUE builds a dense n_pred x n_pred matrix: out of the hot path")
    tau, sigma = float(cov_par["tau"]), float(cov_par["sigma"])
    nugget = delta if family == "gaussian" else tau ** 2 + delta
    return _predict(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, nugget, sigma ** 2 + tau ** 2, ctx)
