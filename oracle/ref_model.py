"""NumPy literal transcription of the reference's R model algebra -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, statement by statement, the pure-R functions of luisdamiano/sparseRGPs that sit on the hot path
(SURVEY.md section 8a, rows a14-a25).  Each function cites the reference file:line it follows.  R's `solve`
(DGESV, LU) maps to numpy.linalg.solve, `chol` (DPOTRF, upper) to the transpose of numpy.linalg.cholesky,
`det` (LU) to numpy.linalg.det, so the numerical route is the reference's, including its quirks
(SURVEY.md Appendix C).  The dense covariance / derivative matrices come from the C restatement of the Rcpp
kernels (ref_kernels.c) exactly as the R code obtains them through `.Call`.

PARITY PINNED to the reference's own R sources, with one stated caveat.  No R installation exists in this image and
the reference ships no tests or fixtures, so the reference's R files (/root/reference/R/*.R) are executed, unmodified
and where they lie, by the mini-R interpreter of oracle/mini_r with the Rcpp exports served by the reference's
compiled C++ (oracle/ref_r.py).  tests/tools/make_golden_r.py records what those R functions return
(tests/golden/r_level.*) and tests/test_reference_r.py holds every function below to it at 1e-10 (Gaussian rows),
1e-9 (Laplace rows, optimiser trajectories).  Caveat: the interpreter is ours, not GNU R -- it implements the language
subset these files use from the R Language Definition (known-answer tests: tests/test_mini_r.py) and reaches LAPACK
through NumPy where R reaches it through its own BLAS; what is pinned is the reference's R SOURCE, statement for
statement, not GNU R's binary.  The first run of this pin found one bug in the interpreter (plogis(log.p=)) and none in
this file.  Further checks: analytic known answers, finite differences, and the independent reduced-form algebra
in oracle/reduced_model.py (tests/test_oracle.py).  The Rcpp layer it calls is pinned bit for bit against the
reference's compiled C++ (tests/test_reference_pin.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln

from . import ref_kernels as rk

make_cov_matC = rk.make_cov_matC
make_cov_mat_ardC = rk.make_cov_mat_ardC
dsig_dthetaC = rk.dsig_dthetaC
dsig_dtheta_ardC = rk.dsig_dtheta_ardC
pos_to_real = rk.pos_to_real
real_to_pos = rk.real_to_pos


# --------------------------------------------------------------------------------------------------
# R primitives
# --------------------------------------------------------------------------------------------------
def solve(a, b=None):
    """R solve(a, b): LAPACK DGESV; solve(a) is the explicit inverse."""
    a = np.asarray(a, dtype=np.float64)
    if b is None:
        return np.linalg.solve(a, np.eye(a.shape[0]))
    return np.linalg.solve(a, np.asarray(b, dtype=np.float64))


def chol(x):
    """R chol(x): upper-triangular R with t(R) %*% R == x; raises like R's `chol` error when not PD."""
    return np.linalg.cholesky(np.asarray(x, dtype=np.float64)).T


def det(x):
    return np.linalg.det(np.asarray(x, dtype=np.float64))


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def _rows(v, M):
    """R's `v * M` for a length-nrow(M) vector v: recycling down the columns == row scaling."""
    return np.asarray(v, dtype=np.float64).reshape(-1, 1) * M


def lnames_for(xy):
    return ["l%d" % (i + 1) for i in range(np.asarray(xy).reshape(len(xy), -1).shape[1])]


# --------------------------------------------------------------------------------------------------
# R-side per-pair derivative closures (transform = TRUE), vectorised over rows x1[i,], x2[i,]
# R/covariance_function_derivatives.R
# --------------------------------------------------------------------------------------------------
def dsqexp_dsigma(x1, x2, cov_par):
    """R/covariance_function_derivatives.R:7-39."""
    sigma, l = cov_par["sigma"], cov_par["l"]
    r2 = np.sum((np.atleast_2d(x1) - np.atleast_2d(x2)) ** 2, axis=1)
    return {"derivative": 2 * sigma * np.exp(-(1 / (2 * l ** 2)) * r2) * sigma, "trans_par": math.log(sigma)}


def dsqexp_dsigma_ard(x1, x2, cov_par):
    """R/covariance_function_derivatives.R:41-78."""
    sigma = cov_par["sigma"]
    x1, x2 = np.atleast_2d(x1), np.atleast_2d(x2)
    l = np.array([cov_par["l%d" % (i + 1)] for i in range(x1.shape[1])])
    s = np.sum((x1 - x2) ** 2 / (l ** 2), axis=1)
    return {"derivative": 2 * sigma * np.exp(-(1 / 2) * s) * sigma, "trans_par": math.log(sigma)}


def dsqexp_dtau(x1, x2, cov_par):
    """R/covariance_function_derivatives.R:82-114."""
    tau = cov_par["tau"]
    eq = np.all(np.atleast_2d(x1) == np.atleast_2d(x2), axis=1)
    return {"derivative": 2 * tau * tau * 1 * eq.astype(np.float64), "trans_par": math.log(tau)}


def dsqexp_dl(x1, x2, cov_par):
    """R/covariance_function_derivatives.R:122-154."""
    sigma, l = cov_par["sigma"], cov_par["l"]
    r2 = np.sum((np.atleast_2d(x1) - np.atleast_2d(x2)) ** 2, axis=1)
    return {"derivative": (sigma ** 2 * np.exp((-1 / (2 * l ** 2)) * r2)) * ((1 / (l ** 3)) * r2) * l,
            "trans_par": math.log(l)}


def _knot_l(x2, cov_par, ard):
    if ard:
        return np.array([cov_par["l%d" % (i + 1)] for i in range(len(x2))])
    return cov_par["l"]


def _dx2(x1, x2, cov_par, transform, bounds, ard):
    sigma = cov_par["sigma"]
    x1 = np.asarray(x1, dtype=np.float64).reshape(-1)
    x2 = np.asarray(x2, dtype=np.float64).reshape(-1)
    l = _knot_l(x2, cov_par, ard)
    bounds = np.asarray(bounds, dtype=np.float64)
    dx2_dx2t = (bounds[:, 1] - bounds[:, 0]) / (((x2 - bounds[:, 0]) * (bounds[:, 1] - x2)) + 1e-4)
    if ard:
        e = np.exp(-1 / 2 * np.sum((x1 - x2) ** 2 / l ** 2))
    else:
        e = np.exp(-np.sum((x1 - x2) ** 2) / (2 * l ** 2))
    der = (1 / (l ** 2)) * (x1 - x2) * sigma ** 2 * e
    if transform:
        with np.errstate(invalid="ignore", divide="ignore"):      # a knot outside the bounds gives NaN, as in R
            tp = np.log((x2 - bounds[:, 0]) + 1e-4) - np.log((bounds[:, 1] - x2) + 1e-4)
        return {"derivative": der * dx2_dx2t, "trans_par": tp}
    return {"derivative": der, "trans_par": x2}


def dsqexp_dx2(x1, x2, cov_par, transform=False, bounds=None):
    """R/covariance_function_derivatives.R:178-236 (the R closure optimize_gp installs as dcov_fun_dknot, not the
    unused Rcpp dsqexp_dx2C): derivative of k(x1, x2) wrt each coordinate of x2, times d x2 / d x2t when
    transform = TRUE (quirk Q12: the 1e-4 guard sits in the Jacobian and in the logit, not in the map)."""
    return _dx2(x1, x2, cov_par, transform, bounds, ard=False)


def dsqexp_dx2_ard(x1, x2, cov_par, transform=False, bounds=None):
    """R/covariance_function_derivatives.R:238-320."""
    return _dx2(x1, x2, cov_par, transform, bounds, ard=True)


def dcov_fun_dknot_for(cov_fun):
    """R/optimize_gp.R:246,261."""
    return {"sqexp": dsqexp_dx2, "ard": dsqexp_dx2_ard}[cov_fun]


def knot_bounds_for(xy):
    """R/vi_functions.R:175-178, R/laplace_approx_gradient.R (same four lines in every gradient function)."""
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    lo, hi = xy.min(axis=0), xy.max(axis=0)
    diffs = hi - lo
    return np.stack([lo - diffs / 10, hi + diffs / 10], axis=1)


def _dsig12_dknot(k, d, cov_par, dcov_fun_dknot, xu, xy, bounds, transform):
    """R/vi_functions.R:427-445: n x m, only column k non-zero."""
    mat = np.zeros((xy.shape[0], xu.shape[0]))
    for i in range(xy.shape[0]):
        mat[i, k] = dcov_fun_dknot(xy[i], xu[k], cov_par, transform, bounds)["derivative"][d]
    return mat


def _dsig22_dknot(k, d, cov_par, dcov_fun_dknot, xu, bounds, transform):
    """R/vi_functions.R:446-474: column k, then row k, both filled with d k(xu[i], xu[k]) / d xu[k, d]."""
    m = xu.shape[0]
    mat = np.zeros((m, m))
    for i in range(m):
        mat[i, k] = dcov_fun_dknot(xu[i], xu[k], cov_par, transform, bounds)["derivative"][d]
    for i in range(m):
        mat[k, i] = dcov_fun_dknot(xu[i], xu[k], cov_par, transform, bounds)["derivative"][d]
    return mat


def _knot_gradient(vi, cov_par, dcov_fun_dknot, xu, xy, knot_opt, transform, B, C, Sigma12, Sigma22, FF, comp2_1):
    """Knot loop shared by delbo_dcov_par (R/vi_functions.R:475-581, vi = True: A = 0 and the trace-term
    derivative is added) and dlogp_dcov_par (R/laplace_approx_gradient.R:1011-1115, vi = False: A = -A2)."""
    m, dd = xu.shape
    n = xy.shape[0]
    bounds = knot_bounds_for(xy)
    grad_knot = np.zeros(m * dd)
    trans_knot = xu.copy()
    p = 0
    for k in range(m):
        if transform:
            trans_knot[k] = dcov_fun_dknot(0, xu[k], cov_par, transform, bounds)["trans_par"]
        for d in range(dd):
            p += 1
            if k not in knot_opt:
                continue
            dK = _dsig12_dknot(k, d, cov_par, dcov_fun_dknot, xu, xy, bounds, transform)
            dS = _dsig22_dknot(k, d, cov_par, dcov_fun_dknot, xu, bounds, transform)
            temp1 = 2 * dK - FF.T @ dS
            A2 = np.sum(temp1 * FF.T, axis=1)
            A = np.zeros(n) if vi else -A2
            comp1 = _comp1(A, B, C, Sigma12, Sigma22, FF, dK, dS)
            comp2 = _comp2(A, comp2_1, Sigma12, Sigma22, FF, dK, dS)
            g = (1 / 2) * comp2 - (1 / 2) * comp1
            if vi:
                g += dtrace_term_dcov_par(cov_par, 0 - A2)
            grad_knot[p - 1] = g
    return grad_knot, trans_knot


def dcov_fun_dtheta_for(cov_fun):
    """R/optimize_gp.R:236-261 (nugget = TRUE)."""
    if cov_fun == "sqexp":
        return {"sigma": dsqexp_dsigma, "l": dsqexp_dl, "tau": dsqexp_dtau}
    if cov_fun == "ard":
        return {"sigma": dsqexp_dsigma_ard, "tau": dsqexp_dtau}
    raise ValueError("Error: invalid covariance function")


def _trans_par(cov_par, dcov_fun_dtheta, lnames):
    """R/vi_functions.R:263-275: log of each parameter."""
    tp = {}
    for name, val in cov_par.items():
        if name in lnames:
            tp[name] = float(pos_to_real(val))
        else:
            tp[name] = dcov_fun_dtheta[name](np.zeros((1, 1)), np.zeros((1, 1)), cov_par)["trans_par"]
    return tp


# --------------------------------------------------------------------------------------------------
# shared assembly as the R drivers do it
# --------------------------------------------------------------------------------------------------
def assemble(cov_par, cov_fun, xy, xu, delta, keep_tau_in_S=False):
    """Sigma12 / Sigma22 as every R driver builds them.

    Gaussian models: Sigma22 = self-cov minus tau^2 I (R/vi_functions.R:733-747, :192-221;
    R/laplace_gradient_ascent.R:1241-1256; R/laplace_approx_gradient.R:782-813).
    Laplace models keep tau^2 on the diagonal (R/newtrap_sparseGP.R:43-60, R/laplace_approx_gradient.R:91-121).
    """
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    if cov_fun == "ard":
        lnames = lnames_for(xy)
        Sigma12 = make_cov_mat_ardC(xy, xu, cov_par, cov_fun, delta, lnames)
        Sigma22 = make_cov_mat_ardC(xu, None, cov_par, cov_fun, delta, lnames)
    else:
        lnames = []
        Sigma12 = make_cov_matC(xy, xu, cov_par, cov_fun, delta)
        Sigma22 = make_cov_matC(xu, None, cov_par, cov_fun, delta)
    if not keep_tau_in_S:
        Sigma22 = Sigma22 - cov_par["tau"] ** 2 * np.eye(xu.shape[0])
    return Sigma12, Sigma22, lnames


def fic_Z(cov_par, Sigma12, Sigma22, delta):
    """Z = sigma^2 + tau^2 + delta - diag(K S^-1 K^T): R/laplace_gradient_ascent.R:1259-1263,
    R/newtrap_sparseGP.R:62-66."""
    Z2 = solve(Sigma22, Sigma12.T)
    Z3 = Sigma12 * Z2.T
    Z4 = np.sum(Z3, axis=1)
    return cov_par["sigma"] ** 2 + cov_par["tau"] ** 2 + delta - Z4


# --------------------------------------------------------------------------------------------------
# trace term: R/vi_functions.R:14-60
# --------------------------------------------------------------------------------------------------
def trace_term_fun(cov_par, Sigma12, Sigma22, delta):
    """R/vi_functions.R:14-27."""
    tau, sigma = cov_par["tau"], cov_par["sigma"]
    Z2 = solve(Sigma22, Sigma12.T)
    Z3 = Sigma12 * Z2.T
    Z4 = np.sum(Z3, axis=1)
    Lambda = sigma ** 2 + delta - Z4
    return -(1 / (2 * tau ** 2)) * np.sum(Lambda)


def dtrace_term_dtau(cov_par, trace_term):
    """R/vi_functions.R:38-44."""
    return -2 * trace_term


def dtrace_term_dcov_par(cov_par, A_trace):
    """R/vi_functions.R:54-60."""
    tau = cov_par["tau"]
    return -(1 / (2 * tau ** 2)) * np.sum(A_trace)


# --------------------------------------------------------------------------------------------------
# Gaussian objectives: R/vi_functions.R:64-121, R/laplace_approx_obj_funs.R:6-52
# --------------------------------------------------------------------------------------------------
def _gauss_obj_core(mu, Z, Sigma12, Sigma22, y):
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    Z = np.broadcast_to(np.asarray(Z, dtype=np.float64).reshape(-1), y.shape)
    ZSig12 = _rows(1 / Z, Sigma12)
    R = chol(Sigma22 + Sigma12.T @ ZSig12)
    logdetR = 2 * np.sum(np.log(np.diag(R)))
    r = _col(y - mu)
    rhs = (r.T @ ZSig12).T
    quad_form_part = -(1 / 2) * (r.T @ (_col(1 / Z) * r)) + \
        (1 / 2) * (r.T @ ZSig12) @ solve(R, solve(R.T, rhs))
    with np.errstate(divide="ignore", invalid="ignore"):
        # log(det(Sigma22)): LU determinant, then log -- under/overflows to -Inf/Inf
        det_part = -(1 / 2) * (np.sum(np.log(Z)) - np.log(det(Sigma22)) + logdetR)
    return float(quad_form_part[0, 0]) + float(det_part) - (len(y) / 2) * math.log(2 * math.pi)


def elbo_fun(mu, Z, Sigma12, Sigma22, y, cov_par, delta, trace_term_fun=trace_term_fun):
    """R/vi_functions.R:64-121."""
    core = _gauss_obj_core(mu, Z, Sigma12, Sigma22, y)
    trace_term = trace_term_fun(cov_par=cov_par, Sigma12=Sigma12, Sigma22=Sigma22, delta=delta)
    return core + trace_term


def obj_fun_norm(mu, Z, Sigma12, Sigma22, y):
    """R/laplace_approx_obj_funs.R:6-52."""
    return _gauss_obj_core(mu, Z, Sigma12, Sigma22, y)


# --------------------------------------------------------------------------------------------------
# Gaussian gradients: R/vi_functions.R:126-420 (VI), R/laplace_approx_gradient.R:720-968 (FIC)
# --------------------------------------------------------------------------------------------------
def _dsig_pair(cov_fun, xy, xu, cov_par, par_name, lnames):
    if cov_fun == "ard":
        dK = dsig_dtheta_ardC(xy, xu, cov_par, cov_fun, par_name, lnames)
        dS = dsig_dtheta_ardC(xu, None, cov_par, cov_fun, par_name, lnames)
    else:
        dK = dsig_dthetaC(xy, xu, cov_par, cov_fun, par_name)
        dS = dsig_dthetaC(xu, None, cov_par, cov_fun, par_name)
    return dK, dS


def _comp1(A, B, C, Sigma12, Sigma22, FF, dSigma12, dSigma22):
    """comp1 block shared verbatim by delbo_dcov_par (R/vi_functions.R:377-391), dlogp_dcov_par
    (R/laplace_approx_gradient.R:937-951) and dlogq_dcov_par (:262-276)."""
    BK = _rows(B, Sigma12)
    comp1_1 = np.sum(A * B) - np.sum(np.diag(C @ Sigma12.T @ _rows(B * A * B, Sigma12)))
    comp1_2_1 = 2 * np.sum(np.diag(solve(Sigma22, Sigma12.T @ _rows(B, dSigma12))))
    comp1_2_2 = np.sum(np.diag(FF @ solve(Sigma22, BK.T).T @ dSigma22))
    comp1_2_3 = 2 * np.sum(np.diag((FF @ BK) @ (C @ Sigma12.T @ _rows(B, dSigma12))))
    comp1_2_4 = np.sum(np.diag((FF @ BK) @ ((C @ Sigma12.T) @ BK) @ solve(Sigma22, dSigma22)))
    return comp1_1 + comp1_2_1 - comp1_2_2 - comp1_2_3 + comp1_2_4


def _comp2(A, comp2_1, Sigma12, Sigma22, FF, dSigma12, dSigma22):
    """comp2 block: R/vi_functions.R:397-400, R/laplace_approx_gradient.R:957-960, :282-285."""
    s = solve(Sigma22, Sigma12.T @ comp2_1)
    comp2_2 = _col(A) * comp2_1 + 2 * dSigma12 @ s - FF.T @ dSigma22 @ s
    return float((comp2_1.T @ comp2_2)[0, 0])


def delbo_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, dcov_fun_dtheta=None, dcov_fun_dknot=None,
                   knot_opt=None, transform=True):
    """R/vi_functions.R:126-592, transform = TRUE for the covariance parameters.  dcov_fun_dknot = None stands for
    R's NA (no knot gradient).  Returns {"gradient": dict by parameter name, "trans_par": dict} plus, with a
    dcov_fun_dknot, "knot_gradient" (length m*d, knot-major like R's p counter) and "trans_knot" (m x d);
    knot_opt holds 0-based knot indices (R: 1-based)."""
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    if dcov_fun_dtheta is None:
        dcov_fun_dtheta = dcov_fun_dtheta_for(cov_fun)
    Sigma12, Sigma22, lnames = assemble(cov_par, cov_fun, xy, xu, delta)          # :186-221
    n = Sigma12.shape[0]
    FF = solve(Sigma22, Sigma12.T)                                                  # :227
    Z = np.repeat(cov_par["tau"] ** 2 + delta, n)                                   # :229
    B = 1 / Z
    R = chol(Sigma22 + Sigma12.T @ _rows(1 / Z, Sigma12))                           # :231
    C = solve(Sigma22 + Sigma12.T @ _rows(B, Sigma12))                              # :239
    ZK = _rows(1 / Z, Sigma12)
    comp2_1 = _col((1 / Z) * (y - mu)) - solve(R, solve(R.T, ZK.T)).T @ (ZK.T @ _col(y - mu))   # :245-246
    current_trace_term = trace_term_fun(cov_par, Sigma12, Sigma22, delta)           # :250-253
    grad = {}
    trans_par = _trans_par(cov_par, dcov_fun_dtheta, lnames)
    for par_name in cov_par:                                                        # :259
        dSigma12, dSigma22 = _dsig_pair(cov_fun, xy, xu, cov_par, par_name, lnames)  # :277-310
        if par_name == "tau":
            dSigma22 = np.zeros((xu.shape[0], xu.shape[0]))                         # :313-316
        if par_name != "tau":                                                       # :322-350
            A1_trace = np.zeros(n)
            if cov_fun != "ard" or par_name not in lnames:
                A1_trace = dcov_fun_dtheta[par_name](xy, xy, cov_par)["derivative"] * np.ones(n)
            temp1 = 2 * dSigma12 - FF.T @ dSigma22
            A2_trace = np.sum(temp1 * FF.T, axis=1)
            A_trace = A1_trace - A2_trace
        A1 = np.zeros(n)                                                            # :355-367
        if par_name == "tau":
            A1 = dcov_fun_dtheta[par_name](xy, xy, cov_par)["derivative"] * np.ones(n)
        A = A1
        comp1 = _comp1(A, B, C, Sigma12, Sigma22, FF, dSigma12, dSigma22)           # :377-391
        comp2 = _comp2(A, comp2_1, Sigma12, Sigma22, FF, dSigma12, dSigma22)        # :397-400
        if par_name == "tau":                                                       # :403-412
            dtrace_term = dtrace_term_dtau(cov_par, current_trace_term)
        else:
            dtrace_term = dtrace_term_dcov_par(cov_par, A_trace)
        grad[par_name] = (1 / 2) * comp2 - (1 / 2) * comp1 + dtrace_term            # :416-417
    if dcov_fun_dknot is None:
        return {"gradient": grad, "trans_par": trans_par}
    gk, tk = _knot_gradient(True, cov_par, dcov_fun_dknot, xu, xy, range(len(xu)) if knot_opt is None else knot_opt,
                            transform, B, C, Sigma12, Sigma22, FF, comp2_1)         # :425-581
    return {"gradient": grad, "knot_gradient": gk, "trans_par": trans_par, "trans_knot": tk}


def dlogp_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6, dcov_fun_dtheta=None, dcov_fun_dknot=None,
                   knot_opt=None, transform=True):
    """R/laplace_approx_gradient.R:720-1126 (FIC Gaussian), transform = TRUE; knot outputs as in delbo_dcov_par."""
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    if dcov_fun_dtheta is None:
        dcov_fun_dtheta = dcov_fun_dtheta_for(cov_fun)
    Sigma12, Sigma22, lnames = assemble(cov_par, cov_fun, xy, xu, delta)            # :780-813
    n = Sigma12.shape[0]
    FF = solve(Sigma22, Sigma12.T)                                                  # :819
    Z3 = Sigma12 * FF.T
    Z4 = np.sum(Z3, axis=1)
    Z = cov_par["sigma"] ** 2 + cov_par["tau"] ** 2 + delta - Z4                    # :822
    B = 1 / Z
    R = chol(Sigma22 + Sigma12.T @ _rows(1 / Z, Sigma12))                           # :825
    C = solve(Sigma22 + Sigma12.T @ _rows(B, Sigma12))                              # :832
    ZK = _rows(1 / Z, Sigma12)
    comp2_1 = _col((1 / Z) * (y - mu)) - solve(R, solve(R.T, ZK.T)).T @ (ZK.T @ _col(y - mu))   # :838-839
    grad = {}
    trans_par = _trans_par(cov_par, dcov_fun_dtheta, lnames)
    for par_name in cov_par:                                                        # :846
        dSigma12, dSigma22 = _dsig_pair(cov_fun, xy, xu, cov_par, par_name, lnames)  # :865-898
        if par_name == "tau":
            dSigma22 = np.zeros((xu.shape[0], xu.shape[0]))                         # :901-904
        A1 = np.zeros(n)                                                            # :908-920
        if cov_fun != "ard" or par_name not in lnames:
            A1 = dcov_fun_dtheta[par_name](xy, xy, cov_par)["derivative"] * np.ones(n)
        temp1 = 2 * dSigma12 - FF.T @ dSigma22                                      # :925
        A2 = np.sum(temp1 * FF.T, axis=1)                                           # :926-930
        A = A1 - A2
        comp1 = _comp1(A, B, C, Sigma12, Sigma22, FF, dSigma12, dSigma22)           # :937-951
        comp2 = _comp2(A, comp2_1, Sigma12, Sigma22, FF, dSigma12, dSigma22)        # :957-960
        grad[par_name] = (1 / 2) * comp2 - (1 / 2) * comp1                          # :964-965
    if dcov_fun_dknot is None:
        return {"gradient": grad, "trans_par": trans_par}
    gk, tk = _knot_gradient(False, cov_par, dcov_fun_dknot, xu, xy, range(len(xu)) if knot_opt is None else knot_opt,
                            transform, B, C, Sigma12, Sigma22, FF, comp2_1)         # :965-1115
    return {"gradient": grad, "knot_gradient": gk, "trans_par": trans_par, "trans_knot": tk}


# --------------------------------------------------------------------------------------------------
# Optimiser loops: norm_grad_ascent_vi (R/vi_functions.R:596-1218) and norm_grad_ascent
# (R/laplace_gradient_ascent.R:1111-1696), fixed knot count, transform = TRUE (hard-coded there, :640)
# --------------------------------------------------------------------------------------------------
FIT_DEFAULTS = {"optim_method": "adadelta", "decay": 0.95, "epsilon": 1e-6, "learn_rate": 1e-2, "eta": 1e3,
                "maxit": 1000, "obj_tol": 1e-3, "grad_tol": float("inf"), "delta": 1e-6}     # R/vi_functions.R:641-643


def _knot_trans_fun(t, bounds):
    """dsqexp_dx2(...)$trans_fun (R/covariance_function_derivatives.R:189-194), applied to one row."""
    return bounds[:, 1] * (1 / (1 + np.exp(-t))) + bounds[:, 0] * (1 / (1 + np.exp(t)))


def _grad_ascent_loop(evaluate, cov_par_start, xu, xy, o, opt_theta, opt_knots):
    """The optimiser skeleton shared verbatim by norm_grad_ascent_vi (R/vi_functions.R:753-1158), norm_grad_ascent
    (R/laplace_gradient_ascent.R:1267-1633) and laplace_grad_ascent (:100-568).  evaluate(cov_par, xu) returns
    (objective, gradient-function result)."""
    decay, eps, eta = o["decay"], o["epsilon"], o["eta"]
    names = list(cov_par_start)
    cov_par = dict(cov_par_start)
    obj, ev = evaluate(cov_par, xu)
    obj_vals = [obj]
    g_theta = np.array([ev["gradient"][k] for k in names]) if opt_theta else np.zeros(1)
    trans = np.array([ev["trans_par"][k] for k in names])
    g_knot = np.asarray(ev["knot_gradient"]) if opt_knots else np.zeros(1)
    if opt_knots:
        bounds = knot_bounds_for(xy)
        xu_trans = np.array(ev["trans_knot"])
    grad_hist, knot_hist, par_hist = [g_theta.copy()], [g_knot.copy()], [np.array([cov_par[k] for k in names])]
    sg2_t = sd2_t = sc_t = np.zeros(len(names))
    sg2_k = sd2_k = sc_k = np.zeros(xu.size)
    it = 1
    while it < o["maxit"] and (np.any(np.abs(np.concatenate([g_theta, g_knot])) > o["grad_tol"]) or
                               (abs(obj - obj_vals[it - 2]) > o["obj_tol"] if it > 1 else True)):
        it += 1
        if o["optim_method"] == "adadelta":
            if opt_theta:
                sg2_t = decay * sg2_t + (1 - decay) * g_theta ** 2
                d_t = ((1 / eta) ** sc_t) * (np.sqrt(sd2_t + eps) / np.sqrt(sg2_t + eps)) * g_theta
                sd2_t = decay * sd2_t + (1 - decay) * d_t ** 2
                trans = trans + d_t
            if opt_knots:
                sg2_k = decay * sg2_k + (1 - decay) * g_knot ** 2
                d_k = ((1 / eta) ** sc_k) * (np.sqrt(sd2_k + eps) / np.sqrt(sg2_k + eps)) * g_knot
                sd2_k = decay * sd2_k + (1 - decay) * d_k ** 2
                xu_trans = xu_trans + d_k.reshape(xu.shape)                     # matrix(..., byrow = TRUE)
        else:                                                                   # "ga"
            if opt_theta:
                trans = trans + o["learn_rate"] * g_theta
            if opt_knots:
                xu_trans = xu_trans + o["learn_rate"] * g_knot.reshape(xu.shape)
        if opt_knots:
            xu = np.vstack([_knot_trans_fun(xu_trans[k], bounds) for k in range(len(xu))])
        if opt_theta:
            cov_par = {k: float(np.exp(trans[j])) for j, k in enumerate(names)}  # real_to_pos / trans_fun = exp
        obj, ev = evaluate(cov_par, xu)
        obj_vals.append(obj)
        if opt_theta:
            g_new = np.array([ev["gradient"][k] for k in names])
            if o["optim_method"] == "adadelta":
                sc_t = decay * sc_t + (1 - decay) * np.abs(np.sign(g_new) - np.sign(g_theta)) / 2
            g_theta = g_new
            trans = np.array([ev["trans_par"][k] for k in names])
        if opt_knots:
            gk_new = np.asarray(ev["knot_gradient"])
            if o["optim_method"] == "adadelta":
                sc_k = decay * sc_k + (1 - decay) * np.abs(np.sign(gk_new) - np.sign(g_knot))   # no /2 (:1147)
            g_knot = gk_new
        grad_hist.append(g_theta.copy())
        knot_hist.append(g_knot.copy())
        par_hist.append(np.array([cov_par[k] for k in names]))
    return {"cov_par": cov_par, "xu": xu, "iter": it, "obj_fun": np.array(obj_vals), "grad": np.array(grad_hist),
            "knot_grad": np.array(knot_hist), "cov_par_history": np.array(par_hist)}


def norm_grad_ascent(cov_par_start, cov_fun, xu, xy, y, mu, opt=None, vi=True, opt_theta=True, opt_knots=False,
                     knot_opt=None):
    """Gradient ascent on the ELBO (vi = True: R/vi_functions.R:703-1158) or the FIC marginal likelihood
    (vi = False: R/laplace_gradient_ascent.R:1219-1633).  opt_theta stands for is.list(dcov_fun_dtheta), opt_knots
    for is.function(dcov_fun_dknot).  Returns cov_par, xu, iter, obj_fun, grad, knot_grad histories; the posterior at
    the knots (the functions' tail) is gauss_posterior_u."""
    o = dict(FIT_DEFAULTS)
    o.update({k: v for k, v in (opt or {}).items() if k in o})
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.array(xu, dtype=np.float64).reshape(len(xu), -1)
    delta = o["delta"]
    grad_fun = delbo_dcov_par if vi else dlogp_dcov_par
    dknot = dcov_fun_dknot_for(cov_fun) if opt_knots else None

    def evaluate(cp, knots):
        Sigma12, Sigma22, _ = assemble(cp, cov_fun, xy, knots, delta)
        if vi:
            Z = np.repeat(cp["tau"] ** 2 + delta, Sigma12.shape[0])
            obj = elbo_fun(mu, Z, Sigma12, Sigma22, y, cp, delta)
        else:
            obj = obj_fun_norm(mu, fic_Z(cp, Sigma12, Sigma22, delta), Sigma12, Sigma22, y)
        return obj, grad_fun(cp, cov_fun, knots, xy, y, mu, delta, dcov_fun_dknot=dknot, knot_opt=knot_opt)

    return _grad_ascent_loop(evaluate, cov_par_start, xu, xy, o, opt_theta, opt_knots)


LAPLACE_FIT_DEFAULTS = dict(FIT_DEFAULTS, maxit_nr=1000, tol_nr=1e-6)            # R/laplace_gradient_ascent.R:77-80


def laplace_grad_ascent(cov_par_start, cov_fun, xu, xy, y, ff, family, mu, muu, opt=None, opt_theta=True,
                        opt_knots=False, knot_opt=None, **kw):
    """R/laplace_gradient_ascent.R:10-628: every evaluation is a Newton mode search (newtrap_sparseGP, warm-started
    from the previous mode) followed by dlogq_dcov_par at that mode; the objective is the last value of the Newton
    history.  Adds fmax, nr_iter, u_mean, u_var of the last Newton run to the result."""
    o = dict(LAPLACE_FIT_DEFAULTS)
    o.update({k: v for k, v in (opt or {}).items() if k in o})
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.array(xu, dtype=np.float64).reshape(len(xu), -1)
    dknot = dcov_fun_dknot_for(cov_fun) if opt_knots else None
    state = {"ff": np.asarray(ff, dtype=np.float64).reshape(-1), "nr_iter": []}

    def evaluate(cp, knots):
        nr = newtrap_sparseGP(state["ff"], family, cp, cov_fun, xy, knots, y, mu, muu, maxit=o["maxit_nr"],
                              tol=o["tol_nr"], delta=o["delta"], **kw)
        state["ff"], state["nr"] = nr["gp"], nr
        state["nr_iter"].append(len(nr["objective_function_values"]))
        g = dlogq_dcov_par(cp, cov_fun, knots, xy, y, nr["gp"], family, mu, o["delta"], dcov_fun_dknot=dknot,
                           knot_opt=knot_opt, **kw)
        return nr["objective_function_values"][-1], g

    res = _grad_ascent_loop(evaluate, cov_par_start, xu, xy, o, opt_theta, opt_knots)
    res.update(fmax=state["ff"], nr_iter=np.array(state["nr_iter"]), u_mean=state["nr"]["u_posterior_mean"],
               u_var=state["nr"]["u_posterior_variance"])
    return res


# --------------------------------------------------------------------------------------------------
# OAT candidate scoring: knot_prop_random_norm_vi (R/vi_functions.R:2108-2304) and knot_prop_random_norm
# (R/knot_proposal_functions.R:1176-1357) without their RNG (the candidate rows are an argument here)
# --------------------------------------------------------------------------------------------------
def oat_candidate_scores(cov_par, cov_fun, xu, xy, y, mu, pseudo_prop, delta=1e-6, vi=True):
    """For each candidate row: the objective with that row appended to the knots, objective only
    (R/vi_functions.R:2211-2245 with elbo_fun; R/knot_proposal_functions.R:1283-1310 with obj_fun_norm and the FIC
    Z).  A failed chol() -- R's try-error, which makes the caller resample -- is reported as NaN."""
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    pseudo_prop = np.asarray(pseudo_prop, dtype=np.float64).reshape(len(pseudo_prop), -1)
    out = np.full(len(pseudo_prop), np.nan)
    for i in range(len(pseudo_prop)):
        pseudo_xu = np.vstack([xu, pseudo_prop[i]])
        Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, xy, pseudo_xu, delta)
        try:
            if vi:
                Z = np.repeat(cov_par["tau"] ** 2 + delta, Sigma12.shape[0])
                out[i] = elbo_fun(mu, Z, Sigma12, Sigma22, y, cov_par, delta)
            else:
                out[i] = obj_fun_norm(mu, fic_Z(cov_par, Sigma12, Sigma22, delta), Sigma12, Sigma22, y)
        except np.linalg.LinAlgError:
            pass
    return out


def laplace_oat_candidate_scores(cov_par, cov_fun, xu, xy, y, fmax, family, mu, muu, pseudo_prop, delta=1e-6,
                                 maxit=1000, tol=1e-6, **kw):
    """Candidate loop of knot_prop_random (R/knot_proposal_functions.R:1096-1120): newtrap_sparseGP warm-started from
    fmax with the candidate appended to the knots (muu extended by muu[1]); the score is the last objective value."""
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    pseudo_prop = np.asarray(pseudo_prop, dtype=np.float64).reshape(len(pseudo_prop), -1)
    muu = np.asarray(muu, dtype=np.float64).reshape(-1)
    out = np.full(len(pseudo_prop), np.nan)
    for i in range(len(pseudo_prop)):
        try:
            nr = newtrap_sparseGP(fmax, family, cov_par, cov_fun, xy, np.vstack([xu, pseudo_prop[i]]), y, mu,
                                  np.concatenate([muu, muu[:1]]), maxit=maxit, tol=tol, delta=delta, **kw)
            out[i] = nr["objective_function_values"][-1]
        except np.linalg.LinAlgError:
            pass
    return out


def knot_prop_choice(xu, pseudo_prop, obj_current, scores):
    """obj_fun_x[which.max(c(rep(obj_current, nrow(xu)), scores)), ] (R/vi_functions.R:2164,2299-2303): the best
    candidate, or the FIRST existing knot when no candidate beats the current objective."""
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    vals = np.concatenate([np.repeat(obj_current, len(xu)), scores])
    vals = np.where(np.isnan(vals), -np.inf, vals)               # which.max skips NA / NaN
    return np.vstack([xu, np.asarray(pseudo_prop).reshape(len(scores), -1)])[int(np.argmax(vals))].reshape(1, -1)


# --------------------------------------------------------------------------------------------------
# One "objective + gradient evaluation" exactly as one optimiser iteration performs it
# --------------------------------------------------------------------------------------------------
def vi_obj_grad(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6):
    """One iteration's evaluation in norm_grad_ascent_vi: R/vi_functions.R:1089-1128."""
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, xy, xu, delta)
    Z = np.repeat(cov_par["tau"] ** 2 + delta, Sigma12.shape[0])
    obj = elbo_fun(mu, Z, Sigma12, Sigma22, y, cov_par, delta)
    g = delbo_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta)
    return obj, g["gradient"]


def fic_obj_grad(cov_par, cov_fun, xu, xy, y, mu, delta=1e-6):
    """One iteration's evaluation in norm_grad_ascent: R/laplace_gradient_ascent.R:1241-1275."""
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, xy, xu, delta)
    Z = fic_Z(cov_par, Sigma12, Sigma22, delta)
    obj = obj_fun_norm(mu, Z, Sigma12, Sigma22, y)
    g = dlogp_dcov_par(cov_par, cov_fun, xu, xy, y, mu, delta)
    return obj, g["gradient"]


# --------------------------------------------------------------------------------------------------
# Likelihood pieces: R/derivative_functions_of_data_likelihoods.R, R/laplace_approx_obj_funs.R:178-184
# --------------------------------------------------------------------------------------------------
def my_logistic(x):
    """R/laplace_approx_obj_funs.R:178-184."""
    return 1 / (1 + np.exp(-np.asarray(x, dtype=np.float64)))


def d2log_py_dff_pois(ff, y=None, m=1.0):
    """R/derivative_functions_of_data_likelihoods.R:7-12."""
    return -m * np.exp(ff)


def d3log_py_dff_pois(ff, y=None, m=1.0):
    """:16-21."""
    return -m * np.exp(ff)


def dlog_py_dff_pois(ff, y, m=1.0):
    """:25-30."""
    return -m * np.exp(ff) + y


def d2log_py_dff_bern(ff, y, **_):
    """:90-112 (quirk Q1 reproduced verbatim)."""
    pi_ff = my_logistic(ff)
    return (1 - 2 * pi_ff) * (y - pi_ff) - (y * (1 - pi_ff) ** 2 + pi_ff ** 2 + y * pi_ff ** 2)


def d3log_py_dff_bern(ff, y, **_):
    """:116-145."""
    pi_ff = my_logistic(ff)
    dpi_dff = pi_ff * (1 - pi_ff)
    return -2 * dpi_dff * (y - pi_ff) - dpi_dff * (1 - 2 * pi_ff) - \
        (2 * y * (1 - pi_ff) * (-dpi_dff) + 2 * pi_ff * dpi_dff + 2 * y * pi_ff * dpi_dff)


def dlog_py_dff_bern(ff, y, **_):
    """:165-184."""
    pi_ff = my_logistic(ff)
    return y * (1 - pi_ff) - pi_ff + y * pi_ff


def _grad_loglik(d1, ff, mu, Sigma12, Sigma22, Z):
    R = chol(Sigma22 + Sigma12.T @ _rows(1 / Z, Sigma12))
    d2 = -1 / Z * (ff - mu) + (_rows(1 / Z, Sigma12) @ solve(R, solve(R.T, Sigma12.T @ _col(1 / Z * (ff - mu))))).reshape(-1)
    return d1 + d2


def grad_loglik_fn_pois(ff, y, mu, Sigma12, Sigma22, Z, m=1.0):
    """:34-61."""
    return _grad_loglik(-m * np.exp(ff) + y, ff, mu, Sigma12, Sigma22, Z)


def grad_loglik_fn_bern(ff, y, mu, Sigma12, Sigma22, Z, **_):
    """:188-235."""
    pi_ff = my_logistic(ff)
    d1 = y * (1 - pi_ff) - pi_ff + y * pi_ff
    return _grad_loglik(d1, ff, mu, Sigma12, Sigma22, Z)


FAMILIES = {
    "bernoulli": dict(d1=dlog_py_dff_bern, d2=d2log_py_dff_bern, d3=d3log_py_dff_bern, grad=grad_loglik_fn_bern),
    "poisson": dict(d1=dlog_py_dff_pois, d2=d2log_py_dff_pois, d3=d3log_py_dff_pois, grad=grad_loglik_fn_pois),
}


# --------------------------------------------------------------------------------------------------
# Sparse Laplace objectives: R/laplace_approx_obj_funs.R:108-174 (Poisson), :189-341 (Bernoulli)
# --------------------------------------------------------------------------------------------------
def _laplace_obj(ff, mu, Z, Sigma12, Sigma22, W, log_py):
    Z2 = 1 + np.sqrt(-W) * Z * np.sqrt(-W)
    ZSig12 = _rows(1 / Z, Sigma12)
    R = chol(Sigma22 + Sigma12.T @ ZSig12)
    sK = _rows(np.sqrt(-W), Sigma12)
    R2 = chol(Sigma22 + sK.T @ _rows(1 / Z2, sK))
    logdetR2 = 2 * np.sum(np.log(np.diag(R2)))
    R_Sigma22 = chol(Sigma22)
    r = _col(ff - mu)
    t = solve(R.T, (r.T @ ZSig12).T)
    quad_form_part = -(1 / 2) * (r.T @ (_col(1 / Z) * r)) + (1 / 2) * (t.T @ t)
    det_part_1 = -(1 / 2) * (-2 * np.sum(np.log(np.diag(R_Sigma22))) + logdetR2)
    det_part_2 = -(1 / 2) * np.sum(np.log(Z2))
    return float(quad_form_part[0, 0]) + log_py + det_part_1 + det_part_2


def obj_fun_pois(ff, mu, Z, Sigma12, Sigma22, y, m=1.0):
    """R/laplace_approx_obj_funs.R:108-174."""
    ff = np.asarray(ff, dtype=np.float64).reshape(-1)
    W = -m * np.exp(ff)
    log_py = float(np.sum(y * np.log(m) - gammaln(np.asarray(y) + 1) - m * np.exp(ff) + y * ff))
    return _laplace_obj(ff, mu, Z, Sigma12, Sigma22, W, log_py)


def obj_fun_bern(ff, mu, Z, Sigma12, Sigma22, y, **_):
    """R/laplace_approx_obj_funs.R:189-341."""
    ff = np.asarray(ff, dtype=np.float64).reshape(-1)
    pi_ff = my_logistic(ff)
    log_pi_ff = -np.logaddexp(0.0, -ff)              # plogis(q = ff, log.p = TRUE)
    log_1minus_pi_ff = -np.logaddexp(0.0, ff)        # plogis(q = -ff, log.p = TRUE)
    W = (1 - 2 * pi_ff) * (y - pi_ff) - (y * (1 - pi_ff) ** 2 + pi_ff ** 2 + y * pi_ff ** 2)
    log_py = float(np.sum(y * log_pi_ff + (1 - y) * log_1minus_pi_ff))
    return _laplace_obj(ff, mu, Z, Sigma12, Sigma22, W, log_py)


OBJ_FUNS = {"bernoulli": obj_fun_bern, "poisson": obj_fun_pois}


# --------------------------------------------------------------------------------------------------
# Newton-Raphson mode finder: R/newtrap_sparseGP.R
# --------------------------------------------------------------------------------------------------
def newtrap_sparseGP_update(ff, W, Z, Sigma12, Sigma22, grad_psi, dlog_py_dff, y, mu, **kw):
    """R/newtrap_sparseGP.R:234-325."""
    ZSig12 = _rows(1 / Z, Sigma12)
    R = chol(Sigma22 + Sigma12.T @ ZSig12)                                          # :250
    R3 = chol(Sigma22 + Sigma12.T @ _rows((Z - 1 / W) ** (-1), Sigma12))            # :251
    e = 1 / (1 - Z * W)
    A11 = (Z / (1 - Z * W)) * dlog_py_dff(ff, y, **kw)                              # :277
    A12 = (1 - Z * W) ** (-1) * (ff - mu)                                           # :278
    A13 = (solve(R.T, _rows(e, Sigma12).T).T @ solve(R.T, ZSig12.T @ _col(ff - mu))).reshape(-1)   # :279-280
    A2 = (solve(R3.T, _rows(e, Sigma12).T).T @ solve(R3.T, Sigma12.T @ _col(e * grad_psi))).reshape(-1)  # :282-283
    return ff + (A11 - A12 + A13 + A2)                                              # :288, :322


def newtrap_sparseGP(start_vals, family, cov_par, cov_fun, xy, xu, y, mu, muu, maxit=1000, tol=1e-6,
                     delta=1e-6, **kw):
    """R/newtrap_sparseGP.R:6-186."""
    fam, obj_fun = FAMILIES[family], OBJ_FUNS[family]
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, xy, xu, delta, keep_tau_in_S=True)   # :43-60
    Z = fic_Z(cov_par, Sigma12, Sigma22, delta)                                     # :62-66
    ff = np.asarray(start_vals, dtype=np.float64).reshape(-1).copy()
    obj_fun_vals = [obj_fun(ff, mu, Z, Sigma12, Sigma22, y, **kw)]                  # :75
    it = 1
    while True:
        it += 1
        W = fam["d2"](ff, y=y, **kw)                                                # :84, :105
        grad_psi = fam["grad"](ff, y, mu, Sigma12, Sigma22, Z, **kw)
        ff = newtrap_sparseGP_update(ff, W, Z, Sigma12, Sigma22, grad_psi, fam["d1"], y, mu, **kw)
        obj_fun_vals.append(obj_fun(ff, mu, Z, Sigma12, Sigma22, y, **kw))
        if not (it < maxit and (abs(obj_fun_vals[-1] - obj_fun_vals[-2]) > tol or np.any(np.abs(grad_psi) > tol))):
            break                                                                   # :100
    ZSig12 = _rows(1 / Z, Sigma12)                                                  # :156
    WmZ_inv = 1 / ((1 / W) - Z)                                                     # :159
    TT = Sigma12.T @ _rows(WmZ_inv, Sigma12)                                        # :162
    R = chol(Sigma22 + Sigma12.T @ ZSig12)                                          # :166
    u_mean = np.asarray(muu, dtype=np.float64).reshape(-1) + (ZSig12.T @ _col(ff - mu)).reshape(-1) - \
        (Sigma12.T @ (ZSig12 @ solve(R, solve(R.T, ZSig12.T @ _col(ff - mu))))).reshape(-1)   # :171-173
    u_var = Sigma22 + TT + TT @ solve(Sigma22 - TT, TT)                             # :176
    return {"gp": ff, "objective_function_values": np.array(obj_fun_vals), "gradient": grad_psi,
            "u_posterior_mean": u_mean, "u_posterior_variance": u_var, "W": W, "Z": Z}


# --------------------------------------------------------------------------------------------------
# Sparse Laplace gradient: R/laplace_approx_gradient.R:25-339
# --------------------------------------------------------------------------------------------------
def dlogq_dcov_par(cov_par, cov_fun, xu, xy, y, ff, family, mu, delta=1e-6, dcov_fun_dtheta=None, dcov_fun_dknot=None,
                   knot_opt=None, transform=True, **kw):
    """R/laplace_approx_gradient.R:25-715, transform = TRUE for the covariance parameters (quirk Q2 verbatim);
    dcov_fun_dknot / knot_opt / knot outputs as in delbo_dcov_par (:345-705)."""
    fam = FAMILIES[family]
    xy = np.asarray(xy, dtype=np.float64).reshape(len(xy), -1)
    xu = np.asarray(xu, dtype=np.float64).reshape(len(xu), -1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    ff = np.asarray(ff, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    if dcov_fun_dtheta is None:
        dcov_fun_dtheta = dcov_fun_dtheta_for(cov_fun)
    Sigma12, Sigma22, lnames = assemble(cov_par, cov_fun, xy, xu, delta, keep_tau_in_S=True)   # :91-121
    n = Sigma12.shape[0]
    FF = solve(Sigma22, Sigma12.T)                                                  # :127
    Z = cov_par["sigma"] ** 2 + cov_par["tau"] ** 2 + delta - np.sum(Sigma12 * FF.T, axis=1)   # :128-130
    W = fam["d2"](ff, y=y, **kw)                                                    # :132
    B = 1 / (Z - (1 / W))                                                           # :133
    R = chol(Sigma22 + Sigma12.T @ _rows(1 / Z, Sigma12))                           # :134
    W3 = fam["d3"](ff, y=y, **kw)                                                   # :136
    C = solve(Sigma22 + Sigma12.T @ _rows(B, Sigma12))                              # :142
    ZK = _rows(1 / Z, Sigma12)
    comp2_1 = _col((1 / Z) * (ff - mu)) - solve(R, solve(R.T, ZK.T)).T @ (ZK.T @ _col(ff - mu))   # :148-149
    grad_log_py_ff = fam["d1"](ff, y, **kw)                                         # :153
    GG = solve(Sigma22, Sigma12.T @ _col(grad_log_py_ff))                           # :155
    D = W - 1 / Z                                                                   # :161
    E2 = np.eye(Sigma22.shape[0]) + solve(R.T, ZK.T) @ solve(R.T, _rows((1 / D) * (1 / Z), Sigma12).T).T   # :163
    RE2 = chol(E2)
    RE = RE2 @ R                                                                    # :165
    REinv = solve(RE)                                                               # :170
    temp = REinv.T @ (_rows((1 / Z) * (1 / D), Sigma12)).T                          # :171-177 (all i at once)
    comp4_1 = np.sum(temp * temp, axis=0)
    comp4 = -(1 / D) + comp4_1                                                      # :178
    grad = {}
    trans_par = _trans_par(cov_par, dcov_fun_dtheta, lnames)
    for par_name in cov_par:                                                        # :185
        dSigma12, dSigma22 = _dsig_pair(cov_fun, xy, xu, cov_par, par_name, lnames)  # :204-237 (dS for tau kept)
        A1 = np.zeros(n)                                                            # :243-255
        if cov_fun != "ard" or par_name not in lnames:
            A1 = dcov_fun_dtheta[par_name](xy, xy, cov_par)["derivative"] * np.ones(n)
        temp1 = 2 * dSigma12 - FF.T @ dSigma22                                      # :258
        A2 = np.sum(temp1 * FF.T, axis=1)
        A = A1 - A2                                                                 # :264
        comp1 = _comp1(A, B, C, Sigma12, Sigma22, FF, dSigma12, dSigma22)           # :270-284
        comp2 = _comp2(A, comp2_1, Sigma12, Sigma22, FF, dSigma12, dSigma22)        # :290-293
        comp3_1 = _col(A * grad_log_py_ff) + 2 * dSigma12 @ GG - FF.T @ dSigma22 @ GG   # :301-303
        BK = _rows(B, Sigma12)
        comp3 = _col(-(1 / W) * B) * comp3_1 + _col(1 / W) * (BK @ (C @ (Sigma12.T @ (_col(B) * comp3_1))))   # :306-307
        grad[par_name] = (1 / 2) * comp2 - (1 / 2) * comp1 - \
            (1 / 2) * float((_col(comp4 * (-W3)).T @ comp3)[0, 0])                  # :333-335
    if dcov_fun_dknot is None:
        return {"gradient": grad, "trans_par": trans_par}
    m, dd = xu.shape                                                                # :345-705
    bounds = knot_bounds_for(xy)
    grad_knot = np.zeros(m * dd)
    trans_knot = xu.copy()
    BK = _rows(B, Sigma12)
    sel = range(m) if knot_opt is None else knot_opt
    p = 0
    for k in range(m):
        if transform:
            trans_knot[k] = dcov_fun_dknot(0, xu[k], cov_par, transform, bounds)["trans_par"]
        for dk in range(dd):
            p += 1
            if k not in sel:
                continue
            dK = _dsig12_dknot(k, dk, cov_par, dcov_fun_dknot, xu, xy, bounds, transform)
            dS = _dsig22_dknot(k, dk, cov_par, dcov_fun_dknot, xu, bounds, transform)
            A = 0 - np.sum((2 * dK - FF.T @ dS) * FF.T, axis=1)                     # A1 = 0
            comp1 = _comp1(A, B, C, Sigma12, Sigma22, FF, dK, dS)
            comp2 = _comp2(A, comp2_1, Sigma12, Sigma22, FF, dK, dS)
            comp3_1 = _col(A * grad_log_py_ff) + 2 * dK @ GG - FF.T @ dS @ GG
            comp3 = _col(-(1 / W) * B) * comp3_1 + _col(1 / W) * (BK @ (C @ (Sigma12.T @ (_col(B) * comp3_1))))
            grad_knot[p - 1] = (1 / 2) * comp2 - (1 / 2) * comp1 - (1 / 2) * float((_col(comp4 * (-W3)).T @ comp3)[0, 0])
    return {"gradient": grad, "knot_gradient": grad_knot, "trans_par": trans_par, "trans_knot": trans_knot}


# --------------------------------------------------------------------------------------------------
# Posterior at the knots and prediction
# --------------------------------------------------------------------------------------------------
def gauss_posterior_u(cov_par, cov_fun, xu, xy, y, mu, muu, delta=1e-6, vi=True):
    """Tail of norm_grad_ascent_vi (R/vi_functions.R:1160-1180) for vi = True, of norm_grad_ascent
    (R/laplace_gradient_ascent.R:1637-1656) otherwise: posterior mean and variance of the process at the knots."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), y.shape)
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, xy, xu, delta)
    if vi:
        Z = np.repeat(cov_par["tau"] ** 2 + delta, Sigma12.shape[0])            # R/vi_functions.R:753
    else:
        Z = fic_Z(cov_par, Sigma12, Sigma22, delta)                               # R/laplace_gradient_ascent.R:1259-1263
    ZSig12 = _rows(1 / Z, Sigma12)
    R1 = chol(Sigma22 + Sigma12.T @ ZSig12)
    rhs = ZSig12.T @ _col(y - mu)
    u_mean = np.asarray(muu, dtype=np.float64).reshape(-1) + rhs.reshape(-1) - \
        (Sigma12.T @ (ZSig12 @ solve(R1, solve(R1.T, rhs)))).reshape(-1)
    W = solve(R1.T, Sigma12.T)
    u_var = Sigma22 - Sigma12.T @ ZSig12 + (ZSig12.T @ W.T) @ (W @ ZSig12)
    return u_mean, u_var


def predict_vi(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, delta=1e-6):
    """R/vi_functions.R:1222-1336, family = "gaussian", full_cov = FALSE."""
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, x_pred, xu, delta)           # Sigma22 = self-cov - tau^2 I
    Sigma22_inv = solve(Sigma22)
    pred_mean = np.asarray(mu, dtype=np.float64).reshape(-1) + \
        (Sigma12 @ solve(Sigma22, _col(np.asarray(u_mean) - np.asarray(muu)))).reshape(-1)
    temp22 = -Sigma22_inv + Sigma22_inv @ u_var @ Sigma22_inv
    Sigma11 = cov_par["tau"] ** 2 + cov_par["sigma"] ** 2 + delta
    pred_var = Sigma11 + np.sum(Sigma12 * (temp22 @ Sigma12.T).T, axis=1)
    return pred_mean, pred_var


def predict_laplace(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, family="gaussian", delta=1e-6):
    """R/laplace_approx_prediction.R:3-123, full_cov = FALSE: Sigma22 keeps tau^2 for non-Gaussian families and
    the predictive variance constant is sigma^2 + tau^2 (no delta)."""
    Sigma12, Sigma22, _ = assemble(cov_par, cov_fun, x_pred, xu, delta, keep_tau_in_S=(family != "gaussian"))
    Sigma22_inv = solve(Sigma22)
    pred_mean = np.asarray(mu, dtype=np.float64).reshape(-1) + \
        (Sigma12 @ solve(Sigma22, _col(np.asarray(u_mean) - np.asarray(muu)))).reshape(-1)
    temp22 = -Sigma22_inv + Sigma22_inv @ u_var @ Sigma22_inv
    pred_var = cov_par["sigma"] ** 2 + cov_par["tau"] ** 2 + np.sum((Sigma12 @ temp22) * Sigma12, axis=1)
    return pred_mean, pred_var
