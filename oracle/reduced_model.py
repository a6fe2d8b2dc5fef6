"""Row-separable ("reduced") form of the sparse-GP objective + gradient -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy statement of the two-pass algorithm the CUDA product implements: every n-dependent
quantity is a sum over data rows of row-local terms once the replicated m x m factors are known, so the rows can
be split into shards whose partial sums are added (the allreduce).  tests/ use it to
  * prove the reduced algebra equal to the literal transcription of the reference (oracle/ref_model.py),
  * exercise the shard -> allreduce -> replicate logic under gloo with world_size 2 on the CPU,
  * localise a parity failure of the CUDA path to one pass / one m x m stage.
It follows SURVEY.md Appendix B.2/B.3, which restate R/vi_functions.R:64-121,126-420 and
R/laplace_approx_gradient.R:720-968; formulas are derived in DESIGN.md.

Parity: held to oracle/ref_model.py, which is pinned to the reference's own R sources (see its header).
"""
from __future__ import annotations

import math

import contextlib

import numpy as np
import scipy.linalg as sla

# Working precision.  `with extended_precision():` switches every array of this module to numpy.longdouble
# (x87 80-bit here) with hand-written Cholesky / triangular inverse, giving a ~1e-19 reference for configurations
# whose conditioning puts the literal (LU, float64) transcription itself at ~1e-8.
_DT = np.float64


@contextlib.contextmanager
def extended_precision():
    global _DT
    old, _DT = _DT, np.longdouble
    try:
        yield
    finally:
        _DT = old


def _arr(a):
    return np.asarray(a, dtype=_DT)


def _cholesky(A):
    if _DT is np.float64:
        return np.linalg.cholesky(A)
    A = np.array(A, dtype=_DT)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        v = A[j:, j] - L[j:, :j] @ L[j, :j]
        L[j:, j] = v / np.sqrt(v[0])
    return L


def _spd_inverse(A):
    """(A^-1, log det A) through the Cholesky factor."""
    L = _cholesky(A)
    n = L.shape[0]
    logdet = 2 * np.sum(np.log(np.diag(L)))
    if _DT is np.float64:
        return sla.cho_solve((L, True), np.eye(n)), logdet
    X = np.zeros_like(L)                      # X = L^-1 by forward substitution, row by row
    for i in range(n):
        X[i, i] = 1 / L[i, i]
        X[i, :i] = -(L[i, :i] @ X[:i, :i]) / L[i, i]
    return X.T @ X, logdet


def _inv(A):
    if _DT is np.float64:
        return np.linalg.inv(A)
    return _spd_inverse(A)[0]


def _solve(A, b):
    if _DT is np.float64:
        return np.linalg.solve(A, b)
    return _spd_inverse(A)[0] @ b


def kernel_matrix(x, u, sigma, l, cov_fun="ard"):
    """K_ij and the per-dimension scaled squared differences D_ijc = ((x_ic - u_jc)/l_c)^2."""
    x = _arr(x).reshape(len(x), -1)
    u = _arr(u).reshape(len(u), -1)
    sigma = _DT(sigma)
    l = np.broadcast_to(_arr(l).reshape(-1), (x.shape[1],))
    D = ((x[:, None, :] - u[None, :, :]) / l[None, None, :]) ** 2
    K = sigma ** 2 * np.exp(-D.sum(axis=2) / 2)
    return K, D


def theta_names(cov_fun, d):
    return ["sigma"] + (["l%d" % (c + 1) for c in range(d)] if cov_fun == "ard" else ["l"]) + ["tau"]


def _coincident(x, u):
    return np.all(x[:, None, :] == u[None, :, :], axis=2)


# ---------------------------------------------------------------- pass 1 (rows) -----------------
def vi_pass1(x, r, u, sigma, l):
    """Per-shard partial sums: G1 = K^T K, b1 = K^T r, s0 = r^T r."""
    K, _ = kernel_matrix(x, u, sigma, l)
    return {"G1": K.T @ K, "b1": K.T @ r, "s0": r @ r, "n": _DT(len(r))}


def add_partials(parts):
    out = {}
    for k in parts[0]:
        out[k] = sum(p[k] for p in parts)
    return out


# ---------------------------------------------------------------- replicated m x m --------------
def vi_mid(p1, u, sigma, l, tau, delta, cov_fun="ard"):
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    m = u.shape[0]
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    sigma, tau, delta = _DT(sigma), _DT(tau), _DT(delta)
    S = Kuu + delta * np.eye(m, dtype=_DT)           # self-cov minus tau^2 I  (R/vi_functions.R:736-741)
    Z = tau ** 2 + delta
    B = 1 / Z
    n = p1["n"]
    Sinv, logdetS = _spd_inverse(S)
    G = B * p1["G1"]
    Cm, logdetA = _spd_inverse(S + G)
    b = B * p1["b1"]
    v = Cm @ b
    beta = Sinv @ (b - G @ v)
    sum_q = np.sum(Sinv * p1["G1"])
    tt = -(1 / (2 * tau ** 2)) * (n * (sigma ** 2 + delta) - sum_q)
    obj = -B * p1["s0"] / 2 + (b @ v) / 2 - (n * np.log(Z) - logdetS + logdetA) / 2 \
        - (n / 2) * np.log(2 * _DT(np.pi) if _DT is np.float64 else 2 * np.arctan(_DT(1)) * 4) + tt
    CGS = Cm @ G @ Sinv
    M = (1 / tau ** 2 - B) * Sinv + B * CGS
    SGS = Sinv @ G @ Sinv
    N = SGS / 2 - Sinv @ G @ CGS / 2 - np.outer(beta, beta) / 2 - (1 / (2 * tau ** 2)) * Sinv @ p1["G1"] @ Sinv
    return dict(obj=obj, tt=tt, B=B, v=v, beta=beta, M=M, N=N, Sinv=Sinv, C=Cm, trCG1=np.sum(Cm * p1["G1"]), n=n)


# ---------------------------------------------------------------- pass 2 (rows) -----------------
def vi_pass2(x, r, u, sigma, l, tau, mid):
    """Per-shard partials: sum_ij Omega_ij dK_ij(theta) for sigma and each length scale, sum alpha^2,
    and the coincident-row correction of quirk Q4 (dK/dlog tau = 2 tau^2 where x_i == u_j bit-exactly)."""
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    K, D = kernel_matrix(x, u, sigma, l)
    alpha = mid["B"] * (r - K @ mid["v"])
    Omega = K @ mid["M"] + np.outer(alpha, mid["beta"])
    P = Omega * K
    tau = _DT(tau)
    g_sigma = 2 * P.sum()
    g_l = np.array([np.sum(P * D[:, :, c]) for c in range(D.shape[2])], dtype=_DT)
    eq = _coincident(x, u)
    g_tau_q4 = _DT(0)
    if eq.any():
        Om_tau = Omega - (1 / tau ** 2) * (K @ mid["Sinv"])
        g_tau_q4 = 2 * tau ** 2 * Om_tau[eq].sum()
    return {"g_sigma": g_sigma, "g_l": g_l, "sum_alpha2": alpha @ alpha, "g_tau_q4": g_tau_q4,
            "g_knot": knot_colsums(P, x, u, l)}


def knot_colsums(P, x, u, l):
    """Row-shard partial of the knot gradient: G[k, c] = sum_i P_ik (x_ic - u_kc) / l_c^2 with P = Omega o K, i.e.
    sum_i Omega_ik dK_ik / du_kc before the Jacobian of the knot transform."""
    l = np.broadcast_to(np.asarray(l, dtype=_DT).reshape(-1), (u.shape[1],))
    return (P.T @ x - P.sum(axis=0)[:, None] * u) / l ** 2


def knot_bounds(x):
    """[min - range/10, max + range/10] per dimension (R/vi_functions.R:175-178)."""
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    lo, hi = x.min(axis=0), x.max(axis=0)
    return np.stack([lo - (hi - lo) / 10, hi + (hi - lo) / 10], axis=1)


def knot_finish(g_knot, N, u, sigma, l, bounds, transform=True):
    """Adds the m x m part, sum_j (N_jk + N_kj) Kuu_jk (u_jc - u_kc) / l_c^2 (dSigma22/du_kc has row k and column k
    filled with the same vector, zero on the diagonal: R/vi_functions.R:446-474), then the Jacobian
    (ub - lb) / ((u - lb)(ub - u) + 1e-4) of the bounded-logit knot transform (quirk Q12)."""
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    lv = np.broadcast_to(np.asarray(l, dtype=_DT).reshape(-1), (u.shape[1],))
    Q = (N + N.T) * Kuu
    g = g_knot + (Q.T @ u - Q.sum(axis=0)[:, None] * u) / lv ** 2
    if transform:
        b = np.asarray(bounds, dtype=_DT)
        g = g * ((b[:, 1] - b[:, 0]) / ((u - b[:, 0]) * (b[:, 1] - u) + _DT(1e-4)))
    return g


def dS_dtheta(u, sigma, l, tau, name, cov_fun="ard"):
    """Self-covariance derivative wrt log theta as dsig_dtheta*C builds it; tau forced to 0 for Gaussian
    models by the caller (R/vi_functions.R:313-316)."""
    Kuu, D = kernel_matrix(u, u, sigma, l)
    if name == "sigma":
        return 2 * Kuu
    if name == "l":
        return Kuu * D.sum(axis=2)
    if name.startswith("l"):
        return Kuu * D[:, :, int(name[1:]) - 1]
    raise KeyError(name)


def vi_finish(p2, mid, u, sigma, l, tau, cov_fun="ard"):
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    sigma, tau = _DT(sigma), _DT(tau)
    d = u.shape[1]
    n, B, N = mid["n"], mid["B"], mid["N"]
    grad = {}
    grad["sigma"] = p2["g_sigma"] + np.sum(N * dS_dtheta(u, sigma, l, tau, "sigma")) \
        - (1 / (2 * tau ** 2)) * 2 * sigma ** 2 * n
    if cov_fun == "ard":
        for c in range(d):
            nm = "l%d" % (c + 1)
            grad[nm] = p2["g_l"][c] + np.sum(N * dS_dtheta(u, sigma, l, tau, nm))
    else:
        grad["l"] = np.sum(p2["g_l"]) + np.sum(N * dS_dtheta(u, sigma, l, tau, "l", cov_fun))
    grad["tau"] = tau ** 2 * p2["sum_alpha2"] - tau ** 2 * (n * B - B ** 2 * mid["trCG1"]) - 2 * mid["tt"] \
        + p2["g_tau_q4"]
    return grad


def vi_obj_grad(x, y, mu, u, sigma, l, tau, delta, cov_fun="ard", shards=1, knots=False, bounds=None):
    """Full evaluation; `shards` splits the rows into that many contiguous blocks (the multi-GPU layout).
    knots = True appends the m x d knot gradient (bounds default to knot_bounds(x))."""
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    r = np.asarray(y, dtype=_DT).reshape(-1) - np.broadcast_to(np.asarray(mu, dtype=_DT).reshape(-1), (len(x),))
    kb, bounds = bounds, shard_bounds(len(x), shards)
    p1 = add_partials([vi_pass1(x[a:b], r[a:b], u, sigma, l) for a, b in bounds])
    mid = vi_mid(p1, u, sigma, l, tau, delta, cov_fun)
    p2 = add_partials([vi_pass2(x[a:b], r[a:b], u, sigma, l, tau, mid) for a, b in bounds])
    grad = vi_finish(p2, mid, u, sigma, l, tau, cov_fun)
    if not knots:
        return mid["obj"], grad
    return mid["obj"], grad, knot_finish(p2["g_knot"], mid["N"], u, sigma, l, knot_bounds(x) if kb is None else kb)


def vi_oat_scores(x, y, mu, u, cand, sigma, l, tau, delta, shards=1):
    """OAT candidate scoring for the VI objective by bordered updates: ONE Gram of [knots | candidates] over the
    data rows (sharded like pass 1), then per candidate O(m^2) Schur-complement algebra instead of a full
    (m+1)-knot evaluation.  Returns (objective with the m knots, objective with each candidate appended)."""
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    cand = np.asarray(cand, dtype=_DT).reshape(len(cand), -1)
    m, T = len(u), len(cand)
    ua = np.vstack([u, cand])
    r = np.asarray(y, dtype=_DT).reshape(-1) - np.broadcast_to(np.asarray(mu, dtype=_DT).reshape(-1), (len(x),))
    p1 = add_partials([vi_pass1(x[a:b], r[a:b], ua, sigma, l) for a, b in shard_bounds(len(x), shards)])
    sigma, tau, delta = _DT(sigma), _DT(tau), _DT(delta)
    Ka, _ = kernel_matrix(ua, ua, sigma, l)
    Sa = Ka + delta * np.eye(m + T, dtype=_DT)
    Z = tau ** 2 + delta
    B = 1 / Z
    n, s0 = p1["n"], p1["s0"]
    G1, b1 = p1["G1"][:m, :m], p1["b1"][:m]
    LS = _cholesky(Sa[:m, :m])
    LA = _cholesky(Sa[:m, :m] + B * G1)
    LSi, LAi = _inv_lower(LS), _inv_lower(LA)
    t1 = LAi @ (B * b1)
    Sinv = LSi.T @ LSi
    logdetS, logdetA = 2 * np.sum(np.log(np.diag(LS))), 2 * np.sum(np.log(np.diag(LA)))
    sumq = np.sum(Sinv * G1)
    log2pi = np.log(2 * _DT(np.pi) if _DT is np.float64 else 8 * np.arctan(_DT(1)))

    def objective(bCb, ldS, ldA, sq):
        tt = -(1 / (2 * tau ** 2)) * (n * (sigma ** 2 + delta) - sq)
        return -B * s0 / 2 + bCb / 2 - (n * np.log(Z) - ldS + ldA) / 2 - (n / 2) * log2pi + tt

    obj0 = objective(t1 @ t1, logdetS, logdetA, sumq)
    sT, gT = Sa[:m, m:], p1["G1"][:m, m:]
    kap, gam, kcr = np.diag(Sa)[m:], np.diag(p1["G1"])[m:], p1["b1"][m:]
    ES = LSi @ sT
    w = LSi.T @ ES
    H = G1 @ w
    EA = LAi @ (sT + B * gT)
    schurS = kap - np.sum(ES * ES, axis=0)
    schurA = kap + B * gam - np.sum(EA * EA, axis=0)
    out = np.full(T, np.nan, dtype=_DT)
    for t in range(T):
        if schurS[t] > 0 and schurA[t] > 0:
            tr_inc = (w[:, t] @ H[:, t] - 2 * (w[:, t] @ gT[:, t]) + gam[t]) / schurS[t]
            quad = (B * kcr[t] - EA[:, t] @ t1) ** 2 / schurA[t]
            out[t] = objective(t1 @ t1 + quad, logdetS + np.log(schurS[t]), logdetA + np.log(schurA[t]), sumq + tr_inc)
    return obj0, out


def _inv_lower(L):
    n = L.shape[0]
    X = np.zeros_like(L)
    for i in range(n):
        X[i, i] = 1 / L[i, i]
        X[i, :i] = -(L[i, :i] @ X[:i, :i]) / L[i, i]
    return X


def shard_bounds(n, world):
    """Contiguous row blocks, sizes differing by at most one (rank r gets rows [lo, hi))."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


# ====================================================================================================
# Gaussian FIC: obj_fun_norm + dlogp_dcov_par in row-separable form
# ====================================================================================================
def fic_rows_q(x, u, sigma, l, Sinv):
    """pass 1a (rows): q_i = K_i S^-1 K_i^T."""
    K, _ = kernel_matrix(x, u, sigma, l)
    return np.sum((K @ Sinv) * K, axis=1)


def fic_pass1(x, r, u, sigma, l, Bv):
    """pass 1b (rows): G_B = K^T diag(B) K, b = K^T (B r), s0 = sum B r^2, s1 = sum log Z."""
    K, _ = kernel_matrix(x, u, sigma, l)
    return {"GB": K.T @ (Bv[:, None] * K), "b": K.T @ (Bv * r), "s0": np.sum(Bv * r * r),
            "s1": np.sum(-np.log(Bv)), "n": _DT(len(r))}


def fic_mid(p1, S, Sinv, logdetS):
    Cm, logdetA = _spd_inverse(S + p1["GB"])
    v = Cm @ p1["b"]
    beta = Sinv @ (p1["b"] - p1["GB"] @ v)
    n = p1["n"]
    obj = -p1["s0"] / 2 + (p1["b"] @ v) / 2 - (p1["s1"] - logdetS + logdetA) / 2 - (n / 2) * np.log(8 * np.arctan(_DT(1)))
    M2 = Cm @ p1["GB"] @ Sinv
    return dict(obj=obj, C=Cm, v=v, beta=beta, M2=M2, n=n)


def fic_rows_rho(x, r, u, sigma, l, Bv, mid):
    """pass 2a (rows): c_i = K_i C K_i^T, alpha_i = B_i (r_i - K_i v), rho_i = alpha_i^2/2 - (B_i - B_i^2 c_i)/2."""
    K, _ = kernel_matrix(x, u, sigma, l)
    c = np.sum((K @ mid["C"]) * K, axis=1)
    alpha = Bv * (r - K @ mid["v"])
    rho = alpha ** 2 / 2 - (Bv - Bv ** 2 * c) / 2
    return alpha, rho


def fic_pass2(x, u, sigma, l, tau, Bv, alpha, rho, Sinv, mid):
    """pass 2b (rows): sum Omega o dK for sigma / l_c, G_rho = K^T diag(rho) K, sum rho, Q4 pairs."""
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    K, D = kernel_matrix(x, u, sigma, l)
    Omega = (-Bv - 2 * rho)[:, None] * (K @ Sinv) + Bv[:, None] * (K @ mid["M2"]) + np.outer(alpha, mid["beta"])
    P = Omega * K
    eq = _coincident(x, u)
    tau = _DT(tau)
    return {"g_sigma": 2 * P.sum(), "g_l": np.array([np.sum(P * D[:, :, c]) for c in range(D.shape[2])], dtype=_DT),
            "Grho": K.T @ (rho[:, None] * K), "sum_rho": rho.sum(),
            "g_tau_q4": 2 * tau ** 2 * Omega[eq].sum() if eq.any() else _DT(0), "g_knot": knot_colsums(P, x, u, l)}


def fic_obj_grad(x, y, mu, u, sigma, l, tau, delta, cov_fun="ard", shards=1, knots=False, bounds=None):
    kb = bounds
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    m, d = u.shape
    r = np.asarray(y, dtype=_DT).reshape(-1) - np.broadcast_to(np.asarray(mu, dtype=_DT).reshape(-1), (len(x),))
    bounds = shard_bounds(len(x), shards)
    sigma, tau, delta = _DT(sigma), _DT(tau), _DT(delta)
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    S = Kuu + delta * np.eye(m, dtype=_DT)
    Sinv, logdetS = _spd_inverse(S)
    Bs = [1 / (sigma ** 2 + tau ** 2 + delta - fic_rows_q(x[a:b], u, sigma, l, Sinv)) for a, b in bounds]
    p1 = add_partials([fic_pass1(x[a:b], r[a:b], u, sigma, l, Bs[k]) for k, (a, b) in enumerate(bounds)])
    mid = fic_mid(p1, S, Sinv, logdetS)
    ar = [fic_rows_rho(x[a:b], r[a:b], u, sigma, l, Bs[k], mid) for k, (a, b) in enumerate(bounds)]
    p2 = add_partials([fic_pass2(x[a:b], u, sigma, l, tau, Bs[k], ar[k][0], ar[k][1], Sinv, mid)
                       for k, (a, b) in enumerate(bounds)])
    GB = p1["GB"]
    N = Sinv @ GB @ Sinv / 2 - Sinv @ GB @ mid["M2"] / 2 - np.outer(mid["beta"], mid["beta"]) / 2 \
        + Sinv @ p2["Grho"] @ Sinv
    grad = {"sigma": p2["g_sigma"] + np.sum(N * dS_dtheta(u, sigma, l, tau, "sigma")) + 2 * sigma ** 2 * p2["sum_rho"]}
    if cov_fun == "ard":
        for c in range(d):
            nm = "l%d" % (c + 1)
            grad[nm] = p2["g_l"][c] + np.sum(N * dS_dtheta(u, sigma, l, tau, nm))
    else:
        grad["l"] = np.sum(p2["g_l"]) + np.sum(N * dS_dtheta(u, sigma, l, tau, "l", cov_fun))
    grad["tau"] = 2 * tau ** 2 * p2["sum_rho"] + p2["g_tau_q4"]
    if not knots:
        return mid["obj"], grad
    return mid["obj"], grad, knot_finish(p2["g_knot"], N, u, sigma, l, knot_bounds(x) if kb is None else kb)


# ====================================================================================================
# Sparse Laplace (Bernoulli / Poisson): Newton mode finder + gradient in matvec / Gram form
# (reference R/newtrap_sparseGP.R, R/laplace_approx_obj_funs.R:108-341,
#  R/laplace_approx_gradient.R:25-339).  One weighted Gram per Newton iteration.
# ====================================================================================================
def _softplus(x):
    return np.where(x > 0, x + np.log1p(np.exp(-np.abs(x))), np.log1p(np.exp(-np.abs(x))))


def lik_terms(family, ff, y, pois_m=1.0):
    """d1, W (= d2), W3 (= d3), log p(y | ff) -- quirk Q1 verbatim for Bernoulli."""
    if family == "bernoulli":
        pi = 1 / (1 + np.exp(-ff))
        d1 = y * (1 - pi) - pi + y * pi
        W = (1 - 2 * pi) * (y - pi) - (y * (1 - pi) ** 2 + pi ** 2 + y * pi ** 2)
        dpi = pi * (1 - pi)
        W3 = -2 * dpi * (y - pi) - dpi * (1 - 2 * pi) - (2 * y * (1 - pi) * (-dpi) + 2 * pi * dpi + 2 * y * pi * dpi)
        logpy = float(np.sum(y * (-_softplus(-ff)) + (1 - y) * (-_softplus(ff))))
    else:
        from scipy.special import gammaln
        ef = pois_m * np.exp(ff)
        d1, W, W3 = y - ef, -ef, -ef
        logpy = float(np.sum(y * np.log(pois_m) - gammaln(y + 1) - ef + y * ff))
    return d1, W, W3, logpy


def laplace_setup(x, u, sigma, l, tau, delta):
    K, _ = kernel_matrix(x, u, sigma, l)
    m = len(u)
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    S = Kuu + (tau ** 2 + delta) * np.eye(m)          # tau^2 kept (R/newtrap_sparseGP.R:51-60)
    Sinv, logdetS = _spd_inverse(S)
    Z = sigma ** 2 + tau ** 2 + delta - np.sum((K @ Sinv) * K, axis=1)
    GZ = K.T @ (K / Z[:, None])
    CZ = _inv(S + GZ)
    return dict(K=K, S=S, Sinv=Sinv, Z=Z, GZ=GZ, CZ=CZ, logdetS=logdetS)


def laplace_obj(st, family, ff, y, mu, pois_m=1.0):
    """Returns (objective, G_omega): the Gram is reused by the next Newton update."""
    K, Z = st["K"], st["Z"]
    d1, W, W3, logpy = lik_terms(family, ff, y, pois_m)
    omega = -W / (1 - W * Z)
    Gw = K.T @ (omega[:, None] * K)
    a = K.T @ ((ff - mu) / Z)
    logdetA = _spd_inverse(st["S"] + Gw)[1]
    obj = -0.5 * float(np.sum((ff - mu) ** 2 / Z)) + 0.5 * float(a @ st["CZ"] @ a) + logpy \
        - 0.5 * (-st["logdetS"] + logdetA) - 0.5 * float(np.sum(np.log(1 - W * Z)))
    return obj, Gw


def laplace_newton(x, y, mu, muu, u, sigma, l, tau, delta, family, ff0, maxit=1000, tol=1e-6, pois_m=1.0):
    st = laplace_setup(x, u, sigma, l, tau, delta)
    K, Z, S = st["K"], st["Z"], st["S"]
    ff = np.array(ff0, dtype=_DT)
    obj, Gw = laplace_obj(st, family, ff, y, mu, pois_m)
    hist = [obj]
    it = 1
    while True:
        it += 1
        d1, W, _, _ = lik_terms(family, ff, y, pois_m)
        e = 1 / (1 - Z * W)
        a = K.T @ ((ff - mu) / Z)
        h = st["CZ"] @ a
        Kh = K @ h
        grad_psi = d1 - (ff - mu) / Z + Kh / Z
        g2 = _solve(S + Gw, K.T @ (e * grad_psi))
        Gw_used = Gw
        ff = ff + Z * e * d1 - e * (ff - mu) + e * Kh + e * (K @ g2)
        obj, Gw = laplace_obj(st, family, ff, y, mu, pois_m)
        hist.append(obj)
        if not (it < maxit and (abs(hist[-1] - hist[-2]) > tol or np.any(np.abs(grad_psi) > tol))):
            break
    a = K.T @ ((ff - mu) / Z)
    u_mean = np.asarray(muu, dtype=_DT) + a - st["GZ"] @ (st["CZ"] @ a)
    u_var = S - Gw_used + Gw_used @ _solve(S + Gw_used, Gw_used)      # TT = -G_omega of the last update
    return dict(gp=ff, hist=np.array(hist), gradient=grad_psi, u_mean=u_mean, u_var=u_var)


def laplace_grad(x, y, mu, u, sigma, l, tau, delta, family, ff, cov_fun="ard", pois_m=1.0):
    x = np.asarray(x, dtype=_DT).reshape(len(x), -1)
    u = np.asarray(u, dtype=_DT).reshape(len(u), -1)
    d = u.shape[1]
    st = laplace_setup(x, u, sigma, l, tau, delta)
    K, Z, S, Sinv, CZ = st["K"], st["Z"], st["S"], st["Sinv"], st["CZ"]
    g, W, W3, _ = lik_terms(family, ff, y, pois_m)
    B = 1 / (Z - 1 / W)
    GB = K.T @ (B[:, None] * K)
    C = _inv(S + GB)
    a = K.T @ ((ff - mu) / Z)
    h = CZ @ a
    alpha = (ff - mu) / Z - (K @ h) / Z
    beta = Sinv @ (a - st["GZ"] @ h)
    GG = Sinv @ (K.T @ g)
    c = np.sum((K @ C) * K, axis=1)
    Dv = W - 1 / Z
    comp4 = -1 / Dv + c / (Z * W - 1) ** 2
    uvec = comp4 * (-W3)
    t = -uvec * B / W + B * (K @ (C @ (K.T @ (B * uvec / W))))
    rho = 0.5 * alpha ** 2 - 0.5 * (B - B ** 2 * c) - 0.5 * t * g
    M2 = C @ GB @ Sinv
    Omega = (-B - 2 * rho)[:, None] * (K @ Sinv) + B[:, None] * (K @ M2) + np.outer(alpha, beta) - np.outer(t, GG)
    _, D = kernel_matrix(x, u, sigma, l)
    P = Omega * K
    Grho = K.T @ (rho[:, None] * K)
    N = 0.5 * Sinv @ GB @ Sinv - 0.5 * Sinv @ GB @ M2 - 0.5 * np.outer(beta, beta) + Sinv @ Grho @ Sinv \
        + 0.5 * np.outer(Sinv @ (K.T @ t), GG)
    sum_rho = float(rho.sum())
    grad = {"sigma": 2 * float(P.sum()) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "sigma"))) + 2 * sigma ** 2 * sum_rho}
    if cov_fun == "ard":
        for cdim in range(d):
            nm = "l%d" % (cdim + 1)
            grad[nm] = float(np.sum(P * D[:, :, cdim])) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, nm)))
    else:
        grad["l"] = float(np.sum(P * D.sum(axis=2))) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "l", cov_fun)))
    equ = _coincident(u, u)
    eqx = _coincident(x, u)
    grad["tau"] = 2 * tau ** 2 * (sum_rho + float(N[equ].sum()) + (float(Omega[eqx].sum()) if eqx.any() else 0.0))
    return grad
