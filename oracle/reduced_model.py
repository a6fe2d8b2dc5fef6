"""Row-separable ("reduced") form of the sparse-GP objective + gradient -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy statement of the two-pass algorithm the CUDA product implements: every n-dependent
quantity is a sum over data rows of row-local terms once the replicated m x m factors are known, so the rows can
be split into shards whose partial sums are added (the allreduce).  tests/ use it to
  * prove the reduced algebra equal to the literal transcription of the reference (oracle/ref_model.py),
  * exercise the shard -> allreduce -> replicate logic under gloo with world_size 2 on the CPU,
  * localise a parity failure of the CUDA path to one pass / one m x m stage.
It follows SURVEY.md Appendix B.2/B.3, which restate R/vi_functions.R:64-121,126-420 and
R/laplace_approx_gradient.R:720-968; formulas are derived in DESIGN.md.

PARITY UNPINNED (see oracle/ref_model.py).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla


def kernel_matrix(x, u, sigma, l, cov_fun="ard"):
    """K_ij and the per-dimension scaled squared differences D_ijc = ((x_ic - u_jc)/l_c)^2."""
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    l = np.broadcast_to(np.asarray(l, dtype=np.float64).reshape(-1), (x.shape[1],))
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
    return {"G1": K.T @ K, "b1": K.T @ r, "s0": float(r @ r), "n": float(len(r))}


def add_partials(parts):
    out = {}
    for k in parts[0]:
        out[k] = sum(p[k] for p in parts)
    return out


# ---------------------------------------------------------------- replicated m x m --------------
def vi_mid(p1, u, sigma, l, tau, delta, cov_fun="ard"):
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    m = u.shape[0]
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    S = Kuu + delta * np.eye(m)                      # self-cov minus tau^2 I  (R/vi_functions.R:736-741)
    Z = tau ** 2 + delta
    B = 1.0 / Z
    n = p1["n"]
    LS = np.linalg.cholesky(S)
    Sinv = sla.cho_solve((LS, True), np.eye(m))
    logdetS = 2 * np.sum(np.log(np.diag(LS)))
    G = B * p1["G1"]
    LA = np.linalg.cholesky(S + G)
    Cm = sla.cho_solve((LA, True), np.eye(m))
    logdetA = 2 * np.sum(np.log(np.diag(LA)))
    b = B * p1["b1"]
    v = Cm @ b
    beta = Sinv @ (b - G @ v)
    sum_q = float(np.sum(Sinv * p1["G1"]))
    tt = -(1 / (2 * tau ** 2)) * (n * (sigma ** 2 + delta) - sum_q)
    obj = -0.5 * B * p1["s0"] + 0.5 * float(b @ v) - 0.5 * (n * math.log(Z) - logdetS + logdetA) \
        - (n / 2) * math.log(2 * math.pi) + tt
    CGS = Cm @ G @ Sinv
    M = (1 / tau ** 2 - B) * Sinv + B * CGS
    SGS = Sinv @ G @ Sinv
    N = 0.5 * SGS - 0.5 * Sinv @ G @ CGS - 0.5 * np.outer(beta, beta) - (1 / (2 * tau ** 2)) * Sinv @ p1["G1"] @ Sinv
    return dict(obj=obj, tt=tt, B=B, v=v, beta=beta, M=M, N=N, Sinv=Sinv, C=Cm, trCG1=float(np.sum(Cm * p1["G1"])),
                n=n)


# ---------------------------------------------------------------- pass 2 (rows) -----------------
def vi_pass2(x, r, u, sigma, l, tau, mid):
    """Per-shard partials: sum_ij Omega_ij dK_ij(theta) for sigma and each length scale, sum alpha^2,
    and the coincident-row correction of quirk Q4 (dK/dlog tau = 2 tau^2 where x_i == u_j bit-exactly)."""
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    K, D = kernel_matrix(x, u, sigma, l)
    alpha = mid["B"] * (r - K @ mid["v"])
    Omega = K @ mid["M"] + np.outer(alpha, mid["beta"])
    P = Omega * K
    g_sigma = 2 * float(P.sum())
    g_l = np.einsum("ij,ijc->c", P, D)
    eq = _coincident(x, u)
    g_tau_q4 = 0.0
    if eq.any():
        Om_tau = Omega - (1 / tau ** 2) * (K @ mid["Sinv"])
        g_tau_q4 = 2 * tau ** 2 * float(Om_tau[eq].sum())
    return {"g_sigma": g_sigma, "g_l": g_l, "sum_alpha2": float(alpha @ alpha), "g_tau_q4": g_tau_q4}


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
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    d = u.shape[1]
    n, B, N = mid["n"], mid["B"], mid["N"]
    grad = {}
    grad["sigma"] = p2["g_sigma"] + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "sigma"))) \
        - (1 / (2 * tau ** 2)) * 2 * sigma ** 2 * n
    if cov_fun == "ard":
        for c in range(d):
            nm = "l%d" % (c + 1)
            grad[nm] = float(p2["g_l"][c]) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, nm)))
    else:
        grad["l"] = float(np.sum(p2["g_l"])) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "l", cov_fun)))
    grad["tau"] = tau ** 2 * p2["sum_alpha2"] - tau ** 2 * (n * B - B ** 2 * mid["trCG1"]) - 2 * mid["tt"] \
        + p2["g_tau_q4"]
    return grad


def vi_obj_grad(x, y, mu, u, sigma, l, tau, delta, cov_fun="ard", shards=1):
    """Full evaluation; `shards` splits the rows into that many contiguous blocks (the multi-GPU layout)."""
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
    r = np.asarray(y, dtype=np.float64).reshape(-1) - np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (len(x),))
    bounds = shard_bounds(len(x), shards)
    p1 = add_partials([vi_pass1(x[a:b], r[a:b], u, sigma, l) for a, b in bounds])
    mid = vi_mid(p1, u, sigma, l, tau, delta, cov_fun)
    p2 = add_partials([vi_pass2(x[a:b], r[a:b], u, sigma, l, tau, mid) for a, b in bounds])
    return mid["obj"], vi_finish(p2, mid, u, sigma, l, tau, cov_fun)


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
    return {"GB": K.T @ (Bv[:, None] * K), "b": K.T @ (Bv * r), "s0": float(np.sum(Bv * r * r)),
            "s1": float(np.sum(-np.log(Bv))), "n": float(len(r))}


def fic_mid(p1, S, Sinv, logdetS):
    m = S.shape[0]
    LA = np.linalg.cholesky(S + p1["GB"])
    Cm = sla.cho_solve((LA, True), np.eye(m))
    logdetA = 2 * np.sum(np.log(np.diag(LA)))
    v = Cm @ p1["b"]
    beta = Sinv @ (p1["b"] - p1["GB"] @ v)
    n = p1["n"]
    obj = -0.5 * p1["s0"] + 0.5 * float(p1["b"] @ v) - 0.5 * (p1["s1"] - logdetS + logdetA) - (n / 2) * math.log(2 * math.pi)
    M2 = Cm @ p1["GB"] @ Sinv
    return dict(obj=obj, C=Cm, v=v, beta=beta, M2=M2, n=n)


def fic_rows_rho(x, r, u, sigma, l, Bv, mid):
    """pass 2a (rows): c_i = K_i C K_i^T, alpha_i = B_i (r_i - K_i v), rho_i = alpha_i^2/2 - (B_i - B_i^2 c_i)/2."""
    K, _ = kernel_matrix(x, u, sigma, l)
    c = np.sum((K @ mid["C"]) * K, axis=1)
    alpha = Bv * (r - K @ mid["v"])
    rho = 0.5 * alpha ** 2 - 0.5 * (Bv - Bv ** 2 * c)
    return alpha, rho


def fic_pass2(x, u, sigma, l, tau, Bv, alpha, rho, Sinv, mid):
    """pass 2b (rows): sum Omega o dK for sigma / l_c, G_rho = K^T diag(rho) K, sum rho, Q4 pairs."""
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    K, D = kernel_matrix(x, u, sigma, l)
    Omega = (-Bv - 2 * rho)[:, None] * (K @ Sinv) + Bv[:, None] * (K @ mid["M2"]) + np.outer(alpha, mid["beta"])
    P = Omega * K
    eq = _coincident(x, u)
    return {"g_sigma": 2 * float(P.sum()), "g_l": np.einsum("ij,ijc->c", P, D), "Grho": K.T @ (rho[:, None] * K),
            "sum_rho": float(rho.sum()), "g_tau_q4": 2 * tau ** 2 * float(Omega[eq].sum()) if eq.any() else 0.0}


def fic_obj_grad(x, y, mu, u, sigma, l, tau, delta, cov_fun="ard", shards=1):
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)
    u = np.asarray(u, dtype=np.float64).reshape(len(u), -1)
    m, d = u.shape
    r = np.asarray(y, dtype=np.float64).reshape(-1) - np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (len(x),))
    bounds = shard_bounds(len(x), shards)
    Kuu, _ = kernel_matrix(u, u, sigma, l)
    S = Kuu + delta * np.eye(m)
    LS = np.linalg.cholesky(S)
    Sinv = sla.cho_solve((LS, True), np.eye(m))
    logdetS = 2 * np.sum(np.log(np.diag(LS)))
    Bs = [1.0 / (sigma ** 2 + tau ** 2 + delta - fic_rows_q(x[a:b], u, sigma, l, Sinv)) for a, b in bounds]
    p1 = add_partials([fic_pass1(x[a:b], r[a:b], u, sigma, l, Bs[k]) for k, (a, b) in enumerate(bounds)])
    mid = fic_mid(p1, S, Sinv, logdetS)
    ar = [fic_rows_rho(x[a:b], r[a:b], u, sigma, l, Bs[k], mid) for k, (a, b) in enumerate(bounds)]
    p2 = add_partials([fic_pass2(x[a:b], u, sigma, l, tau, Bs[k], ar[k][0], ar[k][1], Sinv, mid)
                       for k, (a, b) in enumerate(bounds)])
    GB = p1["GB"]
    N = 0.5 * Sinv @ GB @ Sinv - 0.5 * Sinv @ GB @ mid["M2"] - 0.5 * np.outer(mid["beta"], mid["beta"]) \
        + Sinv @ p2["Grho"] @ Sinv
    grad = {"sigma": p2["g_sigma"] + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "sigma"))) + 2 * sigma ** 2 * p2["sum_rho"]}
    if cov_fun == "ard":
        for c in range(d):
            nm = "l%d" % (c + 1)
            grad[nm] = float(p2["g_l"][c]) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, nm)))
    else:
        grad["l"] = float(np.sum(p2["g_l"])) + float(np.sum(N * dS_dtheta(u, sigma, l, tau, "l", cov_fun)))
    grad["tau"] = 2 * tau ** 2 * p2["sum_rho"] + p2["g_tau_q4"]
    return mid["obj"], grad
