"""Golden vectors from the reference's OWN R sources (rows a14-a25 and f1/f2 of SURVEY.md section 8) -- TEST TOOL.

Runs /root/reference/R/*.R, unmodified, in oracle/mini_r (no R in this image) with the Rcpp exports served by the
reference's compiled C++ (oracle/_ref), on small seeded cases, and records inputs + outputs in
tests/golden/r_level.npz (+ .json index).  tests/test_reference_r.py pins oracle/ref_model.py to these numbers
(and, in the build container, re-runs the R code to check the file is current); tests/test_golden_r_gpu.py holds
the CUDA path to them on the GPU box, where /root/reference does not exist.

    python tests/tools/make_golden_r.py            # rewrite tests/golden/r_level.*
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_r as rr          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "r_level")

# The few lines of R glue below are OURS: they are what the reference's callers do before they reach the functions
# under test (R/vi_functions.R:730-760, R/laplace_gradient_ascent.R:1237-1263, R/newtrap_sparseGP.R:43-66).
R_GLUE = r'''
.srgp_sig <- function(cov_par, cov_fun, xy, xu, delta, keep_tau) {
  if(cov_fun == "ard") {
    lnames <- paste("l", 1:ncol(xy), sep = "")
    Sigma12 <- make_cov_mat_ardC(x = xy, x_pred = xu, cov_par = cov_par, cov_fun = cov_fun, delta = delta, lnames = lnames)
    Sigma22 <- make_cov_mat_ardC(x = xu, x_pred = matrix(), cov_par = cov_par, cov_fun = cov_fun, delta = delta, lnames = lnames)
  } else {
    Sigma12 <- make_cov_matC(x = xy, x_pred = xu, cov_par = cov_par, cov_fun = cov_fun, delta = delta)
    Sigma22 <- make_cov_matC(x = xu, x_pred = matrix(), cov_par = cov_par, cov_fun = cov_fun, delta = delta)
  }
  if(!keep_tau) Sigma22 <- Sigma22 - as.list(cov_par)$tau^2 * diag(nrow(xu))
  list(Sigma12 = Sigma12, Sigma22 = Sigma22)
}
.srgp_fic_Z <- function(cov_par, Sigma12, Sigma22, delta) {
  Z2 <- solve(a = Sigma22, b = t(Sigma12))
  Z3 <- Sigma12 * t(Z2)
  Z4 <- apply(X = Z3, MARGIN = 1, FUN = sum)
  cov_par$sigma^2 + cov_par$tau^2 + delta - Z4
}
.srgp_dtheta <- function(cov_fun) {
  if(cov_fun == "sqexp") return(list("sigma" = dsqexp_dsigma, "l" = dsqexp_dl, "tau" = dsqexp_dtau))
  list("sigma" = dsqexp_dsigma_ard, "tau" = dsqexp_dtau)
}
.srgp_dknot <- function(cov_fun) { if(cov_fun == "sqexp") dsqexp_dx2 else dsqexp_dx2_ard }
.srgp_scores <- numeric()
.srgp_rec_elbo <- function(...) { v <- elbo_fun(...); .srgp_scores <<- c(.srgp_scores, as.numeric(v)); v }
.srgp_rec_norm <- function(...) { v <- obj_fun_norm(...); .srgp_scores <<- c(.srgp_scores, as.numeric(v)); v }
.srgp_reset_scores <- function() { .srgp_scores <<- numeric(); 0 }
.srgp_get_scores <- function() .srgp_scores
'''


def ard_par(sigma, ls, tau):
    cp = {"sigma": float(sigma)}
    for i, l in enumerate(ls):
        cp["l%d" % (i + 1)] = float(l)
    cp["tau"] = float(tau)
    return cp


def gauss_cases():
    out = {}
    rng = np.random.default_rng(2201)
    # config 1 shape: 1-D sqexp, theta = (2, 1, 1), knots (1,3,5,7,9)
    n = 60
    x = np.sort(rng.uniform(0, 10, n)).reshape(-1, 1)
    y = 2.0 * np.sin(x[:, 0]) + 0.5 * np.cos(2.3 * x[:, 0]) + rng.normal(size=n)
    out["g_sqexp_1d"] = dict(xy=x, y=y, mu=np.zeros(n), xu=np.array([1.0, 3, 5, 7, 9]).reshape(-1, 1),
                             cov_par={"sigma": 2.0, "l": 1.0, "tau": 1.0}, cov_fun="sqexp", delta=1e-6, knots=True)
    # 2-D sqexp (isotropic length scale), non-zero mean
    n, d, m = 50, 2, 6
    x = rng.normal(size=(n, d))
    out["g_sqexp_2d_mu"] = dict(xy=x, y=np.cos(x[:, 0]) + 0.2 * rng.normal(size=n), mu=0.3 + 0.1 * x[:, 0],
                                xu=rng.normal(size=(m, d)), cov_par={"sigma": 1.3, "l": 0.8, "tau": 0.35},
                                cov_fun="sqexp", delta=1e-5, knots=True)
    # config 2 shape (airfoil): d = 5 ARD, knots = jittered data rows, delta = 1e-4
    n, d, m = 70, 5, 8
    x = rng.normal(size=(n, d))
    y = np.sin(x[:, 0]) + 0.5 * x[:, 1] - 0.3 * x[:, 2] * x[:, 3] + 0.3 * rng.normal(size=n)
    s = float(np.sqrt(np.var(y) / 2))
    out["g_ard_d5"] = dict(xy=x, y=y, mu=np.zeros(n), xu=x[rng.choice(n, m, replace=False)] + 0.01 * rng.normal(size=(m, d)),
                           cov_par=ard_par(s, [1.0] * d, s), cov_fun="ard", delta=1e-4, knots=True)
    # config 3 shape (OAT step): one knot IS a data row (the nugget pairs of quirk Q4), delta = 1e-3
    n, d, m = 64, 4, 7
    x = rng.normal(size=(n, d))
    y = 0.8 * x[:, 0] - 0.4 * np.tanh(x[:, 1]) + 0.2 * x[:, 2] * x[:, 3] + 0.25 * rng.normal(size=n)
    xu = np.vstack([x[rng.choice(n, m, replace=False)] + 0.05 * rng.normal(size=(m, d)), x[11]])
    out["g_ard_d4_coincident"] = dict(xy=x, y=y, mu=np.zeros(n), xu=xu, cov_par=ard_par(1.0, [1.2, 0.9, 1.5, 1.1], 0.3),
                                      cov_fun="ard", delta=1e-3, knots=False)
    # config 5 shape (headline): d = 8 ARD, X, U ~ N(0, I), l_c = 0.8 + 0.05 c, tau 0.5, delta 1e-6
    n, d, m = 90, 8, 12
    x = rng.normal(size=(n, d))
    cp = ard_par(1.0, [0.8 + 0.05 * (c + 1) for c in range(d)], 0.5)
    out["g_ard_d8"] = dict(xy=x, y=np.sin(x[:, 0]) + 0.5 * x[:, 1] + 0.5 * rng.normal(size=n), mu=np.zeros(n),
                           xu=rng.normal(size=(m, d)), cov_par=cp, cov_fun="ard", delta=1e-6, knots=False)
    return out


def laplace_cases():
    out = {}
    rng = np.random.default_rng(2202)
    n, d, m = 60, 3, 6
    x = rng.normal(size=(n, d))
    f = 1.5 * np.sin(x[:, 0]) + x[:, 1]
    out["l_bern_ard_d3"] = dict(xy=x, y=(rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float), mu=np.zeros(n),
                                muu=np.zeros(m), xu=rng.normal(size=(m, d)), cov_par=ard_par(2.0, [1.5, 1.2, 1.8], 0.1),
                                cov_fun="ard", delta=1e-3, family="bernoulli", knots=True, extra={})
    n, d, m = 50, 1, 5
    x = np.sort(rng.uniform(0, 10, n)).reshape(-1, 1)
    lam = np.exp(0.8 * np.sin(x[:, 0]) + 0.5)
    mm = np.full(n, 0.7)
    out["l_pois_sqexp_1d"] = dict(xy=x, y=rng.poisson(lam * mm).astype(float), mu=np.full(n, 0.2), muu=np.full(m, 0.2),
                                  xu=np.array([1.0, 3, 5, 7, 9]).reshape(-1, 1), cov_par={"sigma": 1.2, "l": 1.5, "tau": 0.05},
                                  cov_fun="sqexp", delta=1e-4, family="poisson", knots=True, extra={"m": mm})
    n, d, m = 80, 8, 10
    x = rng.normal(size=(n, d))
    f = 1.5 * np.sin(x[:, 0]) + x[:, 1] - 0.5 * x[:, 2]
    out["l_bern_ard_d8"] = dict(xy=x, y=(rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(float), mu=np.zeros(n),
                                muu=np.zeros(m), xu=rng.normal(size=(m, d)), cov_par=ard_par(2.0, [1.5] * d, 0.1),
                                cov_fun="ard", delta=1e-3, family="bernoulli", knots=False, extra={})
    return out


def run_gauss(c):
    """Everything the reference computes for one Gaussian case, through its own R functions."""
    m = c["xu"].shape[0]
    rr.session().run(R_GLUE)
    sig = rr.call(".srgp_sig", cov_par=c["cov_par"], cov_fun=c["cov_fun"], xy=c["xy"], xu=c["xu"], delta=c["delta"], keep_tau=False)
    S12, S22 = sig["Sigma12"], sig["Sigma22"]
    res = {"Sigma12": S12, "Sigma22": S22}
    n = S12.shape[0]
    Z_vi = np.full(n, c["cov_par"]["tau"] ** 2 + c["delta"])                         # R/vi_functions.R:753
    Z_fic = rr.call(".srgp_fic_Z", cov_par=c["cov_par"], Sigma12=S12, Sigma22=S22, delta=c["delta"])
    res["Z_fic"] = Z_fic
    tt = rr.call("trace_term_fun", cov_par=c["cov_par"], Sigma12=S12, Sigma22=S22, delta=c["delta"])
    res["trace_term"] = np.asarray(tt).reshape(-1)
    res["dtrace_term_dtau"] = np.asarray(rr.call("dtrace_term_dtau", cov_par=c["cov_par"], trace_term=tt)).reshape(-1)
    res["elbo"] = np.asarray(rr.call("elbo_fun", mu=c["mu"], Z=Z_vi, Sigma12=S12, Sigma22=S22, y=c["y"],
                                     trace_term_fun=rr.rfun("trace_term_fun"), cov_par=c["cov_par"], delta=c["delta"])).reshape(-1)
    res["obj_fun_norm"] = np.asarray(rr.call("obj_fun_norm", mu=c["mu"], Z=Z_fic, Sigma12=S12, Sigma22=S22, y=c["y"])).reshape(-1)
    dth = rr.call(".srgp_dtheta", cov_fun=c["cov_fun"])
    dkn = rr.call(".srgp_dknot", cov_fun=c["cov_fun"]) if c["knots"] else None
    for name, fn in (("vi", "delbo_dcov_par"), ("fic", "dlogp_dcov_par")):
        kw = dict(cov_par=c["cov_par"], cov_fun=c["cov_fun"], dcov_fun_dtheta=dth, knot_opt=np.arange(1, m + 1),
                  xu=c["xu"], xy=c["xy"], y=c["y"], mu=c["mu"], transform=True, delta=c["delta"])
        if dkn is not None:
            kw["dcov_fun_dknot"] = dkn
        g = rr.call(fn, **kw)
        res[name + "_gradient"] = np.array([float(np.asarray(g["gradient"][k]).reshape(-1)[0]) for k in c["cov_par"]])
        res[name + "_trans_par"] = np.array([float(np.asarray(g["trans_par"][k]).reshape(-1)[0]) for k in c["cov_par"]])
        if dkn is not None:
            res[name + "_knot_gradient"] = np.asarray(g["knot_gradient"]).reshape(-1)
            res[name + "_trans_knot"] = np.asarray(g["trans_knot"])
    return res


def run_laplace(c):
    """Newton mode search, objective, gradient and prediction of one sparse Laplace case through the reference's R."""
    sfx = {"bernoulli": "bern", "poisson": "pois"}[c["family"]]
    m = c["xu"].shape[0]
    rr.session().run(R_GLUE)
    n = c["xy"].shape[0]
    nr = rr.call("newtrap_sparseGP", start_vals=np.array(c["mu"], dtype=float), obj_fun=rr.rfun("obj_fun_" + sfx),
                 grad_loglik_fn=rr.rfun("grad_loglik_fn_" + sfx), dlog_py_dff=rr.rfun("dlog_py_dff_" + sfx),
                 d2log_py_dff=rr.rfun("d2log_py_dff_" + sfx), maxit=1000, tol=1e-6, cov_par=c["cov_par"],
                 cov_fun=c["cov_fun"], xy=c["xy"], xu=c["xu"], y=c["y"], mu=c["mu"], muu=c["muu"], delta=c["delta"],
                 **c["extra"])
    res = {"gp": np.asarray(nr["gp"]).reshape(-1),
           "objective_function_values": np.asarray(nr["objective_function_values"]).reshape(-1),
           "newton_gradient": np.asarray(nr["gradient"]).reshape(-1),
           "u_posterior_mean": np.asarray(nr["u_posterior_mean"]).reshape(-1),
           "u_posterior_variance": np.asarray(nr["u_posterior_variance"])}
    dth = rr.call(".srgp_dtheta", cov_fun=c["cov_fun"])
    kw = dict(cov_par=c["cov_par"], cov_fun=c["cov_fun"], dcov_fun_dtheta=dth, knot_opt=np.arange(1, m + 1), xu=c["xu"],
              xy=c["xy"], y=c["y"], ff=res["gp"], dlog_py_dff=rr.rfun("dlog_py_dff_" + sfx),
              d2log_py_dff=rr.rfun("d2log_py_dff_" + sfx), d3log_py_dff=rr.rfun("d3log_py_dff_" + sfx), mu=c["mu"],
              transform=True, delta=c["delta"], **c["extra"])
    if c["knots"]:
        kw["dcov_fun_dknot"] = rr.call(".srgp_dknot", cov_fun=c["cov_fun"])
    g = rr.call("dlogq_dcov_par", **kw)
    res["gradient"] = np.array([float(np.asarray(g["gradient"][k]).reshape(-1)[0]) for k in c["cov_par"]])
    if c["knots"]:
        res["knot_gradient"] = np.asarray(g["knot_gradient"]).reshape(-1)
        res["trans_knot"] = np.asarray(g["trans_knot"])
    rng = np.random.default_rng(7)
    x_pred = c["xy"][:9] + 0.1 * rng.normal(size=(9, c["xy"].shape[1]))
    res["x_pred"] = x_pred
    pr = rr.call("predict_laplace", u_mean=res["u_posterior_mean"], u_var=res["u_posterior_variance"], xu=c["xu"],
                 x_pred=x_pred, cov_fun=c["cov_fun"], cov_par=c["cov_par"], mu=np.full(9, float(c["mu"][0])), muu=c["muu"],
                 full_cov=False, family=c["family"], delta=c["delta"])
    res["pred_mean"], res["pred_var"] = np.asarray(pr["pred_mean"]).reshape(-1), np.asarray(pr["pred_var"]).reshape(-1)
    return res


FIT_IT = 5


def run_fit(c):
    """The reference's optimiser loops (norm_grad_ascent_vi, norm_grad_ascent, laplace_grad_ascent), FIT_IT
    iterations of ADADELTA on log(theta) and the knots, then predict_vi / predict_laplace from the returned posterior."""
    rr.session().run(R_GLUE)
    m = c["xu"].shape[0]
    fam = c.get("family", "gaussian")
    dth = rr.call(".srgp_dtheta", cov_fun=c["cov_fun"])
    dkn = rr.call(".srgp_dknot", cov_fun=c["cov_fun"]) if c["knots"] else np.nan       # NA: theta only
    opt = {"maxit": FIT_IT, "delta": c["delta"], "obj_tol": 0.0}
    common = dict(cov_par_start=c["cov_par"], cov_fun=c["cov_fun"], dcov_fun_dtheta=dth, dcov_fun_dknot=dkn,
                  knot_opt=np.arange(1, m + 1), xu=c["xu"], xy=c["xy"], y=c["y"], mu=c["mu"], muu=c["muu"], opt=opt, verbose=False)
    if c["model"] == "vi":
        out = rr.call("norm_grad_ascent_vi", **common)
    elif c["model"] == "fic":
        out = rr.call("norm_grad_ascent", obj_fun=rr.rfun("obj_fun_norm"), transform=True, **common)
    else:
        sfx = {"bernoulli": "bern", "poisson": "pois"}[fam]
        out = rr.call("laplace_grad_ascent", ff=np.array(c["mu"], dtype=float), grad_loglik_fn=rr.rfun("grad_loglik_fn_" + sfx),
                      dlog_py_dff=rr.rfun("dlog_py_dff_" + sfx), d2log_py_dff=rr.rfun("d2log_py_dff_" + sfx),
                      d3log_py_dff=rr.rfun("d3log_py_dff_" + sfx), obj_fun=rr.rfun("obj_fun_" + sfx), transform=True,
                      **common, **c["extra"])
    names = list(c["cov_par"])
    res = {"cov_par": np.array([float(np.asarray(out["cov_par"][k]).reshape(-1)[0]) for k in names]),
           "xu_final": np.asarray(out["xu"]), "iter": np.asarray(out["iter"]).reshape(-1),
           "obj_fun": np.asarray(out["obj_fun"]).reshape(-1), "grad": np.asarray(out["grad"]),
           "cov_par_history": np.asarray(out["cov_par_history"]),
           "u_mean": np.asarray(out["u_mean"]).reshape(-1), "u_var": np.asarray(out["u_var"])}
    if c["knots"]:
        res["knot_grad"] = np.asarray(out["knot_grad"])
    rng = np.random.default_rng(8)
    x_pred = c["xy"][:7] + 0.1 * rng.normal(size=(7, c["xy"].shape[1]))
    res["x_pred"] = x_pred
    cp_final = dict(zip(names, res["cov_par"].tolist()))
    pred = "predict_vi" if c["model"] == "vi" else "predict_laplace"
    pr = rr.call(pred, u_mean=res["u_mean"], u_var=res["u_var"], xu=res["xu_final"], x_pred=x_pred, cov_fun=c["cov_fun"],
                 cov_par=cp_final, mu=np.full(7, float(c["mu"][0])), muu=c["muu"], full_cov=False, family=fam, delta=c["delta"])
    res["pred_mean"], res["pred_var"] = np.asarray(pr["pred_mean"]).reshape(-1), np.asarray(pr["pred_var"]).reshape(-1)
    return res


def fit_cases():
    g, l = gauss_cases(), laplace_cases()
    out = {}
    for nm, src, model in (("f_vi_sqexp_2d", g["g_sqexp_2d_mu"], "vi"), ("f_fic_sqexp_2d", g["g_sqexp_2d_mu"], "fic"),
                           ("f_vi_ard_d5_theta", dict(g["g_ard_d5"], knots=False), "vi"),
                           ("f_fic_ard_d5", g["g_ard_d5"], "fic"),
                           ("f_lap_bern_d3", l["l_bern_ard_d3"], "laplace"), ("f_lap_pois_1d", l["l_pois_sqexp_1d"], "laplace")):
        c = dict(src, model=model)
        c.setdefault("muu", np.zeros(c["xu"].shape[0]))
        c.setdefault("extra", {})
        out[nm] = c
    return out


def run_oat(c):
    """The reference's candidate-selection functions knot_prop_random_norm_vi (R/vi_functions.R:2108-2304) and
    knot_prop_random_norm (R/knot_proposal_functions.R:1176-1357), run unmodified; their only random input --
    sample.int() choosing TTmax candidate rows of xy -- is replaced by the fixed draw c["draw"] (1-based), and the
    objective function handed in through `...` is a recording wrapper around the reference's own elbo_fun /
    obj_fun_norm, so the per-candidate scores of the loop come back besides the chosen knot."""
    from oracle.mini_r import interp as RI
    I = rr.session()
    I.run(R_GLUE)
    draw = np.asarray(c["draw"], dtype=np.int64)
    I.globalenv.vars["sample.int"] = RI.Builtin("sample.int", lambda I_, pos, kw: RI.Vec(draw.copy()))
    rr.call(".srgp_reset_scores")
    vi = c["model"] == "vi"
    norm_opt = {"xu": c["xu"], "cov_par": c["cov_par"], "xy": c["xy"], "mu": c["mu"], "cov_fun": c["cov_fun"],
                "obj_fun": np.array([c["obj_current"] - 1.0, c["obj_current"]])}
    out = rr.call("knot_prop_random_norm_vi" if vi else "knot_prop_random_norm", norm_opt=norm_opt,
                  opt={"TTmax": len(draw), "delta": c["delta"]}, y=c["y"],
                  obj_fun=rr.rfun(".srgp_rec_elbo" if vi else ".srgp_rec_norm"), cov_fun=c["cov_fun"])
    scores = np.asarray(rr.call(".srgp_get_scores")).reshape(-1)
    del I.globalenv.vars["sample.int"]
    return {"chosen": np.asarray(out).reshape(1, -1), "scores": scores, "pseudo_prop": c["xy"][draw - 1]}


def run_oat_laplace(c, nr):
    """knot_prop_random (R/knot_proposal_functions.R:1001-1175) run unmodified with a fixed sample.int draw: one
    warm-started newtrap_sparseGP per candidate; returns the chosen knot.  nr = the Newton result the step starts from."""
    from oracle.mini_r import interp as RI
    I = rr.session()
    I.run(R_GLUE)
    sfx = {"bernoulli": "bern", "poisson": "pois"}[c["family"]]
    draw = np.asarray(c["draw"], dtype=np.int64)
    I.globalenv.vars["sample.int"] = RI.Builtin("sample.int", lambda I_, pos, kw: RI.Vec(draw.copy()))
    lo = {"xu": c["xu"], "cov_par": c["cov_par"], "xy": c["xy"], "mu": c["mu"], "muu": c["muu"], "cov_fun": c["cov_fun"],
          "fmax": nr["gp"], "obj_fun": nr["objective_function_values"]}
    out = rr.call("knot_prop_random", laplace_opt=lo, opt={"TTmax": len(draw), "delta": c["delta"]}, cov_fun=c["cov_fun"], y=c["y"],
                  obj_fun=rr.rfun("obj_fun_" + sfx), grad_loglik_fn=rr.rfun("grad_loglik_fn_" + sfx),
                  dlog_py_dff=rr.rfun("dlog_py_dff_" + sfx), d2log_py_dff=rr.rfun("d2log_py_dff_" + sfx), maxit=1000, tol=1e-6,
                  **c["extra"])
    del I.globalenv.vars["sample.int"]
    return {"chosen": np.asarray(out).reshape(1, -1), "pseudo_prop": c["xy"][draw - 1], "fmax": nr["gp"],
            "obj_current": np.asarray(nr["objective_function_values"]).reshape(-1)[-1:]}


def oat_cases():
    from oracle import ref_model as rm
    g = gauss_cases()
    out = {}
    for nm, src, model, draw, shift in (("o_vi_ard_d5", g["g_ard_d5"], "vi", [5, 17, 33, 60, 2, 48], 0.0),
                                        ("o_fic_ard_d5", g["g_ard_d5"], "fic", [5, 17, 33, 60, 2, 48], 0.0),
                                        ("o_vi_sqexp_2d", g["g_sqexp_2d_mu"], "vi", [1, 9, 30, 44], 0.0),
                                        ("o_vi_none_better", g["g_sqexp_2d_mu"], "vi", [1, 9, 30, 44], 1e6)):
        c = dict(src, model=model, draw=np.array(draw, dtype=np.float64))
        f = rm.vi_obj_grad if model == "vi" else rm.fic_obj_grad
        # the last objective value of the fit the OAT step starts from (shift > 0: no candidate can beat it, the reference
        # then returns the FIRST existing knot)
        c["obj_current"] = float(f(c["cov_par"], c["cov_fun"], c["xu"], c["xy"], c["y"], c["mu"], c["delta"])[0]) + shift
        out[nm] = c
    return out


def flatten(cases, results):
    arrays, index = {}, {}
    for nm, c in cases.items():
        index[nm] = {"cov_par": c["cov_par"], "cov_fun": c["cov_fun"], "delta": c["delta"], "knots": c["knots"],
                     "family": c.get("family", "gaussian"), "model": c.get("model"), "outputs": sorted(results[nm])}
        for k in ("xy", "y", "mu", "xu", "muu", "draw", "obj_current"):
            if k in c:
                arrays["%s/in/%s" % (nm, k)] = np.asarray(c[k], dtype=np.float64)
        for k, v in c.get("extra", {}).items():
            arrays["%s/in/extra_%s" % (nm, k)] = np.asarray(v, dtype=np.float64)
        for k, v in results[nm].items():
            arrays["%s/out/%s" % (nm, k)] = np.asarray(v, dtype=np.float64)
    return arrays, index


def generate(verbose=False):
    import time
    t0 = time.time()
    cases, results = {}, {}
    for nm, c in gauss_cases().items():
        cases[nm], results[nm] = c, run_gauss(c)
        if verbose:
            print("%-22s %.1f s" % (nm, time.time() - t0))
    for nm, c in laplace_cases().items():
        cases[nm], results[nm] = c, run_laplace(c)
        if verbose:
            print("%-22s %.1f s (%d Newton iterations)" % (nm, time.time() - t0, len(results[nm]["objective_function_values"]) - 1))
    for nm, c in fit_cases().items():
        cases[nm], results[nm] = c, run_fit(c)
        if verbose:
            print("%-22s %.1f s" % (nm, time.time() - t0))
    for nm, c in oat_cases().items():
        cases[nm], results[nm] = c, run_oat(c)
        if verbose:
            print("%-22s %.1f s" % (nm, time.time() - t0))
    for nm, src, draw in (("p_lap_bern_d3", "l_bern_ard_d3", [3, 21, 40, 55]), ("p_lap_pois_1d", "l_pois_sqexp_1d", [4, 18, 33])):
        c = dict(cases[src], draw=np.array(draw, dtype=np.float64), model="laplace")
        cases[nm], results[nm] = c, run_oat_laplace(c, results[src])
        if verbose:
            print("%-22s %.1f s" % (nm, time.time() - t0))
    return flatten(cases, results)


if __name__ == "__main__":
    arrays, index = generate(verbose=True)
    np.savez_compressed(OUT + ".npz", **arrays)
    with open(OUT + ".json", "w") as f:
        json.dump({"generator": "tests/tools/make_golden_r.py", "source": "/root/reference/R/*.R executed by oracle/mini_r; "
                   "Rcpp exports = the reference's src/*.cpp compiled unmodified (oracle/_ref)", "cases": index}, f, indent=1)
    print("wrote %s.npz (%d arrays, %d bytes)" % (OUT, len(arrays), os.path.getsize(OUT + ".npz")))
