"""The CUDA path (through the C ABI) against what the reference's OWN R code returns.

tests/golden/r_level.* = outputs of /root/reference/R/*.R executed unmodified (tests/tools/make_golden_r.py; the GPU
box has no /root/reference and needs none).  Tolerances are BASELINE.json's: relative 1e-8 on objectives and on every
gradient component (scaled by the largest component), 1e-10 on K entries; iterated quantities (Newton modes, optimiser
trajectories) 1e-7."""
import numpy as np
import pytest

from sparsergps_b200 import laplace as Lp
from sparsergps_b200 import rcpp_exports as R
from sparsergps_b200 import vi_functions as vf
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

G = gu.load_r_level()
RTOL = 1e-8


def _cases(prefix):
    return sorted(k for k in G if k.startswith(prefix))


def _close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.max(np.abs(b))), err_msg=what)


def _lnames(cp):
    return [k for k in cp if k.startswith("l")]


@pytest.mark.parametrize("name", _cases("g_"))
def test_gaussian_rows_match_reference_r(ctx, name):
    c = G[name]
    cp, cf, delta, i, o = c["meta"]["cov_par"], c["meta"]["cov_fun"], c["meta"]["delta"], c["in"], c["out"]
    # rows a1-a5 as the R callers use them (Sigma22 = self covariance - tau^2 I)
    if cf == "ard":
        S12 = R.make_cov_mat_ardC(i["xy"], i["xu"], cp, cf, delta, _lnames(cp), ctx=ctx)
        S22 = R.make_cov_mat_ardC(i["xu"], None, cp, cf, delta, _lnames(cp), ctx=ctx) - cp["tau"] ** 2 * np.eye(len(i["xu"]))
    else:
        S12 = R.make_cov_matC(i["xy"], i["xu"], cp, cf, delta, ctx=ctx)
        S22 = R.make_cov_matC(i["xu"], None, cp, cf, delta, ctx=ctx) - cp["tau"] ** 2 * np.eye(len(i["xu"]))
    np.testing.assert_allclose(S12, o["Sigma12"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(S22, o["Sigma22"], rtol=1e-10, atol=1e-14)
    # rows a14, a15
    tt = vf.trace_term_fun(cp, o["Sigma12"], o["Sigma22"], delta, ctx=ctx)
    _close(tt, o["trace_term"][0], what="trace_term_fun")
    _close(vf.dtrace_term_dtau(cp, tt), o["dtrace_term_dtau"][0])
    # rows a17-a20 (+ f1): fused objective + gradient (+ knot gradient)
    dkn = True if c["meta"]["knots"] else None
    for tag, fn, okey in (("vi", vf.delbo_dcov_par, "elbo"), ("fic", vf.dlogp_dcov_par, "obj_fun_norm")):
        g = fn(cp, cf, i["xu"], i["xy"], i["y"], i["mu"], delta, ctx=ctx, dcov_fun_dknot=dkn)
        _close(g["objective"], o[okey][0], what=okey)
        _close([g["gradient"][k] for k in cp], o[tag + "_gradient"], what=tag + " gradient")
        _close([g["trans_par"][k] for k in cp], o[tag + "_trans_par"], 1e-14)
        if dkn:
            _close(g["knot_gradient"], o[tag + "_knot_gradient"], what=tag + " knot gradient")
            _close(g["trans_knot"], o[tag + "_trans_knot"], 1e-12)


@pytest.mark.parametrize("name", _cases("l_"))
def test_laplace_rows_match_reference_r(ctx, name):
    c = G[name]
    cp, cf, delta, fam, i, o = c["meta"]["cov_par"], c["meta"]["cov_fun"], c["meta"]["delta"], c["meta"]["family"], c["in"], c["out"]
    ex = gu.r_case_extra(c)
    m_off = float(ex["m"][0]) if "m" in ex else 1.0
    nr = Lp.newtrap_sparseGP(i["mu"].copy(), fam, cp, cf, i["xy"], i["xu"], i["y"], i["mu"], i["muu"], maxit=1000, tol=1e-6,
                             delta=delta, m=m_off, ctx=ctx)
    h, h_ref = nr["objective_function_values"], o["objective_function_values"]
    assert len(h) == len(h_ref)                                   # same stopping decision as the reference's loop
    _close(h, h_ref, what="Newton objective history")
    _close(nr["gp"], o["gp"], 1e-7, "mode")
    _close(nr["u_posterior_mean"], o["u_posterior_mean"], 1e-7)
    _close(nr["u_posterior_variance"], o["u_posterior_variance"], 1e-6)
    dkn = True if c["meta"]["knots"] else None
    g = Lp.dlogq_dcov_par(cp, cf, i["xu"], i["xy"], i["y"], o["gp"], fam, i["mu"], delta, m=m_off, ctx=ctx, dcov_fun_dknot=dkn)
    _close([g["gradient"][k] for k in cp], o["gradient"], what="dlogq gradient")
    if dkn:
        _close(g["knot_gradient"], o["knot_gradient"], what="dlogq knot gradient")
    pr = Lp.predict_laplace(o["u_posterior_mean"], o["u_posterior_variance"], i["xu"], o["x_pred"], cf, cp,
                            np.full(len(o["x_pred"]), i["mu"][0]), i["muu"], family=fam, delta=delta, ctx=ctx)
    _close(pr["pred_mean"], o["pred_mean"], what="pred_mean")
    _close(pr["pred_var"], o["pred_var"], 1e-7, "pred_var")


@pytest.mark.parametrize("name", _cases("f_"))
def test_optimiser_loops_match_reference_r(ctx, name):
    c = G[name]
    meta, i, o = c["meta"], c["in"], c["out"]
    cp, cf, delta, fam, model, knots = meta["cov_par"], meta["cov_fun"], meta["delta"], meta["family"], meta["model"], meta["knots"]
    opt = {"maxit": int(o["iter"][0]), "delta": delta, "obj_tol": 0.0}
    dkn = True if knots else None
    if model == "laplace":
        ex = gu.r_case_extra(c)
        r = Lp.laplace_grad_ascent(cp, cf, i["xu"], i["xy"], i["y"], i["mu"].copy(), fam, i["mu"], i["muu"], opt,
                                   m=float(ex["m"][0]) if "m" in ex else 1.0, dcov_fun_dknot=dkn, ctx=ctx)
    else:
        f = vf.norm_grad_ascent_vi if model == "vi" else vf.norm_grad_ascent
        r = f(cp, cf, i["xu"], i["xy"], i["y"], i["mu"], i["muu"], opt, dcov_fun_dknot=dkn, ctx=ctx)
    assert r["iter"] == int(o["iter"][0])
    _close(r["obj_fun"], o["obj_fun"], 1e-7, "objective trajectory")
    _close(r["cov_par_history"], o["cov_par_history"], 1e-7, "parameter trajectory")
    _close(r["grad"], o["grad"], 1e-6, "gradient trajectory")
    _close(r["xu"], o["xu_final"], 1e-7, "final knots")
    _close([r["cov_par"][k] for k in cp], o["cov_par"], 1e-7)
    _close(r["u_mean"], o["u_mean"], 1e-6, "u_mean")
    _close(r["u_var"], o["u_var"], 1e-5, "u_var")
    cpf = dict(zip(cp, o["cov_par"].tolist()))
    mu_p = np.full(len(o["x_pred"]), i["mu"][0])
    if model == "vi":
        pr = vf.predict_vi(o["u_mean"], o["u_var"], o["xu_final"], o["x_pred"], cf, cpf, mu_p, i["muu"], delta=delta, ctx=ctx)
    else:
        pr = Lp.predict_laplace(o["u_mean"], o["u_var"], o["xu_final"], o["x_pred"], cf, cpf, mu_p, i["muu"], family=fam,
                                delta=delta, ctx=ctx)
    _close(pr["pred_mean"], o["pred_mean"], what="pred_mean")
    _close(pr["pred_var"], o["pred_var"], 1e-7, "pred_var")
