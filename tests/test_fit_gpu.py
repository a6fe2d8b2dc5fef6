"""Optimiser loops: srgp_gauss_fit against the literal transcription of
norm_grad_ascent_vi / norm_grad_ascent (oracle/ref_model.py:norm_grad_ascent; R/vi_functions.R:596-1218,
R/laplace_gradient_ascent.R:1111-1696).  The whole trajectory is compared: objectives rel 1e-8 per iteration."""
import numpy as np
import pytest

from oracle import ref_model as rm
from sparsergps_b200 import vi_functions as vf
from tests import cases

pytestmark = pytest.mark.gpu


def _start(c):
    cp = dict(c["cov_par"])
    for k in cp:                       # start away from the generating values so the optimiser has work to do
        cp[k] *= 0.7 if k != "tau" else 1.3
    return cp


@pytest.mark.parametrize("vi", [True, False])
@pytest.mark.parametrize("method", ["adadelta", "ga"])
def test_fit_theta_trajectory_matches_reference_loop(ctx, vi, method):
    c = cases.config2(n=800, m=30)
    opt = {"optim_method": method, "maxit": 12, "obj_tol": 1e-6, "delta": c["delta"], "learn_rate": 1e-4}
    start = _start(c)
    ref = rm.norm_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], c["mu"], opt, vi=vi)
    f = vf.norm_grad_ascent_vi if vi else vf.norm_grad_ascent
    got = f(start, "ard", c["xu"], c["x"], c["y"], c["mu"], None, opt, ctx=ctx)
    assert got["iter"] == ref["iter"]
    np.testing.assert_allclose(got["obj_fun"], ref["obj_fun"], rtol=1e-8)
    np.testing.assert_allclose(got["cov_par_history"], ref["cov_par_history"], rtol=1e-8)
    np.testing.assert_allclose(got["grad"], ref["grad"], rtol=1e-6, atol=1e-8 * np.abs(ref["grad"]).max())
    for k in start:
        assert got["cov_par"][k] == pytest.approx(ref["cov_par"][k], rel=1e-8)
    np.testing.assert_array_equal(got["xu"], c["xu"])
    # the reference returns the posterior at the knots with the fit
    um, uv = rm.gauss_posterior_u(ref["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], np.zeros(30), c["delta"], vi=vi)
    np.testing.assert_allclose(got["u_mean"], um, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(got["u_var"], uv, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("vi", [True, False])
def test_fit_with_knots_matches_reference_loop(ctx, vi):
    """sqexp, 1-D, all knots optimised on the bounded-logit scale (quirk Q12 shifts them at the first step)."""
    c = cases.config1(n=400)
    opt = {"maxit": 10, "obj_tol": 1e-9, "delta": c["delta"]}
    start = {"sigma": 1.0, "l": 0.5, "tau": 0.7}
    ref = rm.norm_grad_ascent(start, "sqexp", c["xu"], c["x"], c["y"], c["mu"], opt, vi=vi, opt_knots=True)
    f = vf.norm_grad_ascent_vi if vi else vf.norm_grad_ascent
    got = f(start, "sqexp", c["xu"], c["x"], c["y"], c["mu"], None, opt, dcov_fun_dknot=True, ctx=ctx)
    assert got["iter"] == ref["iter"] == 10
    np.testing.assert_allclose(got["obj_fun"], ref["obj_fun"], rtol=1e-8)
    np.testing.assert_allclose(got["xu"], ref["xu"], rtol=1e-8)
    assert not np.array_equal(got["xu"], c["xu"])
    for k in start:
        assert got["cov_par"][k] == pytest.approx(ref["cov_par"][k], rel=1e-8)


def test_fit_knot_subset_and_stop_rule(ctx):
    c = cases.config2(n=500, m=12)
    start = _start(c)
    opt = {"maxit": 50, "obj_tol": 1.3, "delta": c["delta"]}          # loose tolerance: the stop rule ends the loop (iter 8)
    ref = rm.norm_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], c["mu"], opt, vi=True, opt_knots=True,
                              knot_opt=[2, 7])
    got = vf.norm_grad_ascent_vi(start, "ard", c["xu"], c["x"], c["y"], c["mu"], None, opt, dcov_fun_dknot=True,
                                 knot_opt=[2, 7], ctx=ctx)
    assert got["iter"] == ref["iter"] < 50
    np.testing.assert_allclose(got["obj_fun"], ref["obj_fun"], rtol=1e-8)
    np.testing.assert_allclose(got["xu"], ref["xu"], rtol=1e-8)


def test_fit_at_scale_improves_objective(ctx):
    """A size no CPU loop reaches: 20 ADADELTA iterations at n = 200k, m = 512; the ELBO must increase."""
    c = cases.config5(n=200_000, m=512)
    start = _start(c)
    got = vf.norm_grad_ascent_vi(start, "ard", c["xu"], c["x"], c["y"], 0.0, None,
                                 {"maxit": 20, "obj_tol": 1e-12, "delta": c["delta"]}, ctx=ctx)
    assert got["iter"] == 20 and got["obj_fun"][-1] > got["obj_fun"][0]
    assert np.all(np.isfinite(got["u_mean"]))


@pytest.mark.parametrize("family", ["bernoulli", "poisson"])
def test_laplace_fit_trajectory_matches_reference_loop(ctx, family):
    """laplace_grad_ascent (R/laplace_gradient_ascent.R:10-628): warm-started Newton + dlogq_dcov_par per iteration."""
    from sparsergps_b200 import laplace as Lp
    c = cases.config4(n=500, m=16)
    if family == "poisson":
        c["y"] = np.random.default_rng(3).poisson(np.exp(0.5 * np.sin(c["x"][:, 0]))).astype(np.float64)
    start = dict(c["cov_par"])
    start["sigma"] *= 0.8
    kw = {"m": 1.0} if family == "poisson" else {}
    opt = {"maxit": 6, "obj_tol": 1e-9, "delta": c["delta"], "maxit_nr": 40, "tol_nr": 1e-5}
    ref = rm.laplace_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], np.zeros(500), family, c["mu"], np.zeros(16),
                                 opt, **kw)
    got = Lp.laplace_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], np.zeros(500), family, c["mu"], np.zeros(16),
                                 opt, ctx=ctx)
    assert got["iter"] == ref["iter"] == 6
    np.testing.assert_array_equal(got["nr_iter"], ref["nr_iter"])
    np.testing.assert_allclose(got["obj_fun"], ref["obj_fun"], rtol=1e-8)
    np.testing.assert_allclose(got["cov_par_history"], ref["cov_par_history"], rtol=1e-7)
    np.testing.assert_allclose(got["fmax"], ref["fmax"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(got["u_mean"], ref["u_mean"], rtol=1e-6, atol=1e-8)


def test_laplace_fit_with_knots(ctx):
    from sparsergps_b200 import laplace as Lp
    c = cases.config4(n=400, m=10)
    start = dict(c["cov_par"])
    opt = {"maxit": 4, "obj_tol": 1e-9, "delta": c["delta"], "maxit_nr": 30, "tol_nr": 1e-5}
    ref = rm.laplace_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], np.zeros(400), "bernoulli", c["mu"], np.zeros(10),
                                 opt, opt_knots=True, knot_opt=[1, 8])
    got = Lp.laplace_grad_ascent(start, "ard", c["xu"], c["x"], c["y"], np.zeros(400), "bernoulli", c["mu"], np.zeros(10),
                                 opt, dcov_fun_dknot=True, knot_opt=[1, 8], ctx=ctx)
    assert got["iter"] == ref["iter"]
    np.testing.assert_allclose(got["obj_fun"], ref["obj_fun"], rtol=1e-8)
    np.testing.assert_allclose(got["xu"], ref["xu"], rtol=1e-7)
