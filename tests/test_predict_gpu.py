"""posterior at the knots and prediction against the literal oracle."""
import numpy as np
import pytest

from oracle import ref_model as rm
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vi", [True, False])
@pytest.mark.parametrize("case", ["config1", "config2", "config5"])
def test_posterior_at_knots(ctx, case, vi):
    from sparsergps_b200 import vi_functions as V
    c = {"config1": cases.config1, "config2": cases.config2, "config5": lambda: cases.config5(n=3000, m=200)}[case]()
    m = len(c["xu"])
    muu = np.linspace(-0.2, 0.3, m)
    um_ref, uv_ref = rm.gauss_posterior_u(c["cov_par"], c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], muu, c["delta"], vi=vi)
    um, uv = V.gauss_posterior_u(c["cov_par"], c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], muu, c["delta"], vi=vi, ctx=ctx)
    np.testing.assert_allclose(um, um_ref, rtol=1e-8, atol=1e-9 * np.abs(um_ref).max())
    np.testing.assert_allclose(uv, uv_ref, rtol=1e-7, atol=1e-9 * np.abs(uv_ref).max())


@pytest.mark.parametrize("case", ["config1", "config2", "config5"])
def test_predict_vi_and_laplace(ctx, case):
    from sparsergps_b200 import laplace as Lp
    from sparsergps_b200 import vi_functions as V
    c = {"config1": cases.config1, "config2": cases.config2, "config5": lambda: cases.config5(n=3000, m=200)}[case]()
    cp, m = c["cov_par"], len(c["xu"])
    rng = np.random.default_rng(4)
    x_pred = c["x"][rng.choice(len(c["x"]), 777)] + 0.1 * rng.normal(size=(777, c["x"].shape[1]))
    mu_p, muu = 0.1 * np.ones(777), np.zeros(m)
    um, uv = rm.gauss_posterior_u(cp, c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], muu, c["delta"], vi=True)
    ref = rm.predict_vi(um, uv, c["xu"], x_pred, c["cov_fun"], cp, mu_p, muu, c["delta"])
    got = V.predict_vi(um, uv, c["xu"], x_pred, c["cov_fun"], cp, mu_p, muu, delta=c["delta"], ctx=ctx)
    np.testing.assert_allclose(got["pred_mean"], ref[0], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(got["pred_var"], ref[1], rtol=1e-8, atol=1e-10)
    for fam in ("gaussian", "bernoulli"):
        ref = rm.predict_laplace(um, uv, c["xu"], x_pred, c["cov_fun"], cp, mu_p, muu, fam, c["delta"])
        got = Lp.predict_laplace(um, uv, c["xu"], x_pred, c["cov_fun"], cp, mu_p, muu, family=fam, delta=c["delta"], ctx=ctx)
        np.testing.assert_allclose(got["pred_mean"], ref[0], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(got["pred_var"], ref[1], rtol=1e-8, atol=1e-10)
    assert V.predict_vi(um, uv, c["xu"], x_pred, c["cov_fun"], cp, mu_p, muu, family="poisson", ctx=ctx).startswith("Error")
