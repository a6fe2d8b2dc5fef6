"""K1 / K2 parity: CUDA assembly through the C ABI vs the C oracle, rel <= 1e-10 per entry (north_star)."""
import numpy as np
import pytest

from oracle import ref_kernels as rk
from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-10          # north_star: relative 1e-10 on K entries
ATOL = 1e-300         # entries that underflow to denormals are compared absolutely


def _close(a, b):
    assert a.shape == b.shape
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("d", [1, 2, 3, 5, 8, 11])
@pytest.mark.parametrize("cov_fun", ["sqexp", "exp", "ard"])
def test_make_cov_mat_matches_oracle(ctx, d, cov_fun):
    from sparsergps_b200 import rcpp_exports as R
    rng = np.random.default_rng(100 + d)
    n1, n2 = 517, 93                      # ragged: not multiples of the 256 x 32 tile
    x, xp = rng.normal(size=(n1, d)), rng.normal(size=(n2, d))
    if cov_fun == "ard":
        ln = ["l%d" % (i + 1) for i in range(d)]
        cp = cases.ard_par(1.3, rng.uniform(0.5, 2.0, d), 0.4)
        _close(R.make_cov_mat_ardC(x, xp, cp, "ard", 1e-4, ln, ctx=ctx), rk.make_cov_mat_ardC(x, xp, cp, "ard", 1e-4, ln))
        _close(R.make_cov_mat_ardC(x, None, cp, "ard", 1e-4, ln, ctx=ctx), rk.make_cov_mat_ardC(x, None, cp, "ard", 1e-4, ln))
    else:
        cp = {"sigma": 1.3, "l": 0.9, "tau": 0.4}
        _close(R.make_cov_matC(x, xp, cp, cov_fun, 1e-4, ctx=ctx), rk.make_cov_matC(x, xp, cp, cov_fun, 1e-4))
        _close(R.make_cov_matC(x, None, cp, cov_fun, 1e-4, ctx=ctx), rk.make_cov_matC(x, None, cp, cov_fun, 1e-4))


@pytest.mark.parametrize("cov_fun", ["sqexp", "exp", "ard"])
def test_dsig_dtheta_matches_oracle(ctx, cov_fun):
    from sparsergps_b200 import rcpp_exports as R
    rng = np.random.default_rng(7)
    d, n1, n2 = 4, 300, 70
    x = rng.normal(size=(n1, d))
    xp = np.vstack([rng.normal(size=(n2 - 2, d)), x[5], x[17]])     # coincident rows (quirk Q4)
    if cov_fun == "ard":
        ln = ["l%d" % (i + 1) for i in range(d)]
        cp = cases.ard_par(0.8, [0.7, 1.0, 1.6, 2.2], 0.6)
        for par in ["sigma"] + ln + ["tau"]:
            _close(R.dsig_dtheta_ardC(x, xp, cp, "ard", par, ln, ctx=ctx), rk.dsig_dtheta_ardC(x, xp, cp, "ard", par, ln))
            _close(R.dsig_dtheta_ardC(x, None, cp, "ard", par, ln, ctx=ctx), rk.dsig_dtheta_ardC(x, None, cp, "ard", par, ln))
    else:
        cp = {"sigma": 0.8, "l": 1.4, "tau": 0.6}
        for par in ["sigma", "l", "tau"]:
            _close(R.dsig_dthetaC(x, xp, cp, cov_fun, par, ctx=ctx), rk.dsig_dthetaC(x, xp, cp, cov_fun, par))
            _close(R.dsig_dthetaC(x, None, cp, cov_fun, par, ctx=ctx), rk.dsig_dthetaC(x, None, cp, cov_fun, par))


def test_edge_cases(ctx, capfd):
    from sparsergps_b200 import rcpp_exports as R
    cp = {"sigma": 1.0, "l": 1.0, "tau": 0.5}
    x = np.array([[0.0], [1.0], [40.0], [1e3]])
    # far-apart points underflow gradually like libm
    _close(R.make_cov_matC(x, x[:2], cp, "sqexp", 0.0, ctx=ctx), rk.make_cov_matC(x, x[:2], cp, "sqexp", 0.0))
    # single row / single column
    _close(R.make_cov_matC(x[:1], x, cp, "sqexp", 0.0, ctx=ctx), rk.make_cov_matC(x[:1], x, cp, "sqexp", 0.0))
    # NA sentinel as an array, integer input coerced to double
    xi = np.array([[1, 2], [3, 4]], dtype=np.int32)
    _close(R.make_cov_matC(xi, np.array([[np.nan]]), cp, "sqexp", 1e-6, ctx=ctx), rk.make_cov_matC(xi, None, cp, "sqexp", 1e-6))
    # unknown kernel / parameter: message + 0 x 0 matrix, no exception
    assert R.make_cov_matC(x, None, cp, "matern", 0.0, ctx=ctx).shape == (0, 0)
    assert R.dsig_dthetaC(x, None, cp, "sqexp", "nope", ctx=ctx).shape == (0, 0)
    assert "invalid" in capfd.readouterr().err
    # NaN propagates silently
    xn = np.array([[0.0], [np.nan]])
    out = R.make_cov_matC(xn, x[:2], cp, "sqexp", 0.0, ctx=ctx)
    assert np.isnan(out[1]).all() and np.isfinite(out[0]).all()


def test_config_shapes(ctx):
    """BASELINE.json configs 1-3 at full size: K1 on the n x n / n x m matrices the reference builds."""
    from sparsergps_b200 import rcpp_exports as R
    c = cases.config1()
    _close(R.make_cov_matC(c["x"], None, c["cov_par"], "sqexp", c["delta"], ctx=ctx),
           rk.make_cov_matC(c["x"], None, c["cov_par"], "sqexp", c["delta"]))
    for c in (cases.config2(), cases.config3()):
        ln = ["l%d" % (i + 1) for i in range(c["x"].shape[1])]
        _close(R.make_cov_mat_ardC(c["x"], c["xu"], c["cov_par"], "ard", c["delta"], ln, ctx=ctx),
               rk.make_cov_mat_ardC(c["x"], c["xu"], c["cov_par"], "ard", c["delta"], ln))
        for par in ("l2", "tau"):
            _close(R.dsig_dtheta_ardC(c["x"], c["xu"], c["cov_par"], "ard", par, ln, ctx=ctx),
                   rk.dsig_dtheta_ardC(c["x"], c["xu"], c["cov_par"], "ard", par, ln))


def test_large_properties(ctx):
    """Size-independent properties at a size the oracle would not finish: K(x,x) diag, symmetry, dsigma = 2K."""
    from sparsergps_b200 import rcpp_exports as R
    rng = np.random.default_rng(3)
    n, d = 6000, 8
    x = rng.normal(size=(n, d))
    ln = ["l%d" % (i + 1) for i in range(d)]
    cp = cases.ard_par(1.0, [0.8 + 0.05 * (c + 1) for c in range(d)], 0.5)
    S = R.make_cov_mat_ardC(x, None, cp, "ard", 1e-6, ln, ctx=ctx)
    np.testing.assert_allclose(np.diag(S), 1.0 + 0.25 + 1e-6, rtol=1e-15)
    assert np.array_equal(S, S.T)
    dS = R.dsig_dtheta_ardC(x, None, cp, "ard", "sigma", ln, ctx=ctx)
    K = S - (0.25 + 1e-6) * np.eye(n)
    np.testing.assert_allclose(dS[np.triu_indices(n, 1)], 2 * K[np.triu_indices(n, 1)], rtol=1e-15)
    sub = rng.choice(n, 40, replace=False)
    _close(S[np.ix_(sub, sub)], rk.make_cov_mat_ardC(x[sub], None, cp, "ard", 1e-6, ln))
