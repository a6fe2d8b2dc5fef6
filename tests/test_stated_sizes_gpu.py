"""Parity at the sizes BASELINE.json states ("next round" item 1), against committed goldens.

tests/golden/stated_sizes.json is produced in the build container by tests/tools/make_golden_sizes.py from the CPU
oracle (literal transcription where it fits in memory, the reduced form -- proven equal to it in tests/test_oracle.py
-- in row shards where it does not).  Inputs are regenerated here from the same seeds (tests/cases.py,
bench.workload); only outputs are stored.  Tolerances: north_star's rel 1e-8 on the objective and every gradient
component (config 3: 1e-7 against the long-double yardstick, see test_vi_gpu.py).
"""
import json
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-8
ILL_RTOL = 1e-7
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stated_sizes.json")


def _book():
    with open(GOLDEN) as f:
        return json.load(f)


def _check_grad(grad, ref, names, rtol):
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    for k, nm in enumerate(names):
        assert grad[k] == pytest.approx(ref[k], rel=rtol, abs=rtol * 1e-3 * scale), nm


def _check_knots(kg, summ, rtol):
    kg = np.asarray(kg, dtype=np.float64).reshape(summ["shape"])
    fro = summ["fro"]
    assert float(np.linalg.norm(kg)) == pytest.approx(fro, rel=rtol)
    assert float(np.abs(kg).sum()) == pytest.approx(summ["abs_sum"], rel=rtol)
    assert float(kg.sum()) == pytest.approx(summ["sum"], abs=rtol * summ["abs_sum"])
    for (k, c), v in zip(summ["probe_idx"], summ["probe"]):
        assert kg[k, c] == pytest.approx(v, rel=rtol, abs=rtol * 1e-3 * np.abs(kg).max()), (k, c)


def test_config5_vi_full_size(ctx):
    """n = 1,000,000, m = 1024, d = 8: objective, the 10 gradient components and the 1024 x 8 knot gradient."""
    import bench
    from sparsergps_b200.vi_functions import knot_bounds
    g = _book()["cfg5_vi"]
    x, y, xu, th = bench.workload(g["n"], g["m"], g["d"])
    ctx.set_data(x, y, None)
    obj, grad = ctx.gauss_obj_grad("vi", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"])
    assert obj == pytest.approx(g["obj"], rel=RTOL)
    _check_grad(grad, g["grad"], g["names"], RTOL)
    obj2, grad2, kg, _ = ctx.gauss_obj_grad_knots("vi", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"], knot_bounds(x))
    assert obj2 == pytest.approx(g["obj"], rel=RTOL)
    _check_grad(grad2, g["grad"], g["names"], RTOL)
    _check_knots(kg, g["knot"], RTOL)


def test_config5_fic_quarter_size(ctx):
    """FIC at n = 250,000, m = 1024, d = 8 (the reduced NumPy form needs 4 row passes; 250k keeps the generator at minutes)."""
    import bench
    from sparsergps_b200.vi_functions import knot_bounds
    g = _book()["cfg5_fic"]
    x, y, xu, th = bench.workload(g["n"], g["m"], g["d"])
    ctx.set_data(x, y, None)
    obj, grad, kg, _ = ctx.gauss_obj_grad_knots("fic", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"], knot_bounds(x))
    assert obj == pytest.approx(g["obj"], rel=RTOL)
    _check_grad(grad, g["grad"], g["names"], RTOL)
    _check_knots(kg, g["knot"], RTOL)


def test_config4_newton_and_gradient_full_size(ctx):
    """Bernoulli, n = 100,000, d = 8, m = 512: same Newton iteration count and objective history to tol 1e-5, the mode,
    the posterior at the knots, and dlogq_dcov_par at a closed-form ff and at the mode."""
    from sparsergps_b200 import laplace as Lp
    from tests.tools.make_golden_sizes import cfg4_ff_closed_form
    g = _book()["cfg4"]
    c = cases.config4(n=g["n"], d=g["d"], m=g["m"])
    cp = c["cov_par"]
    fit = Lp.newtrap_sparseGP(np.zeros(g["n"]), "bernoulli", cp, "ard", c["x"], c["xu"], c["y"], c["mu"], np.zeros(g["m"]),
                              maxit=g["maxit"], tol=g["tol"], delta=c["delta"], ctx=ctx)
    h = fit["objective_function_values"]
    # The search ends when an objective step falls below tol = 1e-5, and over its last ~15 iterations the steps hover at
    # 1.0e-5 .. 1.3e-5 on a value of 7e4 (golden tail: 1.14e-5, 1.04e-5, 1.14e-5, 9.0e-6): the stopping iteration is decided
    # by rounding at 1e-11 relative -- 422, 425 and 426 iterations have all been observed for bit-different but equally
    # accurate evaluations.  The count is therefore held to +-8 and the history to 1e-8 on the common prefix (the small
    # cases of test_laplace_gpu.py / test_golden_r_gpu.py, which stop on a clear step, keep the exact-count assertion).
    assert abs(len(h) - g["iterations"]) <= 8
    k = min(len(h), g["iterations"])
    np.testing.assert_allclose(h[:k], g["hist"][:k], rtol=RTOL)
    # one Newton step near the stopping point moves the mode by ~1e-8 (absolute, entries of order 1), the posterior mean
    # and variance at the knots by ~1e-5 of their largest entry: the absolute tolerances below grow with the number of
    # steps the two searches differ by
    dsteps = 1 + abs(len(h) - g["iterations"])
    np.testing.assert_allclose(fit["gp"][:64], g["gp_head"], rtol=1e-7, atol=1e-7 * dsteps)
    assert float(np.linalg.norm(fit["gp"])) == pytest.approx(g["gp_norm"], rel=1e-8 * dsteps)
    # u_mean = muu + a - G_Z C_Z a cancels ~6 digits at this size (|a| ~ n / Z): the reference's own float64 formula is
    # 9e-7 (relative to max |u_mean|) from a long-double evaluation of the same expression, a Cholesky-solve form
    # 3e-10 (tests/tools/make_golden_sizes.py header; measured at the closed-form ff).  The golden holds the reference's
    # formula, so the comparison is stated at 1e-5, not at the 1e-7 of the small cases.
    um_ref = np.asarray(g["u_mean"])
    np.testing.assert_allclose(fit["u_posterior_mean"], um_ref, rtol=0, atol=1e-5 * dsteps * np.abs(um_ref).max())
    # the posterior variance moves by ~1e-5 (absolute) per Newton step near the stopping point, and the stopping
    # iteration itself is decided by rounding (above): held to 2e-5 of its largest entry
    uv_ref = np.asarray(g["u_var_diag"])
    np.testing.assert_allclose(np.diag(fit["u_posterior_variance"]), uv_ref, rtol=1e-5, atol=2e-5 * dsteps * np.abs(uv_ref).max())
    ff = cfg4_ff_closed_form(c["x"])
    got = Lp.dlogq_dcov_par(cp, "ard", c["xu"], c["x"], c["y"], ff, "bernoulli", c["mu"], c["delta"], ctx=ctx)["gradient"]
    # n / Z ~ 1e7 in the row sums: at this size the float64 transcription of the reference is itself 1.7e-8 away from a
    # long-double evaluation of the same gradient, so 1e-8 agreement WITH THE REFERENCE is not defined here.  The product
    # is held to 5e-8 against the float64 oracle and to 2e-8 against the long-double yardstick (measured 1.1e-8 with the
    # 7-slice INT8 passes and the single-precision low level group; 0.9e-8 with that group in double precision).
    _check_grad([got[k] for k in g["names"]], g["grad_at_closed_form_ff"], g["names"], 5e-8)
    if "grad_longdouble_at_closed_form_ff" in g:
        _check_grad([got[k] for k in g["names"]], g["grad_longdouble_at_closed_form_ff"], g["names"], 2e-8)
    # at the mode the two Newton runs agree to ~1e-7 in ff, so the gradient is held to 1e-6
    got = Lp.dlogq_dcov_par(cp, "ard", c["xu"], c["x"], c["y"], fit["gp"], "bernoulli", c["mu"], c["delta"], ctx=ctx)["gradient"]
    _check_grad([got[k] for k in g["names"]], g["grad_at_mode"], g["names"], 1e-6)


@pytest.mark.parametrize("n", [9568, 4784])
@pytest.mark.parametrize("model", ["vi", "fic"])
def test_config3_full_size(ctx, model, n):
    """ccpp-shaped OAT step at its stated size (d = 4, m = 256 + one knot that IS a data row, delta = 1e-3): cond(Sigma22)
    ~ 4e4, so the yardstick is the long-double evaluation and the tolerance 1e-7 (the float64 literal transcription of
    the reference is itself 1e-8 .. 1.5e-7 away from it; at n = 4784 its objective is -inf because det(Sigma22)
    underflows -- quirk Q6 -- while Cholesky log-determinants stay finite)."""
    g = _book()["cfg3_%d" % n]
    c = cases.config3(n=n)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], c["mu"])
    obj, grad = ctx.gauss_obj_grad(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
    ref = g[model]
    assert obj == pytest.approx(ref["longdouble_obj"], rel=ILL_RTOL)
    _check_grad(grad, ref["longdouble_grad"], g["names"], ILL_RTOL)
    # the literal float64 transcription must sit inside a comparable band around the same yardstick
    lit_err = max(abs(a - b) / abs(b) for a, b in zip(ref["literal_grad"], ref["longdouble_grad"]))
    got_err = max(abs(a - b) / abs(b) for a, b in zip(grad, ref["longdouble_grad"]))
    assert got_err <= max(10 * lit_err, ILL_RTOL)
    if np.isfinite(ref["literal_obj"]):
        assert obj == pytest.approx(ref["literal_obj"], rel=ILL_RTOL)
