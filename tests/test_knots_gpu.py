"""Knot-location gradient: delbo_dcov_par / dlogp_dcov_par with a dcov_fun_dknot
(R/vi_functions.R:425-592, R/laplace_approx_gradient.R:965-1126) through the C ABI, against the literal oracle
(selected knots -- its loop is O(m d n m^2)) and the reduced oracle (all knots).  rel <= 1e-8 (north_star)."""
import numpy as np
import pytest

from oracle import reduced_model as red
from oracle import ref_model as rm
from sparsergps_b200 import vi_functions as vf
from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * 1e-3 * scale)


CASES = {
    "config1": (lambda: cases.config1(), None),
    "config2": (lambda: cases.config2(), [0, 7, 21, 40, 63]),
    "config5": (lambda: cases.config5(n=3000, m=300), [3, 150, 299]),
}


@pytest.mark.parametrize("model", ["vi", "fic"])
@pytest.mark.parametrize("case", list(CASES))
def test_knot_gradient_matches_literal_oracle(ctx, model, case):
    make, knot_opt = CASES[case]
    c = make()
    cp, cf = c["cov_par"], c["cov_fun"]
    ours = (vf.delbo_dcov_par if model == "vi" else vf.dlogp_dcov_par)(
        cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"], ctx=ctx, dcov_fun_dknot=True, knot_opt=knot_opt)
    lit = (rm.delbo_dcov_par if model == "vi" else rm.dlogp_dcov_par)(
        cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"], dcov_fun_dknot=rm.dcov_fun_dknot_for(cf),
        knot_opt=knot_opt)
    m, d = c["xu"].shape
    g, g_ref = ours["knot_gradient"].reshape(m, d), lit["knot_gradient"].reshape(m, d)
    sel = list(range(m)) if knot_opt is None else knot_opt
    _close(g[sel], g_ref[sel])
    rest = [k for k in range(m) if k not in sel]
    assert not g[rest].any()                                 # knots outside knot_opt: exactly 0 like the reference
    np.testing.assert_allclose(ours["trans_knot"], lit["trans_knot"], rtol=1e-13, atol=1e-13)
    for nm in cp:                                            # the theta gradient of the same call is unchanged
        assert ours["gradient"][nm] == pytest.approx(lit["gradient"][nm], rel=RTOL,
                                                     abs=RTOL * 1e-3 * max(abs(v) for v in lit["gradient"].values()))


@pytest.mark.parametrize("model", ["vi", "fic"])
def test_knot_gradient_without_transform(ctx, model):
    c = cases.config2(n=700, m=24)
    cp, cf = c["cov_par"], c["cov_fun"]
    opt = [1, 5, 23]
    ours = (vf.delbo_dcov_par if model == "vi" else vf.dlogp_dcov_par)(
        cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"], ctx=ctx, dcov_fun_dknot=True, knot_opt=opt,
        transform=False)
    lit = (rm.delbo_dcov_par if model == "vi" else rm.dlogp_dcov_par)(
        cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"], dcov_fun_dknot=rm.dcov_fun_dknot_for(cf),
        knot_opt=opt, transform=False)
    m, d = c["xu"].shape
    _close(ours["knot_gradient"].reshape(m, d)[opt], lit["knot_gradient"].reshape(m, d)[opt])
    np.testing.assert_array_equal(ours["trans_knot"], c["xu"])


@pytest.mark.parametrize("model", ["vi", "fic"])
@pytest.mark.parametrize("shape", [(20000, 1024, 1312), (11003, 130, 77)])
def test_knot_gradient_all_knots_against_reduced_oracle(ctx, model, shape):
    """Headline knot count (m = 1024, d = 8) and a ragged several-chunk shape; every knot, both models."""
    n, m, seed = shape
    c = cases.config5(n=n, m=m, seed=seed)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    kb = vf.knot_bounds(c["x"])
    obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"], kb)
    f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
    obj_ref, g_ref, kg_ref = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"],
                               knots=True)
    assert obj == pytest.approx(obj_ref, rel=RTOL)
    _close(grad, [g_ref[k] for k in cp])
    # m*d = 8192 entries spanning orders of magnitude: the gradient VECTOR is held to 1e-8 of its largest entry
    # (tests/tools/dbg_knots.py: 6e-11 (VI) / 2e-10 (FIC) here, and closer to a long-double evaluation than this NumPy
    # yardstick wherever long double is affordable)
    kg = kg.reshape(m, -1)
    assert np.abs(kg - kg_ref).max() <= RTOL * np.abs(kg_ref).max()
    big = np.abs(kg_ref) > 1e-2 * np.abs(kg_ref).max()
    np.testing.assert_allclose(kg[big], kg_ref[big], rtol=RTOL)
    # and the plain evaluation still gives the same objective / gradient (the knot epilogue is a separate instantiation)
    obj2, grad2 = ctx.gauss_obj_grad(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
    assert obj2 == pytest.approx(obj, rel=1e-13)     # different epilogues, different summation order
    np.testing.assert_allclose(grad2, grad, rtol=1e-10, atol=1e-12 * np.abs(grad).max())


def test_knot_gradient_finite_difference_at_scale(ctx):
    """Size-independent property: d objective / d u_kc equals the central difference, at a size no oracle reaches."""
    c = cases.config5(n=200_000, m=512)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    l = cases.lvec(cp)
    for model in ("vi", "fic"):
        _, _, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", c["xu"], cp["sigma"], l, cp["tau"], c["delta"], None)
        kg = kg.reshape(512, 8)
        h = 1e-5
        for k, cc in [(0, 0), (100, 3), (511, 7)]:
            def f(sign):
                xu = c["xu"].copy()
                xu[k, cc] += sign * h
                return ctx.gauss_obj_grad(model, "ard", xu, cp["sigma"], l, cp["tau"], c["delta"], want_grad=False)[0]
            fd = (f(+1) - f(-1)) / (2 * h)
            assert kg[k, cc] == pytest.approx(fd, rel=2e-5, abs=2e-5 * np.abs(kg).max()), (model, k, cc)


def test_knot_gradient_argument_errors(ctx):
    from sparsergps_b200 import _lib as L
    c = cases.config1()
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    with pytest.raises(Exception):
        ctx.gauss_obj_grad_knots("vi", "sqexp", c["xu"], cp["sigma"], [cp["l"]], cp["tau"], c["delta"], None, [7])
    with pytest.raises(Exception):
        ctx.gauss_obj_grad_knots("vi", "exp", c["xu"], cp["sigma"], [cp["l"]], cp["tau"], c["delta"], None, None)
