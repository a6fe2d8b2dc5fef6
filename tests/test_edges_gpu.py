"""Edge cases of the fused paths: tiny / ragged sizes, generic-d code path (d > 8), error behaviour."""
import numpy as np
import pytest

from oracle import ref_model as rm
from tests import cases
from tests.test_vi_gpu import _check

pytestmark = pytest.mark.gpu


def _problem(n, m, d, seed, cov_fun="ard"):
    rng = np.random.default_rng(seed)
    spread = max(1.0, 0.3 * m ** (1.0 / d))      # keep the knot spacing near the length scale: cond(S) moderate
    x, xu = spread * rng.normal(size=(n, d)), spread * rng.normal(size=(m, d))
    y = np.sin(x[:, 0]) + 0.3 * rng.normal(size=n)
    if cov_fun == "ard":
        cp = cases.ard_par(1.1, rng.uniform(0.7, 1.6, d), 0.4)
    else:
        cp = {"sigma": 1.1, "l": 1.2, "tau": 0.4}
    return x, y, xu, cp


@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (2, 3, 2), (5, 1, 3), (129, 129, 1), (300, 7, 11), (1000, 40, 16)])
@pytest.mark.parametrize("model", ["vi", "fic"])
def test_tiny_ragged_and_generic_d(ctx, n, m, d, model):
    x, y, xu, cp = _problem(n, m, d, 10 * n + m + d)
    ctx.set_data(x, y, None)
    obj, grad = ctx.gauss_obj_grad(model, "ard", xu, cp["sigma"], cases.lvec(cp), cp["tau"], 1e-4)
    f = rm.vi_obj_grad if model == "vi" else rm.fic_obj_grad
    obj_ref, g_ref = f(cp, "ard", xu, x, y, np.zeros(n), 1e-4)
    if not np.isfinite(obj_ref):
        # quirk Q6: the reference's log(det(Sigma22)) under/overflows (129 knots on a line); the product takes
        # log|S| from the Cholesky factor, as the reduced-form oracle does -- the gradient is unaffected
        from oracle import reduced_model as red
        fr = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
        obj_ref = fr(x, y, np.zeros(n), xu, cp["sigma"], cases.lvec(cp), cp["tau"], 1e-4)[0]
    _check(obj, grad, obj_ref, g_ref, list(cp))


@pytest.mark.parametrize("model", ["vi", "fic"])
def test_sqexp_multi_dim(ctx, model):
    x, y, xu, cp = _problem(800, 30, 3, 5, cov_fun="sqexp")
    ctx.set_data(x, y, None)
    obj, grad = ctx.gauss_obj_grad(model, "sqexp", xu, cp["sigma"], [cp["l"]], cp["tau"], 1e-5)
    f = rm.vi_obj_grad if model == "vi" else rm.fic_obj_grad
    obj_ref, g_ref = f(cp, "sqexp", xu, x, y, np.zeros(800), 1e-5)
    _check(obj, grad, obj_ref, g_ref, ["sigma", "l", "tau"])


def test_laplace_generic_d_and_tiny(ctx):
    from sparsergps_b200 import laplace as Lp
    for n, m, d in ((40, 3, 2), (400, 20, 11)):
        rng = np.random.default_rng(n)
        x, xu = rng.normal(size=(n, d)), rng.normal(size=(m, d))
        y = (rng.uniform(size=n) < 0.5).astype(np.float64)
        cp = cases.ard_par(1.5, [1.3] * d, 0.2)
        ref = rm.newtrap_sparseGP(np.zeros(n), "bernoulli", cp, "ard", x, xu, y, np.zeros(n), np.zeros(m), maxit=12, tol=1e-5, delta=1e-3)
        got = Lp.newtrap_sparseGP(np.zeros(n), "bernoulli", cp, "ard", x, xu, y, np.zeros(n), np.zeros(m), maxit=12, tol=1e-5, delta=1e-3, ctx=ctx)
        np.testing.assert_allclose(got["objective_function_values"], ref["objective_function_values"], rtol=1e-8)
        g_ref = rm.dlogq_dcov_par(cp, "ard", xu, x, y, ref["gp"], "bernoulli", np.zeros(n), 1e-3)["gradient"]
        g = Lp.dlogq_dcov_par(cp, "ard", xu, x, y, ref["gp"], "bernoulli", np.zeros(n), 1e-3, ctx=ctx)["gradient"]
        scale = max(abs(v) for v in g_ref.values())
        for k in g_ref:
            assert g[k] == pytest.approx(g_ref[k], rel=1e-8, abs=1e-9 * scale), k


def test_error_behaviour(ctx):
    from sparsergps_b200 import _lib as L
    from sparsergps_b200.context import Context
    fresh = Context(0)
    xu = np.zeros((3, 2), order="F")
    l = np.ones(2)
    obj = L.cd()
    st = fresh._lib.srgp_gauss_obj_grad(fresh.handle, L.VI, L.ARD, L.ptr(xu), 3, 1.0, L.ptr(l), 0.5, 1e-6, obj, None)
    assert st == L.ERR_STATE and b"before srgp_set_data" in fresh._lib.srgp_last_error()
    fresh.set_data(np.zeros((4, 2)), np.zeros(4), None)
    st = fresh._lib.srgp_gauss_obj_grad(fresh.handle, L.VI, L.EXP, L.ptr(xu), 3, 1.0, L.ptr(l), 0.5, 1e-6, obj, None)
    assert st == L.ERR_UNKNOWN_KERNEL
    st = fresh._lib.srgp_gauss_obj_grad(fresh.handle, 7, L.ARD, L.ptr(xu), 3, 1.0, L.ptr(l), 0.5, 1e-6, obj, None)
    assert st == L.ERR_ARG
    st = fresh._lib.srgp_make_cov_mat(fresh.handle, 9, L.ptr(xu), 3, None, 0, 2, 1.0, L.ptr(l), 0.5, 0.0, L.ptr(np.zeros((3, 3), order="F")))
    assert st == L.ERR_UNKNOWN_KERNEL
    fresh.close()


@pytest.mark.parametrize("model", ["vi", "fic"])
def test_dimension_limits_of_the_fused_passes(ctx, model):
    """include/srgp.h: SRGP_MAX_D_FUSED = 20 (12 with the knot gradient); beyond it SRGP_ERR_ARG, never a wrong answer."""
    from oracle import reduced_model as red
    from sparsergps_b200 import _lib as L
    n, m = 700, 40
    for d, knots, ok in [(20, False, True), (12, True, True), (21, False, False), (13, True, False)]:
        x, y, xu, cp = _problem(n, m, d, 900 + d)
        ctx.set_data(x, y, None)
        args = ("ard", xu, cp["sigma"], cases.lvec(cp), cp["tau"], 1e-4)
        if not ok:
            with pytest.raises(RuntimeError, match="shared memory"):
                ctx.gauss_obj_grad_knots(model, *args, red.knot_bounds(x)) if knots else ctx.gauss_obj_grad(model, *args)
            continue
        f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
        ref = f(x, y, np.zeros(n), xu, cp["sigma"], cases.lvec(cp), cp["tau"], 1e-4, knots=knots)
        if knots:
            obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, *args, red.knot_bounds(x))
            np.testing.assert_allclose(kg.reshape(m, d), ref[2], rtol=1e-7, atol=1e-9 * np.abs(ref[2]).max())
        else:
            obj, grad = ctx.gauss_obj_grad(model, *args)
        _check(obj, grad, ref[0], ref[1], list(cp))


def _gridded(n, n_knots, repeat=1, seed=41):
    """d = 1 integer-valued inputs with knots ON the grid: every data row is bit-identical to `repeat` knots."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, n_knots, size=n).astype(np.float64).reshape(-1, 1)
    y = np.sin(0.7 * x[:, 0]) + 0.3 * rng.normal(size=n)
    xu = np.tile(np.arange(n_knots, dtype=np.float64), repeat).reshape(-1, 1)
    return x, y, xu, {"sigma": 1.3, "l": 1.1, "tau": 0.4}


@pytest.mark.parametrize("model", ["vi", "fic"])
def test_more_coincident_pairs_than_the_old_fixed_list(ctx, model):
    """Quirk Q4 on gridded inputs: 200,000 bit-identical (row, knot) pairs -- three times the 65,536 the
    round-1 list silently truncated at -- must all enter the tau gradient, and the sum must not depend on the order in
    which the atomics filled the list (two evaluations agree bit for bit)."""
    from oracle import reduced_model as red
    n = 200_000
    x, y, xu, cp = _gridded(n, 48)
    ctx.set_data(x, y, None)
    obj, grad = ctx.gauss_obj_grad(model, "sqexp", xu, cp["sigma"], [cp["l"]], cp["tau"], 1e-3)
    f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
    obj_ref, g_ref = f(x, y, np.zeros(n), xu, cp["sigma"], [cp["l"]], cp["tau"], 1e-3, cov_fun="sqexp")
    _check(obj, grad, obj_ref, g_ref, ["sigma", "l", "tau"])
    obj2, grad2 = ctx.gauss_obj_grad(model, "sqexp", xu, cp["sigma"], [cp["l"]], cp["tau"], 1e-3)
    assert obj2 == obj and np.array_equal(grad2, grad)


def test_pair_list_overflow_is_an_error_not_a_truncation(ctx):
    """Each row coincides with THREE duplicated knots: 3 n pairs > n + 65,536 list entries -> SRGP_ERR_STATE."""
    from sparsergps_b200 import _lib as L
    n = 100_000
    x, y, xu, cp = _gridded(n, 16, repeat=3)
    ctx.set_data(x, y, None)
    with pytest.raises(L.SrgpError, match="bit-identical") as e:
        ctx.gauss_obj_grad("vi", "sqexp", xu, cp["sigma"], [cp["l"]], cp["tau"], 1e-2)
    assert e.value.status == L.ERR_STATE
    # the context stays usable and the next evaluation is clean
    x2, y2, xu2, _ = _gridded(5000, 16)
    ctx.set_data(x2, y2, None)
    obj, grad = ctx.gauss_obj_grad("vi", "sqexp", xu2, cp["sigma"], [cp["l"]], cp["tau"], 1e-2)
    assert np.isfinite(obj) and np.all(np.isfinite(grad))
