"""Pin the oracle's R-level algebra (no R here -> 'parity unpinned' for it; the Rcpp layer is
pinned against the compiled reference in tests/test_reference_pin.py):
analytic known answers, quirks, finite differences of the Gaussian objectives, and literal == reduced form."""
import math

import numpy as np
import pytest

from oracle import reduced_model as red
from oracle import ref_kernels as rk
from oracle import ref_model as rm
from tests import cases


def test_known_answers_per_element():
    # SURVEY.md Appendix A known answers
    K = rk.make_cov_matC(np.array([[0.0]]), np.array([[1.0]]), {"sigma": 1, "l": 1, "tau": 0}, "sqexp", 0.0)
    assert K[0, 0] == pytest.approx(0.6065306597126334, rel=1e-15)
    cp = {"sigma": 2, "l1": 1, "l2": 2, "tau": 0.5}
    K = rk.make_cov_mat_ardC(np.array([[1.0, 2.0]]), np.array([[0.0, 0.0]]), cp, "ard", 0.0, ["l1", "l2"])
    assert K[0, 0] == pytest.approx(1.4715177646857693, rel=1e-15)
    x = np.random.default_rng(0).normal(size=(7, 2))
    S = rk.make_cov_mat_ardC(x, None, cp, "ard", 1e-3, ["l1", "l2"])
    np.testing.assert_allclose(np.diag(S), 4 + 0.25 + 1e-3, rtol=1e-15)      # sigma^2 + tau^2 + delta
    np.testing.assert_allclose(S, S.T, rtol=1e-15)
    Kc = rk.make_cov_mat_ardC(x, x, cp, "ard", 1e-3, ["l1", "l2"])             # cross: no nugget
    np.testing.assert_allclose(np.diag(Kc), 4.0, rtol=1e-15)
    dsig = rk.dsig_dtheta_ardC(x, x[:3], cp, "ard", "sigma", ["l1", "l2"])
    np.testing.assert_allclose(dsig, 2 * Kc[:, :3], rtol=1e-15)
    dl = sum(rk.dsig_dtheta_ardC(x, x[:3], {"sigma": 2, "l1": 1.5, "l2": 1.5, "tau": 0.5}, "ard", nm, ["l1", "l2"])
             for nm in ("l1", "l2"))
    dl_iso = rk.dsig_dthetaC(x, x[:3], {"sigma": 2, "l": 1.5, "tau": 0.5}, "sqexp", "l")
    np.testing.assert_allclose(dl, dl_iso, rtol=1e-13, atol=1e-300)
    assert float(rk.real_to_bounded(0.0, 3.0, 1.0)) == pytest.approx(2.0)


def test_quirks():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(6, 3))
    xu = np.vstack([x[2], rng.normal(size=(2, 3)), x[2]])
    cp = {"sigma": 1.1, "l1": 1, "l2": 2, "l3": 3, "tau": 0.7}
    ln = ["l1", "l2", "l3"]
    # Q4: tau derivative = 2 tau^2 wherever rows are bit-equal, in the cross matrix too
    dt = rk.dsig_dtheta_ardC(x, xu, cp, "ard", "tau", ln)
    expect = np.zeros((6, 4)); expect[2, 0] = expect[2, 3] = 2 * 0.49
    np.testing.assert_allclose(dt, expect, rtol=1e-15)
    # Q5: nugget by index equality only; duplicated knots get no off-diagonal nugget
    S = rk.make_cov_mat_ardC(xu, None, cp, "ard", 1e-2, ln)
    assert S[0, 3] == pytest.approx(1.21, rel=1e-15) and S[0, 0] == pytest.approx(1.21 + 0.49 + 1e-2, rel=1e-15)
    # Q8/Q9: exp kernel L1 vs L2 and the unreachable tau branch
    cpe = {"sigma": 1.0, "l": 2.0, "tau": 0.3}
    a, b = np.array([[0.0, 0.0]]), np.array([[3.0, 4.0]])
    assert rk.make_cov_matC(a, b, cpe, "exp", 0)[0, 0] == pytest.approx(math.exp(-7 / 2))
    assert rk.dsig_dthetaC(a, b, cpe, "exp", "sigma")[0, 0] == pytest.approx(2 * math.exp(-5 / 2))
    assert rk.dsig_dthetaC(a, a, cpe, "exp", "tau")[0, 0] == 0.0              # cross: zeros (Q9)
    assert rk.dsig_dthetaC(a, None, cpe, "exp", "tau")[0, 0] == pytest.approx(2 * 0.09)
    # unknown kernel -> 0 x 0 matrix, no exception
    assert rk.make_cov_matC(a, b, cpe, "matern", 0).shape == (0, 0)
    # Q1: Bernoulli W for y = 1 is -pi - pi^2
    f = np.array([0.3]); p = 1 / (1 + np.exp(-f))
    assert rm.d2log_py_dff_bern(f, np.array([1.0]))[0] == pytest.approx(float(-p - p ** 2), rel=1e-14)
    assert rm.d2log_py_dff_bern(f, np.array([0.0]))[0] == pytest.approx(float(-p * (1 - p)), rel=1e-14)


def _fd(fun, cp, h=1e-6):
    out = {}
    for k in cp:
        a, b = dict(cp), dict(cp)
        a[k] = cp[k] * math.exp(h); b[k] = cp[k] * math.exp(-h)
        out[k] = (fun(a) - fun(b)) / (2 * h)
    return out


@pytest.mark.parametrize("case", ["config1", "config2"])
def test_gaussian_gradients_match_finite_differences(case):
    c = getattr(cases, case)(n=300) if case == "config1" else cases.config2(n=400, m=16)
    x, y, mu, xu, cp, cf, delta = c["x"], c["y"], c["mu"], c["xu"], c["cov_par"], c["cov_fun"], c["delta"]

    def vi(p):
        return rm.vi_obj_grad(p, cf, xu, x, y, mu, delta)[0]

    def fic(p):
        S12, S22, _ = rm.assemble(p, cf, x, xu, delta)
        return rm.obj_fun_norm(mu, rm.fic_Z(p, S12, S22, delta), S12, S22, y)

    g_vi = rm.vi_obj_grad(cp, cf, xu, x, y, mu, delta)[1]
    g_fic = rm.fic_obj_grad(cp, cf, xu, x, y, mu, delta)[1]
    fd_vi, fd_fic = _fd(vi, cp), _fd(fic, cp)
    for k in cp:
        assert g_vi[k] == pytest.approx(fd_vi[k], rel=2e-6, abs=1e-6), k
        assert g_fic[k] == pytest.approx(fd_fic[k], rel=2e-6, abs=1e-6), k


@pytest.mark.parametrize("shards", [1, 3])
def test_reduced_form_equals_literal_vi(shards):
    for c in (cases.config2(n=700, m=32), cases.config3(n=600, m=24), cases.config5(n=512, m=128)):
        cp = c["cov_par"]
        obj, g = rm.vi_obj_grad(cp, c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], c["delta"])
        obj2, g2 = red.vi_obj_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"],
                                   c["delta"], shards=shards)
        assert obj2 == pytest.approx(obj, rel=1e-11)
        for k in g:
            assert g2[k] == pytest.approx(g[k], rel=1e-8, abs=1e-9 * max(abs(v) for v in g.values())), k


def test_newton_and_laplace_gradient_run():
    c = cases.config4(n=300, m=12)
    cp = c["cov_par"]
    fit = rm.newtrap_sparseGP(np.zeros(300), "bernoulli", cp, "ard", c["x"], c["xu"], c["y"], c["mu"],
                              np.zeros(12), maxit=400, tol=1e-5, delta=c["delta"])
    h = fit["objective_function_values"]
    assert np.all(np.isfinite(h)) and h[-1] > h[0]
    assert np.max(np.abs(fit["gradient"])) < 1e-3
    g = rm.dlogq_dcov_par(cp, "ard", c["xu"], c["x"], c["y"], fit["gp"], "bernoulli", c["mu"], c["delta"])
    assert all(np.isfinite(v) for v in g["gradient"].values())


def test_extended_precision_reference_agrees_when_well_conditioned():
    c = cases.config2(n=300, m=16)
    cp = c["cov_par"]
    for f in (red.vi_obj_grad, red.fic_obj_grad):
        o64, g64 = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
        with red.extended_precision():
            ox, gx = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"])
        assert type(ox).__name__ == "longdouble"
        assert float(o64) == pytest.approx(float(ox), rel=1e-13)
        for k in g64:
            assert float(g64[k]) == pytest.approx(float(gx[k]), rel=1e-11), k


@pytest.mark.parametrize("model", ["vi", "fic"])
@pytest.mark.parametrize("case", ["config1", "config2"])
def test_reduced_knot_gradient_equals_literal(model, case):
    """Knot-location gradient: column sums of Omega o dK plus the N o dSigma22 part reproduce the reference's
    per-knot loop (R/vi_functions.R:425-592, R/laplace_approx_gradient.R:965-1126), incl. the Q12 Jacobian."""
    c = cases.config1(n=200) if case == "config1" else cases.config2(n=400, m=20)
    cp, cf = c["cov_par"], c["cov_fun"]
    opt = None if case == "config1" else [0, 9, 19]
    lit = (rm.delbo_dcov_par if model == "vi" else rm.dlogp_dcov_par)(
        cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"], dcov_fun_dknot=rm.dcov_fun_dknot_for(cf), knot_opt=opt)
    f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
    l = cases.lvec(cp) if cf == "ard" else cp["l"]
    _, _, kg = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], cf, shards=3, knots=True)
    m, d = c["xu"].shape
    ref = lit["knot_gradient"].reshape(m, d)
    sel = list(range(m)) if opt is None else opt
    np.testing.assert_allclose(kg[sel], ref[sel], rtol=1e-9, atol=1e-11 * np.abs(ref).max())
    rest = [k for k in range(m) if k not in sel]
    assert not ref[rest].any()
    assert lit["trans_knot"].shape == (m, d)


def test_bordered_oat_scores_equal_per_candidate_loop():
    """OAT scoring: one Gram of [knots | candidates] + Schur complements == the reference's loop of full
    (m+1)-knot elbo_fun evaluations (R/vi_functions.R:2211-2298)."""
    c = cases.config2(n=500, m=24)
    cp = c["cov_par"]
    cand = c["x"][[3, 77, 210, 499]]
    lit = rm.oat_candidate_scores(cp, "ard", c["xu"], c["x"], c["y"], c["mu"], cand, c["delta"], vi=True)
    obj0, bordered = red.vi_oat_scores(c["x"], c["y"], c["mu"], c["xu"], cand, cp["sigma"], cases.lvec(cp), cp["tau"],
                                       c["delta"], shards=2)
    np.testing.assert_allclose(bordered, lit, rtol=1e-12)
    assert obj0 == pytest.approx(rm.vi_obj_grad(cp, "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"])[0], rel=1e-12)
    pick = rm.knot_prop_choice(c["xu"], cand, obj0, lit)
    np.testing.assert_array_equal(pick[0], cand[int(np.argmax(lit))])          # every candidate raises the bound
    np.testing.assert_array_equal(rm.knot_prop_choice(c["xu"], cand, 1e9, lit)[0], c["xu"][0])   # none does: first knot


def test_optimiser_loop_transcription_basics():
    """norm_grad_ascent: ascent on the ELBO, the documented stop rule, and the sign-flip state staying inert while
    no gradient component changes sign (R/vi_functions.R:963-965,981-986,1132-1147)."""
    c = cases.config1(n=200)
    start = {"sigma": 1.0, "l": 0.5, "tau": 0.7}
    r = rm.norm_grad_ascent(start, "sqexp", c["xu"], c["x"], c["y"], c["mu"], {"maxit": 8, "obj_tol": 1e-9}, vi=True)
    assert r["iter"] == 8 and np.all(np.diff(r["obj_fun"]) > 0)
    # first ADADELTA step: sd2 = 0, sc = 0  ->  step_j = sqrt(eps) / sqrt(0.05 g_j^2 + eps) * g_j on log(theta)
    g0 = r["grad"][0]
    step = np.sqrt(1e-6) / np.sqrt(0.05 * g0 ** 2 + 1e-6) * g0
    np.testing.assert_allclose(np.log(r["cov_par_history"][1]), np.log(r["cov_par_history"][0]) + step, rtol=1e-12)
    r2 = rm.norm_grad_ascent(start, "sqexp", c["xu"], c["x"], c["y"], c["mu"], {"maxit": 50, "obj_tol": 1e9}, vi=True)
    assert r2["iter"] == 2                                   # iter = 1 always proceeds; |obj_2 - obj_1| <= tol stops
    ga = rm.norm_grad_ascent(start, "sqexp", c["xu"], c["x"], c["y"], c["mu"],
                             {"maxit": 3, "optim_method": "ga", "learn_rate": 1e-4}, vi=False)
    np.testing.assert_allclose(np.log(ga["cov_par_history"][1]), np.log(ga["cov_par_history"][0]) + 1e-4 * ga["grad"][0],
                               rtol=1e-12)
