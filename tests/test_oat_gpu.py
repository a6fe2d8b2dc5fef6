"""OAT candidate scoring: srgp_oat_scores against the reference's per-candidate loop
(oracle/ref_model.py:oat_candidate_scores, R/vi_functions.R:2211-2298, R/knot_proposal_functions.R:1283-1353).
Objectives: rel <= 1e-8 (north_star)."""
import numpy as np
import pytest

from oracle import reduced_model as red
from oracle import ref_model as rm
from sparsergps_b200 import vi_functions as vf
from tests import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _candidates(c, T, seed=5):
    rng = np.random.default_rng(seed)
    return c["x"][rng.choice(len(c["x"]), T, replace=False)]


@pytest.mark.parametrize("vi", [True, False])
@pytest.mark.parametrize("case", ["config1", "config2", "config3"])
def test_oat_scores_match_per_candidate_loop(ctx, case, vi):
    c = {"config1": lambda: cases.config1(), "config2": lambda: cases.config2(n=1503, m=64),
         "config3": lambda: cases.config3(n=2000, m=100, extra_knot_is_data_row=False)}[case]()
    cp, cf = c["cov_par"], c["cov_fun"]
    cand = _candidates(c, 9)
    obj0, scores = vf.oat_candidate_scores(cp, cf, c["xu"], c["x"], c["y"], c["mu"], cand, c["delta"], vi=vi, ctx=ctx)
    ref = rm.oat_candidate_scores(cp, cf, c["xu"], c["x"], c["y"], c["mu"], cand, c["delta"], vi=vi)
    ref0 = (rm.vi_obj_grad if vi else rm.fic_obj_grad)(cp, cf, c["xu"], c["x"], c["y"], c["mu"], c["delta"])[0]
    assert obj0 == pytest.approx(ref0, rel=RTOL)
    np.testing.assert_allclose(scores, ref, rtol=RTOL)
    # adding a knot can only raise the VI bound (Titsias); a property the per-candidate loop shares
    if vi:
        assert np.all(scores >= obj0 - 1e-9 * abs(obj0))
    # and the proposal the optimiser receives is the same row
    norm_opt = dict(xu=c["xu"], cov_par=cp, xy=c["x"], mu=c["mu"], cov_fun=cf, obj_fun=[ref0])
    pick = (vf.knot_prop_random_norm_vi if vi else vf.knot_prop_random_norm)(
        norm_opt, cand, opt={"delta": c["delta"]}, ctx=ctx, y=c["y"])
    np.testing.assert_array_equal(pick, rm.knot_prop_choice(c["xu"], cand, ref0, ref))


def test_oat_vi_headline_knot_count_and_batches(ctx):
    """m = 1024 knots, 140 candidates (two device batches) against the NumPy bordered form; a few of them also
    against a plain (m+1)-knot evaluation through the fused entry point."""
    c = cases.config5(n=20000, m=1024)
    cp = c["cov_par"]
    l = cases.lvec(cp)
    cand = _candidates(c, 140)
    ctx.set_data(c["x"], c["y"], None)
    obj0, scores = ctx.oat_scores("vi", "ard", c["xu"], cand, cp["sigma"], l, cp["tau"], c["delta"])
    ref0, ref = red.vi_oat_scores(c["x"], c["y"], c["mu"], c["xu"], cand, cp["sigma"], l, cp["tau"], c["delta"])
    assert obj0 == pytest.approx(ref0, rel=RTOL)
    np.testing.assert_allclose(scores, ref, rtol=RTOL)
    for t in (0, 77, 139):
        full = ctx.gauss_obj_grad("vi", "ard", np.vstack([c["xu"], cand[t]]), cp["sigma"], l, cp["tau"], c["delta"],
                                  want_grad=False)[0]
        assert scores[t] == pytest.approx(full, rel=1e-10)


def test_oat_duplicate_candidate_is_flagged(ctx):
    """A candidate equal to an existing knot with delta = 0 makes S+ singular: R's solve() / chol() raise and the
    proposal function resamples; here that candidate's VI score is NaN and the others are unaffected.  (FIC scores
    each candidate by a full evaluation whose Cholesky, like LAPACK's, only trips on a non-positive pivot -- whether
    an exact duplicate produces one is a rounding accident there as in the reference, so only the clean candidates
    are checked.)"""
    c = cases.config2(n=600, m=20)
    cp = c["cov_par"]
    l = cases.lvec(cp)
    cand = np.vstack([_candidates(c, 3), c["xu"][4]])
    ctx.set_data(c["x"], c["y"], None)
    _, scores = ctx.oat_scores("vi", "ard", c["xu"], cand, cp["sigma"], l, cp["tau"], 0.0)
    assert np.isnan(scores[3]) and np.all(np.isfinite(scores[:3]))
    for model in ("vi", "fic"):
        _, scores = ctx.oat_scores(model, "ard", c["xu"], cand, cp["sigma"], l, cp["tau"], 0.0)
        _, clean = ctx.oat_scores(model, "ard", c["xu"], cand[:3], cp["sigma"], l, cp["tau"], 0.0)
        np.testing.assert_allclose(scores[:3], clean, rtol=1e-12)


@pytest.mark.parametrize("family", ["bernoulli", "poisson"])
def test_laplace_oat_scores_match_per_candidate_newton(ctx, family):
    """knot_prop_random's candidate loop (R/knot_proposal_functions.R:1096-1120) for the sparse Laplace models."""
    from sparsergps_b200 import laplace as Lp
    c = cases.config4(n=600, m=20)
    if family == "poisson":
        c["y"] = np.random.default_rng(3).poisson(np.exp(0.5 * np.sin(c["x"][:, 0]))).astype(np.float64)
    cp = c["cov_par"]
    kw = {"m": 1.0} if family == "poisson" else {}
    fit = rm.newtrap_sparseGP(np.zeros(600), family, cp, "ard", c["x"], c["xu"], c["y"], c["mu"], np.zeros(20),
                              maxit=40, tol=1e-5, delta=c["delta"], **kw)
    cand = _candidates(c, 5)
    ref = rm.laplace_oat_candidate_scores(cp, "ard", c["xu"], c["x"], c["y"], fit["gp"], family, c["mu"], np.zeros(20),
                                          cand, c["delta"], maxit=40, tol=1e-5, **kw)
    lo = dict(xu=c["xu"], cov_par=cp, xy=c["x"], mu=c["mu"], muu=np.zeros(20), cov_fun="ard", fmax=fit["gp"],
              obj_fun=fit["objective_function_values"])
    pick, scores = Lp.knot_prop_random(lo, cand, family, c["y"], opt={"delta": c["delta"]}, maxit=40, tol=1e-5, ctx=ctx,
                                       return_scores=True)
    np.testing.assert_allclose(scores, ref, rtol=1e-8)
    np.testing.assert_array_equal(pick, rm.knot_prop_choice(c["xu"], cand, fit["objective_function_values"][-1], ref))
