"""The R-facing boundary at the SEXP level: r/shim.c, compiled UNCHANGED against the miniature R C API in
tests/mini_r/ and linked with libsrgp.so into sparseRGPs.so, driven the way R drives a package DLL
(R_init_sparseRGPs -> registered routine table -> .Call with arity check -> SEXP results / R errors).

CPU part: the table R/RcppExports.R binds against (src/RcppExports.cpp:285-304), coercion, list lookup by name,
errors, PROTECT balance, and the host-side routines against the reference's golden vectors.
GPU part (-m gpu): the four matrix builders through `.Call` reproduce the reference's golden matrices; the fused
entry points agree with the oracle.
"""
import numpy as np
import pytest

from tests import cases
from tests import golden_util as G
from tests.mini_r import driver as D

# src/RcppExports.cpp:285-304 -- the routine table of the reference (name, number of arguments)
REFERENCE_TABLE = {
    "_sparseRGPs_real_to_pos": 1, "_sparseRGPs_pos_to_real": 1, "_sparseRGPs_real_to_bounded": 3,
    "_sparseRGPs_dsqexp_dsigmaC": 3, "_sparseRGPs_dsqexp_dsigma_ardC": 4, "_sparseRGPs_dsqexp_dlC": 3,
    "_sparseRGPs_dsqexp_dl_ardC": 5, "_sparseRGPs_dsqexp_dtauC": 3, "_sparseRGPs_dsqexp_dx2C": 5,
    "_sparseRGPs_dsqexp_dx2_ardC": 6, "_sparseRGPs_dexp_dsigmaC": 3, "_sparseRGPs_dexp_dlC": 3,
    "_sparseRGPs_dexp_dtauC": 3, "_sparseRGPs_dsig_dthetaC": 5, "_sparseRGPs_dsig_dtheta_ardC": 6,
    "_sparseRGPs_cov_fun_sqrd_expC": 3, "_sparseRGPs_cov_fun_sqrd_exp_ardC": 4, "_sparseRGPs_cov_fun_expC": 3,
    "_sparseRGPs_make_cov_matC": 5, "_sparseRGPs_make_cov_mat_ardC": 6,
}

# argument order of the R wrappers (R/RcppExports.R:7-127)
ARG_ORDER = {
    "real_to_pos": ["x"], "pos_to_real": ["x"], "real_to_bounded": ["x", "ub", "lb"],
    "cov_fun_sqrd_expC": ["x1", "x2", "cov_par"], "cov_fun_sqrd_exp_ardC": ["x1", "x2", "cov_par", "lnames"],
    "cov_fun_expC": ["x1", "x2", "cov_par"],
    "dsqexp_dsigmaC": ["x1", "x2", "cov_par"], "dsqexp_dsigma_ardC": ["x1", "x2", "cov_par", "lnames"],
    "dsqexp_dlC": ["x1", "x2", "cov_par"], "dsqexp_dl_ardC": ["x1", "x2", "cov_par", "lnames", "comp"],
    "dsqexp_dtauC": ["x1", "x2", "cov_par"], "dsqexp_dx2C": ["x1", "x2", "cov_par", "lb", "ub"],
    "dsqexp_dx2_ardC": ["x1", "x2", "cov_par", "lb", "ub", "lnames"],
    "dexp_dsigmaC": ["x1", "x2", "cov_par"], "dexp_dlC": ["x1", "x2", "cov_par"], "dexp_dtauC": ["x1", "x2", "cov_par"],
    "make_cov_matC": ["x", "x_pred", "cov_par", "cov_fun", "delta"],
    "make_cov_mat_ardC": ["x", "x_pred", "cov_par", "cov_fun", "delta", "lnames"],
    "dsig_dthetaC": ["x", "x_pred", "cov_par", "cov_fun", "par_name"],
    "dsig_dtheta_ardC": ["x", "x_pred", "cov_par", "cov_fun", "par_name", "lnames"],
}

CASES = G.load()
HOST_CASES = [c for c in CASES if c[0] not in G.MATRIX_FNS]
MATRIX_CASES = [c for c in CASES if c[0] in G.MATRIX_FNS]


def r_call(fn, kw):
    """R wrapper `fn(...)` -> .Call('_sparseRGPs_fn', ...) with R-typed arguments."""
    args = []
    for name in ARG_ORDER[fn]:
        v = kw[name]
        if name == "x_pred" and v is None:
            v = D.NAMatrix()                         # matrix(): 1 x 1 logical NA
        elif name in ("x", "x_pred"):
            v = np.asarray(v, dtype=np.float64).reshape(len(v), -1)
        elif name in ("delta", "comp"):
            v = np.array([float(v)])
        args.append(v)
    return D.dot_call("_sparseRGPs_" + fn, *args)


def _check(got, exp):
    if isinstance(exp, dict):
        assert list(got) == ["derivative", "trans_par", "inv_trans_par"]      # names and order of the Rcpp list
        for k in exp:
            np.testing.assert_allclose(got[k], exp[k], rtol=1e-10, atol=1e-300, err_msg=k)
    else:
        assert got.shape == exp.shape
        np.testing.assert_allclose(got, exp, rtol=1e-10, atol=1e-300)


# ------------------------------------------------------------------------------------------------ CPU
def test_shim_compiles_warning_free_and_registers_the_reference_table():
    D.build(force=True)                                # gcc -Wall -Werror on r/shim.c, unchanged
    table = D.routines()
    for name, nargs in REFERENCE_TABLE.items():
        assert table.get(name) == nargs, name
    assert D.lib().mr_dynamic_symbols() == 0           # R_useDynamicSymbols(dll, FALSE), src/RcppExports.cpp:309


@pytest.mark.parametrize("case", HOST_CASES, ids=G.ids(HOST_CASES))
def test_host_routines_through_dot_call_reproduce_golden(case):
    fn, kw, exp = case
    got = r_call(fn, kw)
    _check(got if isinstance(exp, dict) else np.asarray(got).reshape(exp.shape), exp)


def test_integer_vectors_are_coerced_like_rcpp_input_parameter():
    cp = {"sigma": 2.0, "l": 1.0, "tau": 1.0}
    a = D.dot_call("_sparseRGPs_cov_fun_sqrd_expC", np.array([0, 2], dtype=np.int32), np.array([1, 1], dtype=np.int32), cp)
    b = D.dot_call("_sparseRGPs_cov_fun_sqrd_expC", np.array([0.0, 2.0]), np.array([1.0, 1.0]), cp)
    assert a[0] == b[0]
    assert D.dot_call("_sparseRGPs_real_to_pos", np.array([0, 1], dtype=np.int32))[1] == pytest.approx(np.e, rel=1e-15)


def test_list_lookup_is_by_name_and_missing_elements_are_r_errors():
    a = D.dot_call("_sparseRGPs_cov_fun_sqrd_expC", [0.0], [1.0], {"tau": 9.0, "l": 1.5, "sigma": 2.0})
    b = D.dot_call("_sparseRGPs_cov_fun_sqrd_expC", [0.0], [1.0], {"sigma": 2.0, "l": 1.5})
    assert a[0] == b[0]
    with pytest.raises(D.RError, match="Index out of bounds"):
        D.dot_call("_sparseRGPs_cov_fun_sqrd_expC", [0.0], [1.0], {"sigma": 2.0})


def test_dot_call_checks_registration_and_arity():
    with pytest.raises(D.RError, match="Incorrect number of arguments"):
        D.dot_call("_sparseRGPs_real_to_pos", [0.0], [1.0])
    with pytest.raises(D.RError, match="not available"):
        D.dot_call("srgp_make_cov_mat", [0.0])          # the C ABI itself is not an R routine


def test_unknown_kernel_prints_the_message_and_returns_0x0_without_a_gpu(capfd):
    x = np.zeros((3, 2))
    out = D.dot_call("_sparseRGPs_make_cov_matC", x, D.NAMatrix(), {"sigma": 1.0, "l": 1.0, "tau": 1.0}, "matern",
                     np.array([1e-6]))
    assert out.shape == (0, 0)
    assert "Error: invalid covariance function" in capfd.readouterr().err


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", MATRIX_CASES, ids=G.ids(MATRIX_CASES))
def test_matrix_builders_through_dot_call_reproduce_golden(case, capfd):
    fn, kw, exp = case
    got = r_call(fn, kw)
    _check(got, exp)
    assert np.array_equal(got == 0.0, exp == 0.0)


@pytest.mark.gpu
def test_integer_design_matrix_and_real_na_sentinel():
    xi = np.arange(12, dtype=np.int32).reshape(6, 2) % 5
    cp = {"sigma": 1.3, "l": 2.0, "tau": 0.4}
    a = D.dot_call("_sparseRGPs_make_cov_matC", xi, D.NAMatrix(), cp, "sqexp", np.array([1e-4]))
    b = D.dot_call("_sparseRGPs_make_cov_matC", xi.astype(np.float64), np.array([[np.nan]]), cp, "sqexp", np.array([1e-4]))
    assert a.shape == (6, 6) and np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("model", [0, 1])
def test_fused_objective_gradient_entry_point_matches_oracle(model):
    from oracle import ref_model as rm
    c = cases.config2(n=600, m=32)
    ln = [k for k in c["cov_par"] if k.startswith("l")]
    out = D.dot_call("_sparseRGPs_gauss_obj_grad", np.array([model], dtype=np.int32), "ard", c["x"], c["y"], c["mu"],
                     c["xu"], c["cov_par"], np.array([c["delta"]]), ln)
    assert list(out) == ["objective", "gradient"]
    f = rm.vi_obj_grad if model == 0 else rm.fic_obj_grad
    obj, g = f(c["cov_par"], "ard", c["xu"], c["x"], c["y"], c["mu"], c["delta"])
    assert out["objective"][0] == pytest.approx(obj, rel=1e-8)
    gref = np.array([g[k] for k in c["cov_par"]])
    np.testing.assert_allclose(out["gradient"], gref, rtol=1e-8, atol=1e-8 * np.abs(gref).max())


@pytest.mark.gpu
def test_failed_cholesky_is_an_r_error():
    """The OAT code wraps these calls in try() (R/knot_proposal_functions.R:1314-1354)."""
    c = cases.config2(n=300, m=16)
    xu = np.vstack([c["xu"], c["xu"][:1]])               # duplicated knot, delta = 0 -> singular Sigma22
    ln = [k for k in c["cov_par"] if k.startswith("l")]
    with pytest.raises(D.RError, match="sparseRGPs"):
        D.dot_call("_sparseRGPs_gauss_obj_grad", np.array([0], dtype=np.int32), "ard", c["x"], c["y"], c["mu"], xu,
                   c["cov_par"], np.array([0.0]), ln)


@pytest.mark.gpu
def test_trace_term_entry_point():
    from oracle import ref_kernels as rk
    from oracle import ref_model as rm
    c = cases.config2(n=400, m=24)
    ln = [k for k in c["cov_par"] if k.startswith("l")]
    K = rk.make_cov_mat_ardC(c["x"], c["xu"], c["cov_par"], "ard", 0.0, ln)
    S = rk.make_cov_mat_ardC(c["xu"], None, dict(c["cov_par"], tau=0.0), "ard", c["delta"], ln)
    got = D.dot_call("_sparseRGPs_trace_term", [c["cov_par"]["sigma"]], [c["cov_par"]["tau"]], [c["delta"]], K, S)
    assert got[0] == pytest.approx(rm.trace_term_fun(c["cov_par"], K, S, c["delta"]), rel=1e-8)
