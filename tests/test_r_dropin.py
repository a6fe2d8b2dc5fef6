"""The R-level drop-in, end to end: r/patches.R keeps the reference's names and signatures and
routes the bodies through `.Call`.

Two halves, because the reference's R sources exist only in the build container and the GPU only on the box:

* GPU (`-m gpu`, no /root/reference needed): r/patches.R ALONE is sourced into the mini-R interpreter; `.Call` is served
  by r/shim.c (compiled unchanged against tests/mini_r) -> libsrgp.so on the GPU.  Every patched function --
  trace_term_fun, elbo_fun, obj_fun_norm, delbo_dcov_par, dlogp_dcov_par, newtrap_sparseGP, dlogq_dcov_par, predict_vi,
  predict_laplace -- is called with the reference's own argument list on the golden inputs and must return what the
  UNPATCHED reference returned (tests/golden/r_level.*, produced by /root/reference/R/*.R).
* CPU (build container only, skipped where /root/reference is absent): the reference's R files AND r/patches.R are
  sourced; the reference's optimiser loops norm_grad_ascent_vi / norm_grad_ascent / laplace_grad_ascent and its
  predict_* run UNMODIFIED on top of the patched functions, with `.Call` of the fused routines served by the CPU
  oracle standing in for the library.  Same goldens: the callers cannot tell the patched functions from their own.
"""
import glob
import os

import numpy as np
import pytest

from oracle.mini_r import interp as RI
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHES = os.path.join(ROOT, "r", "patches.R")
REF_R_DIR = "/root/reference/R"
G = gu.load_r_level()

# The likelihood-derivative closures of the reference are only IDENTITIES to the patched functions (they select the
# family); without the package they are stand-ins with the package's names.
STUBS = r'''
dlog_py_dff_bern <- function(...) stop("stub"); d2log_py_dff_bern <- function(...) stop("stub")
d3log_py_dff_bern <- function(...) stop("stub"); grad_loglik_fn_bern <- function(...) stop("stub")
obj_fun_bern <- function(...) stop("stub")
dlog_py_dff_pois <- function(...) stop("stub"); d2log_py_dff_pois <- function(...) stop("stub")
d3log_py_dff_pois <- function(...) stop("stub"); grad_loglik_fn_pois <- function(...) stop("stub")
obj_fun_pois <- function(...) stop("stub")
dsqexp_dsigma <- function(...) stop("stub"); dsqexp_dl <- function(...) stop("stub"); dsqexp_dtau <- function(...) stop("stub")
dsqexp_dsigma_ard <- function(...) stop("stub"); dsqexp_dx2 <- function(...) stop("stub"); dsqexp_dx2_ard <- function(...) stop("stub")
'''
GLUE = r'''
.t_dtheta <- function(cov_fun) {
  if(cov_fun == "sqexp") return(list("sigma" = dsqexp_dsigma, "l" = dsqexp_dl, "tau" = dsqexp_dtau))
  list("sigma" = dsqexp_dsigma_ard, "tau" = dsqexp_dtau)
}
.t_dknot <- function(cov_fun) { if(cov_fun == "sqexp") dsqexp_dx2 else dsqexp_dx2_ard }
'''


def _close(a, b, rtol, what=""):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.max(np.abs(b))), err_msg=what)


def _py_arg(v):
    """interpreter value -> what tests/mini_r/driver.to_sexp takes"""
    v = RI.to_py(v)
    if isinstance(v, np.ndarray) and v.dtype == object:
        return [str(s) for s in v]
    if isinstance(v, dict):
        return {k: (float(np.asarray(x).reshape(-1)[0]) if np.asarray(x).size == 1 else np.asarray(x)) for k, x in v.items()}
    return v


def _r_value(v):
    """driver.from_sexp result -> interpreter value (names of atomic vectors kept)"""
    from tests.mini_r import driver as D
    if isinstance(v, dict):
        return RI.RList([_r_value(x) for x in v.values()], list(v.keys()))
    if isinstance(v, D.NamedArray):
        return RI.Vec(np.asarray(v, dtype=np.float64).copy(), names=list(v.names))
    return RI.from_py(v)


def gpu_backend(I, pos):
    from tests.mini_r import driver as D
    name = str(pos[0].v[0])
    try:
        return _r_value(D.dot_call(name, *[_py_arg(a) for a in pos[1:]]))
    except D.RError as e:
        raise RI.RError(str(e))


# ------------------------------------------------------------------------------------------ CPU stand-in for the library
def _cp(lst):
    return {nm: float(RI.as_float(v)[0]) for nm, v in zip(lst.names, lst.items)}


def _named(g, cov_par, cov_fun):
    names = ["sigma"] + ([k for k in cov_par if k.startswith("l")] if cov_fun == "ard" else ["l"]) + ["tau"]
    return RI.Vec(np.array([float(g[k]) for k in names]), names=names)


CALLS = {}


def oracle_backend(I, pos):
    """`.Call` of the fused routines answered by oracle/ref_model.py with the entry points' argument order
    (r/shim.c); the 20 Rcpp routines go to the reference's compiled C++ as in oracle/ref_r.py."""
    from oracle import ref_model as rm
    from oracle import ref_r as rr
    name = str(pos[0].v[0]).replace("_sparseRGPs_", "")
    CALLS[name] = CALLS.get(name, 0) + 1
    a = pos[1:]
    F, M, S = RI.as_float, RI.matrix_of, lambda x: str(x.v[0])
    if name == "trace_term":
        return RI.dbl(rm.trace_term_fun({"sigma": F(a[0])[0], "tau": F(a[1])[0]}, M(a[3]), M(a[4]), F(a[2])[0]))
    if name == "gauss_obj_mats":
        n = M(a[0]).shape[0]
        return RI.dbl(rm.obj_fun_norm(np.broadcast_to(F(a[4]), (n,)) if len(F(a[4])) else np.zeros(n),
                                      np.broadcast_to(F(a[2]), (n,)), M(a[0]), M(a[1]), F(a[3])))
    if name in ("gauss_obj_grad", "gauss_obj_grad_knots"):
        model, cf, xy, y, mu, xu, cp, delta = int(F(a[0])[0]), S(a[1]), M(a[2]), F(a[3]), F(a[4]), M(a[5]), _cp(a[6]), F(a[7])[0]
        fn, obj = (rm.delbo_dcov_par, rm.vi_obj_grad) if model == 0 else (rm.dlogp_dcov_par, rm.fic_obj_grad)
        kw = {}
        if name.endswith("knots"):
            kw = dict(dcov_fun_dknot=rm.dcov_fun_dknot_for(cf), knot_opt=[int(k) - 1 for k in F(a[10])], transform=bool(a[11].v[0]))
        g = fn(cp, cf, xu, xy, y, mu, delta, **kw)
        out = {"objective": RI.dbl(obj(cp, cf, xu, xy, y, mu, delta)[0]), "gradient": _named(g["gradient"], cp, cf)}
        if kw:
            out["knot_gradient"] = RI.Vec(np.asarray(g["knot_gradient"], dtype=np.float64).reshape(-1))
            out["trans_knot"] = RI.from_matrix(g["trans_knot"])
        return RI.RList(list(out.values()), list(out.keys()))
    if name == "laplace_newton":
        fam, cf, xy, y, mu, xu, muu, cp, delta = S(a[0]), S(a[1]), M(a[2]), F(a[3]), F(a[4]), M(a[5]), F(a[6]), _cp(a[7]), F(a[8])[0]
        kw = {"m": F(a[13])[0]} if fam == "poisson" else {}
        nr = rm.newtrap_sparseGP(F(a[10]).copy(), fam, cp, cf, xy, xu, y, mu, muu, maxit=int(F(a[11])[0]), tol=F(a[12])[0],
                                 delta=delta, **kw)
        keys = ["gp", "objective_function_values", "gradient", "u_posterior_mean", "u_posterior_variance"]
        return RI.RList([RI.from_matrix(nr[k]) if k.endswith("variance") else RI.Vec(np.asarray(nr[k], dtype=np.float64).reshape(-1))
                         for k in keys], keys)
    if name in ("laplace_grad", "laplace_grad_knots"):
        fam, cf, xy, y, mu, xu, cp, delta, ff = S(a[0]), S(a[1]), M(a[2]), F(a[3]), F(a[4]), M(a[5]), _cp(a[6]), F(a[7])[0], F(a[9])
        kw = {"m": F(a[10])[0]} if fam == "poisson" else {}
        if name.endswith("knots"):
            kw.update(dcov_fun_dknot=rm.dcov_fun_dknot_for(cf), knot_opt=[int(k) - 1 for k in F(a[12])], transform=bool(a[13].v[0]))
        g = rm.dlogq_dcov_par(cp, cf, xu, xy, y, ff, fam, mu, delta, **kw)
        out = {"gradient": _named(g["gradient"], cp, cf)}
        if name.endswith("knots"):
            out["knot_gradient"] = RI.Vec(np.asarray(g["knot_gradient"], dtype=np.float64).reshape(-1))
            out["trans_knot"] = RI.from_matrix(g["trans_knot"])
        return RI.RList(list(out.values()), list(out.keys()))
    if name == "predict":
        from oracle import ref_kernels as rk
        cf, xp, mup, xu, muu, um, uv, cp = S(a[0]), M(a[1]), F(a[2]), M(a[3]), F(a[4]), F(a[5]), M(a[6]), _cp(a[7])
        ln = [str(s) for s in a[8].v]
        nug, vc = F(a[9])[0], F(a[10])[0]
        mk = (lambda x, p: rk.make_cov_mat_ardC(x, p, dict(cp, tau=0.0), cf, nug, ln)) if cf == "ard" else \
            (lambda x, p: rk.make_cov_matC(x, p, dict(cp, tau=0.0), cf, nug))
        S12, S22 = mk(xp, xu), mk(xu, None)
        Si = np.linalg.solve(S22, np.eye(len(xu)))
        T = -Si + Si @ uv @ Si
        pm = np.broadcast_to(mup, (len(xp),)) + (S12 @ np.linalg.solve(S22, (um - np.broadcast_to(muu, um.shape)).reshape(-1, 1))).reshape(-1)
        return RI.RList([RI.Vec(pm.copy()), RI.Vec(vc + np.sum((S12 @ T) * S12, axis=1))], ["pred_mean", "pred_var"])
    return rr._dot_call(I, pos, {})


# ------------------------------------------------------------------------------------------ sessions
def make_session(backend, with_reference):
    I = RI.Interp()
    I.globalenv.vars[".Call"] = RI.Builtin(".Call", lambda I_, pos, kw: backend(I_, pos))
    if with_reference:
        for path in sorted(glob.glob(os.path.join(REF_R_DIR, "*.R"))):
            I.source(path)
    else:
        I.run(STUBS)
    I.source(PATCHES)
    I.run(GLUE)
    return I


def rcall(I, fname, **kwargs):
    f = I.get_fun(fname, I.globalenv)
    args = [(k, v if isinstance(v, (RI.Vec, RI.RList, RI.Closure, RI.Builtin)) else RI.from_py(v)) for k, v in kwargs.items()]
    return RI.to_py(I.apply_function(f, args))


def rfun(I, name):
    return I.get_fun(name, I.globalenv)


_gpu_session = None


@pytest.fixture
def gpu_r():
    global _gpu_session
    if _gpu_session is None:
        _gpu_session = make_session(gpu_backend, with_reference=False)
    return _gpu_session


def _gauss_mats(c):
    o = c["out"]
    return o["Sigma12"], o["Sigma22"]


def check_gaussian_case(I, name, rtol):
    c = G[name]
    cp, cf, delta, i, o = c["meta"]["cov_par"], c["meta"]["cov_fun"], c["meta"]["delta"], c["in"], c["out"]
    S12, S22 = _gauss_mats(c)
    n, m = S12.shape
    tt = rcall(I, "trace_term_fun", cov_par=cp, Sigma12=S12, Sigma22=S22, delta=delta)
    _close(tt, o["trace_term"], rtol, "trace_term_fun")
    Z_vi = np.full(n, cp["tau"] ** 2 + delta)
    elbo = rcall(I, "elbo_fun", mu=i["mu"], Z=Z_vi, Sigma12=S12, Sigma22=S22, y=i["y"], trace_term_fun=rfun(I, "trace_term_fun"),
                 cov_par=cp, delta=delta)
    _close(elbo, o["elbo"], rtol, "elbo_fun")
    objn = rcall(I, "obj_fun_norm", mu=i["mu"], Z=o["Z_fic"], Sigma12=S12, Sigma22=S22, y=i["y"])
    _close(objn, o["obj_fun_norm"], rtol, "obj_fun_norm")
    dth = I.apply_function(rfun(I, ".t_dtheta"), [("cov_fun", RI.from_py(cf))])
    for tag, fn in (("vi", "delbo_dcov_par"), ("fic", "dlogp_dcov_par")):
        kw = dict(cov_par=cp, cov_fun=cf, dcov_fun_dtheta=dth, knot_opt=np.arange(1, m + 1), xu=i["xu"], xy=i["xy"], y=i["y"],
                  mu=i["mu"], transform=True, delta=delta)
        if c["meta"]["knots"]:
            kw["dcov_fun_dknot"] = I.apply_function(rfun(I, ".t_dknot"), [("cov_fun", RI.from_py(cf))])
        g = rcall(I, fn, **kw)
        assert list(g["gradient"]) == list(cp)                       # named, in cov_par's own order
        _close([g["gradient"][k] for k in cp], o[tag + "_gradient"], rtol, tag + " gradient")
        _close([np.asarray(g["trans_par"][k]).reshape(-1)[0] for k in cp], o[tag + "_trans_par"], 1e-14)
        if c["meta"]["knots"]:
            _close(g["knot_gradient"], o[tag + "_knot_gradient"], rtol, tag + " knot gradient")
            _close(g["trans_knot"], o[tag + "_trans_knot"], 1e-12)


def check_laplace_case(I, name, rtol):
    c = G[name]
    cp, cf, delta, fam, i, o = c["meta"]["cov_par"], c["meta"]["cov_fun"], c["meta"]["delta"], c["meta"]["family"], c["in"], c["out"]
    sfx = {"bernoulli": "bern", "poisson": "pois"}[fam]
    extra = gu.r_case_extra(c)
    m = len(i["xu"])
    nr = rcall(I, "newtrap_sparseGP", start_vals=i["mu"].copy(), obj_fun=rfun(I, "obj_fun_" + sfx),
               grad_loglik_fn=rfun(I, "grad_loglik_fn_" + sfx), dlog_py_dff=rfun(I, "dlog_py_dff_" + sfx),
               d2log_py_dff=rfun(I, "d2log_py_dff_" + sfx), maxit=1000, tol=1e-6, cov_par=cp, cov_fun=cf, xy=i["xy"], xu=i["xu"],
               y=i["y"], mu=i["mu"], muu=i["muu"], delta=delta, **extra)
    assert list(nr) == ["gp", "objective_function_values", "gradient", "u_posterior_mean", "u_posterior_variance"]
    assert len(nr["objective_function_values"]) == len(o["objective_function_values"])
    _close(nr["objective_function_values"], o["objective_function_values"], rtol, "Newton objective history")
    _close(nr["gp"], o["gp"], max(rtol, 1e-7), "mode")
    _close(nr["u_posterior_mean"], o["u_posterior_mean"], max(rtol, 1e-7))
    _close(nr["u_posterior_variance"], o["u_posterior_variance"], max(rtol, 1e-6))
    dth = I.apply_function(rfun(I, ".t_dtheta"), [("cov_fun", RI.from_py(cf))])
    kw = dict(cov_par=cp, cov_fun=cf, dcov_fun_dtheta=dth, knot_opt=np.arange(1, m + 1), xu=i["xu"], xy=i["xy"], y=i["y"],
              ff=o["gp"], dlog_py_dff=rfun(I, "dlog_py_dff_" + sfx), d2log_py_dff=rfun(I, "d2log_py_dff_" + sfx),
              d3log_py_dff=rfun(I, "d3log_py_dff_" + sfx), mu=i["mu"], transform=True, delta=delta, **extra)
    if c["meta"]["knots"]:
        kw["dcov_fun_dknot"] = I.apply_function(rfun(I, ".t_dknot"), [("cov_fun", RI.from_py(cf))])
    g = rcall(I, "dlogq_dcov_par", **kw)
    assert list(g["gradient"]) == list(cp)
    _close([g["gradient"][k] for k in cp], o["gradient"], rtol, "dlogq gradient")
    if c["meta"]["knots"]:
        _close(g["knot_gradient"], o["knot_gradient"], rtol, "dlogq knot gradient")
    pr = rcall(I, "predict_laplace", u_mean=o["u_posterior_mean"], u_var=o["u_posterior_variance"], xu=i["xu"], x_pred=o["x_pred"],
               cov_fun=cf, cov_par=cp, mu=np.full(len(o["x_pred"]), i["mu"][0]), muu=i["muu"], full_cov=False, family=fam, delta=delta)
    assert np.asarray(pr["pred_mean"]).shape == (len(o["x_pred"]), 1)            # the reference returns an n x 1 matrix
    _close(pr["pred_mean"], o["pred_mean"], rtol, "pred_mean")
    _close(pr["pred_var"], o["pred_var"], max(rtol, 1e-7), "pred_var")


def check_predict_vi(I, name, rtol):
    c = G[name]
    meta, i, o = c["meta"], c["in"], c["out"]
    cpf = dict(zip(meta["cov_par"], o["cov_par"].tolist()))
    pr = rcall(I, "predict_vi", u_mean=o["u_mean"], u_var=o["u_var"], xu=o["xu_final"], x_pred=o["x_pred"], cov_fun=meta["cov_fun"],
               cov_par=cpf, mu=np.full(len(o["x_pred"]), i["mu"][0]), muu=i["muu"], full_cov=False, family="gaussian", delta=meta["delta"])
    _close(pr["pred_mean"], o["pred_mean"], rtol, "pred_mean")
    _close(pr["pred_var"], o["pred_var"], max(rtol, 1e-7), "pred_var")


# ------------------------------------------------------------------------------------------ GPU half
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(k for k in G if k.startswith("g_")))
def test_patched_gaussian_functions_on_the_gpu_return_what_the_reference_returns(gpu_r, name):
    check_gaussian_case(gpu_r, name, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(k for k in G if k.startswith("l_")))
def test_patched_laplace_functions_on_the_gpu_return_what_the_reference_returns(gpu_r, name):
    check_laplace_case(gpu_r, name, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(k for k in G if k.startswith("f_vi")))
def test_patched_predict_vi_on_the_gpu(gpu_r, name):
    check_predict_vi(gpu_r, name, 1e-8)


@pytest.mark.gpu
def test_gradient_names_follow_cov_par_order_and_scalar_mu_recycles(gpu_r):
    """a cov_par listed as (sigma, tau, l) must get its gradient by NAME; a scalar mu recycles as in R."""
    c = G["g_sqexp_2d_mu"]
    cp, i, o = c["meta"]["cov_par"], c["in"], c["out"]
    perm = {"sigma": cp["sigma"], "tau": cp["tau"], "l": cp["l"]}
    dth = gpu_r.apply_function(rfun(gpu_r, ".t_dtheta"), [("cov_fun", RI.from_py("sqexp"))])
    g = rcall(gpu_r, "delbo_dcov_par", cov_par=perm, cov_fun="sqexp", dcov_fun_dtheta=dth, knot_opt=np.arange(1, 7), xu=i["xu"],
              xy=i["xy"], y=i["y"], mu=i["mu"], transform=True, delta=c["meta"]["delta"])
    assert list(g["gradient"]) == ["sigma", "tau", "l"]
    ref = dict(zip(cp, o["vi_gradient"]))
    _close([g["gradient"][k] for k in perm], [ref[k] for k in perm], 1e-8)
    g0 = rcall(gpu_r, "delbo_dcov_par", cov_par=cp, cov_fun="sqexp", dcov_fun_dtheta=dth, knot_opt=np.arange(1, 7), xu=i["xu"],
               xy=i["xy"], y=i["y"], mu=np.array([0.25]), transform=True, delta=c["meta"]["delta"])
    g1 = rcall(gpu_r, "delbo_dcov_par", cov_par=cp, cov_fun="sqexp", dcov_fun_dtheta=dth, knot_opt=np.arange(1, 7), xu=i["xu"],
               xy=i["xy"], y=i["y"], mu=np.full(len(i["y"]), 0.25), transform=True, delta=c["meta"]["delta"])
    _close([g0["gradient"][k] for k in cp], [g1["gradient"][k] for k in cp], 1e-13)


# ------------------------------------------------------------------------------------------ CPU half (build container)
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_R_DIR), reason="the reference's R sources exist only in the build container")
_cpu_session = None


@pytest.fixture
def cpu_r():
    global _cpu_session
    if _cpu_session is None:
        from oracle import ref_native
        ref_native.build()
        _cpu_session = make_session(oracle_backend, with_reference=True)
    return _cpu_session


@needs_reference
def test_patches_keep_the_reference_bodies_as_fallbacks(cpu_r):
    for nm in ("delbo_dcov_par", "dlogp_dcov_par", "newtrap_sparseGP", "dlogq_dcov_par", "elbo_fun", "obj_fun_norm",
               "predict_vi", "predict_laplace", "trace_term_fun"):
        assert isinstance(rfun(cpu_r, nm + "_R"), RI.Closure) and isinstance(rfun(cpu_r, nm), RI.Closure)
        assert rfun(cpu_r, nm + "_R") is not rfun(cpu_r, nm)


@needs_reference
@pytest.mark.parametrize("name", ["g_sqexp_2d_mu", "g_ard_d4_coincident"])
def test_patched_functions_with_the_oracle_behind_dot_call(cpu_r, name):
    check_gaussian_case(cpu_r, name, 1e-9)


@needs_reference
@pytest.mark.parametrize("name", sorted(k for k in G if k.startswith("f_")))
def test_unmodified_reference_loops_run_on_the_patched_functions(cpu_r, name):
    """norm_grad_ascent_vi / norm_grad_ascent / laplace_grad_ascent of /root/reference/R, untouched, 5 ADADELTA iterations
    on log(theta) (and the knots), then the reference's predict_* -- every objective, gradient, Newton search and
    prediction inside them now goes through r/patches.R -> .Call."""
    c = G[name]
    meta, i, o = c["meta"], c["in"], c["out"]
    cp, cf, fam = meta["cov_par"], meta["cov_fun"], meta["family"]
    I = cpu_r
    m = len(i["xu"])
    CALLS.clear()
    dth = I.apply_function(rfun(I, ".t_dtheta"), [("cov_fun", RI.from_py(cf))])
    dkn = I.apply_function(rfun(I, ".t_dknot"), [("cov_fun", RI.from_py(cf))]) if meta["knots"] else np.nan
    opt = {"maxit": int(o["iter"][0]), "delta": meta["delta"], "obj_tol": 0.0}
    common = dict(cov_par_start=cp, cov_fun=cf, dcov_fun_dtheta=dth, dcov_fun_dknot=dkn, knot_opt=np.arange(1, m + 1), xu=i["xu"],
                  xy=i["xy"], y=i["y"], mu=i["mu"], muu=i["muu"], opt=opt, verbose=False)
    if meta["model"] == "vi":
        out = rcall(I, "norm_grad_ascent_vi", **common)
    elif meta["model"] == "fic":
        out = rcall(I, "norm_grad_ascent", obj_fun=rfun(I, "obj_fun_norm"), transform=True, **common)
    else:
        sfx = {"bernoulli": "bern", "poisson": "pois"}[fam]
        out = rcall(I, "laplace_grad_ascent", ff=i["mu"].copy(), grad_loglik_fn=rfun(I, "grad_loglik_fn_" + sfx),
                    dlog_py_dff=rfun(I, "dlog_py_dff_" + sfx), d2log_py_dff=rfun(I, "d2log_py_dff_" + sfx),
                    d3log_py_dff=rfun(I, "d3log_py_dff_" + sfx), obj_fun=rfun(I, "obj_fun_" + sfx), transform=True, **common,
                    **gu.r_case_extra(c))
    # the loops really went through the patched functions: one fused gradient (or Newton search + gradient) per iteration
    it = int(o["iter"][0])
    if meta["model"] == "laplace":
        assert CALLS.get("laplace_newton", 0) >= it and CALLS.get("laplace_grad", 0) + CALLS.get("laplace_grad_knots", 0) >= it
    else:
        assert CALLS.get("gauss_obj_grad", 0) + CALLS.get("gauss_obj_grad_knots", 0) >= it and CALLS.get("gauss_obj_mats", 0) >= it
    _close(out["obj_fun"], o["obj_fun"], 1e-8, "objective trajectory")
    _close(out["cov_par_history"], o["cov_par_history"], 1e-8, "parameter trajectory")
    _close(out["xu"], o["xu_final"], 1e-8, "final knots")
    _close(out["u_mean"], o["u_mean"], 1e-7, "u_mean")
    cpf = dict(zip(cp, o["cov_par"].tolist()))
    pred = "predict_vi" if meta["model"] == "vi" else "predict_laplace"
    pr = rcall(I, pred, u_mean=o["u_mean"], u_var=o["u_var"], xu=o["xu_final"], x_pred=o["x_pred"], cov_fun=cf, cov_par=cpf,
               mu=np.full(len(o["x_pred"]), i["mu"][0]), muu=i["muu"], full_cov=False, family=fam, delta=meta["delta"])
    _close(pr["pred_mean"], o["pred_mean"], 1e-9, "pred_mean")
    _close(pr["pred_var"], o["pred_var"], 1e-8, "pred_var")
