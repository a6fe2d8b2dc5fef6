"""The reference's OWN R sources, executed here -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`session()` sources /root/reference/R/*.R, unmodified and where they lie, into the mini-R interpreter of
oracle/mini_r/ (no R installation exists in this image). `R/RcppExports.R` is sourced like every other file; its
`.Call('_sparseRGPs_<name>', PACKAGE = 'sparseRGPs', ...)` lines are served by the reference's compiled C++
(oracle/ref_native.py -> oracle/_ref/libsparseRGPs_ref.so). The helpers below call the reference's R functions with
NumPy arguments and return NumPy results: they are what pins oracle/ref_model.py (tests/test_reference_r.py) and what
tests/tools/make_golden_r.py records into tests/golden/r_level.*.

Needs /root/reference (build container only). Only tests/ and tests/tools/ import this.
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import ref_native as rn
from .mini_r import interp as RI

REF_R_DIR = "/root/reference/R"


def available() -> bool:
    return os.path.isdir(REF_R_DIR) and rn.sources_present()


def _cov_par(lst):
    return {nm: float(RI.as_float(v)[0]) for nm, v in zip(lst.names, lst.items)}


def _mat(x):
    return RI.matrix_of(x)


def _xpred(x):
    a = RI.as_float(x)
    if len(a) >= 1 and np.isnan(a[0]):
        return None                      # matrix(): the 1 x 1 NA sentinel (src/covariance_functionsC.cpp:81)
    return RI.matrix_of(x)


def _str(x):
    return str(x.v[0])


def _pair_list(d):
    def conv(v):
        return RI.Vec(np.atleast_1d(np.asarray(v, dtype=np.float64)).copy())
    return RI.RList([conv(d["derivative"]), conv(d["trans_par"]), conv(d["inv_trans_par"])],
                    ["derivative", "trans_par", "inv_trans_par"])


def _dot_call(I, pos, kw):
    """.Call('_sparseRGPs_<name>', PACKAGE = 'sparseRGPs', ...) -> the reference's compiled C++ (src/RcppExports.cpp)."""
    name = _str(pos[0]).replace("_sparseRGPs_", "")
    a = pos[1:]
    V = RI.as_float
    if name in ("real_to_pos", "pos_to_real"):
        return RI.Vec(getattr(rn, name)(V(a[0])))
    if name == "real_to_bounded":
        return RI.Vec(rn.real_to_bounded(V(a[0]), V(a[1]), V(a[2])))
    if name == "make_cov_matC":
        return RI.from_matrix(rn.make_cov_matC(_mat(a[0]), _xpred(a[1]), _cov_par(a[2]), _str(a[3]), float(V(a[4])[0])))
    if name == "make_cov_mat_ardC":
        return RI.from_matrix(rn.make_cov_mat_ardC(_mat(a[0]), _xpred(a[1]), _cov_par(a[2]), _str(a[3]),
                                                   float(V(a[4])[0]), [str(s) for s in a[5].v]))
    if name == "dsig_dthetaC":
        return RI.from_matrix(rn.dsig_dthetaC(_mat(a[0]), _xpred(a[1]), _cov_par(a[2]), _str(a[3]), _str(a[4])))
    if name == "dsig_dtheta_ardC":
        return RI.from_matrix(rn.dsig_dtheta_ardC(_mat(a[0]), _xpred(a[1]), _cov_par(a[2]), _str(a[3]), _str(a[4]),
                                                  [str(s) for s in a[5].v]))
    if name in ("cov_fun_sqrd_expC", "cov_fun_expC"):
        return RI.dbl(getattr(rn, name)(V(a[0]), V(a[1]), _cov_par(a[2])))
    if name == "cov_fun_sqrd_exp_ardC":
        return RI.dbl(rn.cov_fun_sqrd_exp_ardC(V(a[0]), V(a[1]), _cov_par(a[2]), [str(s) for s in a[3].v]))
    if name in ("dsqexp_dsigmaC", "dsqexp_dlC", "dsqexp_dtauC", "dexp_dsigmaC", "dexp_dlC", "dexp_dtauC"):
        return _pair_list(getattr(rn, name)(V(a[0]), V(a[1]), _cov_par(a[2])))
    if name == "dsqexp_dsigma_ardC":
        return _pair_list(rn.dsqexp_dsigma_ardC(V(a[0]), V(a[1]), _cov_par(a[2]), [str(s) for s in a[3].v]))
    if name == "dsqexp_dl_ardC":
        return _pair_list(rn.dsqexp_dl_ardC(V(a[0]), V(a[1]), _cov_par(a[2]), [str(s) for s in a[3].v], float(V(a[4])[0])))
    if name == "dsqexp_dx2C":
        return _pair_list(rn.dsqexp_dx2C(V(a[0]), V(a[1]), _cov_par(a[2]), V(a[3]), V(a[4])))
    if name == "dsqexp_dx2_ardC":
        return _pair_list(rn.dsqexp_dx2_ardC(V(a[0]), V(a[1]), _cov_par(a[2]), V(a[3]), V(a[4]), [str(s) for s in a[5].v]))
    raise RI.RError("\"%s\" not available for .Call() for package \"sparseRGPs\"" % name)


_session = None


def session():
    """One interpreter with every R file of the reference sourced (like `library(sparseRGPs)`)."""
    global _session
    if _session is None:
        if not available():
            raise RuntimeError("/root/reference is absent: the R-level reference can only run in the build container")
        I = RI.Interp()
        I.globalenv.vars[".Call"] = RI.Builtin(".Call", lambda I_, pos, kw: _dot_call(I_, pos, {}))
        for path in sorted(glob.glob(os.path.join(REF_R_DIR, "*.R"))):
            I.source(path)
        _session = I
    return _session


def call(fname, **kwargs):
    """Call the reference's R function `fname` with NumPy / dict / str arguments; returns NumPy / dict."""
    I = session()
    f = I.get_fun(fname, I.globalenv)
    args = [(k, v if isinstance(v, (RI.Vec, RI.RList, RI.Closure, RI.Builtin)) else RI.from_py(v)) for k, v in kwargs.items()]
    return RI.to_py(I.apply_function(f, args))


def rfun(name):
    I = session()
    return I.get_fun(name, I.globalenv)


def rvalue(src):
    """Evaluate an R expression in the session (e.g. the dcov_fun_dtheta lists of R/optimize_gp.R:240-262)."""
    return session().run(src)


def na_matrix():
    return RI.Vec(np.array([np.nan]), dim=(1, 1))
