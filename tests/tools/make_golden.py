"""Generate tests/golden/rcpp_layer.{json,npz}: input/output vectors of the REFERENCE's compiled Rcpp layer.

Run in the build container (needs /root/reference): `python tests/tools/make_golden.py`. It compiles the reference's
src/*.cpp in place through oracle/ref_native.py (Rcpp stand-in header, see oracle/rcpp_shim/Rcpp.h), calls every
exported function of SURVEY.md 8(b)'s routine table on seeded inputs -- configs 1 and 2 of BASELINE.json at reduced
size, the L1/L2 `exp` kernel, coincident rows, unknown names, the unreachable exp+cross+tau branch -- and stores
inputs and outputs. The GPU box has no /root/reference; tests read only the committed files.

Each case in the JSON: {"fn": name, "args": {...}, "out": key | {"derivative": key, ...}} where array arguments and
outputs are keys into the NPZ and everything else is literal.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_native as rn  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")


class Book:
    def __init__(self):
        self.arrays, self.cases = {}, []

    def arr(self, a):
        a = np.asarray(a, dtype=np.float64)
        for key, have in self.arrays.items():          # inputs recur across cases: store each array once
            if have.shape == a.shape and np.array_equal(have, a, equal_nan=True):
                return key
        key = "a%03d" % len(self.arrays)
        self.arrays[key] = a
        return key

    def call(self, fn, **args):
        real = dict(args)
        enc = {}
        for k, v in args.items():
            if isinstance(v, np.ndarray):
                enc[k] = {"npz": self.arr(v)}
            else:
                enc[k] = v
        out = getattr(rn, fn)(**real)
        if isinstance(out, dict):
            enc_out = {k: self.arr(np.atleast_1d(v)) for k, v in out.items()}
        else:
            enc_out = self.arr(np.atleast_1d(out) if np.ndim(out) == 0 else out)
        self.cases.append({"fn": fn, "args": enc, "out": enc_out})


def main():
    assert rn.sources_present(), "needs /root/reference"
    rn.build(force=True)
    b = Book()
    rng = np.random.default_rng(20261018)

    # ---- transforms (a13)
    t = np.array([-3.0, -0.5, 0.0, 0.25, 2.0, 30.0])
    b.call("real_to_pos", x=t)
    b.call("pos_to_real", x=np.exp(t))
    b.call("real_to_bounded", x=t, ub=np.full(6, 4.0), lb=np.full(6, -1.5))

    # ---- config 1 shape: 1-D sqexp, sorted U(0,10), theta = (2, 1, 1), knots (1,3,5,7,9); one x equals a knot
    x1d = np.sort(rng.uniform(0, 10, 48)).reshape(-1, 1)
    x1d[20, 0] = 5.0
    xu1d = np.array([1.0, 3.0, 5.0, 7.0, 9.0]).reshape(-1, 1)
    cp1 = {"sigma": 2.0, "l": 1.0, "tau": 1.0}
    for kern in ("sqexp", "exp"):
        b.call("make_cov_matC", x=x1d, x_pred=None, cov_par=cp1, cov_fun=kern, delta=1e-6)
        b.call("make_cov_matC", x=x1d, x_pred=xu1d, cov_par=cp1, cov_fun=kern, delta=1e-6)
        for par in ("sigma", "l", "tau"):
            b.call("dsig_dthetaC", x=x1d, x_pred=None, cov_par=cp1, cov_fun=kern, par_name=par)
            b.call("dsig_dthetaC", x=x1d, x_pred=xu1d, cov_par=cp1, cov_fun=kern, par_name=par)

    # ---- d = 3 non-ARD (the exp kernel: L1 distance in K, L2 in dK -- SURVEY a3 / a10), coincident rows
    x3 = rng.normal(size=(37, 3))
    xp3 = np.vstack([rng.normal(size=(9, 3)), x3[4], x3[30]])
    cp3 = {"tau": 0.6, "l": 1.4, "sigma": 0.8}          # list order must not matter (lookup by name)
    for kern in ("sqexp", "exp"):
        b.call("make_cov_matC", x=x3, x_pred=xp3, cov_par=cp3, cov_fun=kern, delta=1e-4)
        b.call("make_cov_matC", x=x3, x_pred=None, cov_par=cp3, cov_fun=kern, delta=1e-4)
        for par in ("sigma", "l", "tau"):
            b.call("dsig_dthetaC", x=x3, x_pred=xp3, cov_par=cp3, cov_fun=kern, par_name=par)
            b.call("dsig_dthetaC", x=x3, x_pred=None, cov_par=cp3, cov_fun=kern, par_name=par)

    # ---- config 2 shape: ARD d = 5, knots = data rows + jitter, two exact coincidences
    d = 5
    x5 = rng.normal(size=(64, d))
    xu5 = x5[rng.choice(64, 12, replace=False)] + 0.01 * rng.normal(size=(12, d))
    xu5[3] = x5[7]
    xu5[9] = x5[50]
    ln = ["l%d" % (c + 1) for c in range(d)]
    cp5 = {"sigma": 1.7}
    for c in range(d):
        cp5[ln[c]] = [0.7, 1.0, 1.6, 2.2, 0.9][c]
    cp5["tau"] = 0.45
    b.call("make_cov_mat_ardC", x=x5, x_pred=xu5, cov_par=cp5, cov_fun="ard", delta=1e-4, lnames=ln)
    b.call("make_cov_mat_ardC", x=xu5, x_pred=None, cov_par=cp5, cov_fun="ard", delta=1e-4, lnames=ln)
    for par in ["sigma"] + ln + ["tau"]:
        b.call("dsig_dtheta_ardC", x=x5, x_pred=xu5, cov_par=cp5, cov_fun="ard", par_name=par, lnames=ln)
        b.call("dsig_dtheta_ardC", x=xu5, x_pred=None, cov_par=cp5, cov_fun="ard", par_name=par, lnames=ln)

    # ---- config 5 shape: d = 8, l_c = 0.8 + 0.05 c, sigma 1, tau 0.5
    d8 = 8
    x8, xu8 = rng.normal(size=(40, d8)), rng.normal(size=(24, d8))
    ln8 = ["l%d" % (c + 1) for c in range(d8)]
    cp8 = {"sigma": 1.0}
    for c in range(d8):
        cp8[ln8[c]] = 0.8 + 0.05 * (c + 1)
    cp8["tau"] = 0.5
    b.call("make_cov_mat_ardC", x=x8, x_pred=xu8, cov_par=cp8, cov_fun="ard", delta=1e-6, lnames=ln8)
    for par in ("sigma", "l3", "l8", "tau"):
        b.call("dsig_dtheta_ardC", x=x8, x_pred=xu8, cov_par=cp8, cov_fun="ard", par_name=par, lnames=ln8)

    # ---- far-apart points (underflow to 0 / denormals), huge and tiny length scales
    xf = np.array([[0.0], [1.0], [40.0], [1e3]])
    for l in (1e-2, 1.0, 1e4):
        cpf = {"sigma": 1.0, "l": l, "tau": 0.5}
        b.call("make_cov_matC", x=xf, x_pred=None, cov_par=cpf, cov_fun="sqexp", delta=0.0)
        b.call("dsig_dthetaC", x=xf, x_pred=None, cov_par=cpf, cov_fun="sqexp", par_name="l")

    # ---- unknown names: message + 0 x 0 (and the unreachable exp + cross + unknown branch: zeros)
    b.call("make_cov_matC", x=x3, x_pred=xp3, cov_par=cp3, cov_fun="matern", delta=1e-4)
    b.call("make_cov_mat_ardC", x=x5, x_pred=xu5, cov_par=cp5, cov_fun="sqexp", delta=1e-4, lnames=ln)
    b.call("dsig_dthetaC", x=x3, x_pred=xp3, cov_par=cp3, cov_fun="sqexp", par_name="nu")
    b.call("dsig_dthetaC", x=x3, x_pred=xp3, cov_par=cp3, cov_fun="exp", par_name="nu")
    b.call("dsig_dthetaC", x=x3, x_pred=None, cov_par=cp3, cov_fun="exp", par_name="nu")
    b.call("dsig_dtheta_ardC", x=x5, x_pred=xu5, cov_par=cp5, cov_fun="ard", par_name="l9", lnames=ln)

    # ---- the 13 per-pair helpers (a1-a3, a6-a10) and the knot-location closures (8 f1)
    pairs = [(x5[0], xu5[0]), (x5[7], xu5[3]), (x5[1], x5[1] + 1e-9)]
    lb, ub = x5.min(0) - 0.5, x5.max(0) + 0.5
    cps = {"sigma": 1.7, "l": 1.3, "tau": 0.45}
    for a, c in pairs:
        b.call("cov_fun_sqrd_expC", x1=a, x2=c, cov_par=cps)
        b.call("cov_fun_expC", x1=a, x2=c, cov_par=cps)
        b.call("cov_fun_sqrd_exp_ardC", x1=a, x2=c, cov_par=cp5, lnames=ln)
        for fn in ("dsqexp_dsigmaC", "dsqexp_dlC", "dsqexp_dtauC", "dexp_dsigmaC", "dexp_dlC", "dexp_dtauC"):
            b.call(fn, x1=a, x2=c, cov_par=cps)
        b.call("dsqexp_dsigma_ardC", x1=a, x2=c, cov_par=cp5, lnames=ln)
        for comp in (1, 3, 5):
            b.call("dsqexp_dl_ardC", x1=a, x2=c, cov_par=cp5, lnames=ln, comp=comp)
    for a, c in [(x5[0], xu5[0]), (x5[10], xu5[5])]:
        b.call("dsqexp_dx2C", x1=a, x2=c, cov_par=cps, lb=lb, ub=ub)
        b.call("dsqexp_dx2_ardC", x1=a, x2=c, cov_par=cp5, lb=lb, ub=ub, lnames=ln)

    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, "rcpp_layer.npz"), **b.arrays)
    with open(os.path.join(OUT_DIR, "rcpp_layer.json"), "w") as f:
        json.dump({"generator": "tests/tools/make_golden.py",
                   "source": "reference src/covariance_functionsC.cpp + src/covariance_function_derivativesC.cpp, "
                             "compiled unmodified with g++ -O2 against oracle/rcpp_shim/Rcpp.h",
                   "cases": b.cases}, f, indent=1)
    print("wrote %d cases, %d arrays" % (len(b.cases), len(b.arrays)))


if __name__ == "__main__":
    main()
