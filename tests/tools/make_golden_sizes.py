#!/usr/bin/env python
"""Goldens at the sizes BASELINE.json states -> tests/golden/stated_sizes.json.

Run in the BUILD container (CPU, minutes): the GPU box only reads the committed JSON.

    python tests/tools/make_golden_sizes.py [cfg5_vi cfg5_fic cfg4 cfg3 ...]     # default: all, merged into the file

Cases (inputs are regenerated from seeds by tests/cases.py / bench.workload, so only outputs are stored):
  cfg5_vi   n = 1,000,000, m = 1024, d = 8 (bench.py's workload): VI objective, 10 gradient components, and a
            checksum + 16 entries of the 1024 x 8 knot gradient -- oracle/reduced_model.py in 100 row shards
            (the literal transcription needs >= 100 GB there; the reduced form is proven equal to it by
            tests/test_oracle.py and is pinned to the reference's R through oracle/ref_model.py).
  cfg5_fic  n = 250,000, m = 1024, d = 8: FIC objective + gradient + knot-gradient checksum (reduced form).
  cfg4      Bernoulli, n = 100,000, d = 8, m = 512: Newton objective history and iteration count to tol 1e-5,
            posterior mean at the knots, 64 entries of the mode, and the dlogq gradient at a closed-form ff.
  cfg3_9568 / cfg3_4784   n = 9568 / 4784, d = 4, m = 256 + 1 (the extra knot is a data row): VI and FIC objective +
            gradient from the LITERAL transcription (oracle/ref_model.py) and from the long-double yardstick.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reduced_model as red          # noqa: E402
from oracle import ref_model as rm               # noqa: E402
from tests import cases                          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "stated_sizes.json")
KNOT_PROBE = [(0, 0), (1, 3), (17, 7), (100, 2), (255, 5), (256, 0), (511, 6), (512, 1), (700, 4), (1023, 7),
              (1023, 0), (33, 3), (64, 6), (901, 2), (400, 5), (5, 1)]


def knot_summary(gk):
    gk = np.asarray(gk, dtype=np.float64)
    probe = [(k, c) for k, c in KNOT_PROBE if k < gk.shape[0] and c < gk.shape[1]]
    return {"shape": list(gk.shape), "sum": float(gk.sum()), "abs_sum": float(np.abs(gk).sum()),
            "fro": float(np.linalg.norm(gk)), "probe_idx": probe, "probe": [float(gk[k, c]) for k, c in probe]}


def cfg4_ff_closed_form(x):
    """A smooth stand-in for the mode: the gradient formula is defined for any ff, and a closed form needs no fixture."""
    return 1.5 * np.sin(x[:, 0]) + x[:, 1] - 0.5 * x[:, 2]


def gen_cfg5_vi():
    import bench
    n, m, d = 1_000_000, 1024, 8
    x, y, xu, th = bench.workload(n, m, d)
    obj, g, gk = red.vi_obj_grad(x, y, np.zeros(1), xu, th["sigma"], th["l"], th["tau"], th["delta"], shards=100, knots=True)
    names = red.theta_names("ard", d)
    return {"n": n, "m": m, "d": d, "source": "oracle/reduced_model.vi_obj_grad, 100 row shards, float64",
            "obj": float(obj), "grad": [float(g[k]) for k in names], "names": names, "knot": knot_summary(gk)}


def gen_cfg5_fic():
    import bench
    n, m, d = 250_000, 1024, 8
    x, y, xu, th = bench.workload(n, m, d)
    obj, g, gk = red.fic_obj_grad(x, y, np.zeros(1), xu, th["sigma"], th["l"], th["tau"], th["delta"], shards=25, knots=True)
    names = red.theta_names("ard", d)
    return {"n": n, "m": m, "d": d, "source": "oracle/reduced_model.fic_obj_grad, 25 row shards, float64",
            "obj": float(obj), "grad": [float(g[k]) for k in names], "names": names, "knot": knot_summary(gk)}


def gen_cfg4():
    n, m, d = 100_000, 512, 8
    c = cases.config4(n=n, d=d, m=m)
    cp = c["cov_par"]
    l = cases.lvec(cp)
    fit = red.laplace_newton(c["x"], c["y"], c["mu"], np.zeros(m), c["xu"], cp["sigma"], l, cp["tau"], c["delta"],
                             "bernoulli", np.zeros(n), maxit=1000, tol=1e-5)
    ff = cfg4_ff_closed_form(c["x"])
    g = red.laplace_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], "bernoulli", ff)
    g_mode = red.laplace_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], "bernoulli", fit["gp"])
    names = red.theta_names("ard", d)
    return {"n": n, "m": m, "d": d, "family": "bernoulli", "tol": 1e-5, "maxit": 1000,
            "source": "oracle/reduced_model.laplace_newton / laplace_grad, float64",
            "hist": [float(v) for v in fit["hist"]], "iterations": int(len(fit["hist"])),
            "gp_head": [float(v) for v in fit["gp"][:64]], "gp_norm": float(np.linalg.norm(fit["gp"])),
            "gp_sum": float(fit["gp"].sum()),
            "u_mean": [float(v) for v in fit["u_mean"]], "u_var_diag": [float(v) for v in np.diag(fit["u_var"])],
            "grad_at_closed_form_ff": [float(g[k]) for k in names],
            "grad_at_mode": [float(g_mode[k]) for k in names], "names": names}


def gen_cfg3(n):
    c = cases.config3(n=n)                       # d = 4, m = 256 + 1 data-row knot
    cp = c["cov_par"]
    l = cases.lvec(cp)
    out = {"n": n, "m": int(len(c["xu"])), "d": 4, "names": list(cp),
           "source": "literal: oracle/ref_model.{vi,fic}_obj_grad (float64); yardstick: oracle/reduced_model in numpy.longdouble"}
    for model, lit, ext in (("vi", rm.vi_obj_grad, red.vi_obj_grad), ("fic", rm.fic_obj_grad, red.fic_obj_grad)):
        o, g = lit(cp, c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], c["delta"])
        with red.extended_precision():
            ox, gx = ext(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"])
        out[model] = {"literal_obj": float(o), "literal_grad": [float(g[k]) for k in cp],
                      "longdouble_obj": float(ox), "longdouble_grad": [float(gx[k]) for k in cp]}
    return out


GENERATORS = {"cfg5_vi": gen_cfg5_vi, "cfg5_fic": gen_cfg5_fic, "cfg4": gen_cfg4,
              "cfg3_9568": lambda: gen_cfg3(9568), "cfg3_4784": lambda: gen_cfg3(4784)}


def main():
    which = sys.argv[1:] or list(GENERATORS)
    book = {}
    if os.path.exists(OUT):
        book = json.load(open(OUT))
    for nm in which:
        t0 = time.time()
        book[nm] = GENERATORS[nm]()
        book[nm]["seconds_to_generate"] = round(time.time() - t0, 1)
        print(nm, "done in %.1f s" % (time.time() - t0), flush=True)
        with open(OUT, "w") as f:
            json.dump(book, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
