"""Diagnostic: per-component errors of the CUDA product, the float64 reduced oracle and the literal oracle, all
against the extended-precision (long double) reduced oracle.  Shows which differences are conditioning (H4)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_model as rm, reduced_model as red
from tests import cases
from sparsergps_b200.context import Context
ctx = Context(0)
for name, c in (("config3", cases.config3(n=2000, m=200)), ("config5", cases.config5(n=3000, m=300)), ("config1", cases.config1())):
    cp = c["cov_par"]; l = cases.lvec(cp)
    ctx.set_data(c["x"], c["y"], c["mu"])
    for model, f in (("vi", rm.vi_obj_grad), ("fic", rm.fic_obj_grad)):
        obj, g = ctx.gauss_obj_grad(model, c["cov_fun"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"])
        o_ref, g_ref = f(cp, c["cov_fun"], c["xu"], c["x"], c["y"], c["mu"], c["delta"])
        fr = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
        o64, g64 = fr(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], cov_fun=c["cov_fun"])
        with red.extended_precision():
            old, gld = fr(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], cov_fun=c["cov_fun"])
        e = lambda a, b: abs(float(a) / float(b) - 1)
        print("%s %s  max grad rel err vs long double:  cuda %.1e  reduced64 %.1e  literal %.1e   | obj: cuda %.1e literal %.1e" % (
            name, model, max(e(g[i], gld[k]) for i, k in enumerate(cp)), max(e(g64[k], gld[k]) for k in cp),
            max(e(g_ref[k], gld[k]) for k in cp), e(obj, old), e(o_ref, old) if np.isfinite(o_ref) else np.nan))
