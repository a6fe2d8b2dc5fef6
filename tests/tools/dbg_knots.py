import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reduced_model as red
from tests import cases
from sparsergps_b200.context import Context
from sparsergps_b200 import vi_functions as vf
ctx = Context(0)
for (n, m, seed) in ((20000, 1024, 1312), (11003, 130, 77), (3000, 300, 1312)):
    c = cases.config5(n=n, m=m, seed=seed)
    cp = c["cov_par"]
    ctx.set_data(c["x"], c["y"], None)
    kb = vf.knot_bounds(c["x"])
    for model in ("vi", "fic"):
        obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"], kb)
        f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
        _, _, ref = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"], knots=True)
        kg = kg.reshape(m, -1)
        err = np.abs(kg - ref)
        sc = np.nanmax(np.abs(ref))
        rel = err / (np.abs(ref) + 1e-300)
        i = np.unravel_index(np.nanargmax(err), err.shape)
        print(n, m, model, "scale %.3e max abs err/scale %.2e  max elementwise rel %.2e  at %s ref %.3e ours %.3e nan %d" % (
            sc, np.nanmax(err) / sc, np.nanmax(rel), i, ref[i], kg[i], np.isnan(kg).sum()))
        if m <= 300:
            with red.extended_precision():
                _, _, refx = f(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], cases.lvec(cp), cp["tau"], c["delta"], knots=True)
            refx = refx.astype(np.float64)
            print("    vs longdouble: ours %.2e   numpy-reduced %.2e" % (np.nanmax(np.abs(kg - refx)) / sc, np.nanmax(np.abs(ref - refx)) / sc))
