"""Driver for compute-sanitizer (memcheck / racecheck) over the hand-rolled mbarrier / TMEM protocols at ragged sizes:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py 11003 1025
    compute-sanitizer --tool racecheck python tools/sanitize_run.py 97 130

One evaluation each of: VI objective+gradient (INT8 Gram + INT8 K*M), VI with the knot gradient (DMMA epilogue), FIC
(weighted INT8 Grams, ROWD mode), a short sparse-Laplace Newton search + gradient (Gram over the materialised K), OAT
candidate scoring and prediction.  n and m are deliberately not multiples of any tile."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sparsergps_b200 import laplace as Lp
from sparsergps_b200 import vi_functions as vf
from sparsergps_b200.context import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 11003
m = int(sys.argv[2]) if len(sys.argv) > 2 else 130
what = sys.argv[3] if len(sys.argv) > 3 else "all"
d = 8
x, y, xu, th = workload(n, m, d)
cp = {"sigma": th["sigma"], **{"l%d" % (c + 1): float(th["l"][c]) for c in range(d)}, "tau": th["tau"]}
ctx = Context(0)
ctx.set_data(x, y, None)
args = ("ard", xu, th["sigma"], th["l"], th["tau"], 1e-4)
if what in ("all", "vi"):
    obj, grad = ctx.gauss_obj_grad("vi", *args)
    print("vi", obj, float(np.linalg.norm(grad)), flush=True)
if what in ("all", "knots"):
    obj, grad, kg, _ = ctx.gauss_obj_grad_knots("vi", *args, vf.knot_bounds(x))
    print("vi+knots", obj, float(np.linalg.norm(kg)), flush=True)
if what in ("all", "fic"):
    obj, grad = ctx.gauss_obj_grad("fic", *args)
    print("fic", obj, float(np.linalg.norm(grad)), flush=True)
if what in ("all", "laplace"):
    yb = (y > np.median(y)).astype(np.float64)
    fit = Lp.newtrap_sparseGP(np.zeros(n), "bernoulli", cp, "ard", x, xu, yb, np.zeros(n), np.zeros(m), maxit=4, tol=1e-5, delta=1e-3, ctx=ctx)
    g = Lp.dlogq_dcov_par(cp, "ard", xu, x, yb, fit["gp"], "bernoulli", np.zeros(n), 1e-3, ctx=ctx)
    print("laplace", fit["objective_function_values"][-1], g["gradient"]["sigma"], flush=True)
if what in ("all", "oat"):
    ctx.set_data(x, y, None)
    obj0, sc = ctx.oat_scores("vi", "ard", xu, x[:5], th["sigma"], th["l"], th["tau"], 1e-4)
    print("oat", obj0, sc[:2], flush=True)
ctx.close()
print("sanitize_run done", n, m, what)
