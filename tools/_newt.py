import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsergps_b200 import laplace as Lp
from sparsergps_b200.context import Context
n, m, d, maxit = 100_000, 512, 8, 60
rng = np.random.default_rng(1311)
x = rng.normal(size=(n, d)); f = 1.5 * np.sin(x[:, 0]) + x[:, 1] - 0.5 * x[:, 2]
y = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(np.float64); xu = rng.normal(size=(m, d))
cp = {"sigma": 2.0}; cp.update({"l%d" % (c + 1): 1.5 for c in range(d)}); cp["tau"] = 0.1
ctx = Context(0)
for rep in range(3):
    ctx.prof_enable(True); ctx.prof_reset()
    t0 = time.perf_counter()
    res = Lp.newtrap_sparseGP(np.zeros(n), "bernoulli", cp, "ard", x, xu, y, np.zeros(n), np.zeros(m), maxit=maxit, tol=1e-5, delta=1e-3, ctx=ctx)
    dt = time.perf_counter() - t0
    it = len(res["objective_function_values"]) - 1
    print("rep", rep, it, "iters %.1f ms -> %.2f ms/iter" % (dt*1e3, dt*1e3/it), " ".join("%s=%.2fms/%d" % (k, ctx.prof_get(k)[1], ctx.prof_get(k)[0]) for k in ("gen","gram","km","dense","reduce","comm")))
