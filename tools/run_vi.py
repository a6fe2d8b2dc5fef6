"""Run the fused VI objective+gradient a few times on synthetic data (profiling driver for ncu).
   python tools/run_vi.py [n] [m] [d] [reps] [model] [knots: 0|1]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sparsergps_b200.context import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
model = sys.argv[5] if len(sys.argv) > 5 else "vi"
knots = len(sys.argv) > 6 and sys.argv[6] == "1"
x, y, xu, th = workload(n, m, d)
ctx = Context(0)
ctx.set_data(x, y, None)
ctx.prof_enable(True)
for r in range(reps):
    ctx.prof_reset()
    t0 = time.perf_counter()
    if knots:
        obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"], None)
    else:
        obj, grad = ctx.gauss_obj_grad(model, "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"])
    dt = time.perf_counter() - t0
    print("rep %d: %.2f ms  obj %.6f |g| %.4e  " % (r, dt * 1e3, obj, np.linalg.norm(grad)) +
          " ".join("%s=%.2fms/%d" % (k, ctx.prof_get(k)[1], ctx.prof_get(k)[0]) for k in ("gen", "gram", "km", "dense", "reduce")))
ctx.close()
