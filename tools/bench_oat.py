"""OAT candidate scoring at the headline shape: T candidates in one bordered pass vs T objective-only evaluations.
   python tools/bench_oat.py [n] [m] [T]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sparsergps_b200.context import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 40
x, y, xu, th = workload(n, m, 8)
cand = x[np.random.default_rng(3).choice(n, T, replace=False)]
ctx = Context(0)
ctx.set_data(x, y, None)
out = {"n": n, "m": m, "T": T}
for rep in range(3):
    ctx.sync()
    t0 = time.perf_counter()
    obj0, scores = ctx.oat_scores("vi", "ard", xu, cand, th["sigma"], th["l"], th["tau"], th["delta"])
    out["bordered_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
k = min(T, 4)
full = [ctx.gauss_obj_grad("vi", "ard", np.vstack([xu, cand[t]]), th["sigma"], th["l"], th["tau"], th["delta"],
                           want_grad=False)[0] for t in range(k)]
out["per_candidate_eval_ms"] = (time.perf_counter() - t0) * 1e3 / k
out["loop_of_T_evals_ms"] = out["per_candidate_eval_ms"] * T
out["max_rel_diff_vs_full_eval"] = float(np.max(np.abs(scores[:k] - np.array(full)) / np.abs(full)))
out["speedup"] = out["loop_of_T_evals_ms"] / out["bordered_ms"]
print(json.dumps(out))
ctx.close()
