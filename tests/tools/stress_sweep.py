"""Randomised shape sweep of the fused VI / FIC evaluations against the reduced-form oracle (test infrastructure):
chunk boundaries (n around multiples of the generator's 8192 / 9472-row chunks), m off the 128 tile, m up to 4096,
d in 1..64, with and without the knot gradient.   python tests/tools/stress_sweep.py [cases] [seed]  (GPU box)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reduced_model as red
from sparsergps_b200.context import Context

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed)
ctx = Context(0)
worst = {"obj": 0.0, "grad": 0.0, "knot": 0.0}
fails = []
N_CHOICES = [1, 127, 8191, 8192, 8193, 9472, 9473, 16384, 18945, 30000, 70001]
M_CHOICES = [1, 5, 64, 127, 128, 129, 255, 300, 513, 1000, 1024, 1025, 2048, 4096]
D_CHOICES = [1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 20]          # include/srgp.h: SRGP_MAX_D_FUSED = 20 (12 with knots)
for it in range(ncases):
    n, m, d = int(rng.choice(N_CHOICES)), int(rng.choice(M_CHOICES)), int(rng.choice(D_CHOICES))
    if n * m * m > 6e11:                       # keep the NumPy oracle within seconds
        n = max(1, int(6e11 / (m * m)))
    model = "vi" if it % 2 == 0 else "fic"
    spread = max(1.0, 0.3 * m ** (1.0 / d))
    x, xu = spread * rng.normal(size=(n, d)), spread * rng.normal(size=(m, d))
    if m > 2 and n > 3:
        xu[1] = x[2]                           # one coincident pair (quirk Q4)
    y = np.sin(x[:, 0]) + 0.3 * rng.normal(size=n)
    sigma, l, tau, delta = 1.1, rng.uniform(0.7, 1.6, d), 0.4, 1e-3
    t0 = time.perf_counter()
    ctx.set_data(x, y, None)
    knots = (it % 3 == 0) and d <= 12
    try:
        if knots:
            obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", xu, sigma, l, tau, delta, red.knot_bounds(x))
        else:
            obj, grad = ctx.gauss_obj_grad(model, "ard", xu, sigma, l, tau, delta)
    except Exception as e:                     # noqa: BLE001
        fails.append({"case": [n, m, d, model, knots], "error": str(e)})
        print("FAIL", n, m, d, model, knots, e, flush=True)
        continue
    tg = time.perf_counter() - t0
    f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
    ref = f(x, y, np.zeros(n), xu, sigma, l, tau, delta, knots=knots) if knots else f(x, y, np.zeros(n), xu, sigma, l, tau, delta)
    obj_r, g_r = ref[0], ref[1]
    names = red.theta_names("ard", d)
    gr = np.array([g_r[k] for k in names])
    e_obj = abs(obj - obj_r) / max(abs(obj_r), 1e-300)
    e_g = float(np.max(np.abs(np.asarray(grad) - gr)) / max(np.max(np.abs(gr)), 1e-300))
    e_k = 0.0
    if knots:
        kr = np.asarray(ref[2])
        e_k = float(np.max(np.abs(np.asarray(kg).reshape(kr.shape) - kr)) / max(np.max(np.abs(kr)), 1e-300))
    worst = {"obj": max(worst["obj"], e_obj), "grad": max(worst["grad"], e_g), "knot": max(worst["knot"], e_k)}
    # the float64 NumPy oracle (explicit inverses) loses ~cond(S) * 1e-13 on the knot gradient; the CUDA path solves
    # through triangular factors and sits 3 digits closer to the long-double yardstick (tests/tools/dbg_knot_m.py,
    # profiles/r01_stress_sweep.txt), so the knot bound scales with cond(S)
    Kuu, _ = red.kernel_matrix(xu, xu, sigma, l)
    cond = float(np.linalg.cond(Kuu + delta * np.eye(m))) if m <= 2048 else float("nan")
    ok = e_obj < 1e-8 and e_g < 1e-7 and e_k < max(1e-7, 1e-8 * (cond if cond == cond else 1e5))
    if not ok:
        fails.append({"case": [n, m, d, model, knots], "e_obj": e_obj, "e_grad": e_g, "e_knot": e_k})
    print("%-4s n=%6d m=%5d d=%3d %s knots=%d  cond(S)=%.0e  e_obj=%.1e e_grad=%.1e e_knot=%.1e  gpu %.0f ms" %
          ("ok" if ok else "BAD", n, m, d, model, knots, cond, e_obj, e_g, e_k, tg * 1e3), flush=True)
print(json.dumps({"cases": ncases, "seed": seed, "worst": worst, "fails": fails}))
ctx.close()
