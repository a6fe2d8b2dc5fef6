"""Debug: knot-gradient error vs m / conditioning, against float64 and long-double reduced oracles."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reduced_model as red
from sparsergps_b200.context import Context

ctx = Context(0)
def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for (n, m, d, model, ext) in [(6000, 1025, 4, "vi", True), (6000, 1024, 4, "vi", True), (6000, 1025, 4, "fic", False),
                              (6000, 1152, 4, "vi", False), (6000, 1025, 8, "vi", False), (6000, 700, 4, "vi", True), (30000, 1025, 4, "vi", False)]:
    rng = np.random.default_rng(n + m + d)
    spread = max(1.0, 0.3 * m ** (1.0 / d))
    x, xu = spread * rng.normal(size=(n, d)), spread * rng.normal(size=(m, d))
    xu[1] = x[2]
    y = np.sin(x[:, 0]) + 0.3 * rng.normal(size=n)
    sigma, l, tau, delta = 1.1, rng.uniform(0.7, 1.6, d), 0.4, 1e-3
    ctx.set_data(x, y, None)
    kb = red.knot_bounds(x)
    obj, grad, kg, _ = ctx.gauss_obj_grad_knots(model, "ard", xu, sigma, l, tau, delta, kb)
    f = red.vi_obj_grad if model == "vi" else red.fic_obj_grad
    o64, g64, k64 = f(x, y, np.zeros(n), xu, sigma, l, tau, delta, knots=True)
    names = red.theta_names("ard", d)
    K, _ = red.kernel_matrix(xu, xu, sigma, l)
    cond = np.linalg.cond(K + delta * np.eye(m))
    line = "n=%d m=%d d=%d %s cond(S)=%.1e | cuda vs f64: grad %.1e knot %.1e" % (
        n, m, d, model, cond, rel(grad, [g64[k] for k in names]), rel(kg.reshape(m, d), k64))
    kgm, k64m = kg.reshape(m, d), np.asarray(k64, dtype=np.float64)
    bad = np.unravel_index(np.argmax(np.abs(kgm - k64m)), kgm.shape)
    line += " worst at knot %d dim %d (cuda %.6e f64 %.6e)" % (bad[0], bad[1], kgm[bad], k64m[bad])
    if ext:
        t0 = time.time()
        with red.extended_precision():
            ox, gx, kx = f(x, y, np.zeros(n), xu, sigma, l, tau, delta, knots=True)
        line += " | vs longdouble: cuda grad %.1e knot %.1e ; f64 grad %.1e knot %.1e (%.0fs)" % (
            rel(grad, [gx[k] for k in names]), rel(kgm, kx), rel([g64[k] for k in names], [gx[k] for k in names]), rel(k64m, kx), time.time() - t0)
    print(line, flush=True)
ctx.close()
