import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import reduced_model as red
from tests import cases
n, m, d = 100000, 512, 8
c = cases.config4(n=n, d=d, m=m); cp = c["cov_par"]; l = cases.lvec(cp)
ff = 1.5 * np.sin(c["x"][:, 0]) + c["x"][:, 1] - 0.5 * c["x"][:, 2]
t0 = time.time()
g64 = red.laplace_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], "bernoulli", ff)
print("f64", time.time() - t0, flush=True)
# u_mean three ways at this ff
st = red.laplace_setup(c["x"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"])
a = st["K"].T @ ((ff - c["mu"]) / st["Z"])
um_oracle = a - st["GZ"] @ (st["CZ"] @ a)
import scipy.linalg as sla
L = np.linalg.cholesky(st["S"] + st["GZ"])
um_stable = st["S"] @ sla.cho_solve((L, True), a)
with red.extended_precision():
    t0 = time.time()
    stx = red.laplace_setup(c["x"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"])
    ax = stx["K"].T @ ((np.asarray(ff, dtype=np.longdouble) - c["mu"]) / stx["Z"])
    um_x = ax - stx["GZ"] @ (stx["CZ"] @ ax)
    print("ld setup", time.time() - t0, flush=True)
    sc = float(np.max(np.abs(um_x)))
    print("u_mean: oracle-formula f64 vs ld", float(np.max(np.abs(um_oracle - um_x))) / sc, " stable f64 vs ld", float(np.max(np.abs(um_stable - um_x))) / sc, flush=True)
    gx = red.laplace_grad(c["x"], c["y"], c["mu"], c["xu"], cp["sigma"], l, cp["tau"], c["delta"], "bernoulli", ff)
    print("ld grad", time.time() - t0, flush=True)
names = red.theta_names("ard", d)
out = {"names": names, "grad_f64": [float(g64[k]) for k in names], "grad_ld": [float(gx[k]) for k in names],
       "um_oracle_err": float(np.max(np.abs(um_oracle - um_x))) / sc, "um_stable_err": float(np.max(np.abs(um_stable - um_x))) / sc,
       "u_mean_ld": [float(v) for v in um_x]}
json.dump(out, open('/tmp/cfg4_ld.json', 'w'))
print("rel f64 vs ld", max(abs(a - b) / abs(b) for a, b in zip(out["grad_f64"], out["grad_ld"])))
