"""Kernel-only timing of K1/K2 (materialising assembly) with inputs resident in HBM.
   python tools/bench_k1.py [n] [m] [d]  -> JSON lines with GB/s against MEASURED_PEAKS.json's hbm_gbs."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsergps_b200 import _lib as L
from sparsergps_b200.context import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = int(sys.argv[3]) if len(sys.argv) > 3 else 8
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
ctx = Context(0)
x = ctx.dev_alloc(8 * n * d)
u = ctx.dev_alloc(8 * m * d)
out = ctx.dev_alloc(8 * n * m)
ctx.fill_normal(x, n * d, 1312)
ctx.fill_normal(u, m * d, 1313)
l = np.array([0.8 + 0.05 * (c + 1) for c in range(d)])
lib = ctx._lib
for name, fn in (("make_cov_mat_ard", lambda: lib.srgp_make_cov_mat_dev(ctx.handle, L.ARD, x, n, u, m, d, 1.0, L.ptr(l), 0.5, 1e-6, out)),
                 ("dsig_dtheta_ard_l3", lambda: lib.srgp_dsig_dtheta_dev(ctx.handle, L.ARD, L.PAR_LC, 2, x, n, u, m, d, 1.0, L.ptr(l), 0.5, out)),
                 ("dsig_dtheta_ard_tau", lambda: lib.srgp_dsig_dtheta_dev(ctx.handle, L.ARD, L.PAR_TAU, 0, x, n, u, m, d, 1.0, L.ptr(l), 0.5, out))):
    for _ in range(3):
        L.check(fn())
    ctx.sync()
    ts = []
    for _ in range(10):
        ctx.timer_start()
        L.check(fn())
        ts.append(ctx.timer_stop_ms())
    ms = float(np.median(ts))
    gb = 8.0 * n * m / 1e9
    # FP64-pipe roofline: 3d (difference, scale, square-accumulate) + 10 (table exp, fastexp.cuh) + 2..4 instructions per entry
    dfma = 33.38e12 / 2      # profiles/r01_microbench.json: 16.7e12 DFMA lane-instructions/s
    instr = 0 if "tau" in name else (3 * d + 12 + (2 if "l3" in name else 0))
    print(json.dumps({"kernel": name, "n": n, "m": m, "d": d, "ms": round(ms, 4), "GBps": round(gb / (ms * 1e-3), 1),
                      "frac_of_measured_hbm": round(gb / (ms * 1e-3) / hbm, 3), "fp64_pipe_instr_per_entry": instr,
                      "frac_of_fp64_pipe": round(instr * n * m / (ms * 1e-3) / dfma, 3),
                      "bound": "hbm write" if instr == 0 else "fp64 pipe (ridge = %d instr per 8-byte entry)" % round(dfma / (hbm * 1e9 / 8)),
                      "output_larger_than_L2": gb > 0.2}))
