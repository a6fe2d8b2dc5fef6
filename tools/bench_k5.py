"""Kernel-only timing of the fused all-theta reduction sum Omega o dK (K5) with Omega resident in HBM.
   python tools/bench_k5.py [n] [m] [d]  -> JSON with GB/s of the 8*n*m-byte Omega read."""
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
try:
    hbm = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    src = "measured"
except Exception:
    hbm, src = 6650.0, "fallback"
ctx = Context(0)
x, u, om = ctx.dev_alloc(8 * n * d), ctx.dev_alloc(8 * m * d), ctx.dev_alloc(8 * n * m)
ctx.fill_normal(x, n * d, 1)
ctx.fill_normal(u, m * d, 2)
ctx.fill_normal(om, n * m, 3)
l = np.array([0.8 + 0.05 * (c + 1) for c in range(d)])
out = np.zeros(d + 2)
fn = lambda: ctx._lib.srgp_omega_dk_reduce_dev(ctx.handle, L.ARD, x, n, u, m, d, 1.0, L.ptr(l), 0.5, om, L.ptr(out))
for _ in range(3):
    L.check(fn())
ts = []
for _ in range(10):
    ctx.timer_start()
    L.check(fn())
    ts.append(ctx.timer_stop_ms())
ms = float(np.median(ts))
gb = 8.0 * n * m / 1e9
# FP64-pipe roofline: 2d + 10 (table exp) + 2d + 5 instructions per entry against the measured DFMA issue rate
dfma = 33.38e12 / 2          # profiles/r01_microbench.json: 33.38 TFLOP/s of DFMA = 16.7e12 lane-instructions/s
instr = 4 * d + 15
print(json.dumps({"kernel": "omega_dk_kernel (sum Omega o dK, all theta)", "n": n, "m": m, "d": d, "ms": round(ms, 4),
                  "GBps": round(gb / (ms * 1e-3), 1), "frac_of_%s_hbm" % src: round(gb / (ms * 1e-3) / hbm, 3),
                  "fp64_pipe_instr_per_entry": instr,
                  "frac_of_fp64_pipe": round(instr * n * m / (ms * 1e-3) / dfma, 3), "bound": "fp64 pipe (ridge = %d instr per 8-byte entry)" % round(dfma / (hbm * 1e9 / 8)),
                  "input_larger_than_L2": gb > 0.2}))
