"""Time the K-chunk generators alone (no DMMA kernels in between).  python tools/bench_gen.py [n] [m] [d]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sparsergps_b200 import _lib as L
from sparsergps_b200.context import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = int(sys.argv[3]) if len(sys.argv) > 3 else 8
x, y, xu, th = workload(n, m, d)
ctx = Context(0)
ctx.set_data(x, y, None)
ms = np.zeros(2)
lv = L.fvec(th["l"])
L.check(ctx._lib.srgp_test_gen(ctx.handle, L.ptr(L.fmat(xu)), m, th["sigma"], L.ptr(lv), 5, L.ptr(ms)))
el = n * m
print("gen_rowmajor: %.3f ms (%.1f Gelem/s, %.2f TB/s written)   gen_colmajor: %.3f ms (%.1f Gelem/s)" %
      (ms[0], el / ms[0] / 1e6, 8 * el / ms[0] / 1e9, ms[1], el / ms[1] / 1e6))
