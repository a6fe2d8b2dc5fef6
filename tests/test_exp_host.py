"""Accuracy of the device exp (csrc/kmath.cuh) compiled for the host: <= 2.5e-16 relative against long-double exp,
gradual underflow reproduced."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    so = os.path.join(str(tmp_path), "host_exp.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "host_exp.cpp")])
    return ctypes.CDLL(so)


def test_exp_nonpos_accuracy(tmp_path):
    lib = _build(tmp_path)
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 745.2, 500_000), -10.0 ** rng.uniform(-20, 0, 50_000),
                        [0.0, -0.0, -745.13, -745.14, -746.0, -800.0, -1e300, -np.inf, -708.4, -709.0, -1e-320]])
    out = np.empty_like(x)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.host_exp_nonpos(x.ctypes.data_as(dp), ctypes.c_long(len(x)), out.ctypes.data_as(dp))
    ref = np.exp(x.astype(np.longdouble))
    refd = ref.astype(np.float64)
    normal = refd > 2.3e-308
    rel = np.abs((out.astype(np.longdouble) - ref) / np.where(normal, ref, 1))[normal]
    assert float(rel.max()) <= 2.5e-16            # ~1 ulp
    sub = ~normal
    assert np.max(np.abs(out[sub] - refd[sub])) <= 4.95e-324      # within one denormal step
    nan = np.array([np.nan])
    o = np.empty(1)
    lib.host_exp_nonpos(nan.ctypes.data_as(dp), ctypes.c_long(1), o.ctypes.data_as(dp))
    assert np.isnan(o[0])
