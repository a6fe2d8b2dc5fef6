"""NumPy statement of the error-free INT8 splitting used by the tensor-core row passes -- TEST INFRASTRUCTURE.

Restates, with plain integer arithmetic, what csrc/tc_i8.cuh / csrc/gauss_i8.cu do on the INT8 tensor cores
(DESIGN.md section 3a), so that the scheme itself -- digit extraction, level sums, overflow bound, dropped pairs,
level weights -- is checked on the CPU, independently of any kernel:

    q = rint(v 2^(8 NS - 2))         fixed point of v in [-1, 1]: NS = 7 slices (54 bits, the product's default) or 8 (62)
    q = sum_t d_t 256^t              balanced digits d_t in [-128, 127]; slice s = NS - 1 - t
    bytes of (q + B) ^ B, B = 0x80..80 (NS bytes)   are exactly those digits (two's complement INT8)
    sum_r a_r b_r ~= sum_{sa+sb<NS} 2^(-12-8(sa+sb)) sum_r da_sa[r] db_sb[r]

There is no reference-file citation here: the reference computes these products in FP64 (R's %*%); this module is
the checker of OUR replacement for that arithmetic.  Only tests/ import it.
"""
from __future__ import annotations

import numpy as np

NS = 7                                  # SRGP_I8_NS of csrc/tc_i8.cuh; set_slices(8) models the validation build
BIAS = np.uint64(0x0080808080808080)


def set_slices(ns):
    global NS, BIAS
    assert ns in (7, 8)
    NS = ns
    BIAS = np.uint64(0x8080808080808080 if ns == 8 else 0x0080808080808080)


def fixed_point(v):
    v = np.asarray(v, dtype=np.float64)
    assert np.all(np.abs(v) <= 1.0)
    return np.rint(v * 2.0 ** (8 * NS - 2)).astype(np.int64)


def digits_carry_chain(q):
    """Balanced base-256 digits by the textbook carry chain: out[..., s], slice s = NS - 1 - t."""
    q = np.array(q, dtype=np.int64, copy=True)
    out = np.zeros(q.shape + (NS,), dtype=np.int8)
    for t in range(NS):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        out[..., NS - 1 - t] = d.astype(np.int8)
    assert np.all(q == 0)
    return out


def digits_bias_trick(q):
    """The same digits as the bytes of (q + B) ^ B (what split_quad does on the GPU)."""
    y = (np.asarray(q, dtype=np.int64).astype(np.uint64) + BIAS) ^ BIAS
    by = y[..., None] >> (np.arange(NS, dtype=np.uint64) * np.uint64(8)) & np.uint64(255)      # byte t
    return by.astype(np.uint8).view(np.int8)[..., ::-1].copy()                                   # slice s = NS - 1 - t


def join(digits):
    """Inverse: exact q from the slices (what join_quad does)."""
    d = np.asarray(digits, dtype=np.int64)
    q = np.zeros(d.shape[:-1], dtype=np.int64)
    for s in range(NS):
        q = q * 256 + d[..., s]
    return q


def level_sums(da, db):
    """INT32 level accumulators of C = A B^T: da [M, K, NS], db [N, K, NS] -> lev [NS, M, N] (Python ints would also do;
    int64 here, with the INT32 overflow bound asserted)."""
    M, K, _ = da.shape
    lev = np.zeros((NS, M, db.shape[0]), dtype=np.int64)
    for sa in range(NS):
        for sb in range(NS - sa):
            lev[sa + sb] += da[:, :, sa].astype(np.int64) @ db[:, :, sb].astype(np.int64).T
    assert np.max(np.abs(lev)) < 2 ** 31, "INT32 accumulator overflow: more than 8192 rows per accumulation?"
    return lev


def combine(lev, scale=1.0):
    """FP64 value of the levels as the kernels form it: two exact 64-bit integers (levels 0..3, 4..NS-1), then
    2^(-12-8(NS-1)) lo + 2^-36 hi, least significant first."""
    hi = ((lev[0] * 256 + lev[1]) * 256 + lev[2]) * 256 + lev[3]
    lo = lev[4]
    for L in range(5, NS):
        lo = lo * 256 + lev[L]
    return scale * (2.0 ** (-12 - 8 * (NS - 1)) * lo.astype(np.float64) + 2.0 ** -36 * hi.astype(np.float64))


def matmul_nt(a, b):
    """A B^T for a [M, K], b [N, K] with entries in [-1, 1], by the INT8 scheme (rows of K <= 8192)."""
    return combine(level_sums(digits_bias_trick(fixed_point(a)), digits_bias_trick(fixed_point(b))))


def row_scales(m):
    """Per-row power-of-two scale of slice_mop_kernel: max |row| / 2^e in [0.5, 1)."""
    mx = np.max(np.abs(m), axis=1)
    e = np.where((mx > 0) & np.isfinite(mx), np.floor(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)
    return 2.0 ** e
