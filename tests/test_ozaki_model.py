"""The INT8 error-free splitting of the tensor-core row passes, checked on the CPU with plain
integer arithmetic (oracle/ozaki_model.py): digit extraction by the byte-bias trick equals the carry chain, the
split is invertible, the INT32 accumulators cannot overflow within the kernels' row limit, and the combined value
agrees with a long-double product to better than FP64 rounding of the operands."""
import numpy as np
import pytest

from oracle import ozaki_model as oz


@pytest.fixture(autouse=True, params=[7, 8], ids=["ns7", "ns8"])
def slices(request):
    """7 = the product's default (csrc/tc_i8.cuh SRGP_I8_NS), 8 = the validation build."""
    oz.set_slices(request.param)
    yield request.param
    oz.set_slices(7)


def test_bias_trick_digits_equal_the_carry_chain_and_invert():
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.uniform(-1, 1, 5000), np.exp(-rng.uniform(0, 40, 5000)), [0.0, 1.0, -1.0, 2.0 ** -(8 * oz.NS - 2), -2.0 ** -(8 * oz.NS - 2),
                        0.5, 127 / 256, 128 / 256, -128 / 256]])
    q = oz.fixed_point(v)
    a, b = oz.digits_carry_chain(q), oz.digits_bias_trick(q)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(oz.join(b), q)
    assert b.min() >= -128 and b.max() <= 127 and abs(int(b[..., 0].max())) <= 64      # top slice: |d| <= 64
    assert oz.join(oz.digits_bias_trick(oz.fixed_point(np.array([1.0]))))[0] == 1 << (8 * oz.NS - 2)   # the coincidence marker of pass 2


def test_fixed_point_is_at_least_double_precision_for_large_entries():
    """Exact (the double itself) for v >= 2^-9 with 8 slices, v >= 2^-1 ... 2^-2 with 7; absolute 2^-(bits+1) below."""
    rng = np.random.default_rng(2)
    bits = 8 * oz.NS - 2
    v = np.exp(-rng.uniform(0, 6 if oz.NS == 8 else 1.38, 10000))                     # >= 2^-9 / >= 2^-2
    q = oz.fixed_point(v)
    np.testing.assert_array_equal(q.astype(np.float64) * 2.0 ** -bits, v)             # exact: 53 bits fit below 2^bits
    tiny = np.exp(-rng.uniform(2, 60, 1000))
    assert np.max(np.abs(oz.fixed_point(tiny).astype(np.float64) * 2.0 ** -bits - tiny)) <= 2.0 ** -(bits + 1)


def test_accumulators_stay_below_2_to_31_at_the_row_limit():
    # worst case: every digit at its extreme, 8192 rows, NS pairs on the last level
    K = 8192
    da = np.full((1, K, oz.NS), -128, dtype=np.int8)
    da[:, :, 0] = 64
    lev = oz.level_sums(da, da)
    assert np.max(np.abs(lev)) < 2 ** 31
    assert int(lev[oz.NS - 1, 0, 0]) == 2 * 64 * -128 * K + (oz.NS - 2) * 128 * 128 * K


@pytest.mark.parametrize("shape", [(16, 8192, 24), (33, 1024, 17)])
def test_gram_and_product_match_long_double(shape):
    M, K, N = shape
    rng = np.random.default_rng(3)
    a = np.exp(-12 * rng.uniform(0, 1, (M, K)) ** 2)           # the profile of K / sigma^2 at the headline config
    b = np.exp(-12 * rng.uniform(0, 1, (N, K)) ** 2)
    got = oz.matmul_nt(a, b)
    ref = (a.astype(np.longdouble) @ b.astype(np.longdouble).T)
    err = np.max(np.abs(got - ref.astype(np.float64)) / np.abs(ref.astype(np.float64)))
    plain = np.max(np.abs(a @ b.T - ref.astype(np.float64)) / np.abs(ref.astype(np.float64)))
    assert err < 4e-16 and err <= 2 * max(plain, 1.2e-16)                              # as good as a correctly rounded FP64 product
    # Gram: exactly symmetric, as the kernels' diagonal tiles rely on
    g = oz.matmul_nt(a, a)
    np.testing.assert_array_equal(g, g.T)


def test_signed_operand_with_row_scales_and_cancellation():
    """Pass 2: T = K Mop^T with Mop rows of very different magnitude and mixed signs, scaled per row by a power of two."""
    rng = np.random.default_rng(4)
    K, m = 1024, 24
    k = np.exp(-8 * rng.uniform(0, 1, (40, K)) ** 2)
    mop = rng.normal(size=(m, K)) * 10.0 ** rng.uniform(-6, 6, (m, 1))
    sc = oz.row_scales(mop)
    assert np.all(np.max(np.abs(mop), axis=1) / sc < 1.0) and np.all(np.max(np.abs(mop), axis=1) / sc >= 0.5)
    got = oz.matmul_nt(k, mop / sc[:, None]) * sc[None, :]
    ref = (k.astype(np.longdouble) @ mop.astype(np.longdouble).T).astype(np.float64)
    bound = (np.abs(k) @ np.abs(mop).T)                                                # the natural scale of each sum
    assert np.max(np.abs(got - ref) / bound) < 1e-15
