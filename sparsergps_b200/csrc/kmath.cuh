// kmath.cuh -- FP64 exp for the covariance kernels.
//
// Every exponent on this path is <= 0 (-r^2/(2 l^2), -sum((d/l)^2)/2, -r/l), and exp is evaluated once per
// covariance entry, so it sits on the FP64 pipe next to the DMMA work.  exp_nonpos uses a 64-entry table of
// 2^(j/64) with a degree-5 polynomial on |r| <= ln2/128: 10 FP64 instructions and one cached 8-byte load,
// about 1 ulp (measured <= 2.3e-16 relative against long-double exp, tests/test_exp_host.py).
// Gradual underflow is reproduced (results below 2^-1022 are denormal, below exp(-745.14) zero) so entries
// agree with the reference's libm exp at the stated 1e-10 relative tolerance all the way down.
//
// The same source compiles for the host (g++ -DSRGP_KMATH_HOST) so the accuracy test runs without a GPU.
#pragma once
#include <stdint.h>

#ifdef SRGP_KMATH_HOST
#include <cmath>
#include <cstring>
#define SRGP_HD inline
#define SRGP_FMA(a, b, c) std::fma((a), (b), (c))
static const double srgp_exp_tab[64] = {
#include "exp_table.inc"
};
#define SRGP_TAB(j) srgp_exp_tab[(j)]
static inline int64_t srgp_d2ll(double x) { int64_t v; std::memcpy(&v, &x, 8); return v; }
static inline double srgp_ll2d(int64_t v) { double x; std::memcpy(&x, &v, 8); return x; }
#else
#define SRGP_HD __device__ __forceinline__
#define SRGP_FMA(a, b, c) fma((a), (b), (c))
__device__ const double srgp_exp_tab[64] = {
#include "exp_table.inc"
};
#define SRGP_TAB(j) __ldg(&srgp_exp_tab[(j)])
#define srgp_d2ll(x) __double_as_longlong(x)
#define srgp_ll2d(v) __longlong_as_double(v)
#endif

namespace srgp {

SRGP_HD double exp_nonpos(double x)
{
    const double L2E64 = 0x1.71547652b82fep+6;       // 64 / ln 2
    const double LN2_64_HI = 0x1.62e42fee00000p-7;   // ln2/64, 32 significant bits: n * HI is exact
    const double LN2_64_LO = 0x1.a39ef35793c76p-39;
    const double SHIFT = 0x1.8p52;
    if (!(x > -746.0)) return (x == x) ? 0.0 : x;    // underflow to zero; NaN propagates
    const double t = SRGP_FMA(x, L2E64, SHIFT);
    const int n = (int)(uint32_t)(uint64_t)srgp_d2ll(t);   // round-to-nearest integer sits in the low mantissa bits
    const double nd = t - SHIFT;
    double r = SRGP_FMA(-nd, LN2_64_HI, x);
    r = SRGP_FMA(-nd, LN2_64_LO, r);
    const int j = n & 63;
    const int k = n >> 6;                             // floor(n / 64)
    double q = SRGP_FMA(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);   // 1/120, 1/24
    q = SRGP_FMA(r, q, 0x1.5555555555555p-3);                              // 1/6
    q = SRGP_FMA(r, q, 0.5);
    const double s = SRGP_FMA(r * r, q, r);           // exp(r) - 1
    const double tj = SRGP_TAB(j);
    double res = SRGP_FMA(tj, s, tj);                 // 2^(j/64) * exp(r), in [1, 2)
    if (k >= -1000) return srgp_ll2d(srgp_d2ll(res) + ((int64_t)k << 52));
    // gradual underflow: scale in two exact steps
    res = srgp_ll2d(srgp_d2ll(res) + ((int64_t)(k + 1000) << 52));
    return res * 0x1p-1000;
}

}  // namespace srgp
