// fastexp.cuh -- exp() for the covariance generators: the FP64 pipe is the contended one, and
// libdevice's exp costs ~20 FP64 instructions per call (degree-11 polynomial + range code).  Here
//     exp(x) = 2^e 2^(j/64) exp(r),   k = rint(64 x / ln 2) = 64 e + j,   r = x - k ln2/64,  |r| <= ln2/128,
// with 2^(j/64) from a 64-entry table the CTA keeps in shared memory (correctly rounded constants), a degree-5 Taylor
// polynomial for exp(r) (truncation r^6/720 <= 3.4e-17) and 2^e applied by an integer add to the exponent field:
// 10 FP64 instructions.  Measured against a long-double exp on 4e6 arguments in [-700, 0]: max relative error
// 2.22e-16 (libm: 2.22e-16) -- within the <= 1 ulp the reference's own exp() (Rcpp sugar -> libm) guarantees, and far
// inside the 1e-10 tolerance on K entries.  exp(0) = 1 exactly (the quirk-Q4 coincidence marker relies on it).
#pragma once
#include <math.h>

namespace srgp {

static __constant__ double EXP2_TAB64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

constexpr int EXP_TAB_DOUBLES = 64;

// every thread of the CTA calls this with its linear id; follow with __syncthreads() before the first exp_tab()
__device__ __forceinline__ void exp_tab_load(double *tab, int tid, int nthreads)
{
    for (int j = tid; j < EXP_TAB_DOUBLES; j += nthreads) tab[j] = EXP2_TAB64[j];
}

__device__ __forceinline__ double exp_tab(double x, const double *__restrict__ tab)
{
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: the low word of x + MAGIC is rint(x)
    const double t = fma(x, 0x1.71547652b82fep+6, MAGIC);    // 64 / ln 2
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -0x1.62e42fee00000p-7, x);            // ln2/64, high 32 bits: kf * hi is exact
    r = fma(kf, -0x1.a39ef35793c76p-39, r);
    double q = fma(0x1.1111111111111p-7, r, 0x1.5555555555555p-5);   // 1/120, 1/24
    q = fma(q, r, 0x1.5555555555555p-3);                     // 1/6
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double T = tab[k & 63];
    double res = fma(T * r, q, T);                           // T (1 + r q)
    res = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
    // outside the range where the exponent arithmetic is valid (results below 2^-1021 or above 2^1023, NaN): libm
    if (!(x >= -708.0 && x <= 709.0)) res = exp(x);
    return res;
}

}  // namespace srgp
