/*
 * oracle/ref_kernels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, no Rcpp) of the per-element covariance kernels of
 * luisdamiano/sparseRGPs:
 *     src/covariance_functionsC.cpp              (make_cov_matC, make_cov_mat_ardC)
 *     src/covariance_function_derivativesC.cpp   (dsig_dthetaC, dsig_dtheta_ardC, helpers)
 * Every function cites the reference file:line whose operation order it follows.
 *
 * PARITY PINNED (this file only): the reference ships no golden vectors and R is
 * not installed here, but its two C++ sources compile unmodified against a small
 * Rcpp stand-in header (oracle/rcpp_shim/Rcpp.h -> oracle/_ref/, oracle/ref_native.py).
 * tests/test_reference_pin.py checks this restatement BIT FOR BIT against that
 * compiled reference on fresh seeded inputs and against the vectors recorded from
 * it (tests/golden/rcpp_layer.*, tests/tools/make_golden.py), plus the analytic known
 * answers of tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never does.
 *
 * Matrices are R-style column-major doubles, no padding: x(i,c) = x[i + n*c].
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

enum { ORA_SQEXP = 0, ORA_EXP = 1, ORA_ARD = 2 };
enum { ORA_SIGMA = 0, ORA_L = 1, ORA_TAU = 2, ORA_LC = 3 };

/* ---- parameter transforms: src/covariance_function_derivativesC.cpp:11-29 ---- */
void ora_real_to_pos(const double *x, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = exp(x[i]);                 /* :12 */
}
void ora_pos_to_real(const double *x, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = log(x[i]);                 /* :20 */
}
void ora_real_to_bounded(const double *x, const double *ub, const double *lb,
                         int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++)                                      /* :28 */
        out[i] = (ub[i] * exp(x[i]) + lb[i]) / (exp(x[i]) + 1);
}

/* ---- strided "row view" helpers (x(i,_) of a column-major matrix) ---- */
typedef struct { const double *p; int64_t stride; } rowv;
static inline double rv(rowv r, int c) { return r.p[(int64_t)c * r.stride]; }

/* sum(pow(x1 - x2, 2)): Rcpp sugar sum = sequential double accumulation */
static double sum_sq_diff(rowv a, rowv b, int d)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) s += pow(rv(a, c) - rv(b, c), 2);
    return s;
}
/* sum(pow((x1 - x2) / l, 2)): divide by l_c, then square, then sum */
static double sum_sq_scaled_diff(rowv a, rowv b, int d, const double *l)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) s += pow((rv(a, c) - rv(b, c)) / l[c], 2);
    return s;
}
static double sum_abs_diff(rowv a, rowv b, int d)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) s += fabs(rv(a, c) - rv(b, c));
    return s;
}
static int all_equal(rowv a, rowv b, int d)
{
    for (int c = 0; c < d; c++) if (!(rv(a, c) == rv(b, c))) return 0;
    return 1;
}

/* ---- per-pair covariance functions ---- */
/* src/covariance_functionsC.cpp:5-12 */
static double cov_sqexp(rowv a, rowv b, int d, double sigma, double l)
{
    return pow(sigma, 2) * exp(-1 / (2 * pow(l, 2)) * sum_sq_diff(a, b, d));
}
/* src/covariance_functionsC.cpp:16-42 */
static double cov_ard(rowv a, rowv b, int d, double sigma, const double *l)
{
    return pow(sigma, 2) * exp(-sum_sq_scaled_diff(a, b, d, l) / 2);
}
/* src/covariance_functionsC.cpp:45-52  (L1 distance) */
static double cov_exp(rowv a, rowv b, int d, double sigma, double l)
{
    return pow(sigma, 2) * exp(-1 / l * sum_abs_diff(a, b, d));
}

/* ---- per-pair log-parameter derivatives ---- */
/* src/covariance_function_derivativesC.cpp:35-52 */
static double dsqexp_dsigma(rowv a, rowv b, int d, double sigma, double l)
{
    double dsigma_dsigmat = sigma;
    return 2 * sigma * exp(-(1 / (2 * pow(l, 2))) * sum_sq_diff(a, b, d)) * dsigma_dsigmat;
}
/* :55-83 */
static double dsqexp_dsigma_ard(rowv a, rowv b, int d, double sigma, const double *l)
{
    double dsigma_dsigmat = sigma;
    return 2 * sigma * exp(-(sum_sq_scaled_diff(a, b, d, l) / 2)) * dsigma_dsigmat;
}
/* :86-104 */
static double dsqexp_dl(rowv a, rowv b, int d, double sigma, double l)
{
    double dl_dlt = l;
    double r2 = sum_sq_diff(a, b, d);
    return (pow(sigma, 2) * exp((-1 / (2 * pow(l, 2))) * r2)) * ((1 / (pow(l, 3))) * r2) * dl_dlt;
}
/* :107-139  (comp is 0-based here; the reference decrements its 1-based double at :121) */
static double dsqexp_dl_ard(rowv a, rowv b, int d, double sigma, const double *l, int comp)
{
    double dl_dlt = l[comp];
    return (pow(sigma, 2) * exp(-(sum_sq_scaled_diff(a, b, d, l) / 2))) *
           ((1 / (pow(l[comp], 3))) * (pow((rv(a, comp) - rv(b, comp)), 2))) * dl_dlt;
}
/* :142-171 and :272-301  (value equality of every coordinate) */
static double dk_dtau(rowv a, rowv b, int d, double tau)
{
    double dtau_dtaut = tau;
    if (all_equal(a, b, d)) return 2 * tau * dtau_dtaut;
    return 0;
}
/* :232-249 (L2 distance, unlike cov_exp) */
static double dexp_dsigma(rowv a, rowv b, int d, double sigma, double l)
{
    double dsigma_dsigmat = sigma;
    return 2 * sigma * exp(-(1 / (l)) * sqrt(sum_sq_diff(a, b, d))) * dsigma_dsigmat;
}
/* :252-269 */
static double dexp_dl(rowv a, rowv b, int d, double sigma, double l)
{
    double dl_dlt = l;
    double r = sqrt(sum_sq_diff(a, b, d));
    return (pow(sigma, 2) * exp((-1 / (l)) * r)) * ((1 / (pow(l, 2))) * r) * dl_dlt;
}

/* ---- exported scalar helpers (contiguous length-d vectors) ---- */
double ora_cov_fun_sqrd_exp(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return cov_sqexp(a, b, d, sigma, l); }
double ora_cov_fun_sqrd_exp_ard(const double *x1, const double *x2, int d, double sigma, const double *l)
{ rowv a = {x1, 1}, b = {x2, 1}; return cov_ard(a, b, d, sigma, l); }
double ora_cov_fun_exp(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return cov_exp(a, b, d, sigma, l); }
double ora_dsqexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return dsqexp_dsigma(a, b, d, sigma, l); }
double ora_dsqexp_dsigma_ard(const double *x1, const double *x2, int d, double sigma, const double *l)
{ rowv a = {x1, 1}, b = {x2, 1}; return dsqexp_dsigma_ard(a, b, d, sigma, l); }
double ora_dsqexp_dl(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return dsqexp_dl(a, b, d, sigma, l); }
double ora_dsqexp_dl_ard(const double *x1, const double *x2, int d, double sigma, const double *l, int comp0)
{ rowv a = {x1, 1}, b = {x2, 1}; return dsqexp_dl_ard(a, b, d, sigma, l, comp0); }
double ora_dk_dtau(const double *x1, const double *x2, int d, double tau)
{ rowv a = {x1, 1}, b = {x2, 1}; return dk_dtau(a, b, d, tau); }
double ora_dexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return dexp_dsigma(a, b, d, sigma, l); }
double ora_dexp_dl(const double *x1, const double *x2, int d, double sigma, double l)
{ rowv a = {x1, 1}, b = {x2, 1}; return dexp_dl(a, b, d, sigma, l); }

/* knot-coordinate derivatives (Rcpp versions, unused from R):
 * src/covariance_function_derivativesC.cpp:176-227.  out has length d. */
void ora_dsqexp_dx2(const double *x1, const double *x2, int d, double sigma, double l,
                    const double *lb, const double *ub, double *deriv, double *tx2_out)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) s += pow(x1[c] - x2[c], 2);
    for (int c = 0; c < d; c++) {
        double tx2 = log((x2[c] - lb[c]) / (ub[c] - x2[c]));                       /* :180 */
        double dx2_dtx2 = (exp(tx2) * (ub[c] - lb[c])) / pow((exp(tx2) + 1), 2);   /* :181 */
        deriv[c] = (1 / (pow(l, 2))) * (x1[c] - x2[c]) * pow(sigma, 2) *
                   (exp(-s / (2 * pow(l, 2)))) * dx2_dtx2;                         /* :188-189 */
        if (tx2_out) tx2_out[c] = tx2;
    }
}
void ora_dsqexp_dx2_ard(const double *x1, const double *x2, int d, double sigma, const double *l,
                        const double *lb, const double *ub, double *deriv, double *tx2_out)
{
    rowv a = {x1, 1}, b = {x2, 1};
    double s = sum_sq_scaled_diff(a, b, d, l);
    for (int c = 0; c < d; c++) {
        double tx2 = log((x2[c] - lb[c]) / (ub[c] - x2[c]));                       /* :214 */
        double dx2_dtx2 = (exp(tx2) * (ub[c] - lb[c])) / pow((exp(tx2) + 1), 2);   /* :215 */
        deriv[c] = (1 / (pow(l[c], 2))) * (x1[c] - x2[c]) * pow(sigma, 2) *
                   (exp(-(s / 2))) * dx2_dtx2;                                     /* :222-223 */
        if (tx2_out) tx2_out[c] = tx2;
    }
}

/*
 * Dense covariance assembly.
 *   kernel ORA_SQEXP / ORA_EXP : src/covariance_functionsC.cpp:72-169 (l[0] is the length scale)
 *   kernel ORA_ARD             : src/covariance_functionsC.cpp:191-252 (l[0..d-1])
 * x_pred == NULL is the reference's "x_pred(0,0) is NA" self-covariance branch:
 * n1 x n1 output, tau^2 + delta added where i == j (index equality, :90, :210).
 * Returns 0, or 1 for an unknown kernel (reference: message on Rcerr + 0x0 matrix).
 */
int ora_make_cov_mat(int kernel, const double *x, int64_t n1, const double *x_pred, int64_t n2,
                     int d, double sigma, const double *l, double tau, double delta, double *out)
{
    if (kernel != ORA_SQEXP && kernel != ORA_EXP && kernel != ORA_ARD) return 1;
    const int self = (x_pred == NULL);
    const double *xb = self ? x : x_pred;
    const int64_t nb = self ? n1 : n2;
    for (int64_t i = 0; i < n1; i++) {
        rowv a = {x + i, n1};
        for (int64_t j = 0; j < nb; j++) {
            rowv b = {xb + j, nb};
            double k;
            if (kernel == ORA_SQEXP)    k = cov_sqexp(a, b, d, sigma, l[0]);
            else if (kernel == ORA_EXP) k = cov_exp(a, b, d, sigma, l[0]);
            else                        k = cov_ard(a, b, d, sigma, l);
            if (self && i == j) k = k + pow(tau, 2) + delta;
            out[i + n1 * j] = k;
        }
    }
    return 0;
}

/*
 * Dense d(Sigma)/d(log theta).
 *   sqexp / exp : src/covariance_function_derivativesC.cpp:307-552
 *   ard         : src/covariance_function_derivativesC.cpp:555-722
 * par: ORA_SIGMA, ORA_L (sqexp/exp), ORA_TAU, ORA_LC (ard; comp0 = 0-based component).
 * Quirk: kernel == ORA_EXP, cross matrix, par == ORA_TAU returns zeros
 * because the reference returns `mat` at :520 before reaching the tau branch.
 * Returns 0; 1 = unknown kernel; 2 = unknown parameter for the kernel
 * (reference: message on Rcerr + 0x0 matrix in both cases).
 */
int ora_dsig_dtheta(int kernel, int par, int comp0, const double *x, int64_t n1,
                    const double *x_pred, int64_t n2, int d, double sigma, const double *l,
                    double tau, double *out)
{
    if (kernel != ORA_SQEXP && kernel != ORA_EXP && kernel != ORA_ARD) return 1;
    if (kernel == ORA_EXP && x_pred != NULL && par != ORA_SIGMA && par != ORA_L) {
        /* Q9: anything but sigma / l (tau included, and unknown names) hits `return mat` at :520 */
        memset(out, 0, sizeof(double) * (size_t)n1 * (size_t)n2);
        return 0;
    }
    if (kernel == ORA_ARD) {
        if (!(par == ORA_SIGMA || par == ORA_TAU || (par == ORA_LC && comp0 >= 0 && comp0 < d))) return 2;
    } else {
        if (!(par == ORA_SIGMA || par == ORA_TAU || par == ORA_L)) return 2;
    }
    const int self = (x_pred == NULL);
    const double *xb = self ? x : x_pred;
    const int64_t nb = self ? n1 : n2;
    for (int64_t i = 0; i < n1; i++) {
        rowv a = {x + i, n1};
        for (int64_t j = 0; j < nb; j++) {
            rowv b = {xb + j, nb};
            double v;
            if (par == ORA_TAU) v = dk_dtau(a, b, d, tau);
            else if (kernel == ORA_SQEXP)
                v = (par == ORA_SIGMA) ? dsqexp_dsigma(a, b, d, sigma, l[0]) : dsqexp_dl(a, b, d, sigma, l[0]);
            else if (kernel == ORA_EXP)
                v = (par == ORA_SIGMA) ? dexp_dsigma(a, b, d, sigma, l[0]) : dexp_dl(a, b, d, sigma, l[0]);
            else
                v = (par == ORA_SIGMA) ? dsqexp_dsigma_ard(a, b, d, sigma, l)
                                       : dsqexp_dl_ard(a, b, d, sigma, l, comp0);
            out[i + n1 * j] = v;
        }
    }
    return 0;
}
