// scalars.cpp -- the scalar Rcpp exports of the reference (3 transforms + 13 per-pair helpers).
//
// These routines work on one pair of length-d vectors; R code never calls the 13 helpers (0 call sites,
// SURVEY.md section 8b) and calls the transforms on scalars, so they stay on the host.  They exist so that
// the R shim can register all 20 `_sparseRGPs_*` symbols of src/RcppExports.cpp:285-304.
#include <cmath>

#include "../../include/srgp.h"

namespace {

double sq_dist(const double *a, const double *b, int d)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) {
        double t = a[c] - b[c];
        s += t * t;
    }
    return s;
}

double ard_dist(const double *a, const double *b, int d, const double *l)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) {
        double t = (a[c] - b[c]) / l[c];
        s += t * t;
    }
    return s;
}

bool same_point(const double *a, const double *b, int d)
{
    for (int c = 0; c < d; c++)
        if (!(a[c] == b[c])) return false;
    return true;
}

}  // namespace

extern "C" {

void srgp_real_to_pos(const double *x, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = std::exp(x[i]);
}

void srgp_pos_to_real(const double *x, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = std::log(x[i]);
}

void srgp_real_to_bounded(const double *x, const double *ub, const double *lb, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        double e = std::exp(x[i]);
        out[i] = (ub[i] * e + lb[i]) / (e + 1.0);
    }
}

double srgp_cov_fun_sqrd_exp(const double *x1, const double *x2, int d, double sigma, double l)
{
    return sigma * sigma * std::exp(-sq_dist(x1, x2, d) / (2.0 * l * l));
}

double srgp_cov_fun_sqrd_exp_ard(const double *x1, const double *x2, int d, double sigma, const double *l)
{
    return sigma * sigma * std::exp(-ard_dist(x1, x2, d, l) / 2.0);
}

double srgp_cov_fun_exp(const double *x1, const double *x2, int d, double sigma, double l)
{
    double s = 0.0;
    for (int c = 0; c < d; c++) s += std::fabs(x1[c] - x2[c]);  // L1 distance, as the reference
    return sigma * sigma * std::exp(-s / l);
}

double srgp_dsqexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l)
{
    return 2.0 * srgp_cov_fun_sqrd_exp(x1, x2, d, sigma, l);
}

double srgp_dsqexp_dsigma_ard(const double *x1, const double *x2, int d, double sigma, const double *l)
{
    return 2.0 * srgp_cov_fun_sqrd_exp_ard(x1, x2, d, sigma, l);
}

double srgp_dsqexp_dl(const double *x1, const double *x2, int d, double sigma, double l)
{
    double r2 = sq_dist(x1, x2, d);
    return sigma * sigma * std::exp(-r2 / (2.0 * l * l)) * (r2 / (l * l));
}

double srgp_dsqexp_dl_ard(const double *x1, const double *x2, int d, double sigma, const double *l, int comp0)
{
    double t = (x1[comp0] - x2[comp0]) / l[comp0];
    return srgp_cov_fun_sqrd_exp_ard(x1, x2, d, sigma, l) * (t * t);
}

double srgp_dsqexp_dtau(const double *x1, const double *x2, int d, double tau)
{
    return same_point(x1, x2, d) ? 2.0 * tau * tau : 0.0;
}

double srgp_dexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l)
{
    return 2.0 * sigma * sigma * std::exp(-std::sqrt(sq_dist(x1, x2, d)) / l);  // L2 distance (reference quirk)
}

double srgp_dexp_dl(const double *x1, const double *x2, int d, double sigma, double l)
{
    double r = std::sqrt(sq_dist(x1, x2, d));
    return sigma * sigma * std::exp(-r / l) * (r / l);
}

double srgp_dexp_dtau(const double *x1, const double *x2, int d, double tau)
{
    return same_point(x1, x2, d) ? 2.0 * tau * tau : 0.0;
}

void srgp_dsqexp_dx2(const double *x1, const double *x2, int d, double sigma, double l, const double *lb,
                     const double *ub, double *deriv, double *trans_par)
{
    double k = srgp_cov_fun_sqrd_exp(x1, x2, d, sigma, l);
    for (int c = 0; c < d; c++) {
        double tx2 = std::log((x2[c] - lb[c]) / (ub[c] - x2[c]));
        double e = std::exp(tx2);
        double jac = e * (ub[c] - lb[c]) / ((e + 1.0) * (e + 1.0));
        deriv[c] = (x1[c] - x2[c]) / (l * l) * k * jac;
        if (trans_par) trans_par[c] = tx2;
    }
}

void srgp_dsqexp_dx2_ard(const double *x1, const double *x2, int d, double sigma, const double *l,
                         const double *lb, const double *ub, double *deriv, double *trans_par)
{
    double k = srgp_cov_fun_sqrd_exp_ard(x1, x2, d, sigma, l);
    for (int c = 0; c < d; c++) {
        double tx2 = std::log((x2[c] - lb[c]) / (ub[c] - x2[c]));
        double e = std::exp(tx2);
        double jac = e * (ub[c] - lb[c]) / ((e + 1.0) * (e + 1.0));
        deriv[c] = (x1[c] - x2[c]) / (l[c] * l[c]) * k * jac;
        if (trans_par) trans_par[c] = tx2;
    }
}

double srgp_dtrace_term_dtau(double trace_term) { return -2.0 * trace_term; }

}  // extern "C"
