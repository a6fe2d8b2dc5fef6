/*
 * r/shim.c -- the R-side glue a sparseRGPs maintainer adds to bind libsrgp.so behind the package's own
 * `.Call` routines (replaces src/RcppExports.cpp + the two Rcpp kernel files of the reference).
 *
 * R (Rinternals.h, libR) is not installed in this repository's image; this file is compiled
 * and run unchanged against the miniature R C API of tests/mini_r/ (tests/test_r_shim.py: registration table,
 * coercion, golden vectors of the reference through .Call, R errors, PROTECT balance).  It is marshalling only;
 * every numerical statement lives behind include/srgp.h.  Build inside the package:   PKG_LIBS = -L<dir> -lsrgp   (src/Makevars), keep R/RcppExports.R as is.
 *
 * Registered routine names and arities are those of src/RcppExports.cpp:285-304, so R/RcppExports.R loads
 * unchanged.  Error conventions: status != 0 -> Rf_error, EXCEPT unknown kernel /
 * parameter names, which print the reference's message and return a 0 x 0 matrix.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "srgp.h"

static srgp_ctx *g_ctx = NULL;

static srgp_ctx *ctx(void)
{
    if (!g_ctx && srgp_ctx_create(0, &g_ctx) != SRGP_OK) Rf_error("sparseRGPs: %s", srgp_last_error());
    return g_ctx;
}

static double list_get(SEXP lst, const char *name)
{
    SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < Rf_xlength(lst); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(lst, i));
    Rf_error("Index out of bounds: [index='%s'].", name);   /* Rcpp's index_out_of_bounds */
    return NA_REAL;
}

/* overwrite the named element of a list of scalars (the list must be freshly duplicated by the caller) */
static void set_list(SEXP lst, const char *name, double v)
{
    SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < Rf_xlength(lst); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) { SET_VECTOR_ELT(lst, i, Rf_ScalarReal(v)); return; }
}

static int kernel_id(SEXP cov_fun)
{
    const char *s = CHAR(STRING_ELT(cov_fun, 0));
    if (!strcmp(s, "sqexp")) return SRGP_SQEXP;
    if (!strcmp(s, "exp")) return SRGP_EXP;
    if (!strcmp(s, "ard")) return SRGP_ARD;
    return -1;
}

static SEXP empty_matrix(const char *msg)
{
    REprintf("%s", msg);
    return Rf_allocMatrix(REALSXP, 0, 0);
}

/* x_pred = matrix() arrives as a 1 x 1 logical NA (src/covariance_functionsC.cpp:81); a zero-length x_pred has no
   element 0 to look at and is a cross-covariance with zero columns */
static int is_empty(SEXP x_pred) { return Rf_xlength(x_pred) > 0 && (ISNA(REAL(x_pred)[0]) || ISNAN(REAL(x_pred)[0])); }

/* length-scale array of the kernel from the named list; d is checked BEFORE the fixed-size array is filled */
static void get_l(SEXP cov_par, int kernel, int d, SEXP lnames, double *l)
{
    if (kernel == SRGP_ARD) {
        if (d > SRGP_MAX_D) Rf_error("sparseRGPs: d = %d exceeds %d", d, SRGP_MAX_D);
        if (Rf_length(lnames) < d) Rf_error("sparseRGPs: lnames has %d entries for d = %d", Rf_length(lnames), d);
        for (int c = 0; c < d; c++) l[c] = list_get(cov_par, CHAR(STRING_ELT(lnames, c)));
    } else {
        l[0] = list_get(cov_par, "l");
    }
}

/* R recycles a mean of length 1 (`y - mu`); length n passes through, length 0 is the zero mean, anything else is the
   R error a length mismatch deserves.  v must already be a REALSXP. */
static const double *recycled(SEXP v, int n, const char *what)
{
    const int len = Rf_length(v);
    if (len == n) return REAL(v);
    if (len == 0) return NULL;
    if (len == 1) {
        double *out = (double *)R_alloc(n > 0 ? n : 1, sizeof(double));
        for (int i = 0; i < n; i++) out[i] = REAL(v)[0];
        return out;
    }
    Rf_error("sparseRGPs: %s has length %d, expected 1 or %d", what, len, n);
    return NULL;
}

/* gradient in the library's canonical order with its names: c(sigma, l | l1..ld, tau).  The R side reorders it to
   names(cov_par) (the reference fills its gradient by name, R/vi_functions.R:163,259-261,416). */
static SEXP named_gradient(const double *g, int kernel, int d, SEXP lnames)
{
    const int p = (kernel == SRGP_ARD) ? d + 2 : 3;
    SEXP out = PROTECT(Rf_allocVector(REALSXP, p)), nm = PROTECT(Rf_allocVector(STRSXP, p));
    for (int k = 0; k < p; k++) REAL(out)[k] = g[k];
    SET_STRING_ELT(nm, 0, Rf_mkChar("sigma"));
    if (kernel == SRGP_ARD) for (int c = 0; c < d; c++) SET_STRING_ELT(nm, 1 + c, STRING_ELT(lnames, c));
    else SET_STRING_ELT(nm, 1, Rf_mkChar("l"));
    SET_STRING_ELT(nm, p - 1, Rf_mkChar("tau"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}

static SEXP named_list(int n, const char **names, SEXP *vals)
{
    SEXP out = PROTECT(Rf_allocVector(VECSXP, n)), nm = PROTECT(Rf_allocVector(STRSXP, n));
    for (int i = 0; i < n; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(nm, i, Rf_mkChar(names[i])); }
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}

static int family_id(SEXP family) { return strcmp(CHAR(STRING_ELT(family, 0)), "poisson") ? SRGP_BERNOULLI : SRGP_POISSON; }

static SEXP assemble(SEXP x, SEXP x_pred, SEXP cov_par, int kernel, double delta, SEXP lnames, int par, int comp0,
                     int derivative)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    x_pred = PROTECT(Rf_coerceVector(x_pred, REALSXP));
    const int n1 = Rf_nrows(x), d = Rf_ncols(x);
    const int self = is_empty(x_pred);
    const int n2 = self ? n1 : Rf_nrows(x_pred);
    double l[SRGP_MAX_D];
    get_l(cov_par, kernel, d, lnames, l);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n1, n2));
    int st;
    if (!derivative)
        st = srgp_make_cov_mat(ctx(), kernel, REAL(x), n1, self ? NULL : REAL(x_pred), n2, d,
                               list_get(cov_par, "sigma"), l, self ? list_get(cov_par, "tau") : 0.0, delta, REAL(out));
    else
        st = srgp_dsig_dtheta(ctx(), kernel, par, comp0, REAL(x), n1, self ? NULL : REAL(x_pred), n2, d,
                              list_get(cov_par, "sigma"), l, par == SRGP_PAR_TAU ? list_get(cov_par, "tau") : 0.0,
                              REAL(out));
    UNPROTECT(3);
    if (st == SRGP_ERR_UNKNOWN_PAR) return empty_matrix("Error: invalid parameter name for chosen covariance function");
    if (st != SRGP_OK) Rf_error("sparseRGPs: %s", srgp_last_error());
    return out;
}

SEXP _sparseRGPs_make_cov_matC(SEXP x, SEXP x_pred, SEXP cov_par, SEXP cov_fun, SEXP delta)
{
    const int k = kernel_id(cov_fun);
    if (k != SRGP_SQEXP && k != SRGP_EXP) return empty_matrix("Error: invalid covariance function");
    return assemble(x, x_pred, cov_par, k, Rf_asReal(delta), R_NilValue, 0, 0, 0);
}

SEXP _sparseRGPs_make_cov_mat_ardC(SEXP x, SEXP x_pred, SEXP cov_par, SEXP cov_fun, SEXP delta, SEXP lnames)
{
    if (kernel_id(cov_fun) != SRGP_ARD) return empty_matrix("Error: invalid covariance function");
    return assemble(x, x_pred, cov_par, SRGP_ARD, Rf_asReal(delta), lnames, 0, 0, 0);
}

static int par_id(const char *name, int kernel, SEXP lnames, int *comp0)
{
    *comp0 = -1;
    if (!strcmp(name, "sigma")) return SRGP_PAR_SIGMA;
    if (!strcmp(name, "tau")) return SRGP_PAR_TAU;
    if (kernel == SRGP_ARD) {
        int found = -1;   /* the reference keeps the LAST match (covariance_function_derivativesC.cpp:596-605) */
        for (int c = 0; c < Rf_length(lnames); c++)
            if (!strcmp(name, CHAR(STRING_ELT(lnames, c)))) found = c;
        if (found >= 0) { *comp0 = found; return SRGP_PAR_LC; }
        return 99;
    }
    return strcmp(name, "l") ? 99 : SRGP_PAR_L;
}

SEXP _sparseRGPs_dsig_dthetaC(SEXP x, SEXP x_pred, SEXP cov_par, SEXP cov_fun, SEXP par_name)
{
    const int k = kernel_id(cov_fun);
    if (k != SRGP_SQEXP && k != SRGP_EXP) return empty_matrix("Error: invalid covariance function");
    int comp0;
    const int par = par_id(CHAR(STRING_ELT(par_name, 0)), k, R_NilValue, &comp0);
    return assemble(x, x_pred, cov_par, k, 0.0, R_NilValue, par, comp0, 1);
}

SEXP _sparseRGPs_dsig_dtheta_ardC(SEXP x, SEXP x_pred, SEXP cov_par, SEXP cov_fun, SEXP par_name, SEXP lnames)
{
    if (kernel_id(cov_fun) != SRGP_ARD) return empty_matrix("Error: invalid covariance function");
    int comp0;
    const int par = par_id(CHAR(STRING_ELT(par_name, 0)), SRGP_ARD, lnames, &comp0);
    return assemble(x, x_pred, cov_par, SRGP_ARD, 0.0, lnames, par, comp0, 1);
}

/* transforms (vectors in, vectors out) */
SEXP _sparseRGPs_real_to_pos(SEXP x)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_xlength(x)));
    srgp_real_to_pos(REAL(x), Rf_xlength(x), REAL(out));
    UNPROTECT(2);
    return out;
}
SEXP _sparseRGPs_pos_to_real(SEXP x)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_xlength(x)));
    srgp_pos_to_real(REAL(x), Rf_xlength(x), REAL(out));
    UNPROTECT(2);
    return out;
}
SEXP _sparseRGPs_real_to_bounded(SEXP x, SEXP ub, SEXP lb)
{
    x = PROTECT(Rf_coerceVector(x, REALSXP));
    ub = PROTECT(Rf_coerceVector(ub, REALSXP));
    lb = PROTECT(Rf_coerceVector(lb, REALSXP));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_xlength(x)));
    srgp_real_to_bounded(REAL(x), REAL(ub), REAL(lb), Rf_xlength(x), REAL(out));
    UNPROTECT(4);
    return out;
}

/* Rcpp's input_parameter<NumericVector> coerces integer / logical vectors to double: the registered entry points
   (generated at the end of this section) coerce and PROTECT the vector arguments, then call impl_<name>. */

/* scalar covariance functions: (x1, x2, cov_par[, lnames]) -> double */
static SEXP impl_cov_fun_sqrd_expC(SEXP x1, SEXP x2, SEXP cov_par)
{
    return Rf_ScalarReal(srgp_cov_fun_sqrd_exp(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"),
                                               list_get(cov_par, "l")));
}
static SEXP impl_cov_fun_expC(SEXP x1, SEXP x2, SEXP cov_par)
{
    return Rf_ScalarReal(srgp_cov_fun_exp(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"),
                                          list_get(cov_par, "l")));
}
static SEXP impl_cov_fun_sqrd_exp_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames)
{
    double l[SRGP_MAX_D];
    get_l(cov_par, SRGP_ARD, Rf_length(x1), lnames, l);
    return Rf_ScalarReal(srgp_cov_fun_sqrd_exp_ard(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), l));
}

/* per-pair derivative helpers return list(derivative, trans_par, inv_trans_par) like the Rcpp originals */
static SEXP deriv_list(double deriv, double par)
{
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(deriv));
    SET_VECTOR_ELT(out, 1, Rf_ScalarReal(log(par)));
    /* inv_trans_par = real_to_pos(par) on the UNtransformed value, i.e. exp(par)
       (src/covariance_function_derivativesC.cpp:49,80,101,136,168; golden vectors in tests/golden) */
    double inv;
    srgp_real_to_pos(&par, 1, &inv);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal(inv));
    SET_STRING_ELT(nm, 0, Rf_mkChar("derivative"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("trans_par"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("inv_trans_par"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}
#define PAIR_SCALAR(NAME, FN, PARNAME)                                                                         \
    static SEXP impl_##NAME(SEXP x1, SEXP x2, SEXP cov_par)                                                      \
    {                                                                                                          \
        return deriv_list(FN(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), list_get(cov_par, "l")), \
                          list_get(cov_par, PARNAME));                                                         \
    }
PAIR_SCALAR(dsqexp_dsigmaC, srgp_dsqexp_dsigma, "sigma")
PAIR_SCALAR(dsqexp_dlC, srgp_dsqexp_dl, "l")
PAIR_SCALAR(dexp_dsigmaC, srgp_dexp_dsigma, "sigma")
PAIR_SCALAR(dexp_dlC, srgp_dexp_dl, "l")
static SEXP impl_dsqexp_dtauC(SEXP x1, SEXP x2, SEXP cov_par)
{
    return deriv_list(srgp_dsqexp_dtau(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "tau")), list_get(cov_par, "tau"));
}
static SEXP impl_dexp_dtauC(SEXP x1, SEXP x2, SEXP cov_par)
{
    return deriv_list(srgp_dexp_dtau(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "tau")), list_get(cov_par, "tau"));
}
static SEXP impl_dsqexp_dsigma_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames)
{
    double l[SRGP_MAX_D];
    get_l(cov_par, SRGP_ARD, Rf_length(x1), lnames, l);
    return deriv_list(srgp_dsqexp_dsigma_ard(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), l),
                      list_get(cov_par, "sigma"));
}
static SEXP impl_dsqexp_dl_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames, SEXP comp)
{
    double l[SRGP_MAX_D];
    get_l(cov_par, SRGP_ARD, Rf_length(x1), lnames, l);
    const int c0 = (int)Rf_asReal(comp) - 1;   /* 1-based in R (covariance_function_derivativesC.cpp:121) */
    return deriv_list(srgp_dsqexp_dl_ard(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), l, c0), l[c0]);
}
static SEXP dx2_list(SEXP x2, SEXP lb, SEXP ub, const double *deriv, const double *tp, int d)
{
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SEXP a = PROTECT(Rf_allocVector(REALSXP, d)), b = PROTECT(Rf_allocVector(REALSXP, d)), c = PROTECT(Rf_allocVector(REALSXP, d));
    memcpy(REAL(a), deriv, sizeof(double) * d);
    memcpy(REAL(b), tp, sizeof(double) * d);
    srgp_real_to_bounded(REAL(x2), REAL(ub), REAL(lb), d, REAL(c));
    SET_VECTOR_ELT(out, 0, a); SET_VECTOR_ELT(out, 1, b); SET_VECTOR_ELT(out, 2, c);
    SET_STRING_ELT(nm, 0, Rf_mkChar("derivative")); SET_STRING_ELT(nm, 1, Rf_mkChar("trans_par"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("inv_trans_par"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}
static SEXP impl_dsqexp_dx2C(SEXP x1, SEXP x2, SEXP cov_par, SEXP lb, SEXP ub)
{
    double deriv[SRGP_MAX_D], tp[SRGP_MAX_D];
    if (Rf_length(x1) > SRGP_MAX_D) Rf_error("sparseRGPs: d = %d exceeds %d", Rf_length(x1), SRGP_MAX_D);
    srgp_dsqexp_dx2(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), list_get(cov_par, "l"), REAL(lb),
                    REAL(ub), deriv, tp);
    return dx2_list(x2, lb, ub, deriv, tp, Rf_length(x1));
}
static SEXP impl_dsqexp_dx2_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lb, SEXP ub, SEXP lnames)
{
    double l[SRGP_MAX_D], deriv[SRGP_MAX_D], tp[SRGP_MAX_D];
    get_l(cov_par, SRGP_ARD, Rf_length(x1), lnames, l);
    srgp_dsqexp_dx2_ard(REAL(x1), REAL(x2), Rf_length(x1), list_get(cov_par, "sigma"), l, REAL(lb), REAL(ub), deriv, tp);
    return dx2_list(x2, lb, ub, deriv, tp, Rf_length(x1));
}

/* registered entry points of the per-pair helpers: coerce x1, x2 (and lb, ub) like Rcpp, keep them protected */
SEXP _sparseRGPs_cov_fun_sqrd_expC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_cov_fun_sqrd_expC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_cov_fun_expC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_cov_fun_expC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_cov_fun_sqrd_exp_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_cov_fun_sqrd_exp_ardC(x1, x2, cov_par, lnames);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dsqexp_dtauC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dsqexp_dtauC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dexp_dtauC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dexp_dtauC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dsqexp_dsigma_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dsqexp_dsigma_ardC(x1, x2, cov_par, lnames);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dsqexp_dl_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lnames, SEXP comp)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dsqexp_dl_ardC(x1, x2, cov_par, lnames, comp);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dsqexp_dx2C(SEXP x1, SEXP x2, SEXP cov_par, SEXP lb, SEXP ub)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    lb = PROTECT(Rf_coerceVector(lb, REALSXP));
    ub = PROTECT(Rf_coerceVector(ub, REALSXP));
    SEXP r = impl_dsqexp_dx2C(x1, x2, cov_par, lb, ub);
    UNPROTECT(4);
    return r;
}
SEXP _sparseRGPs_dsqexp_dx2_ardC(SEXP x1, SEXP x2, SEXP cov_par, SEXP lb, SEXP ub, SEXP lnames)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    lb = PROTECT(Rf_coerceVector(lb, REALSXP));
    ub = PROTECT(Rf_coerceVector(ub, REALSXP));
    SEXP r = impl_dsqexp_dx2_ardC(x1, x2, cov_par, lb, ub, lnames);
    UNPROTECT(4);
    return r;
}
SEXP _sparseRGPs_dsqexp_dsigmaC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dsqexp_dsigmaC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dsqexp_dlC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dsqexp_dlC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dexp_dsigmaC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dexp_dsigmaC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}
SEXP _sparseRGPs_dexp_dlC(SEXP x1, SEXP x2, SEXP cov_par)
{
    x1 = PROTECT(Rf_coerceVector(x1, REALSXP));
    x2 = PROTECT(Rf_coerceVector(x2, REALSXP));
    SEXP r = impl_dexp_dlC(x1, x2, cov_par);
    UNPROTECT(2);
    return r;
}

/* ---- new fused entry points (r/patches.R delegates to them; the R signatures stay as they are) ---------- */
/* .Call("_sparseRGPs_gauss_obj_grad", model, cov_fun, xy, y, mu, xu, cov_par, delta, lnames)
     -> list(objective =, gradient = named numeric)  : one iteration's elbo_fun + delbo_dcov_par (model 0) or
        obj_fun_norm + dlogp_dcov_par (model 1).  A failed Cholesky raises an R error (SRGP_ERR_NOT_PD), which
        keeps the try() paths of R/knot_proposal_functions.R:1314-1354 working. */
SEXP _sparseRGPs_gauss_obj_grad(SEXP model, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP cov_par,
                                SEXP delta, SEXP lnames)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    double gbuf[SRGP_MAX_D + 2];
    double obj = NA_REAL;
    const int st = srgp_gauss_obj_grad_host(ctx(), Rf_asInteger(model), k, REAL(xy), n, d, REAL(y),
                                            recycled(mu, n, "mu"), REAL(xu), m,
                                            list_get(cov_par, "sigma"), l, list_get(cov_par, "tau"), Rf_asReal(delta),
                                            &obj, gbuf);
    if (st != SRGP_OK) { UNPROTECT(4); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP grad = PROTECT(named_gradient(gbuf, k, d, lnames));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(obj));
    SET_VECTOR_ELT(out, 1, grad);
    SET_STRING_ELT(nm, 0, Rf_mkChar("objective"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("gradient"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(7);
    return out;
}

/* .Call("_sparseRGPs_gauss_obj_grad_knots", model, cov_fun, xy, y, mu, xu, cov_par, delta, lnames, knot_bounds,
         knot_opt, transform) -> list(objective, gradient, knot_gradient, trans_knot): the branch of delbo_dcov_par /
   dlogp_dcov_par taken when dcov_fun_dknot is a function (R/vi_functions.R:425-592,
   R/laplace_approx_gradient.R:965-1126).  knot_bounds is the d x 2 matrix of R/vi_functions.R:175-178; knot_opt is
   R's 1-based index vector. */
SEXP _sparseRGPs_gauss_obj_grad_knots(SEXP model, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP cov_par,
                                      SEXP delta, SEXP lnames, SEXP knot_bounds, SEXP knot_opt, SEXP transform)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    knot_bounds = PROTECT(Rf_coerceVector(knot_bounds, REALSXP));
    knot_opt = PROTECT(Rf_coerceVector(knot_opt, INTSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu);
    const int tr = Rf_asLogical(transform), n_opt = Rf_length(knot_opt);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    int *opt0 = (int *)R_alloc(n_opt > 0 ? n_opt : 1, sizeof(int));
    for (int t = 0; t < n_opt; t++) opt0[t] = INTEGER(knot_opt)[t] - 1;
    double gbuf[SRGP_MAX_D + 2];
    SEXP kgrad = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)m * d));
    SEXP tknot = PROTECT(Rf_allocMatrix(REALSXP, m, d));
    double obj = NA_REAL;
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_gauss_obj_grad_knots(ctx(), Rf_asInteger(model), k, REAL(xu), m, list_get(cov_par, "sigma"), l,
                                       list_get(cov_par, "tau"), Rf_asReal(delta),
                                       tr ? REAL(knot_bounds) : NULL, tr ? REAL(knot_bounds) + d : NULL,  /* cbind(lb, ub) */
                                       opt0, n_opt, &obj, gbuf, REAL(kgrad), REAL(tknot));
    if (st != SRGP_OK) { UNPROTECT(8); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP grad = PROTECT(named_gradient(gbuf, k, d, lnames));
    const char *names[] = {"objective", "gradient", "knot_gradient", "trans_knot"};
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4)), nm = PROTECT(Rf_allocVector(STRSXP, 4));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(obj));
    SET_VECTOR_ELT(out, 1, grad);
    SET_VECTOR_ELT(out, 2, kgrad);
    SET_VECTOR_ELT(out, 3, tknot);
    for (int t = 0; t < 4; t++) SET_STRING_ELT(nm, t, Rf_mkChar(names[t]));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(11);
    return out;
}

/* .Call("_sparseRGPs_oat_scores", model, cov_fun, xy, y, mu, xu, pseudo_prop, cov_par, delta, lnames)
     -> list(objective = <objective with xu>, scores = numeric(nrow(pseudo_prop))): the candidate loop of
   knot_prop_random_norm_vi (R/vi_functions.R:2211-2298, model 0) / knot_prop_random_norm
   (R/knot_proposal_functions.R:1283-1353, model 1).  NaN scores mark candidates whose chol()/solve() would have
   raised; the R side resamples those exactly as before. */
SEXP _sparseRGPs_oat_scores(SEXP model, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP pseudo_prop,
                            SEXP cov_par, SEXP delta, SEXP lnames)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    pseudo_prop = PROTECT(Rf_coerceVector(pseudo_prop, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu), T = Rf_nrows(pseudo_prop);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    SEXP scores = PROTECT(Rf_allocVector(REALSXP, T));
    double obj0 = NA_REAL;
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_oat_scores(ctx(), Rf_asInteger(model), k, REAL(xu), m, REAL(pseudo_prop), T,
                             list_get(cov_par, "sigma"), l, list_get(cov_par, "tau"), Rf_asReal(delta), &obj0,
                             REAL(scores));
    if (st != SRGP_OK) { UNPROTECT(6); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, Rf_ScalarReal(obj0));
    SET_VECTOR_ELT(out, 1, scores);
    SET_STRING_ELT(nm, 0, Rf_mkChar("objective"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("scores"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(8);
    return out;
}

/* .Call("_sparseRGPs_gauss_fit", model, cov_fun, xy, y, mu, xu, cov_par_start, delta, lnames, optim_method,
         optim_par = c(decay, epsilon, eta, learn_rate), maxit, obj_tol, grad_tol, opt_theta, opt_knots, knot_bounds,
         knot_opt) -> list(cov_par, xu, iter, obj_fun, grad, cov_par_history): the loops of norm_grad_ascent_vi
   (R/vi_functions.R:963-1158) / norm_grad_ascent (R/laplace_gradient_ascent.R:1453-1633). */
SEXP _sparseRGPs_gauss_fit(SEXP model, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP cov_par, SEXP delta,
                           SEXP lnames, SEXP optim_method, SEXP optim_par, SEXP maxit, SEXP obj_tol, SEXP grad_tol,
                           SEXP opt_theta, SEXP opt_knots, SEXP knot_bounds, SEXP knot_opt)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    knot_bounds = PROTECT(Rf_coerceVector(knot_bounds, REALSXP));
    knot_opt = PROTECT(Rf_coerceVector(knot_opt, INTSXP));
    SEXP xu_out = PROTECT(Rf_duplicate(Rf_coerceVector(xu, REALSXP)));          /* in/out */
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu_out);
    const int nl = (k == SRGP_ARD) ? d : 1, p = nl + 2, mi = Rf_asInteger(maxit), n_opt = Rf_length(knot_opt);
    double sigma = list_get(cov_par, "sigma"), tau = list_get(cov_par, "tau"), l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    srgp_fit_opt o;
    o.optim_method = strcmp(CHAR(STRING_ELT(optim_method, 0)), "ga") ? SRGP_OPT_ADADELTA : SRGP_OPT_GA;
    o.decay = REAL(optim_par)[0]; o.epsilon = REAL(optim_par)[1]; o.eta = REAL(optim_par)[2];
    o.learn_rate = REAL(optim_par)[3];
    o.maxit = mi; o.obj_tol = Rf_asReal(obj_tol); o.grad_tol = Rf_asReal(grad_tol);
    o.opt_theta = Rf_asLogical(opt_theta); o.opt_knots = Rf_asLogical(opt_knots);
    int *opt0 = (int *)R_alloc(n_opt > 0 ? n_opt : 1, sizeof(int));
    for (int t = 0; t < n_opt; t++) opt0[t] = INTEGER(knot_opt)[t] - 1;
    double *obj_h = (double *)R_alloc(mi, sizeof(double));
    double *par_h = (double *)R_alloc((size_t)mi * p, sizeof(double)), *grad_h = (double *)R_alloc((size_t)mi * p, sizeof(double));
    int iter = 0;
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_gauss_fit(ctx(), Rf_asInteger(model), k, REAL(xu_out), m, &sigma, l, &tau, Rf_asReal(delta), &o,
                            o.opt_knots ? REAL(knot_bounds) : NULL, o.opt_knots ? REAL(knot_bounds) + d : NULL,
                            opt0, n_opt, &iter, obj_h, par_h, grad_h);
    if (st != SRGP_OK) { UNPROTECT(6); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    /* cov_par: same names and order as cov_par_start */
    SEXP cp = PROTECT(Rf_duplicate(cov_par));
    set_list(cp, "sigma", sigma); set_list(cp, "tau", tau);
    if (k == SRGP_ARD) for (int c = 0; c < d; c++) set_list(cp, CHAR(STRING_ELT(lnames, c)), l[c]);
    else set_list(cp, "l", l[0]);
    SEXP objv = PROTECT(Rf_allocVector(REALSXP, iter));
    SEXP gh = PROTECT(Rf_allocMatrix(REALSXP, iter, p)), ph = PROTECT(Rf_allocMatrix(REALSXP, iter, p));
    for (int i = 0; i < iter; i++) {
        REAL(objv)[i] = obj_h[i];
        for (int j = 0; j < p; j++) {               /* row-major history -> R's column-major matrix */
            REAL(gh)[i + (size_t)iter * j] = grad_h[(size_t)i * p + j];
            REAL(ph)[i + (size_t)iter * j] = par_h[(size_t)i * p + j];
        }
    }
    const char *names[] = {"cov_par", "xu", "iter", "obj_fun", "grad", "cov_par_history"};
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 6)), nm = PROTECT(Rf_allocVector(STRSXP, 6));
    SET_VECTOR_ELT(out, 0, cp);
    SET_VECTOR_ELT(out, 1, xu_out);
    SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(iter));
    SET_VECTOR_ELT(out, 3, objv);
    SET_VECTOR_ELT(out, 4, gh);
    SET_VECTOR_ELT(out, 5, ph);
    for (int t = 0; t < 6; t++) SET_STRING_ELT(nm, t, Rf_mkChar(names[t]));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(12);
    return out;
}

/* .Call("_sparseRGPs_trace_term", sigma, tau, delta, Sigma12, Sigma22): body of trace_term_fun */
SEXP _sparseRGPs_trace_term(SEXP sigma, SEXP tau, SEXP delta, SEXP Sigma12, SEXP Sigma22)
{
    double out = NA_REAL;
    Sigma12 = PROTECT(Rf_coerceVector(Sigma12, REALSXP));
    Sigma22 = PROTECT(Rf_coerceVector(Sigma22, REALSXP));
    const int st = srgp_trace_term(ctx(), Rf_asReal(sigma), Rf_asReal(tau), Rf_asReal(delta), REAL(Sigma12),
                                   Rf_nrows(Sigma12), Rf_ncols(Sigma12), REAL(Sigma22), &out);
    UNPROTECT(2);
    if (st != SRGP_OK) Rf_error("sparseRGPs: %s", srgp_last_error());
    return Rf_ScalarReal(out);
}

/* .Call("_sparseRGPs_laplace_newton", family, cov_fun, xy, y, mu, xu, muu, cov_par, delta, lnames, start_vals,
         maxit, tol, pois_m) -> the list newtrap_sparseGP returns (R/newtrap_sparseGP.R:183-184) */
SEXP _sparseRGPs_laplace_newton(SEXP family, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP muu, SEXP cov_par,
                                SEXP delta, SEXP lnames, SEXP start_vals, SEXP maxit, SEXP tol, SEXP pois_m)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    muu = PROTECT(Rf_coerceVector(muu, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu), mi = Rf_asInteger(maxit);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    if (srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu")) != SRGP_OK)
        Rf_error("sparseRGPs: %s", srgp_last_error());
    SEXP ff = PROTECT(Rf_duplicate(Rf_coerceVector(start_vals, REALSXP)));
    SEXP hist = PROTECT(Rf_allocVector(REALSXP, mi + 1)), gpsi = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP um = PROTECT(Rf_allocVector(REALSXP, m)), uv = PROTECT(Rf_allocMatrix(REALSXP, m, m));
    int nit = 0;
    const int fam = strcmp(CHAR(STRING_ELT(family, 0)), "poisson") ? SRGP_BERNOULLI : SRGP_POISSON;
    const int st = srgp_laplace_newton(ctx(), fam, k, REAL(xu), m, recycled(muu, m, "muu"),
                                       list_get(cov_par, "sigma"), l, list_get(cov_par, "tau"), Rf_asReal(delta),
                                       Rf_asReal(pois_m), mi, Rf_asReal(tol), REAL(ff), REAL(hist), &nit, REAL(gpsi),
                                       REAL(um), REAL(uv));
    if (st != SRGP_OK) { UNPROTECT(10); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP h = PROTECT(Rf_lengthgets(hist, nit));
    const char *nms[] = {"gp", "objective_function_values", "gradient", "u_posterior_mean", "u_posterior_variance"};
    SEXP vals[] = {ff, h, gpsi, um, uv};
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 5)), nm = PROTECT(Rf_allocVector(STRSXP, 5));
    for (int i = 0; i < 5; i++) { SET_VECTOR_ELT(out, i, vals[i]); SET_STRING_ELT(nm, i, Rf_mkChar(nms[i])); }
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(13);
    return out;
}

/* .Call("_sparseRGPs_predict", cov_fun, x_pred, mu_pred, xu, muu, u_mean, u_var, cov_par, lnames, s22_nugget,
         var_const) -> list(pred_mean, pred_var): bodies of predict_vi / predict_laplace with full_cov = FALSE */
SEXP _sparseRGPs_predict(SEXP cov_fun, SEXP x_pred, SEXP mu_pred, SEXP xu, SEXP muu, SEXP u_mean, SEXP u_var,
                         SEXP cov_par, SEXP lnames, SEXP s22_nugget, SEXP var_const)
{
    x_pred = PROTECT(Rf_coerceVector(x_pred, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    mu_pred = PROTECT(Rf_coerceVector(mu_pred, REALSXP));
    muu = PROTECT(Rf_coerceVector(muu, REALSXP));
    u_mean = PROTECT(Rf_coerceVector(u_mean, REALSXP));
    u_var = PROTECT(Rf_coerceVector(u_var, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(x_pred), d = Rf_ncols(x_pred), m = Rf_nrows(xu);
    if (Rf_length(u_mean) != m || Rf_xlength(u_var) != (R_xlen_t)m * m)
        Rf_error("sparseRGPs: u_mean / u_var do not match the %d knots", m);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    SEXP pm = PROTECT(Rf_allocVector(REALSXP, n)), pv = PROTECT(Rf_allocVector(REALSXP, n));
    const int st = srgp_predict(ctx(), k, REAL(x_pred), n, d, recycled(mu_pred, n, "mu_pred"), REAL(xu),
                                m, recycled(muu, m, "muu"), REAL(u_mean), REAL(u_var),
                                list_get(cov_par, "sigma"), l, Rf_asReal(s22_nugget), Rf_asReal(var_const), REAL(pm),
                                REAL(pv));
    if (st != SRGP_OK) { UNPROTECT(8); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, pm); SET_VECTOR_ELT(out, 1, pv);
    SET_STRING_ELT(nm, 0, Rf_mkChar("pred_mean")); SET_STRING_ELT(nm, 1, Rf_mkChar("pred_var"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(10);
    return out;
}

/* ---- sparse Laplace gradient, optimiser loop, OAT scores; posterior at the knots; objectives from matrices ------ */
/* .Call("_sparseRGPs_laplace_grad", family, cov_fun, xy, y, mu, xu, cov_par, delta, lnames, ff, pois_m)
     -> list(gradient = named numeric): dlogq_dcov_par with dcov_fun_dknot = NA (R/laplace_approx_gradient.R:25-339) */
SEXP _sparseRGPs_laplace_grad(SEXP family, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP cov_par, SEXP delta,
                              SEXP lnames, SEXP ff, SEXP pois_m)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    ff = PROTECT(Rf_coerceVector(ff, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu);
    double l[SRGP_MAX_D], gbuf[SRGP_MAX_D + 2];
    get_l(cov_par, k, d, lnames, l);
    if (Rf_length(ff) != n) Rf_error("sparseRGPs: ff has length %d, expected %d", Rf_length(ff), n);
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_laplace_grad(ctx(), family_id(family), k, REAL(xu), m, list_get(cov_par, "sigma"), l,
                               list_get(cov_par, "tau"), Rf_asReal(delta), Rf_asReal(pois_m), REAL(ff), gbuf);
    if (st != SRGP_OK) { UNPROTECT(5); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP grad = PROTECT(named_gradient(gbuf, k, d, lnames));
    const char *nms[] = {"gradient"};
    SEXP out = named_list(1, nms, &grad);
    UNPROTECT(6);
    return out;
}

/* .Call("_sparseRGPs_laplace_grad_knots", family, cov_fun, xy, y, mu, xu, cov_par, delta, lnames, ff, pois_m,
         knot_bounds, knot_opt, transform) -> list(gradient, knot_gradient, trans_knot): dlogq_dcov_par with a
   dcov_fun_dknot (R/laplace_approx_gradient.R:345-705) */
SEXP _sparseRGPs_laplace_grad_knots(SEXP family, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP cov_par,
                                    SEXP delta, SEXP lnames, SEXP ff, SEXP pois_m, SEXP knot_bounds, SEXP knot_opt,
                                    SEXP transform)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    ff = PROTECT(Rf_coerceVector(ff, REALSXP));
    knot_bounds = PROTECT(Rf_coerceVector(knot_bounds, REALSXP));
    knot_opt = PROTECT(Rf_coerceVector(knot_opt, INTSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu);
    const int tr = Rf_asLogical(transform), n_opt = Rf_length(knot_opt);
    double l[SRGP_MAX_D], gbuf[SRGP_MAX_D + 2];
    get_l(cov_par, k, d, lnames, l);
    if (Rf_length(ff) != n) Rf_error("sparseRGPs: ff has length %d, expected %d", Rf_length(ff), n);
    int *opt0 = (int *)R_alloc(n_opt > 0 ? n_opt : 1, sizeof(int));
    for (int t = 0; t < n_opt; t++) opt0[t] = INTEGER(knot_opt)[t] - 1;
    SEXP kgrad = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)m * d));
    SEXP tknot = PROTECT(Rf_allocMatrix(REALSXP, m, d));
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_laplace_grad_knots(ctx(), family_id(family), k, REAL(xu), m, list_get(cov_par, "sigma"), l,
                                     list_get(cov_par, "tau"), Rf_asReal(delta), Rf_asReal(pois_m), REAL(ff),
                                     tr ? REAL(knot_bounds) : NULL, tr ? REAL(knot_bounds) + d : NULL, opt0, n_opt, gbuf,
                                     REAL(kgrad), REAL(tknot));
    if (st != SRGP_OK) { UNPROTECT(9); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP grad = PROTECT(named_gradient(gbuf, k, d, lnames));
    const char *nms[] = {"gradient", "knot_gradient", "trans_knot"};
    SEXP vals[] = {grad, kgrad, tknot};
    SEXP out = named_list(3, nms, vals);
    UNPROTECT(10);
    return out;
}

/* .Call("_sparseRGPs_laplace_fit", family, cov_fun, xy, y, mu, xu, muu, cov_par_start, delta, lnames, ff, pois_m,
         optim_method, optim_par = c(decay, epsilon, eta, learn_rate), maxit, obj_tol, grad_tol, maxit_nr, tol_nr,
         opt_theta, opt_knots, knot_bounds, knot_opt)
     -> list(cov_par, xu, iter, obj_fun, grad, cov_par_history, nr_iter, fmax, u_mean, u_var): the loop of
   laplace_grad_ascent (R/laplace_gradient_ascent.R:10-628) */
SEXP _sparseRGPs_laplace_fit(SEXP family, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP muu, SEXP cov_par,
                             SEXP delta, SEXP lnames, SEXP ff, SEXP pois_m, SEXP optim_method, SEXP optim_par, SEXP maxit,
                             SEXP obj_tol, SEXP grad_tol, SEXP maxit_nr, SEXP tol_nr, SEXP opt_theta, SEXP opt_knots,
                             SEXP knot_bounds, SEXP knot_opt)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    muu = PROTECT(Rf_coerceVector(muu, REALSXP));
    optim_par = PROTECT(Rf_coerceVector(optim_par, REALSXP));
    knot_bounds = PROTECT(Rf_coerceVector(knot_bounds, REALSXP));
    knot_opt = PROTECT(Rf_coerceVector(knot_opt, INTSXP));
    SEXP xu_out = PROTECT(Rf_duplicate(Rf_coerceVector(xu, REALSXP)));          /* in/out */
    SEXP ff_out = PROTECT(Rf_duplicate(Rf_coerceVector(ff, REALSXP)));          /* in/out */
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu_out);
    const int nl = (k == SRGP_ARD) ? d : 1, p = nl + 2, mi = Rf_asInteger(maxit), n_opt = Rf_length(knot_opt);
    double sigma = list_get(cov_par, "sigma"), tau = list_get(cov_par, "tau"), l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    if (Rf_length(ff_out) != n) Rf_error("sparseRGPs: ff has length %d, expected %d", Rf_length(ff_out), n);
    if (Rf_length(optim_par) < 4) Rf_error("sparseRGPs: optim_par needs c(decay, epsilon, eta, learn_rate)");
    srgp_fit_opt o;
    o.optim_method = strcmp(CHAR(STRING_ELT(optim_method, 0)), "ga") ? SRGP_OPT_ADADELTA : SRGP_OPT_GA;
    o.decay = REAL(optim_par)[0]; o.epsilon = REAL(optim_par)[1]; o.eta = REAL(optim_par)[2];
    o.learn_rate = REAL(optim_par)[3];
    o.maxit = mi; o.obj_tol = Rf_asReal(obj_tol); o.grad_tol = Rf_asReal(grad_tol);
    o.opt_theta = Rf_asLogical(opt_theta); o.opt_knots = Rf_asLogical(opt_knots);
    int *opt0 = (int *)R_alloc(n_opt > 0 ? n_opt : 1, sizeof(int));
    for (int t = 0; t < n_opt; t++) opt0[t] = INTEGER(knot_opt)[t] - 1;
    double *obj_h = (double *)R_alloc(mi > 0 ? mi : 1, sizeof(double));
    double *par_h = (double *)R_alloc((size_t)(mi > 0 ? mi : 1) * p, sizeof(double));
    double *grad_h = (double *)R_alloc((size_t)(mi > 0 ? mi : 1) * p, sizeof(double));
    int *nr_h = (int *)R_alloc(mi > 0 ? mi : 1, sizeof(int));
    SEXP um = PROTECT(Rf_allocVector(REALSXP, m)), uv = PROTECT(Rf_allocMatrix(REALSXP, m, m));
    int iter = 0;
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_laplace_fit(ctx(), family_id(family), k, REAL(xu_out), m, recycled(muu, m, "muu"), &sigma, l, &tau,
                              Rf_asReal(delta), Rf_asReal(pois_m), Rf_asInteger(maxit_nr), Rf_asReal(tol_nr), &o,
                              o.opt_knots ? REAL(knot_bounds) : NULL, o.opt_knots ? REAL(knot_bounds) + d : NULL, opt0,
                              n_opt, REAL(ff_out), &iter, obj_h, par_h, grad_h, nr_h, REAL(um), REAL(uv));
    if (st != SRGP_OK) { UNPROTECT(11); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    SEXP cp = PROTECT(Rf_duplicate(cov_par));      /* same names and order as cov_par_start */
    set_list(cp, "sigma", sigma); set_list(cp, "tau", tau);
    if (k == SRGP_ARD) for (int c = 0; c < d; c++) set_list(cp, CHAR(STRING_ELT(lnames, c)), l[c]);
    else set_list(cp, "l", l[0]);
    SEXP objv = PROTECT(Rf_allocVector(REALSXP, iter)), nrv = PROTECT(Rf_allocVector(INTSXP, iter));
    SEXP gh = PROTECT(Rf_allocMatrix(REALSXP, iter, p)), ph = PROTECT(Rf_allocMatrix(REALSXP, iter, p));
    for (int i = 0; i < iter; i++) {
        REAL(objv)[i] = obj_h[i];
        INTEGER(nrv)[i] = nr_h[i];
        for (int j = 0; j < p; j++) {               /* row-major history -> R's column-major matrix */
            REAL(gh)[i + (size_t)iter * j] = grad_h[(size_t)i * p + j];
            REAL(ph)[i + (size_t)iter * j] = par_h[(size_t)i * p + j];
        }
    }
    const char *nms[] = {"cov_par", "xu", "iter", "obj_fun", "grad", "cov_par_history", "nr_iter", "fmax", "u_mean", "u_var"};
    SEXP it = PROTECT(Rf_ScalarInteger(iter));
    SEXP vals[] = {cp, xu_out, it, objv, gh, ph, nrv, ff_out, um, uv};
    SEXP out = named_list(10, nms, vals);
    UNPROTECT(17);
    return out;
}

/* .Call("_sparseRGPs_laplace_oat_scores", family, cov_fun, xy, y, mu, xu, pseudo_prop, cov_par, delta, lnames, fmax,
         pois_m, maxit_nr, tol_nr) -> list(scores): the candidate loop of knot_prop_random
   (R/knot_proposal_functions.R:1096-1120); NaN marks a candidate whose chol() would have raised */
SEXP _sparseRGPs_laplace_oat_scores(SEXP family, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP pseudo_prop,
                                    SEXP cov_par, SEXP delta, SEXP lnames, SEXP fmax, SEXP pois_m, SEXP maxit_nr,
                                    SEXP tol_nr)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    pseudo_prop = PROTECT(Rf_coerceVector(pseudo_prop, REALSXP));
    fmax = PROTECT(Rf_coerceVector(fmax, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu), T = Rf_nrows(pseudo_prop);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    if (Rf_length(fmax) != n) Rf_error("sparseRGPs: fmax has length %d, expected %d", Rf_length(fmax), n);
    SEXP scores = PROTECT(Rf_allocVector(REALSXP, T));
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_laplace_oat_scores(ctx(), family_id(family), k, REAL(xu), m, REAL(pseudo_prop), T,
                                     list_get(cov_par, "sigma"), l, list_get(cov_par, "tau"), Rf_asReal(delta),
                                     Rf_asReal(pois_m), Rf_asInteger(maxit_nr), Rf_asReal(tol_nr), REAL(fmax), REAL(scores));
    if (st != SRGP_OK) { UNPROTECT(7); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    const char *nms[] = {"scores"};
    SEXP out = named_list(1, nms, &scores);
    UNPROTECT(7);
    return out;
}

/* .Call("_sparseRGPs_gauss_posterior_u", model, cov_fun, xy, y, mu, xu, muu, cov_par, delta, lnames)
     -> list(u_mean, u_var): the tail of norm_grad_ascent_vi (R/vi_functions.R:1160-1180, model 0) / norm_grad_ascent
   (R/laplace_gradient_ascent.R:1637-1656, model 1) */
SEXP _sparseRGPs_gauss_posterior_u(SEXP model, SEXP cov_fun, SEXP xy, SEXP y, SEXP mu, SEXP xu, SEXP muu, SEXP cov_par,
                                   SEXP delta, SEXP lnames)
{
    xy = PROTECT(Rf_coerceVector(xy, REALSXP));
    xu = PROTECT(Rf_coerceVector(xu, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    muu = PROTECT(Rf_coerceVector(muu, REALSXP));
    const int k = kernel_id(cov_fun), n = Rf_nrows(xy), d = Rf_ncols(xy), m = Rf_nrows(xu);
    double l[SRGP_MAX_D];
    get_l(cov_par, k, d, lnames, l);
    SEXP um = PROTECT(Rf_allocVector(REALSXP, m)), uv = PROTECT(Rf_allocMatrix(REALSXP, m, m));
    int st = srgp_set_data(ctx(), REAL(xy), n, d, REAL(y), recycled(mu, n, "mu"));
    if (st == SRGP_OK)
        st = srgp_gauss_posterior_u(ctx(), Rf_asInteger(model), k, REAL(xu), m, recycled(muu, m, "muu"),
                                    list_get(cov_par, "sigma"), l, list_get(cov_par, "tau"), Rf_asReal(delta), REAL(um),
                                    REAL(uv));
    if (st != SRGP_OK) { UNPROTECT(7); Rf_error("sparseRGPs: %s", srgp_last_error()); }
    const char *nms[] = {"u_mean", "u_var"};
    SEXP vals[] = {um, uv};
    SEXP out = named_list(2, nms, vals);
    UNPROTECT(7);
    return out;
}

/* .Call("_sparseRGPs_gauss_obj_mats", Sigma12, Sigma22, Z, y, mu) -> double: the body of obj_fun_norm
   (R/laplace_approx_obj_funs.R:6-52) = elbo_fun without its trace term (R/vi_functions.R:64-121) */
SEXP _sparseRGPs_gauss_obj_mats(SEXP Sigma12, SEXP Sigma22, SEXP Z, SEXP y, SEXP mu)
{
    Sigma12 = PROTECT(Rf_coerceVector(Sigma12, REALSXP));
    Sigma22 = PROTECT(Rf_coerceVector(Sigma22, REALSXP));
    Z = PROTECT(Rf_coerceVector(Z, REALSXP));
    y = PROTECT(Rf_coerceVector(y, REALSXP));
    mu = PROTECT(Rf_coerceVector(mu, REALSXP));
    const int n = Rf_nrows(Sigma12), m = Rf_ncols(Sigma12);
    if (Rf_length(y) != n) Rf_error("sparseRGPs: y has length %d, expected %d", Rf_length(y), n);
    if (Rf_length(Z) != 1 && Rf_length(Z) != n) Rf_error("sparseRGPs: Z has length %d, expected 1 or %d", Rf_length(Z), n);
    if (Rf_length(mu) > 1 && Rf_length(mu) != n) Rf_error("sparseRGPs: mu has length %d, expected 1 or %d", Rf_length(mu), n);
    double obj = NA_REAL;
    const int st = srgp_gauss_obj_mats(ctx(), REAL(Sigma12), n, m, REAL(Sigma22), REAL(Z), Rf_length(Z), REAL(y),
                                       Rf_length(mu) ? REAL(mu) : NULL, Rf_length(mu), &obj);
    UNPROTECT(5);
    if (st != SRGP_OK) Rf_error("sparseRGPs: %s", srgp_last_error());
    return Rf_ScalarReal(obj);
}

static const R_CallMethodDef CallEntries[] = {
    {"_sparseRGPs_real_to_pos", (DL_FUNC)&_sparseRGPs_real_to_pos, 1},
    {"_sparseRGPs_pos_to_real", (DL_FUNC)&_sparseRGPs_pos_to_real, 1},
    {"_sparseRGPs_real_to_bounded", (DL_FUNC)&_sparseRGPs_real_to_bounded, 3},
    {"_sparseRGPs_dsqexp_dsigmaC", (DL_FUNC)&_sparseRGPs_dsqexp_dsigmaC, 3},
    {"_sparseRGPs_dsqexp_dsigma_ardC", (DL_FUNC)&_sparseRGPs_dsqexp_dsigma_ardC, 4},
    {"_sparseRGPs_dsqexp_dlC", (DL_FUNC)&_sparseRGPs_dsqexp_dlC, 3},
    {"_sparseRGPs_dsqexp_dl_ardC", (DL_FUNC)&_sparseRGPs_dsqexp_dl_ardC, 5},
    {"_sparseRGPs_dsqexp_dtauC", (DL_FUNC)&_sparseRGPs_dsqexp_dtauC, 3},
    {"_sparseRGPs_dsqexp_dx2C", (DL_FUNC)&_sparseRGPs_dsqexp_dx2C, 5},
    {"_sparseRGPs_dsqexp_dx2_ardC", (DL_FUNC)&_sparseRGPs_dsqexp_dx2_ardC, 6},
    {"_sparseRGPs_dexp_dsigmaC", (DL_FUNC)&_sparseRGPs_dexp_dsigmaC, 3},
    {"_sparseRGPs_dexp_dlC", (DL_FUNC)&_sparseRGPs_dexp_dlC, 3},
    {"_sparseRGPs_dexp_dtauC", (DL_FUNC)&_sparseRGPs_dexp_dtauC, 3},
    {"_sparseRGPs_dsig_dthetaC", (DL_FUNC)&_sparseRGPs_dsig_dthetaC, 5},
    {"_sparseRGPs_dsig_dtheta_ardC", (DL_FUNC)&_sparseRGPs_dsig_dtheta_ardC, 6},
    {"_sparseRGPs_cov_fun_sqrd_expC", (DL_FUNC)&_sparseRGPs_cov_fun_sqrd_expC, 3},
    {"_sparseRGPs_cov_fun_sqrd_exp_ardC", (DL_FUNC)&_sparseRGPs_cov_fun_sqrd_exp_ardC, 4},
    {"_sparseRGPs_cov_fun_expC", (DL_FUNC)&_sparseRGPs_cov_fun_expC, 3},
    {"_sparseRGPs_make_cov_matC", (DL_FUNC)&_sparseRGPs_make_cov_matC, 5},
    {"_sparseRGPs_make_cov_mat_ardC", (DL_FUNC)&_sparseRGPs_make_cov_mat_ardC, 6},
    {"_sparseRGPs_gauss_obj_grad", (DL_FUNC)&_sparseRGPs_gauss_obj_grad, 9},
    {"_sparseRGPs_gauss_obj_grad_knots", (DL_FUNC)&_sparseRGPs_gauss_obj_grad_knots, 12},
    {"_sparseRGPs_oat_scores", (DL_FUNC)&_sparseRGPs_oat_scores, 10},
    {"_sparseRGPs_gauss_fit", (DL_FUNC)&_sparseRGPs_gauss_fit, 18},
    {"_sparseRGPs_trace_term", (DL_FUNC)&_sparseRGPs_trace_term, 5},
    {"_sparseRGPs_laplace_newton", (DL_FUNC)&_sparseRGPs_laplace_newton, 14},
    {"_sparseRGPs_predict", (DL_FUNC)&_sparseRGPs_predict, 11},
    {"_sparseRGPs_laplace_grad", (DL_FUNC)&_sparseRGPs_laplace_grad, 11},
    {"_sparseRGPs_laplace_grad_knots", (DL_FUNC)&_sparseRGPs_laplace_grad_knots, 14},
    {"_sparseRGPs_laplace_fit", (DL_FUNC)&_sparseRGPs_laplace_fit, 23},
    {"_sparseRGPs_laplace_oat_scores", (DL_FUNC)&_sparseRGPs_laplace_oat_scores, 14},
    {"_sparseRGPs_gauss_posterior_u", (DL_FUNC)&_sparseRGPs_gauss_posterior_u, 10},
    {"_sparseRGPs_gauss_obj_mats", (DL_FUNC)&_sparseRGPs_gauss_obj_mats, 5},
    {NULL, NULL, 0}};

void R_init_sparseRGPs(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_sparseRGPs(DllInfo *dll)
{
    (void)dll;
    if (g_ctx) srgp_ctx_destroy(g_ctx);
    g_ctx = NULL;
}
