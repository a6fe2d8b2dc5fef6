// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Plain-pointer C entry points in front of the REFERENCE's own C++ functions, compiled unmodified from
//   /root/reference/src/covariance_functionsC.cpp
//   /root/reference/src/covariance_function_derivativesC.cpp
// against oracle/rcpp_shim/Rcpp.h (see that header for what the shim is and is not). This file contains no
// arithmetic of its own: it converts pointers to the shim's NumericMatrix / List / String types -- the job
// src/RcppExports.cpp:10-282 does with SEXPs -- calls the reference function, and copies the result out.
// It also binds the two R closures the reference resolves through Rcpp::Function (R/RcppExports.R:12-14,24-26)
// to the same translation unit's C++ functions, which is what the installed package does.
//
// Built by oracle/ref_native.py (build()) into oracle/_ref/libsparseRGPs_ref.so (git-ignored, travels with gpurun).
// Only tests/ (and tests/tools/make_golden.py) load it.
#include <Rcpp.h>

#include <cstring>
#include <string>

using namespace Rcpp;

// ---- the reference's exported functions (signatures: src/RcppExports.cpp:10-282) ---------------------------------
NumericVector real_to_pos(NumericVector x);
NumericVector pos_to_real(NumericVector x);
NumericVector real_to_bounded(NumericVector x, NumericVector ub, NumericVector lb);
double cov_fun_sqrd_expC(NumericVector x1, NumericVector x2, List cov_par);
double cov_fun_sqrd_exp_ardC(NumericVector x1, NumericVector x2, List cov_par, StringVector lnames);
double cov_fun_expC(NumericVector x1, NumericVector x2, List cov_par);
NumericMatrix make_cov_matC(NumericMatrix x, NumericMatrix x_pred, List cov_par, String cov_fun, double delta);
NumericMatrix make_cov_mat_ardC(NumericMatrix x, NumericMatrix x_pred, List cov_par, String cov_fun, double delta,
                                StringVector lnames);
List dsqexp_dsigmaC(NumericVector x1, NumericVector x2, List cov_par);
List dsqexp_dsigma_ardC(NumericVector x1, NumericVector x2, List cov_par, StringVector lnames);
List dsqexp_dlC(NumericVector x1, NumericVector x2, List cov_par);
List dsqexp_dl_ardC(NumericVector x1, NumericVector x2, List cov_par, StringVector lnames, double comp);
List dsqexp_dtauC(NumericVector x1, NumericVector x2, List cov_par);
List dsqexp_dx2C(NumericVector x1, NumericVector x2, List cov_par, NumericVector lb, NumericVector ub);
List dsqexp_dx2_ardC(NumericVector x1, NumericVector x2, List cov_par, NumericVector lb, NumericVector ub,
                     StringVector lnames);
List dexp_dsigmaC(NumericVector x1, NumericVector x2, List cov_par);
List dexp_dlC(NumericVector x1, NumericVector x2, List cov_par);
List dexp_dtauC(NumericVector x1, NumericVector x2, List cov_par);
NumericMatrix dsig_dthetaC(NumericMatrix x, NumericMatrix x_pred, List cov_par, String cov_fun, String par_name);
NumericMatrix dsig_dtheta_ardC(NumericMatrix x, NumericMatrix x_pred, List cov_par, String cov_fun, String par_name,
                               StringVector lnames);

namespace {

thread_local std::string g_err;

void bind_closures() {
  static bool done = false;
  if (done) return;
  done = true;
  // R/RcppExports.R:12  real_to_pos <- function(x) .Call('_sparseRGPs_real_to_pos', x)
  shim::global_env()["real_to_pos"] = [](const std::vector<RObject>& a) {
    return RObject(::real_to_pos(a.at(0).vec()));
  };
  // R/RcppExports.R:24  real_to_bounded <- function(x, ub, lb) .Call('_sparseRGPs_real_to_bounded', x, ub, lb)
  shim::global_env()["real_to_bounded"] = [](const std::vector<RObject>& a) {
    return RObject(::real_to_bounded(a.at(0).vec(), a.at(1).vec(), a.at(2).vec()));
  };
}

List make_list(const char* const* names, const double* values, int npar) {
  List l;
  for (int i = 0; i < npar; i++) l.push_back(names[i], RObject(values[i]));
  return l;
}
StringVector make_names(const char* const* lnames, int nl) {
  StringVector s;
  for (int i = 0; i < nl; i++) s.push_back(String(lnames[i]));
  return s;
}
// R's `matrix()` sentinel: 1x1 logical NA, coerced to NA_real_ by Rcpp's input_parameter.
NumericMatrix make_x_pred(const double* x_pred, int n2, int d) {
  if (x_pred == nullptr) {
    NumericMatrix na(1, 1);
    na(0, 0) = std::numeric_limits<double>::quiet_NaN();
    return na;
  }
  return NumericMatrix(n2, d, x_pred);
}
NumericVector make_vec(const double* p, int n) { return p ? NumericVector(p, p + n) : NumericVector(); }

int copy_out(const NumericMatrix& m, double* out, long long out_cap, int* nrow, int* ncol) {
  *nrow = m.nrow();
  *ncol = m.ncol();
  long long cnt = (long long)m.nrow() * (long long)m.ncol();
  if (cnt > out_cap) {
    g_err = "output buffer too small";
    return 2;
  }
  if (cnt) std::memcpy(out, m.begin(), sizeof(double) * (size_t)cnt);
  return 0;
}

template <typename F> int guarded(F f) {
  try {
    bind_closures();
    return f();
  } catch (const std::exception& e) {   // BEGIN_RCPP / END_RCPP turn these into R errors
    g_err = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

int ref_transform(int which, const double* x, const double* ub, const double* lb, int n, double* out) {
  return guarded([&] {
    NumericVector r = which == 0   ? ::real_to_pos(make_vec(x, n))
                      : which == 1 ? ::pos_to_real(make_vec(x, n))
                                   : ::real_to_bounded(make_vec(x, n), make_vec(ub, n), make_vec(lb, n));
    for (int i = 0; i < r.size(); i++) out[i] = r[i];
    return 0;
  });
}

// which: 0 cov_fun_sqrd_expC, 1 cov_fun_sqrd_exp_ardC, 2 cov_fun_expC
int ref_cov_fun(int which, const double* x1, const double* x2, int d, const char* const* names, const double* values,
                int npar, const char* const* lnames, int nl, double* out) {
  return guarded([&] {
    List cp = make_list(names, values, npar);
    NumericVector a = make_vec(x1, d), b = make_vec(x2, d);
    *out = which == 0   ? cov_fun_sqrd_expC(a, b, cp)
           : which == 1 ? cov_fun_sqrd_exp_ardC(a, b, cp, make_names(lnames, nl))
                        : cov_fun_expC(a, b, cp);
    return 0;
  });
}

// which: 0 dsqexp_dsigmaC, 1 dsqexp_dsigma_ardC, 2 dsqexp_dlC, 3 dsqexp_dl_ardC, 4 dsqexp_dtauC, 5 dsqexp_dx2C,
//        6 dsqexp_dx2_ardC, 7 dexp_dsigmaC, 8 dexp_dlC, 9 dexp_dtauC.
// Each output holds `*len` doubles (1, or d for the dx2 pair): list elements derivative / trans_par / inv_trans_par.
int ref_pair_derivative(int which, const double* x1, const double* x2, int d, const char* const* names,
                        const double* values, int npar, const char* const* lnames, int nl, double comp,
                        const double* lb, const double* ub, double* derivative, double* trans_par,
                        double* inv_trans_par, int* len) {
  return guarded([&] {
    List cp = make_list(names, values, npar);
    NumericVector a = make_vec(x1, d), b = make_vec(x2, d);
    StringVector ln = make_names(lnames, nl);
    List r;
    switch (which) {
      case 0: r = dsqexp_dsigmaC(a, b, cp); break;
      case 1: r = dsqexp_dsigma_ardC(a, b, cp, ln); break;
      case 2: r = dsqexp_dlC(a, b, cp); break;
      case 3: r = dsqexp_dl_ardC(a, b, cp, ln, comp); break;
      case 4: r = dsqexp_dtauC(a, b, cp); break;
      case 5: r = dsqexp_dx2C(a, b, cp, make_vec(lb, d), make_vec(ub, d)); break;
      case 6: r = dsqexp_dx2_ardC(a, b, cp, make_vec(lb, d), make_vec(ub, d), ln); break;
      case 7: r = dexp_dsigmaC(a, b, cp); break;
      case 8: r = dexp_dlC(a, b, cp); break;
      case 9: r = dexp_dtauC(a, b, cp); break;
      default: g_err = "unknown function index"; return 3;
    }
    NumericVector dv = r["derivative"].vec(), tp = r["trans_par"].vec(), ip = r["inv_trans_par"].vec();
    *len = dv.size();
    for (int i = 0; i < dv.size(); i++) derivative[i] = dv[i];
    for (int i = 0; i < tp.size() && i < dv.size(); i++) trans_par[i] = tp[i];
    for (int i = 0; i < ip.size() && i < dv.size(); i++) inv_trans_par[i] = ip[i];
    return 0;
  });
}

// lnames == NULL -> make_cov_matC, else make_cov_mat_ardC. x_pred == NULL -> the `matrix()` sentinel.
int ref_make_cov_mat(const double* x, int n1, int d, const double* x_pred, int n2, const char* const* names,
                     const double* values, int npar, const char* cov_fun, double delta, const char* const* lnames,
                     int nl, double* out, long long out_cap, int* nrow, int* ncol) {
  return guarded([&] {
    NumericMatrix X(n1, d, x), XP = make_x_pred(x_pred, n2, d);
    List cp = make_list(names, values, npar);
    NumericMatrix m = lnames ? make_cov_mat_ardC(X, XP, cp, String(cov_fun), delta, make_names(lnames, nl))
                             : make_cov_matC(X, XP, cp, String(cov_fun), delta);
    return copy_out(m, out, out_cap, nrow, ncol);
  });
}

// lnames == NULL -> dsig_dthetaC, else dsig_dtheta_ardC.
int ref_dsig_dtheta(const double* x, int n1, int d, const double* x_pred, int n2, const char* const* names,
                    const double* values, int npar, const char* cov_fun, const char* par_name,
                    const char* const* lnames, int nl, double* out, long long out_cap, int* nrow, int* ncol) {
  return guarded([&] {
    NumericMatrix X(n1, d, x), XP = make_x_pred(x_pred, n2, d);
    List cp = make_list(names, values, npar);
    NumericMatrix m = lnames ? dsig_dtheta_ardC(X, XP, cp, String(cov_fun), String(par_name), make_names(lnames, nl))
                             : dsig_dthetaC(X, XP, cp, String(cov_fun), String(par_name));
    return copy_out(m, out, out_cap, nrow, ncol);
  });
}

}  // extern "C"
