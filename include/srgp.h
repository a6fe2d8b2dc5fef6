/*
 * srgp.h -- plain-C ABI of the B200-native sparseRGPs hot path (libsrgp.so).
 *
 * This is the drop-in boundary: every entry point takes plain pointers and sizes (no torch / Rcpp types) and
 * is what the reference's `.Call` layer for this path binds to.  Each declaration cites the reference
 * interface it replaces (paths relative to the reference repo).  INTEGRATION.md shows the R-side `.Call`
 * shim and the ctypes binding.
 *
 * Conventions
 *   - All matrices are R matrices: column-major IEEE doubles, no padding; x(i,c) = x[i + n*c].
 *   - Pointers are HOST pointers unless the name ends in `_dev`; the library copies to/from the GPU inside
 *     the call and never retains a caller pointer after returning.
 *   - Every function returns an int status (SRGP_OK == 0) unless stated; srgp_last_error() gives the message
 *     (thread-local).  There is NO CPU fallback: without a usable sm_100 GPU every compute entry point
 *     returns SRGP_ERR_CUDA.
 *   - `l` is the length-scale array: 1 entry for SRGP_SQEXP / SRGP_EXP, d entries (l1..ld) for SRGP_ARD.
 *   - Derivatives are with respect to log(theta) (the reference's transform = TRUE).
 */
#ifndef SRGP_H
#define SRGP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRGP_VERSION 100

enum srgp_status {
    SRGP_OK = 0,
    SRGP_ERR_ARG = 1,            /* bad argument (NULL pointer, negative size, d > SRGP_MAX_D ...) */
    SRGP_ERR_UNKNOWN_KERNEL = 2, /* reference: message on Rcerr + 0x0 matrix, no R error
                                    (src/covariance_functionsC.cpp:161-168) */
    SRGP_ERR_UNKNOWN_PAR = 3,    /* reference: message + 0x0 matrix (src/covariance_function_derivativesC.cpp:418-422) */
    SRGP_ERR_CUDA = 4,           /* CUDA runtime / launch failure, or no device */
    SRGP_ERR_NOT_PD = 5,         /* an m x m Cholesky failed; the R shim must raise an R error (R's chol() does) */
    SRGP_ERR_STATE = 6,          /* call order (e.g. obj_grad before set_data) */
    SRGP_ERR_COMM = 7,           /* NCCL failure / libnccl not loadable */
    SRGP_ERR_NUMERIC = 8         /* non-finite objective / gradient inside srgp_gauss_fit (R: "missing value where
                                    TRUE/FALSE needed" in the loop condition) */
};

enum srgp_kernel { SRGP_SQEXP = 0, SRGP_EXP = 1, SRGP_ARD = 2 };          /* cov_fun = "sqexp" | "exp" | "ard" */
enum srgp_par { SRGP_PAR_SIGMA = 0, SRGP_PAR_L = 1, SRGP_PAR_TAU = 2, SRGP_PAR_LC = 3 };
enum srgp_model { SRGP_VI = 0, SRGP_FIC = 1 };
enum srgp_family { SRGP_BERNOULLI = 0, SRGP_POISSON = 1 };

#define SRGP_MAX_D 64           /* input dimension limit of the assembly kernels (srgp_make_cov_mat, srgp_dsig_dtheta,
                                   trace / Omega o dK reductions) */
#define SRGP_MAX_D_FUSED 20     /* fused objective+gradient / Laplace / predict passes: the K*M epilogue stages the
                                   coordinates of its row tile and two knot blocks in shared memory, 3 KB per
                                   dimension beside the 4-stage operand pipeline (227 KB per CTA) */
#define SRGP_MAX_D_FUSED_KNOTS 12 /* the same passes with the knot-location gradient (+2 KB per dimension).
                                   Larger d returns SRGP_ERR_ARG with the byte count in srgp_last_error(). The
                                   reference's own examples use d <= 8 (airfoil 5, ccpp 4, copper wire 8). */
#define SRGP_UNIQUE_ID_BYTES 128

typedef struct srgp_ctx srgp_ctx;

/* ---------------------------------------------------------------- library / context ---------- */
int srgp_version(void);
const char *srgp_last_error(void);
/* One context per GPU (one process per GPU under torchrun; several contexts per process also work).
   Holds streams, scratch, the resident data shard and the NCCL communicator. */
int srgp_ctx_create(int device, srgp_ctx **out);
void srgp_ctx_destroy(srgp_ctx *ctx);
int srgp_ctx_sync(srgp_ctx *ctx);

/* ---------------------------------------------------------------- K1: covariance assembly ---- */
/* Replaces make_cov_matC (src/covariance_functionsC.cpp:72-169) and make_cov_mat_ardC (:191-252).
   x is n1 x d.  x_pred == NULL is the reference's `matrix()` (1x1 NA) sentinel: the n1 x n1 self-covariance
   with tau^2 + delta added where i == j.  Otherwise x_pred is n2 x d and out is the n1 x n2 cross-covariance. */
int srgp_make_cov_mat(srgp_ctx *ctx, int kernel, const double *x, int64_t n1, const double *x_pred, int64_t n2,
                      int d, double sigma, const double *l, double tau, double delta, double *out);

/* ---------------------------------------------------------------- K2: d Sigma / d log theta -- */
/* Replaces dsig_dthetaC (src/covariance_function_derivativesC.cpp:307-552) and dsig_dtheta_ardC (:555-722).
   par = SRGP_PAR_SIGMA | SRGP_PAR_L (sqexp, exp) | SRGP_PAR_LC with comp0 = 0-based component (ard) | SRGP_PAR_TAU.
   Reproduces: tau derivative = 2 tau^2 where all coordinates are bit-equal (cross matrices too);
   kernel == SRGP_EXP with a cross matrix returns zeros for anything but sigma / l. */
int srgp_dsig_dtheta(srgp_ctx *ctx, int kernel, int par, int comp0, const double *x, int64_t n1,
                     const double *x_pred, int64_t n2, int d, double sigma, const double *l, double tau,
                     double *out);

/* Device-resident variants (inputs and output already in HBM; `stream-ordered`, call srgp_ctx_sync to wait).
   Used by bench.py for kernel-only timing. */
int srgp_make_cov_mat_dev(srgp_ctx *ctx, int kernel, const double *x_dev, int64_t n1, const double *x_pred_dev,
                          int64_t n2, int d, double sigma, const double *l, double tau, double delta,
                          double *out_dev);
int srgp_dsig_dtheta_dev(srgp_ctx *ctx, int kernel, int par, int comp0, const double *x_dev, int64_t n1,
                         const double *x_pred_dev, int64_t n2, int d, double sigma, const double *l, double tau,
                         double *out_dev);

/* ---------------------------------------------------------------- scalar Rcpp exports -------- */
/* The remaining registered routines (src/RcppExports.cpp:285-304) are per-pair scalars with no R call sites;
   they are host code.  x1, x2 are contiguous length-d vectors. */
void srgp_real_to_pos(const double *x, int64_t n, double *out);                                   /* derivativesC.cpp:11 */
void srgp_pos_to_real(const double *x, int64_t n, double *out);                                   /* :19 */
void srgp_real_to_bounded(const double *x, const double *ub, const double *lb, int64_t n, double *out); /* :27 */
double srgp_cov_fun_sqrd_exp(const double *x1, const double *x2, int d, double sigma, double l);   /* functionsC.cpp:5 */
double srgp_cov_fun_sqrd_exp_ard(const double *x1, const double *x2, int d, double sigma, const double *l); /* :16 */
double srgp_cov_fun_exp(const double *x1, const double *x2, int d, double sigma, double l);        /* :45 */
double srgp_dsqexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l);      /* derivativesC.cpp:35 */
double srgp_dsqexp_dsigma_ard(const double *x1, const double *x2, int d, double sigma, const double *l); /* :55 */
double srgp_dsqexp_dl(const double *x1, const double *x2, int d, double sigma, double l);          /* :86 */
double srgp_dsqexp_dl_ard(const double *x1, const double *x2, int d, double sigma, const double *l, int comp0); /* :107 */
double srgp_dsqexp_dtau(const double *x1, const double *x2, int d, double tau);                    /* :142 */
double srgp_dexp_dsigma(const double *x1, const double *x2, int d, double sigma, double l);        /* :232 */
double srgp_dexp_dl(const double *x1, const double *x2, int d, double sigma, double l);            /* :252 */
double srgp_dexp_dtau(const double *x1, const double *x2, int d, double tau);                      /* :272 */
void srgp_dsqexp_dx2(const double *x1, const double *x2, int d, double sigma, double l, const double *lb,
                     const double *ub, double *deriv, double *trans_par);                          /* :176 */
void srgp_dsqexp_dx2_ard(const double *x1, const double *x2, int d, double sigma, const double *l,
                         const double *lb, const double *ub, double *deriv, double *trans_par);    /* :197 */

/* ---------------------------------------------------------------- K5: trace terms ------------ */
/* trace_term_fun(cov_par, Sigma12, Sigma22, delta)  (R/vi_functions.R:14-27) on materialised inputs:
   -(1/(2 tau^2)) * sum_i (sigma^2 + delta - Sigma12[i,] Sigma22^-1 Sigma12[i,]^T).  Sigma12 is n x m. */
int srgp_trace_term(srgp_ctx *ctx, double sigma, double tau, double delta, const double *Sigma12, int64_t n,
                    int64_t m, const double *Sigma22, double *out);
/* dtrace_term_dcov_par(cov_par, A_trace) (R/vi_functions.R:54-60): -(1/(2 tau^2)) * sum(A_trace). */
int srgp_dtrace_term_dcov_par(srgp_ctx *ctx, double tau, const double *A_trace, int64_t n, double *out);
/* dtrace_term_dtau(cov_par, trace_term) (R/vi_functions.R:38-44): -2 * trace_term. */
double srgp_dtrace_term_dtau(double trace_term);
/* sum_ij Omega_ij * dSigma12_ij/dlog(theta) for every theta at once, never materialising dSigma12:
   the reduction every `dSigma12_dtheta` use in R/vi_functions.R:344-398 collapses to.
   Omega is n x m (host).  out has p entries ordered sigma, l (or l1..ld), tau. */
int srgp_omega_dk_reduce(srgp_ctx *ctx, int kernel, const double *x, int64_t n, const double *xu, int64_t m,
                         int d, double sigma, const double *l, double tau, const double *Omega, double *out);
/* Same with x, xu and Omega already in HBM (kernel-only timing: the kernel reads 8*n*m bytes of Omega once). */
int srgp_omega_dk_reduce_dev(srgp_ctx *ctx, int kernel, const double *x_dev, int64_t n, const double *xu_dev,
                             int64_t m, int d, double sigma, const double *l, double tau, const double *Omega_dev,
                             double *out);

/* ---------------------------------------------------------------- fused objective + gradient - */
/* Resident data shard: xy (n x d), y (n), mu (n, NULL = 0).  Under multi-GPU each rank passes ITS rows. */
int srgp_set_data(srgp_ctx *ctx, const double *xy, int64_t n, int d, const double *y, const double *mu);
/* Device-side generation of the synthetic benchmark workload (BASELINE.json config 5 recipe):
   not part of the reference API; lets bench.py time with inputs resident in HBM. */
int srgp_set_data_dev(srgp_ctx *ctx, const double *xy_dev, int64_t n, int d, const double *y_dev,
                      const double *mu_dev);

/* One optimiser iteration's evaluation for the sparse Gaussian models with fixed knots:
     model == SRGP_VI  : elbo_fun (R/vi_functions.R:64-121) + delbo_dcov_par (:126-420)
     model == SRGP_FIC : obj_fun_norm (R/laplace_approx_obj_funs.R:6-52) + dlogp_dcov_par
                         (R/laplace_approx_gradient.R:720-968)
   xu is m x d.  grad (may be NULL) receives p = d + 2 (ard) or 3 (sqexp) entries ordered like the
   reference's cov_par list: sigma, l / l1..ld, tau; all with respect to log(theta).
   With a communicator attached the row sums are all-reduced (one fused allreduce per pass) and every rank
   returns the global objective and gradient. */
int srgp_gauss_obj_grad(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, double sigma,
                        const double *l, double tau, double delta, double *obj, double *grad);
/* Same, but uploads xy / y / mu first: the one-shot call with the reference's argument list. */
int srgp_gauss_obj_grad_host(srgp_ctx *ctx, int model, int kernel, const double *xy, int64_t n, int d,
                             const double *y, const double *mu, const double *xu, int64_t m, double sigma,
                             const double *l, double tau, double delta, double *obj, double *grad);

/* The same evaluation plus the knot-location gradient: delbo_dcov_par / dlogp_dcov_par called with
   dcov_fun_dknot = dsqexp_dx2 / dsqexp_dx2_ard (R/vi_functions.R:425-592, R/laplace_approx_gradient.R:965-1126,
   closures R/covariance_function_derivatives.R:178-320; installed by R/optimize_gp.R:246,261).
     knot_lb, knot_ub : d bounds each, the reference's knot_bounds = [min - range/10, max + range/10] of xy per
                        dimension (R/vi_functions.R:175-178), computed by the caller from the WHOLE data set;
                        both NULL = transform FALSE (plain d/du, trans_knot = xu).
     knot_opt, n_opt  : 0-based indices of the knots being optimised (R: 1-based `knot_opt`); NULL = all.
                        Other knots get gradient 0.
     knot_grad        : m*d entries, knot-major ([k*d + c]) like the reference's `knot_gradient`.
     trans_knot       : m x d column-major (may be NULL), the reference's `trans_knot`:
                        log(u - lb + 1e-4) - log(ub - u + 1e-4).
   The Jacobian keeps the reference's guard, (ub - lb) / ((u - lb)(ub - u) + 1e-4)  (SURVEY.md quirk Q12).
   grad is required.  Cost: the pass-2 epilogue also sums P_ik (x_ic - u_kc) per knot; no extra pass. */
int srgp_gauss_obj_grad_knots(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, double sigma,
                              const double *l, double tau, double delta, const double *knot_lb,
                              const double *knot_ub, const int *knot_opt, int64_t n_opt, double *obj, double *grad,
                              double *knot_grad, double *trans_knot);

/* OAT candidate scoring: the loops of knot_prop_random_norm_vi (R/vi_functions.R:2211-2298, model SRGP_VI, elbo_fun)
   and knot_prop_random_norm (R/knot_proposal_functions.R:1283-1353, model SRGP_FIC, obj_fun_norm with the FIC Z):
   scores[t] = objective with candidate row t appended to the knots, theta fixed, on the resident shard.
   cand is n_cand x d column-major (the reference samples TTmax rows of xy that are not knots; the sampling stays
   with the caller).  obj0 (may be NULL) receives the objective with the m knots alone.
   SRGP_VI: all candidates are scored from ONE pass over the data (Gram of [knots | candidates]) plus bordered
   Cholesky updates, 128 candidates per pass.  SRGP_FIC: one objective-only evaluation per candidate.
   A candidate whose bordered / full Cholesky fails (the reference's try-error, after which it resamples that
   candidate with jitter) gets scores[t] = NaN; a failure with the CURRENT knots returns SRGP_ERR_NOT_PD. */
int srgp_oat_scores(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, const double *cand,
                    int64_t n_cand, double sigma, const double *l, double tau, double delta, double *obj0,
                    double *scores);

/* The optimiser loops norm_grad_ascent_vi (R/vi_functions.R:596-1218, model SRGP_VI) and norm_grad_ascent
   (R/laplace_gradient_ascent.R:1111-1696, model SRGP_FIC) with a fixed number of knots, transform = TRUE (hard-coded
   in the reference, R/vi_functions.R:640).  Defaults of the reference (`opt_master`, :641-643): adadelta, decay 0.95,
   epsilon 1e-6, learn_rate 1e-2, eta 1e3, maxit 1000, obj_tol 1e-3, grad_tol Inf. */
enum srgp_optim_method { SRGP_OPT_ADADELTA = 0, SRGP_OPT_GA = 1 };
typedef struct srgp_fit_opt {
    int optim_method;                            /* "adadelta" | "ga" */
    double decay, epsilon, eta, learn_rate;      /* optim_par */
    int maxit;
    double obj_tol, grad_tol;                    /* grad_tol = INFINITY disables the gradient test, as in R */
    int opt_theta;                               /* is.list(dcov_fun_dtheta): optimise log(theta) */
    int opt_knots;                               /* is.function(dcov_fun_dknot): optimise the knots too */
} srgp_fit_opt;
/* In/out: xu (m x d column-major), *sigma, l (d entries for ard, 1 otherwise), *tau.  knot_lb / knot_ub / knot_opt as
   in srgp_gauss_obj_grad_knots (required when opt_knots).  *iter_out = the reference's `iter` (evaluations done);
   obj_hist needs maxit entries (`obj_fun`); par_hist / grad_hist (may be NULL) maxit x p, one row per iteration
   (`cov_par_history`, `grad`).  Each iteration is one fused evaluation on the resident shard; the posterior at the
   knots that the reference returns with the fit is srgp_gauss_posterior_u at the returned theta / xu. */
int srgp_gauss_fit(srgp_ctx *ctx, int model, int kernel, double *xu, int64_t m, double *sigma, double *l, double *tau,
                   double delta, const srgp_fit_opt *opt, const double *knot_lb, const double *knot_ub,
                   const int *knot_opt, int64_t n_opt, int *iter_out, double *obj_hist, double *par_hist,
                   double *grad_hist);

/* ---------------------------------------------------------------- posterior at the knots, prediction ---- */
/* Posterior of the process at the knots on the resident shard: the tail of norm_grad_ascent_vi
   (R/vi_functions.R:1160-1180, model SRGP_VI) / norm_grad_ascent (R/laplace_gradient_ascent.R:1637-1656, SRGP_FIC).
   muu (m, NULL = 0) is the prior mean at the knots; u_mean (m) and u_var (m x m) are global under multi-GPU. */
int srgp_gauss_posterior_u(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, const double *muu,
                           double sigma, const double *l, double tau, double delta, double *u_mean, double *u_var);
/* predict_vi (R/vi_functions.R:1222-1336) and predict_laplace (R/laplace_approx_prediction.R:3-123) with
   full_cov = FALSE:  pred_mean = mu_pred + K S22^-1 (u_mean - muu),
                      pred_var_i = var_const + K_i (-S22^-1 + S22^-1 u_var S22^-1) K_i^T,  K = k(x_pred, xu).
   s22_nugget is the diagonal added to K_uu: delta for Gaussian models, tau^2 + delta for the other families;
   var_const is tau^2 + sigma^2 + delta (predict_vi) or sigma^2 + tau^2 (predict_laplace).  K is never built. */
int srgp_predict(srgp_ctx *ctx, int kernel, const double *x_pred, int64_t n_pred, int d, const double *mu_pred,
                 const double *xu, int64_t m, const double *muu, const double *u_mean, const double *u_var,
                 double sigma, const double *l, double s22_nugget, double var_const, double *pred_mean,
                 double *pred_var);

/* ---------------------------------------------------------------- sparse Laplace ------------- */
/* newtrap_sparseGP (R/newtrap_sparseGP.R:6-186) on the resident shard.  ff (n) in: start values, out: mode.
   obj_hist receives up to maxit objective values, *n_iter their count.  grad_psi (n), u_mean (m),
   u_var (m x m) may be NULL.  pois_m is the Poisson offset `m` (ignored for Bernoulli). */
int srgp_laplace_newton(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, const double *muu,
                        double sigma, const double *l, double tau, double delta, double pois_m, int maxit,
                        double tol, double *ff, double *obj_hist, int *n_iter, double *grad_psi, double *u_mean,
                        double *u_var);
/* dlogq_dcov_par (R/laplace_approx_gradient.R:25-339) at the mode ff. */
int srgp_laplace_grad(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                      const double *l, double tau, double delta, double pois_m, const double *ff, double *grad);

/* dlogq_dcov_par with dcov_fun_dknot (R/laplace_approx_gradient.R:345-705): the same gradient plus the knot-location
   gradient at the mode ff.  Knot arguments and outputs as in srgp_gauss_obj_grad_knots. */
int srgp_laplace_grad_knots(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                            const double *l, double tau, double delta, double pois_m, const double *ff,
                            const double *knot_lb, const double *knot_ub, const int *knot_opt, int64_t n_opt,
                            double *grad, double *knot_grad, double *trans_knot);

/* laplace_grad_ascent (R/laplace_gradient_ascent.R:10-628): the optimiser loop of srgp_gauss_fit around the sparse
   Laplace evaluation -- a Newton mode search (newtrap_sparseGP with maxit_nr / tol_nr, warm-started from the previous
   mode) and dlogq_dcov_par at that mode per iteration; the objective is the last Newton objective value.
   In/out: xu, *sigma, l, *tau and ff (n: start values in, `fmax` out).  nr_iter (may be NULL) needs maxit entries
   (`nr_iter`: Newton iterations per step); u_mean (m) / u_var (m x m) (may be NULL) are those of the last Newton run.
   Other arguments as in srgp_gauss_fit.  Reference defaults: maxit_nr 1000, tol_nr 1e-6 (:77-80). */
int srgp_laplace_fit(srgp_ctx *ctx, int family, int kernel, double *xu, int64_t m, const double *muu, double *sigma,
                     double *l, double *tau, double delta, double pois_m, int maxit_nr, double tol_nr,
                     const srgp_fit_opt *opt, const double *knot_lb, const double *knot_ub, const int *knot_opt,
                     int64_t n_opt, double *ff, int *iter_out, double *obj_hist, double *par_hist, double *grad_hist,
                     int *nr_iter, double *u_mean, double *u_var);

/* Candidate loop of knot_prop_random (R/knot_proposal_functions.R:1096-1120) for the sparse Laplace models:
   scores[t] = last Newton objective with candidate row t appended to the knots, every search warm-started from
   fmax (n, the current mode).  cand is n_cand x d column-major.  NaN = that candidate's Cholesky failed. */
int srgp_laplace_oat_scores(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, const double *cand,
                            int64_t n_cand, double sigma, const double *l, double tau, double delta, double pois_m,
                            int maxit_nr, double tol_nr, const double *fmax, double *scores);

/* ---------------------------------------------------------------- multi-GPU ------------------ */
/* Row sharding over ranks with NCCL sum-allreduce of the pass partials (m x m Gram, m-vectors, scalars,
   gradient partials).  Rank 0 calls srgp_comm_unique_id and distributes the bytes out of band
   (bench.py: torch.distributed broadcast); every rank then calls srgp_comm_init. */
int srgp_comm_unique_id(char id[SRGP_UNIQUE_ID_BYTES]);
int srgp_comm_init(srgp_ctx *ctx, int world, int rank, const char id[SRGP_UNIQUE_ID_BYTES]);
int srgp_comm_destroy(srgp_ctx *ctx);

/* ---------------------------------------------------------------- objectives from materialised matrices ---- */
/* Bodies of obj_fun_norm(ff, mu, Z, Sigma12, Sigma22, y, ...) (R/laplace_approx_obj_funs.R:6-52) and of elbo_fun
   (R/vi_functions.R:64-121) WITHOUT its trace term (elbo_fun adds trace_term_fun(...), see r/patches.R):
     -1/2 r' Z^-1 r + 1/2 b' (S22 + G)^-1 b - 1/2 (sum log Z - log|S22| + log|S22 + G|) - n/2 log 2 pi
   with G = S12' diag(1/Z) S12, b = S12' (r / Z), r = y - mu.  Sigma12 is n x m, Sigma22 m x m (column-major, as the R
   callers built them with make_cov_mat*C); Z has length nz = 1 or n, mu length nmu = 0 (zero mean), 1 or n (R's
   recycling).  Drops the resident data shard of the context (call srgp_set_data again before a fused evaluation). */
int srgp_gauss_obj_mats(srgp_ctx *ctx, const double *Sigma12, int64_t n, int64_t m, const double *Sigma22,
                        const double *Z, int64_t nz, const double *y, const double *mu, int64_t nmu, double *obj);

/* ---------------------------------------------------------------- instrumentation ------------ */
/* Device memory helpers for callers without a CUDA binding (bench.py, tests). */
int srgp_dev_alloc(srgp_ctx *ctx, int64_t bytes, void **out_dev);
int srgp_dev_free(srgp_ctx *ctx, void *dev);
int srgp_memcpy_h2d(srgp_ctx *ctx, void *dst_dev, const void *src, int64_t bytes);
int srgp_memcpy_d2h(srgp_ctx *ctx, void *dst, const void *src_dev, int64_t bytes);
int srgp_fill_normal_dev(srgp_ctx *ctx, double *dst_dev, int64_t n, uint64_t seed, double mean, double sd);
/* CUDA-event timers on the context's own stream (torch.cuda.Event would not see it). */
int srgp_timer_start(srgp_ctx *ctx);
int srgp_timer_stop_ms(srgp_ctx *ctx, double *ms);
/* Per-kernel accounting: when enabled, every launch of a named hot kernel is bracketed by CUDA events;
   srgp_prof_get returns launches and accumulated milliseconds since the last reset. */
enum srgp_prof_id {
    SRGP_PROF_ASSEMBLE = 0, SRGP_PROF_GEN = 1, SRGP_PROF_GRAM = 2, SRGP_PROF_KM = 3, SRGP_PROF_DENSE = 4,
    SRGP_PROF_REDUCE = 5, SRGP_PROF_COMM = 6, SRGP_PROF_COUNT = 8
};
int srgp_prof_enable(srgp_ctx *ctx, int on);
int srgp_prof_reset(srgp_ctx *ctx);
int srgp_prof_get(srgp_ctx *ctx, int id, int64_t *launches, double *ms);
/* Total kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t srgp_launch_count(srgp_ctx *ctx);
/* Flush L2 by writing a scratch buffer larger than the 126 MB L2 (timing hygiene between iterations). */
int srgp_flush_l2(srgp_ctx *ctx);
/* Measured INT8 tensor-pipe peak (bench.py's roofline denominator; MEASURED_PEAKS.json has no INT8 entry): every SM
   issues `iters` x 32 resident tcgen05.mma.kind::i8 128x128x32.  tops = 1e-12 INT8 op/s from CUDA-event time,
   cycles_per_mma = SM cycles per MMA on SM 0 (64 = 8192 MAC/clk/SM).  No reference counterpart. */
int srgp_probe_i8_peak(srgp_ctx *ctx, int iters, double *tops, double *cycles_per_mma);
/* Digit slices per operand of the INT8 row passes (compile-time SRGP_I8_NS of csrc/tc_i8.cuh: 7, or 8 in the validation
   build): NS (NS + 1) / 2 exact INT8 MMAs stand for one FP64 MMA.  Returns the count, not a status. */
int srgp_i8_slices(void);

#ifdef __cplusplus
}
#endif
#endif /* SRGP_H */
