// fit.cu -- the optimiser loops around the fused evaluation.
//
// Reference: norm_grad_ascent_vi (R/vi_functions.R:596-1218) and norm_grad_ascent
// (R/laplace_gradient_ascent.R:1111-1696): gradient ascent ("ga") or ADADELTA with sign-flip damping on
// log(theta) and, optionally, on the bounded-logit knot coordinates; every iteration re-assembles Sigma12 / Sigma22,
// evaluates the objective and then the gradient.  Here one iteration = one fused evaluation on the resident shard
// (srgp_gauss_obj_grad / _knots); the p + m*d numbers of optimiser state live on the host and never leave the
// library, so a fit costs one API call instead of 2 * iter R-level rebuilds.
//
// Transcribed details (R/vi_functions.R line numbers; the FIC loop is the same text):
//   stop rule :963-965  iter < maxit && (any(|g_theta, g_knot| > grad_tol) || (iter > 1 ? |obj - obj_prev| > obj_tol : TRUE))
//   ADADELTA  :981-986  sg2 = decay sg2 + (1 - decay) g^2; step = (1/eta)^sc * sqrt(sd2 + eps)/sqrt(sg2 + eps) * g;
//                       sd2 = decay sd2 + (1 - decay) step^2; trans += step          (ascent)
//   knots     :1031-1038 xu_trans (never re-derived from xu after the first evaluation) += step, then
//                       xu = ub/(1 + exp(-t)) + lb/(1 + exp(t))  (R/covariance_function_derivatives.R:189-194)
//   theta     :1042-1054 theta = exp(trans); after the evaluation trans is reset to log(theta) (:1134)
//   sign flip :1132-1133 sc_theta = decay sc + (1 - decay) |sign(g_new) - sign(g_old)| / 2
//             :1146-1147 sc_knot  = decay sc + (1 - decay) |sign(g_new) - sign(g_old)|        (no / 2, as written)
#include <math.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "gauss.cuh"
#include "srgp.h"

using namespace srgp;

namespace {
inline double sgn(double v) { return (v > 0.0) - (v < 0.0); }   // R's sign(); NaN is caught before it gets here
}

// One optimiser run.  th: p parameters (in/out) in the reference's order; xu: m x d column-major (in/out);
// evaluate(obj, g_theta[p], g_knot[m*d]) runs one evaluation at the current th / xu.
typedef std::function<int(double *, double *, double *)> EvalFn;

static int fit_loop(const srgp_fit_opt *opt, int p, int64_t m, int d, std::vector<double> &th, double *xu,
                    const double *knot_lb, const double *knot_ub, const int *knot_opt, int64_t n_opt,
                    const EvalFn &evaluate_raw, int *iter_out, double *obj_hist, double *par_hist, double *grad_hist)
{
    const int64_t md = m * d;
    const bool ok = opt->opt_knots != 0, ot = opt->opt_theta != 0;
    std::vector<char> keep;
    if (ok && knot_opt) {
        keep.assign((size_t)m, 0);
        for (int64_t t = 0; t < n_opt; t++) {
            if (knot_opt[t] < 0 || knot_opt[t] >= m) {
                set_error("knot_opt[%lld] = %d outside [0, %lld)", (long long)t, knot_opt[t], (long long)m);
                return SRGP_ERR_ARG;
            }
            keep[knot_opt[t]] = 1;
        }
    }
    std::vector<double> trans(p), g(p, 0.0), gnew(p), sg2(p, 0.0), sd2(p, 0.0), sc(p, 0.0);
    std::vector<double> gk(ok ? md : 0, 0.0), gknew(ok ? md : 0), ksg2(ok ? md : 0, 0.0), ksd2(ok ? md : 0, 0.0),
        ksc(ok ? md : 0, 0.0), xt(ok ? md : 0);

    auto evaluate = [&](double *obj) -> int {
        int rc = evaluate_raw(obj, gnew.data(), ok ? gknew.data() : nullptr);
        if (rc != SRGP_OK) return rc;
        if (ok && !keep.empty())
            for (int64_t k = 0; k < m; k++)
                if (!keep[k])
                    for (int c = 0; c < d; c++) gknew[k * d + c] = 0.0;
        bool bad = !isfinite(*obj);
        for (int j = 0; j < p && !bad; j++) bad = !isfinite(gnew[j]);
        for (int64_t j = 0; j < (ok ? md : 0) && !bad; j++) bad = !isfinite(gknew[j]);
        if (bad) {   // R: "missing value where TRUE/FALSE needed" in the while() condition
            set_error("objective or gradient is not finite");
            return SRGP_ERR_NUMERIC;
        }
        return SRGP_OK;
    };
    auto record = [&](int it, double obj) {
        obj_hist[it - 1] = obj;
        if (par_hist)
            for (int j = 0; j < p; j++) par_hist[(size_t)(it - 1) * p + j] = th[j];
        if (grad_hist)
            for (int j = 0; j < p; j++) grad_hist[(size_t)(it - 1) * p + j] = ot ? g[j] : 0.0;
    };

    double obj = NAN;
    SRGP_TRY(evaluate(&obj));
    if (ot) g = gnew;
    if (ok) {
        gk = gknew;
        // trans_knot of the first evaluation: inv_trans_fun with its 1e-4 guards (quirk Q12); xu is m x d column-major
        for (int64_t k = 0; k < m; k++)
            for (int c = 0; c < d; c++) {
                const double u = xu[k + m * c];
                xt[k * d + c] = log((u - knot_lb[c]) + 1e-4) - log((knot_ub[c] - u) + 1e-4);
            }
    }
    for (int j = 0; j < p; j++) trans[j] = log(th[j]);
    int it = 1;
    record(it, obj);
    const double decay = opt->decay, eps = opt->epsilon, ieta = 1.0 / opt->eta, lr = opt->learn_rate;
    for (;;) {
        bool big = false;
        if (ot)
            for (int j = 0; j < p; j++) big = big || fabs(g[j]) > opt->grad_tol;
        if (ok)
            for (int64_t j = 0; j < md && !big; j++) big = fabs(gk[j]) > opt->grad_tol;
        const bool moving = it > 1 ? fabs(obj - obj_hist[it - 2]) > opt->obj_tol : true;
        if (!(it < opt->maxit && (big || moving))) break;
        it++;
        if (ot)
            for (int j = 0; j < p; j++) {
                double step;
                if (opt->optim_method == SRGP_OPT_ADADELTA) {
                    sg2[j] = decay * sg2[j] + (1.0 - decay) * g[j] * g[j];
                    step = pow(ieta, sc[j]) * (sqrt(sd2[j] + eps) / sqrt(sg2[j] + eps)) * g[j];
                    sd2[j] = decay * sd2[j] + (1.0 - decay) * step * step;
                } else {
                    step = lr * g[j];
                }
                trans[j] += step;
            }
        if (ok) {
            for (int64_t j = 0; j < md; j++) {
                double step;
                if (opt->optim_method == SRGP_OPT_ADADELTA) {
                    ksg2[j] = decay * ksg2[j] + (1.0 - decay) * gk[j] * gk[j];
                    step = pow(ieta, ksc[j]) * (sqrt(ksd2[j] + eps) / sqrt(ksg2[j] + eps)) * gk[j];
                    ksd2[j] = decay * ksd2[j] + (1.0 - decay) * step * step;
                } else {
                    step = lr * gk[j];
                }
                xt[j] += step;
            }
            for (int64_t k = 0; k < m; k++)
                for (int c = 0; c < d; c++) {
                    const double t = xt[k * d + c];
                    xu[k + m * c] = knot_ub[c] * (1.0 / (1.0 + exp(-t))) + knot_lb[c] * (1.0 / (1.0 + exp(t)));
                }
        }
        if (ot)
            for (int j = 0; j < p; j++) th[j] = exp(trans[j]);
        SRGP_TRY(evaluate(&obj));
        if (ot) {
            if (opt->optim_method == SRGP_OPT_ADADELTA)
                for (int j = 0; j < p; j++) sc[j] = decay * sc[j] + (1.0 - decay) * fabs(sgn(gnew[j]) - sgn(g[j])) / 2.0;
            g = gnew;
            for (int j = 0; j < p; j++) trans[j] = log(th[j]);
        }
        if (ok) {
            if (opt->optim_method == SRGP_OPT_ADADELTA)
                for (int64_t j = 0; j < md; j++) ksc[j] = decay * ksc[j] + (1.0 - decay) * fabs(sgn(gknew[j]) - sgn(gk[j]));
            gk = gknew;
        }
        record(it, obj);
    }
    *iter_out = it;
    return SRGP_OK;
}

static bool fit_args_ok(const srgp_fit_opt *opt, const double *knot_lb, const double *knot_ub, const int *knot_opt,
                        int64_t n_opt)
{
    return opt && opt->maxit >= 1 && (opt->optim_method == SRGP_OPT_ADADELTA || opt->optim_method == SRGP_OPT_GA) &&
           !(opt->opt_knots && (!knot_lb || !knot_ub)) && n_opt >= 0 && !(n_opt > 0 && !knot_opt);
}

extern "C" int srgp_gauss_fit(srgp_ctx *ctx, int model, int kernel, double *xu, int64_t m, double *sigma, double *l,
                              double *tau, double delta, const srgp_fit_opt *opt, const double *knot_lb,
                              const double *knot_ub, const int *knot_opt, int64_t n_opt, int *iter_out,
                              double *obj_hist, double *par_hist, double *grad_hist)
{
    if (!ctx || !xu || !sigma || !l || !tau || !iter_out || !obj_hist || m <= 0 ||
        !fit_args_ok(opt, knot_lb, knot_ub, knot_opt, n_opt)) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("srgp_gauss_fit called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    const int d = ctx->d;
    const bool ard = (kernel == SRGP_ARD);
    const int nl = ard ? d : 1, p = nl + 2;
    std::vector<double> th(p), lfull(std::max(d, 1));   // theta in the reference's order: sigma, l / l1..ld, tau
    th[0] = *sigma;
    for (int c = 0; c < nl; c++) th[1 + c] = l[c];
    th[p - 1] = *tau;
    const bool ok = opt->opt_knots != 0;
    EvalFn eval = [&](double *obj, double *g, double *gk) -> int {
        for (int c = 0; c < d; c++) lfull[c] = th[1 + (ard ? c : 0)];
        return gauss_eval(ctx, model, kernel, xu, m, th[0], lfull.data(), th[p - 1], delta, obj, g, ok, knot_lb, knot_ub,
                          gk);
    };
    SRGP_TRY(fit_loop(opt, p, m, d, th, xu, knot_lb, knot_ub, knot_opt, n_opt, eval, iter_out, obj_hist, par_hist,
                      grad_hist));
    *sigma = th[0];
    for (int c = 0; c < nl; c++) l[c] = th[1 + c];
    *tau = th[p - 1];
    return SRGP_OK;
}

// laplace_grad_ascent (R/laplace_gradient_ascent.R:10-628): the same optimiser skeleton; one evaluation = a Newton
// mode search warm-started from the previous mode (newtrap_sparseGP, :475-487) + dlogq_dcov_par at that mode
// (:493-508); the objective is the last value of the Newton history (:489).
extern "C" int srgp_laplace_fit(srgp_ctx *ctx, int family, int kernel, double *xu, int64_t m, const double *muu,
                                double *sigma, double *l, double *tau, double delta, double pois_m, int maxit_nr,
                                double tol_nr, const srgp_fit_opt *opt, const double *knot_lb, const double *knot_ub,
                                const int *knot_opt, int64_t n_opt, double *ff, int *iter_out, double *obj_hist,
                                double *par_hist, double *grad_hist, int *nr_iter, double *u_mean, double *u_var)
{
    if (!ctx || !xu || !sigma || !l || !tau || !ff || !iter_out || !obj_hist || m <= 0 || maxit_nr < 1 ||
        !fit_args_ok(opt, knot_lb, knot_ub, knot_opt, n_opt)) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("srgp_laplace_fit called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    const int d = ctx->d;
    const bool ard = (kernel == SRGP_ARD);
    const int nl = ard ? d : 1, p = nl + 2;
    std::vector<double> th(p), lfull(std::max(d, 1)), nr_hist((size_t)maxit_nr + 1);
    th[0] = *sigma;
    for (int c = 0; c < nl; c++) th[1 + c] = l[c];
    th[p - 1] = *tau;
    const bool ok = opt->opt_knots != 0;
    int evals = 0;
    EvalFn eval = [&](double *obj, double *g, double *gk) -> int {
        for (int c = 0; c < d; c++) lfull[c] = th[1 + (ard ? c : 0)];
        int n_nr = 0;
        SRGP_TRY(srgp_laplace_newton(ctx, family, kernel, xu, m, muu, th[0], lfull.data(), th[p - 1], delta, pois_m,
                                     maxit_nr, tol_nr, ff, nr_hist.data(), &n_nr, nullptr, u_mean, u_var));
        *obj = nr_hist[n_nr - 1];
        if (nr_iter) nr_iter[evals] = n_nr;
        evals++;
        if (ok)
            return srgp_laplace_grad_knots(ctx, family, kernel, xu, m, th[0], lfull.data(), th[p - 1], delta, pois_m, ff,
                                           knot_lb, knot_ub, nullptr, 0, g, gk, nullptr);
        return srgp_laplace_grad(ctx, family, kernel, xu, m, th[0], lfull.data(), th[p - 1], delta, pois_m, ff, g);
    };
    SRGP_TRY(fit_loop(opt, p, m, d, th, xu, knot_lb, knot_ub, knot_opt, n_opt, eval, iter_out, obj_hist, par_hist,
                      grad_hist));
    *sigma = th[0];
    for (int c = 0; c < nl; c++) l[c] = th[1 + c];
    *tau = th[p - 1];
    return SRGP_OK;
}

// knot_prop_random (R/knot_proposal_functions.R:1001-1175), candidate loop :1096-1120: for every candidate row a Newton
// mode search with the knots [U; c], warm-started from the current mode; score = last Newton objective value.
// A failed Cholesky (the reference's try-error, after which it resamples) gives NaN.
extern "C" int srgp_laplace_oat_scores(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m,
                                       const double *cand, int64_t n_cand, double sigma, const double *l, double tau,
                                       double delta, double pois_m, int maxit_nr, double tol_nr, const double *fmax,
                                       double *scores)
{
    if (!ctx || !xu || !cand || !l || !fmax || !scores || m <= 0 || n_cand <= 0 || maxit_nr < 1) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("srgp_laplace_oat_scores called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    const int d = ctx->d;
    const int64_t ma = m + 1, n = ctx->n;
    std::vector<double> ua((size_t)ma * d), ff((size_t)std::max<int64_t>(n, 1)), hist((size_t)maxit_nr + 1);
    for (int64_t t = 0; t < n_cand; t++) {
        for (int c = 0; c < d; c++) {
            for (int64_t k = 0; k < m; k++) ua[k + (size_t)ma * c] = xu[k + m * c];
            ua[m + (size_t)ma * c] = cand[t + n_cand * c];
        }
        std::copy(fmax, fmax + n, ff.begin());
        int n_nr = 0;
        // muu only enters u_mean (not requested here): R passes c(muu, muu[1])
        const int rc = srgp_laplace_newton(ctx, family, kernel, ua.data(), ma, nullptr, sigma, l, tau, delta, pois_m,
                                           maxit_nr, tol_nr, ff.data(), hist.data(), &n_nr, nullptr, nullptr, nullptr);
        if (rc == SRGP_ERR_NOT_PD) scores[t] = NAN;
        else if (rc != SRGP_OK) return rc;
        else scores[t] = hist[n_nr - 1];
    }
    return SRGP_OK;
}
