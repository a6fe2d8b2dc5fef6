## r/patches.R -- R-level delegation: same names and signatures as the reference, bodies call the GPU library.
## Source after library(sparseRGPs) (or drop into R/); nothing else in the package changes: optimize_gp, predict_gp,
## norm_grad_ascent_vi, norm_grad_ascent, laplace_grad_ascent and the knot proposal functions run UNMODIFIED on top.
## Every patched function keeps the reference's own body as <name>_R and falls back to it for argument combinations
## the GPU path does not take (user-supplied likelihood derivatives, a vector Poisson offset, full_cov = TRUE, ...).
## Checked end to end by tests/test_r_dropin.py (reference R sources + this file in the mini-R interpreter; `.Call`
## served by r/shim.c -> libsrgp.so on the GPU, or by the CPU oracle in the build container).

.srgp_keep <- function(name)
{
  ## the reference's own body, kept once under <name>_R (absent when this file is sourced without the package)
  if(exists(name) && !exists(paste(name, "_R", sep = "")))
    assign(paste(name, "_R", sep = ""), get(name), envir = globalenv())
}
for(.srgp_nm in c("trace_term_fun", "elbo_fun", "obj_fun_norm", "delbo_dcov_par", "dlogp_dcov_par",
                  "newtrap_sparseGP", "dlogq_dcov_par", "predict_vi", "predict_laplace")) .srgp_keep(.srgp_nm)

.srgp_lnames <- function(cov_fun, d) if(cov_fun == "ard") paste("l", 1:d, sep = "") else character()

## bounds of the knot transform, R/vi_functions.R:175-178
.srgp_knot_bounds <- function(xy)
{
  lo <- apply(X = xy, MARGIN = 2, FUN = min); hi <- apply(X = xy, MARGIN = 2, FUN = max)
  cbind(lo - (hi - lo)/10, hi + (hi - lo)/10)
}

## "bernoulli" / "poisson" when the likelihood derivative passed by the caller IS the package's own
## (R/derivative_functions_of_data_likelihoods.R:7-235), otherwise NA -> the R body runs
.srgp_family <- function(dlog_py_dff)
{
  if(exists("dlog_py_dff_bern") && identical(dlog_py_dff, dlog_py_dff_bern)) return("bernoulli")
  if(exists("dlog_py_dff_pois") && identical(dlog_py_dff, dlog_py_dff_pois)) return("poisson")
  NA
}

## the Poisson exposure `m` of `...`: a scalar (or constant vector) goes to the library, anything else to R
.srgp_pois_m <- function(args)
{
  if(is.null(args$m)) return(1)
  m <- as.numeric(args$m)
  if(length(m) == 1 || all(m == m[1])) return(m[1])
  NA
}

## ------------------------------------------------------------------------------------------------------------------
## R/vi_functions.R:14-27 -- signature unchanged
trace_term_fun <- function(cov_par, Sigma12, Sigma22, delta)
{
  .Call('_sparseRGPs_trace_term', cov_par$sigma, cov_par$tau, delta, Sigma12, Sigma22, PACKAGE = 'sparseRGPs')
}

## R/vi_functions.R:38-44, :54-60 are scalar arithmetic on the host and stay as they are:
##   dtrace_term_dtau(cov_par, trace_term)     = -2 * trace_term
##   dtrace_term_dcov_par(cov_par, A_trace)    = -(1/(2 * tau^2)) * sum(A_trace)

## R/laplace_approx_obj_funs.R:6-52 -- signature unchanged.  Sigma12 / Sigma22 / Z are what the caller built.
obj_fun_norm <- function(ff = NA, mu, Z, Sigma12, Sigma22, y, ...)
{
  .Call('_sparseRGPs_gauss_obj_mats', Sigma12, Sigma22, as.numeric(Z), as.numeric(y), as.numeric(mu),
        PACKAGE = 'sparseRGPs')
}

## R/vi_functions.R:64-121 -- signature unchanged (the default of trace_term_fun is resolved like the reference's
## callers do: they always pass it, R/vi_functions.R:755-762)
elbo_fun <- function(ff = NA, mu, Z, Sigma12, Sigma22, y, trace_term_fun = NULL, cov_par, ...)
{
  args <- list(...)
  ttf <- if(is.function(trace_term_fun)) trace_term_fun else get("trace_term_fun", envir = globalenv())
  .Call('_sparseRGPs_gauss_obj_mats', Sigma12, Sigma22, as.numeric(Z), as.numeric(y), as.numeric(mu),
        PACKAGE = 'sparseRGPs') +
    ttf(cov_par = cov_par, Sigma12 = Sigma12, Sigma22 = Sigma22, delta = args$delta)
}

## Shared body of the two Gaussian gradients: dcov_fun_dknot = NA -> fused objective + gradient; a function -> the same
## evaluation plus the knot-location gradient (R/vi_functions.R:425-592 / R/laplace_approx_gradient.R:965-1126).
## `fallback` is the NAME of the reference body (looked up only when it is needed).
.srgp_gauss_grad <- function(model, fallback, cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y,
                             ff, mu, transform, delta, ...)
{
  if(!is.list(dcov_fun_dtheta) || !transform)
    return(get(fallback)(cov_par = cov_par, cov_fun = cov_fun, dcov_fun_dtheta = dcov_fun_dtheta,
                    dcov_fun_dknot = dcov_fun_dknot, knot_opt = knot_opt, xu = xu, xy = xy, y = y, ff = ff, mu = mu,
                    transform = transform, delta = delta, ...))
  lnames <- .srgp_lnames(cov_fun, ncol(xy))
  if(!is.function(dcov_fun_dknot))
  {
    res <- .Call('_sparseRGPs_gauss_obj_grad', model, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par, delta,
                 lnames, PACKAGE = 'sparseRGPs')
    ## the library answers in its canonical order c(sigma, l.., tau) WITH names; the reference fills by name in
    ## cov_par's own order (R/vi_functions.R:163,259-261,416)
    return(list("gradient" = res$gradient[names(cov_par)], "trans_par" = lapply(cov_par, log),
                "objective" = res$objective))
  }
  res <- .Call('_sparseRGPs_gauss_obj_grad_knots', model, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par,
               delta, lnames, .srgp_knot_bounds(xy), as.integer(knot_opt), transform, PACKAGE = 'sparseRGPs')
  list("gradient" = res$gradient[names(cov_par)], "knot_gradient" = res$knot_gradient,
       "trans_par" = lapply(cov_par, log), "trans_knot" = res$trans_knot, "objective" = res$objective)
}

## R/vi_functions.R:126 -- signature unchanged
delbo_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
  .srgp_gauss_grad(0L, "delbo_dcov_par_R", cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff,
                   mu, transform, delta, ...)

## R/laplace_approx_gradient.R:720 -- same for the FIC gradient (model 1)
dlogp_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
  .srgp_gauss_grad(1L, "dlogp_dcov_par_R", cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff,
                   mu, transform, delta, ...)

## R/newtrap_sparseGP.R:6 -- signature unchanged.  The family is read off the likelihood derivative the caller passes.
newtrap_sparseGP <- function(start_vals, obj_fun, grad_loglik_fn, dlog_py_dff, d2log_py_dff, maxit = 1000,
                             tol = 1e-6, cov_par, cov_fun, xy, xu, y, mu, muu, delta = 1e-6, ...)
{
  family <- .srgp_family(dlog_py_dff)
  pois_m <- .srgp_pois_m(list(...))
  if(is.na(family) || is.na(pois_m))
    return(newtrap_sparseGP_R(start_vals = start_vals, obj_fun = obj_fun, grad_loglik_fn = grad_loglik_fn,
                              dlog_py_dff = dlog_py_dff, d2log_py_dff = d2log_py_dff, maxit = maxit, tol = tol,
                              cov_par = cov_par, cov_fun = cov_fun, xy = xy, xu = xu, y = y, mu = mu, muu = muu,
                              delta = delta, ...))
  .Call('_sparseRGPs_laplace_newton', family, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, as.numeric(muu),
        cov_par, delta, .srgp_lnames(cov_fun, ncol(xu)), as.numeric(start_vals), as.integer(maxit), tol, pois_m,
        PACKAGE = 'sparseRGPs')
}

## R/laplace_approx_gradient.R:25 -- signature unchanged
dlogq_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff,
                           dlog_py_dff, d2log_py_dff, d3log_py_dff, mu, transform = TRUE, delta = 1e-6, ...)
{
  family <- .srgp_family(dlog_py_dff)
  pois_m <- .srgp_pois_m(list(...))
  if(is.na(family) || is.na(pois_m) || !is.list(dcov_fun_dtheta) || !transform)
    return(dlogq_dcov_par_R(cov_par = cov_par, cov_fun = cov_fun, dcov_fun_dtheta = dcov_fun_dtheta,
                            dcov_fun_dknot = dcov_fun_dknot, knot_opt = knot_opt, xu = xu, xy = xy, y = y, ff = ff,
                            dlog_py_dff = dlog_py_dff, d2log_py_dff = d2log_py_dff, d3log_py_dff = d3log_py_dff,
                            mu = mu, transform = transform, delta = delta, ...))
  lnames <- .srgp_lnames(cov_fun, ncol(xy))
  if(!is.function(dcov_fun_dknot))
  {
    res <- .Call('_sparseRGPs_laplace_grad', family, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par, delta,
                 lnames, as.numeric(ff), pois_m, PACKAGE = 'sparseRGPs')
    return(list("gradient" = res$gradient[names(cov_par)], "trans_par" = lapply(cov_par, log)))
  }
  res <- .Call('_sparseRGPs_laplace_grad_knots', family, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par,
               delta, lnames, as.numeric(ff), pois_m, .srgp_knot_bounds(xy), as.integer(knot_opt), transform,
               PACKAGE = 'sparseRGPs')
  list("gradient" = res$gradient[names(cov_par)], "knot_gradient" = res$knot_gradient,
       "trans_par" = lapply(cov_par, log), "trans_knot" = res$trans_knot)
}

## Shared body of predict_vi / predict_laplace with full_cov = FALSE (R/vi_functions.R:1288-1324,
## R/laplace_approx_prediction.R:79-119): Sigma12 is never built.
.srgp_predict <- function(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, s22_nugget, var_const)
{
  res <- .Call('_sparseRGPs_predict', cov_fun, x_pred, as.numeric(mu), xu, as.numeric(muu), as.numeric(u_mean), u_var,
               cov_par, .srgp_lnames(cov_fun, ncol(xu)), s22_nugget, var_const, PACKAGE = 'sparseRGPs')
  list("pred_mean" = matrix(res$pred_mean, ncol = 1), "pred_var" = res$pred_var)
}

## R/vi_functions.R:1222 -- signature unchanged
predict_vi <- function(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, full_cov = FALSE, family = "gaussian",
                       delta = 1e-6)
{
  if(full_cov || family != "gaussian")
    return(predict_vi_R(u_mean = u_mean, u_var = u_var, xu = xu, x_pred = x_pred, cov_fun = cov_fun, cov_par = cov_par,
                        mu = mu, muu = muu, full_cov = full_cov, family = family, delta = delta))
  .srgp_predict(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, delta, cov_par$tau^2 + cov_par$sigma^2 + delta)
}

## R/laplace_approx_prediction.R:3 -- signature unchanged
predict_laplace <- function(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, full_cov = FALSE,
                            family = "gaussian", delta = 1e-6)
{
  if(full_cov)
    return(predict_laplace_R(u_mean = u_mean, u_var = u_var, xu = xu, x_pred = x_pred, cov_fun = cov_fun,
                             cov_par = cov_par, mu = mu, muu = muu, full_cov = full_cov, family = family, delta = delta))
  nugget <- if(family == "gaussian") delta else cov_par$tau^2 + delta
  .srgp_predict(u_mean, u_var, xu, x_pred, cov_fun, cov_par, mu, muu, nugget, cov_par$sigma^2 + cov_par$tau^2)
}

## ------------------------------------------------------------------------------------------------------------------
## Optional whole-loop replacements (one `.Call` instead of an R loop); the functions above already make the
## unmodified loops run on the GPU, these remove the per-iteration R round trip as well.
##
## R/vi_functions.R:2211-2298 and R/knot_proposal_functions.R:1283-1353 -- inside knot_prop_random_norm_vi /
## knot_prop_random_norm, the `for(i in 1:nrow(pseudo_prop))` loop that rebuilds Sigma12 / Sigma22 and calls obj_fun
## once per candidate becomes one call; the sampling above it and the which.max below it stay as they are:
##
##   res <- .Call('_sparseRGPs_oat_scores', 0L,      # 1L in knot_prop_random_norm
##                norm_opt$cov_fun, norm_opt$xy, as.numeric(y), as.numeric(norm_opt$mu), xu, pseudo_prop,
##                norm_opt$cov_par, delta, lnames, PACKAGE = 'sparseRGPs')
##   bad <- which(is.nan(res$scores))                # the reference's try-error branch: resample with jitter
##   obj_fun_vals <- c(obj_fun_vals, res$scores)
##
## knot_prop_random (R/knot_proposal_functions.R:1096-1120), sparse Laplace models:
##
##   res <- .Call('_sparseRGPs_laplace_oat_scores', family, laplace_opt$cov_fun, laplace_opt$xy, as.numeric(y),
##                as.numeric(laplace_opt$mu), xu, pseudo_prop, laplace_opt$cov_par, delta, lnames, laplace_opt$fmax,
##                pois_m, as.integer(maxit_nr), tol_nr, PACKAGE = 'sparseRGPs')
##
## R/vi_functions.R:596 / R/laplace_gradient_ascent.R:1111 -- norm_grad_ascent_vi / norm_grad_ascent: the
## `while(iter < maxit && ...)` loops (:963-1158 / :1453-1633) become ONE call, the u_mean / u_var tail another.
## With `o` = the merged opt_master list:
##
##   fit <- .Call('_sparseRGPs_gauss_fit', 0L,        # 1L in norm_grad_ascent
##                cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par_start, o$delta, lnames,
##                o$optim_method, c(o$decay, o$epsilon, o$eta, o$learn_rate), as.integer(o$maxit), o$obj_tol, o$grad_tol,
##                is.list(dcov_fun_dtheta), is.function(dcov_fun_dknot), knot_bounds, as.integer(knot_opt),
##                PACKAGE = 'sparseRGPs')
##   ## fit: list(cov_par, xu, iter, obj_fun, grad, cov_par_history)
##   post <- .Call('_sparseRGPs_gauss_posterior_u', 0L, cov_fun, xy, as.numeric(y), as.numeric(mu), fit$xu,
##                 as.numeric(muu), fit$cov_par, o$delta, lnames, PACKAGE = 'sparseRGPs')   # list(u_mean, u_var)
##
## R/laplace_gradient_ascent.R:10 -- laplace_grad_ascent, the same for the sparse Laplace models:
##
##   fit <- .Call('_sparseRGPs_laplace_fit', family, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, as.numeric(muu),
##                cov_par_start, o$delta, lnames, as.numeric(ff), pois_m, o$optim_method,
##                c(o$decay, o$epsilon, o$eta, o$learn_rate), as.integer(o$maxit), o$obj_tol, o$grad_tol,
##                as.integer(o$maxit_nr), o$tol_nr, is.list(dcov_fun_dtheta), is.function(dcov_fun_dknot), knot_bounds,
##                as.integer(knot_opt), PACKAGE = 'sparseRGPs')
##   ## fit: list(cov_par, xu, iter, obj_fun, grad, cov_par_history, nr_iter, fmax, u_mean, u_var)
