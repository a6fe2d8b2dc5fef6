## r/patches.R -- optional R-level delegation: same signatures as the reference, bodies call the fused GPU path.
## Source after library(sparseRGPs) (or drop into R/).  Nothing else in the package changes.

## R/vi_functions.R:14-27 -- signature unchanged
trace_term_fun <- function(cov_par, Sigma12, Sigma22, delta)
{
  .Call('_sparseRGPs_trace_term', cov_par$sigma, cov_par$tau, delta, Sigma12, Sigma22, PACKAGE = 'sparseRGPs')
}

## R/vi_functions.R:38-44, :54-60 are scalar arithmetic on the host and stay as they are:
##   dtrace_term_dtau(cov_par, trace_term)     = -2 * trace_term
##   dtrace_term_dcov_par(cov_par, A_trace)    = -(1/(2 * tau^2)) * sum(A_trace)

## Shared body: dcov_fun_dknot = NA -> fused objective + gradient; a function -> the same evaluation plus the
## knot-location gradient (R/vi_functions.R:425-592 / R/laplace_approx_gradient.R:965-1126), knot_bounds as in
## R/vi_functions.R:175-178.  The original R bodies (renamed *_R) remain the fall-back for argument combinations
## the GPU path does not take (dcov_fun_dtheta not a list).
.srgp_gauss_grad <- function(model, fallback, cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y,
                             ff, mu, transform, delta, ...)
{
  if(!is.list(dcov_fun_dtheta) || (!is.function(dcov_fun_dknot) && !transform))
    return(fallback(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff, mu, transform,
                    delta, ...))
  lnames <- if(cov_fun == "ard") paste("l", 1:ncol(xy), sep = "") else character()
  if(!is.function(dcov_fun_dknot))
  {
    res <- .Call('_sparseRGPs_gauss_obj_grad', model, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par, delta,
                 lnames, PACKAGE = 'sparseRGPs')
    grad <- res$gradient; names(grad) <- names(cov_par)
    return(list("gradient" = grad, "trans_par" = lapply(cov_par, log), "objective" = res$objective))
  }
  lo <- apply(X = xy, MARGIN = 2, FUN = min); hi <- apply(X = xy, MARGIN = 2, FUN = max)
  knot_bounds <- cbind(lo - (hi - lo)/10, hi + (hi - lo)/10)
  res <- .Call('_sparseRGPs_gauss_obj_grad_knots', model, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par,
               delta, lnames, knot_bounds, as.integer(knot_opt), transform, PACKAGE = 'sparseRGPs')
  grad <- res$gradient; names(grad) <- names(cov_par)
  list("gradient" = grad, "knot_gradient" = res$knot_gradient, "trans_par" = lapply(cov_par, log),
       "trans_knot" = res$trans_knot, "objective" = res$objective)
}

## R/vi_functions.R:126 -- signature unchanged
delbo_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
  .srgp_gauss_grad(0L, delbo_dcov_par_R, cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff,
                   mu, transform, delta, ...)

## R/laplace_approx_gradient.R:720 -- same for the FIC gradient (model 1)
dlogp_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
  .srgp_gauss_grad(1L, dlogp_dcov_par_R, cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff,
                   mu, transform, delta, ...)

## R/vi_functions.R:2211-2298 and R/knot_proposal_functions.R:1283-1353 -- inside knot_prop_random_norm_vi /
## knot_prop_random_norm, the `for(i in 1:nrow(pseudo_prop))` loop that rebuilds Sigma12 / Sigma22 and calls obj_fun
## once per candidate becomes one call; the sampling above it and the which.max below it stay as they are:
##
##   res <- .Call('_sparseRGPs_oat_scores', 0L,      # 1L in knot_prop_random_norm
##                norm_opt$cov_fun, norm_opt$xy, as.numeric(y), as.numeric(norm_opt$mu), xu, pseudo_prop,
##                norm_opt$cov_par, delta, lnames, PACKAGE = 'sparseRGPs')
##   bad <- which(is.nan(res$scores))                # the reference's try-error branch: resample with jitter
##   while(length(bad) > 0) {
##     pseudo_prop[bad,] <- xy_setminus_xu[sample.int(nrow(xy_setminus_xu), length(bad)),, drop = FALSE] +
##       rnorm(n = length(bad) * ncol(xu), mean = 0, sd = 1e-6)
##     res$scores[bad] <- .Call('_sparseRGPs_oat_scores', 0L, norm_opt$cov_fun, norm_opt$xy, as.numeric(y),
##                              as.numeric(norm_opt$mu), xu, pseudo_prop[bad,, drop = FALSE], norm_opt$cov_par, delta,
##                              lnames, PACKAGE = 'sparseRGPs')$scores
##     bad <- bad[is.nan(res$scores[bad])]
##   }
##   obj_fun_vals <- c(obj_fun_vals, res$scores)

## R/vi_functions.R:596 / R/laplace_gradient_ascent.R:1111 -- norm_grad_ascent_vi / norm_grad_ascent: the
## `while(iter < maxit && ...)` loops (:963-1158 / :1453-1633) become ONE call; argument handling above the loop and
## the u_mean / u_var tail (-> '_sparseRGPs_gauss_posterior_u') stay.  With `o` = the merged opt_master list:
##
##   fit <- .Call('_sparseRGPs_gauss_fit', 0L,        # 1L in norm_grad_ascent
##                cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par_start, o$delta, lnames,
##                o$optim_method, c(o$decay, o$epsilon, o$eta, o$learn_rate), as.integer(o$maxit), o$obj_tol, o$grad_tol,
##                is.list(dcov_fun_dtheta), is.function(dcov_fun_dknot), knot_bounds, as.integer(knot_opt),
##                PACKAGE = 'sparseRGPs')
##   ## fit: list(cov_par, xu, iter, obj_fun, grad, cov_par_history)
