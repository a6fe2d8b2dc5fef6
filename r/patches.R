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

## R/vi_functions.R:126 -- signature unchanged; fixed knots (dcov_fun_dknot = NA) go to the GPU, the knot-gradient
## branch keeps the original body (delbo_dcov_par_R is the renamed original).
delbo_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
{
  if(is.function(dcov_fun_dknot) || !transform || !is.list(dcov_fun_dtheta))
    return(delbo_dcov_par_R(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff, mu,
                            transform, delta, ...))
  lnames <- if(cov_fun == "ard") paste("l", 1:ncol(xy), sep = "") else character()
  res <- .Call('_sparseRGPs_gauss_obj_grad', 0L, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par, delta,
               lnames, PACKAGE = 'sparseRGPs')
  grad <- res$gradient; names(grad) <- names(cov_par)
  list("gradient" = grad, "trans_par" = lapply(cov_par, log), "objective" = res$objective)
}

## R/laplace_approx_gradient.R:720 -- same for the FIC gradient (model 1)
dlogp_dcov_par <- function(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot = NA, knot_opt, xu, xy, y, ff = NA,
                           mu, transform = TRUE, delta = 1e-6, ...)
{
  if(is.function(dcov_fun_dknot) || !transform || !is.list(dcov_fun_dtheta))
    return(dlogp_dcov_par_R(cov_par, cov_fun, dcov_fun_dtheta, dcov_fun_dknot, knot_opt, xu, xy, y, ff, mu,
                            transform, delta, ...))
  lnames <- if(cov_fun == "ard") paste("l", 1:ncol(xy), sep = "") else character()
  res <- .Call('_sparseRGPs_gauss_obj_grad', 1L, cov_fun, xy, as.numeric(y), as.numeric(mu), xu, cov_par, delta,
               lnames, PACKAGE = 'sparseRGPs')
  grad <- res$gradient; names(grad) <- names(cov_par)
  list("gradient" = grad, "trans_par" = lapply(cov_par, log), "objective" = res$objective)
}
