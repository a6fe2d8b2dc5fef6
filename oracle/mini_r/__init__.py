"""mini_r: a small interpreter for the subset of R the reference's R/*.R files use -- TEST INFRASTRUCTURE.

Why: rows a14-a25 of SURVEY.md 8(a) (trace_term_fun, elbo_fun, delbo_dcov_par, obj_fun_norm, dlogp_dcov_par,
newtrap_sparseGP, obj_fun_bern/pois, dlogq_dcov_par, the likelihood derivatives) are plain R. No R interpreter
exists in this image, so the reference's model algebra could only be TRANSCRIBED (oracle/ref_model.py), and a
transcription can silently differ from its source. This package executes the reference's own R source files,
unmodified and where they lie (/root/reference/R/*.R), with the Rcpp exports bound to the reference's compiled C++
(oracle/ref_native.py) and LAPACK reached through NumPy where R reaches LAPACK. Its outputs pin oracle/ref_model.py
and are recorded as golden vectors (tests/golden/r_level.*) for the GPU tests.

What it is not: R. Only the language subset and base functions those files exercise are implemented (rparse.py,
interp.py, builtins.py); semantics follow the R Language Definition and the base help pages, and are checked on
small known-answer snippets in tests/test_mini_r.py.
"""
