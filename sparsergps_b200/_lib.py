"""ctypes binding of libsrgp.so (include/srgp.h).  Fails loudly: no CPU fallback, no oracle import."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsrgp.so")

dp = C.POINTER(C.c_double)
i64 = C.c_int64
ci = C.c_int
cd = C.c_double
vp = C.c_void_p

OK, ERR_ARG, ERR_UNKNOWN_KERNEL, ERR_UNKNOWN_PAR, ERR_CUDA, ERR_NOT_PD, ERR_STATE, ERR_COMM, ERR_NUMERIC = range(9)
OPT_ADADELTA, OPT_GA = 0, 1
SQEXP, EXP, ARD = 0, 1, 2
PAR_SIGMA, PAR_L, PAR_TAU, PAR_LC = 0, 1, 2, 3
VI, FIC = 0, 1
BERNOULLI, POISSON = 0, 1
UNIQUE_ID_BYTES = 128
PROF = {"assemble": 0, "gen": 1, "gram": 2, "km": 3, "dense": 4, "reduce": 5, "comm": 6}

KERNELS = {"sqexp": SQEXP, "exp": EXP, "ard": ARD}


class FitOpt(C.Structure):
    """srgp_fit_opt (include/srgp.h); defaults = the reference's opt_master (R/vi_functions.R:641-643)."""
    _fields_ = [("optim_method", ci), ("decay", cd), ("epsilon", cd), ("eta", cd), ("learn_rate", cd), ("maxit", ci),
                ("obj_tol", cd), ("grad_tol", cd), ("opt_theta", ci), ("opt_knots", ci)]


# name -> (restype, argtypes); must list every symbol include/srgp.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "srgp_version": (ci, []),
    "srgp_last_error": (C.c_char_p, []),
    "srgp_ctx_create": (ci, [ci, C.POINTER(vp)]),
    "srgp_ctx_destroy": (None, [vp]),
    "srgp_ctx_sync": (ci, [vp]),
    "srgp_make_cov_mat": (ci, [vp, ci, dp, i64, dp, i64, ci, cd, dp, cd, cd, dp]),
    "srgp_dsig_dtheta": (ci, [vp, ci, ci, ci, dp, i64, dp, i64, ci, cd, dp, cd, dp]),
    "srgp_make_cov_mat_dev": (ci, [vp, ci, vp, i64, vp, i64, ci, cd, dp, cd, cd, vp]),
    "srgp_dsig_dtheta_dev": (ci, [vp, ci, ci, ci, vp, i64, vp, i64, ci, cd, dp, cd, vp]),
    "srgp_real_to_pos": (None, [dp, i64, dp]),
    "srgp_pos_to_real": (None, [dp, i64, dp]),
    "srgp_real_to_bounded": (None, [dp, dp, dp, i64, dp]),
    "srgp_cov_fun_sqrd_exp": (cd, [dp, dp, ci, cd, cd]),
    "srgp_cov_fun_sqrd_exp_ard": (cd, [dp, dp, ci, cd, dp]),
    "srgp_cov_fun_exp": (cd, [dp, dp, ci, cd, cd]),
    "srgp_dsqexp_dsigma": (cd, [dp, dp, ci, cd, cd]),
    "srgp_dsqexp_dsigma_ard": (cd, [dp, dp, ci, cd, dp]),
    "srgp_dsqexp_dl": (cd, [dp, dp, ci, cd, cd]),
    "srgp_dsqexp_dl_ard": (cd, [dp, dp, ci, cd, dp, ci]),
    "srgp_dsqexp_dtau": (cd, [dp, dp, ci, cd]),
    "srgp_dexp_dsigma": (cd, [dp, dp, ci, cd, cd]),
    "srgp_dexp_dl": (cd, [dp, dp, ci, cd, cd]),
    "srgp_dexp_dtau": (cd, [dp, dp, ci, cd]),
    "srgp_dsqexp_dx2": (None, [dp, dp, ci, cd, cd, dp, dp, dp, dp]),
    "srgp_dsqexp_dx2_ard": (None, [dp, dp, ci, cd, dp, dp, dp, dp, dp]),
    "srgp_trace_term": (ci, [vp, cd, cd, cd, dp, i64, i64, dp, dp]),
    "srgp_dtrace_term_dcov_par": (ci, [vp, cd, dp, i64, dp]),
    "srgp_dtrace_term_dtau": (cd, [cd]),
    "srgp_omega_dk_reduce": (ci, [vp, ci, dp, i64, dp, i64, ci, cd, dp, cd, dp, dp]),
    "srgp_omega_dk_reduce_dev": (ci, [vp, ci, vp, i64, vp, i64, ci, cd, dp, cd, vp, dp]),
    "srgp_set_data": (ci, [vp, dp, i64, ci, dp, dp]),
    "srgp_set_data_dev": (ci, [vp, vp, i64, ci, vp, vp]),
    "srgp_gauss_obj_grad": (ci, [vp, ci, ci, dp, i64, cd, dp, cd, cd, dp, dp]),
    "srgp_gauss_obj_grad_host": (ci, [vp, ci, ci, dp, i64, ci, dp, dp, dp, i64, cd, dp, cd, cd, dp, dp]),
    "srgp_gauss_obj_grad_knots": (ci, [vp, ci, ci, dp, i64, cd, dp, cd, cd, dp, dp, C.POINTER(ci), i64, dp, dp, dp, dp]),
    "srgp_gauss_fit": (ci, [vp, ci, ci, dp, i64, dp, dp, dp, cd, C.POINTER(FitOpt), dp, dp, C.POINTER(ci), i64,
                            C.POINTER(ci), dp, dp, dp]),
    "srgp_oat_scores": (ci, [vp, ci, ci, dp, i64, dp, i64, cd, dp, cd, cd, dp, dp]),
    "srgp_gauss_posterior_u": (ci, [vp, ci, ci, dp, i64, dp, cd, dp, cd, cd, dp, dp]),
    "srgp_predict": (ci, [vp, ci, dp, i64, ci, dp, dp, i64, dp, dp, dp, cd, dp, cd, cd, dp, dp]),
    "srgp_laplace_newton": (ci, [vp, ci, ci, dp, i64, dp, cd, dp, cd, cd, cd, ci, cd, dp, dp, C.POINTER(ci), dp, dp, dp]),
    "srgp_laplace_grad": (ci, [vp, ci, ci, dp, i64, cd, dp, cd, cd, cd, dp, dp]),
    "srgp_laplace_grad_knots": (ci, [vp, ci, ci, dp, i64, cd, dp, cd, cd, cd, dp, dp, dp, C.POINTER(ci), i64, dp, dp, dp]),
    "srgp_laplace_fit": (ci, [vp, ci, ci, dp, i64, dp, dp, dp, dp, cd, cd, ci, cd, C.POINTER(FitOpt), dp, dp,
                              C.POINTER(ci), i64, dp, C.POINTER(ci), dp, dp, dp, C.POINTER(ci), dp, dp]),
    "srgp_laplace_oat_scores": (ci, [vp, ci, ci, dp, i64, dp, i64, cd, dp, cd, cd, cd, ci, cd, dp, dp]),
    "srgp_gauss_obj_mats": (ci, [vp, dp, i64, i64, dp, dp, i64, dp, dp, i64, dp]),
    "srgp_comm_unique_id": (ci, [C.c_char_p]),
    "srgp_comm_init": (ci, [vp, ci, ci, C.c_char_p]),
    "srgp_comm_destroy": (ci, [vp]),
    "srgp_dev_alloc": (ci, [vp, i64, C.POINTER(vp)]),
    "srgp_dev_free": (ci, [vp, vp]),
    "srgp_memcpy_h2d": (ci, [vp, vp, vp, i64]),
    "srgp_memcpy_d2h": (ci, [vp, vp, vp, i64]),
    "srgp_fill_normal_dev": (ci, [vp, vp, i64, C.c_uint64, cd, cd]),
    "srgp_timer_start": (ci, [vp]),
    "srgp_timer_stop_ms": (ci, [vp, dp]),
    "srgp_prof_enable": (ci, [vp, ci]),
    "srgp_prof_reset": (ci, [vp]),
    "srgp_prof_get": (ci, [vp, ci, C.POINTER(i64), dp]),
    "srgp_launch_count": (i64, [vp]),
    "srgp_flush_l2": (ci, [vp]),
    "srgp_probe_i8_peak": (ci, [vp, ci, dp, dp]),
    "srgp_i8_slices": (ci, []),
}


# test hooks (csrc/srgp_internal.h) -- not part of the drop-in ABI
INTERNAL_SIGNATURES = {
    "srgp_test_gemm": (ci, [vp, ci, ci, ci, ci, ci, cd, dp, ci, dp, ci, cd, dp, ci, ci, ci, dp]),
    "srgp_test_chol_inverse": (ci, [vp, ci, dp, dp, dp, dp, C.POINTER(ci), ci, dp]),
}


class SrgpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libsrgp status %d: %s" % (status, message))
        self.status = status


class NotPositiveDefinite(SrgpError):
    """An m x m Cholesky failed on the device (R's chol() error)."""


_lib = None


def load():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "sparsergps_b200/libsrgp.so is missing: run `python -m sparsergps_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for table in (SIGNATURES, INTERNAL_SIGNATURES):
            for name, (res, args) in table.items():
                f = getattr(lib, name)
                f.restype = res
                f.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != OK:
        msg = load().srgp_last_error().decode("utf-8", "replace")
        if status == ERR_NOT_PD:
            raise NotPositiveDefinite(status, msg)
        raise SrgpError(status, msg)


def fmat(a):
    """R matrix -> column-major float64 (copy only when needed)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def fvec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


def ptr(a):
    return a.ctypes.data_as(dp)
