/* srgp_internal.h -- test hooks into building blocks of libsrgp.so.  NOT part of the drop-in ABI
   (include/srgp.h); used only by tests/ to check each device building block in isolation. */
#ifndef SRGP_INTERNAL_H
#define SRGP_INTERNAL_H
#include "../../include/srgp.h"
#ifdef __cplusplus
extern "C" {
#endif
/* C = alpha op(A) op(B) + beta C through the DMMA tile engine; host pointers; M,N % 128 == 0, K % 16 == 0. */
int srgp_test_gemm(srgp_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha, const double *A,
                   int lda, const double *B, int ldb, double beta, double *C, int ldc, int lower_only, int reps,
                   double *ms_out);
/* Cholesky + inverse + logdet of an m x m SPD matrix (host, column-major, ld = m). */
int srgp_test_chol_inverse(srgp_ctx *ctx, int m, const double *A, double *L_out, double *Ainv_out,
                           double *logdet_out, int *info_out, int reps, double *ms_out);
#ifdef __cplusplus
}
#endif
#endif
