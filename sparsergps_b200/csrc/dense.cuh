// dense.cuh -- replicated m x m stage (K6): DMMA GEMM, blocked Cholesky, triangular inverse, small BLAS-1/2.
// All matrices are column-major with leading dimension `ld` = mp = round_up(m, 128); the padding block is
// zero (identity on the diagonal where a factorisation needs it) so the tile engine never sees a tail.
#pragma once
#include "common.cuh"

namespace srgp {
namespace dense {

constexpr int NB = 128;   // Cholesky / inverse block = GEMM tile
constexpr int GEMV_SCRATCH = 16;   // gemv scratch is GEMV_SCRATCH * mp doubles

struct BatchDesc {
    int batch = 1;
    int64_t strideA = 0, strideB = 0, strideC = 0, strideCt = 0;
};

// k-range restriction for triangular operands (zero k-tiles are skipped)
enum { KMODE_FULL = 0,
       KMODE_A_UPPER = 1,    // op(A)[i, k] = 0 for k < i      -> k starts at the row tile
       KMODE_B_LOWER = 2,    // op(B)^T[n, k] = 0 for k > n    -> k ends after the column tile
       KMODE_AB_UPPER = 3 }; // both of the above kind (W W^T) -> k starts at max(row tile, column tile)

// C = alpha * op(A) * op(B) + beta * C,  op(A) is M x K, op(B) is K x N (BLAS semantics, column-major).
// M, N multiples of 128, K multiple of 16.  lower_only: skip tiles strictly above the block diagonal.
int gemm(srgp_ctx *ctx, cudaStream_t s, char transA, char transB, int M, int N, int K, double alpha,
         const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc,
         BatchDesc bd = BatchDesc(), bool lower_only = false, int kmode = KMODE_FULL, double *Ct = nullptr,
         int64_t ldct = 0);

// In-place blocked Cholesky A = L L^T of the leading mp x mp matrix (lower triangle referenced; on exit the
// lower triangle holds L, diagonal blocks have a zeroed strict upper part).  dinv (2*mp*128 + mp/128 doubles)
// receives the inverses of the diagonal blocks of L, their transposes, and the per-block log-det parts.  *info (device int) is set to 1 + column index on a non-positive
// pivot (R's chol() error); logdet (device double) receives 2 * sum_{i<m} log L_ii.
int potrf(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double *dinv, int *info, double *logdet);

// Linv = L^-1 (lower) and LinvT = L^-T (upper), both mp x mp and fully written, given L (lower, as left by
// potrf) and dinv.  tmp is mp x mp scratch.
int trtri(srgp_ctx *ctx, cudaStream_t s, const double *L, int mp, const double *dinv, double *Linv, double *LinvT,
          double *tmp);

// Ainv = L^-T L^-1 (full symmetric) from LinvT.
// y = alpha AT^T x + beta y0 given the transpose AT of the matrix (one launch; see dense.cu)
int gemv_t(srgp_ctx *ctx, cudaStream_t s, int mp, double alpha, const double *AT, const double *x, double beta,
           const double *y0, double *y);
int lauum(srgp_ctx *ctx, cudaStream_t s, const double *LinvT, int mp, double *Ainv);

// Convenience: A (mp x mp, destroyed -> L) ; Ainv, logdet as above.  Linv / LinvT / tmp are mp x mp scratch.
int chol_inverse(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double *dinv, double *Linv, double *LinvT,
                 double *tmp, double *Ainv, int *info, double *logdet);

// ---- small kernels ----------------------------------------------------------------------------------
// Put `v` on the padding diagonal (i >= m) and zero the rest of the padding rows / columns.
int pad_identity(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double v);
// C = a * A + b * B (elementwise, mp x mp); B may be null when b == 0.  add_diag is added for i == j < m.
int axpby(srgp_ctx *ctx, cudaStream_t s, int mp, int m, double a, const double *A, double b, const double *B,
          double add_diag, double *C);
// C += a * x y^T  (mp x mp, vectors of length mp)
int ger(srgp_ctx *ctx, cudaStream_t s, int mp, double a, const double *x, const double *y, double *C);
// y = alpha * A x + beta * y0 (A mp x mp symmetric or general column-major; y0 may be null)
int gemv(srgp_ctx *ctx, cudaStream_t s, int mp, double alpha, const double *A, const double *x, double beta,
         const double *y0, double *y, double *scratch /* 16 * mp doubles */);
// *out = sum_ij A_ij * B_ij over the leading m x m block   (deterministic two-stage reduction)
int dot_mm(srgp_ctx *ctx, cudaStream_t s, int mp, int m, const double *A, const double *B, double *out,
           double *scratch /* >= 1024 doubles */);
// *out = sum_i x_i * y_i, i < m
int dot_v(srgp_ctx *ctx, cudaStream_t s, int m, const double *x, const double *y, double *out);

}  // namespace dense
}  // namespace srgp
