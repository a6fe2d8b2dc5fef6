// dense.cu -- replicated m x m stage (K6): what R's chol / solve / det / %*% do on m x m matrices in
// R/vi_functions.R:96-118,227-246, R/laplace_approx_obj_funs.R:30-48, R/newtrap_sparseGP.R:244-289.
#include "dense.cuh"
#include "gemm.cuh"

namespace srgp {
namespace dense {

using namespace gemm;

// ------------------------------------------------------------------------------------------------
// DMMA GEMM
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
    const double *A, *B;
    double *C;
    int64_t lda, ldb, ldc, sA, sB, sC;
    int ktiles;
    double alpha, beta;
    int lower_only;
    double *Ct;          // optional: also store the transposed result, Ct[c + r*ldct]
    int64_t ldct, sCt;
    int kmode;           // zero-skipping for triangular operands, see dense.cuh
};

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 1) gemm_kernel(GemmArgs g)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tm = blockIdx.x, tn = blockIdx.y;
    if (g.lower_only && tn > tm) return;
    const double *A = g.A + blockIdx.z * g.sA + (A_KC ? (int64_t)tm * BM * g.lda : (int64_t)tm * BM);
    const double *B = g.B + blockIdx.z * g.sB + (B_KC ? (int64_t)tn * BN * g.ldb : (int64_t)tn * BN);
    // k-range: skip k-tiles where a triangular operand is known to be zero
    int kt0 = 0, kt1 = g.ktiles;
    if (g.kmode == KMODE_A_UPPER) kt0 = tm * (BM / BK);
    else if (g.kmode == KMODE_B_LOWER) kt1 = min(g.ktiles, (tn + 1) * (BN / BK));
    else if (g.kmode == KMODE_AB_UPPER) kt0 = max(tm, tn) * (BM / BK);
    A += A_KC ? (int64_t)kt0 * BK : (int64_t)kt0 * BK * g.lda;
    B += B_KC ? (int64_t)kt0 * BK : (int64_t)kt0 * BK * g.ldb;
    pipeline_init(sm);
    uint32_t it = 0;
    const int nk = max(0, kt1 - kt0);
    if (is_producer()) {
        reg_dec<PRODUCER_REGS>();
        if (is_producer_lead()) producer_issue<A_KC, B_KC, false>(sm, A, g.lda, B, g.ldb, nullptr, nk, it);
        return;
    }
    reg_inc<CONSUMER_REGS>();
    double acc[8][4][2];
    zero_acc(acc);
    consumer_mma<A_KC, B_KC, false>(sm, nk, it, acc);
    double *C = g.C + blockIdx.z * g.sC + (int64_t)tm * BM + (int64_t)tn * BN * g.ldc;
    double *Ct = g.Ct ? g.Ct + blockIdx.z * g.sCt + (int64_t)tn * BN + (int64_t)tm * BM * g.ldct : nullptr;
    // beta != 0: the old values of two fragment rows (16 independent loads) are fetched before anything is stored --
    // a load-update-store per element would serialise 64 global-memory latencies behind the stores the compiler
    // cannot prove disjoint (trailing update of the Cholesky: 48 -> 27 us per 128-wide step)
#pragma unroll
    for (int mp2 = 0; mp2 < 8; mp2 += 2) {
        double cold[2][4][2];
        if (g.beta != 0.0) {
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
                        cold[q][ni][e] = C[frag_row(mp2 + q) + (int64_t)(frag_col(ni) + e) * g.ldc];
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int mi = mp2 + q;
            const int r = frag_row(mi);
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int c = frag_col(ni);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double v = g.alpha * acc[mi][ni][e];
                    if (g.beta != 0.0) v += g.beta * cold[q][ni][e];
                    C[r + (int64_t)(c + e) * g.ldc] = v;
                    if (Ct) Ct[(c + e) + (int64_t)r * g.ldct] = v;
                }
            }
        }
    }
}

template <bool A_KC, bool B_KC>
static int launch_gemm(srgp_ctx *ctx, cudaStream_t s, dim3 grid, const GemmArgs &g)
{
    static DeviceOnce once;   // per template instantiation
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(gemm_kernel<A_KC, B_KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    gemm_kernel<A_KC, B_KC><<<grid, THREADS, sizeof(Smem), s>>>(g);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

int gemm(srgp_ctx *ctx, cudaStream_t s, char transA, char transB, int M, int N, int K, double alpha,
         const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc,
         BatchDesc bd, bool lower_only, int kmode, double *Ct, int64_t ldct)
{
    if (M % BM || N % BN || K % BK || M <= 0 || N <= 0 || K < 0) {
        set_error("gemm: extents %d x %d x %d are not tile multiples", M, N, K);
        return SRGP_ERR_ARG;
    }
    GemmArgs g{A, B, C, lda, ldb, ldc, bd.strideA, bd.strideB, bd.strideC, K / BK, alpha, beta, lower_only ? 1 : 0,
               Ct, ldct, bd.strideCt, kmode};
    dim3 grid(M / BM, N / BN, bd.batch);
    const bool akc = (transA == 'T'), bkc = (transB == 'N');
    if (!akc && !bkc) return launch_gemm<false, false>(ctx, s, grid, g);
    if (!akc && bkc) return launch_gemm<false, true>(ctx, s, grid, g);
    if (akc && !bkc) return launch_gemm<true, false>(ctx, s, grid, g);
    return launch_gemm<true, true>(ctx, s, grid, g);
}

// ------------------------------------------------------------------------------------------------
// The m x m products of the Cholesky / inverse on MANY SMs.  On sm_100 a DMMA CTA and a DFMA CTA run at the same FP64 rate
// (profiles/r01_microbench.json), so what a 128 x 128 x 128 tile costs is the 17 us ONE SM needs for its 4.2 MFLOP.  The
// trailing update of a factorisation step, the recursive-doubling products of the triangular inverse and W W^T therefore go
// in 64 x 64 tiles, 256 threads x (4 x 4) outputs, plain DFMA from shared memory: four times the CTAs of the 128 x 128 DMMA
// tiles, i.e. all 148 SMs instead of 16 .. 64 at m = 1024.
//   C (+)= alpha A B^T with A (M x K), B (N x K) column-major (MN-contiguous); beta in {0, 1}; lower_only skips the tiles
//   strictly above the diagonal; kmode skips the k range where a triangular operand is zero.  C must not alias A or B (the
//   in-place panel has its own kernel below).
// ------------------------------------------------------------------------------------------------
constexpr int ST = 64, SK = 16;      // tile, k-chunk

struct SmallNT {
    const double *A, *B;
    double *C, *Ct;              // Ct: optional second, transposed store (Ct[c + r ldct])
    int64_t lda, ldb, ldc, ldct, sA, sB, sC, sCt;   // leading dimensions, batch strides (blockIdx.z)
    int K;
    double alpha;
    int beta_one, lower_only, kmode;
};

__global__ void __launch_bounds__(256) small_nt_kernel(SmallNT g)
{
    const int tm = blockIdx.x, tn = blockIdx.y;
    if (g.lower_only && tn > tm) return;
    __shared__ double As[SK][ST + 4], Bs[SK][ST + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double *Ab = g.A + blockIdx.z * g.sA + (int64_t)tm * ST, *Bb = g.B + blockIdx.z * g.sB + (int64_t)tn * ST;
    // k range: skip where a triangular operand is known to be zero (dense.cuh)
    int k_lo = 0, k_hi = g.K;
    if (g.kmode == KMODE_A_UPPER) k_lo = tm * ST;
    else if (g.kmode == KMODE_B_LOWER) k_hi = min(g.K, (tn + 1) * ST);
    else if (g.kmode == KMODE_AB_UPPER) k_lo = max(tm, tn) * ST;
    double acc[4][4] = {};
    // 64 rows x 16 k of each operand: thread -> (row = threadIdx.x % 64, four k's).  The next k-chunk is fetched into
    // registers while the current one is multiplied (these kernels are short dependent launches: latency, not rate).
    const int r = threadIdx.x & 63, kq = threadIdx.x >> 6;
    double pa[4], pb[4];
    if (k_lo < k_hi) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            pa[q] = Ab[r + (int64_t)(k_lo + kq * 4 + q) * g.lda];
            pb[q] = Bb[r + (int64_t)(k_lo + kq * 4 + q) * g.ldb];
        }
    }
    for (int k0 = k_lo; k0 < k_hi; k0 += SK) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            As[kq * 4 + q][r] = pa[q];
            Bs[kq * 4 + q][r] = pb[q];
        }
        __syncthreads();
        if (k0 + SK < k_hi) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                pa[q] = Ab[r + (int64_t)(k0 + SK + kq * 4 + q) * g.lda];
                pb[q] = Bb[r + (int64_t)(k0 + SK + kq * 4 + q) * g.ldb];
            }
        }
#pragma unroll
        for (int k = 0; k < SK; k++) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) av[i] = As[k][tx + 16 * i], bv[i] = Bs[k][ty + 16 * i];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    double *Cb = g.C + blockIdx.z * g.sC + (int64_t)tm * ST + (int64_t)tn * ST * g.ldc;
    double *Ctb = g.Ct ? g.Ct + blockIdx.z * g.sCt + (int64_t)tn * ST + (int64_t)tm * ST * g.ldct : nullptr;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = tx + 16 * i, c = ty + 16 * j;
            double *p = Cb + r + (int64_t)c * g.ldc;
            const double v = g.beta_one ? fma(g.alpha, acc[i][j], *p) : g.alpha * acc[i][j];
            *p = v;
            if (Ctb) Ctb[c + (int64_t)r * g.ldct] = v;
        }
}

// C (+)= alpha A B^T in 64 x 64 DFMA tiles (M, N multiples of 64, K of 16)
int gemm_nt_small(srgp_ctx *ctx, cudaStream_t s, int M, int N, int K, double alpha, const double *A, int64_t lda,
                  const double *B, int64_t ldb, bool beta_one, double *C, int64_t ldc, BatchDesc bd, bool lower_only, int kmode,
                  double *Ct, int64_t ldct)
{
    SmallNT g{A, B, C, Ct, lda, ldb, ldc, ldct, bd.strideA, bd.strideB, bd.strideC, bd.strideCt, K, alpha,
              beta_one ? 1 : 0, lower_only ? 1 : 0, kmode};
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    small_nt_kernel<<<dim3(M / ST, N / ST, bd.batch), 256, 0, s>>>(g);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

// The panel L21 = A21 X^T (X = L11^-1, 128 x 128, ld = 128) IN PLACE: a CTA owns 32 rows of A21 with all 128 columns, reads
// them completely, then stores -- no other CTA touches those rows.
__global__ void __launch_bounds__(256)
panel_kernel(double *A, int64_t lda, const double *__restrict__ X)
{
    __shared__ double As[SK][32 + 4], Xs[SK][NB + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double *Ab = A + (int64_t)blockIdx.x * 32;
    double acc[2][8] = {};
    const int r = threadIdx.x & 31, kq = threadIdx.x >> 5;          // A: 32 rows x 16 k, two k's per thread
    const int j = threadIdx.x & 127, kh = threadIdx.x >> 7;         // X: 128 columns x 16 k, eight k's per thread
    double pa[2], px[8];                                            // the next k-chunk, fetched while the current one is used
    pa[0] = Ab[r + (int64_t)(kq * 2) * lda];
    pa[1] = Ab[r + (int64_t)(kq * 2 + 1) * lda];
#pragma unroll
    for (int q = 0; q < 8; q++) px[q] = X[j + (kh * 8 + q) * NB];
    for (int k0 = 0; k0 < NB; k0 += SK) {
        As[kq * 2][r] = pa[0];
        As[kq * 2 + 1][r] = pa[1];
#pragma unroll
        for (int q = 0; q < 8; q++) Xs[kh * 8 + q][j] = px[q];
        __syncthreads();
        if (k0 + SK < NB) {
            pa[0] = Ab[r + (int64_t)(k0 + SK + kq * 2) * lda];
            pa[1] = Ab[r + (int64_t)(k0 + SK + kq * 2 + 1) * lda];
#pragma unroll
            for (int q = 0; q < 8; q++) px[q] = X[j + (k0 + SK + kh * 8 + q) * NB];
        }
#pragma unroll
        for (int k = 0; k < SK; k++) {
            const double a0 = As[k][tx], a1 = As[k][tx + 16];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const double b = Xs[k][ty + 16 * jj];
                acc[0][jj] = fma(a0, b, acc[0][jj]);
                acc[1][jj] = fma(a1, b, acc[1][jj]);
            }
        }
        __syncthreads();                                                    // also: every read of A precedes the stores below
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        Ab[tx + (int64_t)(ty + 16 * j) * lda] = acc[0][j];
        Ab[tx + 16 + (int64_t)(ty + 16 * j) * lda] = acc[1][j];
    }
}

// ------------------------------------------------------------------------------------------------
// Cholesky: diagonal block factor + inverse in shared memory (one CTA), panel / trailing updates in 64 x 64 DFMA tiles
// ------------------------------------------------------------------------------------------------
constexpr int DLD = NB + 1;   // 129: odd stride, column-major block in shared memory

// Factor one 128 x 128 diagonal block and invert its Cholesky factor, in ONE CTA of 256 threads.
// Thread (ti, tk) = (tid & 15, tid >> 4) keeps the 8 x 8 sub-block {(ti + 16 ii, tk + 16 kk)} in registers;
// per column j only the scaled column (Cholesky) or the scaled row of the running inverse travels through a
// double-buffered shared vector, followed by 64 register FMAs -- one __syncthreads per column.  The column /
// block loops are split as j = 16 jb + jt with jb unrolled so every register index is static.
//   outputs: L (lower, strict upper zeroed) back into A; X = L^-1 into dinv (lower) and X^T into dinvT (upper).
__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double *__restrict__ A, int64_t ld, int col0, int m, double *__restrict__ dinv,
                  double *__restrict__ dinvT, int *__restrict__ info, double *__restrict__ logdet_part)
{
    extern __shared__ double sh[];
    double *Ls = sh;                   // NB x DLD: the factor, for the inverse phase
    double *vec = Ls + NB * DLD;       // 2 x NB double-buffered broadcast vector
    double *diag = vec + 2 * NB;       // NB pivots
    double *rdiag = diag + NB;         // NB reciprocal pivots (no division in either column loop)
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int ti = tid & 15, tk = tid >> 4;
    const int lane = tid & 31;
    double a[8][8];
#pragma unroll
    for (int ii = 0; ii < 8; ii++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const int i = ti + 16 * ii, k = tk + 16 * kk;
            a[ii][kk] = (i >= k) ? A[i + (int64_t)k * ld] : 0.0;
        }
    // ---- Cholesky (right-looking) ----
#pragma unroll
    for (int jb = 0; jb < 8; jb++) {
        for (int jt = 0; jt < 16; jt++) {
            const int j = jb * 16 + jt;
            double *cv = vec + (j & 1) * NB;
            // the 16 owners of column j sit in one half-warp (tk == jt); the pivot owner is its lane ti == jt
            const double piv = __shfl_sync(0xffffffffu, a[jb][jb], (lane & 16) + jt);
            if (tk == jt) {
                // 1/sqrt and one multiply instead of sqrt followed by a division: the pivot computation is the
                // serial part of every column step (rsqrt is within 1 ulp; L_jj = piv / sqrt(piv) within 2)
                const double rj = rsqrt(piv);
                const double dj = piv * rj;
                if (ti == jt) {
                    if (!(piv > 0.0) && *info == 0) *info = col0 + j + 1;
                    diag[j] = dj;
                    rdiag[j] = rj;
                }
#pragma unroll
                for (int ii = 0; ii < 8; ii++) {
                    const int i = ti + 16 * ii;
                    double v = 0.0;
                    if (ii > jb || (ii == jb && ti > jt)) v = a[ii][jb] * rj;
                    else if (ii == jb && ti == jt) v = dj;
                    if (ii >= jb) a[ii][jb] = v;
                    cv[i] = (i > j) ? v : 0.0;
                }
            }
            __syncthreads();
            double li[8], lk[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                li[q] = cv[ti + 16 * q];
                lk[q] = cv[tk + 16 * q];
            }
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
                if (kk < jb) continue;                           // columns of finished blocks
                const bool kact = (kk > jb) || (tk > jt);        // k > j
#pragma unroll
                for (int ii = 0; ii < 8; ii++) {
                    if (ii < kk) continue;                       // strictly upper sub-blocks are never read
                    if (kact) a[ii][kk] = fma(-li[ii], lk[kk], a[ii][kk]);
                }
            }
        }
    }
    __syncthreads();
    // ---- log-determinant part, L to global and to shared memory ----
    {
        double v = 0.0;
        if (tid < NB && col0 + tid < m) v = log(diag[tid]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[tid >> 5] = v;
    }
#pragma unroll
    for (int ii = 0; ii < 8; ii++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const int i = ti + 16 * ii, k = tk + 16 * kk;
            const double v = (i >= k) ? a[ii][kk] : 0.0;
            A[i + (int64_t)k * ld] = v;
            Ls[i + k * DLD] = v;
        }
    __syncthreads();
    if (tid == 0) *logdet_part = 2.0 * (red[0] + red[1] + red[2] + red[3]);
    // ---- X = L^-1 by forward substitution on the identity held in registers ----
    //   row j of X: x_j = x_j / L_jj ;  rows i > j: x_i -= L_ij x_j
#pragma unroll
    for (int ii = 0; ii < 8; ii++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) a[ii][kk] = (ii == kk && ti == tk) ? 1.0 : 0.0;
#pragma unroll
    for (int jb = 0; jb < 8; jb++) {
        for (int jt = 0; jt < 16; jt++) {
            const int j = jb * 16 + jt;
            double *rv = vec + (j & 1) * NB;
            if (ti == jt) {                                      // owners of row j (16 threads, one per tk)
                const double inv = rdiag[j];
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    double v = 0.0;
                    if (kk <= jb) {
                        v = a[jb][kk] * inv;
                        a[jb][kk] = v;
                    }
                    rv[tk + 16 * kk] = v;
                }
            }
            __syncthreads();
            double li[8], rk[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                li[q] = Ls[ti + 16 * q + j * DLD];
                rk[q] = rv[tk + 16 * q];
            }
#pragma unroll
            for (int ii = 0; ii < 8; ii++) {
                if (ii < jb) continue;
                const bool iact = (ii > jb) || (ti > jt);        // i > j
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    if (kk > jb) continue;                       // X is lower triangular: columns k <= j only
                    if (iact) a[ii][kk] = fma(-li[ii], rk[kk], a[ii][kk]);
                }
            }
        }
    }
#pragma unroll
    for (int ii = 0; ii < 8; ii++)
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const int i = ti + 16 * ii, k = tk + 16 * kk;
            const double v = (i >= k) ? a[ii][kk] : 0.0;
            dinv[i + k * NB] = v;
            dinvT[k + i * NB] = v;
        }
}

__global__ void sum_parts_kernel(const double *parts, int n, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += parts[i];
        *out = s;
    }
}

int potrf(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double *dinv, int *info, double *logdet)
{
    static DeviceOnce once;
    const size_t smem = sizeof(double) * (NB * DLD + 4 * NB);   // factor + 2 broadcast vectors + pivots + reciprocals
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nb = mp / NB;
    // dinv area layout (caller allocates 2*mp*NB + nb doubles): [X_kk blocks | X_kk^T blocks | logdet partials]
    double *dinvT = dinv + (size_t)mp * NB;
    double *parts = dinv + (size_t)2 * mp * NB;
    for (int kb = 0; kb < nb; kb++) {
        double *Akk = A + (size_t)kb * NB * ((size_t)mp + 1);
        double *Dk = dinv + (size_t)kb * NB * NB;
        {
            KernelScope ks(ctx, SRGP_PROF_DENSE, s);
            potrf_diag_kernel<<<1, 256, smem, s>>>(Akk, mp, kb * NB, m, Dk, dinvT + (size_t)kb * NB * NB, info, parts + kb);
            SRGP_LAUNCH_CHECK();
        }
        const int Mr = mp - (kb + 1) * NB;
        if (Mr > 0) {
            double *A21 = Akk + NB;
            double *A22 = Akk + NB * ((size_t)mp + 1);
            // L21 = A21 * L11^-T  (in place: a CTA reads its rows completely before writing them)
            {
                KernelScope ks(ctx, SRGP_PROF_DENSE, s);
                panel_kernel<<<Mr / 32, 256, 0, s>>>(A21, mp, Dk);
                SRGP_LAUNCH_CHECK();
            }
            // A22 -= L21 L21^T on the lower block triangle
            SRGP_TRY(gemm_nt_small(ctx, s, Mr, Mr, NB, -1.0, A21, mp, A21, mp, true, A22, mp, BatchDesc(), true, KMODE_FULL, nullptr, 0));
        }
    }
    {
        KernelScope ks(ctx, SRGP_PROF_DENSE, s);
        sum_parts_kernel<<<1, 32, 0, s>>>(parts, nb, logdet);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

__global__ void scatter_diag_blocks_kernel(const double *__restrict__ dinv, double *__restrict__ M, int mp)
{
    // one CTA per diagonal block
    const int kb = blockIdx.x;
    const double *src = dinv + (size_t)kb * NB * NB;
    double *dst = M + (size_t)kb * NB * ((size_t)mp + 1);
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int i = idx % NB, j = idx / NB;
        dst[i + (size_t)j * mp] = src[i + j * NB];
    }
}

// X = L^-1 (lower) and W = X^T (upper) by recursive doubling over the diagonal:
//   W12 = -(W11 L21^T) X22^T ,  X21 = W12^T
// Both products are "NT" GEMMs on MN-contiguous operands (the tile engine's fast path); the second one stores
// its result a second time transposed, which is how X21 is obtained.  Zero k-tiles of the triangular
// operands are skipped.
int trtri(srgp_ctx *ctx, cudaStream_t s, const double *L, int mp, const double *dinv, double *Linv, double *LinvT,
          double *tmp)
{
    const int nb = mp / NB;
    const double *dinvT = dinv + (size_t)mp * NB;
    SRGP_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * (size_t)mp * mp, s));
    SRGP_CUDA(cudaMemsetAsync(LinvT, 0, sizeof(double) * (size_t)mp * mp, s));
    {
        KernelScope ks(ctx, SRGP_PROF_DENSE, s, 2);
        scatter_diag_blocks_kernel<<<nb, 256, 0, s>>>(dinv, Linv, mp);
        SRGP_LAUNCH_CHECK();
        scatter_diag_blocks_kernel<<<nb, 256, 0, s>>>(dinvT, LinvT, mp);
        SRGP_LAUNCH_CHECK();
    }
    const int64_t ld = mp;
    auto at = [&](const double *base, int rb, int cb) {
        return const_cast<double *>(base) + (size_t)rb * NB + (size_t)cb * NB * ld;
    };
    auto pair = [&](int b0, int sl, int sr, int batch) -> int {
        // left block [b0, b0+sl), right block [b0+sl, b0+sl+sr) (in units of NB); `batch` pairs, stride 2*sl
        BatchDesc bd;
        bd.batch = batch;
        bd.strideA = bd.strideB = bd.strideC = bd.strideCt = (int64_t)2 * sl * NB * (ld + 1);
        const int S = sl * NB, Mr = sr * NB;
        // T' = W11 * L21^T   (S x Mr, at tmp[b0, b0+sl])
        SRGP_TRY(gemm_nt_small(ctx, s, S, Mr, S, 1.0, at(LinvT, b0, b0), ld, at(L, b0 + sl, b0), ld, false,
                               at(tmp, b0, b0 + sl), ld, bd, false, KMODE_A_UPPER, nullptr, 0));
        // W12 = -T' * X22^T  (S x Mr, at LinvT[b0, b0+sl]);  X21 = W12^T at Linv[b0+sl, b0]
        SRGP_TRY(gemm_nt_small(ctx, s, S, Mr, Mr, -1.0, at(tmp, b0, b0 + sl), ld, at(Linv, b0 + sl, b0 + sl), ld, false,
                               at(LinvT, b0, b0 + sl), ld, bd, false, KMODE_B_LOWER, at(Linv, b0 + sl, b0), ld));
        return SRGP_OK;
    };
    for (int sblk = 1; sblk < nb; sblk *= 2) {
        const int full = nb / (2 * sblk);
        if (full > 0) SRGP_TRY(pair(0, sblk, sblk, full));
        const int b0 = full * 2 * sblk, rem = nb - b0;
        if (rem > sblk) SRGP_TRY(pair(b0, sblk, rem - sblk, 1));
    }
    return SRGP_OK;
}

// Ainv = X^T X = W W^T (full symmetric); k >= max(i, j) because W is upper triangular.
int lauum(srgp_ctx *ctx, cudaStream_t s, const double *LinvT, int mp, double *Ainv)
{
    // the result is symmetric: the lower block triangle is computed (136 of 256 tiles at m = 1024: one wave instead of two)
    // and every tile is stored a second time transposed (on the diagonal tiles both stores carry the same bits)
    return gemm_nt_small(ctx, s, mp, mp, mp, 1.0, LinvT, mp, LinvT, mp, false, Ainv, mp, BatchDesc(), true, KMODE_AB_UPPER,
                         Ainv, mp);
}

int chol_inverse(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double *dinv, double *Linv, double *LinvT,
                 double *tmp, double *Ainv, int *info, double *logdet)
{
    SRGP_TRY(potrf(ctx, s, A, mp, m, dinv, info, logdet));
    SRGP_TRY(trtri(ctx, s, A, mp, dinv, Linv, LinvT, tmp));
    return lauum(ctx, s, LinvT, mp, Ainv);
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void pad_identity_kernel(double *A, int mp, int m, double v)
{
    const int64_t total = (int64_t)mp * mp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % mp), j = (int)(idx / mp);
        if (i >= m || j >= m) A[idx] = (i == j) ? v : 0.0;
    }
}

int pad_identity(srgp_ctx *ctx, cudaStream_t s, double *A, int mp, int m, double v)
{
    if (mp == m) return SRGP_OK;
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    pad_identity_kernel<<<ctx->sm_count, 256, 0, s>>>(A, mp, m, v);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

__global__ void axpby_kernel(int mp, int m, double a, const double *__restrict__ A, double b,
                             const double *__restrict__ B, double add_diag, double *__restrict__ C)
{
    const int64_t total = (int64_t)mp * mp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % mp), j = (int)(idx / mp);
        double v = a * A[idx];
        if (b != 0.0) v += b * B[idx];
        if (i == j && i < m) v += add_diag;
        C[idx] = v;
    }
}

int axpby(srgp_ctx *ctx, cudaStream_t s, int mp, int m, double a, const double *A, double b, const double *B,
          double add_diag, double *C)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    axpby_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(mp, m, a, A, b, B, add_diag, C);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

__global__ void ger_kernel(int mp, double a, const double *__restrict__ x, const double *__restrict__ y,
                           double *__restrict__ C)
{
    const int64_t total = (int64_t)mp * mp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % mp), j = (int)(idx / mp);
        C[idx] += a * x[i] * y[j];
    }
}

int ger(srgp_ctx *ctx, cudaStream_t s, int mp, double a, const double *x, const double *y, double *C)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    ger_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(mp, a, x, y, C);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

constexpr int GEMV_SPLIT = GEMV_SCRATCH;

__global__ void gemv_partial_kernel(int mp, const double *__restrict__ A, const double *__restrict__ x,
                                    double *__restrict__ part)
{
    const int row = blockIdx.x * 128 + threadIdx.x;
    const int cols = mp / GEMV_SPLIT, c0 = blockIdx.y * cols;
    double acc = 0.0;
    for (int c = c0; c < c0 + cols; c++) acc = fma(A[row + (int64_t)c * mp], x[c], acc);
    part[blockIdx.y * mp + row] = acc;
}

__global__ void gemv_finish_kernel(int mp, double alpha, const double *__restrict__ part, double beta,
                                   const double *__restrict__ y0, double *__restrict__ y)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= mp) return;
    double acc = 0.0;
    for (int sidx = 0; sidx < GEMV_SPLIT; sidx++) acc += part[sidx * mp + row];
    double v = alpha * acc;
    if (beta != 0.0) v += beta * y0[row];
    y[row] = v;
}

int gemv(srgp_ctx *ctx, cudaStream_t s, int mp, double alpha, const double *A, const double *x, double beta,
         const double *y0, double *y, double *scratch)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s, 2);
    gemv_partial_kernel<<<dim3(mp / 128, GEMV_SPLIT), 128, 0, s>>>(mp, A, x, scratch);
    SRGP_LAUNCH_CHECK();
    gemv_finish_kernel<<<ceil_div(mp, 256), 256, 0, s>>>(mp, alpha, scratch, beta, y0, y);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = alpha AT^T x + beta y0 from the TRANSPOSE of the matrix (column-major, ld = mp): output row i is the dot product of
// column i of AT with x, one warp per row, coalesced -- one launch instead of the split-k pair of gemv().  The chains use it
// wherever the transpose is at hand (L^-1 and its transposed copy from trtri, symmetric Grams and inverses).
__global__ void __launch_bounds__(256) gemv_t_kernel(int mp, double alpha, const double *__restrict__ AT,
                                                     const double *__restrict__ x, double beta,
                                                     const double *__restrict__ y0, double *__restrict__ y)
{
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const double *col = AT + (int64_t)row * mp;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int k = lane; k < mp; k += 128) {          // mp is a multiple of 128
        a0 = fma(col[k], x[k], a0);
        a1 = fma(col[k + 32], x[k + 32], a1);
        a2 = fma(col[k + 64], x[k + 64], a2);
        a3 = fma(col[k + 96], x[k + 96], a3);
    }
    const double acc = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) {
        double v = alpha * acc;
        if (beta != 0.0) v += beta * y0[row];
        y[row] = v;
    }
}

int gemv_t(srgp_ctx *ctx, cudaStream_t s, int mp, double alpha, const double *AT, const double *x, double beta,
           const double *y0, double *y)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    gemv_t_kernel<<<mp / 8, 256, 0, s>>>(mp, alpha, AT, x, beta, y0, y);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

constexpr int DOT_BLOCKS = 128;

__global__ void dot_mm_partial_kernel(int mp, int m, const double *__restrict__ A, const double *__restrict__ B,
                                      double *__restrict__ part)
{
    double acc = 0.0;
    const int64_t total = (int64_t)mp * m;   // columns j < m
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % mp);
        if (i < m) acc = fma(A[idx], B[idx], acc);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void sum_small_kernel(const double *__restrict__ part, int n, double *__restrict__ out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) *out = acc;
}

int dot_mm(srgp_ctx *ctx, cudaStream_t s, int mp, int m, const double *A, const double *B, double *out,
           double *scratch)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s, 2);
    dot_mm_partial_kernel<<<DOT_BLOCKS, 256, 0, s>>>(mp, m, A, B, scratch);
    SRGP_LAUNCH_CHECK();
    sum_small_kernel<<<1, 128, 0, s>>>(scratch, DOT_BLOCKS, out);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

__global__ void dot_v_kernel(int m, const double *__restrict__ x, const double *__restrict__ y,
                             double *__restrict__ out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) acc = fma(x[i], y[i], acc);
    acc = block_sum(acc);
    if (threadIdx.x == 0) *out = acc;
}

int dot_v(srgp_ctx *ctx, cudaStream_t s, int m, const double *x, const double *y, double *out)
{
    KernelScope ks(ctx, SRGP_PROF_DENSE, s);
    dot_v_kernel<<<1, 256, 0, s>>>(m, x, y, out);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

}  // namespace dense
}  // namespace srgp

// ------------------------------------------------------------------------------------------------
// test hooks (declared in srgp_internal.h; not part of the drop-in ABI)
// ------------------------------------------------------------------------------------------------
#include "srgp_internal.h"
using namespace srgp;

extern "C" int srgp_test_gemm(srgp_ctx *ctx, int transA, int transB, int M, int N, int K, double alpha,
                              const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc,
                              int lower_only, int reps, double *ms_out)
{
    SRGP_TRY(use_device(ctx));
    const size_t a_elems = (size_t)lda * (transA ? M : K), b_elems = (size_t)ldb * (transB ? K : N);
    const size_t c_elems = (size_t)ldc * N;
    DevBuf dA, dB, dC;
    DevBufScope scope;
    scope.own(dA), scope.own(dB), scope.own(dC);
    SRGP_TRY(dA.reserve(a_elems * 8));
    SRGP_TRY(dB.reserve(b_elems * 8));
    SRGP_TRY(dC.reserve(c_elems * 8));
    // stream-ordered copies: ctx->stream is non-blocking, so legacy-stream cudaMemcpy would not order with it
    SRGP_CUDA(cudaMemcpyAsync(dA.p, A, a_elems * 8, cudaMemcpyHostToDevice, ctx->stream));
    SRGP_CUDA(cudaMemcpyAsync(dB.p, B, b_elems * 8, cudaMemcpyHostToDevice, ctx->stream));
    SRGP_CUDA(cudaMemcpyAsync(dC.p, C, c_elems * 8, cudaMemcpyHostToDevice, ctx->stream));
    int st = dense::gemm(ctx, ctx->stream, transA ? 'T' : 'N', transB ? 'T' : 'N', M, N, K, alpha, dA.d(), lda,
                         dB.d(), ldb, beta, dC.d(), ldc, dense::BatchDesc(), lower_only != 0);
    if (st == SRGP_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("gemm failed: %s", cudaGetErrorString(e));
            st = SRGP_ERR_CUDA;
        }
    }
    if (st == SRGP_OK) {
        SRGP_CUDA(cudaMemcpyAsync(C, dC.p, c_elems * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (st == SRGP_OK && reps > 0 && ms_out) {
        // timing run (result discarded): beta = 0 so repeated launches are idempotent
        SRGP_CUDA(cudaEventRecord(ctx->tim0, ctx->stream));
        for (int r = 0; r < reps; r++)
            dense::gemm(ctx, ctx->stream, transA ? 'T' : 'N', transB ? 'T' : 'N', M, N, K, alpha, dA.d(), lda, dB.d(),
                        ldb, 0.0, dC.d(), ldc, dense::BatchDesc(), lower_only != 0);
        SRGP_CUDA(cudaEventRecord(ctx->tim1, ctx->stream));
        SRGP_CUDA(cudaEventSynchronize(ctx->tim1));
        float f;
        SRGP_CUDA(cudaEventElapsedTime(&f, ctx->tim0, ctx->tim1));
        *ms_out = f / reps;
    }
    return st;
}

extern "C" int srgp_test_chol_inverse(srgp_ctx *ctx, int m, const double *A, double *L_out, double *Ainv_out,
                                      double *logdet_out, int *info_out, int reps, double *ms_out)
{
    SRGP_TRY(use_device(ctx));
    const int mp = (int)round_up(m, dense::NB);
    const size_t mm = (size_t)mp * mp;
    DevBuf dA, dA0, dDinv, dLinv, dLinvT, dTmp, dAinv, dMisc;
    DevBufScope scope;
    for (DevBuf *b : {&dA, &dA0, &dDinv, &dLinv, &dLinvT, &dTmp, &dAinv, &dMisc}) scope.own(*b);
    SRGP_TRY(dA.reserve(mm * 8));
    SRGP_TRY(dA0.reserve(mm * 8));
    SRGP_TRY(dDinv.reserve(((size_t)2 * mp * dense::NB + mp / dense::NB) * 8));
    SRGP_TRY(dLinvT.reserve(mm * 8));
    SRGP_TRY(dLinv.reserve(mm * 8));
    SRGP_TRY(dTmp.reserve(mm * 8));
    SRGP_TRY(dAinv.reserve(mm * 8));
    SRGP_TRY(dMisc.reserve(64));
    SRGP_CUDA(cudaMemsetAsync(dA0.p, 0, mm * 8, ctx->stream));
    SRGP_CUDA(cudaMemcpy2DAsync(dA0.p, (size_t)mp * 8, A, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice,
                                ctx->stream));
    int *info = reinterpret_cast<int *>(dMisc.d() + 1);
    double *logdet = dMisc.d();
    int st = SRGP_OK;
    const int total = 1 + (reps > 0 ? reps : 0);
    for (int r = 0; r < total && st == SRGP_OK; r++) {
        if (r == 1) SRGP_CUDA(cudaEventRecord(ctx->tim0, ctx->stream));
        SRGP_CUDA(cudaMemcpyAsync(dA.p, dA0.p, mm * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        SRGP_CUDA(cudaMemsetAsync(dMisc.p, 0, 64, ctx->stream));
        st = dense::pad_identity(ctx, ctx->stream, dA.d(), mp, m, 1.0);
        if (st == SRGP_OK)
            st = dense::chol_inverse(ctx, ctx->stream, dA.d(), mp, m, dDinv.d(), dLinv.d(), dLinvT.d(), dTmp.d(),
                                     dAinv.d(), info, logdet);
    }
    if (st == SRGP_OK && reps > 0) {
        SRGP_CUDA(cudaEventRecord(ctx->tim1, ctx->stream));
        SRGP_CUDA(cudaEventSynchronize(ctx->tim1));
        float f = 0.f;
        SRGP_CUDA(cudaEventElapsedTime(&f, ctx->tim0, ctx->tim1));
        if (ms_out) *ms_out = f / reps;
    }
    if (st == SRGP_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("chol_inverse failed: %s", cudaGetErrorString(e));
            st = SRGP_ERR_CUDA;
        }
    }
    if (st == SRGP_OK) {
        if (L_out)
            cudaMemcpy2DAsync(L_out, (size_t)m * 8, dA.p, (size_t)mp * 8, (size_t)m * 8, m, cudaMemcpyDeviceToHost,
                              ctx->stream);
        if (Ainv_out)
            cudaMemcpy2DAsync(Ainv_out, (size_t)m * 8, dAinv.p, (size_t)mp * 8, (size_t)m * 8, m,
                              cudaMemcpyDeviceToHost, ctx->stream);
        double misc[8];
        cudaMemcpyAsync(misc, dMisc.p, 64, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        if (logdet_out) *logdet_out = misc[0];
        if (info_out) *info_out = *reinterpret_cast<int *>(&misc[1]);
    }
    return st;
}
