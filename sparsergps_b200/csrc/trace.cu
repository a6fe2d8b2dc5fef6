// trace.cu -- K5: trace-term entry points on MATERIALISED inputs (API parity with the R functions) and the
// fused all-theta reduction sum_ij Omega_ij dSigma12_ij/dlog(theta) that never materialises dSigma12.
//
//   trace_term_fun(cov_par, Sigma12, Sigma22, delta)     R/vi_functions.R:14-27
//   dtrace_term_dcov_par(cov_par, A_trace)               R/vi_functions.R:54-60
//   every `dSigma12_dtheta` use in delbo_dcov_par        R/vi_functions.R:344-398  -> srgp_omega_dk_reduce
#include <math.h>

#include "dense.cuh"
#include "fastexp.cuh"
#include "gauss.cuh"

namespace srgp {

__global__ void __launch_bounds__(256)
sum_vec_partial_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ part)
{
    __shared__ double red[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += x[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int k = 0; k < 8; k++) v += red[k];
        part[blockIdx.x] = v;
    }
}

__global__ void sum_small_vec_kernel(const double *__restrict__ part, int n, double *__restrict__ out, int accumulate)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double v = 0.0;
        for (int i = 0; i < n; i++) v += part[i];
        *out = accumulate ? *out + v : v;
    }
}

// *out (+)= sum x[0..n)   -- warp-shuffle tree, deterministic
static int sum_vec(srgp_ctx *ctx, const double *x, int64_t n, double *scratch, double *out, bool accumulate)
{
    cudaStream_t s = ctx->stream;
    KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
    sum_vec_partial_kernel<<<128, 256, 0, s>>>(x, n, scratch);
    SRGP_LAUNCH_CHECK();
    sum_small_vec_kernel<<<1, 32, 0, s>>>(scratch, 128, out, accumulate ? 1 : 0);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

// ------------------------------------------------------------------------------------------------
// sum_ij Omega_ij dK_ij/dlog(theta): one thread per row (coalesced Omega reads, the only HBM traffic),
// knots broadcast from shared memory, K and the scaled differences regenerated in registers.
//   part[block][0] = sum Omega K, [1 + c] = sum Omega K ((x_c - u_c)/l_c)^2, [1 + d] = sum Omega [x == u]
// ------------------------------------------------------------------------------------------------
constexpr int OD_THREADS = 128;
constexpr int OD_RPT = 2;                       // rows per thread: the knot loads from shared memory serve both
constexpr int OD_ROWS = OD_THREADS * OD_RPT;    // rows per CTA
constexpr int OD_COLS = 32;
constexpr int OD_STRIDE = SRGP_MAX_D + 8;

// FP64-pipe budget per entry (d = 8): 8 DADD + 8 DFMA (distance) + 10 (exp_tab, fastexp.cuh) + 8 DMUL + 8 DFMA
// (per-dimension sums) + 5 = 47 instructions for 8 bytes of Omega: on B200 (37 TF/s FP64 vs 6.5 TB/s HBM, ridge
// = 22 instructions per 8-byte entry) this kernel is bound by the FP64 pipe, not by HBM.
template <int DT>
__global__ void __launch_bounds__(OD_THREADS)
omega_dk_kernel(const double *__restrict__ Omega, int64_t ldo, const double *__restrict__ X, int64_t ldx,
                int64_t rows, const double *__restrict__ U, int m, int d_rt, GenParams p,
                double *__restrict__ part, int first)
{
    extern __shared__ double su[];   // [OD_COLS][d] scaled knots
    __shared__ double red[OD_THREADS / 32][OD_STRIDE];
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, OD_THREADS);
    const int d = DT > 0 ? DT : d_rt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = lane; t < OD_STRIDE; t += 32) red[warp][t] = 0.0;
    const double logs2 = log(p.sigma2);
    int64_t irow[OD_RPT];
    bool iv[OD_RPT];
    double xi[OD_RPT][DT > 0 ? DT : 1];
#pragma unroll
    for (int q = 0; q < OD_RPT; q++) {
        irow[q] = (int64_t)blockIdx.x * OD_ROWS + q * OD_THREADS + threadIdx.x;
        iv[q] = irow[q] < rows;
        if (DT > 0) {
#pragma unroll
            for (int c = 0; c < DT; c++) xi[q][c] = iv[q] ? X[irow[q] + ldx * c] * p.invl[c] : 0.0;
        }
    }
    double g0 = 0.0, gt = 0.0, gl[DT > 0 ? DT : 1];
#pragma unroll
    for (int c = 0; c < (DT > 0 ? DT : 1); c++) gl[c] = 0.0;
    const int col_tiles = (m + OD_COLS - 1) / OD_COLS;
    for (int jt = blockIdx.y; jt < col_tiles; jt += gridDim.y) {
        const int j0 = jt * OD_COLS;
        __syncthreads();
        for (int t = threadIdx.x; t < OD_COLS * d; t += OD_THREADS) {
            const int jj = t / d, c = t - jj * d;
            su[t] = (j0 + jj < m) ? U[j0 + jj + (int64_t)m * c] * p.invl[c] : 0.0;
        }
        __syncthreads();
        const int jmax = min(OD_COLS, m - j0);
        for (int jj = 0; jj < jmax; jj++) {
            if (DT > 0) {
                double uj[DT > 0 ? DT : 1];
#pragma unroll
                for (int c = 0; c < DT; c++) uj[c] = su[jj * DT + c];
#pragma unroll
                for (int q = 0; q < OD_RPT; q++) {
                    const double om = iv[q] ? __ldcs(Omega + irow[q] + ldo * (int64_t)(j0 + jj)) : 0.0;
                    double t[DT > 0 ? DT : 1], sq = 0.0;
#pragma unroll
                    for (int c = 0; c < DT; c++) {
                        t[c] = xi[q][c] - uj[c];
                        sq = fma(t[c], t[c], sq);
                    }
                    const double pk = om * exp_tab(fma(-0.5, sq, logs2), etab);
                    g0 += pk;
                    // identical scaled coordinates <=> sq == 0 (exact differences): quirk Q4 pairs
                    if (sq == 0.0) gt += om;
#pragma unroll
                    for (int c = 0; c < DT; c++) gl[c] = fma(pk * t[c], t[c], gl[c]);
                }
            } else {
                for (int q = 0; q < OD_RPT; q++) {
                    if (!iv[q]) continue;
                    const double om = __ldcs(Omega + irow[q] + ldo * (int64_t)(j0 + jj));
                    double sq = 0.0;
                    for (int c = 0; c < d; c++) {
                        // __dmul_rn: no FMA contraction, so identical points give exactly t = 0 (quirk Q4)
                        const double t = __dmul_rn(X[irow[q] + ldx * c], p.invl[c]) - su[jj * d + c];
                        sq = fma(t, t, sq);
                    }
                    const double pk = om * exp_tab(fma(-0.5, sq, logs2), etab);
                    g0 += pk;
                    if (sq == 0.0) gt += om;
                    for (int c = 0; c < d; c++) {
                        const double t = __dmul_rn(X[irow[q] + ldx * c], p.invl[c]) - su[jj * d + c];
                        atomicAdd(&red[warp][1 + c], pk * t * t);   // generic-d slow path (d > 8)
                    }
                }
            }
        }
    }
    // warp-shuffle reductions, then the warps through shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g0 += __shfl_xor_sync(0xffffffffu, g0, o);
        gt += __shfl_xor_sync(0xffffffffu, gt, o);
    }
    if (DT > 0) {
#pragma unroll
        for (int c = 0; c < DT; c++) {
            double v = gl[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][1 + c] = v;
        }
    }
    if (lane == 0) {
        red[warp][0] = g0;
        red[warp][1 + d] = gt;
    }
    __syncthreads();
    if (threadIdx.x < 2 + d) {
        double v = 0.0;
        for (int w = 0; w < OD_THREADS / 32; w++) v += red[w][threadIdx.x];
        double *slot = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * OD_STRIDE + threadIdx.x;
        *slot = first ? v : (*slot + v);
    }
}

__global__ void sum_slots_kernel(const double *__restrict__ part, int slots, int stride, int count,
                                 double *__restrict__ out)
{
    const int e = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < slots; i += 32) s += part[(int64_t)i * stride + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0 && e < count) out[e] = s;
}

// device-resident core: Omega_dev is rows x m with leading dimension ldo; results accumulate into slots.
static int omega_dk_dev(srgp_ctx *ctx, const GenParams &gp, const double *x_dev, int64_t ldx, int64_t rows,
                        const double *u_dev, int m, int d, const double *omega_dev, int64_t ldo, double *part,
                        int grid_x, int grid_y, int first)
{
    cudaStream_t s = ctx->stream;
    const size_t smem = sizeof(double) * OD_COLS * d;
    KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
    dim3 grid(grid_x, grid_y);
#define CALL(D) omega_dk_kernel<D><<<grid, OD_THREADS, smem, s>>>(omega_dev, ldo, x_dev, ldx, rows, u_dev, m, d, gp, part, first)
    switch (d) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 3: CALL(3); break;
    case 4: CALL(4); break;
    case 5: CALL(5); break;
    case 6: CALL(6); break;
    case 7: CALL(7); break;
    case 8: CALL(8); break;
    default: CALL(0); break;
    }
#undef CALL
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

static void finish_omega(int kernel, int d, double sigma, double tau, const double *raw, double *out)
{
    // raw: [sum Omega K, sum Omega K D_c ..., sum Omega over identical pairs]
    out[0] = 2.0 * raw[0];                                   // d/dlog sigma = 2 K
    if (kernel == SRGP_ARD) {
        for (int c = 0; c < d; c++) out[1 + c] = raw[1 + c];
        out[1 + d] = 2.0 * tau * tau * raw[1 + d];
    } else {
        double g = 0.0;
        for (int c = 0; c < d; c++) g += raw[1 + c];
        out[1] = g;
        out[2] = 2.0 * tau * tau * raw[1 + d];
    }
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_trace_term(srgp_ctx *ctx, double sigma, double tau, double delta, const double *Sigma12,
                               int64_t n, int64_t m, const double *Sigma22, double *out)
{
    if (!ctx || !Sigma12 || !Sigma22 || !out || n <= 0 || m <= 0 || m > 32768) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    GaussWS *w = gauss_ws(ctx);
    SRGP_TRY(plan(ctx, w, (int)m, 1));
    const int mp = w->mp;
    const size_t mm = (size_t)mp * mp;
    double *S = w->mat(GaussWS::M_T1), *Sinv = w->mat(GaussWS::M_SINV);
    // Sigma22 -> padded device matrix, Cholesky, explicit inverse (the reference: solve(Sigma22, t(Sigma12)))
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + GaussWS::S_INFO, 0, 16, s));
    SRGP_CUDA(cudaMemcpy2DAsync(S, (size_t)mp * 8, Sigma22, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice, s));
    SRGP_TRY(dense::pad_identity(ctx, s, S, mp, (int)m, 1.0));
    SRGP_TRY(dense::chol_inverse(ctx, s, S, mp, (int)m, w->dinv(0), w->mat(GaussWS::M_LINV), w->mat(GaussWS::M_X1),
                                 w->mat(GaussWS::M_TMP), Sinv, w->info(0), w->sc(GaussWS::S_LOGDET_S)));
    (void)mm;
    // stream Sigma12 through the chunk buffer: Z4_i = Sigma12[i,] Sigma22^-1 Sigma12[i,]^T, summed
    SRGP_TRY(ctx->tmp0.reserve((size_t)w->rows2 * 8));
    double *qsum = w->sc(GaussWS::S_SUMQ);
    for (int64_t r0 = 0; r0 < n; r0 += w->rows2) {
        const int rows = (int)std::min<int64_t>(w->rows2, n - r0);
        if (rows < w->rows2 || mp > m) SRGP_CUDA(cudaMemsetAsync(w->chunk.p, 0, (size_t)w->rows2 * mp * 8, s));
        SRGP_CUDA(cudaMemcpy2DAsync(w->chunk.p, (size_t)w->rows2 * 8, Sigma12 + r0, (size_t)n * 8, (size_t)rows * 8, m,
                                    cudaMemcpyHostToDevice, s));
        SRGP_TRY(rowform_chunk(ctx, w, Sinv, rows, ctx->tmp0.d()));
        SRGP_TRY(sum_vec(ctx, ctx->tmp0.d(), rows, w->part2.d(), qsum, r0 > 0));
    }
    SRGP_TRY(fetch_scalars(ctx, w));
    const double z4 = w->h_scal[GaussWS::S_SUMQ];
    *out = -(1.0 / (2.0 * tau * tau)) * ((double)n * (sigma * sigma + delta) - z4);
    return SRGP_OK;
}

extern "C" int srgp_dtrace_term_dcov_par(srgp_ctx *ctx, double tau, const double *A_trace, int64_t n, double *out)
{
    if (!ctx || !A_trace || !out || n < 0) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    SRGP_TRY(ctx->tmp0.reserve(std::max<size_t>(8, (size_t)n * 8)));
    SRGP_TRY(ctx->tmp1.reserve(256 * 8));
    SRGP_CUDA(cudaMemcpyAsync(ctx->tmp0.p, A_trace, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    SRGP_TRY(sum_vec(ctx, ctx->tmp0.d(), n, ctx->tmp1.d(), ctx->tmp1.d() + 200, false));
    double v = 0.0;
    SRGP_CUDA(cudaMemcpyAsync(&v, ctx->tmp1.d() + 200, 8, cudaMemcpyDeviceToHost, s));
    SRGP_CUDA(cudaStreamSynchronize(s));
    *out = -(1.0 / (2.0 * tau * tau)) * v;
    return SRGP_OK;
}

static int omega_check(srgp_ctx *ctx, int kernel, const void *x, int64_t n, const void *xu, int64_t m, int d,
                       const double *l, const void *Omega, const double *out)
{
    if (!ctx || !x || !xu || !l || !Omega || !out || n <= 0 || m <= 0 || d <= 0 || d > SRGP_MAX_D) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    return SRGP_OK;
}

extern "C" int srgp_omega_dk_reduce_dev(srgp_ctx *ctx, int kernel, const double *x_dev, int64_t n,
                                        const double *xu_dev, int64_t m, int d, double sigma, const double *l,
                                        double tau, const double *Omega_dev, double *out)
{
    SRGP_TRY(omega_check(ctx, kernel, x_dev, n, xu_dev, m, d, l, Omega_dev, out));
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    const int gx = (int)ceil_div(n, OD_ROWS);
    const int gy = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(m, OD_COLS), ceil_div(ctx->sm_count * 8, gx)));
    SRGP_TRY(ctx->tmp1.reserve(((size_t)gx * gy * OD_STRIDE + OD_STRIDE) * 8));
    SRGP_TRY(omega_dk_dev(ctx, gp, x_dev, n, n, xu_dev, (int)m, d, Omega_dev, n, ctx->tmp1.d(), gx, gy, 1));
    double *res = ctx->tmp1.d() + (size_t)gx * gy * OD_STRIDE;
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        sum_slots_kernel<<<d + 2, 32, 0, s>>>(ctx->tmp1.d(), gx * gy, OD_STRIDE, d + 2, res);
        SRGP_LAUNCH_CHECK();
    }
    double raw[OD_STRIDE];
    SRGP_CUDA(cudaMemcpyAsync(raw, res, (size_t)(d + 2) * 8, cudaMemcpyDeviceToHost, s));
    SRGP_CUDA(cudaStreamSynchronize(s));
    finish_omega(kernel, d, sigma, tau, raw, out);
    return SRGP_OK;
}

extern "C" int srgp_omega_dk_reduce(srgp_ctx *ctx, int kernel, const double *x, int64_t n, const double *xu,
                                    int64_t m, int d, double sigma, const double *l, double tau,
                                    const double *Omega, double *out)
{
    SRGP_TRY(omega_check(ctx, kernel, x, n, xu, m, d, l, Omega, out));
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    // rows are streamed in blocks so that the device copy of Omega stays bounded (256 MiB)
    const int64_t rows_blk = std::max<int64_t>(OD_ROWS, ((int64_t(256) << 20) / (8 * m)) / OD_ROWS * OD_ROWS);
    SRGP_TRY(ctx->in_x.reserve((size_t)n * d * 8));
    SRGP_TRY(ctx->in_xp.reserve((size_t)m * d * 8));
    SRGP_TRY(ctx->out_mat.reserve((size_t)std::min<int64_t>(rows_blk, n) * m * 8));
    SRGP_CUDA(cudaMemcpyAsync(ctx->in_x.p, x, (size_t)n * d * 8, cudaMemcpyHostToDevice, s));
    SRGP_CUDA(cudaMemcpyAsync(ctx->in_xp.p, xu, (size_t)m * d * 8, cudaMemcpyHostToDevice, s));
    const int gx = (int)ceil_div(std::min<int64_t>(rows_blk, n), OD_ROWS);
    const int gy = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(m, OD_COLS), ceil_div(ctx->sm_count * 8, gx)));
    SRGP_TRY(ctx->tmp1.reserve(((size_t)gx * gy * OD_STRIDE + OD_STRIDE) * 8));
    int first = 1;
    for (int64_t r0 = 0; r0 < n; r0 += rows_blk) {
        const int64_t rows = std::min<int64_t>(rows_blk, n - r0);
        SRGP_CUDA(cudaMemcpy2DAsync(ctx->out_mat.p, (size_t)rows * 8, Omega + r0, (size_t)n * 8, (size_t)rows * 8, m,
                                    cudaMemcpyHostToDevice, s));
        SRGP_TRY(omega_dk_dev(ctx, gp, ctx->in_x.d() + r0, n, rows, ctx->in_xp.d(), (int)m, d, ctx->out_mat.d(), rows,
                              ctx->tmp1.d(), gx, gy, first));
        first = 0;
    }
    double *res = ctx->tmp1.d() + (size_t)gx * gy * OD_STRIDE;
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        sum_slots_kernel<<<d + 2, 32, 0, s>>>(ctx->tmp1.d(), gx * gy, OD_STRIDE, d + 2, res);
        SRGP_LAUNCH_CHECK();
    }
    double raw[OD_STRIDE];
    SRGP_CUDA(cudaMemcpyAsync(raw, res, (size_t)(d + 2) * 8, cudaMemcpyDeviceToHost, s));
    SRGP_CUDA(cudaStreamSynchronize(s));
    finish_omega(kernel, d, sigma, tau, raw, out);
    return SRGP_OK;
}
