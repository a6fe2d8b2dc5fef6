// gauss_fic.cu -- Gaussian FIC objective + gradient on the resident shard.
//
// Reference: obj_fun_norm (R/laplace_approx_obj_funs.R:6-52), Z construction
// (R/laplace_gradient_ascent.R:1259-1263), dlogp_dcov_par (R/laplace_approx_gradient.R:720-968).
// Reduced form (oracle/reduced_model.py::fic_obj_grad == literal transcription to 1e-13):
//   S = K_uu + delta I;  q_i = K_i S^-1 K_i^T;  Z_i = sigma^2 + tau^2 + delta - q_i;  B_i = 1/Z_i;  r = y - mu
//   G_B = K^T diag(B) K, b = K^T (B r), C = (S + G_B)^-1, v = C b, beta = S^-1 (b - G_B v), M2 = C G_B S^-1
//   obj = -sum B r^2/2 + b^T v/2 - (sum log Z - log|S| + log|S + G_B|)/2 - n log(2 pi)/2
//   c_i = K_i C K_i^T, alpha_i = B_i (r_i - K_i v), w_i = B_i - B_i^2 c_i, rho_i = alpha_i^2/2 - w_i/2
//   Omega = diag(-B - 2 rho) K S^-1 + diag(B) K M2 + alpha beta^T
//   N = S^-1 G_B S^-1/2 - S^-1 G_B M2/2 - beta beta^T/2 + S^-1 G_rho S^-1,  G_rho = K^T diag(rho) K
//   g_theta = sum Omega o dK + sum N o dS + A1(theta) sum rho   (A1 = 2 sigma^2, 0, 2 tau^2; dS(tau) = 0)
// With C (S + G_B) = I:  M2 = S^-1 - C,  Omega = -2 diag(rho) K S^-1 - diag(B) K C + alpha beta^T,
//   N = M2/2 - beta beta^T/2 + S^-1 G_rho S^-1,  and
//   sum Omega o K o D_c = sum_i ( -2 rho_i R_ic - B_i U_ic + alpha_i E_ic ),
//   R_ic = sum_j (K S^-1)_ij K_ij D_ijc,  U_ic = sum_j (K C)_ij K_ij D_ijc,  E_ic = sum_j beta_j K_ij D_ijc  (D_ij0 = 1)
// -- per-row sums of the two products the row forms q_i = R_i0 and c_i = U_i0 need anyway.
// Row passes: K S^-1 (q, R), weighted Gram G_B, K C (c, U, E, K v), weighted Gram G_rho: 2 K*M + 2 SYRK.
// With the knot-location gradient (column sums of Omega o K need rho, alpha first): q, G_B, c / K v, the two
// explicit Omega passes, G_rho.
#include <math.h>

#include "dense.cuh"
#include "gauss.cuh"

namespace srgp {

using W = GaussWS;

__device__ __forceinline__ double block_reduce_256(double v, double *red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += red[k];
    return t;
}

// Z, B, B r from q; per-block partial sums of B r^2 and log Z  (part[2 * block + {0, 1}])
__global__ void __launch_bounds__(256)
fic_rows1_kernel(const double *__restrict__ q, const double *__restrict__ r, int64_t n, double zconst,
                 double *__restrict__ B, double *__restrict__ Br, double *__restrict__ part)
{
    __shared__ double red[8];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double z = zconst - q[i];
        const double b = 1.0 / z, ri = r[i];
        B[i] = b;
        Br[i] = b * ri;
        s0 = fma(b * ri, ri, s0);
        s1 += log(z);
    }
    s0 = block_reduce_256(s0, red);
    s1 = block_reduce_256(s1, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = s0;
        part[2 * blockIdx.x + 1] = s1;
    }
}

// alpha, rho, rs1 = -B - 2 rho from c, K v; per-block partial sums of rho; rhos = rho + B / 2 = (alpha^2 + B^2 c) / 2 >= 0,
// the weights of the shifted Gram that gives G_rho from ONE slice set (see pass 2c)
__global__ void __launch_bounds__(256)
fic_rows2_kernel(const double *__restrict__ c, const double *__restrict__ kv, const double *__restrict__ r,
                 const double *__restrict__ B, int64_t n, double *__restrict__ alpha, double *__restrict__ rho,
                 double *__restrict__ rs1, double *__restrict__ rhos, double *__restrict__ part)
{
    __shared__ double red[8];
    double sr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double b = B[i];
        const double a = b * (r[i] - kv[i]);
        const double wi = b - b * b * c[i];
        const double rh = 0.5 * a * a - 0.5 * wi;
        alpha[i] = a;
        rho[i] = rh;
        rs1[i] = -b - 2.0 * rh;
        rhos[i] = 0.5 * (a * a + b * b * c[i]);
        sr += rh;
    }
    sr = block_reduce_256(sr, red);
    if (threadIdx.x == 0) part[blockIdx.x] = sr;
}

// out[c] = sum over the blocks of part[c * ROW_BLOCKS + block]   (one block per c)
__global__ void sum_part_rows_kernel(const double *__restrict__ part, int count, double *__restrict__ out)
{
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[blockIdx.x * count + i];
    s = block_reduce_256(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void sum_strided_kernel(const double *__restrict__ part, int count, int stride, int offset,
                                   double *__restrict__ out)
{
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[i * stride + offset];
    s = block_reduce_256(s, red);
    if (threadIdx.x == 0) *out = s;
}

constexpr int ROW_BLOCKS = 128;

// g[c] = sum_i ( -2 rho_i R_ic - B_i U_ic + alpha_i E_ic ),  c = 0..d  (= sum Omega o K o D_c with D_0 = 1):
// one block column per c; per-block partials part[c * ROW_BLOCKS + block]
__global__ void __launch_bounds__(256)
fic_rowd_reduce_kernel(const double *__restrict__ R, const double *__restrict__ U, const double *__restrict__ E,
                       int64_t stride, const double *__restrict__ B, const double *__restrict__ rho,
                       const double *__restrict__ alpha, int64_t n, double *__restrict__ part)
{
    __shared__ double red[8];
    const int c = blockIdx.y;
    const double *Rc = R + (int64_t)c * stride, *Uc = U + (int64_t)c * stride, *Ec = E + (int64_t)c * stride;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += -2.0 * rho[i] * Rc[i] - B[i] * Uc[i] + alpha[i] * Ec[i];
    s = block_reduce_256(s, red);
    if (threadIdx.x == 0) part[c * ROW_BLOCKS + blockIdx.x] = s;
}

int gauss_fic(srgp_ctx *ctx, GaussWS *w, int kernel, double sigma, const double *l, double tau, double delta,
              double *obj, double *grad)
{
    cudaStream_t s = ctx->stream;
    const int mp = w->mp, m = w->m, d = w->d;
    const int64_t n = ctx->n;
    const size_t mm = (size_t)mp * mp;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    w->k_reuse = true;     // two K*M passes (1a, 2a) over this K: the second reads the image of the first (gauss_i8.cu)
    const double zconst = sigma * sigma + tau * tau + delta;

    double *GB = w->red1.d(), *b = GB + mm, *tail = b + mp;   // allreduce buffer of pass 1
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W::S_INFO, 0, 16, s));
    SRGP_TRY(coin_reset(ctx, w));

    double *S = w->mat(W::M_S), *Sinv = w->mat(W::M_SINV), *A = w->mat(W::M_A), *C = w->mat(W::M_C);
    double *Linv = w->mat(W::M_LINV), *LinvT = w->mat(W::M_X1), *tmp = w->mat(W::M_TMP);
    double *CG = w->mat(W::M_CG), *M2 = w->mat(W::M_CGS), *SG = w->mat(W::M_SG), *SGS = w->mat(W::M_SGS);
    double *N = w->mat(W::M_N), *T1 = w->mat(W::M_T1), *T2 = w->mat(W::M_T2), *Grho = w->mat(W::M_X2);
    double *v = w->vec(W::V_V), *gv = w->vec(W::V_GV), *tv = w->vec(W::V_TMP), *beta = w->vec(W::V_BETA);
    double *gsc = w->gemv_scratch();
    double *q = w->rowv(0, n), *Bv = w->rowv(1, n), *Br = w->rowv(2, n), *kv = w->rowv(3, n);
    double *alpha = w->rowv(4, n), *rho = w->rowv(5, n), *rhos = w->rowv(6, n);
    double *cq = q, *rs1 = Br;   // reuse: q is dead once B exists, B r once b is reduced

    // ---- S, S^-1 ----------------------------------------------------------------------------------------
    SRGP_TRY(assemble_dev_ld(ctx, s, kernel, w->U.d(), m, d, sigma, l, delta, S, mp));
    SRGP_TRY(dense::pad_identity(ctx, s, S, mp, m, 1.0));
    SRGP_CUDA(cudaMemcpyAsync(T1, S, mm * 8, cudaMemcpyDeviceToDevice, s));
    double *LinvS = w->mat(W::M_L1), *LinvTS = w->mat(W::M_L2);   // kept: vector solves with S use the factors
    SRGP_TRY(dense::chol_inverse(ctx, s, T1, mp, m, w->dinv(0), LinvS, LinvTS, tmp, Sinv, w->info(0),
                                 w->sc(W::S_LOGDET_S)));

    // ---- pass 1a: q_i (and, when a gradient without knot terms is wanted, R_ic = sum_j (K S^-1)_ij K_ij D_ijc) ;
    //      rows: Z, B, B r, sum B r^2, sum log Z --------------------------------------------------------------------
    // Omega = -2 diag(rho) K S^-1 - diag(B) K C + alpha beta^T (M2 = S^-1 - C), so sum Omega o K o D_c splits into
    // per-row sums of the two products the row forms q_i, c_i need anyway: no third / fourth K * M pass.  The
    // knot-location gradient needs column sums of Omega o K, i.e. rho and alpha BEFORE the products: it keeps the
    // two explicit Omega passes.
    const bool rowd_path = grad && !w->want_knots;
    const int64_t rstride = (int64_t)W::row_stride(n);
    double *Rd = nullptr, *Ud = nullptr, *Ed = nullptr;
    if (rowd_path) {
        SRGP_TRY(w->rowd.reserve((size_t)(3 * (d + 1) + 1) * rstride * 8));
        Rd = w->rowd.d();
        Ud = Rd + (int64_t)(d + 1) * rstride;
        Ed = Ud + (int64_t)(d + 1) * rstride;      // followed by K v
        SRGP_TRY(gauss_rowd(ctx, w, gp, Sinv, nullptr, nullptr, Rd, rstride));
        q = Rd;
    } else {
        SRGP_TRY(gauss_rowform(ctx, w, gp, Sinv, nullptr, q, nullptr));
    }
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 3);
        fic_rows1_kernel<<<ROW_BLOCKS, 256, 0, s>>>(q, w->r.d(), n, zconst, Bv, Br, w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_strided_kernel<<<1, 256, 0, s>>>(w->part2.d(), ROW_BLOCKS, 2, 0, tail);
        SRGP_LAUNCH_CHECK();
        sum_strided_kernel<<<1, 256, 0, s>>>(w->part2.d(), ROW_BLOCKS, 2, 1, tail + 1);
        SRGP_LAUNCH_CHECK();
    }
    SRGP_TRY(set_scalar(ctx, tail + 2, (double)n));
    // ---- pass 1b: G_B, b ---------------------------------------------------------------------------------
    SRGP_TRY(gauss_pass1(ctx, w, gp, Bv, Br, GB, b, true));        // B = 1 / Z > 0
    SRGP_TRY(comm_allreduce(ctx, GB, mm + mp + 3, s));
    SRGP_TRY(copy_scalar(ctx, w->sc(W::S_X), tail, 3));   // s0, s1, n (global)

    // ---- m x m: C, v, beta, M2 -----------------------------------------------------------------------------
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, S, 1.0, GB, 0.0, A));
    SRGP_TRY(dense::chol_inverse(ctx, s, A, mp, m, w->dinv(1), Linv, LinvT, tmp, C, w->info(1),
                                 w->sc(W::S_LOGDET_A)));
    // vector solves through the triangular factors (forward error ~ sqrt(cond) instead of cond)
    double *t1 = w->vec(W::V_T1), *t2 = w->vec(W::V_T2);
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, LinvT, b, 0.0, nullptr, t1));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, Linv, t1, 0.0, nullptr, v));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, GB, v, 0.0, nullptr, gv));
    SRGP_TRY(axpby_vec(ctx, mp, 1.0, b, -1.0, gv, tv));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, LinvTS, tv, 0.0, nullptr, t2));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, LinvS, t2, 0.0, nullptr, beta));
    SRGP_TRY(dense::dot_v(ctx, s, m, t1, t1, w->sc(W::S_BV)));   // b^T (S+G_B)^-1 b = |L^-1 b|^2
    if (grad) {
        // M2 = C G_B S^-1 = S^-1 - C   (G_B = A - S, C A = I): no product needed
        SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, Sinv, -1.0, C, 0.0, M2));

        // ---- pass 2a: c_i, (K v)_i ; rows: alpha, rho, -B - 2 rho, sum rho -----------------------------------
        double *p2 = w->sc(W::S_P2);
        if (rowd_path) {
            // pairs recorded by pass 1a carry (K S^-1)_ij; restart the list so that it holds (K C)_ij only
            SRGP_TRY(coin_reset(ctx, w));
            SRGP_TRY(gauss_rowd(ctx, w, gp, C, beta, v, Ud, rstride));
            cq = Ud;
            kv = Ed + (int64_t)(d + 1) * rstride;
        } else {
            SRGP_TRY(gauss_rowform(ctx, w, gp, C, v, cq, kv));
        }
        {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
            fic_rows2_kernel<<<ROW_BLOCKS, 256, 0, s>>>(cq, kv, w->r.d(), Bv, n, alpha, rho, rs1, rhos, w->part2.d());
            SRGP_LAUNCH_CHECK();
            sum_strided_kernel<<<1, 256, 0, s>>>(w->part2.d(), ROW_BLOCKS, 1, 0, p2 + d + 2);
            SRGP_LAUNCH_CHECK();
        }
        if (rowd_path) {
            // ---- sum Omega o K o D_c from the per-row sums; quirk Q4 pairs ---------------------------------------
            {
                KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
                fic_rowd_reduce_kernel<<<dim3(ROW_BLOCKS, d + 1), 256, 0, s>>>(Rd, Ud, Ed, rstride, Bv, rho, alpha, n,
                                                                              w->part2.d());
                SRGP_LAUNCH_CHECK();
                sum_part_rows_kernel<<<d + 1, 256, 0, s>>>(w->part2.d(), ROW_BLOCKS, p2);
                SRGP_LAUNCH_CHECK();
            }
            SRGP_TRY(coin_fix_fic(ctx, w, gp, Sinv, Bv, rho, alpha, beta, p2 + 1 + d));
        } else {
            // ---- pass 2b: the two Omega terms share the per-CTA slots ------------------------------------------
            SRGP_TRY(gauss_pass2(ctx, w, gp, Sinv, rs1, alpha, beta, nullptr, false));
            SRGP_TRY(gauss_pass2(ctx, w, gp, M2, Bv, nullptr, beta, p2, true));
            // quirk Q4: sum of Omega_ij over bit-identical (row, knot) pairs (both terms were recorded)
            SRGP_TRY(coin_fix(ctx, w, gp, Sinv, 0.0, p2 + 1 + d));
        }
        // ---- pass 2c: G_rho = K^T diag(rho) K.  rho has either sign, but rho_i = alpha_i^2/2 - B_i (1 - B_i c_i)/2 with
        //      c_i = K_i C K_i^T >= 0, so rho_i + B_i/2 = (alpha_i^2 + B_i^2 c_i)/2 >= 0:  G_rho = K^T diag(rho + B/2) K - G_B/2
        //      is ONE slice set sqrt(rho + B/2) K (half the generator output and launches of the two-set weighted Gram)
        //      minus the Gram pass 1b left behind (already summed over the ranks; the shift is linear, so the allreduce
        //      of the shifted partial Grams commutes with it) ----------------------------------------------------------
        double *red2 = w->mat(W::M_T2);   // allreduce buffer of pass 2: [G_rho + G_B/2 | p2 (d + 4)]
        SRGP_TRY(gauss_pass1(ctx, w, gp, rhos, rhos, red2, tv, true));
        SRGP_TRY(copy_scalar(ctx, red2 + mm, p2, W::p2_len(d)));
        SRGP_TRY(comm_allreduce(ctx, red2, mm + W::p2_len(d), s));
        SRGP_TRY(copy_scalar(ctx, p2, red2 + mm, W::p2_len(d)));
        SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, red2, -0.5, GB, 0.0, Grho));

        // ---- N = S^-1 G_B S^-1/2 - S^-1 G_B M2/2 - beta beta^T/2 + S^-1 G_rho S^-1; with M2 = S^-1 - C and
        //      S^-1 G_B C = S^-1 (A - S) C = S^-1 - C the first two terms collapse to M2/2 --------------------------
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, Sinv, mp, Grho, mp, 0.0, SG, mp));
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, SG, mp, Sinv, mp, 0.0, SGS, mp));
        SRGP_TRY(dense::axpby(ctx, s, mp, m, 0.5, M2, 1.0, SGS, 0.0, N));
        SRGP_TRY(dense::ger(ctx, s, mp, -0.5, beta, beta, N));
        SRGP_TRY(ns_reduce(ctx, w, gp, N, S, delta, w->sc(W::S_NS), s));
        if (w->want_knots) SRGP_TRY(knot_finish(ctx, w, gp, N, S));
    }
    SRGP_TRY(fetch_scalars(ctx, w));
    if (grad) SRGP_TRY(coin_check(w));

    const double *h = w->h_scal;
    const double s0 = h[W::S_X], s1 = h[W::S_X + 1], ntot = h[W::S_X + 2];
    *obj = -0.5 * s0 + 0.5 * h[W::S_BV] - 0.5 * (s1 - h[W::S_LOGDET_S] + h[W::S_LOGDET_A]) - 0.5 * ntot * log(2.0 * M_PI);
    if (grad) {
        const double *p2 = h + W::S_P2, *ns = h + W::S_NS;
        const double sum_rho = p2[d + 2];
        grad[0] = 2.0 * p2[0] + 2.0 * ns[0] + 2.0 * sigma * sigma * sum_rho;
        if (kernel == SRGP_ARD) {
            for (int c = 0; c < d; c++) grad[1 + c] = p2[1 + c] + ns[1 + c];
        } else {
            double g = 0.0;
            for (int c = 0; c < d; c++) g += p2[1 + c] + ns[1 + c];
            grad[1] = g;
        }
        const int ti = (kernel == SRGP_ARD) ? 1 + d : 2;
        grad[ti] = 2.0 * tau * tau * sum_rho + 2.0 * tau * tau * p2[1 + d];
    }
    return SRGP_OK;
}

}  // namespace srgp

// ====================================================================================================
// posterior at the knots, and prediction
// ====================================================================================================
namespace srgp {

// u_mean = muu + b - G v, u_var = S - G + G C G with G = K^T diag(1/Z) K, b = K^T ((y - mu)/Z), C = (S + G)^-1,
// v = C b.  Z = tau^2 + delta (VI: R/vi_functions.R:753,1160-1180) or sigma^2 + tau^2 + delta - q_i
// (FIC: R/laplace_gradient_ascent.R:1259-1263,1637-1656).
int gauss_posterior(srgp_ctx *ctx, GaussWS *w, int model, int kernel, double sigma, const double *l, double tau,
                    double delta, double *u_plus /* dev, mp: b - G v */, double *u_var_dev /* dev, mp x mp */)
{
    cudaStream_t s = ctx->stream;
    const int mp = w->mp, m = w->m, d = w->d;
    const int64_t n = ctx->n;
    const size_t mm = (size_t)mp * mp;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    w->k_reuse = false;
    double *G = w->red1.d(), *b = G + mm;
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W::S_INFO, 0, 16, s));
    double *S = w->mat(W::M_S), *Sinv = w->mat(W::M_SINV), *A = w->mat(W::M_A), *C = w->mat(W::M_C);
    double *Linv = w->mat(W::M_LINV), *LinvT = w->mat(W::M_X1), *tmp = w->mat(W::M_TMP), *T1 = w->mat(W::M_T1);
    double *v = w->vec(W::V_V), *gv = w->vec(W::V_GV), *t1 = w->vec(W::V_T1), *gsc = w->gemv_scratch();
    SRGP_TRY(assemble_dev_ld(ctx, s, kernel, w->U.d(), m, d, sigma, l, delta, S, mp));
    SRGP_TRY(dense::pad_identity(ctx, s, S, mp, m, 1.0));
    if (model == SRGP_VI) {
        const double B = 1.0 / (tau * tau + delta);
        SRGP_TRY(gauss_pass1(ctx, w, gp, nullptr, w->r.d(), G, b));
        SRGP_TRY(comm_allreduce(ctx, G, mm + mp, s));
        SRGP_TRY(dense::axpby(ctx, s, mp, m, B, G, 0.0, nullptr, 0.0, G));
        SRGP_TRY(axpby_vec(ctx, mp, B, b, 0.0, nullptr, b));
    } else {
        SRGP_CUDA(cudaMemcpyAsync(T1, S, mm * 8, cudaMemcpyDeviceToDevice, s));
        SRGP_TRY(dense::chol_inverse(ctx, s, T1, mp, m, w->dinv(0), Linv, LinvT, tmp, Sinv, w->info(0),
                                     w->sc(W::S_LOGDET_S)));
        double *q = w->rowv(0, n), *Bv = w->rowv(1, n), *Br = w->rowv(2, n);
        SRGP_TRY(gauss_rowform(ctx, w, gp, Sinv, nullptr, q, nullptr));
        {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            fic_rows1_kernel<<<ROW_BLOCKS, 256, 0, s>>>(q, w->r.d(), n, sigma * sigma + tau * tau + delta, Bv, Br,
                                                        w->part2.d());
            SRGP_LAUNCH_CHECK();
        }
        SRGP_TRY(gauss_pass1(ctx, w, gp, Bv, Br, G, b, true));
        SRGP_TRY(comm_allreduce(ctx, G, mm + mp, s));
    }
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, S, 1.0, G, 0.0, A));
    SRGP_TRY(dense::chol_inverse(ctx, s, A, mp, m, w->dinv(1), Linv, LinvT, tmp, C, w->info(1), w->sc(W::S_LOGDET_A)));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, Linv, b, 0.0, nullptr, t1, gsc));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, LinvT, t1, 0.0, nullptr, v, gsc));
    SRGP_TRY(dense::gemv(ctx, s, mp, -1.0, G, v, 1.0, b, u_plus, gsc));       // b - G v
    (void)gv;
    // u_var = S - G + G C G
    double *GC = w->mat(W::M_CG);
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, G, mp, C, mp, 0.0, GC, mp));
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, GC, mp, G, mp, 0.0, u_var_dev, mp));
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, u_var_dev, -1.0, G, 0.0, u_var_dev));
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, u_var_dev, 1.0, S, 0.0, u_var_dev));
    return SRGP_OK;
}

__global__ void predict_finish_kernel(const double *__restrict__ kw, const double *__restrict__ kq, int64_t n,
                                      const double *__restrict__ mu, double var_const, double *__restrict__ mean,
                                      double *__restrict__ var)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        mean[i] = (mu ? mu[i] : 0.0) + kw[i];
        var[i] = var_const + kq[i];
    }
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_gauss_posterior_u(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m,
                                      const double *muu, double sigma, const double *l, double tau, double delta,
                                      double *u_mean, double *u_var)
{
    if (!ctx || !xu || !l || !u_mean || !u_var || m <= 0 || m > 32768 || (model != SRGP_VI && model != SRGP_FIC)) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("srgp_gauss_posterior_u called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    SRGP_TRY(use_device(ctx));
    GaussWS *w = gauss_ws(ctx);
    SRGP_TRY(plan(ctx, w, (int)m, ctx->d));
    cudaStream_t s = ctx->stream;
    SRGP_TRY(upload_knots(w, xu, (size_t)m * ctx->d * 8, s));
    double *uplus = w->vec(GaussWS::V_T3), *uvar = w->mat(GaussWS::M_N);
    SRGP_TRY(gauss_posterior(ctx, w, model, kernel, sigma, l, tau, delta, uplus, uvar));
    std::vector<double> up((size_t)w->mp);
    SRGP_CUDA(cudaMemcpyAsync(up.data(), uplus, (size_t)w->mp * 8, cudaMemcpyDeviceToHost, s));
    SRGP_CUDA(cudaMemcpy2DAsync(u_var, (size_t)m * 8, uvar, (size_t)w->mp * 8, (size_t)m * 8, m, cudaMemcpyDeviceToHost, s));
    SRGP_TRY(fetch_scalars(ctx, w));
    for (int64_t j = 0; j < m; j++) u_mean[j] = (muu ? muu[j] : 0.0) + up[j];
    return SRGP_OK;
}

extern "C" int srgp_predict(srgp_ctx *ctx, int kernel, const double *x_pred, int64_t n_pred, int d,
                            const double *mu_pred, const double *xu, int64_t m, const double *muu,
                            const double *u_mean, const double *u_var, double sigma, const double *l,
                            double s22_nugget, double var_const, double *pred_mean, double *pred_var)
{
    if (!ctx || !x_pred || !xu || !u_mean || !u_var || !l || !pred_mean || !pred_var || n_pred <= 0 || m <= 0 ||
        m > 32768 || d <= 0 || d > SRGP_MAX_D) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    GaussWS *w = gauss_ws(ctx);
    // the prediction points temporarily take the place of the resident shard for one row-form pass
    const double *Xsave = ctx->Xp;
    const int64_t nsave = ctx->n;
    const int dsave = ctx->d;
    SRGP_TRY(ctx->in_x.reserve((size_t)n_pred * d * 8));
    SRGP_CUDA(cudaMemcpyAsync(ctx->in_x.p, x_pred, (size_t)n_pred * d * 8, cudaMemcpyHostToDevice, s));
    ctx->Xp = ctx->in_x.d();
    ctx->data_version++;
    ctx->n = n_pred;
    ctx->d = d;
    int st = plan(ctx, w, (int)m, d);
    const int mp = w->mp;
    const size_t mm = (size_t)mp * mp;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    w->k_reuse = false;
    double *S = w->mat(GaussWS::M_S), *Sinv = w->mat(GaussWS::M_SINV), *UV = w->mat(GaussWS::M_A);
    double *T1 = w->mat(GaussWS::M_T1), *T2 = w->mat(GaussWS::M_T2), *Tm = w->mat(GaussWS::M_MOP);
    double *dv = w->vec(GaussWS::V_B), *wv = w->vec(GaussWS::V_V), *gsc = w->gemv_scratch();
    std::vector<double> diff((size_t)mp, 0.0);
    for (int64_t j = 0; j < m; j++) diff[j] = u_mean[j] - (muu ? muu[j] : 0.0);
    if (st == SRGP_OK) st = ctx->tmp0.reserve((size_t)n_pred * 8 * 4 + 64) ;
    auto run = [&]() -> int {
        SRGP_CUDA(cudaMemsetAsync(w->scal.d() + GaussWS::S_INFO, 0, 16, s));
        SRGP_TRY(upload_knots(w, xu, (size_t)m * d * 8, s));
        SRGP_CUDA(cudaMemcpyAsync(dv, diff.data(), (size_t)mp * 8, cudaMemcpyHostToDevice, s));
        SRGP_CUDA(cudaMemsetAsync(UV, 0, mm * 8, s));
        SRGP_CUDA(cudaMemcpy2DAsync(UV, (size_t)mp * 8, u_var, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice, s));
        // Sigma22 (nugget delta for Gaussian models, tau^2 + delta otherwise), its inverse
        SRGP_TRY(assemble_dev_ld(ctx, s, kernel, w->U.d(), m, d, sigma, l, s22_nugget, S, mp));
        SRGP_TRY(dense::pad_identity(ctx, s, S, mp, (int)m, 1.0));
        SRGP_CUDA(cudaMemcpyAsync(T1, S, mm * 8, cudaMemcpyDeviceToDevice, s));
        SRGP_TRY(dense::chol_inverse(ctx, s, T1, mp, (int)m, w->dinv(0), w->mat(GaussWS::M_LINV), w->mat(GaussWS::M_X1),
                                     w->mat(GaussWS::M_TMP), Sinv, w->info(0), w->sc(GaussWS::S_LOGDET_S)));
        // w = S^-1 (u_mean - muu) through the factors; T = -S^-1 + S^-1 u_var S^-1
        SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, w->mat(GaussWS::M_LINV), dv, 0.0, nullptr, w->vec(GaussWS::V_T1), gsc));
        SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, w->mat(GaussWS::M_X1), w->vec(GaussWS::V_T1), 0.0, nullptr, wv, gsc));
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, Sinv, mp, UV, mp, 0.0, T2, mp));      // u_var symmetric
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, T2, mp, Sinv, mp, 0.0, Tm, mp));
        SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, Tm, -1.0, Sinv, 0.0, Tm));
        // padding block of Tm must not contribute: K has zero columns there, nothing to do
        double *kq = ctx->tmp0.d(), *kw = kq + n_pred, *pm = kw + n_pred, *pv = pm + n_pred;
        SRGP_TRY(gauss_rowform(ctx, w, gp, Tm, wv, kq, kw));
        const double *mu_dev = nullptr;
        if (mu_pred) {
            SRGP_TRY(ctx->tmp1.reserve((size_t)n_pred * 8));
            SRGP_CUDA(cudaMemcpyAsync(ctx->tmp1.p, mu_pred, (size_t)n_pred * 8, cudaMemcpyHostToDevice, s));
            mu_dev = ctx->tmp1.d();
        }
        {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            predict_finish_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(kw, kq, n_pred, mu_dev, var_const, pm, pv);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaMemcpyAsync(pred_mean, pm, (size_t)n_pred * 8, cudaMemcpyDeviceToHost, s));
        SRGP_CUDA(cudaMemcpyAsync(pred_var, pv, (size_t)n_pred * 8, cudaMemcpyDeviceToHost, s));
        return fetch_scalars(ctx, w);
    };
    if (st == SRGP_OK) st = run();
    if (st != SRGP_OK) cudaStreamSynchronize(s);
    ctx->Xp = Xsave;
    ctx->data_version++;
    ctx->n = nsave;
    ctx->d = dsave;
    return st;
}
