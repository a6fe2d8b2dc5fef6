// laplace.cu -- K7: sparse Laplace models (Bernoulli, Poisson) on the resident shard.
//
//   srgp_laplace_newton : newtrap_sparseGP + newtrap_sparseGP_update (reference R/newtrap_sparseGP.R:6-186,234-325),
//                         objectives obj_fun_bern / obj_fun_pois (R/laplace_approx_obj_funs.R:189-341,108-174),
//                         grad_loglik_fn_* and d{1,2,3}log_py_dff_* (R/derivative_functions_of_data_likelihoods.R)
//   srgp_laplace_grad   : dlogq_dcov_par (R/laplace_approx_gradient.R:25-339)
//
// theta is fixed during the Newton loop, so the shard's K (rows x m, row-major) is generated once and kept in
// HBM; one iteration is then: elementwise likelihood terms, ONE weighted Gram K^T diag(omega) K on DMMA
// (omega = -W / (1 - Z W) serves both the update's R3 and the objective's R2 -- SURVEY.md App. B.4), an m x m
// Cholesky, and four matrix-vector passes over K.  Two NCCL allreduces per iteration:
// [G_omega | K^T((ff - mu)/Z) | 3 scalars] and [K^T(e grad_psi) | #rows with |grad_psi| > tol].
// The reference's quirks are kept verbatim: Bernoulli W for y = 1 (Q1), 2 dK GG (Q2), tau^2 kept in Sigma22.
// Algebra checked against the literal transcription by tests/test_oracle.py (oracle/reduced_model.py).
#include <math.h>

#include <vector>

#include "dense.cuh"
#include "gauss.cuh"
#include "gemm.cuh"

namespace srgp {

using W_ = GaussWS;

__device__ __forceinline__ double softplus(double x)   // log(1 + e^x), stable
{
    const double t = log1p(exp(-fabs(x)));
    return x > 0.0 ? x + t : t;
}

__device__ __forceinline__ double block_sum_256(double v, double *red)
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

// Likelihood terms at ff (R/derivative_functions_of_data_likelihoods.R:7-30,90-183):
//   d1, W (second derivative, quirk Q1 kept), W3 (third), e = 1/(1 - Z W), omega = -W/(1 - Z W), rz = (ff - mu)/Z
//   part[3 b + {0,1,2}] = sum (ff-mu)^2/Z, log p(y|ff), sum log(1 - W Z)
template <int FAMILY>
__global__ void __launch_bounds__(256)
lap_rows_kernel(const double *__restrict__ ff, const double *__restrict__ y, const double *__restrict__ mu,
                const double *__restrict__ Z, int64_t n, double pois_m, double *__restrict__ d1o,
                double *__restrict__ Wo, double *__restrict__ W3o, double *__restrict__ eo,
                double *__restrict__ omo, double *__restrict__ rzo, double *__restrict__ part)
{
    __shared__ double red[8];
    double s_quad = 0.0, s_lpy = 0.0, s_lz2 = 0.0, s_neg = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double f = ff[i], yi = y[i], z = Z[i], fm = f - (mu ? mu[i] : 0.0);
        double d1, w, w3, lpy;
        if (FAMILY == SRGP_BERNOULLI) {
            const double pi = 1.0 / (1.0 + exp(-f));
            const double dpi = pi * (1.0 - pi);
            d1 = yi * (1.0 - pi) - pi + yi * pi;
            w = (1.0 - 2.0 * pi) * (yi - pi) - (yi * (1.0 - pi) * (1.0 - pi) + pi * pi + yi * pi * pi);
            w3 = -2.0 * dpi * (yi - pi) - dpi * (1.0 - 2.0 * pi) -
                 (2.0 * yi * (1.0 - pi) * (-dpi) + 2.0 * pi * dpi + 2.0 * yi * pi * dpi);
            lpy = yi * (-softplus(-f)) + (1.0 - yi) * (-softplus(f));
        } else {
            const double ef = pois_m * exp(f);
            d1 = yi - ef;
            w = -ef;
            w3 = -ef;
            lpy = yi * log(pois_m) - lgamma(yi + 1.0) - ef + yi * f;
        }
        const double z2 = 1.0 - w * z;
        d1o[i] = d1;
        Wo[i] = w;
        if (W3o) W3o[i] = w3;
        eo[i] = 1.0 / z2;
        omo[i] = -w / z2;
        rzo[i] = fm / z;
        s_quad = fma(fm, fm / z, s_quad);
        s_lpy += lpy;
        s_lz2 += log(z2);
        if (!(-w / z2 >= 0.0)) s_neg += 1.0;      // omega < 0 or NaN: the one-slice-set Gram (sqrt(omega) K) does not apply
    }
    s_quad = block_sum_256(s_quad, red);
    s_lpy = block_sum_256(s_lpy, red);
    s_lz2 = block_sum_256(s_lz2, red);
    s_neg = block_sum_256(s_neg, red);
    if (threadIdx.x == 0) {
        part[4 * blockIdx.x] = s_quad;
        part[4 * blockIdx.x + 1] = s_lpy;
        part[4 * blockIdx.x + 2] = s_lz2;
        part[4 * blockIdx.x + 3] = s_neg;
    }
}

// part[g][j] = sum over the rows of group g of K[i][j] v_i   (K row-major, ld = mp; one thread per knot, eight rows in
// flight per thread: with 8 CTAs of 128 threads per SM that is 64 KB of loads in flight per SM, what the HBM latency needs)
__global__ void __launch_bounds__(128)
kt_v_kernel(const double *__restrict__ K, int mp, int64_t n, const double *__restrict__ v, double *__restrict__ part)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t i0 = (int64_t)blockIdx.y * per, i1 = min(n, i0 + per);
    double a[8] = {};
    int64_t i = i0;
    for (; i + 7 < i1; i += 8) {
        double k[8];
#pragma unroll
        for (int q = 0; q < 8; q++) k[q] = __ldcs(K + (i + q) * mp + j);
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = fma(k[q], v[i + q], a[q]);
    }
    for (; i < i1; i++) a[0] = fma(K[i * mp + j], v[i], a[0]);
    part[(int64_t)blockIdx.y * mp + j] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

// out[j] = sum_g part[g][j]: 32 columns x 8 group lanes per CTA (launch with 256 threads, mp / 32 CTAs), fixed order
__global__ void __launch_bounds__(256)
sum_groups_kernel(const double *__restrict__ part, int groups, int mp, double *__restrict__ out)
{
    __shared__ double sm[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double s0 = 0.0, s1 = 0.0;
    int g = ty;
    for (; g + 8 < groups; g += 16) {
        s0 += part[(int64_t)g * mp + j];
        s1 += part[(int64_t)(g + 8) * mp + j];
    }
    if (g < groups) s0 += part[(int64_t)g * mp + j];
    sm[ty][tx] = s0 + s1;
    __syncthreads();
    if (ty == 0) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += sm[k][tx];
        out[j] = s;
    }
}

// out1[i] = K[i,:] h1, out2[i] = K[i,:] h2 (h2 / out2 may be null): one warp per row, coalesced over knots
__global__ void __launch_bounds__(256)
k_v_kernel(const double *__restrict__ K, int mp, int m, int64_t n, const double *__restrict__ h1,
           const double *__restrict__ h2, double *__restrict__ out1, double *__restrict__ out2)
{
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (int64_t)gridDim.x * 8) {
        const double *row = K + i * mp;
        double a = 0.0, b = 0.0;
        for (int j = lane; j < m; j += 32) {
            const double k = row[j];
            a = fma(k, h1[j], a);
            if (h2) b = fma(k, h2[j], b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            out1[i] = a;
            if (out2) out2[i] = b;
        }
    }
}

// grad_psi = d1 - rz + Kh / Z ; egp = e * grad_psi ; part[b] = #rows with |grad_psi| > tol (NaN counts)
__global__ void __launch_bounds__(256)
lap_gradpsi_kernel(const double *__restrict__ d1, const double *__restrict__ rz, const double *__restrict__ Kh,
                   const double *__restrict__ Z, const double *__restrict__ e, int64_t n, double tol,
                   double *__restrict__ gpsi, double *__restrict__ egp, double *__restrict__ part)
{
    __shared__ double red[8];
    double cnt = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double g = d1[i] - rz[i] + Kh[i] / Z[i];
        gpsi[i] = g;
        egp[i] = e[i] * g;
        if (!(fabs(g) <= tol)) cnt += 1.0;
    }
    cnt = block_sum_256(cnt, red);
    if (threadIdx.x == 0) part[blockIdx.x] = cnt;
}

// ff <- ff + Z e d1 - e (ff - mu) + e Kh + e Kg      (A11 - A12 + A13 + A2, R/newtrap_sparseGP.R:277-288)
__global__ void lap_update_kernel(double *__restrict__ ff, const double *__restrict__ mu, const double *__restrict__ Z,
                                  const double *__restrict__ e, const double *__restrict__ d1,
                                  const double *__restrict__ Kh, const double *__restrict__ Kg, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double f = ff[i], fm = f - (mu ? mu[i] : 0.0), ei = e[i];
        ff[i] = f + (Z[i] * ei * d1[i] - ei * fm + ei * Kh[i] + ei * Kg[i]);
    }
}

// Z = zconst - q ; invZ = 1 / Z
__global__ void lap_z_kernel(const double *__restrict__ q, int64_t n, double zconst, double *__restrict__ Z,
                             double *__restrict__ invZ)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double z = zconst - q[i];
        Z[i] = z;
        invZ[i] = 1.0 / z;
    }
}

// gradient rows, stage 1: alpha = rz - Kh/Z ; comp4 = -1/D + c/(Z W - 1)^2, D = W - 1/Z ; u = comp4 (-W3) ;
// vec1 = B u / W
__global__ void lap_grad_rows1_kernel(const double *__restrict__ rz, const double *__restrict__ Kh,
                                      const double *__restrict__ Z, const double *__restrict__ Wv,
                                      const double *__restrict__ W3, const double *__restrict__ B,
                                      const double *__restrict__ c, int64_t n, double *__restrict__ alpha,
                                      double *__restrict__ u, double *__restrict__ vec1)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double z = Z[i], w = Wv[i];
        alpha[i] = rz[i] - Kh[i] / z;
        const double Dv = w - 1.0 / z, zw1 = z * w - 1.0;
        const double comp4 = -1.0 / Dv + c[i] / (zw1 * zw1);
        const double ui = comp4 * (-W3[i]);
        u[i] = ui;
        vec1[i] = B[i] * ui / w;
    }
}

// stage 2: t = -u B / W + B Kc2 ; rho = alpha^2/2 - (B - B^2 c)/2 - t g/2 ; rs1 = -B - 2 rho ; negt = -t
__global__ void __launch_bounds__(256)
lap_grad_rows2_kernel(const double *__restrict__ u, const double *__restrict__ B, const double *__restrict__ Wv,
                      const double *__restrict__ Kc2, const double *__restrict__ alpha,
                      const double *__restrict__ c, const double *__restrict__ g, int64_t n, double *__restrict__ t,
                      double *__restrict__ negt, double *__restrict__ rho, double *__restrict__ rs1,
                      double *__restrict__ part)
{
    __shared__ double red[8];
    double sr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double b = B[i], a = alpha[i];
        const double ti = -u[i] * b / Wv[i] + b * Kc2[i];
        const double rh = 0.5 * a * a - 0.5 * (b - b * b * c[i]) - 0.5 * ti * g[i];
        t[i] = ti;
        negt[i] = -ti;
        rho[i] = rh;
        rs1[i] = -b - 2.0 * rh;
        sr += rh;
    }
    sr = block_sum_256(sr, red);
    if (threadIdx.x == 0) part[blockIdx.x] = sr;
}

__global__ void sum_strided2_kernel(const double *__restrict__ part, int count, int stride, int offset,
                                    double *__restrict__ out)
{
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[i * stride + offset];
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) *out = s;
}

// out[k] = sum_i part[i * stride + k], k = blockIdx.x (one launch for all the partial sums of a row kernel)
__global__ void sum_strided_multi_kernel(const double *__restrict__ part, int count, int stride, double *__restrict__ out)
{
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[i * stride + blockIdx.x];
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

constexpr int LROW_BLOCKS = 128;
constexpr int KTV_GROUPS = 296;   // x mp / 128 column blocks: 8 CTAs per SM at m = 512
static_assert(KTV_GROUPS <= PART2_VEC_GROUPS, "part2 is sized by plan() in gauss.cu");

struct Lap {
    srgp_ctx *ctx;
    GaussWS *w;
    cudaStream_t s;
    int mp, m, d;
    int64_t n;
    GenParams gp;
    int family;
    double pois_m;
    // row vectors
    double *ff, *Z, *invZ, *d1, *Wv, *W3, *e, *om, *rz, *Kh, *Kg, *gpsi, *egp, *t0, *t1, *t2, *t3, *t4;
    // matrices / vectors
    double *S, *Sinv, *GZ, *CZ, *A, *Linv, *LinvT, *tmp;
    double *av, *hv, *cv, *g2v, *tv, *gsc;
    bool om_nonneg = true;   // Newton loop: omega >= 0 so far (checked on the device every stage): one slice set sqrt(omega) K

    int ktv(const double *v, double *out)
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        kt_v_kernel<<<dim3(mp / 128, KTV_GROUPS), 128, 0, s>>>(w->Kmat.d(), mp, n, v, w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_groups_kernel<<<mp / 32, 256, 0, s>>>(w->part2.d(), KTV_GROUPS, mp, out);
        SRGP_LAUNCH_CHECK();
        return SRGP_OK;
    }
    int kv(const double *h1, const double *h2, double *o1, double *o2)
    {
        if (n == 0) return SRGP_OK;
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        k_v_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(w->Kmat.d(), mp, m, n, h1, h2, o1, o2);
        SRGP_LAUNCH_CHECK();
        return SRGP_OK;
    }
    // scal4 = [s_quad, log p, sum log Z2, #rows with omega < 0 or NaN]
    int rows(double *W3out, double *scal4)
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        if (family == SRGP_BERNOULLI)
            lap_rows_kernel<SRGP_BERNOULLI><<<LROW_BLOCKS, 256, 0, s>>>(ff, ctx->yp, ctx->mup, Z, n, pois_m, d1, Wv, W3out,
                                                                         e, om, rz, w->part2.d());
        else
            lap_rows_kernel<SRGP_POISSON><<<LROW_BLOCKS, 256, 0, s>>>(ff, ctx->yp, ctx->mup, Z, n, pois_m, d1, Wv, W3out, e,
                                                                       om, rz, w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_strided_multi_kernel<<<4, 256, 0, s>>>(w->part2.d(), LROW_BLOCKS, 4, scal4);
        SRGP_LAUNCH_CHECK();
        return SRGP_OK;
    }
};

// common set-up: plan, knots, K, S (tau^2 kept), S^-1, Z, G_Z, C_Z
static int lap_setup(Lap &L, srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                     const double *l, double tau, double delta, double pois_m)
{
    GaussWS *w = gauss_ws(ctx);
    SRGP_TRY(plan(ctx, w, (int)m, ctx->d));
    L.ctx = ctx;
    L.w = w;
    L.s = ctx->stream;
    L.mp = w->mp;
    L.m = w->m;
    L.d = w->d;
    L.n = ctx->n;
    L.family = family;
    L.pois_m = pois_m;
    cudaStream_t s = L.s;
    SRGP_TRY(upload_knots(w, xu, (size_t)m * ctx->d * 8, s));
    fill_gen(L.gp, kernel, L.d, sigma, l);
    w->k_reuse = true;    // up to four row-form / K*M passes over this K (gauss_i8.cu)
    L.om_nonneg = getenv("SRGP_LAP_TWO_SETS") == nullptr;   // test switch: start with the two-slice-set Gram (lap_stage_checked)
    const int64_t n = L.n;
    double **rv[] = {&L.ff, &L.Z, &L.invZ, &L.d1, &L.Wv, &L.W3, &L.e, &L.om, &L.rz, &L.Kh, &L.Kg, &L.gpsi, &L.egp,
                     &L.t0, &L.t1, &L.t2, &L.t3, &L.t4};
    for (int k = 0; k < 18; k++) *rv[k] = w->rowv(k, n);
    L.S = w->mat(W_::M_S);
    L.Sinv = w->mat(W_::M_SINV);
    L.GZ = w->mat(W_::M_GZ);
    L.CZ = w->mat(W_::M_CZ);
    L.A = w->mat(W_::M_A);
    L.Linv = w->mat(W_::M_LINV);
    L.LinvT = w->mat(W_::M_X1);
    L.tmp = w->mat(W_::M_TMP);
    L.av = w->vec(W_::V_B);
    L.hv = w->vec(W_::V_V);
    L.cv = w->vec(W_::V_GV);
    L.g2v = w->vec(W_::V_TMP);
    L.tv = w->vec(W_::V_T1);
    L.gsc = w->gemv_scratch();
    const int mp = L.mp;
    const size_t mm = (size_t)mp * mp;
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W_::S_INFO, 0, 16, s));
    SRGP_TRY(coin_reset(ctx, w));
    SRGP_TRY(materialise_k(ctx, w, L.gp));
    // S keeps tau^2 + delta on its diagonal in the Laplace models (R/newtrap_sparseGP.R:51-60)
    SRGP_TRY(assemble_dev_ld(ctx, s, kernel, w->U.d(), m, L.d, sigma, l, tau * tau + delta, L.S, mp));
    SRGP_TRY(dense::pad_identity(ctx, s, L.S, mp, (int)m, 1.0));
    SRGP_CUDA(cudaMemcpyAsync(w->mat(W_::M_T1), L.S, mm * 8, cudaMemcpyDeviceToDevice, s));
    SRGP_TRY(dense::chol_inverse(ctx, s, w->mat(W_::M_T1), mp, (int)m, w->dinv(0), L.Linv, L.LinvT, L.tmp, L.Sinv,
                                 w->info(0), w->sc(W_::S_LOGDET_S)));
    // Z_i = sigma^2 + tau^2 + delta - K_i S^-1 K_i^T (R/newtrap_sparseGP.R:62-66)
    SRGP_TRY(gauss_rowform(ctx, w, L.gp, L.Sinv, nullptr, L.t0, nullptr));
    if (n > 0) {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        lap_z_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(L.t0, n, sigma * sigma + tau * tau + delta, L.Z, L.invZ);
        SRGP_LAUNCH_CHECK();
    }
    // G_Z = K^T diag(1/Z) K ; C_Z = (S + G_Z)^-1 (the R of grad_loglik_fn / update / objective)
    double *buf = w->red1.d();
    // (INT8 tensor-core Gram over regenerated K, as every Gram of the Gaussian models: 1 / Z > 0, so one slice set
    // sqrt(1 / Z) K serves both operands; its K^T r by-product is not needed here)
    SRGP_TRY(gauss_pass1(ctx, w, L.gp, L.invZ, L.invZ, buf, w->vec(W_::V_T1), true));
    SRGP_TRY(comm_allreduce(ctx, buf, mm, s));
    SRGP_CUDA(cudaMemcpyAsync(L.GZ, buf, mm * 8, cudaMemcpyDeviceToDevice, s));
    SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, L.S, 1.0, L.GZ, 0.0, L.A));
    SRGP_TRY(dense::chol_inverse(ctx, s, L.A, mp, (int)m, w->dinv(1), L.Linv, L.LinvT, L.tmp, L.CZ, w->info(1),
                                 w->sc(W_::S_LOGDET_A)));
    return SRGP_OK;
}

// objective pieces at the current ff: rows, a = K^T rz, G_omega; allreduce; factor S + G_omega.
// On return: red1 = [G_omega | a | s_quad, log p, sum log Z2, #omega<0, n], Linv / LinvT of S + G_omega, S_X+0..4 scalars,
// S_LOGDET_A = log|S + G_omega|, S_BV = a^T C_Z a.
static int lap_objective_stage(Lap &L, double *W3out)
{
    GaussWS *w = L.w;
    cudaStream_t s = L.s;
    const int mp = L.mp, m = L.m;
    const size_t mm = (size_t)mp * mp;
    double *buf = w->red1.d(), *a = buf + mm, *tail = a + mp;
    SRGP_TRY(L.rows(W3out, tail));
    SRGP_TRY(set_scalar(L.ctx, tail + 4, (double)L.n));
    // G_omega = K^T diag(omega) K and a = K^T rz in one INT8 pass.  omega = -W / (1 - Z W) >= 0 for the reference's
    // likelihoods (W < 0 also with quirk Q1, Z > 0), so one slice set sqrt(omega) K serves both operands -- half the
    // generator output and half the launches of the two-set form (omega K, K).  Nothing guarantees the sign for arbitrary
    // y, so the row kernel counts the rows with omega < 0 (or NaN), the count travels with the allreduce, and
    // lap_stage_checked() repeats the stage with two slice sets if it is not zero.
    w->pass1_kmat = w->Kmat.d();     // K is materialised for the matrix-vector products: the generator reads it (WMODE 3)
    const int rc1 = gauss_pass1(L.ctx, w, L.gp, L.om, L.rz, buf, a, L.om_nonneg);
    w->pass1_kmat = nullptr;
    SRGP_TRY(rc1);
    SRGP_TRY(comm_allreduce(L.ctx, buf, mm + mp + 5, s));
    SRGP_TRY(copy_scalar(L.ctx, w->sc(W_::S_X), tail, 5));
    SRGP_CUDA(cudaMemcpyAsync(L.av, a, (size_t)mp * 8, cudaMemcpyDeviceToDevice, s));
    // h = C_Z a ; a^T h
    SRGP_TRY(dense::gemv_t(L.ctx, s, mp, 1.0, L.CZ, L.av, 0.0, nullptr, L.hv));      // C_Z is symmetric
    SRGP_TRY(dense::dot_v(L.ctx, s, m, L.av, L.hv, w->sc(W_::S_BV)));
    // factor S + G_omega
    SRGP_TRY(dense::axpby(L.ctx, s, mp, m, 1.0, L.S, 1.0, buf, 0.0, L.A));
    SRGP_TRY(dense::potrf(L.ctx, s, L.A, mp, m, w->dinv(1), w->info(1), w->sc(W_::S_LOGDET_A)));
    SRGP_TRY(dense::trtri(L.ctx, s, L.A, mp, w->dinv(1), L.Linv, L.LinvT, L.tmp));
    return SRGP_OK;
}

static double lap_objective_value(const double *h)
{
    // -quad/2 + a^T C_Z a/2 + log p(y|ff) - (-log|S| + log|S + G_omega|)/2 - sum log(1 - W Z)/2
    return -0.5 * h[W_::S_X] + 0.5 * h[W_::S_BV] + h[W_::S_X + 1] - 0.5 * (-h[W_::S_LOGDET_S] + h[W_::S_LOGDET_A]) -
           0.5 * h[W_::S_X + 2];
}

// One objective stage + the scalars on the host.  A negative (or NaN) omega found while the one-slice-set Gram was
// in use invalidates that Gram (and possibly its Cholesky flag): every rank sees the same summed count and repeats the
// stage with two slice sets, which it keeps for the rest of the call.
static int lap_stage_checked(Lap &L)
{
    SRGP_TRY(lap_objective_stage(L, nullptr));
    int rc = fetch_scalars(L.ctx, L.w);
    if (L.om_nonneg && (rc == SRGP_OK || rc == SRGP_ERR_NOT_PD) && !(L.w->h_scal[W_::S_X + 3] == 0.0)) {
        L.om_nonneg = false;
        SRGP_CUDA(cudaMemsetAsync(L.w->info(1), 0, sizeof(int), L.s));
        SRGP_TRY(lap_objective_stage(L, nullptr));
        rc = fetch_scalars(L.ctx, L.w);
    }
    return rc;
}

}  // namespace srgp

using namespace srgp;

static int lap_check(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, const double *l)
{
    if (!ctx || !xu || !l || m <= 0 || m > 32768) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("Laplace entry point called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    if (family != SRGP_BERNOULLI && family != SRGP_POISSON) {
        set_error("unknown family %d", family);
        return SRGP_ERR_ARG;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    return SRGP_OK;
}

extern "C" int srgp_laplace_newton(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m,
                                   const double *muu, double sigma, const double *l, double tau, double delta,
                                   double pois_m, int maxit, double tol, double *ff, double *obj_hist, int *n_iter,
                                   double *grad_psi, double *u_mean, double *u_var)
{
    SRGP_TRY(lap_check(ctx, family, kernel, xu, m, l));
    if (!ff || !obj_hist || !n_iter || maxit < 2) {
        set_error("bad argument (ff, obj_hist, n_iter must be given; maxit >= 2)");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    Lap L;
    SRGP_TRY(lap_setup(L, ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m));
    GaussWS *w = L.w;
    cudaStream_t s = L.s;
    const int mp = L.mp;
    const size_t mm = (size_t)mp * mp;
    const int64_t n = L.n;
    double *Gw = w->red1.d(), *GwPrev = w->mat(W_::M_GWP);
    SRGP_CUDA(cudaMemcpyAsync(L.ff, ff, (size_t)n * 8, cudaMemcpyHostToDevice, s));

    SRGP_TRY(lap_stage_checked(L));
    obj_hist[0] = lap_objective_value(w->h_scal);
    int it = 1;
    while (true) {
        it++;
        // ---- update (R/newtrap_sparseGP.R:84-92,105-112): grad_psi, then ff <- ff + A11 - A12 + A13 + A2 ----
        SRGP_TRY(L.kv(L.hv, nullptr, L.Kh, nullptr));
        {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            lap_gradpsi_kernel<<<LROW_BLOCKS, 256, 0, s>>>(L.d1, L.rz, L.Kh, L.Z, L.e, n, tol, L.gpsi, L.egp, w->part2.d());
            SRGP_LAUNCH_CHECK();
        }
        double *cbuf = w->vec(W_::V_T2);   // [c (mp) | count]  -- V_T2 and V_T3 are adjacent
        {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            sum_strided2_kernel<<<1, 256, 0, s>>>(w->part2.d(), LROW_BLOCKS, 1, 0, cbuf + mp);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_TRY(L.ktv(L.egp, cbuf));      // c = K^T (e grad_psi); reuses part2 as scratch, the count is already out
        SRGP_TRY(comm_allreduce(ctx, cbuf, mp + 1, s));
        SRGP_TRY(copy_scalar(ctx, w->sc(W_::S_X + 8), cbuf + mp, 1));
        // g2 = (S + G_omega)^-1 c = L^-T (L^-1 c)
        SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, L.LinvT, cbuf, 0.0, nullptr, L.tv));
        SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, L.Linv, L.tv, 0.0, nullptr, L.g2v));
        SRGP_TRY(L.kv(L.g2v, nullptr, L.Kg, nullptr));
        SRGP_CUDA(cudaMemcpyAsync(GwPrev, Gw, mm * 8, cudaMemcpyDeviceToDevice, s));
        if (n > 0) {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            lap_update_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(L.ff, ctx->mup, L.Z, L.e, L.d1, L.Kh, L.Kg, n);
            SRGP_LAUNCH_CHECK();
        }
        // ---- objective at the new ff (its Gram and factor serve the next update) ----
        SRGP_TRY(lap_stage_checked(L));
        obj_hist[it - 1] = lap_objective_value(w->h_scal);
        const bool grad_big = w->h_scal[W_::S_X + 8] > 0.0;
        const double dobj = fabs(obj_hist[it - 1] - obj_hist[it - 2]);
        if (!(it < maxit && (dobj > tol || !(dobj == dobj) || grad_big))) break;   // R/newtrap_sparseGP.R:100
    }
    *n_iter = it;
    SRGP_CUDA(cudaMemcpyAsync(ff, L.ff, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    if (grad_psi) SRGP_CUDA(cudaMemcpyAsync(grad_psi, L.gpsi, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    if (u_mean) {
        // u_mean = muu + a - G_Z C_Z a with a = K^T ((ff - mu)/Z) at the final ff (R/newtrap_sparseGP.R:171-173)
        SRGP_TRY(dense::gemv(ctx, s, mp, -1.0, L.GZ, L.hv, 1.0, L.av, L.tv, L.gsc));
        std::vector<double> um((size_t)mp);
        SRGP_CUDA(cudaMemcpyAsync(um.data(), L.tv, (size_t)mp * 8, cudaMemcpyDeviceToHost, s));
        SRGP_CUDA(cudaStreamSynchronize(s));
        for (int64_t j = 0; j < m; j++) u_mean[j] = (muu ? muu[j] : 0.0) + um[j];
    }
    if (u_var) {
        // u_var = S + TT + TT (S - TT)^-1 TT with TT = -G_omega of the LAST update (R/newtrap_sparseGP.R:159-176)
        double *Cw = w->mat(W_::M_C), *T1 = w->mat(W_::M_T1), *T2 = w->mat(W_::M_T2);
        SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, L.S, 1.0, GwPrev, 0.0, L.A));
        SRGP_TRY(dense::chol_inverse(ctx, s, L.A, mp, (int)m, w->dinv(1), L.Linv, L.LinvT, L.tmp, Cw, w->info(1),
                                     w->sc(W_::S_LOGDET_A)));
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, GwPrev, mp, Cw, mp, 0.0, T1, mp));
        SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, T1, mp, GwPrev, mp, 0.0, T2, mp));
        SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, L.S, -1.0, GwPrev, 0.0, T1));
        SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, T1, 1.0, T2, 0.0, T1));
        SRGP_CUDA(cudaMemcpy2DAsync(u_var, (size_t)m * 8, T1, (size_t)mp * 8, (size_t)m * 8, m, cudaMemcpyDeviceToHost, s));
    }
    SRGP_TRY(fetch_scalars(ctx, w));
    return SRGP_OK;
}

// dlogq_dcov_par on the resident shard; knot_grad (host, m * d, knot-major; null = none) adds the knot-location
// gradient, which reuses Omega (the two pass-2 launches collect its per-knot column sums) and N.
// om_nonneg: G_B = K^T diag(omega) K from one slice set sqrt(omega) K (as in the Newton loop, lap_objective_stage); the
// rows with omega < 0 are counted on the device and *retry is set when there were any (the caller repeats with two sets).
static int laplace_grad_body(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                             const double *l, double tau, double delta, double pois_m, const double *ff, double *grad,
                             const double *knot_lb, const double *knot_ub, double *knot_grad, bool om_nonneg, bool *retry)
{
    SRGP_TRY(lap_check(ctx, family, kernel, xu, m, l));
    if (!ff || !grad) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    Lap L;
    SRGP_TRY(lap_setup(L, ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m));
    GaussWS *w = L.w;
    cudaStream_t s = L.s;
    const int mp = L.mp, d = L.d;
    const size_t mm = (size_t)mp * mp;
    const int64_t n = L.n;
    SRGP_CUDA(cudaMemcpyAsync(L.ff, ff, (size_t)n * 8, cudaMemcpyHostToDevice, s));

    // B = 1/(Z - 1/W) = omega ; G_B, a = K^T rz, K^T g in one allreduce (R/laplace_approx_gradient.R:127-155)
    double *GB = w->red1.d(), *a = GB + mm, *ktg = a + mp, *tail = ktg + mp;
    SRGP_TRY(L.rows(L.W3, tail));
    SRGP_TRY(L.ktv(L.d1, ktg));
    w->pass1_kmat = w->Kmat.d();
    const int rc1 = gauss_pass1(ctx, w, L.gp, L.om, L.rz, GB, a, om_nonneg);
    w->pass1_kmat = nullptr;
    SRGP_TRY(rc1);
    SRGP_TRY(comm_allreduce(ctx, GB, mm + 2 * mp + 4, s));
    SRGP_TRY(copy_scalar(ctx, w->sc(W_::S_X + 3), tail + 3, 1));      // rows with omega < 0 or NaN, over all ranks
    double *C = w->mat(W_::M_C), *M2 = w->mat(W_::M_CGS), *SG = w->mat(W_::M_SG);
    double *SGS = w->mat(W_::M_SGS), *N = w->mat(W_::M_N), *Grho = w->mat(W_::M_X2);
    double *beta = w->vec(W_::V_BETA), *GG = w->vec(W_::V_T4), *c2 = w->vec(W_::V_T5), *Cc2 = w->vec(W_::V_T6);
    double *skt = w->vec(W_::V_T7);
    SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, L.S, 1.0, GB, 0.0, L.A));
    SRGP_TRY(dense::chol_inverse(ctx, s, L.A, mp, (int)m, w->dinv(1), L.Linv, L.LinvT, L.tmp, C, w->info(1),
                                 w->sc(W_::S_LOGDET_A)));
    // h = C_Z a ; alpha = rz - K h / Z ; beta = S^-1 (a - G_Z h) ; GG = S^-1 K^T g
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, L.CZ, a, 0.0, nullptr, L.hv, L.gsc));
    SRGP_TRY(L.kv(L.hv, nullptr, L.Kh, nullptr));
    SRGP_TRY(dense::gemv(ctx, s, mp, -1.0, L.GZ, L.hv, 1.0, a, L.tv, L.gsc));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, L.Sinv, L.tv, 0.0, nullptr, beta, L.gsc));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, L.Sinv, ktg, 0.0, nullptr, GG, L.gsc));
    // c_i = K_i C K_i^T (row forms), comp4, u, vec1 = B u / W
    double *cq = L.t0, *alpha = L.t1, *u = L.t2, *vec1 = L.t3, *rho = L.t4;
    double *tt = L.gpsi, *negt = L.egp, *rs1 = L.Kg;
    SRGP_TRY(gauss_rowform(ctx, w, L.gp, C, nullptr, cq, nullptr));
    if (n > 0) {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        lap_grad_rows1_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(L.rz, L.Kh, L.Z, L.Wv, L.W3, L.om, cq, n, alpha, u, vec1);
        SRGP_LAUNCH_CHECK();
    }
    // t = -u B / W + B K C K^T (B u / W)   (R/laplace_approx_gradient.R:306-307 contracted with comp4 (-W3))
    SRGP_TRY(L.ktv(vec1, c2));
    SRGP_TRY(comm_allreduce(ctx, c2, mp, s));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, C, c2, 0.0, nullptr, Cc2, L.gsc));
    SRGP_TRY(L.kv(Cc2, nullptr, L.Kh, nullptr));   // Kh is free now: holds K C c2
    double *p2 = w->sc(W_::S_P2);
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        lap_grad_rows2_kernel<<<LROW_BLOCKS, 256, 0, s>>>(u, L.om, L.Wv, L.Kh, alpha, cq, L.d1, n, tt, negt, rho, rs1,
                                                          w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_strided2_kernel<<<1, 256, 0, s>>>(w->part2.d(), LROW_BLOCKS, 1, 0, p2 + d + 2);
        SRGP_LAUNCH_CHECK();
    }
    // M2 = C G_B S^-1 = S^-1 - C (C (S + G_B) = I) ; Omega = diag(-B - 2 rho) K S^-1 + diag(B) K M2 + alpha beta^T - t GG^T
    SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, L.Sinv, -1.0, C, 0.0, M2));
    w->want_knots = knot_grad != nullptr;
    w->knot_transform = knot_grad && knot_lb && knot_ub;
    if (w->knot_transform)
        for (int c = 0; c < d; c++) {
            w->knot_lb[c] = knot_lb[c];
            w->knot_ub[c] = knot_ub[c];
        }
    int rc2 = gauss_pass2(ctx, w, L.gp, L.Sinv, rs1, alpha, beta, nullptr, false);
    if (rc2 == SRGP_OK) rc2 = gauss_pass2(ctx, w, L.gp, M2, L.om, negt, GG, p2, true);
    w->want_knots = false;
    SRGP_TRY(rc2);
    SRGP_TRY(coin_fix(ctx, w, L.gp, L.Sinv, 0.0, p2 + 1 + d));
    // G_rho, K^T t -> one allreduce with the gradient partials
    double *red2 = w->mat(W_::M_T2);   // [G_rho (mm)] then M_X1.. is LinvT: use a separate tail buffer
    double *tail2 = w->vec(W_::V_T2);  // [K^T t (mp) | p2 (d + 4)]  (V_T2, V_T3 adjacent: 2 mp >= mp + d + 4)
    SRGP_TRY(gauss_pass1(ctx, w, L.gp, rho, tt, red2, tail2));     // G_rho and K^T t
    SRGP_TRY(copy_scalar(ctx, tail2 + mp, p2, W_::p2_len(d)));
    SRGP_TRY(comm_allreduce(ctx, red2, mm, s));
    SRGP_TRY(comm_allreduce(ctx, tail2, mp + W_::p2_len(d), s));
    SRGP_TRY(copy_scalar(ctx, p2, tail2 + mp, W_::p2_len(d)));
    SRGP_CUDA(cudaMemcpyAsync(Grho, red2, mm * 8, cudaMemcpyDeviceToDevice, s));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, L.Sinv, tail2, 0.0, nullptr, skt, L.gsc));
    // N = S^-1 G_B S^-1/2 - S^-1 G_B M2/2 - beta beta^T/2 + S^-1 G_rho S^-1 + (S^-1 K^T t) GG^T/2, and
    // S^-1 G_B (S^-1 - M2) = S^-1 G_B C = S^-1 - C = M2, so the first two terms are M2/2
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, L.Sinv, mp, Grho, mp, 0.0, SG, mp));
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'T', mp, mp, mp, 1.0, SG, mp, L.Sinv, mp, 0.0, SGS, mp));
    SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 0.5, M2, 1.0, SGS, 0.0, N));
    SRGP_TRY(dense::ger(ctx, s, mp, -0.5, beta, beta, N));
    SRGP_TRY(dense::ger(ctx, s, mp, 0.5, skt, GG, N));
    SRGP_TRY(ns_reduce(ctx, w, L.gp, N, L.S, tau * tau + delta, w->sc(W_::S_NS), s));
    if (knot_grad) SRGP_TRY(knot_finish(ctx, w, L.gp, N, L.S));
    {
        const int rcf = fetch_scalars(ctx, w);
        if (rcf == SRGP_ERR_CUDA) return rcf;
        if (om_nonneg && !(w->h_scal[W_::S_X + 3] == 0.0)) {
            *retry = true;
            return SRGP_OK;
        }
        SRGP_TRY(rcf);
    }
    SRGP_TRY(coin_check(w));
    if (knot_grad) {
        SRGP_CUDA(cudaMemcpyAsync(knot_grad, w->knotsum.d() + (int64_t)d * mp, (size_t)m * d * 8, cudaMemcpyDeviceToHost, s));
        SRGP_CUDA(cudaStreamSynchronize(s));
    }

    const double *h = w->h_scal, *hp2 = h + W_::S_P2, *ns = h + W_::S_NS;
    const double sum_rho = hp2[d + 2];
    grad[0] = 2.0 * hp2[0] + 2.0 * ns[0] + 2.0 * sigma * sigma * sum_rho;
    if (kernel == SRGP_ARD) {
        for (int c = 0; c < d; c++) grad[1 + c] = hp2[1 + c] + ns[1 + c];
    } else {
        double g = 0.0;
        for (int c = 0; c < d; c++) g += hp2[1 + c] + ns[1 + c];
        grad[1] = g;
    }
    const int ti = (kernel == SRGP_ARD) ? 1 + d : 2;
    // dS(tau) = 2 tau^2 on identical knot pairs is NOT zeroed in the Laplace gradient (R/laplace_approx_gradient.R:214-237)
    grad[ti] = 2.0 * tau * tau * (sum_rho + ns[1 + d] + hp2[1 + d]);
    return SRGP_OK;
}

static int laplace_grad_impl(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                             const double *l, double tau, double delta, double pois_m, const double *ff, double *grad,
                             const double *knot_lb, const double *knot_ub, double *knot_grad)
{
    bool retry = false;
    int rc = laplace_grad_body(ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m, ff, grad, knot_lb, knot_ub, knot_grad,
                               getenv("SRGP_LAP_TWO_SETS") == nullptr, &retry);
    if (rc == SRGP_OK && retry)
        rc = laplace_grad_body(ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m, ff, grad, knot_lb, knot_ub, knot_grad,
                               false, &retry);
    return rc;
}

extern "C" int srgp_laplace_grad(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                                 const double *l, double tau, double delta, double pois_m, const double *ff,
                                 double *grad)
{
    return laplace_grad_impl(ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m, ff, grad, nullptr, nullptr,
                             nullptr);
}

extern "C" int srgp_laplace_grad_knots(srgp_ctx *ctx, int family, int kernel, const double *xu, int64_t m, double sigma,
                                       const double *l, double tau, double delta, double pois_m, const double *ff,
                                       const double *knot_lb, const double *knot_ub, const int *knot_opt,
                                       int64_t n_opt, double *grad, double *knot_grad, double *trans_knot)
{
    if (!knot_grad || (knot_lb == nullptr) != (knot_ub == nullptr) || n_opt < 0 || (n_opt > 0 && !knot_opt)) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (knot_opt)
        for (int64_t t = 0; t < n_opt; t++)
            if (knot_opt[t] < 0 || knot_opt[t] >= m) {
                set_error("knot_opt[%lld] = %d outside [0, %lld)", (long long)t, knot_opt[t], (long long)m);
                return SRGP_ERR_ARG;
            }
    SRGP_TRY(laplace_grad_impl(ctx, family, kernel, xu, m, sigma, l, tau, delta, pois_m, ff, grad, knot_lb, knot_ub,
                               knot_grad));
    const int d = ctx->d;
    if (knot_opt) {
        std::vector<char> keep((size_t)m, 0);
        for (int64_t t = 0; t < n_opt; t++) keep[knot_opt[t]] = 1;
        for (int64_t k = 0; k < m; k++)
            if (!keep[k])
                for (int c = 0; c < d; c++) knot_grad[k * d + c] = 0.0;
    }
    if (trans_knot)
        for (int c = 0; c < d; c++)
            for (int64_t k = 0; k < m; k++) {
                const double u = xu[k + m * c];
                trans_knot[k + m * c] = knot_lb ? log((u - knot_lb[c]) + 1e-4) - log((knot_ub[c] - u) + 1e-4) : u;
            }
    return SRGP_OK;
}

// ====================================================================================================
// Objective from MATERIALISED matrices: the bodies of obj_fun_norm (R/laplace_approx_obj_funs.R:6-52) and of elbo_fun
// without its trace term (R/vi_functions.R:64-121; the R patch adds trace_term_fun's value, r/patches.R):
//   -1/2 r^T Z^-1 r + 1/2 b^T (S22 + G)^-1 b - 1/2 (sum log Z - log|S22| + log|S22 + G|) - n/2 log 2 pi,
//   G = S12^T diag(1/Z) S12, b = S12^T (r / Z), r = y - mu.
// The callers that stay in R (norm_grad_ascent_vi :755,910,1114, the knot proposal loops) hand over the matrices they
// built with make_cov_mat*C; here they are uploaded, S12 is laid out row-major [n][mp] like the Laplace path's K, and
// the Gram runs on the same DMMA kernel.  log|S22| comes from the Cholesky factor (the reference's log(det()) is the
// same number until det() under/overflows, quirk Q6).  The resident data shard of the context is dropped.
// ====================================================================================================
namespace srgp {

// Kr[(r0 + i) * mp + j] = S12c[i + j * rows]  (one chunk of rows; columns j >= m stay zero)
__global__ void __launch_bounds__(256)
transpose_chunk_kernel(const double *__restrict__ src, int rows, int m, int mp, double *__restrict__ dst)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += 8) {
        const int i = i0 + threadIdx.x, j = j0 + jj;
        tile[jj][threadIdx.x] = (i < rows && j < m) ? src[i + (int64_t)j * rows] : 0.0;
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += 8) {
        const int i = i0 + ii, j = j0 + threadIdx.x;
        if (i < rows && j < mp) dst[(int64_t)i * mp + j] = tile[threadIdx.x][ii];
    }
}

// invZ, rz = (y - mu) / Z (Z, mu recycled like R vectors of length 1 or n); part[2 b + {0, 1}] = sum r^2 / Z, sum log Z
__global__ void __launch_bounds__(256)
mats_rows_kernel(const double *__restrict__ y, const double *__restrict__ mu, int64_t nmu, const double *__restrict__ Z,
                 int64_t nz, int64_t n, double *__restrict__ invZ, double *__restrict__ rz, double *__restrict__ part)
{
    __shared__ double red[8];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double z = Z[nz == 1 ? 0 : i], r = y[i] - (nmu == 0 ? 0.0 : mu[nmu == 1 ? 0 : i]);
        invZ[i] = 1.0 / z;
        rz[i] = r / z;
        s0 = fma(r, r / z, s0);
        s1 += log(z);
    }
    s0 = block_sum_256(s0, red);
    s1 = block_sum_256(s1, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = s0;
        part[2 * blockIdx.x + 1] = s1;
    }
}

}  // namespace srgp

extern "C" int srgp_gauss_obj_mats(srgp_ctx *ctx, const double *Sigma12, int64_t n, int64_t m, const double *Sigma22,
                                   const double *Z, int64_t nz, const double *y, const double *mu, int64_t nmu,
                                   double *obj)
{
    if (!ctx || !Sigma12 || !Sigma22 || !Z || !y || !obj || n <= 0 || m <= 0 || m > 32768 || (nz != 1 && nz != n) ||
        (nmu != 0 && nmu != 1 && nmu != n) || (nmu > 0 && !mu)) {
        set_error("bad argument (Z must have length 1 or n, mu length 0, 1 or n)");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    cudaStream_t s = ctx->stream;
    GaussWS *w = gauss_ws(ctx);
    SRGP_TRY(plan(ctx, w, (int)m, 1));
    ctx->have_data = false;                        // the row buffers below replace the resident shard
    ctx->n = n;
    const int mp = w->mp;
    const size_t mm = (size_t)mp * mp;
    SRGP_TRY(w->rowa.reserve(GaussWS::row_stride(n) * 8 * GaussWS::NROWV));
    SRGP_CUDA(cudaMemsetAsync(w->rowa.p, 0, GaussWS::row_stride(n) * 8 * GaussWS::NROWV, s));
    SRGP_TRY(w->scal.reserve(GaussWS::NSCAL * 8));
    SRGP_TRY(w->part2.reserve((size_t)std::max(256, KTV_GROUPS * mp) * 8));
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W_::S_INFO, 0, 16, s));
    double *invZ = w->rowv(0, n), *rz = w->rowv(1, n);
    // y, mu, Z -> device
    SRGP_TRY(ctx->y.reserve((size_t)n * 8));
    SRGP_TRY(ctx->mu.reserve((size_t)std::max<int64_t>(nmu, 1) * 8));
    SRGP_TRY(ctx->tmp1.reserve((size_t)nz * 8));
    SRGP_CUDA(cudaMemcpyAsync(ctx->y.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    if (nmu > 0) SRGP_CUDA(cudaMemcpyAsync(ctx->mu.p, mu, (size_t)nmu * 8, cudaMemcpyHostToDevice, s));
    SRGP_CUDA(cudaMemcpyAsync(ctx->tmp1.p, Z, (size_t)nz * 8, cudaMemcpyHostToDevice, s));
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 3);
        mats_rows_kernel<<<LROW_BLOCKS, 256, 0, s>>>(ctx->y.d(), ctx->mu.d(), nmu, ctx->tmp1.d(), nz, n, invZ, rz, w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_strided2_kernel<<<1, 256, 0, s>>>(w->part2.d(), LROW_BLOCKS, 2, 0, w->sc(W_::S_X));
        SRGP_LAUNCH_CHECK();
        sum_strided2_kernel<<<1, 256, 0, s>>>(w->part2.d(), LROW_BLOCKS, 2, 1, w->sc(W_::S_X + 1));
        SRGP_LAUNCH_CHECK();
    }
    // Sigma22 -> S (identity on the padding), log|S| from its Cholesky factor
    double *S = w->mat(W_::M_S), *A = w->mat(W_::M_A), *T1 = w->mat(W_::M_T1);
    SRGP_CUDA(cudaMemsetAsync(S, 0, mm * 8, s));
    SRGP_CUDA(cudaMemcpy2DAsync(S, (size_t)mp * 8, Sigma22, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice, s));
    SRGP_TRY(dense::pad_identity(ctx, s, S, mp, (int)m, 1.0));
    SRGP_CUDA(cudaMemcpyAsync(T1, S, mm * 8, cudaMemcpyDeviceToDevice, s));
    SRGP_TRY(dense::potrf(ctx, s, T1, mp, (int)m, w->dinv(0), w->info(0), w->sc(W_::S_LOGDET_S)));
    // Sigma12 (column-major n x m on the host) -> K row-major [rows][mp], zero rows up to the Gram kernel's quantum
    const int quantum = gemm::BK * w->splits;
    const int64_t rows_alloc = round_up(n, quantum);
    SRGP_TRY(w->Kmat.reserve((size_t)rows_alloc * mp * 8));
    SRGP_CUDA(cudaMemsetAsync(w->Kmat.p, 0, (size_t)rows_alloc * mp * 8, s));
    const int crows = 16384;
    SRGP_TRY(ctx->out_mat.reserve((size_t)crows * m * 8));
    for (int64_t r0 = 0; r0 < n; r0 += crows) {
        const int rows = (int)std::min<int64_t>(crows, n - r0);
        SRGP_CUDA(cudaMemcpy2DAsync(ctx->out_mat.p, (size_t)rows * 8, Sigma12 + r0, (size_t)n * 8, (size_t)rows * 8, m,
                                    cudaMemcpyHostToDevice, s));
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        transpose_chunk_kernel<<<dim3(ceil_div(rows, 32), ceil_div(mp, 32)), dim3(32, 8), 0, s>>>(
            ctx->out_mat.d(), rows, (int)m, mp, w->Kmat.d() + (size_t)r0 * mp);
        SRGP_LAUNCH_CHECK();
    }
    // G = K^T diag(1/Z) K, b = K^T (r / Z); A = S + G; b^T A^-1 b = |L^-1 b|^2; log|A|
    double *buf = w->red1.d(), *bv = w->vec(W_::V_B), *t1 = w->vec(W_::V_T1);
    SRGP_TRY(gram_materialised(ctx, w, invZ, buf));
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        kt_v_kernel<<<dim3(mp / 128, KTV_GROUPS), 128, 0, s>>>(w->Kmat.d(), mp, n, rz, w->part2.d());
        SRGP_LAUNCH_CHECK();
        sum_groups_kernel<<<mp / 32, 256, 0, s>>>(w->part2.d(), KTV_GROUPS, mp, bv);
        SRGP_LAUNCH_CHECK();
    }
    SRGP_TRY(dense::axpby(ctx, s, mp, (int)m, 1.0, S, 1.0, buf, 0.0, A));
    SRGP_TRY(dense::chol_inverse(ctx, s, A, mp, (int)m, w->dinv(1), w->mat(W_::M_LINV), w->mat(W_::M_X1), w->mat(W_::M_TMP),
                                 w->mat(W_::M_C), w->info(1), w->sc(W_::S_LOGDET_A)));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, w->mat(W_::M_LINV), bv, 0.0, nullptr, t1, w->gemv_scratch()));
    SRGP_TRY(dense::dot_v(ctx, s, (int)m, t1, t1, w->sc(W_::S_BV)));
    SRGP_TRY(fetch_scalars(ctx, w));
    const double *h = w->h_scal;
    *obj = -0.5 * h[W_::S_X] + 0.5 * h[W_::S_BV] - 0.5 * (h[W_::S_X + 1] - h[W_::S_LOGDET_S] + h[W_::S_LOGDET_A]) -
           0.5 * (double)n * log(2.0 * M_PI);
    ctx->n = 0;
    return SRGP_OK;
}
