// gauss_oat.cu -- OAT candidate scoring.
//
// Reference: knot_prop_random_norm_vi (R/vi_functions.R:2108-2304) and knot_prop_random_norm
// (R/knot_proposal_functions.R:1176-1357): for each of TTmax candidate rows c the objective is re-evaluated with
// the knots [U; c] -- TTmax full (m+1)-knot rebuilds of Sigma12, Sigma22 and both Cholesky factors.
//
// VI (elbo_fun, R/vi_functions.R:64-121): Z does not depend on the knots, so appending ONE knot borders S, G1, b1:
//     S+  = [S s; s^T kap]      s = k(U, c), kap = sigma^2 + delta
//     G1+ = [G1 g; g^T gam]     g = K^T k_c, gam = k_c^T k_c, k_c = k(X, c);   b1+ = [b1; k_c^T r]
// One Gram of [knots | candidates] over the data rows (the ordinary pass 1 with m + T knots, one allreduce) gives
// G1, g_t, gam_t, k_ct^T r for every candidate at once; with A = S + B G1, a = s + B g, alp = kap + B gam:
//     log|S+| = log|S| + log sS,   sS = kap - |L_S^-1 s|^2
//     log|A+| = log|A| + log sA,   sA = alp - |L_A^-1 a|^2
//     b+^T A+^-1 b+ = b^T A^-1 b + (B k_c^T r - (L_A^-1 a)^T (L_A^-1 b))^2 / sA
//     tr(S+^-1 G1+) = tr(S^-1 G1) + (w^T G1 w - 2 w^T g + gam) / sS,   w = S^-1 s
// i.e. five m x m by m x T products and T column reductions.  A Schur complement that is not positive beyond its
// rounding noise is the bordered form of R's solve() / chol() error (the reference resamples that candidate): the
// score comes back as NaN.
// oracle/reduced_model.py:vi_oat_scores is the NumPy statement; tests compare with the literal per-candidate loop.
//
// FIC (obj_fun_norm with Z_i = sigma^2 + tau^2 + delta - q_i): every Z_i changes with the candidate, the weighted
// Gram K^T diag(1/Z) K changes in full rank, so candidates are scored by objective-only evaluations on the
// resident shard (no re-upload, K never materialised).
#include <math.h>

#include <vector>

#include "dense.cuh"
#include "gauss.cuh"

namespace srgp {

using W = GaussWS;
constexpr int OAT_T = 128;   // candidates per device batch = one column panel of the tile engine

// sT = S_all[0:m, m:m+T], gT = G_all[0:m, m:m+T] as zero-padded mp x 128 panels; per-candidate scalars
// sc[0][t] = kap_t, sc[1][t] = gam_t, sc[2][t] = k_ct^T r
__global__ void oat_extract_kernel(const double *__restrict__ S_all, const double *__restrict__ G_all,
                                   const double *__restrict__ b_all, int mp, int m, int T, double *__restrict__ sT,
                                   double *__restrict__ gT, double *__restrict__ sc)
{
    const int t = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mp) return;
    const bool v = (i < m && t < T);
    const int64_t src = i + (int64_t)(m + t) * mp;
    sT[i + (int64_t)t * mp] = v ? S_all[src] : 0.0;
    gT[i + (int64_t)t * mp] = v ? G_all[src] : 0.0;
    if (i == 0) {
        const int64_t dg = (int64_t)(m + t) * mp + (m + t);
        sc[t] = t < T ? S_all[dg] : 1.0;
        sc[OAT_T + t] = t < T ? G_all[dg] : 0.0;
        sc[2 * OAT_T + t] = t < T ? b_all[m + t] : 0.0;
    }
}

__global__ void oat_axpby_kernel(const double *__restrict__ x, double b, const double *__restrict__ y, int64_t n,
                                 double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] + b * y[i];
}

// one CTA per candidate t: out[0..4][t] = |ES_t|^2, |EA_t|^2, w_t.H_t, w_t.g_t, EA_t.t1
__global__ void __launch_bounds__(256)
oat_finish_kernel(const double *__restrict__ ES, const double *__restrict__ EA, const double *__restrict__ wv,
                  const double *__restrict__ H, const double *__restrict__ gT, const double *__restrict__ t1, int mp,
                  double *__restrict__ out)
{
    __shared__ double red[8][5];
    const int t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < mp; i += 256) {
        const int64_t o = i + (int64_t)t * mp;
        const double es = ES[o], ea = EA[o], w = wv[o];
        a[0] = fma(es, es, a[0]);
        a[1] = fma(ea, ea, a[1]);
        a[2] = fma(w, H[o], a[2]);
        a[3] = fma(w, gT[o], a[3]);
        a[4] = fma(ea, t1[i], a[4]);
    }
#pragma unroll
    for (int q = 0; q < 5; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
        if (lane == 0) red[warp][q] = a[q];
    }
    __syncthreads();
    if (tid < 5) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[w][tid];
        out[tid * OAT_T + t] = v;
    }
}

// w is planned for m + T knots and w->U holds [U; C].  obj0: objective with the m knots; objs[T].
static int gauss_oat_vi(srgp_ctx *ctx, GaussWS *w, int kernel, double sigma, const double *l, double tau, double delta,
                        int m, int T, double *obj0, double *objs)
{
    cudaStream_t s = ctx->stream, s2 = ctx->stream2;
    const int mp = w->mp, m_all = w->m, d = w->d;
    const size_t mm = (size_t)mp * mp;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    w->k_reuse = false;
    const double Z = tau * tau + delta, B = 1.0 / Z, itau2 = 1.0 / (tau * tau);

    double *G_all = w->red1.d(), *b_all = G_all + mm, *tail = b_all + mp;
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W::S_INFO, 0, 16, s));
    double *S_all = w->mat(W::M_S), *Spad = w->mat(W::M_X2), *Sinv = w->mat(W::M_SINV), *A = w->mat(W::M_A);
    double *C = w->mat(W::M_C), *Linv = w->mat(W::M_LINV), *LinvT = w->mat(W::M_X1), *tmp = w->mat(W::M_TMP);
    double *T1 = w->mat(W::M_T1), *LinvS = w->mat(W::M_L1), *LinvTS = w->mat(W::M_L2);
    double *sT = w->mat(W::M_T2), *gT = w->mat(W::M_CG), *ES = w->mat(W::M_CGS), *wv = w->mat(W::M_SG);
    double *H = w->mat(W::M_SGS), *aT = w->mat(W::M_N), *EA = w->mat(W::M_MOP);
    double *bv = w->vec(W::V_B), *t1 = w->vec(W::V_T1), *gsc = w->gemv_scratch();
    // [3][128] candidate scalars, then [5][128] reductions = 1024 doubles: vectors V_T2 .. V_T2 + 7 when mp = 128
    double *sc = w->vec(W::V_T2);
    static_assert(W::V_T2 + 8 <= W::NVECS, "candidate scalars must stay inside the vector block");

    // ---- side stream: S over [U; C]; the m x m leading block (identity beyond m) is factorised ------------------
    SRGP_TRY(stream_fork(ctx));
    SRGP_TRY(assemble_dev_ld(ctx, s2, kernel, w->U.d(), m_all, d, sigma, l, delta, S_all, mp));
    SRGP_TRY(dense::pad_identity(ctx, s2, S_all, mp, m_all, 1.0));
    SRGP_CUDA(cudaMemcpyAsync(Spad, S_all, mm * 8, cudaMemcpyDeviceToDevice, s2));
    SRGP_TRY(dense::pad_identity(ctx, s2, Spad, mp, m, 1.0));
    SRGP_CUDA(cudaMemcpyAsync(T1, Spad, mm * 8, cudaMemcpyDeviceToDevice, s2));
    SRGP_TRY(dense::chol_inverse(ctx, s2, T1, mp, m, w->dinv(0), LinvS, LinvTS, w->mat(W::M_L3), Sinv, w->info(0),
                                 w->sc(W::S_LOGDET_S)));

    // ---- pass 1 over the data rows with m + T knots --------------------------------------------------------------
    SRGP_TRY(gauss_pass1(ctx, w, gp, nullptr, w->r.d(), G_all, b_all));
    SRGP_TRY(copy_scalar(ctx, tail, w->sc(W::S_S0), 1));
    SRGP_TRY(set_scalar(ctx, tail + 1, (double)ctx->n));
    SRGP_TRY(comm_allreduce(ctx, G_all, mm + mp + 2, s));
    SRGP_TRY(copy_scalar(ctx, w->sc(W::S_S0TOT), tail, 1));
    SRGP_TRY(copy_scalar(ctx, w->sc(W::S_NTOT), tail + 1, 1));
    SRGP_TRY(stream_join(ctx));

    // ---- borders, then restrict G1 / b1 to the m knots ------------------------------------------------------------
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        oat_extract_kernel<<<dim3((unsigned)ceil_div(mp, 256), OAT_T), 256, 0, s>>>(S_all, G_all, b_all, mp, m, T, sT, gT, sc);
        SRGP_LAUNCH_CHECK();
    }
    SRGP_TRY(dense::pad_identity(ctx, s, G_all, mp, m, 0.0));
    SRGP_TRY(axpby_vec(ctx, mp, B, b_all, 0.0, nullptr, bv));
    if (mp > m) SRGP_CUDA(cudaMemsetAsync(bv + m, 0, (size_t)(mp - m) * 8, s));
    // A = S + B G1, C = A^-1 (only the factors are used), t1 = L_A^-1 b
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, Spad, B, G_all, 0.0, A));
    SRGP_TRY(dense::chol_inverse(ctx, s, A, mp, m, w->dinv(1), Linv, LinvT, tmp, C, w->info(1), w->sc(W::S_LOGDET_A)));
    SRGP_TRY(dense::gemv(ctx, s, mp, 1.0, Linv, bv, 0.0, nullptr, t1, gsc));
    SRGP_TRY(dense::dot_v(ctx, s, m, t1, t1, w->sc(W::S_BV)));
    SRGP_TRY(dense::dot_mm(ctx, s, mp, m, Sinv, G_all, w->sc(W::S_SUMQ), w->nspart.d()));
    // ---- per-candidate panels ---------------------------------------------------------------------------------
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'N', mp, OAT_T, mp, 1.0, LinvS, mp, sT, mp, 0.0, ES, mp));
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'N', mp, OAT_T, mp, 1.0, LinvTS, mp, ES, mp, 0.0, wv, mp));
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'N', mp, OAT_T, mp, 1.0, G_all, mp, wv, mp, 0.0, H, mp));
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        const int64_t np = (int64_t)mp * OAT_T;
        oat_axpby_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(sT, B, gT, np, aT);
        SRGP_LAUNCH_CHECK();
    }
    SRGP_TRY(dense::gemm(ctx, s, 'N', 'N', mp, OAT_T, mp, 1.0, Linv, mp, aT, mp, 0.0, EA, mp));
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        oat_finish_kernel<<<OAT_T, 256, 0, s>>>(ES, EA, wv, H, gT, t1, mp, sc + 3 * OAT_T);
        SRGP_LAUNCH_CHECK();
    }
    std::vector<double> h8(8 * OAT_T);
    SRGP_CUDA(cudaMemcpyAsync(h8.data(), sc, h8.size() * 8, cudaMemcpyDeviceToHost, s));
    SRGP_TRY(fetch_scalars(ctx, w));   // synchronises; a failed factorisation of S or A (the CURRENT knots) is an error

    const double *h = w->h_scal;
    const double n = h[W::S_NTOT], s0 = h[W::S_S0TOT];
    auto objective = [&](double bCb, double ldS, double ldA, double sumq) {
        const double tt = -(0.5 * itau2) * (n * (sigma * sigma + delta) - sumq);
        return -0.5 * B * s0 + 0.5 * bCb - 0.5 * (n * log(Z) - ldS + ldA) - 0.5 * n * log(2.0 * M_PI) + tt;
    };
    *obj0 = objective(h[W::S_BV], h[W::S_LOGDET_S], h[W::S_LOGDET_A], h[W::S_SUMQ]);
    const double *kap = &h8[0], *gam = &h8[OAT_T], *kcr = &h8[2 * OAT_T];
    const double *es2 = &h8[3 * OAT_T], *ea2 = &h8[4 * OAT_T], *wH = &h8[5 * OAT_T], *wg = &h8[6 * OAT_T],
                 *eat1 = &h8[7 * OAT_T];
    for (int t = 0; t < T; t++) {
        // a Schur complement inside its own rounding noise ((m + 1) eps x the diagonal entry) is a numerically
        // duplicated knot: R's solve() / chol() stop there ("computationally singular"), the caller resamples
        const double sS = kap[t] - es2[t], alp = kap[t] + B * gam[t], sA = alp - ea2[t];
        const double noise = (double)(m + 1) * 2.220446049250313e-16;
        if (!(sS > noise * kap[t]) || !(sA > noise * alp)) {
            objs[t] = NAN;
            continue;
        }
        const double q = B * kcr[t] - eat1[t];
        objs[t] = objective(h[W::S_BV] + q * q / sA, h[W::S_LOGDET_S] + log(sS), h[W::S_LOGDET_A] + log(sA),
                            h[W::S_SUMQ] + (wH[t] - 2.0 * wg[t] + gam[t]) / sS);
    }
    return SRGP_OK;
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_oat_scores(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, const double *cand,
                               int64_t n_cand, double sigma, const double *l, double tau, double delta, double *obj0,
                               double *scores)
{
    if (!ctx || !xu || !cand || !l || !scores || m <= 0 || n_cand <= 0) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    if (!ctx->have_data) {
        set_error("srgp_oat_scores called before srgp_set_data");
        return SRGP_ERR_STATE;
    }
    if (kernel != SRGP_SQEXP && kernel != SRGP_ARD) {
        set_error("Error: invalid covariance function (the sparse Gaussian models take \"sqexp\" or \"ard\")");
        return SRGP_ERR_UNKNOWN_KERNEL;
    }
    if (model != SRGP_VI && model != SRGP_FIC) {
        set_error("unknown model %d", model);
        return SRGP_ERR_ARG;
    }
    if (m + 1 > 32768) {
        set_error("m = %lld knots exceeds the supported 32767 + 1 candidate", (long long)m);
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(use_device(ctx));
    GaussWS *w = gauss_ws(ctx);
    const int d = ctx->d;
    double o0 = NAN;
    if (model == SRGP_VI) {
        for (int64_t c0 = 0; c0 < n_cand; c0 += OAT_T) {
            const int T = (int)std::min<int64_t>(OAT_T, n_cand - c0), ma = (int)m + T;
            if (ma > 32768) {
                set_error("m + candidates = %d exceeds the supported 32768", ma);
                return SRGP_ERR_ARG;
            }
            std::vector<double> ua((size_t)ma * d);      // [U; C] column-major
            for (int c = 0; c < d; c++) {
                for (int64_t k = 0; k < m; k++) ua[k + (size_t)ma * c] = xu[k + m * c];
                for (int t = 0; t < T; t++) ua[m + t + (size_t)ma * c] = cand[c0 + t + n_cand * c];
            }
            SRGP_TRY(plan(ctx, w, ma, d));
            SRGP_TRY(upload_knots(w, ua.data(), ua.size() * 8, ctx->stream));
            SRGP_CUDA(cudaStreamSynchronize(ctx->stream));   // ua is pageable and dies with this iteration
            SRGP_TRY(gauss_oat_vi(ctx, w, kernel, sigma, l, tau, delta, (int)m, T, &o0, scores + c0));
        }
        if (obj0) *obj0 = o0;
        return SRGP_OK;
    }
    // FIC: objective-only evaluations with [U; c] on the resident shard
    if (obj0) {
        SRGP_TRY(plan(ctx, w, (int)m, d));
        SRGP_TRY(upload_knots(w, xu, (size_t)m * d * 8, ctx->stream));
        SRGP_TRY(gauss_fic(ctx, w, kernel, sigma, l, tau, delta, obj0, nullptr));
    }
    const int ma = (int)m + 1;
    std::vector<double> ua((size_t)ma * d);
    for (int64_t t = 0; t < n_cand; t++) {
        for (int c = 0; c < d; c++) {
            for (int64_t k = 0; k < m; k++) ua[k + (size_t)ma * c] = xu[k + m * c];
            ua[m + (size_t)ma * c] = cand[t + n_cand * c];
        }
        SRGP_TRY(plan(ctx, w, ma, d));
        SRGP_TRY(upload_knots(w, ua.data(), ua.size() * 8, ctx->stream));
        SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
        const int rc = gauss_fic(ctx, w, kernel, sigma, l, tau, delta, scores + t, nullptr);
        if (rc == SRGP_ERR_NOT_PD) scores[t] = NAN;      // R: try-error -> the caller resamples this candidate
        else if (rc != SRGP_OK) return rc;
    }
    return SRGP_OK;
}
