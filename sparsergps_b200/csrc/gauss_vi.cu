// gauss_vi.cu -- Gaussian VI (Titsias) objective + gradient: the stage sequence behind srgp_gauss_obj_grad.
//
// Reference: elbo_fun (R/vi_functions.R:64-121), trace_term_fun (:14-27), delbo_dcov_par (:126-420).
// Reduced form (checked against the literal transcription by tests/test_oracle.py):
//   Z = tau^2 + delta, B = 1/Z, r = y - mu, S = K_uu + delta I, G1 = K^T K, b1 = K^T r, G = B G1
//   C = (S + G)^-1, b = B b1, v = C b, beta = S^-1 (b - G v), sum_q = tr(S^-1 G1)
//   tt  = -(n (sigma^2 + delta) - sum_q) / (2 tau^2)
//   obj = -B s0/2 + b^T v/2 - (n log Z - log|S| + log|S + G|)/2 - n log(2 pi)/2 + tt
//   M   = (1/tau^2 - B) S^-1 + B C G S^-1 = S^-1/tau^2 - B C ;  Omega = K (M - B v beta^T) + B r beta^T
//   N   = S^-1 G S^-1/2 - S^-1 G C G S^-1/2 - beta beta^T/2 - S^-1 G1 S^-1/(2 tau^2)
//       = (S^-1 - C)/2 - beta beta^T/2 - S^-1 G1 S^-1/(2 tau^2)          (G = A - S, A C = I)
//   g_sigma = 2 sum Omega o K + 2 sum N o K_uu - sigma^2 n / tau^2
//   g_l_c   = sum Omega o K o D_c + sum N o K_uu o D_c(u)
//   g_tau   = tau^2 sum alpha^2 - tau^2 (n B - B^2 tr(C G1)) - 2 tt  [+ quirk Q4 pairs],
//             sum alpha^2 = B^2 (s0 - 2 b1^T v + v^T G1 v)
#include <math.h>

#include "dense.cuh"
#include "gauss.cuh"
#include "gauss_i8.cuh"

namespace srgp {

using W = GaussWS;

int stream_fork(srgp_ctx *ctx)
{
    SRGP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    SRGP_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    return SRGP_OK;
}

int stream_join(srgp_ctx *ctx)
{
    SRGP_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream2));
    SRGP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return SRGP_OK;
}

int fetch_scalars(srgp_ctx *ctx, GaussWS *w)
{
    SRGP_CUDA(cudaMemcpyAsync(w->h_scal, w->scal.p, W::NSCAL * 8, cudaMemcpyDeviceToHost, ctx->stream));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("device execution failed: %s", cudaGetErrorString(e));
        return SRGP_ERR_CUDA;
    }
    const int *info = reinterpret_cast<const int *>(w->h_scal + W::S_INFO);
    for (int k = 0; k < 4; k++)
        if (info[k] != 0) {
            set_error("the leading minor of order %d is not positive definite (Cholesky %d of the m x m stage)",
                      info[k], k);
            return SRGP_ERR_NOT_PD;
        }
    return SRGP_OK;
}

// After fetch_scalars of an evaluation that recorded quirk-Q4 pairs: the overflow flag travelled in the pass-2 allreduce,
// so it is the sum over the ranks and every rank fails together.
int coin_check(const GaussWS *w)
{
    if (w->h_scal[W::S_P2 + w->d + 3] > 0.0) {
        set_error("more than %d bit-identical (data row, knot) pairs on one shard (duplicated knots on gridded inputs?): "
                  "the tau gradient of quirk Q4 would be truncated", w->coin_cap);
        return SRGP_ERR_STATE;
    }
    return SRGP_OK;
}

int gauss_vi(srgp_ctx *ctx, GaussWS *w, int kernel, double sigma, const double *l, double tau, double delta,
             double *obj, double *grad)
{
    cudaStream_t s = ctx->stream;
    const int mp = w->mp, m = w->m, d = w->d;
    const size_t mm = (size_t)mp * mp;
    GenParams gp;
    fill_gen(gp, kernel, d, sigma, l);
    w->k_reuse = false;   // one K*M pass: its chunks stay in L2 (gauss_i8.cu)
    const double Z = tau * tau + delta, B = 1.0 / Z, itau2 = 1.0 / (tau * tau);

    double *G1 = w->red1.d(), *b1 = G1 + mm, *tail = b1 + mp;
    SRGP_CUDA(cudaMemsetAsync(w->scal.d() + W::S_INFO, 0, 16, s));
    SRGP_TRY(coin_reset(ctx, w));

    double *S = w->mat(W::M_S), *Sinv = w->mat(W::M_SINV), *A = w->mat(W::M_A), *C = w->mat(W::M_C);
    double *Linv = w->mat(W::M_LINV), *tmp = w->mat(W::M_TMP);
    double *SG = w->mat(W::M_SG), *SGS = w->mat(W::M_SGS), *N = w->mat(W::M_N), *Mop = w->mat(W::M_MOP);
    double *T1 = w->mat(W::M_T1), *LinvT = w->mat(W::M_X1);
    double *bv = w->vec(W::V_B), *v = w->vec(W::V_V), *gv = w->vec(W::V_GV), *tv = w->vec(W::V_TMP);
    double *beta = w->vec(W::V_BETA), *gsc = w->gemv_scratch();
    cudaStream_t s2 = ctx->stream2;

    // ---- side stream: S = K_uu + delta I (self-covariance minus tau^2 I: R/vi_functions.R:736-741), identity on
    //      the padding; Cholesky, S^-1, log|S|.  Independent of the data rows, so it overlaps pass 1.  The fork is taken
    //      here, but the ~45 launches of that chain are enqueued AFTER pass 1's: the host needs ~0.2 ms for them, and the
    //      first generator of pass 1 would otherwise start that much later (timeline of a 125 000-row shard). ----
    SRGP_TRY(stream_fork(ctx));

    // ---- pass 1 (main stream) --------------------------------------------------------------------------------
    SRGP_TRY(gauss_pass1(ctx, w, gp, nullptr, w->r.d(), G1, b1));
    SRGP_TRY(assemble_dev_ld(ctx, s2, kernel, w->U.d(), m, d, sigma, l, delta, S, mp));
    SRGP_TRY(dense::pad_identity(ctx, s2, S, mp, m, 1.0));
    SRGP_CUDA(cudaEventRecord(ctx->ev_aux, s2));                    // S itself is ready long before its inverse
    SRGP_CUDA(cudaMemcpyAsync(T1, S, mm * 8, cudaMemcpyDeviceToDevice, s2));
    SRGP_TRY(dense::chol_inverse(ctx, s2, T1, mp, m, w->dinv(0), w->mat(W::M_L1), w->mat(W::M_L2), w->mat(W::M_L3), Sinv,
                                 w->info(0), w->sc(W::S_LOGDET_S)));
    // pass 2's K does not depend on the m x m stage: its generators start now, on their own stream, and work while that
    // stage (latency-bound, tensor pipe idle) runs; pass 2 then finds most of its chunks waiting
    if (grad) SRGP_TRY(gauss_pregen_k2(ctx, w, gp));
    SRGP_TRY(copy_scalar(ctx, tail, w->sc(W::S_S0), 1));
    SRGP_TRY(set_scalar(ctx, tail + 1, (double)ctx->n));
    SRGP_TRY(comm_allreduce(ctx, G1, mm + mp + 2, s));
    SRGP_TRY(copy_scalar(ctx, w->sc(W::S_S0TOT), tail, 1));
    SRGP_TRY(copy_scalar(ctx, w->sc(W::S_NTOT), tail + 1, 1));
    SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_aux, 0));

    // ---- replicated m x m stage: the chain pass 2 waits for ------------------------------------------------------
    // A = S + B G1 ; C = A^-1
    SRGP_TRY(dense::axpby(ctx, s, mp, m, 1.0, S, B, G1, 0.0, A));
    SRGP_TRY(dense::chol_inverse(ctx, s, A, mp, m, w->dinv(1), Linv, LinvT, tmp, C, w->info(1),
                                 w->sc(W::S_LOGDET_A)));
    // b = B b1 ; v = C b ; gv = G1 v ; beta = S^-1 (b - B gv)
    // (vector solves go through the triangular factors, L^-T (L^-1 x): their forward error scales with
    //  sqrt(cond) instead of cond for the explicit inverse -- matters for the OAT-shaped configs, cond(S) ~ 1e4+)
    double *t1 = w->vec(W::V_T1), *t2 = w->vec(W::V_T2);
    SRGP_TRY(axpby_vec(ctx, mp, B, b1, 0.0, nullptr, bv));
    // (gemv_t: one launch per product, from the transposed copy trtri leaves behind; G1 is symmetric)
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, LinvT, bv, 0.0, nullptr, t1));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, Linv, t1, 0.0, nullptr, v));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, G1, v, 0.0, nullptr, gv));
    SRGP_TRY(axpby_vec(ctx, mp, 1.0, bv, -B, gv, tv));
    // the factorisation of A above needed S only; the factors and the inverse of S (side stream) are needed from here on.
    // On a short shard (8-GPU runs) that chain, slowed by the co-running pass 1, ends after pass 1 does.
    SRGP_TRY(stream_join(ctx));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, w->mat(W::M_L2), tv, 0.0, nullptr, t2));
    SRGP_TRY(dense::gemv_t(ctx, s, mp, 1.0, w->mat(W::M_L1), t2, 0.0, nullptr, beta));
    if (grad) {
        // Mop = (1/tau^2 - B) S^-1 + B^2 C G1 S^-1 - B beta v^T.  With C (S + B G1) = I,  B C G1 S^-1 = S^-1 - C, so
        // Mop = S^-1 / tau^2 - B C - B beta v^T: no product on the chain pass 2 waits for
        SRGP_TRY(dense::axpby(ctx, s, mp, m, itau2, Sinv, -B, C, 0.0, Mop));
        SRGP_TRY(dense::ger(ctx, s, mp, -B, beta, v, Mop));
    }
    // ---- side stream: everything of the m x m stage that pass 2 does not need (scalars, N, sum N o dS) ----------
    SRGP_TRY(stream_fork(ctx));
    SRGP_TRY(dense::dot_mm(ctx, s2, mp, m, Sinv, G1, w->sc(W::S_SUMQ), w->nspart.d()));
    SRGP_TRY(dense::dot_mm(ctx, s2, mp, m, C, G1, w->sc(W::S_TRCG1), w->nspart.d()));
    SRGP_TRY(dense::dot_v(ctx, s2, m, t1, t1, w->sc(W::S_BV)));   // b^T (S+G)^-1 b = |L^-1 b|^2
    SRGP_TRY(dense::dot_v(ctx, s2, m, b1, v, w->sc(W::S_B1V)));
    SRGP_TRY(dense::dot_v(ctx, s2, m, v, gv, w->sc(W::S_VGV)));
    if (grad) {
        // N = (B/2 - 1/(2 tau^2)) X - (B^2/2) S^-1 G1 C G1 S^-1 - beta beta^T/2 with X = S^-1 G1 S^-1.  Since
        // B G1 = A - S and A C = I:  B^2 S^-1 G1 C G1 S^-1 = B X - (S^-1 - C), hence
        // N = -X / (2 tau^2) + (S^-1 - C)/2 - beta beta^T/2   (two products instead of five, no cancelling pair)
        SRGP_TRY(dense::gemm(ctx, s2, 'N', 'T', mp, mp, mp, 1.0, Sinv, mp, G1, mp, 0.0, SG, mp));
        SRGP_TRY(dense::gemm(ctx, s2, 'N', 'T', mp, mp, mp, 1.0, SG, mp, Sinv, mp, 0.0, SGS, mp));
        SRGP_TRY(dense::axpby(ctx, s2, mp, m, 0.5, Sinv, -0.5, C, 0.0, N));
        SRGP_TRY(dense::axpby(ctx, s2, mp, m, 1.0, N, -0.5 * itau2, SGS, 0.0, N));
        SRGP_TRY(dense::ger(ctx, s2, mp, -0.5, beta, beta, N));
        SRGP_TRY(ns_reduce(ctx, w, gp, N, S, delta, w->sc(W::S_NS), s2));

        // ---- pass 2 (main stream) ---------------------------------------------------------------------------
        double *ra = w->rowv(0, ctx->n);
        SRGP_TRY(scale_vec(ctx, w->r.d(), ctx->n, B, ra));
        double *p2 = w->sc(W::S_P2);
        SRGP_TRY(gauss_pass2(ctx, w, gp, Mop, nullptr, ra, beta, p2, false));
        // quirk Q4: d K_ij / d log tau = 2 tau^2 on bit-identical (row, knot) pairs; the tau-gradient picks up
        // sum (Omega_ij - (K S^-1)_ij / tau^2) there (the trace-term part of Omega does not apply to tau)
        SRGP_TRY(coin_fix(ctx, w, gp, Sinv, itau2, p2 + 1 + d));
        SRGP_TRY(comm_allreduce(ctx, p2, W::p2_len(d), s));
    }
    SRGP_TRY(stream_join(ctx));
    if (grad && w->want_knots) SRGP_TRY(knot_finish(ctx, w, gp, N, S));
    SRGP_TRY(fetch_scalars(ctx, w));
    if (grad) SRGP_TRY(coin_check(w));

    // ---- host: a handful of scalars ---------------------------------------------------------------------
    const double *h = w->h_scal;
    const double n = h[W::S_NTOT], s0 = h[W::S_S0TOT];
    const double tt = -(0.5 * itau2) * (n * (sigma * sigma + delta) - h[W::S_SUMQ]);
    *obj = -0.5 * B * s0 + 0.5 * h[W::S_BV] - 0.5 * (n * log(Z) - h[W::S_LOGDET_S] + h[W::S_LOGDET_A]) -
           0.5 * n * log(2.0 * M_PI) + tt;
    if (grad) {
        const double *p2 = h + W::S_P2, *ns = h + W::S_NS;
        grad[0] = 2.0 * p2[0] + 2.0 * ns[0] - itau2 * sigma * sigma * n;
        if (kernel == SRGP_ARD) {
            for (int c = 0; c < d; c++) grad[1 + c] = p2[1 + c] + ns[1 + c];
        } else {
            double g = 0.0;
            for (int c = 0; c < d; c++) g += p2[1 + c] + ns[1 + c];
            grad[1] = g;
        }
        const int ti = (kernel == SRGP_ARD) ? 1 + d : 2;
        const double sum_alpha2 = B * B * (s0 - 2.0 * h[W::S_B1V] + h[W::S_VGV]);
        grad[ti] = tau * tau * sum_alpha2 - tau * tau * (n * B - B * B * h[W::S_TRCG1]) - 2.0 * tt +
                   2.0 * tau * tau * p2[1 + d];
    }
    return SRGP_OK;
}

}  // namespace srgp
