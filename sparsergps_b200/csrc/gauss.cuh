// gauss.cuh -- workspace of the fused Gaussian (VI / FIC) and Laplace pipelines.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace srgp {

struct GenParams {
    double sigma2;                 // sigma^2
    double invl[SRGP_MAX_D];       // 1 / l_c (sqexp: the same 1 / l in every dimension)
};

// Chunk buffers of the INT8 pass 1.  With two, generator c + 2 cannot start before Gram c has finished and the two kernels
// alternate instead of overlapping (timeline of pass 1: 207 us per chunk = 64 us generator + 144 us Gram, the serial sum).
constexpr int PASS1_BUFS = 4;
constexpr int PART2_VEC_GROUPS = 320;   // part2 holds at least this many mp-vectors (laplace.cu: the K^T v partials)

struct GaussWS {
    // plan (depends on m, d and the SM count)
    bool planned = false;
    int m = 0, mp = 0, d = 0, nt = 0, pairs = 0;
    int splits = 1, rows1 = 0, gen_groups = 1;    // pass 1: SYRK split-K over the chunk rows
    int cgroups = 1, rblocks = 1, rows2 = 0;      // pass 2: row blocks x column groups
    size_t chunk_elems = 0;                       // doubles per chunk buffer (PASS1_BUFS buffers)
    bool big_chunks = false;                      // shard of >= 2 GiB of K: 256 MB chunk buffers for the INT8 pass 1 (plan)

    DevBuf U;        // knots, m x d column-major
    DevBuf chunk;    // L2-resident K chunk (row-major in pass 1, column-major in pass 2)
    const double *pass1_kmat = nullptr;   // set around a gauss_pass1 call: the shard's K as a row-major matrix (ld = mp) the
                                          // sqrt-weight generator may read instead of recomputing exp (laplace.cu)
    DevBuf Gpart;    // pass-1 Gram slots [pairs][splits][128*128] in fragment order
    DevBuf b1part;   // pass-1 K^T r slots [gen_groups][mp]
    DevBuf red1;     // pass-1 allreduce buffer: [G1 mp*mp | b1 mp | 16 scalars]
    DevBuf r;        // residual y - mu (n)
    DevBuf rowa;     // per-row vectors (NROWV x n)
    DevBuf mats;     // NMATS m x m matrices (ld = mp) + two diagonal-block-inverse areas
    DevBuf vecs;     // NVECS m-vectors + gemv scratch
    DevBuf scal;     // device scalars
    DevBuf part2;    // pass-2 per-CTA partial sums
    DevBuf coin;     // bit-identical (row, knot) pairs found in pass 2 (quirk Q4): [64 B header | cap x (i, j) | cap x omega]
    DevBuf coinrow;  // per-row sums of the pair values (all zero between evaluations): the deterministic summation order
    int coin_cap = 0;      // pairs the list can hold = rows of the largest shard seen + 65536 (coin_reset)
    int64_t coinrow_n = 0;
    DevBuf rowpart;  // per-column-group row sums of one chunk (row-form passes)
    DevBuf nspart;   // scratch of ns_reduce (runs on the side stream)
    DevBuf rowdpart; // per-(column group, warp column) partials of one chunk in the per-row / per-dimension mode
    DevBuf rowd;     // FIC: per-row, per-dimension sums of passes 1a and 2a
    DevBuf knotpart; // pass 2 with knot gradients: [knot_slots][d][mp] column sums of P o (x - u) / l
    int knot_slots = 0;   // rblocks x 4 (INT8 epilogue: one slot per row block and lane quadrant), rblocks (DMMA epilogue)
    DevBuf knotsum;  // [d][mp] sums over the shard (allreduced), then the m x d knot gradient (knot-major)
    // knot-gradient request of the current call (set by the entry point, read by gauss_pass2 / knot_finish)
    bool want_knots = false, knot_transform = false;
    double knot_lb[SRGP_MAX_D], knot_ub[SRGP_MAX_D];
    DevBuf i8scal;   // INT8 weighted Gram: max |w| over the shard (device scalar)
    DevBuf i8buf;    // INT8 tensor-core passes: digit slices of Mop (8 x mp x mp bytes) + per-column scales
    // The pass-2 operand image of the WHOLE shard (digit slices of exp(-d^2/2), blocks = 128-row blocks, k = knots: NS
    // bytes per entry, 7.2 GB at n = 1e6, m = 1024), written by the first K*M / row-form pass of an evaluation that
    // announces several of them (k_reuse) and read by the later ones.  Absent when it does not fit (the passes then
    // regenerate their chunks).  Keyed on everything K depends on, see K2Key.
    DevBuf k2;
    bool k2_valid = false, k_reuse = false;
    // VI: the image is generated AHEAD of pass 2, on the generator stream, while the m x m stage keeps the tensor pipe idle
    // (gauss_pregen_k2); k2_ev[c] fires when chunk c is complete, k2_pending until pass 2 has waited for them
    std::vector<cudaEvent_t> k2_ev;
    bool k2_pending = false;
    struct K2Key {
        const double *Xp;
        int64_t n;
        int mp, m, d;
        uint64_t uver, xver;
        double invl[SRGP_MAX_D];
    } k2_key = {};
    uint64_t u_version = 0;    // bumped by upload_knots
    DevBuf Kmat;     // Laplace: the shard's K, row-major [rows][mp], kept for the whole Newton loop (theta fixed)
    double *h_scal = nullptr;   // pinned mirror of scal

    enum { NMATS = 24, NVECS = 16, NROWV = 20, NSCAL = 256, COIN_SLACK = 65536 };
    enum Mat { M_S = 0, M_SINV, M_A, M_C, M_LINV, M_TMP, M_CG, M_CGS, M_SG, M_SGS, M_N, M_MOP, M_T1, M_T2, M_X1, M_X2,
               M_GZ, M_CZ, M_GWP, M_L1, M_L2, M_L3, M_L4, M_L5 };
    enum Vec { V_B = 0, V_V, V_GV, V_TMP, V_BETA, V_T1, V_T2, V_T3, V_T4, V_T5, V_T6, V_T7 };
    enum Scal {
        S_S0 = 0, S_LOGDET_S, S_LOGDET_A, S_SUMQ, S_TRCG1, S_BV, S_B1V, S_VGV, S_INFO, S_INFO_HI, S_NTOT, S_S0TOT, S_Q4,
        S_P2 = 32,      // pass-2 sums: 1 + d entries, + q4 at S_P2 + 1 + d, sum rho at + 2 + d, pair-list overflow flag at
                        // + 3 + d  (allreduced together: P2_LEN(d) entries)
        S_NS = 128,     // N o dS sums: 2 + d entries
        S_X = 224       // model-specific extras
    };

    double *mat(int i) const { return mats.d() + (size_t)i * mp * mp; }
    double *dinv(int which) const { return mats.d() + (size_t)NMATS * mp * mp + (size_t)which * ((size_t)2 * mp * 128 + 256); }
    double *vec(int i) const { return vecs.d() + (size_t)i * mp; }
    double *gemv_scratch() const { return vecs.d() + (size_t)NVECS * mp; }
    // per-row vectors: stride padded to 128 doubles (16-byte aligned bulk copies of weight tiles) and followed
    // by a zeroed tail, because the weighted SYRK reads weights for the zero rows that pad a chunk
    static size_t row_stride(int64_t n) { return ((size_t)n + 127) / 128 * 128 + 1024; }
    double *rowv(int i, int64_t n) const { return rowa.d() + (size_t)i * row_stride(n); }
    double *sc(int i) const { return scal.d() + i; }
    int *info(int which) const { return reinterpret_cast<int *>(scal.d() + S_INFO) + which; }
    static int p2_len(int d) { return d + 4; }
    int *coin_count() const { return reinterpret_cast<int *>(coin.p); }
    int *coin_list() const { return reinterpret_cast<int *>(coin.p) + 16; }
    double *coin_omega() const
    {
        return reinterpret_cast<double *>(reinterpret_cast<char *>(coin.p) + 64 + (size_t)coin_cap * 2 * sizeof(int));
    }
    void release();
};

GaussWS *gauss_ws(srgp_ctx *ctx);
int plan(srgp_ctx *ctx, GaussWS *w, int m, int d);
void fill_gen(GenParams &p, int kernel, int d, double sigma, const double *l);
// host knots -> w->U (every change of the knots goes through here: it invalidates the resident K image)
int upload_knots(GaussWS *w, const double *host, size_t bytes, cudaStream_t s);

// pass 1 over the resident shard: G (mp x mp, both triangles) = K^T diag(rowweight) K, b1 = K^T (rvec)
// (rowweight may be null = 1; when given, rvec must already contain the weight).
// weight_nonneg: rowweight >= 0 everywhere (FIC's B = 1 / Z): lets the INT8 pass slice sqrt(w) K once.
int gauss_pass1(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *rowweight, const double *rvec,
                double *G, double *b1, bool weight_nonneg = false);
// pass 2: out[0] = sum_ij P_ij, out[1 + c] = sum_ij P_ij ((x_ic - u_jc) / l_c)^2 with
// P = (rs_i (K Mop^T)_ij + ra_i beta_j) K_ij; bit-identical (row, knot) pairs are appended to w->coin.
// accumulate_slots: add to the per-CTA slots of a previous gauss_pass2 call instead of restarting them.
int gauss_pass2(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs,
                const double *ra, const double *beta, double *out, bool accumulate_slots);
// Row quadratic forms over the shard: rowq_i = K_i Mop K_i^T (Mop symmetric), rowkv_i = K_i v (v may be null).
int gauss_rowform(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *vvec,
                  double *rowq, double *rowkv);
int rowform_chunk(srgp_ctx *ctx, GaussWS *w, const double *Mop, int rows_valid, double *rowq);
// Per-row, per-dimension sums over the shard (see gauss.cu): out[slot * stride + i]
int gauss_rowd(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *beta,
               const double *vvec, double *out, int64_t stride);
// Laplace: materialise the shard's K row-major into w->Kmat (rows padded with zeros to the SYRK quantum)
int materialise_k(srgp_ctx *ctx, GaussWS *w, const GenParams &gp);
// G = Kmat^T diag(rowweight) Kmat over the materialised shard (rowweight may be null)
int gram_materialised(srgp_ctx *ctx, GaussWS *w, const double *rowweight, double *G);
// out[0] = sum N o Kuu, out[1 + c] = sum N o Kuu o D_c, out[1 + d] = sum of N over bit-identical knot pairs
int ns_reduce(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *N, const double *S, double nugget,
              double *out, cudaStream_t s);
// quirk Q4: *out = sum over recorded pairs of (omega_p - coef * (K S^-1)_{i_p j_p})
// Empties the pair list before a pass that records pairs (grows it to the shard's row count + slack; clears the
// overflow flag p2[d + 3]).  A list that still overflows (many rows coinciding with SEVERAL duplicated knots) is an error
// reported by fetch_scalars on every rank, never a silently truncated tau gradient.
int coin_reset(srgp_ctx *ctx, GaussWS *w);
int coin_check(const GaussWS *w);
int coin_fix(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Sinv, double coef, double *out);
// FIC variant: pairs recorded by gauss_rowd(C, ...) carry (K C)_ij; *out = sum of Omega_ij over the pairs
int coin_fix_fic(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Sinv, const double *B, const double *rho,
                 const double *alpha, const double *beta, double *out);
// knot-location gradient from the pass-2 column sums (w->knotpart) and N: see knot_finish_kernel
int knot_finish(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *N, const double *S);
int scale_vec(srgp_ctx *ctx, const double *x, int64_t n, double a, double *out);
int axpby_vec(srgp_ctx *ctx, int n, double a, const double *x, double b, const double *z, double *y);
int set_scalar(srgp_ctx *ctx, double *dst, double v);
int copy_scalar(srgp_ctx *ctx, double *dst, const double *src, int count);

int comm_allreduce(srgp_ctx *ctx, double *buf, size_t count, cudaStream_t s);
// D2H of the scalar block + stream sync; maps a failed m x m Cholesky to SRGP_ERR_NOT_PD
int fetch_scalars(srgp_ctx *ctx, GaussWS *w);
// side stream helpers: fork makes ctx->stream2 wait for everything issued so far on ctx->stream; join makes
// ctx->stream wait for everything issued so far on ctx->stream2
int stream_fork(srgp_ctx *ctx);
int stream_join(srgp_ctx *ctx);
int gauss_vi(srgp_ctx *ctx, GaussWS *w, int kernel, double sigma, const double *l, double tau, double delta,
             double *obj, double *grad);
int gauss_fic(srgp_ctx *ctx, GaussWS *w, int kernel, double sigma, const double *l, double tau, double delta,
              double *obj, double *grad);
// One evaluation on the resident shard behind srgp_gauss_obj_grad / srgp_gauss_obj_grad_knots (argument checks,
// plan, knot upload, model dispatch; knot_grad is a HOST buffer of m * d doubles when knots is set)
int gauss_eval(srgp_ctx *ctx, int model, int kernel, const double *xu, int64_t m, double sigma, const double *l,
               double tau, double delta, double *obj, double *grad, bool knots, const double *knot_lb,
               const double *knot_ub, double *knot_grad);
// K1 with an explicit output leading dimension (assemble.cu)
int assemble_dev_ld(srgp_ctx *ctx, cudaStream_t s, int kernel, const double *x_dev, int64_t n1, int d, double sigma,
                    const double *l, double nugget, double *out_dev, int64_t ldo);

}  // namespace srgp
