// gauss_i8.cuh -- entry points of the INT8 tensor-core (tcgen05) row passes, see gauss_i8.cu.
#pragma once
#include "gauss.cuh"

namespace srgp {

// pass 1: G = K^T diag(rowweight) K (mp x mp, both triangles; rowweight null = 1), b1 = K^T rvec over the resident shard
// weight_nonneg: the caller guarantees rowweight >= 0 (one slice set sqrt(w) e serves both operands)
int gauss_pass1_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *rowweight, const double *rvec, double *G,
                   double *b1, bool weight_nonneg = false);
// pass 2 (MODE_GRAD of gauss.cu: sum P, sum P o D_c, coincident pairs) on the INT8 tensor cores; same per-CTA slots
constexpr int PART_STRIDE_I8 = SRGP_MAX_D + 8;
bool i8_pass2_supported(const GaussWS *w);
// Start generating the pass-2 operand image of the whole shard on the generator stream (after everything issued so far on
// ctx->stream); the next K*M pass with the same parameters waits chunk by chunk.  A no-op when the image does not fit.
int gauss_pregen_k2(srgp_ctx *ctx, GaussWS *w, const GenParams &gp);
int gauss_pass2_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs, const double *ra,
                   const double *beta, double *out, bool accumulate_slots);
int gauss_rowd_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *beta, const double *vvec,
                  double *out, int64_t stride);
int gauss_rowform_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *vvec, double *rowq,
                     double *rowkv);
void gram_combine_rowd(cudaStream_t s, const double *part, int groups, int nslots, int64_t ld, int rows, double *out,
                       int64_t out_stride);
void gram_sum_part(cudaStream_t s, const double *part, int slots, int stride, int count, double *out);
// out[j] = sum_g part[g][j] (gauss.cu)
void gram_sum_rows(cudaStream_t s, const double *part, int groups, int mp, double *out);

}  // namespace srgp
