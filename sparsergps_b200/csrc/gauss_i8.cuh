// gauss_i8.cuh -- entry points of the INT8 tensor-core (tcgen05) row passes, see gauss_i8.cu.
#pragma once
#include "gauss.cuh"

namespace srgp {

// true unless SRGP_TENSOR=dmma (diagnostic switch that keeps the FP64 DMMA kernels for the unweighted Gram)
bool i8_enabled();
// unweighted pass 1: G = K^T K (mp x mp, both triangles), b1 = K^T rvec over the resident shard
int gauss_pass1_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *rvec, double *G, double *b1);
// out[j] = sum_g part[g][j] (gauss.cu)
void gram_sum_rows(cudaStream_t s, const double *part, int groups, int mp, double *out);

}  // namespace srgp
