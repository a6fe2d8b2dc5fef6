// gauss_i8.cu -- the row passes of the fused Gaussian pipeline on the INT8 tensor cores (tcgen05 + TMEM).
//
// Pass 1 (this file, first half): G1 = K^T K and b1 = K^T r.  The generator emits k_ij / sigma^2 = exp(-d_ij^2 / 2)
// in (0, 1] as 8 INT8 digit slices, already in the shared-memory operand image of tc_i8.cuh; the Gram kernel runs
// one CTA per (128 x 64 tile of the lower block triangle, row split), keeps all 8 significance levels of the tile
// in TMEM over its row range, converts INT32 -> FP64 once per launch and adds into its own slot (deterministic).
// Everything is exact except the final FP64 summation of levels, splits and chunks (see tc_i8.cuh for the bound),
// so the result is at least as accurate as the DMMA SYRK it replaces (profiles/r01_ozaki_*.json).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gauss.cuh"
#include "gauss_i8.cuh"
#include "tc_i8.cuh"

namespace srgp {

using namespace i8;

// ------------------------------------------------------------------------------------------------
// generator for pass 1: operand rows = knots (block = 128 knots), k index = data rows of the chunk
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128)
gen_slices_knotrows_kernel(const double *__restrict__ X, int64_t ldx, const double *__restrict__ r, int64_t r0,
                           int rows_valid, int rows_padded, const double *__restrict__ U, int m, int mp, int d_rt,
                           GenParams p, int8_t *__restrict__ slices, size_t slice_stride, double *__restrict__ b1part,
                           int first)
{
    extern __shared__ double sx[];   // [64][d] scaled rows, then [64] residuals
    const int d = DT > 0 ? DT : d_rt;
    double *sr = sx + 64 * d;
    const int j = blockIdx.x * 128 + threadIdx.x;
    const bool jvalid = j < m;
    double uj[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
        for (int c = 0; c < DT; c++) uj[c] = jvalid ? U[j + (int64_t)m * c] * p.invl[c] : 0.0;
    }
    const int KB = rows_padded / BK;
    const int kb_per_group = (KB + gridDim.y - 1) / gridDim.y;
    const int kb_begin = blockIdx.y * kb_per_group, kb_end = min(KB, kb_begin + kb_per_group);
    double bacc = 0.0;
    for (int kb = kb_begin; kb < kb_end; kb++) {
        const int it0 = kb * BK;
        __syncthreads();
        for (int t = threadIdx.x; t < BK * d; t += 128) {
            const int ii = t / d, c = t - ii * d;
            const int i = it0 + ii;
            sx[t] = (i < rows_valid) ? X[r0 + i + ldx * c] * p.invl[c] : 0.0;
        }
        if (threadIdx.x < BK) {
            const int i = it0 + threadIdx.x;
            sr[threadIdx.x] = (i < rows_valid) ? r[r0 + i] : 0.0;
        }
        __syncthreads();
        int8_t *dst = slices + ((size_t)blockIdx.x * KB + kb) * IMG_BLOCK + (size_t)threadIdx.x * 16;
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16++) {
            uint32_t w[NS][4];
#pragma unroll
            for (int s = 0; s < NS; s++) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
            for (int e0 = 0; e0 < 16; e0 += 4) {
                // 4 rows in flight per thread: independent distance / exp chains hide the FP64 latency
                double sq[4] = {0.0, 0.0, 0.0, 0.0};
                const int ii = c16 * 16 + e0;
                if (DT > 0) {
#pragma unroll
                    for (int c = 0; c < DT; c++) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const double t = sx[(ii + q) * DT + c] - uj[c];
                            sq[q] = fma(t, t, sq[q]);
                        }
                    }
                } else {
                    for (int c = 0; c < d; c++) {
                        const double ujc = jvalid ? U[j + (int64_t)m * c] * p.invl[c] : 0.0;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const double t = sx[(ii + q) * d + c] - ujc;
                            sq[q] = fma(t, t, sq[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    double ev = 0.0;
                    if (jvalid && it0 + ii + q < rows_valid) {
                        ev = exp(-0.5 * sq[q]);
                        bacc = fma(p.sigma2 * ev, sr[ii + q], bacc);
                    }
                    split_digits(ev, e0 + q, w);
                }
            }
#pragma unroll
            for (int s = 0; s < NS; s++)
                *reinterpret_cast<uint4 *>(dst + s * slice_stride + c16 * 2048) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        }
    }
    double *slot = b1part + (int64_t)blockIdx.y * mp + j;
    *slot = first ? bacc : (*slot + bacc);
}

// ------------------------------------------------------------------------------------------------
// Gram kernel: slot[tile][split] (+)= scale * sum_L 2^(-12-8L) * (INT32 level L of this launch)
// tile t -> (I, J): rows 128 I .., columns 64 J .., J <= 2 I + 1 (lower block triangle incl. the diagonal blocks)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_to_ij(int t, int &I, int &J)
{
    // tiles before row I: I (I + 1)
    I = (int)((sqrtf(4.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((I + 1) * (I + 2) <= t) I++;
    while (I * (I + 1) > t) I--;
    J = t - I * (I + 1);
}

__global__ void __launch_bounds__(THREADS, 1)
i8_gram_kernel(const int8_t *__restrict__ slices, size_t slice_stride, int KB, int nsplit, double scale,
               double *__restrict__ Gpart, int first)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars &bars = *reinterpret_cast<Bars *>(smem + STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int I, J;
    tile_to_ij(blockIdx.x / nsplit, I, J);
    const int split = blockIdx.x % nsplit;
    const int kb_per = KB / nsplit, kb0 = split * kb_per;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        mbar_init(&bars.tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_all(&bars.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < kb_per; ++it) {
                const int st = it % STAGES;
                if (it >= STAGES) mbar_wait(&bars.empty[st], ((it / STAGES) - 1) & 1);
                mbar_expect_tx(&bars.full[st], STAGE_BYTES);
                const uint32_t sbase = smem_u32(smem + st * STAGE_BYTES);
                const size_t a_off = ((size_t)I * KB + kb0 + it) * IMG_BLOCK;
                const size_t b_off = ((size_t)(J >> 1) * KB + kb0 + it) * IMG_BLOCK + (size_t)(J & 1) * 1024;
                for (int s = 0; s < NS; ++s) {
                    bulk_g2s(sbase + s * A_TILE, slices + s * slice_stride + a_off, A_TILE, &bars.full[st]);
                    for (int c = 0; c < 4; ++c)
                        bulk_g2s(sbase + NS * A_TILE + s * B_TILE + c * 1024, slices + s * slice_stride + b_off + c * 2048,
                                 1024, &bars.full[st]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < kb_per; ++it) {
                const int st = it % STAGES;
                mbar_wait(&bars.full[st], (it / STAGES) & 1);
                tc_fence_after();
                issue_stage(smem_u32(smem + st * STAGE_BYTES), tmem_base, it == 0);
                mma_commit(&bars.empty[st]);
            }
            mma_commit(&bars.tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(&bars.tmem_full, 0);
        tc_fence_after();
        const int row = q * 32 + lane;
        double *out = Gpart + ((size_t)blockIdx.x * BM + row) * BN;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            double acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = 0.0;
#pragma unroll 1
            for (int L = NS - 1; L >= 0; --L) {       // least significant level first
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(L * BN + half * 32), v);
                const double wgt = scale * exp2(-12.0 - 8.0 * L);
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] = fma(wgt, (double)(int)v[c], acc[c]);
            }
            double2 *o2 = reinterpret_cast<double2 *>(out + half * 32);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double2 prev = first ? make_double2(0.0, 0.0) : o2[c];
                o2[c] = make_double2(prev.x + acc[2 * c], prev.y + acc[2 * c + 1]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

// Sum the split slots and scatter to the full symmetric matrix (column-major, ld = mp, both triangles).
__global__ void __launch_bounds__(128)
i8_gram_finalize_kernel(const double *__restrict__ Gpart, int nsplit, int mp, double *__restrict__ G)
{
    int I, J;
    tile_to_ij(blockIdx.x, I, J);
    const int r = I * BM + threadIdx.x;
    const double *base = Gpart + ((size_t)blockIdx.x * nsplit * BM + threadIdx.x) * BN;
    for (int c = 0; c < BN; ++c) {
        double v = 0.0;
        for (int s = 0; s < nsplit; ++s) v += base[(size_t)s * BM * BN + c];
        const int col = J * BN + c;
        G[r + (int64_t)col * mp] = v;
        if (J < 2 * I) G[col + (int64_t)r * mp] = v;      // strictly below the diagonal block: mirror
    }
}

bool i8_enabled()
{
    static const bool on = [] {
        const char *e = getenv("SRGP_TENSOR");
        return !(e && (e[0] == 'd' || e[0] == 'D'));    // SRGP_TENSOR=dmma: diagnostic, keeps the DMMA kernels
    }();
    return on;
}

template <int DT>
static void launch_gen_knotrows(cudaStream_t s, dim3 grid, size_t smem, const double *X, int64_t ldx, const double *r,
                                int64_t r0, int rows_valid, int rows_padded, const double *U, int m, int mp, int d,
                                const GenParams &p, int8_t *slices, size_t slice_stride, double *b1part, int first)
{
    gen_slices_knotrows_kernel<DT><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p, slices,
                                                         slice_stride, b1part, first);
}

#define SRGP_D_SWITCH_I8(d, CALL)                    \
    switch (d) {                                     \
    case 1: CALL(1); break;                          \
    case 2: CALL(2); break;                          \
    case 3: CALL(3); break;                          \
    case 4: CALL(4); break;                          \
    case 5: CALL(5); break;                          \
    case 6: CALL(6); break;                          \
    case 7: CALL(7); break;                          \
    case 8: CALL(8); break;                          \
    default: CALL(0); break;                         \
    }

// Unweighted pass 1 on the INT8 tensor cores: G = K^T K (mp x mp, both triangles), b1 = K^T rvec.
int gauss_pass1_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *rvec, double *G, double *b1)
{
    cudaStream_t s = ctx->stream;
    static DeviceOnce once;
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(i8_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int mp = w->mp, m = w->m, d = w->d;
    const int tiles = w->nt * (w->nt + 1);
    const int nsplit = std::max(1, std::min(16, ctx->sm_count / tiles));
    const int quantum = BK * nsplit;
    // rows per launch: the chunk buffer holds 8 slices x rows x mp bytes (= the FP64 chunk it replaces), and one
    // INT32 accumulator may sum at most MAX_ROWS_PER_SPLIT rows
    int64_t rows1 = std::min<int64_t>((int64_t)w->chunk_elems / mp, (int64_t)MAX_ROWS_PER_SPLIT * nsplit);
    rows1 = std::max<int64_t>(quantum, rows1 / quantum * quantum);
    SRGP_TRY(w->Gpart.reserve((size_t)tiles * nsplit * BM * BN * 8));
    int first = 1;
    if (ctx->n == 0) {
        SRGP_CUDA(cudaMemsetAsync(w->Gpart.p, 0, (size_t)tiles * nsplit * BM * BN * 8, s));
        SRGP_CUDA(cudaMemsetAsync(w->b1part.p, 0, (size_t)w->gen_groups * mp * 8, s));
    }
    cudaStream_t sg = getenv("SRGP_NO_OVERLAP") ? s : ctx->stream3;
    SRGP_CUDA(cudaEventRecord(ctx->ev_fork, s));
    SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_fork, 0));
    int cidx = 0;
    for (int64_t r0 = 0; r0 < ctx->n; r0 += rows1, cidx++) {
        const int rows_valid = (int)std::min<int64_t>(rows1, ctx->n - r0);
        const int rows_padded = (int)round_up(rows_valid, quantum);
        const int b = cidx & 1;
        int8_t *slices = reinterpret_cast<int8_t *>(w->chunk.d() + (size_t)b * w->chunk_elems);
        const size_t slice_stride = (size_t)rows_padded * mp;
        if (cidx >= 2) SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_used[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GEN, sg);
            dim3 grid(mp / 128, w->gen_groups);
            const size_t smem = sizeof(double) * BK * (d + 1);
#define CALL(D) launch_gen_knotrows<D>(sg, grid, smem, ctx->Xp, ctx->n, rvec, r0, rows_valid, rows_padded, w->U.d(), m, mp, d, gp, slices, slice_stride, w->b1part.d(), first)
            SRGP_D_SWITCH_I8(d, CALL)
#undef CALL
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_gen[b], sg));
        SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_gen[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GRAM, s);
            i8_gram_kernel<<<tiles * nsplit, THREADS, SMEM_BYTES, s>>>(slices, slice_stride, rows_padded / BK, nsplit,
                                                                      gp.sigma2 * gp.sigma2, w->Gpart.d(), first);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_used[b], s));
        first = 0;
    }
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        i8_gram_finalize_kernel<<<tiles, 128, 0, s>>>(w->Gpart.d(), nsplit, mp, G);
        SRGP_LAUNCH_CHECK();
        gram_sum_rows(s, w->b1part.d(), w->gen_groups, mp, b1);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

}  // namespace srgp
