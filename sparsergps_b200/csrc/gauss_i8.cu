// gauss_i8.cu -- the row passes of the fused Gaussian / Laplace pipelines on the INT8 tensor cores (tcgen05 + TMEM).
// Reference functions served: the n x m products inside elbo_fun / delbo_dcov_par (R/vi_functions.R:64-121, 126-420),
// obj_fun_norm / dlogp_dcov_par (R/laplace_approx_obj_funs.R:6-52, R/laplace_approx_gradient.R:720-968) and
// dlogq_dcov_par (R/laplace_approx_gradient.R:25-339); the algebra is in gauss_vi.cu / gauss_fic.cu / laplace.cu,
// the scheme and its error bound in tc_i8.cuh and DESIGN.md section 3a.
//
// Pass 1: G = K^T diag(w) K (w optional) and b1 = K^T r.  The generator emits k_ij / sigma^2 = exp(-d_ij^2 / 2) in
// (0, 1] as 8 INT8 digit slices, already in the shared-memory operand image of tc_i8.cuh (plus a second slice set
// w_i k_ij / (sigma^2 2^ew) for a weighted Gram); the Gram kernel runs one CTA per (128 x 64 tile of the lower block
// triangle, row split), keeps all 8 significance levels of the tile in TMEM over its row range, converts
// INT32 -> FP64 once per launch and adds into its own slot (deterministic).  Everything is exact except the final
// FP64 summation of levels, splits and chunks, so the result is at least as accurate as the DMMA SYRK it replaces
// (profiles/r01_ozaki_*.json).
// Pass 2 and the row forms: T = K Mop^T with Mop sliced per output column; one epilogue thread per data row forms the
// gradient sums (gauss_pass2), the per-row / per-dimension sums (gauss_rowd) or the row quadratic forms (gauss_rowform).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "fastexp.cuh"
#include "gauss.cuh"
#include "gauss_i8.cuh"
#include "tc_i8.cuh"

namespace srgp {

using namespace i8;

// ------------------------------------------------------------------------------------------------
// generator for pass 1: operand rows = knots (block = 128 knots), k index = data rows of the chunk
// ------------------------------------------------------------------------------------------------
// power of two >= |x| (1 for x = 0 or a non-finite x): the scale of a weighted operand
__device__ __forceinline__ double pow2_ceil(double x)
{
    return (x > 0.0 && x < INFINITY) ? scalbn(1.0, ilogb(x) + 1) : 1.0;
}

// WEIGHTED: a second slice set holds w_i e_ij / 2^ew (2^ew >= max_i |w_i|, *wmax on the device) for the weighted Grams
// K^T diag(w) K = (diag(w) K)^T K of the FIC model; w may have either sign.
template <int DT, bool WEIGHTED>
__global__ void __launch_bounds__(128)
gen_slices_knotrows_kernel(const double *__restrict__ X, int64_t ldx, const double *__restrict__ r, int64_t r0,
                           int rows_valid, int rows_padded, const double *__restrict__ U, int m, int mp, int d_rt,
                           GenParams p, int8_t *__restrict__ slices, size_t slice_stride, double *__restrict__ b1part,
                           int first, const double *__restrict__ rw, const double *__restrict__ wmax,
                           int8_t *__restrict__ slices_w)
{
    extern __shared__ double sx[];   // [64][d] scaled rows, then [64] residuals, then [64] scaled row weights
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, 128);
    const int d = DT > 0 ? DT : d_rt;
    double *sr = sx + 64 * d;
    double *sw = sr + 64;
    const double winv = WEIGHTED ? 1.0 / pow2_ceil(*wmax) : 0.0;
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
            if (WEIGHTED) sw[threadIdx.x] = (i < rows_valid) ? rw[r0 + i] * winv : 0.0;
        }
        __syncthreads();
        const size_t img = ((size_t)blockIdx.x * KB + kb) * IMG_BLOCK + (size_t)threadIdx.x * 16;
        int8_t *dst = slices + img;
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16++) {
            uint32_t w[NS][4], ww[WEIGHTED ? NS : 1][4];
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
                double ev[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    ev[q] = 0.0;
                    if (jvalid && it0 + ii + q < rows_valid) {
                        ev[q] = exp_tab(-0.5 * sq[q], etab);
                        bacc = fma(p.sigma2 * ev[q], sr[ii + q], bacc);
                    }
                }
                split_quad(ev[0], ev[1], ev[2], ev[3], e0 >> 2, w);
                if (WEIGHTED)
                    split_quad(ev[0] * sw[ii], ev[1] * sw[ii + 1], ev[2] * sw[ii + 2], ev[3] * sw[ii + 3], e0 >> 2,
                               reinterpret_cast<uint32_t (&)[NS][4]>(ww));
            }
#pragma unroll
            for (int s = 0; s < NS; s++) {
                *reinterpret_cast<uint4 *>(dst + s * slice_stride + c16 * 2048) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
                if (WEIGHTED)
                    *reinterpret_cast<uint4 *>(slices_w + img + s * slice_stride + c16 * 2048) =
                        make_uint4(ww[s][0], ww[s][1], ww[s][2], ww[s][3]);
            }
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
i8_gram_kernel(const int8_t *__restrict__ slices_a, const int8_t *__restrict__ slices, size_t slice_stride, int KB,
               int nsplit, double scale, const double *__restrict__ wmax, double *__restrict__ Gpart, int first)
{   // slices_a: A operand (rows 128 I ..): the weighted slice set, or `slices` itself; slices: B operand (columns 64 J ..)
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
            for (int it = 0; it < 2 * kb_per; ++it) {
                const int st = it % STAGES;
                if (it >= STAGES) mbar_wait(&bars.empty[st], ((it / STAGES) - 1) & 1);
                load_stage(smem_u32(smem + st * STAGE_BYTES), &bars.full[st], slices_a, slice_stride,
                           ((size_t)I * KB + kb0) * IMG_BLOCK, slices, slice_stride, ((size_t)(J >> 1) * KB + kb0) * IMG_BLOCK,
                           J & 1, it);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < 2 * kb_per; ++it) {
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
        if (wmax) scale *= pow2_ceil(*wmax);
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

// ------------------------------------------------------------------------------------------------
// Gram kernel on 128 x 128 tiles, two sweeps (tc_i8.cuh): slot[tile][split] (+)= scale * sum_L 2^(-12-8L) level_L.
// tile t -> (I, J), J <= I, 128-row / 128-column knot blocks; on the diagonal tiles of an unweighted Gram the A and B
// operands are the same block and are loaded once.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile2_to_ij(int t, int &I, int &J)
{
    I = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((I + 1) * (I + 2) / 2 <= t) I++;
    while (I * (I + 1) / 2 > t) I--;
    J = t - I * (I + 1) / 2;
}

constexpr int GRAM2_STAGES = 4;
constexpr int GRAM2_SMEM = GRAM2_STAGES * STAGE2_BYTES + (int)sizeof(Bars) + 16;

__global__ void __launch_bounds__(THREADS, 1)
i8_gram2_kernel(const int8_t *__restrict__ slices_a, const int8_t *__restrict__ slices, size_t slice_stride, int KB,
                int nsplit, double scale, const double *__restrict__ wmax, double *__restrict__ Gpart, int first)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars &bars = *reinterpret_cast<Bars *>(smem + GRAM2_STAGES * STAGE2_BYTES);
    uint64_t *tmem_empty = reinterpret_cast<uint64_t *>(smem + GRAM2_STAGES * STAGE2_BYTES + sizeof(Bars));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int I, J;
    tile2_to_ij(blockIdx.x / nsplit, I, J);
    const int split = blockIdx.x % nsplit;
    const int kb_per = KB / nsplit, kb0 = split * kb_per, ksteps = 2 * kb_per;
    const bool same = (I == J) && (slices_a == slices);

    if (threadIdx.x == 0) {
        for (int s = 0; s < GRAM2_STAGES; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        mbar_init(&bars.tmem_full, 1);
        mbar_init(tmem_empty, 4);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_all(&bars.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;
            for (int sw = 0; sw < 2; ++sw)
                for (int it = 0; it < ksteps; ++it, ++g) {
                    const int st = g % GRAM2_STAGES;
                    if (g >= GRAM2_STAGES) mbar_wait(&bars.empty[st], ((g / GRAM2_STAGES) - 1) & 1);
                    load_stage2(smem_u32(smem + st * STAGE2_BYTES), &bars.full[st], slices_a, slice_stride,
                                ((size_t)I * KB + kb0) * IMG_BLOCK, slices, slice_stride, ((size_t)J * KB + kb0) * IMG_BLOCK, same,
                                it, sweep_slices(sw));
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int g = 0;
            for (int sw = 0; sw < 2; ++sw) {
                if (sw > 0) {                           // the epilogue must have drained sweep 0 from TMEM
                    mbar_wait(tmem_empty, 0);
                    tc_fence_after();
                }
                for (int it = 0; it < ksteps; ++it, ++g) {
                    const int st = g % GRAM2_STAGES;
                    mbar_wait(&bars.full[st], (g / GRAM2_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem + st * STAGE2_BYTES);
                    const uint32_t b_base = same ? a_base : a_base + NS * A_TILE;
                    if (sw == 0) issue_stage2<0>(a_base, b_base, tmem_base, it == 0);
                    else issue_stage2<1>(a_base, b_base, tmem_base, it == 0);
                    mma_commit(&bars.empty[st]);
                }
                mma_commit(&bars.tmem_full);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        if (wmax) scale *= pow2_ceil(*wmax);
        double *out = Gpart + ((size_t)blockIdx.x * BM + row) * BN2;
#pragma unroll 1
        for (int sw = 0; sw < 2; ++sw) {
            mbar_wait(&bars.tmem_full, sw & 1);
            tc_fence_after();
            const int nl = sweep_levels(sw), l0 = 4 * sw;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                double acc[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c] = 0.0;
#pragma unroll 1
                for (int k = nl - 1; k >= 0; --k) {        // least significant level first
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * BN2 + c4 * 32), v);
                    const double wgt = scale * exp2(-12.0 - 8.0 * (l0 + k));
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = fma(wgt, (double)(int)v[c], acc[c]);
                }
                double2 *o2 = reinterpret_cast<double2 *>(out + c4 * 32);
                const bool overwrite = first && sw == 0;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    double2 prev = overwrite ? make_double2(0.0, 0.0) : o2[c];
                    o2[c] = make_double2(prev.x + acc[2 * c], prev.y + acc[2 * c + 1]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

// Sum the split slots of the 128 x 128 tiles and scatter to the full symmetric matrix (column-major, ld = mp).
__global__ void __launch_bounds__(128)
i8_gram2_finalize_kernel(const double *__restrict__ Gpart, int nsplit, int mp, double *__restrict__ G)
{
    int I, J;
    tile2_to_ij(blockIdx.x, I, J);
    const int r = I * BM + threadIdx.x;
    const double *base = Gpart + ((size_t)blockIdx.x * nsplit * BM + threadIdx.x) * BN2;
    for (int c = 0; c < BN2; ++c) {
        double v = 0.0;
        for (int s = 0; s < nsplit; ++s) v += base[(size_t)s * BM * BN2 + c];
        const int col = J * BN2 + c;
        G[r + (int64_t)col * mp] = v;
        if (J < I) G[col + (int64_t)r * mp] = v;          // strictly below the diagonal block: mirror
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2 on the INT8 tensor cores:  T = K Mop^T  (T_ij' = sum_j K_ij Mop[j' + j mp]), then the fused
// "never materialise dK" reduction of km_reduce_kernel<MODE_GRAD> (gauss.cu):
//   Omega_ij = rs_i T_ij + ra_i beta_j ;  P_ij = Omega_ij K_ij ;  slot[0] += sum P ;  slot[1 + c] += sum P D_ijc
// K enters the MMA as digit slices of exp(-d^2/2) (generator below); Mop as digit slices of Mop[n, :] / 2^e_n with
// one power-of-two scale per output column n (slice_mop_kernel).  The epilogue owns one data row per thread: it
// drains the 8 levels of 32 columns from TMEM, rebuilds K_ij and D_ijc from the row's coordinates in registers
// and the tile's knots in shared memory (the FP64 pipe is idle while the tensor cores run INT8), and keeps the
// 1 + d sums in registers across all column tiles of the CTA.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(128)
gen_slices_datarows_kernel(const double *__restrict__ X, int64_t ldx, int64_t r0, int rows_valid,
                           const double *__restrict__ U, int m, int mp, GenParams p, int8_t *__restrict__ slices,
                           size_t slice_stride)
{
    extern __shared__ double su[];   // [64][DT] scaled knots of one k-block
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, 128);
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool ivalid = i < rows_valid;
    double xi[DT];
#pragma unroll
    for (int c = 0; c < DT; c++) xi[c] = ivalid ? X[r0 + i + ldx * c] * p.invl[c] : 0.0;
    const int KBm = mp / BK;
    for (int kb = blockIdx.y; kb < KBm; kb += gridDim.y) {
        __syncthreads();
        for (int t = threadIdx.x; t < BK * DT; t += 128) {
            const int jj = t / DT, c = t - jj * DT;
            const int j = kb * BK + jj;
            su[t] = (j < m) ? U[j + (int64_t)m * c] * p.invl[c] : 0.0;
        }
        __syncthreads();
        int8_t *dst = slices + ((size_t)blockIdx.x * KBm + kb) * IMG_BLOCK + (size_t)threadIdx.x * 16;
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16++) {
            uint32_t w[NS][4];
#pragma unroll
            for (int s = 0; s < NS; s++) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
            for (int e0 = 0; e0 < 16; e0 += 4) {
                double sq[4] = {0.0, 0.0, 0.0, 0.0};
                const int jj = c16 * 16 + e0;
#pragma unroll
                for (int c = 0; c < DT; c++) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const double t = xi[c] - su[(jj + q) * DT + c];
                        sq[q] = fma(t, t, sq[q]);
                    }
                }
                double ev[4];
#pragma unroll
                for (int q = 0; q < 4; q++) ev[q] = (ivalid && kb * BK + jj + q < m) ? exp_tab(-0.5 * sq[q], etab) : 0.0;
                split_quad(ev[0], ev[1], ev[2], ev[3], e0 >> 2, w);
            }
#pragma unroll
            for (int s = 0; s < NS; s++)
                *reinterpret_cast<uint4 *>(dst + s * slice_stride + c16 * 2048) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        }
    }
}

// colscale[n] = 2^e_n with max_k |Mop[n, k]| / 2^e_n in [0.5, 1): one pass over Mop (element (n, k) at n + k mp), the k
// range split over blockIdx.y and merged with an integer atomicMax on the bit patterns (non-negative doubles order like
// their bits).  colbits must be zeroed before the launch.  grid (mp / 128, 16), 128 threads.
__global__ void __launch_bounds__(128)
mop_rowmax_kernel(const double *__restrict__ Mop, int mp, unsigned long long *__restrict__ colbits)
{
    const int n = blockIdx.x * 128 + threadIdx.x;
    const int per = mp / gridDim.y, k0 = blockIdx.y * per;
    double mx = 0.0;
    for (int k = k0; k < k0 + per; k++) mx = fmax(mx, fabs(Mop[n + (int64_t)k * mp]));     // fmax drops NaN
    atomicMax(colbits + n, (unsigned long long)__double_as_longlong(mx));
}

// Mop -> digit slices of Mop[n, :] / 2^e_n in the operand image (rows = n); colscale[n] = 2^e_n replaces the max bits.
// grid (mp / 128, mp / 64), 128 threads.
__global__ void __launch_bounds__(128)
slice_mop_kernel(const double *__restrict__ Mop, int mp, int8_t *__restrict__ slices, size_t slice_stride,
                 const unsigned long long *__restrict__ colbits, double *__restrict__ colscale)
{
    const int n = blockIdx.x * 128 + threadIdx.x, kb = blockIdx.y;
    const double mx = __longlong_as_double((long long)colbits[n]);
    int ex = 0;
    if (mx > 0.0 && mx < INFINITY) ex = ilogb(mx) + 1;
    const double inv = scalbn(1.0, -ex);
    if (kb == 0) colscale[n] = scalbn(1.0, ex);
    const int KBm = mp / BK;
    int8_t *dst = slices + ((size_t)blockIdx.x * KBm + kb) * IMG_BLOCK + (size_t)threadIdx.x * 16;
#pragma unroll 1
    for (int c16 = 0; c16 < 4; c16++) {
        uint32_t w[NS][4];
#pragma unroll
        for (int s = 0; s < NS; s++) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; k++)               // NaN / Inf (failed factorisation upstream) -> finite; info flags report it
                v[k] = fmin(1.0, fmax(-1.0, Mop[n + (int64_t)(kb * BK + c16 * 16 + e + k) * mp] * inv));
            split_quad(v[0], v[1], v[2], v[3], e >> 2, w);
        }
#pragma unroll
        for (int s = 0; s < NS; s++)
            *reinterpret_cast<uint4 *>(dst + s * slice_stride + c16 * 2048) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

struct KmI8Args {
    const int8_t *kslices;   // K digit slices of the chunk: blocks = 128-row blocks, k = knots
    size_t kstride;
    const int8_t *mslices;   // Mop digit slices: blocks = 128 output columns n, k = knots
    size_t mstride;
    const double *colscale;  // mp
    int KBm, mp, m;
    const double *X;         // resident rows (column-major, ld = ldx), chunk starts at r0
    int64_t ldx, r0;
    int rows_valid;
    const double *U;         // m x d
    const double *rs, *ra, *beta;
    double invl[8];
    double sigma2;
    int tiles_per_cta;       // 64-column tiles per CTA and row block
    int nsub;                // row blocks per CTA (processed one after the other): rb = sub * gridDim.x + blockIdx.x
    double *part;            // [gridDim.y][gridDim.x][PART_STRIDE_I8] accumulated across launches
    int first;
    int *coin_count, *coin_list;
    double *coin_omega;
    int coin_cap;
    // ROWD mode (per-row, per-dimension sums of the FIC model, see KmArgs::rowd_part in gauss.cu):
    //   rowd_part[((group * nslots + slot) * ld + row], group = blockIdx.y * 2 + column half of the epilogue thread
    //   slot c (0..d): sum_j T_ij K_ij D_ijc (D_ij0 = 1); slot d+1+c (beta given): sum_j beta_j K_ij D_ijc;
    //   last slot (vvec given): sum_j K_ij v_j.  Coincident pairs are recorded with T_ij.
    const double *vvec;
    double *rowd_part;
    int nslots;
    int64_t ld;
    int nodims;              // ROWD without the per-dimension slots (gauss_rowform): slot 0 = sum_j T_ij K_ij, slot 1 = K v
    int cluster;             // i8_km2_kernel: 2 = adjacent row blocks run as tcgen05 CTA pairs (cta_group::2), 1 = single CTAs
    int debug;               // measurement only (SRGP_KM_DEBUG): bit 0 = skip the FP64 epilogue arithmetic
};

// Rare path of quirk Q4 (same contract as record_if_coincident in gauss.cu): decided by the reference's own test,
// all coordinates bit-identical (src/covariance_function_derivativesC.cpp:157-163).
__device__ __noinline__ void record_if_coincident_i8(const double *X, int64_t ldx, int64_t i_shard, const double *U, int m,
                                                     int j, int d, int *coin_count, int *coin_list, double *coin_omega,
                                                     int coin_cap, double omega_ij)
{
    for (int c = 0; c < d; c++)
        if (X[i_shard + ldx * c] != U[j + (int64_t)m * c]) return;
    const int slot = atomicAdd(coin_count, 1);
    if (slot < coin_cap) {
        coin_list[2 * slot] = (int)i_shard;
        coin_list[2 * slot + 1] = j;
        coin_omega[slot] = omega_ij;
    }
}

// pass-2 CTA: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue.  Two epilogue warps share a TMEM lane
// quadrant (32 rows) and take 32 of the tile's 64 columns each, so every thread drains its 8 levels x 32 columns at
// once, releases TMEM, and only then does the FP64 work -- which overlaps the next tile's MMAs.
constexpr int KM_THREADS = 320;
constexpr int KM_EPI_THREADS = 256;
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int DT, bool ROWD>
__global__ void __launch_bounds__(KM_THREADS, 1) i8_km_kernel(KmI8Args a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars &bars = *reinterpret_cast<Bars *>(smem + STAGES * STAGE_BYTES);
    uint64_t *tmem_empty = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + sizeof(Bars));
    double *us = reinterpret_cast<double *>(tmem_empty + 2);   // [64][DT] scaled knots of the current column tile
    double *bt = us + BN * DT;                                  // [64] beta
    double *cs = bt + BN;                                       // [64] sigma^2 * column scale
    double *vv = cs + BN;                                       // [64] v (ROWD)
    double *red = vv + BN;                                      // [8][PART_STRIDE_I8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jt0 = blockIdx.y * a.tiles_per_cta;
    const int KBm = a.KBm;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.empty[s], 1);
        }
        mbar_init(&bars.tmem_full, 1);
        mbar_init(tmem_empty, 8);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_all(&bars.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tt = 0; tt < a.nsub * a.tiles_per_cta; ++tt) {
                const int jt = jt0 + tt % a.tiles_per_cta;
                const int rb = (tt / a.tiles_per_cta) * gridDim.x + blockIdx.x;
                for (int ks = 0; ks < 2 * KBm; ++ks, ++it) {
                    const int st = it % STAGES;
                    if (it >= STAGES) mbar_wait(&bars.empty[st], ((it / STAGES) - 1) & 1);
                    load_stage(smem_u32(smem + st * STAGE_BYTES), &bars.full[st], a.kslices, a.kstride,
                               (size_t)rb * KBm * IMG_BLOCK, a.mslices, a.mstride, (size_t)(jt >> 1) * KBm * IMG_BLOCK, jt & 1, ks);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < a.nsub * a.tiles_per_cta; ++t) {
                if (t > 0) {                           // the epilogue must have drained the previous tile from TMEM
                    mbar_wait(tmem_empty, (t - 1) & 1);
                    tc_fence_after();
                }
                for (int ks = 0; ks < 2 * KBm; ++ks, ++it) {
                    const int st = it % STAGES;
                    mbar_wait(&bars.full[st], (it / STAGES) & 1);
                    tc_fence_after();
                    issue_stage(smem_u32(smem + st * STAGE_BYTES), tmem_base, ks == 0);
                    mma_commit(&bars.empty[st]);
                }
                mma_commit(&bars.tmem_full);
            }
        }
    } else {
        // ===== epilogue: thread = one data row of the block x 32 columns of the tile =====
        const int ew = warp - 2, q = warp & 3, half = ew >> 2, et = threadIdx.x - 64;
        const int row = q * 32 + lane;
        double s0 = 0.0, sc[DT];
#pragma unroll
        for (int c = 0; c < DT; c++) sc[c] = 0.0;
        double rt[ROWD ? DT + 1 : 1], rbt[ROWD ? DT + 1 : 1], rkv = 0.0;   // ROWD: this thread's row sums over its columns
#pragma unroll
        for (int c = 0; c < (ROWD ? DT + 1 : 1); c++) rt[c] = rbt[c] = 0.0;
        int i = 0;
        bool iv = false;
        int64_t ig = 0;
        double xi[DT], rsi = 1.0, rai = 0.0;
        for (int t = 0; t < a.nsub * a.tiles_per_cta; ++t) {
            if (t % a.tiles_per_cta == 0) {                     // next row block of this CTA
                i = ((t / a.tiles_per_cta) * gridDim.x + blockIdx.x) * BM + row;
                iv = i < a.rows_valid;
                ig = a.r0 + i;
#pragma unroll
                for (int c = 0; c < DT; c++) xi[c] = iv ? a.X[ig + a.ldx * c] * a.invl[c] : 0.0;
                rsi = (a.rs && iv) ? a.rs[ig] : 1.0;
                rai = (a.ra && iv) ? a.ra[ig] : 0.0;
            }
            const int j0 = (jt0 + t % a.tiles_per_cta) * BN;
            epi_bar();                                          // everyone is done with the previous tile's us / bt / cs
            for (int e = et; e < BN * DT; e += KM_EPI_THREADS) {
                const int jj = e / DT, c = e - jj * DT;
                us[e] = (j0 + jj < a.m) ? a.U[j0 + jj + (int64_t)a.m * c] * a.invl[c] : 0.0;
            }
            if (et < BN) {
                bt[et] = (a.beta && j0 + et < a.m) ? a.beta[j0 + et] : 0.0;
                cs[et] = a.sigma2 * a.colscale[j0 + et];
                if (ROWD) vv[et] = (a.vvec && j0 + et < a.m) ? a.vvec[j0 + et] : 0.0;
            }
            epi_bar();
            mbar_wait(&bars.tmem_full, t & 1);
            tc_fence_after();
            double T[32];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32 + g * 16);
                long long acc[16];
                drain16<NS - 4>(tcol + 4 * BN, acc);            // levels 4..NS-1
#pragma unroll
                for (int c = 0; c < 16; ++c) T[g * 16 + c] = W_LEVELS_LO * (double)acc[c];
                drain16<4>(tcol, acc);                          // levels 0..3
#pragma unroll
                for (int c = 0; c < 16; ++c) T[g * 16 + c] = fma(W_LEVELS_HI, (double)acc[c], T[g * 16 + c]);
            }
            tc_fence_before();                                  // TMEM is free for the next tile's MMAs
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            if (iv) {
                // K_ij / sigma^2 comes back from the generator's digit slices (exact to 2^-62; 8 coalesced 16-byte loads
                // per 16 columns) instead of being recomputed: exp and the distance chain are ~33 FP64 instructions per
                // entry, and FP64 work is what competes with the tensor pipe (profiles/r01_ncu_i8_km.txt).
                const int8_t *kimg = a.kslices + ((size_t)(i / BM) * KBm) * IMG_BLOCK + (size_t)row * 16;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int jg = j0 + half * 32 + g * 16;     // first column of this 16-column group
                    uint4 w[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        w[s] = __ldg(reinterpret_cast<const uint4 *>(kimg + s * a.kstride + (size_t)(jg / BK) * IMG_BLOCK +
                                                                     (size_t)((jg % BK) / 16) * 2048));
#pragma unroll                                                  // T[] stays in registers only if c is a compile-time index
                    for (int e = 0; e < 16; ++e) {
                        const int jj = half * 32 + g * 16 + e;
                        long long q4[4];
                        if ((e & 3) == 0) join_quad(w, e >> 2, q4);
                        if (ROWD) {
                            if (j0 + jj < a.m) {
                                const long long qd = q4[e & 3];
                                const double kij = a.sigma2 * FIX_INV * (double)qd, tij = cs[jj] * T[g * 16 + e];
                                const double tk = tij * kij, bk = bt[jj] * kij;
                                rt[0] += tk;
                                rbt[0] += bk;
                                rkv = fma(kij, vv[jj], rkv);
                                if (!a.nodims) {
#pragma unroll
                                    for (int k = 0; k < DT; k++) {
                                        const double tt = xi[k] - us[jj * DT + k], d2 = tt * tt;
                                        rt[(ROWD ? 1 + k : 0)] = fma(tk, d2, rt[(ROWD ? 1 + k : 0)]);
                                        rbt[(ROWD ? 1 + k : 0)] = fma(bk, d2, rbt[(ROWD ? 1 + k : 0)]);
                                    }
                                }
                                if (qd == FIX_ONE && !a.nodims)   // the row forms (gauss_rowform) record no pairs
                                    record_if_coincident_i8(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                            a.coin_omega, a.coin_cap, tij);
                            }
                        } else if (j0 + jj < a.m) {
                            const long long qd = q4[e & 3];
                            const double om = fma(rsi, cs[jj] * T[g * 16 + e], rai * bt[jj]);
                            const double pk = om * (a.sigma2 * FIX_INV * (double)qd);
                            s0 += pk;
#pragma unroll
                            for (int k = 0; k < DT; k++) {
                                const double tt = xi[k] - us[jj * DT + k];
                                sc[k] = fma(pk, tt * tt, sc[k]);
                            }
                            if (qd == FIX_ONE)                  // exp(0) = 1: candidate for the bit-identical test
                                record_if_coincident_i8(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                        a.coin_omega, a.coin_cap, om);
                        }
                    }
                }
            }
        }
        if (ROWD) {
            // one partial per (column group, column half): nothing to reduce across threads, stores only
            double *part = a.rowd_part + ((int64_t)(blockIdx.y * 2 + half) * a.nslots) * a.ld + i;
            if (a.nodims) {
                part[0] = rt[0];
            } else {
#pragma unroll
                for (int c = 0; c < DT + 1; c++) {
                    part[(int64_t)c * a.ld] = rt[ROWD ? c : 0];
                    if (a.beta) part[(int64_t)(DT + 1 + c) * a.ld] = rbt[ROWD ? c : 0];
                }
            }
            if (a.vvec) part[(int64_t)(a.nslots - 1) * a.ld] = rkv;
        }
        // CTA reduction: warp shuffles, then the 8 epilogue warps through shared memory -> this CTA's slot
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
#pragma unroll
        for (int c = 0; c < DT; c++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sc[c] += __shfl_xor_sync(0xffffffffu, sc[c], o);
        if (lane == 0) {
            red[ew * PART_STRIDE_I8] = s0;
#pragma unroll
            for (int c = 0; c < DT; c++) red[ew * PART_STRIDE_I8 + 1 + c] = sc[c];
        }
        epi_bar();
        if (!ROWD && et < 1 + DT) {
            double v = 0.0;
            for (int k = 0; k < 8; k++) v += red[k * PART_STRIDE_I8 + et];
            double *slot = a.part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * PART_STRIDE_I8 + et;
            *slot = a.first ? v : (*slot + v);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2 on 128 x 128 tiles, two sweeps per column block (tc_i8.cuh): every quantity the epilogue forms is linear in
// T = K Mop^T, so the two level groups of a tile are two "virtual tiles" whose contributions add; the terms that do not
// involve T (ra_i beta_j, the beta / v row sums of the ROWD mode) ride on sweep 0 only.
// CTA = 384 threads = 3 warpgroups: warp 0 TMA producer, warp 1 MMA issuer (warps 2, 3 idle), warps 4..11 epilogue.  The
// epilogue holds the 64 drained columns of its row in registers (so TMEM is released before the FP64 work starts):
// setmaxnreg moves registers from the first warpgroup to the two epilogue warpgroups.
// Operand ring: 5 units of 32 KB = 4 A-slice tiles + 4 B-slice tiles of one 32-byte k-step.  Sweep 0 needs slices 0..3 =
// one unit per k-step; sweep 1 needs all NS slices = two units (slices 0..3, then 4..NS-1).
// ------------------------------------------------------------------------------------------------
constexpr int KM2_THREADS = 384;
constexpr int KM2_EPI_REGS = 232, KM2_AUX_REGS = 40;
constexpr int KM2_MAX_UNITS = 7;
// operand ring of one CTA: units of 4 A-slice tiles (4 KB each) + 4 B-slice tiles of one 32-byte k-step.
//   single CTA: B tiles are 128 columns x 32 B = 4 KB -> 32 KB units, 5 of them;
//   CTA pair (cta_group::2): each CTA holds the 64-column half of the B tiles (2 KB) -> 24 KB units, 7 of them.
template <bool PAIR> struct Km2Cfg {
    static constexpr int B_TILE_BYTES = PAIR ? A_TILE / 2 : A_TILE;
    static constexpr int UNIT_BYTES = 4 * A_TILE + 4 * B_TILE_BYTES;
    static constexpr int UNITS = PAIR ? 7 : 5;
    static constexpr int RING_BYTES = UNITS * UNIT_BYTES;
};

struct Bars2 {
    uint64_t full[KM2_MAX_UNITS], empty[KM2_MAX_UNITS], tmem_full, tmem_empty;
    uint32_t tmem_slot, pad;
};

// a record of quirk Q4 that carries only the part of Omega_ij that is linear in T (the second sweep of a pair): the knot
// index is stored as ~j and coin_fix* add no T-free term for it
__device__ __noinline__ void record_if_coincident_i8_part(const double *X, int64_t ldx, int64_t i_shard, const double *U, int m,
                                                          int j, int d, int *coin_count, int *coin_list, double *coin_omega,
                                                          int coin_cap, double omega_ij, int partial)
{
    for (int c = 0; c < d; c++)
        if (X[i_shard + ldx * c] != U[j + (int64_t)m * c]) return;
    const int slot = atomicAdd(coin_count, 1);
    if (slot < coin_cap) {
        coin_list[2 * slot] = (int)i_shard;
        coin_list[2 * slot + 1] = partial ? ~j : j;
        coin_omega[slot] = omega_ij;
    }
}

// PAIR: the two CTAs of a cluster (adjacent row blocks, same column group) form one tcgen05 CTA pair.
template <int DT, bool ROWD, bool PAIR>
__global__ void __launch_bounds__(KM2_THREADS, 1) i8_km2_kernel(KmI8Args a)
{
    using Cfg = Km2Cfg<PAIR>;
    constexpr int NU = Cfg::UNITS, UB = Cfg::UNIT_BYTES, BT = Cfg::B_TILE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars2 &bars = *reinterpret_cast<Bars2 *>(smem + Cfg::RING_BYTES);
    double *us = reinterpret_cast<double *>(smem + Cfg::RING_BYTES + sizeof(Bars2));   // [128][DT] scaled knots
    double *bt = us + BN2 * DT;                                 // [128] beta
    double *cs = bt + BN2;                                      // [128] sigma^2 * column scale
    double *vv = cs + BN2;                                      // [128] v (ROWD)
    double *red = vv + BN2;                                     // [8][PART_STRIDE_I8]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb0 = blockIdx.y * a.tiles_per_cta;               // first 128-column block of this CTA
    const int KBm = a.KBm, ksteps = 2 * KBm;
    const int nvt = 2 * a.tiles_per_cta;                        // virtual tiles: (column block, sweep)
    const int rb = blockIdx.x;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;       // 0 = leader of the pair
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        for (int u = 0; u < NU; ++u) {
            // the leader of a pair issues for both CTAs: its "full" also counts the peer's relay (one wait per unit for
            // the single issuing thread, whose waits are on the critical path of every k-step)
            mbar_init(&bars.full[u], (PAIR && leader) ? 2 : 1);
            mbar_init(&bars.empty[u], 1);
        }
        mbar_init(&bars.tmem_full, 1);
        mbar_init(&bars.tmem_empty, PAIR ? 16 : 8);             // epilogue warps of every CTA that shares the MMAs
        mbar_fence_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_all_pair(&bars.tmem_slot);
        else tmem_alloc_all(&bars.tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                               // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(KM2_AUX_REGS));
        if (warp == 0 && lane == 0) {
            // ===== TMA producer (every CTA: its own K rows, its share of the Mop columns) =====
            const size_t a_blk = (size_t)rb * KBm * IMG_BLOCK;
            int g = 0;                                          // unit counter
            for (int vt = 0; vt < nvt; ++vt) {
                const int sw = vt & 1;
                const size_t b_blk = (size_t)(jb0 + (vt >> 1)) * KBm * IMG_BLOCK;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const size_t off = (size_t)(ks >> 1) * IMG_BLOCK + (size_t)(ks & 1) * A_TILE;
                    for (int part = 0; part <= sw; ++part, ++g) {
                        const int u = g % NU;
                        if (g >= NU) mbar_wait(&bars.empty[u], ((g / NU) - 1) & 1);
                        const int s0 = 4 * part, ns = part == 0 ? 4 : NS - 4;
                        const uint32_t ub = smem_u32(smem + u * UB);
                        mbar_expect_tx(&bars.full[u], (uint32_t)(ns * (A_TILE + BT)));
                        for (int s = 0; s < ns; ++s) {
                            bulk_g2s(ub + s * A_TILE, a.kslices + (size_t)(s0 + s) * a.kstride + a_blk + off, A_TILE, &bars.full[u]);
                            const int8_t *bsrc = a.mslices + (size_t)(s0 + s) * a.mstride + b_blk + off;
                            const uint32_t bdst = ub + 4 * A_TILE + s * BT;
                            if (PAIR) {                         // columns 64 crank .. 64 crank + 63: 1 KB of each 16-byte k-chunk
                                bulk_g2s(bdst, bsrc + crank * 1024, 1024, &bars.full[u]);
                                bulk_g2s(bdst + 1024, bsrc + 2048 + crank * 1024, 1024, &bars.full[u]);
                            } else {
                                bulk_g2s(bdst, bsrc, A_TILE, &bars.full[u]);
                            }
                        }
                    }
                }
            }
        } else if (warp == 1 && lane == 0 && !leader) {
            // ===== peer of a pair: tell the leader when this CTA's half of a unit has landed =====
            int g = 0;
            for (int vt = 0; vt < nvt; ++vt)
                for (int ks = 0; ks < ksteps; ++ks)
                    for (int part = 0; part <= (vt & 1); ++part, ++g) {
                        mbar_wait(&bars.full[g % NU], (g / NU) & 1);
                        mbar_arrive_remote(&bars.full[g % NU], 0);
                    }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer (the leader issues for both CTAs of a pair) =====
            constexpr uint32_t B_LBO = PAIR ? 1024 : 2048;
            int g = 0;
            for (int vt = 0; vt < nvt; ++vt) {
                const int sw = vt & 1;
                if (vt > 0) {                                   // the epilogue must have drained the previous virtual tile
                    mbar_wait(&bars.tmem_empty, (vt - 1) & 1);
                    tc_fence_after();
                }
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int u0 = g % NU;
                    mbar_wait(&bars.full[u0], (g / NU) & 1);
                    const uint32_t b0 = smem_u32(smem + u0 * UB);
                    const uint32_t keep = ks == 0 ? 0u : 1u;
                    if (sw == 0) {
                        tc_fence_after();
                        const uint64_t da0 = make_desc(b0, 2048, 128), db0 = make_desc(b0 + 4 * A_TILE, B_LBO, 128);
#pragma unroll
                        for (int sb = 0; sb < 4; ++sb)
#pragma unroll
                            for (int sa = 0; sa < 4; ++sa)
                                if (sa + sb < 4) {
                                    const uint64_t da = da0 + (uint64_t)((sa * A_TILE) >> 4), db = db0 + (uint64_t)((sb * BT) >> 4);
                                    if (PAIR) mma_i8_n128_pair(tmem_base + (uint32_t)(sa + sb) * BN2, da, db, sb == 0 ? keep : 1u);
                                    else mma_i8_n128(tmem_base + (uint32_t)(sa + sb) * BN2, da, db, sb == 0 ? keep : 1u);
                                }
                        if (PAIR) mma_commit_pair(&bars.empty[u0]);
                        else mma_commit(&bars.empty[u0]);
                        g += 1;
                    } else {
                        const int u1 = (g + 1) % NU;
                        mbar_wait(&bars.full[u1], ((g + 1) / NU) & 1);
                        tc_fence_after();
                        const uint32_t b1 = smem_u32(smem + u1 * UB);
                        const uint64_t dlo_a = make_desc(b0, 2048, 128), dlo_b = make_desc(b0 + 4 * A_TILE, B_LBO, 128);
                        const uint64_t dhi_a = make_desc(b1, 2048, 128), dhi_b = make_desc(b1 + 4 * A_TILE, B_LBO, 128);
#pragma unroll
                        for (int sb = 0; sb < NS; ++sb)
#pragma unroll
                            for (int sa = 0; sa < NS; ++sa) {
                                const int L = sa + sb;
                                if (L >= 4 && L < NS) {
                                    const uint64_t da = (sa < 4 ? dlo_a : dhi_a) + (uint64_t)(((sa & 3) * A_TILE) >> 4);
                                    const uint64_t db = (sb < 4 ? dlo_b : dhi_b) + (uint64_t)(((sb & 3) * BT) >> 4);
                                    if (PAIR) mma_i8_n128_pair(tmem_base + (uint32_t)(L - 4) * BN2, da, db, sb == 0 ? keep : 1u);
                                    else mma_i8_n128(tmem_base + (uint32_t)(L - 4) * BN2, da, db, sb == 0 ? keep : 1u);
                                }
                            }
                        if (PAIR) {
                            mma_commit_pair(&bars.empty[u0]);
                            mma_commit_pair(&bars.empty[u1]);
                        } else {
                            mma_commit(&bars.empty[u0]);
                            mma_commit(&bars.empty[u1]);
                        }
                        g += 2;
                    }
                }
                if (PAIR) mma_commit_pair(&bars.tmem_full);
                else mma_commit(&bars.tmem_full);
            }
        }
    } else {
        // ===== epilogue: thread = one data row of the block x 64 columns of the 128-column block =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(KM2_EPI_REGS));
        const int ew = warp - 4, q = warp & 3, half = ew >> 2, et = threadIdx.x - 128;
        const int row = q * 32 + lane;
        double s0 = 0.0, sc[DT];
#pragma unroll
        for (int c = 0; c < DT; c++) sc[c] = 0.0;
        double rt[ROWD ? DT + 1 : 1], rbt[ROWD ? DT + 1 : 1], rkv = 0.0;
#pragma unroll
        for (int c = 0; c < (ROWD ? DT + 1 : 1); c++) rt[c] = rbt[c] = 0.0;
        const int i = rb * BM + row;
        const bool iv = i < a.rows_valid;
        const int64_t ig = a.r0 + i;
        double xi[DT];
#pragma unroll
        for (int c = 0; c < DT; c++) xi[c] = iv ? a.X[ig + a.ldx * c] * a.invl[c] : 0.0;
        const double rsi = (a.rs && iv) ? a.rs[ig] : 1.0;
        const double rai = (a.ra && iv) ? a.ra[ig] : 0.0;
        for (int vt = 0; vt < nvt; ++vt) {
            const int sw = vt & 1;
            const int j0 = (jb0 + (vt >> 1)) * BN2;
            if (sw == 0) {
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with the previous block's us / bt / cs
                for (int e = et; e < BN2 * DT; e += 256) {
                    const int jj = e / DT, c = e - jj * DT;
                    us[e] = (j0 + jj < a.m) ? a.U[j0 + jj + (int64_t)a.m * c] * a.invl[c] : 0.0;
                }
                if (et < BN2) {
                    bt[et] = (a.beta && j0 + et < a.m) ? a.beta[j0 + et] : 0.0;
                    cs[et] = a.sigma2 * a.colscale[j0 + et];
                    if (ROWD) vv[et] = (a.vvec && j0 + et < a.m) ? a.vvec[j0 + et] : 0.0;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&bars.tmem_full, vt & 1);
            tc_fence_after();
            double T[64];
            {
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
                const double wgt = sw == 0 ? W_LEVELS_HI : W_LEVELS_LO;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    long long acc[16];
                    if (sw == 0) drain16_n128<4>(tcol + g * 16, acc);
                    else drain16_n128<NS - 4>(tcol + g * 16, acc);
#pragma unroll
                    for (int c = 0; c < 16; ++c) T[g * 16 + c] = wgt * (double)acc[c];
                }
            }
            tc_fence_before();                                  // TMEM is free for the next virtual tile's MMAs
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) mbar_arrive_remote(&bars.tmem_empty, 0);
                else mbar_arrive(&bars.tmem_empty);
            }
            if (a.debug & 1) {
                s0 += T[0] + T[17] + T[34] + T[51];
            } else if (iv) {
                const int8_t *kimg = a.kslices + ((size_t)rb * KBm) * IMG_BLOCK + (size_t)row * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int jg = j0 + half * 64 + g * 16;     // first column of this 16-column group
                    uint4 w[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        w[s] = __ldg(reinterpret_cast<const uint4 *>(kimg + s * a.kstride + (size_t)(jg / BK) * IMG_BLOCK +
                                                                     (size_t)((jg % BK) / 16) * 2048));
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int jj = half * 64 + g * 16 + e;
                        long long q4[4];
                        if ((e & 3) == 0) join_quad(w, e >> 2, q4);
                        if (j0 + jj < a.m) {
                            const long long qd = q4[e & 3];
                            const double kij = a.sigma2 * FIX_INV * (double)qd;
                            if (ROWD) {
                                const double tij = cs[jj] * T[g * 16 + e];
                                const double tk = tij * kij;
                                rt[0] += tk;
                                if (sw == 0) {
                                    rbt[0] = fma(bt[jj], kij, rbt[0]);
                                    rkv = fma(kij, vv[jj], rkv);
                                }
                                if (!a.nodims) {
                                    const double bk = sw == 0 ? bt[jj] * kij : 0.0;
#pragma unroll
                                    for (int k = 0; k < DT; k++) {
                                        const double tt = xi[k] - us[jj * DT + k], d2 = tt * tt;
                                        rt[(ROWD ? 1 + k : 0)] = fma(tk, d2, rt[(ROWD ? 1 + k : 0)]);
                                        rbt[(ROWD ? 1 + k : 0)] = fma(bk, d2, rbt[(ROWD ? 1 + k : 0)]);
                                    }
                                }
                                if (qd == FIX_ONE && !a.nodims)
                                    record_if_coincident_i8_part(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                                 a.coin_omega, a.coin_cap, tij, sw);
                            } else {
                                const double om = fma(rsi, cs[jj] * T[g * 16 + e], sw == 0 ? rai * bt[jj] : 0.0);
                                const double pk = om * kij;
                                s0 += pk;
#pragma unroll
                                for (int k = 0; k < DT; k++) {
                                    const double tt = xi[k] - us[jj * DT + k];
                                    sc[k] = fma(pk, tt * tt, sc[k]);
                                }
                                if (qd == FIX_ONE)
                                    record_if_coincident_i8_part(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                                 a.coin_omega, a.coin_cap, om, sw);
                            }
                        }
                    }
                }
            }
        }
        if (ROWD) {
            double *part = a.rowd_part + ((int64_t)(blockIdx.y * 2 + half) * a.nslots) * a.ld + i;
            if (a.nodims) {
                part[0] = rt[0];
            } else {
#pragma unroll
                for (int c = 0; c < DT + 1; c++) {
                    part[(int64_t)c * a.ld] = rt[ROWD ? c : 0];
                    if (a.beta) part[(int64_t)(DT + 1 + c) * a.ld] = rbt[ROWD ? c : 0];
                }
            }
            if (a.vvec) part[(int64_t)(a.nslots - 1) * a.ld] = rkv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
#pragma unroll
        for (int c = 0; c < DT; c++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sc[c] += __shfl_xor_sync(0xffffffffu, sc[c], o);
        if (lane == 0) {
            red[ew * PART_STRIDE_I8] = s0;
#pragma unroll
            for (int c = 0; c < DT; c++) red[ew * PART_STRIDE_I8 + 1 + c] = sc[c];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!ROWD && et < 1 + DT) {
            double v = 0.0;
            for (int k = 0; k < 8; k++) v += red[k * PART_STRIDE_I8 + et];
            double *slot = a.part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * PART_STRIDE_I8 + et;
            *slot = a.first ? v : (*slot + v);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

bool i8_wide_tiles()
{
    static const bool on = [] {
        const char *e = getenv("SRGP_I8_TILE");
        return !(e && atoi(e) == 64);
    }();
    return on;
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
                                const GenParams &p, int8_t *slices, size_t slice_stride, double *b1part, int first,
                                const double *rw, const double *wmax, int8_t *slices_w)
{
    if (rw)
        gen_slices_knotrows_kernel<DT, true><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                                   slices, slice_stride, b1part, first, rw, wmax, slices_w);
    else
        gen_slices_knotrows_kernel<DT, false><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                                    slices, slice_stride, b1part, first, nullptr, nullptr,
                                                                    nullptr);
}

// *out = max_i |w_i| over the shard (one block; n is at most a few million)
__global__ void __launch_bounds__(1024) absmax_kernel(const double *__restrict__ w, int64_t n, double *__restrict__ out)
{
    __shared__ double red[32];
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) mx = fmax(mx, fabs(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) *out = mx;
    }
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

// Pass 1 on the INT8 tensor cores: G = K^T diag(rowweight) K (mp x mp, both triangles; rowweight may be null = 1),
// b1 = K^T rvec.  With weights the chunk holds two slice sets (e and w e / 2^ew), so it covers half the rows.
int gauss_pass1_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *rowweight, const double *rvec, double *G,
                   double *b1)
{
    cudaStream_t s = ctx->stream;
    static DeviceOnce once, once2;
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(i8_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int mp = w->mp, m = w->m, d = w->d;
    // 128 x 128 tiles in two sweeps (default) or the 128 x 64 single-sweep tiles (SRGP_I8_TILE=64: the A/B switch the
    // measurements in profiles/ were taken with)
    const bool wide = i8_wide_tiles();
    if (wide && once2.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(i8_gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM2_SMEM));
    const int tiles = wide ? w->nt * (w->nt + 1) / 2 : w->nt * (w->nt + 1);
    const int nsplit = std::max(1, std::min(16, ctx->sm_count / tiles));
    const int quantum = BK * nsplit;
    const int sets = rowweight ? 2 : 1;
    // rows per launch: the chunk buffer holds sets x 8 slices x rows x mp bytes, and one INT32 accumulator may sum at
    // most MAX_ROWS_PER_SPLIT rows
    int64_t rows1 = std::min<int64_t>((int64_t)w->chunk_elems / mp / sets, (int64_t)MAX_ROWS_PER_SPLIT * nsplit);
    rows1 = std::max<int64_t>(quantum, rows1 / quantum * quantum);
    const size_t tile_elems = (size_t)BM * (wide ? BN2 : BN);
    SRGP_TRY(w->Gpart.reserve((size_t)tiles * nsplit * tile_elems * 8));
    double *wmax = nullptr;
    if (rowweight) {
        SRGP_TRY(w->i8scal.reserve(64));
        wmax = w->i8scal.d();
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        absmax_kernel<<<1, 1024, 0, s>>>(rowweight, ctx->n, wmax);
        SRGP_LAUNCH_CHECK();
    }
    int first = 1;
    if (ctx->n == 0) {
        SRGP_CUDA(cudaMemsetAsync(w->Gpart.p, 0, (size_t)tiles * nsplit * tile_elems * 8, s));
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
        int8_t *slices_w = rowweight ? slices + slice_stride * NS : slices;
        if (cidx >= 2) SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_used[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GEN, sg);
            dim3 grid(mp / 128, w->gen_groups);
            const size_t smem = sizeof(double) * BK * (d + 2);
#define CALL(D) launch_gen_knotrows<D>(sg, grid, smem, ctx->Xp, ctx->n, rvec, r0, rows_valid, rows_padded, w->U.d(), m, mp, d, gp, slices, slice_stride, w->b1part.d(), first, rowweight, wmax, slices_w)
            SRGP_D_SWITCH_I8(d, CALL)
#undef CALL
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_gen[b], sg));
        SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_gen[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GRAM, s);
            if (wide)
                i8_gram2_kernel<<<tiles * nsplit, THREADS, GRAM2_SMEM, s>>>(slices_w, slices, slice_stride, rows_padded / BK,
                                                                            nsplit, gp.sigma2 * gp.sigma2, wmax, w->Gpart.d(), first);
            else
                i8_gram_kernel<<<tiles * nsplit, THREADS, SMEM_BYTES, s>>>(slices_w, slices, slice_stride, rows_padded / BK, nsplit,
                                                                          gp.sigma2 * gp.sigma2, wmax, w->Gpart.d(), first);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_used[b], s));
        first = 0;
    }
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        if (wide) i8_gram2_finalize_kernel<<<tiles, 128, 0, s>>>(w->Gpart.d(), nsplit, mp, G);
        else i8_gram_finalize_kernel<<<tiles, 128, 0, s>>>(w->Gpart.d(), nsplit, mp, G);
        SRGP_LAUNCH_CHECK();
        gram_sum_rows(s, w->b1part.d(), w->gen_groups, mp, b1);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

template <int DT>
static void launch_gen_datarows(cudaStream_t s, dim3 grid, const double *X, int64_t ldx, int64_t r0, int rows_valid,
                                const double *U, int m, int mp, const GenParams &p, int8_t *slices, size_t slice_stride)
{
    gen_slices_datarows_kernel<DT><<<grid, 128, sizeof(double) * BK * DT, s>>>(X, ldx, r0, rows_valid, U, m, mp, p, slices,
                                                                              slice_stride);
}

template <int DT, bool ROWD>
static cudaError_t launch_km_i8(cudaStream_t s, dim3 grid, int device, const KmI8Args &a)
{
    const size_t smem = STAGES * STAGE_BYTES + sizeof(Bars) + 16 + sizeof(double) * (BN * DT + 3 * BN + 8 * PART_STRIDE_I8);
    static DeviceOnce once;
    if (once.need(device)) {
        cudaError_t e = cudaFuncSetAttribute(i8_km_kernel<DT, ROWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    i8_km_kernel<DT, ROWD><<<grid, KM_THREADS, smem, s>>>(a);
    return cudaSuccess;
}

template <int DT, bool ROWD, bool PAIR>
static cudaError_t launch_km2_i8(cudaStream_t s, dim3 grid, int device, const KmI8Args &a)
{
    const size_t smem = Km2Cfg<PAIR>::RING_BYTES + sizeof(Bars2) + sizeof(double) * (BN2 * DT + 3 * BN2 + 8 * PART_STRIDE_I8);
    static DeviceOnce once;
    if (once.need(device)) {
        cudaError_t e = cudaFuncSetAttribute(i8_km2_kernel<DT, ROWD, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(KM2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1;      // adjacent row blocks
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, i8_km2_kernel<DT, ROWD, PAIR>, a);
}

// mp <= 16384: one INT32 level accumulator sums up to NS pairs x 2^14 x mp over the knots (tc_i8.cuh)
bool i8_pass2_supported(const GaussWS *w)
{
    return i8_enabled() && !w->want_knots && w->d >= 1 && w->d <= 8 && w->mp <= 16384;
}

static int km_pass_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs, const double *ra,
                      const double *beta, double *out, bool accumulate_slots, bool rowd, const double *vvec, double *rowd_out,
                      int64_t rowd_stride, bool nodims = false);

// Pass 2 (gradient sums) on the INT8 tensor cores; same contract and the same per-CTA slots as gauss_pass2.
int gauss_pass2_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs, const double *ra,
                   const double *beta, double *out, bool accumulate_slots)
{
    return km_pass_i8(ctx, w, gp, Mop, rs, ra, beta, out, accumulate_slots, false, nullptr, nullptr, 0);
}

// Per-row, per-dimension sums (gauss_rowd of gauss.cu) on the INT8 tensor cores: out[slot * stride + i].
int gauss_rowd_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *beta, const double *vvec,
                  double *out, int64_t stride)
{
    return km_pass_i8(ctx, w, gp, Mop, nullptr, nullptr, beta, nullptr, false, true, vvec, out, stride);
}

// Row quadratic forms (gauss_rowform of gauss.cu): rowq_i = K_i Mop K_i^T, rowkv_i = K_i v (v may be null).
int gauss_rowform_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *vvec, double *rowq,
                     double *rowkv)
{
    // slot 0 -> rowq, slot 1 -> rowkv: the two outputs are addressed as out[slot * stride + i] with stride = rowkv - rowq
    return km_pass_i8(ctx, w, gp, Mop, nullptr, nullptr, nullptr, nullptr, false, true, vvec, rowq, vvec ? rowkv - rowq : 0, true);
}

static int km_pass_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs, const double *ra,
                      const double *beta, double *out, bool accumulate_slots, bool rowd, const double *vvec, double *rowd_out,
                      int64_t rowd_stride, bool nodims)
{
    cudaStream_t s = ctx->stream;
    const int mp = w->mp, m = w->m, d = w->d;
    const int KBm = mp / BK;
    const int slots = w->rblocks * w->cgroups;
    const size_t mstride = (size_t)mp * mp;
    SRGP_TRY(w->i8buf.reserve(mstride * NS + (size_t)mp * 16));
    int8_t *mslices = reinterpret_cast<int8_t *>(w->i8buf.p);
    double *colscale = reinterpret_cast<double *>(mslices + mstride * NS);
    unsigned long long *colbits = reinterpret_cast<unsigned long long *>(colscale + mp);
    {
        SRGP_CUDA(cudaMemsetAsync(colbits, 0, (size_t)mp * 8, s));
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        mop_rowmax_kernel<<<dim3(mp / 128, 16), 128, 0, s>>>(Mop, mp, colbits);
        SRGP_LAUNCH_CHECK();
        slice_mop_kernel<<<dim3(mp / 128, KBm), 128, 0, s>>>(Mop, mp, mslices, mstride, colbits, colscale);
        SRGP_LAUNCH_CHECK();
    }
    int first = accumulate_slots ? 0 : 1;
    if (!rowd && ctx->n == 0 && first) SRGP_CUDA(cudaMemsetAsync(w->part2.p, 0, (size_t)slots * PART_STRIDE_I8 * 8, s));
    const int nslots = nodims ? (vvec ? 2 : 1) : (d + 1) * (beta ? 2 : 1) + (vvec ? 1 : 0), groups = w->cgroups * 2;
    if (rowd) SRGP_TRY(w->rowdpart.reserve((size_t)groups * nslots * w->rows2 * 8));
    cudaStream_t sg = getenv("SRGP_NO_OVERLAP") ? s : ctx->stream3;
    SRGP_CUDA(cudaEventRecord(ctx->ev_fork, s));
    SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_fork, 0));
    int cidx = 0;
    for (int64_t r0 = 0; r0 < ctx->n; r0 += w->rows2, cidx++) {
        const int rows_valid = (int)std::min<int64_t>(w->rows2, ctx->n - r0);
        const int b = cidx & 1;
        int8_t *kslices = reinterpret_cast<int8_t *>(w->chunk.d() + (size_t)b * w->chunk_elems);
        const size_t kstride = (size_t)w->rows2 * mp;
        if (cidx >= 2) SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_used[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GEN, sg);
            dim3 grid(w->rblocks, std::min(KBm, 16));
#define CALL(D) launch_gen_datarows<D>(sg, grid, ctx->Xp, ctx->n, r0, rows_valid, w->U.d(), m, mp, gp, kslices, kstride)
            switch (d) {
            case 1: CALL(1); break;
            case 2: CALL(2); break;
            case 3: CALL(3); break;
            case 4: CALL(4); break;
            case 5: CALL(5); break;
            case 6: CALL(6); break;
            case 7: CALL(7); break;
            default: CALL(8); break;
            }
#undef CALL
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_gen[b], sg));
        SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_gen[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_KM, s);
            KmI8Args a;
            a.kslices = kslices;
            a.kstride = kstride;
            a.mslices = mslices;
            a.mstride = mstride;
            a.colscale = colscale;
            a.KBm = KBm;
            a.mp = mp;
            a.m = m;
            a.X = ctx->Xp;
            a.ldx = ctx->n;
            a.r0 = r0;
            a.rows_valid = rows_valid;
            a.U = w->U.d();
            a.rs = rs;
            a.ra = ra;
            a.beta = beta;
            for (int c = 0; c < 8; c++) a.invl[c] = gp.invl[c];
            a.sigma2 = gp.sigma2;
            // The kernel can take several row blocks per CTA one after the other (all CTAs then sweep a fraction of the
            // chunk's row blocks at a time, which shortens the re-read distance of the K slices in L2).  Measured: no
            // change in kernel time (the pass is bound by the shared-memory port and the FP64 pipe, not by HBM), so one
            // row block per CTA it stays.
            const int nsub = 1;
            a.nsub = nsub;
            const bool wide = i8_wide_tiles();
            a.tiles_per_cta = wide ? (mp / BN2) / w->cgroups : (mp / BN) / (w->cgroups * nsub);
            a.cluster = (wide && (w->rblocks / nsub) % 2 == 0 && getenv("SRGP_PAIR")) ? 2 : 1;
            a.debug = getenv("SRGP_KM_DEBUG") ? atoi(getenv("SRGP_KM_DEBUG")) : 0;
            a.part = w->part2.d();
            a.first = first;
            a.coin_count = w->coin_count();
            a.coin_list = w->coin_list();
            a.coin_omega = w->coin_omega();
            a.coin_cap = w->coin_cap;
            a.vvec = vvec;
            a.rowd_part = rowd ? w->rowdpart.d() : nullptr;
            a.nslots = nslots;
            a.ld = w->rows2;
            a.nodims = nodims ? 1 : 0;
            dim3 grid(w->rblocks / nsub, w->cgroups * nsub);
            cudaError_t e = cudaSuccess;
#define CALL2(D, R) (a.cluster == 2 ? launch_km2_i8<D, R, true>(s, grid, ctx->device, a) : launch_km2_i8<D, R, false>(s, grid, ctx->device, a))
#define CALL(D)                                                                                                          \
    e = wide ? (rowd ? CALL2(D, true) : CALL2(D, false))                                                                 \
             : (rowd ? launch_km_i8<D, true>(s, grid, ctx->device, a) : launch_km_i8<D, false>(s, grid, ctx->device, a))
            switch (d) {
            case 1: CALL(1); break;
            case 2: CALL(2); break;
            case 3: CALL(3); break;
            case 4: CALL(4); break;
            case 5: CALL(5); break;
            case 6: CALL(6); break;
            case 7: CALL(7); break;
            default: CALL(8); break;
            }
#undef CALL
#undef CALL2
            SRGP_CUDA(e);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_used[b], s));
        if (rowd) {
            KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
            gram_combine_rowd(s, w->rowdpart.d(), groups, nslots, w->rows2, rows_valid, rowd_out + r0, rowd_stride);
            SRGP_LAUNCH_CHECK();
        }
        first = 0;
    }
    if (out) {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        gram_sum_part(s, w->part2.d(), slots, PART_STRIDE_I8, d + 1, out);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

}  // namespace srgp
