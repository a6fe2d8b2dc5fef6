// gauss_i8.cu -- the row passes of the fused Gaussian / Laplace pipelines on the INT8 tensor cores (tcgen05 + TMEM).
// Reference functions served: the n x m products inside elbo_fun / delbo_dcov_par (R/vi_functions.R:64-121, 126-420),
// obj_fun_norm / dlogp_dcov_par (R/laplace_approx_obj_funs.R:6-52, R/laplace_approx_gradient.R:720-968) and
// dlogq_dcov_par (R/laplace_approx_gradient.R:25-339); the algebra is in gauss_vi.cu / gauss_fic.cu / laplace.cu,
// the scheme and its error bound in tc_i8.cuh and DESIGN.md section 3a.
//
// Pass 1: G = K^T diag(w) K (w optional) and b1 = K^T r.  The generator emits k_ij / sigma^2 = exp(-d_ij^2 / 2) in
// (0, 1] as NS INT8 digit slices, already in the shared-memory operand image of tc_i8.cuh (plus a second slice set
// w_i k_ij / (sigma^2 2^ew) for a weighted Gram); the Gram kernel runs one CTA per (128 x 128 tile of the lower block
// triangle, row split), two sweeps over its row range (levels 0..3, then 4..NS-1, 4 TMEM accumulators each), converts
// INT32 -> FP64 once per sweep and adds into its own slot (deterministic).  Everything is exact except the final
// FP64 summation of levels, splits and chunks, so the result is at least as accurate as the DMMA SYRK it replaces
// (profiles/r01_ozaki_*.json).
// Pass 2 and the row forms: T = K Mop^T with Mop sliced per output column; one epilogue thread per data row forms the
// gradient sums (gauss_pass2), the per-row / per-dimension sums (gauss_rowd) or the row quadratic forms (gauss_rowform).
#include <math.h>
#include <stdlib.h>
#include <string.h>

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

// WMODE 1: a second slice set holds w_i e_ij / 2^ew (2^ew >= max_i |w_i|, *wmax on the device) for the weighted Grams
// K^T diag(w) K = (diag(w) K)^T K of the FIC model; w may have either sign.
// WMODE 2 (w >= 0, e.g. FIC's B = 1 / Z): ONE slice set sqrt(w_i) e_ij / s with s = the power of two >= sqrt(max w):
// K^T diag(w) K = s^2 (that set)^T (that set) is then an unweighted Gram -- half the generator output, twice the rows per
// chunk, and the diagonal tiles load their operand once.
// WMODE 3 = WMODE 2 with K read from the shard's materialised row-major matrix Kr (ld = mp; the Laplace Newton loop keeps it
// for its matrix-vector products and K does not change between its iterations): no distance, no exp -- one coalesced 8-byte
// load per entry, and 1 / sigma^2 folded into the row scale.
template <int DT, int WMODE>
__global__ void __launch_bounds__(128)
gen_slices_knotrows_kernel(const double *__restrict__ X, int64_t ldx, const double *__restrict__ r, int64_t r0,
                           int rows_valid, int rows_padded, const double *__restrict__ U, int m, int mp, int d_rt,
                           GenParams p, int8_t *__restrict__ slices, double *__restrict__ b1part,
                           int first, const double *__restrict__ rw, const double *__restrict__ wmax,
                           int8_t *__restrict__ slices_w, const double *__restrict__ Kr)
{
    extern __shared__ double sx[];   // [64][d] scaled rows, then [64] residuals, then [64] scaled row weights
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, 128);
    const int d = DT > 0 ? DT : d_rt;
    double *sr = sx + 64 * d;
    double *sw = sr + 64;
    constexpr bool WEIGHTED = WMODE == 1, FROMK = WMODE == 3;
    const double winv = WMODE == 1 ? 1.0 / pow2_ceil(*wmax)
                      : WMODE == 2 ? 1.0 / pow2_ceil(sqrt(*wmax))
                      : WMODE == 3 ? 1.0 / (pow2_ceil(sqrt(*wmax)) * p.sigma2) : 0.0;
    const int j = blockIdx.x * 128 + threadIdx.x;
    const bool jvalid = j < m;
    double uj[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
        for (int c = 0; c < DT; c++) uj[c] = jvalid ? U[j + (int64_t)m * c] * p.invl[c] : 0.0;
    }
    const int KB = rows_padded / BK, KST = rows_padded / KS;
    const int kb_per_group = (KB + gridDim.y - 1) / gridDim.y;
    const int kb_begin = blockIdx.y * kb_per_group, kb_end = min(KB, kb_begin + kb_per_group);
    double bacc = 0.0;
    for (int kb = kb_begin; kb < kb_end; kb++) {
        const int it0 = kb * BK;
        __syncthreads();
        if (!FROMK)
            for (int t = threadIdx.x; t < BK * d; t += 128) {
                const int ii = t / d, c = t - ii * d;
                const int i = it0 + ii;
                sx[t] = (i < rows_valid) ? X[r0 + i + ldx * c] * p.invl[c] : 0.0;
            }
        if (threadIdx.x < BK) {
            const int i = it0 + threadIdx.x;
            sr[threadIdx.x] = (i < rows_valid) ? r[r0 + i] : 0.0;
            if (WMODE == 1) sw[threadIdx.x] = (i < rows_valid) ? rw[r0 + i] * winv : 0.0;
            if (WMODE == 2 || WMODE == 3) sw[threadIdx.x] = (i < rows_valid) ? sqrt(fmax(rw[r0 + i], 0.0)) * winv : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16++) {
            uint32_t w[NS][4], ww[WEIGHTED ? NS : 1][4];
#pragma unroll
            for (int e0 = 0; e0 < 16; e0 += 4) {
                // 4 rows in flight per thread: independent distance / exp chains hide the FP64 latency
                double sq[4] = {0.0, 0.0, 0.0, 0.0};
                const int ii = c16 * 16 + e0;
                if (FROMK) {
                    double kv[4];
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        kv[q] = (jvalid && it0 + ii + q < rows_valid) ? __ldcs(Kr + (size_t)(r0 + it0 + ii + q) * mp + j) : 0.0;
#pragma unroll
                    for (int q = 0; q < 4; q++) bacc = fma(kv[q], sr[ii + q], bacc);
                    split_quad(kv[0] * sw[ii], kv[1] * sw[ii + 1], kv[2] * sw[ii + 2], kv[3] * sw[ii + 3], e0 >> 2, w);
                    continue;
                }
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
                if (WMODE == 2) split_quad(ev[0] * sw[ii], ev[1] * sw[ii + 1], ev[2] * sw[ii + 2], ev[3] * sw[ii + 3], e0 >> 2, w);
                else split_quad(ev[0], ev[1], ev[2], ev[3], e0 >> 2, w);
                if (WEIGHTED)
                    split_quad(ev[0] * sw[ii], ev[1] * sw[ii + 1], ev[2] * sw[ii + 2], ev[3] * sw[ii + 3], e0 >> 2,
                               reinterpret_cast<uint32_t (&)[NS][4]>(ww));
            }
            const size_t img = img_off(blockIdx.x, KST, kb * 2 + (c16 >> 1)) + (size_t)(c16 & 1) * 2048 + (size_t)threadIdx.x * 16;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                *reinterpret_cast<uint4 *>(slices + img + s * A_TILE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
                if (WEIGHTED)
                    *reinterpret_cast<uint4 *>(slices_w + img + s * A_TILE) = make_uint4(ww[s][0], ww[s][1], ww[s][2], ww[s][3]);
            }
        }
    }
    double *slot = b1part + (int64_t)blockIdx.y * mp + j;
    *slot = first ? bacc : (*slot + bacc);
}

// ------------------------------------------------------------------------------------------------
// Operand ring shared by the Gram and the K*M kernel: units of 4 A-slice tiles + 4 B-slice tiles of one 32-byte k-step.
// Sweep 0 needs slices 0..3 = one unit per k-step; sweep 1 needs all NS slices = two units (slices 0..3, then 4..NS-1).
// A unit is filled by ONE bulk copy per operand (the slices of a (block, k-step) are contiguous in the image).
//   single CTA: B tiles are 128 columns x 32 B = 4 KB -> 32 KB units;
//   CTA pair (cta_group::2): each CTA holds the 64-column half of the B tiles (2 KB) -> 24 KB units.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_UNITS = 8;
template <bool PAIR> struct RingCfg {
    static constexpr int B_TILE_BYTES = PAIR ? A_TILE / 2 : A_TILE;
    static constexpr int UNIT_BYTES = 4 * A_TILE + 4 * B_TILE_BYTES;
    static constexpr uint32_t B_LBO = PAIR ? 1024 : 2048;
};
struct Bars2 {
    uint64_t full[MAX_UNITS], empty[MAX_UNITS], tmem_full, tmem_empty;
    uint64_t go;             // issuer 0 -> issuer 1: the first k-step of a virtual tile (the overwriting MMAs) has been issued
    uint32_t tmem_slot, pad;
};

// part 0 / 1 of k-step ks of an A block (and of a B block unless `same`): slices [4 part, 4 part + ns)
template <bool PAIR>
__device__ __forceinline__ void load_unit(uint32_t ub, uint64_t *full, const int8_t *a_img, const int8_t *b_img, bool same,
                                          int part, uint32_t crank)
{
    constexpr int BT = RingCfg<PAIR>::B_TILE_BYTES;
    const int ns = part == 0 ? 4 : NS - 4;
    mbar_expect_tx(full, (uint32_t)(ns * (A_TILE + (same ? 0 : BT))));
    bulk_g2s(ub, a_img + part * 4 * A_TILE, (uint32_t)(ns * A_TILE), full);
    // pair image of the B operand: [64-column half][slice][2 KB] per (block, k-step), see slice_mop_kernel
    if (!same) bulk_g2s(ub + 4 * A_TILE, b_img + (PAIR ? crank * (NS * BT) : 0u) + part * 4 * BT, (uint32_t)(ns * BT), full);
}

template <bool PAIR>
__device__ __forceinline__ void mma_n128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    if (PAIR) mma_i8_n128_pair(tmem_d, da, db, accumulate);
    else mma_i8_n128(tmem_d, da, db, accumulate);
}

// one k-step of sweep 0: levels 0..3 from the slices 0..3 of one unit (a0 / b0: shared addresses of its A and B tiles)
template <bool PAIR>
__device__ __forceinline__ void issue_sweep0(uint32_t a0, uint32_t b0, uint32_t tmem_base, uint32_t keep)
{
    constexpr int BT = RingCfg<PAIR>::B_TILE_BYTES;
    const uint64_t da0 = make_desc(a0, 2048, 128), db0 = make_desc(b0, RingCfg<PAIR>::B_LBO, 128);
#pragma unroll
    for (int sb = 0; sb < 4; ++sb)
#pragma unroll
        for (int sa = 0; sa < 4; ++sa)
            if (sa + sb < 4)        // accumulator L is first touched by the pair (sa = L, sb = 0)
                mma_n128<PAIR>(tmem_base + (uint32_t)(sa + sb) * BN2, da0 + (uint64_t)((sa * A_TILE) >> 4),
                               db0 + (uint64_t)((sb * BT) >> 4), sb == 0 ? keep : 1u);
}

// one k-step of sweep 1: levels 4..NS-1 -> accumulators 0..NS-5, slices 0..3 from unit "lo", 4..NS-1 from unit "hi"
template <bool PAIR>
__device__ __forceinline__ void issue_sweep1(uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t tmem_base,
                                             uint32_t keep)
{
    constexpr int BT = RingCfg<PAIR>::B_TILE_BYTES;
    constexpr uint32_t B_LBO = RingCfg<PAIR>::B_LBO;
    const uint64_t dlo_a = make_desc(a_lo, 2048, 128), dlo_b = make_desc(b_lo, B_LBO, 128);
    const uint64_t dhi_a = make_desc(a_hi, 2048, 128), dhi_b = make_desc(b_hi, B_LBO, 128);
#pragma unroll
    for (int sb = 0; sb < NS; ++sb)
#pragma unroll
        for (int sa = 0; sa < NS; ++sa) {
            const int L = sa + sb;
            if (L >= 4 && L < NS) {
                const uint64_t da = (sa < 4 ? dlo_a : dhi_a) + (uint64_t)(((sa & 3) * A_TILE) >> 4);
                const uint64_t db = (sb < 4 ? dlo_b : dhi_b) + (uint64_t)(((sb & 3) * BT) >> 4);
                mma_n128<PAIR>(tmem_base + (uint32_t)(L - 4) * BN2, da, db, sb == 0 ? keep : 1u);
            }
        }
}

// The MMA stream of one virtual tile (sweep sw over ksteps k-steps), as seen by issuer X (0 or 1).  The tensor pipe takes
// the next MMA only when the previous one has started, so every cycle the issuing thread spends elsewhere is a bubble: a
// SATISFIED mbarrier wait per k-step costs ~130 clk = 20 % of a 10-MMA k-step (tools/probes/acc_rotation.cu).  Two threads
// in different warps therefore take the k-steps in turns -- one waits for its next unit while the MMAs of the other run
// (64.0 clk per MMA in the probe).  The sums are exact integers, so the order in which the two streams interleave does
// not matter, except that the overwriting MMAs of k-step 0 must come first: issuer 1 starts a virtual tile only after
// issuer 0 has issued k-step 0 (tcgen05 fences + `go`).  Each issuer commits the units it consumed and, after its last
// k-step, the tile (tmem_full counts 2).
// Each issuer has its OWN ring of NR units (k-step ks lives in ring ks & 1): an mbarrier distinguishes only two
// consecutive phases, so a barrier whose successive phases were waited on by different threads could be passed one lap
// early by a thread that ran ahead.  With one consumer (and one relay, in a CTA pair) per barrier every wait is for the
// phase right after the one that thread saw last.  The unit index within ring X counts the units of X over the whole
// kernel: xbase = units of ring X before this virtual tile.
__host__ __device__ constexpr int ring_units_before(int vt, int ksteps) { return (vt >> 1) * 3 * (ksteps / 2) + (vt & 1) * (ksteps / 2); }

template <bool PAIR>
__device__ __forceinline__ void issue_vt(uint8_t *smem, Bars2 &bars, uint32_t tmem_base, int X, int vt, int sw, int xbase,
                                         int ksteps, int NR, bool same)
{
    constexpr int UB = RingCfg<PAIR>::UNIT_BYTES;
    if (X == 0) {
        if (vt > 0) {                                       // the epilogue must have drained the previous virtual tile
            mbar_wait(&bars.tmem_empty, (vt - 1) & 1);
            tc_fence_after();
        }
    } else {
        mbar_wait(&bars.go, vt & 1);
        tc_fence_after();
    }
    for (int ks = X; ks < ksteps; ks += 2) {
        const int g = xbase + (sw ? 2 : 1) * (ks >> 1);
        const int u0 = X * NR + g % NR;
        mbar_wait(&bars.full[u0], (g / NR) & 1);
        const uint32_t a0 = smem_u32(smem + u0 * UB), b0 = same ? a0 : a0 + 4 * A_TILE;
        const uint32_t keep = ks == 0 ? 0u : 1u;
        if (sw == 0) {
            tc_fence_after();
            issue_sweep0<PAIR>(a0, b0, tmem_base, keep);
            if (PAIR) mma_commit_pair(&bars.empty[u0]);
            else mma_commit(&bars.empty[u0]);
        } else {
            const int u1 = X * NR + (g + 1) % NR;
            mbar_wait(&bars.full[u1], ((g + 1) / NR) & 1);
            tc_fence_after();
            const uint32_t a1 = smem_u32(smem + u1 * UB), b1 = same ? a1 : a1 + 4 * A_TILE;
            issue_sweep1<PAIR>(a0, b0, a1, b1, tmem_base, keep);
            if (PAIR) {
                mma_commit_pair(&bars.empty[u0]);
                mma_commit_pair(&bars.empty[u1]);
            } else {
                mma_commit(&bars.empty[u0]);
                mma_commit(&bars.empty[u1]);
            }
        }
        if (ks == 0) {                                      // X == 0
            tc_fence_before();
            mbar_arrive(&bars.go);
        }
    }
    if (PAIR) mma_commit_pair(&bars.tmem_full);
    else mma_commit(&bars.tmem_full);
}

// Producer side of the two rings: k-step ks of a virtual tile goes to ring ks & 1 (one unit in sweep 0, two in sweep 1).
// gx[X] counts the units of ring X since the start of the kernel.
template <bool PAIR>
__device__ __forceinline__ void produce_kstep(uint8_t *smem, Bars2 &bars, int (&gx)[2], int NR, int ks, int sw, const int8_t *a_img,
                                              const int8_t *b_img, bool same, uint32_t crank)
{
    constexpr int UB = RingCfg<PAIR>::UNIT_BYTES;
    const int X = ks & 1;
    for (int part = 0; part <= sw; ++part) {
        const int g = gx[X]++;
        const int u = X * NR + g % NR;
        if (g >= NR) mbar_wait(&bars.empty[u], ((g / NR) - 1) & 1);
        load_unit<PAIR>(smem_u32(smem + u * UB), &bars.full[u], a_img, b_img, same, part, crank);
    }
}

// ------------------------------------------------------------------------------------------------
// Gram kernel on 128 x 128 tiles, two sweeps (tc_i8.cuh)// ------------------------------------------------------------------------------------------------
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

constexpr int GRAM2_NR = 3;                                    // units per issuer ring: 2 x 3 x 32 KB
constexpr int GRAM2_SMEM = 2 * GRAM2_NR * RingCfg<false>::UNIT_BYTES + (int)sizeof(Bars2);

__global__ void __launch_bounds__(THREADS, 1)
i8_gram2_kernel(const int8_t *__restrict__ slices_a, const int8_t *__restrict__ slices, int KST, int nsplit, double scale,
                const double *__restrict__ wmax, int wsqrt, double *__restrict__ Gpart, int first)
{   // slices_a: A operand (rows 128 I ..): the weighted slice set, or `slices` itself; slices: B operand (columns 128 J ..)
    constexpr int NR = GRAM2_NR, NU = 2 * NR, UB = RingCfg<false>::UNIT_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars2 &bars = *reinterpret_cast<Bars2 *>(smem + NU * UB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int I, J;
    tile2_to_ij(blockIdx.x / nsplit, I, J);
    const int split = blockIdx.x % nsplit;
    const int ksteps = KST / nsplit, ks0 = split * ksteps;
    const bool same = (I == J) && (slices_a == slices);

    if (threadIdx.x == 0) {
        for (int u = 0; u < NU; ++u) {
            mbar_init(&bars.full[u], 1);
            mbar_init(&bars.empty[u], 1);
        }
        mbar_init(&bars.tmem_full, 2);                  // both issuers commit
        mbar_init(&bars.tmem_empty, 4);
        mbar_init(&bars.go, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_all(&bars.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars.tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int gx[2] = {0, 0};
            for (int sw = 0; sw < 2; ++sw)
                for (int it = 0; it < ksteps; ++it)
                    produce_kstep<false>(smem, bars, gx, NR, it, sw, slices_a + img_off(I, KST, ks0 + it),
                                         slices + img_off(J, KST, ks0 + it), same, 0u);
        }
    } else if (warp == 1 || warp == 2) {
        if (lane == 0)
            for (int sw = 0; sw < 2; ++sw)
                issue_vt<false>(smem, bars, tmem_base, warp - 1, sw, sw, ring_units_before(sw, ksteps), ksteps, NR, same);
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        if (wmax) {                                             // the power-of-two scale(s) of the weighted operand(s)
            const double sc1 = pow2_ceil(wsqrt ? sqrt(*wmax) : *wmax);
            scale *= wsqrt ? sc1 * sc1 : sc1;
        }
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
            if (lane == 0) mbar_arrive(&bars.tmem_empty);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tmem_base);
    }
}

// Sum the split slots of the 128 x 128 tiles and scatter to the full symmetric matrix (column-major, ld = mp).
// One CTA per 32 x 32 piece of a tile (grid: tiles x 16): the slots are row-major per tile, so a warp reads 32 consecutive
// columns of one row (256 B) and the piece goes through shared memory for the column-major stores.  Splits are summed in
// slot order (deterministic).
__global__ void __launch_bounds__(256)
i8_gram2_finalize_kernel(const double *__restrict__ Gpart, int nsplit, int mp, double *__restrict__ G)
{
    __shared__ double sm[32][33];
    int I, J;
    tile2_to_ij(blockIdx.x, I, J);
    const int r0 = (blockIdx.y >> 2) * 32, c0 = (blockIdx.y & 3) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double *base = Gpart + (size_t)blockIdx.x * nsplit * BM * BN2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k;
        const double *p = base + (size_t)r * BN2 + c0 + tx;
        double v = 0.0;
        for (int s = 0; s < nsplit; ++s) v += p[(size_t)s * BM * BN2];
        sm[ty + 8 * k][tx] = v;                            // [row][column]
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = ty + 8 * k;                          // column of the piece; tx = row
        const double v = sm[tx][c];
        const int row = I * BM + r0 + tx, col = J * BN2 + c0 + c;
        G[row + (int64_t)col * mp] = v;
    }
    if (J < I) {                                           // strictly below the diagonal block: mirror (coalesced along columns)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = ty + 8 * k;
            const double v = sm[r][tx];
            const int row = I * BM + r0 + r, col = J * BN2 + c0 + tx;
            G[col + (int64_t)row * mp] = v;
        }
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
                           const double *__restrict__ U, int m, int mp, GenParams p, int8_t *__restrict__ slices)
{
    extern __shared__ double su[];   // [64][DT] scaled knots of one k-block
    __shared__ double etab[EXP_TAB_DOUBLES];
    exp_tab_load(etab, threadIdx.x, 128);
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool ivalid = i < rows_valid;
    double xi[DT];
#pragma unroll
    for (int c = 0; c < DT; c++) xi[c] = ivalid ? X[r0 + i + ldx * c] * p.invl[c] : 0.0;
    const int KBm = mp / BK, KST = mp / KS;
    for (int kb = blockIdx.y; kb < KBm; kb += gridDim.y) {
        __syncthreads();
        for (int t = threadIdx.x; t < BK * DT; t += 128) {
            const int jj = t / DT, c = t - jj * DT;
            const int j = kb * BK + jj;
            su[t] = (j < m) ? U[j + (int64_t)m * c] * p.invl[c] : 0.0;
        }
        __syncthreads();
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
            int8_t *dst = slices + img_off(blockIdx.x, KST, kb * 2 + (c16 >> 1)) + (size_t)(c16 & 1) * 2048 + (size_t)threadIdx.x * 16;
#pragma unroll
            for (int s = 0; s < NS; s++)
                *reinterpret_cast<uint4 *>(dst + s * A_TILE) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
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
// pair = 1: the image a CTA pair reads -- per (block, k-step) [64-column half][slice][16-byte k-chunk][8 x 8 rows][16 B],
// so the half of the B tiles a CTA of the pair holds is one contiguous piece (LBO = 1024).
// grid (mp / 128, mp / 64), 128 threads.
__global__ void __launch_bounds__(128)
slice_mop_kernel(const double *__restrict__ Mop, int mp, int8_t *__restrict__ slices, int pair,
                 const unsigned long long *__restrict__ colbits, double *__restrict__ colscale)
{
    const int n = blockIdx.x * 128 + threadIdx.x, kb = blockIdx.y;
    const double mx = __longlong_as_double((long long)colbits[n]);
    int ex = 0;
    if (mx > 0.0 && mx < INFINITY) ex = ilogb(mx) + 1;
    const double inv = scalbn(1.0, -ex);
    if (kb == 0) colscale[n] = scalbn(1.0, ex);
    const int KST = mp / KS;
    const int half = threadIdx.x >> 6, r = threadIdx.x & 63;
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
        int8_t *dst = slices + img_off(blockIdx.x, KST, kb * 2 + (c16 >> 1));
        const size_t sstride = pair ? A_TILE / 2 : A_TILE;
        dst += pair ? (size_t)half * (NS * (A_TILE / 2)) + (size_t)(c16 & 1) * 1024 + (size_t)r * 16
                    : (size_t)(c16 & 1) * 2048 + (size_t)threadIdx.x * 16;
#pragma unroll
        for (int s = 0; s < NS; s++)
            *reinterpret_cast<uint4 *>(dst + s * sstride) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

struct KmI8Args {
    const int8_t *kslices;   // K digit slices of the chunk: blocks = 128-row blocks, k = knots
    const int8_t *mslices;   // Mop digit slices: blocks = 128 output columns n, k = knots
    const double *colscale;  // mp
    int mp, m;
    const double *X;         // resident rows (column-major, ld = ldx), chunk starts at r0
    int64_t ldx, r0;
    int rows_valid;
    const double *U;         // m x d
    const double *rs, *ra, *beta;
    double invl[8];
    double sigma2;
    int tiles_per_cta;       // 128-column blocks per CTA
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
    int cluster;             // 2 = adjacent row blocks run as tcgen05 CTA pairs (cta_group::2), 1 = single CTAs
    double *knot_part;       // KNOT: [row block x 4 lane quadrants][d][mp] column sums of P_ij (x_ic - u_jc) / l_c
};

// ------------------------------------------------------------------------------------------------
// pass 2 on 128 x 128 tiles, two sweeps per column block (tc_i8.cuh): every quantity the epilogue forms is linear in
// T = K Mop^T, so the two level groups of a tile are two "virtual tiles" whose contributions add; the terms that do not
// involve T (ra_i beta_j, the beta / v row sums of the ROWD mode) ride on sweep 0 only.
// CTA = 384 threads = 3 warpgroups: warp 0 TMA producer, warp 1 MMA issuer (warps 2, 3 idle), warps 4..11 epilogue.  The
// epilogue holds the 64 drained columns of its row in registers (so TMEM is released before the FP64 work starts):
// setmaxnreg moves registers from the first warpgroup to the two epilogue warpgroups.
// Operand rings: RingCfg units (above), 2 x 3 of 32 KB for a single CTA, 2 x 4 of 24 KB for a CTA pair.
// ------------------------------------------------------------------------------------------------
constexpr int KM2_THREADS = 384;
constexpr int KM2_EPI_REGS = 232, KM2_AUX_REGS = 40;
template <bool PAIR> struct Km2Cfg : RingCfg<PAIR> {
    static constexpr int NR = PAIR ? 4 : 3;                     // units per issuer ring
    static constexpr int UNITS = 2 * NR;
    static constexpr int RING_BYTES = UNITS * RingCfg<PAIR>::UNIT_BYTES;
};

// Rare path of quirk Q4 (same contract as record_if_coincident in gauss.cu): decided by the reference's own test, all
// coordinates bit-identical (src/covariance_function_derivativesC.cpp:157-163).
// `partial`: a record that carries only the part of Omega_ij that is linear in T (the second sweep of a pair): the knot
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

// Knot-location gradient: per-knot COLUMN sums over the rows of P_ij (x_ic - u_jc) / l_c.  The
// epilogue owns one row per thread, so a column sum is a sum over the 32 lanes of a warp: 16 values per thread (2 columns x 8
// dimensions) go through four halving exchanges and one full exchange -- 16 shuffles per thread instead of 80 -- after which
// lane L holds the sum of value index L >> 1, i.e. column (L >> 4) of the pair, dimension (L >> 1) & 7, in both lanes of a pair.
template <typename T>
__device__ __forceinline__ T warp_colsum16(T (&v)[16], int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const T send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T send = b3 ? v[i] : v[i + 4], keep = b3 ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = b2 ? v[i] : v[i + 2], keep = b2 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const T send = b1 ? v[0] : v[1], keep = b1 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// the 8 column-pair sums of one 16-column round -> this warp's slot of knot_part ([slot][d][mp]); even lanes write
template <int DT, typename T>
__device__ __forceinline__ void knot_round_out(const T (&kn)[8], int lane, double *slot, int mp, int col0, bool overwrite)
{
    const int k = (lane >> 1) & 7;
    if ((lane & 1) == 0 && k < DT) {
        double *p = slot + (int64_t)k * mp + col0 + (lane >> 4);
#pragma unroll
        for (int b = 0; b < 8; ++b) p[2 * b] = overwrite ? (double)kn[b] : p[2 * b] + (double)kn[b];
    }
}

// A column block is two virtual tiles: sweep 0 (levels 0..3), then sweep 1 (levels 4..NS-1).
__host__ __device__ constexpr int km_sweep(int vt) { return vt & 1; }
// units of one issuer ring before virtual tile vt
__host__ __device__ constexpr int km_units_before(int vt, int ksteps) { return ring_units_before(vt, ksteps); }

// PAIR: the two CTAs of a cluster (adjacent row blocks, same column group) form one tcgen05 CTA pair.
template <int DT, bool ROWD, bool PAIR, bool KNOT = false>
__global__ void __launch_bounds__(KM2_THREADS, 1) i8_km2_kernel(KmI8Args a)
{
    static_assert(!KNOT || (!ROWD && !PAIR), "the knot-gradient epilogue exists for the gradient mode on single CTAs");
    using Cfg = Km2Cfg<PAIR>;
    constexpr int NR = Cfg::NR, NU = Cfg::UNITS;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars2 &bars = *reinterpret_cast<Bars2 *>(smem + Cfg::RING_BYTES);
    double *us = reinterpret_cast<double *>(smem + Cfg::RING_BYTES + sizeof(Bars2));   // [128][DT] scaled knots
    double *bt = us + BN2 * DT;                                 // [128] beta
    double *cs = bt + BN2;                                      // [128] sigma^2 * column scale
    double *vv = cs + BN2;                                      // [128] v (ROWD)
    double *red = vv + BN2;                                     // [8][PART_STRIDE_I8]
    float *usf = reinterpret_cast<float *>(red + 8 * PART_STRIDE_I8);   // [128][DT] the same knots in single precision (sweep 1)
    float *csf = usf + BN2 * DT;                                // [128] sigma^2 * column scale * weight of the low level group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb0 = blockIdx.y * a.tiles_per_cta;               // first 128-column block of this CTA
    const int ksteps = a.mp / KS;
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
        mbar_init(&bars.tmem_full, 2);                          // both issuers commit
        mbar_init(&bars.tmem_empty, PAIR ? 16 : 8);             // epilogue warps of every CTA that shares the MMAs
        mbar_init(&bars.go, 1);
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
            int gx[2] = {0, 0};
            for (int vt = 0; vt < nvt; ++vt)
                for (int ks = 0; ks < ksteps; ++ks)
                    produce_kstep<PAIR>(smem, bars, gx, NR, ks, km_sweep(vt), a.kslices + img_off(rb, ksteps, ks),
                                        a.mslices + img_off(jb0 + (vt >> 1), ksteps, ks), false, crank);
        } else if (PAIR && !leader) {
            // ===== peer of a pair: tell the leader when this CTA's half of a unit has landed (one relay per issuer ring, so
            // that each barrier has a single waiter) =====
            if (lane == 0 && warp <= 2) {
                const int X = warp - 1, total = km_units_before(nvt, ksteps);
                for (int g = 0; g < total; ++g) {
                    const int u = X * NR + g % NR;
                    mbar_wait(&bars.full[u], (g / NR) & 1);
                    mbar_arrive_remote_relaxed(&bars.full[u], 0);
                }
            }
        } else if ((warp == 1 || warp == 2) && lane == 0) {
            // ===== two MMA issuers (in the leader of a pair: for both CTAs), see issue_vt =====
            for (int vt = 0; vt < nvt; ++vt)
                issue_vt<PAIR>(smem, bars, tmem_base, warp - 1, vt, km_sweep(vt), km_units_before(vt, ksteps), ksteps, NR, false);
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
        float xif[DT];
#pragma unroll
        for (int c = 0; c < DT; c++) xif[c] = (float)xi[c];
        const float rsf = (float)rsi, kscale = (float)(a.sigma2 * KF_SCALE);
        for (int vt = 0; vt < nvt; ++vt) {
            const int sw = km_sweep(vt);
            const int j0 = (jb0 + (vt >> 1)) * BN2;
            if ((vt & 1) == 0) {
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with the previous block's us / bt / cs
                for (int e = et; e < BN2 * DT; e += 256) {
                    const int jj = e / DT, c = e - jj * DT;
                    us[e] = (j0 + jj < a.m) ? a.U[j0 + jj + (int64_t)a.m * c] * a.invl[c] : 0.0;
                    usf[e] = (float)us[e];
                }
                if (et < BN2) {
                    bt[et] = (a.beta && j0 + et < a.m) ? a.beta[j0 + et] : 0.0;
                    cs[et] = a.sigma2 * a.colscale[j0 + et];
                    csf[et] = (float)(cs[et] * W_LEVELS_LO);
                    if (ROWD) vv[et] = (a.vvec && j0 + et < a.m) ? a.vvec[j0 + et] : 0.0;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&bars.tmem_full, vt & 1);
            tc_fence_after();
            const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
            const int8_t *kimg = a.kslices + img_off(rb, ksteps, 0) + (size_t)row * 16;
            if (sw == 1) {
                // ===== sweep 1: levels 4..NS-1, i.e. the part of T below 2^-32 of the operand scales (typically 2^-25 of
                // them).  Single precision carries it to 2^-24 of ITS size per operation, ~2^-49.7 of the operand scales
                // over the ~8 roundings below -- the size of the slice pairs NS = 7 drops anyway (2^-49.3); measured: the
                // ill-conditioned config-4 gradient moves from 0.9e-8 to 1.1e-8 of the long-double value, the float64
                // transcription of the reference sits at 1.7e-8 (tests/test_stated_sizes_gpu.py).  FP32 instructions do not
                // share the datapath of FP64 and the tensor pipe, whereas every FP64 instruction of this epilogue is a
                // bubble in the MMA stream (profiles/r02_km2_bound.txt: 28.3 -> 22.8 ms).  K_ij comes from its top 4 slices.
                float Tl[64];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t v[NS - 4][16];
#pragma unroll
                    for (int k = 0; k < NS - 4; ++k) tmem_ld16_nowait(tcol + (uint32_t)(g * 16 + k * BN2), v[k]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float t = (float)(int)v[0][c];
#pragma unroll
                        for (int k = 1; k < NS - 4; ++k) t = fmaf(t, 256.0f, (float)(int)v[k][c]);
                        Tl[g * 16 + c] = t;
                    }
                }
                tc_fence_before();                              // TMEM is free for the next virtual tile's MMAs
                __syncwarp();
                if (lane == 0) {
                    if (PAIR && !leader) mbar_arrive_remote_relaxed(&bars.tmem_empty, 0);
                    else mbar_arrive(&bars.tmem_empty);
                }
                double *kslot = KNOT ? a.knot_part + ((int64_t)(rb * 4 + q) * DT) * a.mp : nullptr;
                if (KNOT || iv) {                               // KNOT: every lane takes part in the column sums (shuffles)
                    float s0f = 0.0f, scf[DT], rtf[ROWD ? DT + 1 : 1];
#pragma unroll
                    for (int c = 0; c < DT; c++) scf[c] = 0.0f;
#pragma unroll
                    for (int c = 0; c < (ROWD ? DT + 1 : 1); c++) rtf[c] = 0.0f;
                    // the 64 columns in 4 rounds of 16 through ONE copy of the code (fully unrolled, the two sweeps are
                    // 180 KB of instructions and the instruction cache hit rate drops to 69 %): Tl[] can only be indexed with
                    // compile-time constants, so each round works on Tl[0..15] and then moves the rest down
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        const int jg = j0 + half * 64 + g * 16;
                        uint4 w[4];
#pragma unroll
                        for (int s = 0; s < 4; ++s)
                            w[s] = __ldg(reinterpret_cast<const uint4 *>(kimg + (size_t)(jg / KS) * KSTEP_BYTES + s * A_TILE +
                                                                         (size_t)((jg % KS) / 16) * 2048));
                        float vb[KNOT ? 16 : 1], kn[KNOT ? 8 : 1];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int jj = half * 64 + g * 16 + e;
                            int qt[4];
                            if ((e & 3) == 0) join_quad_top(w, e >> 2, qt);
                            const bool ok = iv && j0 + jj < a.m;
                            if (KNOT || ok) {
                                const float kf = kscale * (float)qt[e & 3];
                                const float tf = csf[jj] * Tl[e];                // T_ij (low part), scaled
                                const float pf = ok ? (ROWD ? tf : rsf * tf) * kf : 0.0f;
                                if (ROWD) rtf[0] += pf;
                                else s0f += pf;
                                if (!ROWD || !a.nodims) {
#pragma unroll
                                    for (int k = 0; k < DT; k++) {
                                        const float tt = xif[k] - usf[jj * DT + k];
                                        if (ROWD) rtf[ROWD ? 1 + k : 0] = fmaf(pf, tt * tt, rtf[ROWD ? 1 + k : 0]);
                                        else scf[k] = fmaf(pf, tt * tt, scf[k]);
                                        if (KNOT) vb[KNOT ? (e & 1) * 8 + k : 0] = pf * tt;
                                    }
                                    if (ok && qt[e & 3] == KF_ONE)   // top digits of exp(0) = 1: candidate for the bit-identical test
                                        record_if_coincident_i8_part(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                                     a.coin_omega, a.coin_cap, (double)(ROWD ? tf : rsf * tf), 1);
                                }
                            }
                            if (KNOT) {
#pragma unroll
                                for (int k = DT; k < 8; k++) vb[KNOT ? (e & 1) * 8 + k : 0] = 0.0f;
                                if (e & 1) kn[KNOT ? e >> 1 : 0] = warp_colsum16(reinterpret_cast<float (&)[16]>(vb), lane);
                            }
                        }
                        if (KNOT) knot_round_out<DT>(reinterpret_cast<const float (&)[8]>(kn), lane, kslot, a.mp, jg, false);
#pragma unroll
                        for (int c = 0; c < 48; ++c) Tl[c] = Tl[c + 16];
                    }
                    s0 += (double)s0f;
#pragma unroll
                    for (int c = 0; c < DT; c++) sc[c] += (double)scf[c];
                    if (ROWD) {
#pragma unroll
                        for (int c = 0; c < DT + 1; c++) rt[ROWD ? c : 0] += (double)rtf[ROWD ? c : 0];
                    }
                }
                continue;
            }
            // ===== sweep 0: levels 0..3 in double precision, with the terms that do not involve T =====
            double T[64];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                long long acc[16];
                drain16_n128<4>(tcol + g * 16, acc);
#pragma unroll
                for (int c = 0; c < 16; ++c) T[g * 16 + c] = W_LEVELS_HI * (double)acc[c];
            }
            tc_fence_before();                                  // TMEM is free for the next virtual tile's MMAs
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) mbar_arrive_remote_relaxed(&bars.tmem_empty, 0);
                else mbar_arrive(&bars.tmem_empty);
            }
            double *kslot = KNOT ? a.knot_part + ((int64_t)(rb * 4 + q) * DT) * a.mp : nullptr;
            if (KNOT || iv) {
#pragma unroll 1                                                // 4 rounds of 16 columns through one copy of the code, as above
                for (int g = 0; g < 4; ++g) {
                    const int jg = j0 + half * 64 + g * 16;     // first column of this 16-column group
                    uint4 w[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        w[s] = __ldg(reinterpret_cast<const uint4 *>(kimg + (size_t)(jg / KS) * KSTEP_BYTES + s * A_TILE +
                                                                     (size_t)((jg % KS) / 16) * 2048));
                    double vb[KNOT ? 16 : 1], kn[KNOT ? 8 : 1];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int jj = half * 64 + g * 16 + e;
                        long long q4[4];
                        if ((e & 3) == 0) join_quad(w, e >> 2, q4);
                        const bool ok = iv && j0 + jj < a.m;
                        if (KNOT || ok) {
                            const long long qd = q4[e & 3];
                            const double kij = a.sigma2 * FIX_INV * (double)qd;
                            if (ROWD) {
                                const double tij = cs[jj] * T[e];
                                const double tk = tij * kij;
                                rt[0] += tk;
                                rbt[0] = fma(bt[jj], kij, rbt[0]);
                                rkv = fma(kij, vv[jj], rkv);
                                if (!a.nodims) {
                                    const double bk = bt[jj] * kij;
#pragma unroll
                                    for (int k = 0; k < DT; k++) {
                                        const double tt = xi[k] - us[jj * DT + k], d2 = tt * tt;
                                        rt[(ROWD ? 1 + k : 0)] = fma(tk, d2, rt[(ROWD ? 1 + k : 0)]);
                                        rbt[(ROWD ? 1 + k : 0)] = fma(bk, d2, rbt[(ROWD ? 1 + k : 0)]);
                                    }
                                }
                                if (qd == FIX_ONE && !a.nodims)
                                    record_if_coincident_i8_part(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                                 a.coin_omega, a.coin_cap, tij, 0);
                            } else {
                                const double om = fma(rsi, cs[jj] * T[e], rai * bt[jj]);
                                const double pk = ok ? om * kij : 0.0;
                                s0 += pk;
#pragma unroll
                                for (int k = 0; k < DT; k++) {
                                    const double tt = xi[k] - us[jj * DT + k];
                                    sc[k] = fma(pk, tt * tt, sc[k]);
                                    if (KNOT) vb[KNOT ? (e & 1) * 8 + k : 0] = pk * tt;
                                }
                                if (ok && qd == FIX_ONE)
                                    record_if_coincident_i8_part(a.X, a.ldx, ig, a.U, a.m, j0 + jj, DT, a.coin_count, a.coin_list,
                                                                 a.coin_omega, a.coin_cap, om, 0);
                            }
                        }
                        if (KNOT) {
#pragma unroll
                            for (int k = DT; k < 8; k++) vb[KNOT ? (e & 1) * 8 + k : 0] = 0.0;
                            if (e & 1) kn[KNOT ? e >> 1 : 0] = warp_colsum16(reinterpret_cast<double (&)[16]>(vb), lane);
                        }
                    }
                    if (KNOT) knot_round_out<DT>(reinterpret_cast<const double (&)[8]>(kn), lane, kslot, a.mp, jg, a.first != 0);
#pragma unroll
                    for (int c = 0; c < 48; ++c) T[c] = T[c + 16];
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

// ------------------------------------------------------------------------------------------------
// The resident pass-2 image (GaussWS::k2): when the caller announces several K*M / row-form passes over the same K
// (GaussWS::k_reuse: FIC has two, the Laplace gradient four), the first of them generates its chunks straight into an
// image of the whole shard and the later ones read it -- exp and the digit split run once.  (Leaving the image behind from
// pass 1 by transposing its chunks was measured too: 14 GB of extra traffic through the L2 -> SM path that bounds the
// tensor kernels cost the Gram pass what the K*M pass gained, so a single-pass evaluation -- VI -- keeps its chunks in L2.)
// ------------------------------------------------------------------------------------------------
static bool k2_enabled()
{
    static const bool on = [] {
        const char *e = getenv("SRGP_K2");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

static size_t k2_chunk_bytes(const GaussWS *w) { return (size_t)w->rblocks * (w->mp / KS) * KSTEP_BYTES; }

static void k2_make_key(const srgp_ctx *ctx, const GaussWS *w, const GenParams &gp, GaussWS::K2Key &k)
{
    memset(&k, 0, sizeof(k));
    k.Xp = ctx->Xp;
    k.n = ctx->n;
    k.mp = w->mp, k.m = w->m, k.d = w->d;
    k.uver = w->u_version, k.xver = ctx->data_version;
    for (int c = 0; c < w->d && c < SRGP_MAX_D; c++) k.invl[c] = gp.invl[c];   // sigma^2 is not part of the image
}

// the image holds this shard's K for these parameters
static bool k2_have(const srgp_ctx *ctx, const GaussWS *w, const GenParams &gp)
{
    if (!w->k2_valid) return false;
    GaussWS::K2Key k;
    k2_make_key(ctx, w, gp, k);
    return memcmp(&k, &w->k2_key, sizeof(k)) == 0;
}

// Make room for the image of the current shard; false when it is switched off or does not fit (the caller then works
// from its double-buffered chunks as before).  Invalidates the previous contents.
static bool k2_reserve(srgp_ctx *ctx, GaussWS *w, bool ahead = false)
{
    w->k2_valid = false;
    w->k2_pending = false;
    if (!k2_enabled() || !(w->k_reuse || ahead) || ctx->n <= 0 || !i8_pass2_supported(w)) return false;
    const size_t chunks = (size_t)((ctx->n + w->rows2 - 1) / w->rows2);
    const size_t bytes = chunks * k2_chunk_bytes(w);
    if (bytes > w->k2.cap) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
        if (bytes > free_b + w->k2.cap || free_b + w->k2.cap - bytes < ((size_t)4 << 30)) return false;   // keep 4 GB of headroom
        if (w->k2.reserve(bytes) != SRGP_OK) {
            cudaGetLastError();
            return false;
        }
    }
    return true;
}

static void k2_commit(const srgp_ctx *ctx, GaussWS *w, const GenParams &gp)
{
    k2_make_key(ctx, w, gp, w->k2_key);
    w->k2_valid = true;
}

// Experiment switch (profiles/r02_km2_bound.txt): SRGP_PAIR=1 runs adjacent row blocks of the K*M pass as tcgen05 CTA pairs
// (cta_group::2, the Mop tiles fetched once per pair).  Measured equal to single CTAs (22.9 ms per evaluation either way:
// with the operand ring deep enough the pass is bound by its drains and its FP64 epilogue, not by operand traffic), so the
// simpler single-CTA protocol is the default.
static bool km_pairs()
{
    static const bool on = [] {
        const char *e = getenv("SRGP_PAIR");
        return e && atoi(e) == 1;
    }();
    return on;
}

template <int DT>
static void launch_gen_knotrows(cudaStream_t s, dim3 grid, size_t smem, const double *X, int64_t ldx, const double *r,
                                int64_t r0, int rows_valid, int rows_padded, const double *U, int m, int mp, int d,
                                const GenParams &p, int8_t *slices, double *b1part, int first,
                                const double *rw, const double *wmax, int8_t *slices_w, bool wsqrt, const double *Kr)
{
    if (rw && wsqrt && Kr)
        gen_slices_knotrows_kernel<0, 3><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                               slices, b1part, first, rw, wmax, nullptr, Kr);
    else if (rw && wsqrt)
        gen_slices_knotrows_kernel<DT, 2><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                                slices, b1part, first, rw, wmax, nullptr, nullptr);
    else if (rw)
        gen_slices_knotrows_kernel<DT, 1><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                                slices, b1part, first, rw, wmax, slices_w, nullptr);
    else
        gen_slices_knotrows_kernel<DT, 0><<<grid, 128, smem, s>>>(X, ldx, r, r0, rows_valid, rows_padded, U, m, mp, d, p,
                                                                slices, b1part, first, nullptr, nullptr, nullptr, nullptr);
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
                   double *b1, bool weight_nonneg)
{
    cudaStream_t s = ctx->stream;
    static DeviceOnce once;
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(i8_gram2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM2_SMEM));
    const int mp = w->mp, m = w->m, d = w->d;
    const int tiles = w->nt * (w->nt + 1) / 2;
    const int nsplit = std::max(1, std::min(16, ctx->sm_count / tiles));
    const int quantum = BK * nsplit;
    const bool wsqrt = rowweight && weight_nonneg;                  // one slice set sqrt(w) e instead of two (e, w e)
    const int sets = (rowweight && !wsqrt) ? 2 : 1;
    // rows per launch: the chunk buffer holds sets x 8 slices x rows x mp bytes, and one INT32 accumulator may sum at
    // most MAX_ROWS_PER_SPLIT rows
    int64_t rows1 = std::min<int64_t>((int64_t)w->chunk_elems / mp / sets, (int64_t)MAX_ROWS_PER_SPLIT * nsplit);
    rows1 = std::max<int64_t>(quantum, rows1 / quantum * quantum);
    const size_t tile_elems = (size_t)BM * BN2;
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
    static const int nb1 = getenv("SRGP_PASS1_BUFS") ? std::min(PASS1_BUFS, std::max(2, atoi(getenv("SRGP_PASS1_BUFS")))) : PASS1_BUFS;
    int cidx = 0;
    for (int64_t r0 = 0; r0 < ctx->n; r0 += rows1, cidx++) {
        const int rows_valid = (int)std::min<int64_t>(rows1, ctx->n - r0);
        const int rows_padded = (int)round_up(rows_valid, quantum);
        const int b = cidx % nb1;
        int8_t *slices = reinterpret_cast<int8_t *>(w->chunk.d() + (size_t)b * w->chunk_elems);
        int8_t *slices_w = sets == 2 ? slices + (size_t)rows_padded * mp * NS : slices;
        if (cidx >= nb1) SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_used[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GEN, sg);
            dim3 grid(mp / 128, w->gen_groups);
            const size_t smem = sizeof(double) * BK * (d + 2);
#define CALL(D) launch_gen_knotrows<D>(sg, grid, smem, ctx->Xp, ctx->n, rvec, r0, rows_valid, rows_padded, w->U.d(), m, mp, d, gp, slices, w->b1part.d(), first, rowweight, wmax, slices_w, wsqrt, w->pass1_kmat)
            SRGP_D_SWITCH_I8(d, CALL)
#undef CALL
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_gen[b], sg));
        SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_gen[b], 0));
        {
            KernelScope ks(ctx, SRGP_PROF_GRAM, s);
            i8_gram2_kernel<<<tiles * nsplit, THREADS, GRAM2_SMEM, s>>>(slices_w, slices, rows_padded / KS, nsplit,
                                                                        gp.sigma2 * gp.sigma2, wmax, wsqrt ? 1 : 0, w->Gpart.d(), first);
            SRGP_LAUNCH_CHECK();
        }
        SRGP_CUDA(cudaEventRecord(ctx->ev_used[b], s));
        first = 0;
    }
    {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        i8_gram2_finalize_kernel<<<dim3(tiles, 16), 256, 0, s>>>(w->Gpart.d(), nsplit, mp, G);
        SRGP_LAUNCH_CHECK();
        gram_sum_rows(s, w->b1part.d(), w->gen_groups, mp, b1);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

template <int DT>
static void launch_gen_datarows(cudaStream_t s, dim3 grid, const double *X, int64_t ldx, int64_t r0, int rows_valid,
                                const double *U, int m, int mp, const GenParams &p, int8_t *slices)
{
    gen_slices_datarows_kernel<DT><<<grid, 128, sizeof(double) * BK * DT, s>>>(X, ldx, r0, rows_valid, U, m, mp, p, slices);
}

template <int DT, bool ROWD, bool PAIR, bool KNOT = false>
static cudaError_t launch_km2_i8(cudaStream_t s, dim3 grid, int device, const KmI8Args &a)
{
    const size_t smem = Km2Cfg<PAIR>::RING_BYTES + sizeof(Bars2) + sizeof(double) * (BN2 * DT + 3 * BN2 + 8 * PART_STRIDE_I8) +
                        sizeof(float) * (BN2 * DT + BN2);
    static DeviceOnce once;
    if (once.need(device)) {
        cudaError_t e = cudaFuncSetAttribute(i8_km2_kernel<DT, ROWD, PAIR, KNOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
    return cudaLaunchKernelEx(&cfg, i8_km2_kernel<DT, ROWD, PAIR, KNOT>, a);
}

// mp <= 8192: one INT32 level accumulator sums up to NS pairs x 2^14 x mp over the knots (tc_i8.cuh)
bool i8_pass2_supported(const GaussWS *w)
{
    return w->d >= 1 && w->d <= 8 && w->mp <= MAX_ROWS_PER_SPLIT;
}

static int km_pass_i8(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, const double *Mop, const double *rs, const double *ra,
                      const double *beta, double *out, bool accumulate_slots, bool rowd, const double *vvec, double *rowd_out,
                      int64_t rowd_stride, bool nodims = false);

// one pass-2 chunk (rows r0 .. of the shard) of digit slices into `kslices`, on stream sg
static int gen_datarows_chunk(srgp_ctx *ctx, GaussWS *w, const GenParams &gp, cudaStream_t sg, int64_t r0, int rows_valid,
                              int8_t *kslices)
{
    const int mp = w->mp, m = w->m, d = w->d;
    KernelScope ks(ctx, SRGP_PROF_GEN, sg);
    dim3 grid(w->rblocks, std::min(mp / BK, 16));
#define CALL(D) launch_gen_datarows<D>(sg, grid, ctx->Xp, ctx->n, r0, rows_valid, w->U.d(), m, mp, gp, kslices)
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
    return SRGP_OK;
}

int gauss_pregen_k2(srgp_ctx *ctx, GaussWS *w, const GenParams &gp)
{
    if (getenv("SRGP_NO_OVERLAP") || k2_have(ctx, w, gp) || !k2_reserve(ctx, w, true)) return SRGP_OK;
    cudaStream_t s = ctx->stream, sg = ctx->stream4;
    const size_t chunks = (size_t)((ctx->n + w->rows2 - 1) / w->rows2);
    while (w->k2_ev.size() < chunks) {
        cudaEvent_t e;
        SRGP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        w->k2_ev.push_back(e);
    }
    SRGP_CUDA(cudaEventRecord(ctx->ev_fork, s));
    SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_fork, 0));
    size_t c = 0;
    for (int64_t r0 = 0; r0 < ctx->n; r0 += w->rows2, c++) {
        const int rows_valid = (int)std::min<int64_t>(w->rows2, ctx->n - r0);
        SRGP_TRY(gen_datarows_chunk(ctx, w, gp, sg, r0, rows_valid, reinterpret_cast<int8_t *>(w->k2.p) + c * k2_chunk_bytes(w)));
        SRGP_CUDA(cudaEventRecord(w->k2_ev[c], sg));
    }
    k2_commit(ctx, w, gp);
    w->k2_pending = true;
    return SRGP_OK;
}

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
    const size_t mbytes = (size_t)mp * mp * NS;
    SRGP_TRY(w->i8buf.reserve(mbytes + (size_t)mp * 16));
    int8_t *mslices = reinterpret_cast<int8_t *>(w->i8buf.p);
    double *colscale = reinterpret_cast<double *>(mslices + mbytes);
    // adjacent row blocks as tcgen05 CTA pairs (cta_group::2: the Mop tiles are fetched once per pair)
    const bool knot = w->want_knots && !rowd;                      // the knot-location gradient rides on the gradient sums
    const bool pair = km_pairs() && w->rblocks % 2 == 0 && !knot;
    if (knot) {
        w->knot_slots = w->rblocks * 4;                           // one slot per (row block, lane quadrant of the epilogue)
        const size_t kbytes = (size_t)w->knot_slots * d * mp * 8;
        SRGP_TRY(w->knotpart.reserve(kbytes));
        if (ctx->n == 0 && !accumulate_slots) SRGP_CUDA(cudaMemsetAsync(w->knotpart.p, 0, kbytes, s));
    }
    unsigned long long *colbits = reinterpret_cast<unsigned long long *>(colscale + mp);
    {
        SRGP_CUDA(cudaMemsetAsync(colbits, 0, (size_t)mp * 8, s));
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s, 2);
        mop_rowmax_kernel<<<dim3(mp / 128, 16), 128, 0, s>>>(Mop, mp, colbits);
        SRGP_LAUNCH_CHECK();
        slice_mop_kernel<<<dim3(mp / 128, KBm), 128, 0, s>>>(Mop, mp, mslices, pair ? 1 : 0, colbits, colscale);
        SRGP_LAUNCH_CHECK();
    }
    int first = accumulate_slots ? 0 : 1;
    if (!rowd && ctx->n == 0 && first) SRGP_CUDA(cudaMemsetAsync(w->part2.p, 0, (size_t)slots * PART_STRIDE_I8 * 8, s));
    const int nslots = nodims ? (vvec ? 2 : 1) : (d + 1) * (beta ? 2 : 1) + (vvec ? 1 : 0), groups = w->cgroups * 2;
    if (rowd) SRGP_TRY(w->rowdpart.reserve((size_t)groups * nslots * w->rows2 * 8));
    cudaStream_t sg = getenv("SRGP_NO_OVERLAP") ? s : ctx->stream3;
    // K comes from the resident image when an earlier pass of this evaluation left it; otherwise this pass generates its
    // chunks -- straight into the image when there is room for it (fill2), else into the two chunk buffers
    const bool have2 = k2_have(ctx, w, gp);
    const bool fill2 = !have2 && k2_reserve(ctx, w);
    SRGP_CUDA(cudaEventRecord(ctx->ev_fork, s));
    SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_fork, 0));
    int cidx = 0;
    for (int64_t r0 = 0; r0 < ctx->n; r0 += w->rows2, cidx++) {
        const int rows_valid = (int)std::min<int64_t>(w->rows2, ctx->n - r0);
        const int b = cidx & 1;
        int8_t *kslices = (have2 || fill2) ? reinterpret_cast<int8_t *>(w->k2.p) + (size_t)cidx * k2_chunk_bytes(w)
                                           : reinterpret_cast<int8_t *>(w->chunk.d() + (size_t)b * w->chunk_elems);
        if (!have2) {
            if (cidx >= 2 && !fill2) SRGP_CUDA(cudaStreamWaitEvent(sg, ctx->ev_used[b], 0));
            SRGP_TRY(gen_datarows_chunk(ctx, w, gp, sg, r0, rows_valid, kslices));
            SRGP_CUDA(cudaEventRecord(ctx->ev_gen[b], sg));
            SRGP_CUDA(cudaStreamWaitEvent(s, ctx->ev_gen[b], 0));
        } else if (w->k2_pending) {
            SRGP_CUDA(cudaStreamWaitEvent(s, w->k2_ev[cidx], 0));   // generated ahead (gauss_pregen_k2)
        }
        {
            KernelScope ks(ctx, SRGP_PROF_KM, s);
            KmI8Args a;
            a.kslices = kslices;
            a.mslices = mslices;
            a.colscale = colscale;
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
            // A short last chunk (row shards of multi-GPU runs: 125 000 rows = 13.2 chunks) would cost a full launch: its few
            // row blocks each walk all their column tiles.  Spread it over more column groups instead -- the slots are
            // addressed by the linear CTA index and all of them are summed at the end, so any grid within `slots` will do
            // once the first launch has initialised them.
            int cg = w->cgroups, rb = w->rblocks;
            if (!first && !rowd && !knot && !pair) {
                const int rb_valid = (rows_valid + BM - 1) / BM;
                while (cg * 2 <= mp / BN2 && (mp / BN2) % (cg * 2) == 0 && rb_valid * cg * 2 <= slots) cg *= 2;
                if (cg != w->cgroups) rb = rb_valid;
            }
            a.tiles_per_cta = (mp / BN2) / cg;
            a.cluster = pair ? 2 : 1;
            a.knot_part = knot ? w->knotpart.d() : nullptr;
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
            dim3 grid(rb, cg);
            cudaError_t e = cudaSuccess;
#define CALL2(D, R) (pair ? launch_km2_i8<D, R, true>(s, grid, ctx->device, a) : launch_km2_i8<D, R, false>(s, grid, ctx->device, a))
#define CALL(D) e = knot ? launch_km2_i8<D, false, false, true>(s, grid, ctx->device, a) : rowd ? CALL2(D, true) : CALL2(D, false)
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
    if (fill2) k2_commit(ctx, w, gp);
    w->k2_pending = false;
    if (out) {
        KernelScope ks(ctx, SRGP_PROF_REDUCE, s);
        gram_sum_part(s, w->part2.d(), slots, PART_STRIDE_I8, d + 1, out);
        SRGP_LAUNCH_CHECK();
    }
    return SRGP_OK;
}

}  // namespace srgp
