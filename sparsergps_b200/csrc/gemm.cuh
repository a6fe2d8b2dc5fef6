// gemm.cuh -- FP64 tensor-core (DMMA) tile engine shared by the Gram pass, the K*M pass and the m x m stage.
//
// FP64 on sm_100a tensor cores is `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4); tcgen05 has no f64 kind, so the
// accumulators live in registers and the Blackwell-specific part is the operand feed: a dedicated producer
// warp streams operand tiles global(L2) -> shared with the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP)
// completing on mbarriers, through a 4-stage full/empty pipeline; 8 consumer warps issue DMMA only.
//
// CTA tile 128 x 128 x 16, consumer warp tile 64 x 32 (8 x 4 DMMA tiles, 64 accumulator doubles / thread).
// An operand is either MN-contiguous (element (mn, k) at p[mn + k*ld]; 16 bulk copies of 1 KB per stage,
// shared layout [16][128+4]) or K-contiguous (element at p[k + mn*ld]; 128 bulk copies of 128 B, layout
// [128][16+4]).  The +4 padding makes every DMMA fragment load bank-conflict free (8-byte banks:
// 132 = 4 mod 16 and 20 = 4 mod 16, so a half-warp's 4 x 4 (k, mn) addresses hit 16 distinct banks).
//
// All extents are multiples of the tile (buffers are zero/identity padded by the callers), operand base
// pointers and leading dimensions are 16-byte aligned, so there are no tails anywhere in the main loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srgp {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDMN = BM + 4;   // MN-contiguous shared row stride (doubles)
constexpr int LDKC = BK + 4;   // K-contiguous shared row stride (doubles)
constexpr int TILE_DOUBLES = BM * LDKC;   // 2560 >= BK * LDMN = 2112
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
// + one producer WARPGROUP.  Only its first warp issues copies; the group exists so that registers can move between
// the roles (setmaxnreg works on warpgroups): a CTA of 384 threads starts with 168 registers per thread, the producer
// group drops to 40 and the two consumer groups rise to 232 -- enough for 64 accumulators plus double-buffered
// fragments, where 168 made ptxas rotate accumulators through ~100 moves per k-tile and spill in the main loop.
constexpr int PRODUCER_THREADS = 128;
constexpr int THREADS = CONSUMER_THREADS + PRODUCER_THREADS;
constexpr int CONSUMER_REGS = 232, PRODUCER_REGS = 40;

struct __align__(128) Smem {
    double a[STAGES][TILE_DOUBLES];
    double b[STAGES][TILE_DOUBLES];
    double w[STAGES][BK];
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
template <int N>
__device__ __forceinline__ void reg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_THREADS) : "memory"); }

// ---- pipeline state -------------------------------------------------------------------------------
// `it` counts k-tiles issued/consumed since the barriers were initialised; producer and consumers advance
// it identically, so one CTA can run several main loops back to back without re-initialising.
__device__ __forceinline__ void pipeline_init(Smem &sm)
{
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

__device__ __forceinline__ bool is_producer() { return threadIdx.x >= CONSUMER_THREADS; }
__device__ __forceinline__ bool is_producer_lead() { return (threadIdx.x >> 5) == CONSUMER_WARPS; }

// Accumulator fragment coordinates of this thread inside the 128 x 128 CTA tile.
__device__ __forceinline__ int frag_row(int mi)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp & 1) * 64 + mi * 8 + (lane >> 2);
}
__device__ __forceinline__ int frag_col(int ni)   // column of element 0; element 1 is +1
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return (warp >> 1) * 32 + ni * 8 + 2 * (lane & 3);
}

// acc += op(A)[128 x K] * op(B)[128 x K]^T over `ktiles` k-tiles starting at the given tile origins.
//   A_KC / B_KC : operand is K-contiguous (see header).   WEIGHT : B fragments are scaled by w[k].
// A, B point at the tile origin: MN-contiguous -> &p[mn0 + k0*ld]; K-contiguous -> &p[k0 + mn0*ld].
// The two halves of the main loop.  Kernels split into roles right after their common prologue:
//     if (is_producer()) { reg_dec<PRODUCER_REGS>(); if (is_producer_lead()) producer_issue(...); ... return; }
//     reg_inc<CONSUMER_REGS>(); ... consumer_mma(...); epilogue
// A, B point at the tile origin: MN-contiguous -> &p[mn0 + k0*ld]; K-contiguous -> &p[k0 + mn0*ld].
//   A_KC / B_KC : operand is K-contiguous (see header).   WEIGHT : B fragments are scaled by w[k].
template <bool A_KC, bool B_KC, bool WEIGHT>
__device__ __forceinline__ void producer_issue(Smem &sm, const double *__restrict__ A, int64_t lda,
                                               const double *__restrict__ B, int64_t ldb, const double *__restrict__ w,
                                               int ktiles, uint32_t &it)
{
    const int lane = threadIdx.x & 31;
    constexpr uint32_t bytes = (uint32_t)(2 * BM * BK * sizeof(double)) + (WEIGHT ? BK * sizeof(double) : 0);
    for (int kt = 0; kt < ktiles; kt++, it++) {
        const int s = it % STAGES;
        const uint32_t round = it / STAGES;
        if (round > 0) mbar_wait(&sm.empty[s], (round - 1) & 1);
        if (lane == 0) mbar_expect_tx(&sm.full[s], bytes);
        __syncwarp();
        const int64_t k0 = (int64_t)kt * BK;
        if (!A_KC) {
            if (lane < BK) bulk_g2s(&sm.a[s][lane * LDMN], A + (k0 + lane) * lda, BM * sizeof(double), &sm.full[s]);
        } else {
#pragma unroll
            for (int r = 0; r < BM / 32; r++) {
                const int row = lane + 32 * r;
                bulk_g2s(&sm.a[s][row * LDKC], A + k0 + (int64_t)row * lda, BK * sizeof(double), &sm.full[s]);
            }
        }
        if (!B_KC) {
            if (lane >= 32 - BK) {
                const int row = lane - (32 - BK);
                bulk_g2s(&sm.b[s][row * LDMN], B + (k0 + row) * ldb, BN * sizeof(double), &sm.full[s]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < BN / 32; r++) {
                const int row = lane + 32 * r;
                bulk_g2s(&sm.b[s][row * LDKC], B + k0 + (int64_t)row * ldb, BK * sizeof(double), &sm.full[s]);
            }
        }
        if (WEIGHT && lane == 0) bulk_g2s(&sm.w[s][0], w + k0, BK * sizeof(double), &sm.full[s]);
    }
}

// acc += op(A)[128 x K] * op(B)[128 x K]^T over `ktiles` k-tiles (consumer warps)
template <bool A_KC, bool B_KC, bool WEIGHT>
__device__ __forceinline__ void consumer_mma(Smem &sm, int ktiles, uint32_t &it, double (&acc)[8][4][2])
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int am = (warp & 1) * 64 + g;    // + mi*8
    const int bn = (warp >> 1) * 32 + g;   // + ni*8
    for (int kt = 0; kt < ktiles; kt++, it++) {
        const int s = it % STAGES;
        mbar_wait(&sm.full[s], (it / STAGES) & 1);
        const double *as = sm.a[s];
        const double *bs = sm.b[s];
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            const int k = kk * 4 + t;
            double af[8], bf[4];
#pragma unroll
            for (int mi = 0; mi < 8; mi++) af[mi] = A_KC ? as[(am + mi * 8) * LDKC + k] : as[k * LDMN + am + mi * 8];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) bf[ni] = B_KC ? bs[(bn + ni * 8) * LDKC + k] : bs[k * LDMN + bn + ni * 8];
            if (WEIGHT) {
                const double wk = sm.w[s][k];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bf[ni] *= wk;
            }
#pragma unroll
            for (int mi = 0; mi < 8; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
}

__device__ __forceinline__ void zero_acc(double (&acc)[8][4][2])
{
#pragma unroll
    for (int mi = 0; mi < 8; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

}  // namespace gemm
}  // namespace srgp
