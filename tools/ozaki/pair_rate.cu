// Microbenchmark (NOT product code): tcgen05.mma.cta_group::2.kind::i8 on a CTA pair (M = 256 = 2 x 128 rows, the B operand
// split between the two SMs), operands resident in shared memory.  Per CTA and MMA the shared-memory port then carries the
// 4 KB A tile plus HALF the B tile, so N = 64 should cost 40 clk instead of 48 and N = 128 should reach the 64 clk of the
// tensor pipe itself (tools/ozaki/mma_rate.cu: cta_group::1 N = 64 -> 48 clk, N = 128 -> 64 clk).
//   phase 1: the INT32 levels of both CTAs equal those of single-CTA MMAs on the same data (layout / descriptor check);
//   phase 2: cycles per MMA for N = 64 (7 levels resident) and N = 128 (4 accumulators).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_rate pair_rate.cu && ./pair_rate
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "tc_i8_r01.cuh"
using namespace srgp::i8;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <int N> __host__ __device__ constexpr uint32_t idesc_pair() { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }
template <int N> __host__ __device__ constexpr uint32_t idesc_single() { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

template <int N>
__device__ __forceinline__ void mma_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc_pair<N>()), "r"(acc), "r"(0u) : "memory");
}
template <int N>
__device__ __forceinline__ void mma_single(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc_single<N>()), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared-memory images (per CTA): A slices [NS][4 KB] (own 128 rows), then B slices.
//   PAIR: B holds this CTA's HALF of the N rows per slice: [NS][N/2 x 32 B], k-chunk stride (LBO) = N/2 * 16.
//   single: B holds all N rows per slice: [NS][N x 32 B], LBO = N * 16.
// LEVELS accumulators of N columns: level(sa, sb) = (sa + sb) % LEVELS (N = 128 has room for 4 only: timing + a
// folded-level check, which is still a complete test of the operand layouts).
template <int N, int LEVELS, bool PAIR>
__global__ void __launch_bounds__(192, 1) pair_kernel(int iters, const uint8_t *__restrict__ dataA, const uint8_t *__restrict__ dataB,
                                                     int *__restrict__ dump, long long *__restrict__ cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_rank() : (blockIdx.x & 1);
    const int pair_id = blockIdx.x >> 1;
    constexpr int BROWS = PAIR ? N / 2 : N, BT = BROWS * 32;
    // A: rows of this CTA; B: global image is [slice][c(2)][row N][16 B]; copy this CTA's rows
    const uint8_t *gA = dataA + (size_t)rank * NS * A_TILE;
    for (int i = threadIdx.x; i < NS * A_TILE / 16; i += 192) reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(gA)[i];
    for (int i = threadIdx.x; i < NS * BT / 16; i += 192) {
        const int s = i / (BT / 16), rem = i % (BT / 16), c = rem / BROWS, r = rem % BROWS;
        const int grow = PAIR ? (int)rank * BROWS + r : r;
        reinterpret_cast<uint4 *>(smem + NS * A_TILE)[i] = reinterpret_cast<const uint4 *>(dataB)[(s * 2 + c) * N + grow];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc_all(&slot);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot, sbase = smem_u32(smem);
    if (warp == 0 && lane == 0 && (!PAIR || rank == 0)) {
        const uint64_t da0 = make_desc(sbase, 2048, 128), db0 = make_desc(sbase + NS * A_TILE, BROWS * 16, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int sb = 0; sb < NS; ++sb)
#pragma unroll
                for (int sa = 0; sa < NS; ++sa)
                    if (sa + sb < NS) {
                        const uint64_t da = da0 + (uint64_t)((sa * A_TILE) >> 4), db = db0 + (uint64_t)((sb * BT) >> 4);
                        const int L = (sa + sb) % LEVELS;
                        const uint32_t acc = (it > 0 || sa + sb >= LEVELS || sb > 0) ? 1u : 0u;     // first touch of accumulator L: (sa = L, sb = 0)
                        if (PAIR) mma_pair<N>(tm + (uint32_t)L * N, da, db, acc);
                        else mma_single<N>(tm + (uint32_t)L * N, da, db, acc);
                    }
        }
        if (PAIR)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                         "h"((uint16_t)3) : "memory");
        else
            mma_commit(&bar);
        cycles[blockIdx.x] = clock64() - t0;          // issue time; completion below
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        mbar_wait(&bar, 0);
        if (!PAIR || rank == 0) cycles[blockIdx.x] += clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (dump && warp >= 2 && pair_id == 0) {
        const int q = warp & 3, row = q * 32 + lane;
        for (int L = 0; L < LEVELS; ++L)
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(L * N + c0), v);
                for (int c = 0; c < 32; ++c) dump[(((int)rank * LEVELS + L) * 128 + row) * N + c0 + c] = (int)v[c];
            }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
        else tmem_free_all(tm);
    }
}

template <int N, int LEVELS, bool PAIR>
static void launch(int grid, int iters, const uint8_t *dA, const uint8_t *dB, int *dump, long long *dc)
{
    const size_t smem = (size_t)NS * A_TILE + (size_t)NS * (PAIR ? N / 2 : N) * 32;
    CK(cudaFuncSetAttribute(pair_kernel<N, LEVELS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, pair_kernel<N, LEVELS, PAIR>, iters, dA, dB, dump, dc));
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
}

template <int N, int LEVELS>
static int run(int sms, const uint8_t *dA, const uint8_t *dB, int *dump, long long *dc)
{
    const size_t ne = (size_t)2 * LEVELS * 128 * N;
    int *ref = (int *)malloc(ne * 4), *got = (int *)malloc(ne * 4);
    int bad_total = 0;
    for (int iters : {1, 3}) {
        launch<N, LEVELS, false>(2, iters, dA, dB, dump, dc);
        CK(cudaMemcpy(ref, dump, ne * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemset(dump, 0, ne * 4));
        launch<N, LEVELS, true>(2, iters, dA, dB, dump, dc);
        CK(cudaMemcpy(got, dump, ne * 4, cudaMemcpyDeviceToHost));
        int bad = 0; long long nz = 0;
        for (size_t i = 0; i < ne; i++) { bad += ref[i] != got[i]; nz += ref[i] != 0; }
        printf("{\"check\": \"pair (cta_group::2, M=256) == 2 x single (M=128)\", \"N\": %d, \"k_steps\": %d, \"mismatches\": %d, \"entries\": %zu, \"nonzero\": %lld}\n", N, iters, bad, ne, nz);
        bad_total += bad;
    }
    long long hc[256];
    const int iters = 3000, grid = sms & ~1;
    for (int pair = 0; pair < 2; ++pair) {
        for (int rep = 0; rep < 2; ++rep) {
            if (pair) launch<N, LEVELS, true>(grid, iters, dA, dB, nullptr, dc); else launch<N, LEVELS, false>(grid, iters, dA, dB, nullptr, dc);
        }
        CK(cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost));
        const double per = (double)hc[0] / iters / NPAIRS;
        printf("{\"variant\": \"%s N=%d\", \"cycles_per_mma\": %.1f, \"tensor_pipe_share\": %.3f}\n", pair ? "cta_group::2 M=256" : "cta_group::1 M=128", N,
               per, (N / 2.0) / per);
    }
    free(ref); free(got);
    return bad_total;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t bytesA = (size_t)2 * NS * A_TILE, bytesB = (size_t)NS * 2 * 128 * 16;
    uint8_t *hA = (uint8_t *)malloc(bytesA), *hB = (uint8_t *)malloc(bytesB);
    srand(11);
    for (size_t i = 0; i < bytesA; i++) hA[i] = (uint8_t)(rand() & 0xff);
    for (size_t i = 0; i < bytesB; i++) hB[i] = (uint8_t)(rand() & 0xff);
    uint8_t *dA, *dB; CK(cudaMalloc(&dA, bytesA)); CK(cudaMalloc(&dB, bytesB));
    CK(cudaMemcpy(dA, hA, bytesA, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, bytesB, cudaMemcpyHostToDevice));
    int *dump; CK(cudaMalloc(&dump, (size_t)2 * 8 * 128 * 128 * 4));
    long long *dc; CK(cudaMalloc(&dc, 256 * 8));
    int bad = 0;
    bad += run<64, 7>(sms, dA, dB, dump, dc);
    bad += run<128, 4>(sms, dA, dB, dump, dc);
    return bad ? 1 : 0;
}
