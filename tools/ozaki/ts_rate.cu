// Microbenchmark (NOT product code): does feeding the A operand of tcgen05.mma.kind::i8 from TENSOR MEMORY lift the
// shared-memory-port bound of the 128 x 64 tiles?  With NS = 7 slices the 7 level accumulators take 448 of the 512 TMEM
// columns; the remaining 64 hold the 7 A slice tiles of ONE k-step (128 lanes x 32 bytes = 8 columns each), copied from
// shared memory by tcgen05.cp.128x256b once per k-step and then read by 7, 6, ... 1 MMAs.
//   phase 1: correctness -- the TS (A from TMEM) k-step must give the INT32 levels of the SS (A from smem) k-step, over
//            several k-steps with different data per stage (also proves cp/mma ordering without explicit waits);
//   phase 2: cycles per k-step, SS vs TS, operands resident in shared memory (no loads, no epilogue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_rate ts_rate.cu && ./ts_rate
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "tc_i8_r01.cuh"
using namespace srgp::i8;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

constexpr int NST = 2;                                   // smem stages with different data
constexpr uint32_t A_TMEM_COL = NS * BN;                 // 448: first column of the A slices

__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void issue_stage_ts(uint32_t sbase, uint32_t tmem_base, bool fresh)
{
    const uint64_t da0 = make_desc(sbase, 2048, 128);
    const uint64_t db0 = make_desc(sbase + NS * A_TILE, 1024, 128);
    const uint32_t keep = fresh ? 0u : 1u;
#pragma unroll
    for (int sa = 0; sa < NS; ++sa) cp_128x256b(tmem_base + A_TMEM_COL + 8u * sa, da0 + (uint64_t)((sa * A_TILE) >> 4));
#pragma unroll
    for (int sb = 0; sb < NS; ++sb)
#pragma unroll
        for (int sa = 0; sa < NS; ++sa)
            if (sa + sb < NS)
                mma_i8_ts(tmem_base + (uint32_t)(sa + sb) * BN, tmem_base + A_TMEM_COL + 8u * sa,
                          db0 + (uint64_t)((sb * B_TILE) >> 4), sb == 0 ? keep : 1u);
}

// mode 0: SS, mode 1: TS.  dump != nullptr: after `iters` k-steps (stage = it % NST) write all 7 x 64 level columns
__global__ void __launch_bounds__(192, 1) ts_kernel(int mode, int iters, const uint8_t *__restrict__ data, int *__restrict__ dump,
                                                    long long *__restrict__ cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NST * STAGE_BYTES / 16; i += 192)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(data)[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc_all(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot, sbase = smem_u32(smem);
    if (warp == 0 && lane == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = sbase + (uint32_t)(it % NST) * STAGE_BYTES;
            if (mode == 0) issue_stage(st, tm, it == 0);
            else issue_stage_ts(st, tm, it == 0);
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (dump && warp >= 2 && blockIdx.x == 0) {
        const int q = warp & 3, row = q * 32 + lane;
        for (int L = 0; L < NS; ++L)
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(L * BN + half * 32), v);
                for (int c = 0; c < 32; ++c) dump[(L * 128 + row) * 64 + half * 32 + c] = (int)v[c];
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_free_all(tm); }
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t bytes = (size_t)NST * STAGE_BYTES;
    uint8_t *h = (uint8_t *)malloc(bytes);
    srand(7);
    for (size_t i = 0; i < bytes; i++) h[i] = (uint8_t)(rand() & 0xff);
    uint8_t *d; CK(cudaMalloc(&d, bytes)); CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    int *dump; CK(cudaMalloc(&dump, NS * 128 * 64 * 4));
    long long *dc; CK(cudaMalloc(&dc, sms * 8));
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const size_t ne = (size_t)NS * 128 * 64;
    int *ref = (int *)malloc(ne * 4), *got = (int *)malloc(ne * 4);
    int bad_total = 0;
    for (int iters : {1, 2, 5}) {
        ts_kernel<<<1, 192, bytes>>>(0, iters, d, dump, dc); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ref, dump, ne * 4, cudaMemcpyDeviceToHost));
        ts_kernel<<<1, 192, bytes>>>(1, iters, d, dump, dc); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got, dump, ne * 4, cudaMemcpyDeviceToHost));
        int bad = 0; long long nz = 0;
        for (size_t i = 0; i < ne; i++) { bad += ref[i] != got[i]; nz += ref[i] != 0; }
        printf("{\"check\": \"TS levels == SS levels\", \"k_steps\": %d, \"mismatches\": %d, \"entries\": %zu, \"nonzero\": %lld}\n", iters, bad, ne, nz);
        bad_total += bad;
    }
    long long hc[256];
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) { ts_kernel<<<sms, 192, bytes>>>(mode, iters, d, nullptr, dc); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(hc, dc, sms * 8, cudaMemcpyDeviceToHost));
        printf("{\"variant\": \"%s\", \"slices\": %d, \"cycles_per_kstep\": %.1f, \"cycles_per_mma\": %.1f, \"mma_pipe_share\": %.3f}\n",
               mode == 0 ? "SS: A and B from shared memory" : "TS: A slices copied to TMEM once per k-step (tcgen05.cp.128x256b)",
               NS, (double)hc[0] / iters, (double)hc[0] / iters / NPAIRS, NPAIRS * 32.0 * iters / hc[0]);
    }
    return bad_total ? 1 : 0;
}
