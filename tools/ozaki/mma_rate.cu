// Microbenchmark (NOT product code): issue rate of tcgen05.mma.kind::i8 from shared-memory operands that are already
// resident, for the two tile shapes of the Ozaki scheme:  N = 64 with 8 resident levels (36 pairs / k-step) and
// N = 128 with 4 resident levels (10 + 26 pairs in two sweeps).  No loads, no epilogue: cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_i8_r01.cuh"
using namespace srgp::i8;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <int N>
__device__ __forceinline__ void mma_n(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t acc)
{
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// variant 0: 36 MMAs 128x64x32 per iteration (8 A tiles of 4 KB, 8 B tiles of 2 KB, levels 0..7 x 64 columns)
// variant 1: 36 MMAs 128x128x32 per iteration (8 A + 8 B tiles of 4 KB = one 32-byte k-step, 4 accumulators x 128 columns)
template <int VARIANT>
__global__ void __launch_bounds__(64, 1) rate_kernel(int iters, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 196608 / 4; i += 64) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc_all(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot, sbase = smem_u32(smem);
    if (warp == 0 && lane == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (VARIANT == 0) {
                issue_stage(sbase + (it & 3) * STAGE_BYTES, tm, it == 0);
            } else {
                const uint32_t st = sbase + (it % 3) * 65536;
                const uint64_t da0 = make_desc(st, 2048, 128), db0 = make_desc(st + 32768, 2048, 128);
#pragma unroll
                for (int sb = 0; sb < NS; ++sb)
#pragma unroll
                    for (int sa = 0; sa < NS; ++sa)
                        if (sa + sb < NS) {
                            // 4 KB tiles: [c(2)][r1(16)][r0(8)][16]
                            const uint64_t da = da0 + (uint64_t)((sa * 4096) >> 4), db = db0 + (uint64_t)((sb * 4096) >> 4);
                            mma_n<128>(tm + (uint32_t)((sa + sb) & 3) * 128, da, db, (it > 0 || sb > 0) ? 1u : 0u);
                        }
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_free_all(tm); }
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    long long *dc; CK(cudaMalloc(&dc, 148 * 8));
    long long hc[148];
    const int iters = 2000;
    CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024));
    CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 1024));
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            if (v == 0) rate_kernel<0><<<prop.multiProcessorCount, 64, 196608 + 1024>>>(iters, dc);
            else rate_kernel<1><<<prop.multiProcessorCount, 64, 196608 + 1024>>>(iters, dc);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
        const double mmas = 36.0 * iters, ops = mmas * 128.0 * (v == 0 ? 64 : 128) * 32 * 2;
        printf("{\"variant\": \"%s\", \"cycles_per_mma\": %.1f, \"int8_ops_per_clk_per_sm\": %.0f, \"tops_at_1965MHz\": %.0f}\n",
               v == 0 ? "N=64, 8 levels" : "N=128, 4 levels", hc[0] / mmas, ops / hc[0], ops / hc[0] * 148 * 1.965e9 / 1e12);
    }
    return 0;
}
