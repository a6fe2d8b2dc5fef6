// acc_rotation.cu -- does the tcgen05.mma.kind::i8 rate (M = 128, N = 128, K = 32, resident operands) depend on how many
// TMEM accumulators the MMA stream rotates over, and on a commit + mbarrier wait every 10 MMAs?  One issuing thread per SM.
#include <cstdio>
#include <cstdint>
#include "../../sparsergps_b200/csrc/tc_i8.cuh"
using namespace srgp::i8;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int SMEM = 131072;

#define PROBE_MMA(L, SA, SB, PRED) \
    "add.u32 td, %1, 128*" #L ";\n add.u64 da, %2, 256*" #SA ";\n add.u64 db, %3, 256*" #SB ";\n" \
    "tcgen05.mma.cta_group::1.kind::i8 [td], da, db, %4, {%8, %8, %8, %8}, " #PRED ";\n"
// the 10 MMAs of sweep 0 with an mbarrier test in flight: the test is issued first, its predicate is consumed last
__device__ __forceinline__ uint32_t issue_sweep0_probe(uint32_t tm, uint64_t da0, uint64_t db0, uint32_t keep, uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred pk, pw, pt;\n.reg .b64 da, db;\n.reg .b32 td;\n"
                 "setp.ne.b32 pk, %5, 0;\nsetp.eq.b32 pt, 0, 0;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 pw, [%6], %7;\n"
                 PROBE_MMA(0, 0, 0, pk) PROBE_MMA(1, 1, 0, pk) PROBE_MMA(2, 2, 0, pk) PROBE_MMA(3, 3, 0, pk)
                 PROBE_MMA(1, 0, 1, pt) PROBE_MMA(2, 1, 1, pt) PROBE_MMA(3, 2, 1, pt)
                 PROBE_MMA(2, 0, 2, pt) PROBE_MMA(3, 1, 2, pt) PROBE_MMA(3, 0, 3, pt)
                 "selp.u32 %0, 1, 0, pw;\n}"
                 : "=r"(ok) : "r"(tm), "l"(da0), "l"(db0), "r"(IDESC2), "r"(keep), "r"(smem_u32(bar)), "r"(parity), "r"(0u) : "memory");
    return ok;
}

// PATTERN: 0 = rotate over R accumulators (12 MMAs per iteration); 1 = the product's sweep-0 order (10 MMAs: acc
// 0,1,2,3,1,2,3,2,3,3); 2 = the product's sweep-1 order (18 MMAs over 3 accumulators).  Fully unrolled, compile-time
// descriptors offsets: the issue sequence is the product's.
template <int PATTERN, int R, int SYNC>
__global__ void __launch_bounds__(64, 1) k(int iters, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    __shared__ int flag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < SMEM / 4; i += 64) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { flag = 0; mbar_init(&bar, SYNC == 6 ? 2 : 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc_all(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot, sbase = smem_u32(smem);
    constexpr int CNT = PATTERN == 0 ? 12 : PATTERN == 1 ? 10 : 18;
    if ((warp == 0 || (SYNC == 6 && warp == 1)) && lane == 0) {
        const long long t0 = clock64();
        for (int it = (SYNC == 6 ? warp : 0); it < iters; it += (SYNC == 6 ? 2 : 1)) {
            const uint32_t st = sbase + (uint32_t)(it & 1) * 65536u;
            const uint64_t da0 = make_desc(st, 2048, 128), db0 = make_desc(st + 32768, 2048, 128);
            const uint32_t keep = it > 0 ? 1u : 0u;
            if (PATTERN == 0) {
#pragma unroll
                for (int q = 0; q < 12; ++q)
                    mma_i8_n128(tm + (uint32_t)(q % R) * 128u, da0 + (uint64_t)(((q & 7) * 4096) >> 4),
                                db0 + (uint64_t)((((q >> 2) & 7) * 4096) >> 4), keep);
            } else if (PATTERN == 1 && SYNC == 4) {
                if (!issue_sweep0_probe(tm, da0, db0, keep, &bar3, 1)) mbar_wait(&bar3, 1);
            } else if (PATTERN == 1) {
#pragma unroll
                for (int sb = 0; sb < 4; ++sb)
#pragma unroll
                    for (int sa = 0; sa < 4; ++sa)
                        if (sa + sb < 4)
                            mma_i8_n128(tm + (uint32_t)(sa + sb) * 128u, da0 + (uint64_t)((sa * 4096) >> 4),
                                        db0 + (uint64_t)((sb * 4096) >> 4), sb == 0 ? keep : 1u);
            } else {
#pragma unroll
                for (int sb = 0; sb < 7; ++sb)
#pragma unroll
                    for (int sa = 0; sa < 7; ++sa)
                        if (sa + sb >= 4 && sa + sb < 7)
                            mma_i8_n128(tm + (uint32_t)(sa + sb - 4) * 128u, da0 + (uint64_t)((sa * 4096) >> 4),
                                        db0 + (uint64_t)((sb * 4096) >> 4), sb == 0 ? keep : 1u);
            }
            if (SYNC) mma_commit(&bar2);                   // what the product does per k-step (nobody waits here)
            if ((SYNC >= 2 && SYNC < 4) || SYNC == 6) mbar_wait(&bar3, 1);            // a wait that succeeds at once (parity of the phase before the first)
            if (SYNC == 3) tc_fence_after();
            if (SYNC == 5) {                                   // a satisfied spin on a plain shared-memory word
                while (*reinterpret_cast<volatile int *>(&flag) <= it - 1000000) { }
                tc_fence_after();
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        if (warp == 0) cycles[blockIdx.x * 2] = clock64() - t0;
        cycles[blockIdx.x * 2 + 1] = (long long)iters * CNT;
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_free_all(tm); }
}

template <int PATTERN, int R, int SYNC>
int run(const char *name, long long *cyc)
{
    CK(cudaFuncSetAttribute(k<PATTERN, R, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int rep = 0; rep < 2; rep++) { k<PATTERN, R, SYNC><<<148, 64, SMEM>>>(3000, cyc); CK(cudaDeviceSynchronize()); }
    long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("{\"variant\": \"%s\", \"cycles_per_mma\": %.1f}\n", name, (double)h[0] / h[1]);
    return 0;
}

int main()
{
    long long *cyc; CK(cudaMalloc(&cyc, 148 * 16));
    run<0, 4, 0>("rotate over 4 accumulators", cyc);
    run<0, 3, 0>("rotate over 3", cyc);
    run<0, 2, 0>("rotate over 2", cyc);
    run<0, 1, 0>("same accumulator", cyc);
    run<1, 4, 0>("sweep-0 order (10 MMAs / k-step)", cyc);
    run<2, 3, 0>("sweep-1 order (18 MMAs / k-step)", cyc);
    run<1, 4, 1>("sweep-0 order + commit per k-step", cyc);
    run<2, 3, 1>("sweep-1 order + commit per k-step", cyc);
    run<1, 4, 2>("sweep-0 order + commit + satisfied mbarrier wait per k-step", cyc);
    run<2, 3, 2>("sweep-1 order + commit + satisfied mbarrier wait per k-step", cyc);
    run<1, 4, 4>("sweep-0 order + commit + satisfied mbarrier TEST issued before the MMAs, consumed after", cyc);
    run<1, 4, 5>("sweep-0 order + commit + satisfied spin on a shared-memory word + tcgen05.fence per k-step", cyc);
    run<2, 3, 5>("sweep-1 order + commit + satisfied spin on a shared-memory word + tcgen05.fence per k-step", cyc);
    run<1, 4, 6>("sweep-0 order, TWO issuing threads alternate k-steps, each: satisfied mbarrier wait + 10 MMAs + commit", cyc);
    run<2, 3, 6>("sweep-1 order, TWO issuing threads alternate k-steps, each: satisfied mbarrier wait + 18 MMAs + commit", cyc);
    run<1, 4, 3>("sweep-0 order + commit + satisfied wait + tcgen05.fence per k-step", cyc);
    return 0;
}
