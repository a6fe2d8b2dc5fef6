// probe.cu -- measured INT8 tensor-pipe peak for the roofline denominator of bench.py.
// MEASURED_PEAKS.json has no INT8 entry, so the denominator of the INT8 row passes is measured in the same run: every
// SM issues tcgen05.mma.kind::i8 M = 128, N = 128, K = 32 from shared-memory operands that are already resident (no
// loads, no epilogue), 4 accumulators x 128 TMEM columns -- the shape at which the shared-memory operand reads
// (8 KB per MMA at 128 B/clk) no longer exceed the MMA's own 64 cycles, i.e. the tensor pipe itself.
// (tools/ozaki/mma_rate.cu is the standalone version; profiles/r01_ozaki_mma_rate.json: 64.0 clk per MMA.)
#include "common.cuh"
#include "tc_i8.cuh"

namespace srgp {
using namespace i8;

constexpr int PROBE_SMEM = 131072;          // 2 x (8 A tiles + 8 B tiles of 4 KB)
constexpr uint32_t IDESC_N128 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(64, 1) i8_rate_probe_kernel(int iters, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < PROBE_SMEM / 4; i += 64) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc_all(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy fills -> async-proxy (MMA) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot, sbase = smem_u32(smem);
    if (warp == 0 && lane == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = sbase + (uint32_t)(it & 1) * 65536u;
            const uint64_t da0 = make_desc(st, 2048, 128), db0 = make_desc(st + 32768, 2048, 128);
#pragma unroll
            for (int k = 0; k < 32; ++k) {                           // 32 MMAs per iteration, operands rotate over 8 + 8 tiles
                const uint64_t da = da0 + (uint64_t)(((k & 7) * 4096) >> 4), db = db0 + (uint64_t)((((k >> 2) & 7) * 4096) >> 4);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}" ::"r"(tm + (uint32_t)(k & 3) * 128u),
                             "l"(da), "l"(db), "r"(IDESC_N128), "r"((it > 0 || k > 3) ? 1u : 0u), "r"(0u) : "memory");
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_free_all(tm);
    }
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_i8_slices(void) { return i8::NS; }

// tops: INT8 operations per second / 1e12 over the whole GPU from CUDA-event time; cycles_per_mma: SM-clock cycles per
// 128 x 128 x 32 MMA seen by the issuing thread of SM 0 (64 = pipe peak, 8192 MAC/clk/SM).
extern "C" int srgp_probe_i8_peak(srgp_ctx *ctx, int iters, double *tops, double *cycles_per_mma)
{
    SRGP_TRY(use_device(ctx));
    if (iters <= 0) iters = 4000;
    static DeviceOnce once;
    if (once.need(ctx->device))
        SRGP_CUDA(cudaFuncSetAttribute(i8_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM));
    SRGP_TRY(ctx->tmp0.reserve((size_t)ctx->sm_count * 8));
    long long *dc = static_cast<long long *>(ctx->tmp0.p);
    cudaStream_t s = ctx->stream;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {                               // first repetition = warm-up
        SRGP_CUDA(cudaEventRecord(ctx->tim0, s));
        ctx->launches++;
        i8_rate_probe_kernel<<<ctx->sm_count, 64, PROBE_SMEM, s>>>(iters, dc);
        SRGP_LAUNCH_CHECK();
        SRGP_CUDA(cudaEventRecord(ctx->tim1, s));
        SRGP_CUDA(cudaEventSynchronize(ctx->tim1));
        float ms = 0.f;
        SRGP_CUDA(cudaEventElapsedTime(&ms, ctx->tim0, ctx->tim1));
        if (rep > 0 && ms < best) best = ms;
    }
    long long c0 = 0;
    SRGP_CUDA(cudaMemcpy(&c0, dc, 8, cudaMemcpyDeviceToHost));
    const double mmas = 32.0 * iters;
    if (tops) *tops = mmas * ctx->sm_count * (2.0 * 128 * 128 * 32) / (best * 1e-3) / 1e12;
    if (cycles_per_mma) *cycles_per_mma = (double)c0 / mmas;
    return SRGP_OK;
}
