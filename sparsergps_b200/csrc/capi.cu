// capi.cu -- context life cycle, device-memory helpers, timers, per-kernel accounting.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace srgp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int use_device(srgp_ctx *ctx)
{
    if (!ctx) {
        set_error("null context");
        return SRGP_ERR_ARG;
    }
    SRGP_CUDA(cudaSetDevice(ctx->device));
    return SRGP_OK;
}

__global__ void fill_kernel(double *p, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// Counter-based N(mean, sd): splitmix64 -> two uniforms -> Box-Muller.  Deterministic in (seed, index).
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_normal_kernel(double *p, int64_t n, uint64_t seed, double mean, double sd)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t a = splitmix64(seed * 0x100000001B3ull + 2 * (uint64_t)i);
        uint64_t b = splitmix64(seed * 0x100000001B3ull + 2 * (uint64_t)i + 1);
        double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
        double u2 = (b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
        p[i] = mean + sd * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

}  // namespace srgp

using namespace srgp;

extern "C" int srgp_version(void) { return SRGP_VERSION; }
extern "C" const char *srgp_last_error(void) { return g_err; }

extern "C" void srgp_ctx_destroy(srgp_ctx *ctx);

extern "C" int srgp_ctx_create(int device, srgp_ctx **out)
{
    if (!out) {
        set_error("null output pointer");
        return SRGP_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return SRGP_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (%d devices)", device, count);
        return SRGP_ERR_ARG;
    }
    SRGP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SRGP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libsrgp is built for sm_100a only", device, prop.major, prop.minor);
        return SRGP_ERR_CUDA;
    }
    srgp_ctx *ctx = new srgp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // a failed step must not leak the context and what it already owns (srgp_ctx_destroy skips null handles)
    auto build = [&]() -> int {
        SRGP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        SRGP_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        SRGP_CUDA(cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking));
        // VI generates pass 2's K while the m x m stage runs (gauss_pregen_k2): that work goes to the lowest-priority
        // stream so the short, dependent kernels of the stage do not queue behind waves of generator CTAs (measured gain
        // <= 1 %, at the run-to-run noise: 45.0 -> 44.5 ms at 1e6 rows).  The chunk generators of the passes themselves
        // stay on stream3 at the default priority: starving them delays the pass that waits for them (FIC 98.7 -> 102.4 ms
        // with every generator on the low priority).
        int prio_least = 0, prio_greatest = 0;
        SRGP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        SRGP_CUDA(cudaStreamCreateWithPriority(&ctx->stream4, cudaStreamNonBlocking, prio_least));
        for (int k = 0; k < 4; k++) {
            SRGP_CUDA(cudaEventCreateWithFlags(&ctx->ev_gen[k], cudaEventDisableTiming));
            SRGP_CUDA(cudaEventCreateWithFlags(&ctx->ev_used[k], cudaEventDisableTiming));
        }
        SRGP_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        SRGP_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        SRGP_CUDA(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
        SRGP_CUDA(cudaEventCreate(&ctx->tim0));
        SRGP_CUDA(cudaEventCreate(&ctx->tim1));
        return SRGP_OK;
    };
    const int st = build();
    if (st != SRGP_OK) {
        srgp_ctx_destroy(ctx);
        return st;
    }
    *out = ctx;
    return SRGP_OK;
}

extern "C" int srgp_comm_destroy(srgp_ctx *ctx);

extern "C" void srgp_ctx_destroy(srgp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    srgp_comm_destroy(ctx);
    if (ctx->ws && ctx->ws_free) ctx->ws_free(ctx->ws);
    srgp::DevBuf *bufs[] = {&ctx->in_x, &ctx->in_xp, &ctx->out_mat, &ctx->tmp0, &ctx->tmp1,
                            &ctx->flush, &ctx->X, &ctx->y, &ctx->mu};
    for (auto *b : bufs) b->release();
    for (auto &slot : ctx->prof)
        for (auto &pr : slot.pending) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    cudaEvent_t evs[] = {ctx->ev_fork, ctx->ev_join, ctx->ev_aux, ctx->tim0, ctx->tim1, ctx->tl_ref,
                         ctx->ev_gen[0], ctx->ev_gen[1], ctx->ev_gen[2], ctx->ev_gen[3],
                         ctx->ev_used[0], ctx->ev_used[1], ctx->ev_used[2], ctx->ev_used[3]};
    for (auto e : evs)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->stream4) cudaStreamDestroy(ctx->stream4);
    delete ctx;
}

extern "C" int srgp_ctx_sync(srgp_ctx *ctx)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream2));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream3));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream4));
    return SRGP_OK;
}

extern "C" int srgp_dev_alloc(srgp_ctx *ctx, int64_t bytes, void **out_dev)
{
    SRGP_TRY(use_device(ctx));
    if (!out_dev || bytes <= 0) {
        set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    SRGP_CUDA(cudaMalloc(out_dev, (size_t)bytes));
    return SRGP_OK;
}

extern "C" int srgp_dev_free(srgp_ctx *ctx, void *dev)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaFree(dev));
    return SRGP_OK;
}

extern "C" int srgp_memcpy_h2d(srgp_ctx *ctx, void *dst_dev, const void *src, int64_t bytes)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaMemcpyAsync(dst_dev, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SRGP_OK;
}

extern "C" int srgp_memcpy_d2h(srgp_ctx *ctx, void *dst, const void *src_dev, int64_t bytes)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaMemcpyAsync(dst, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SRGP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SRGP_OK;
}

extern "C" int srgp_fill_normal_dev(srgp_ctx *ctx, double *dst_dev, int64_t n, uint64_t seed, double mean,
                                    double sd)
{
    SRGP_TRY(use_device(ctx));
    ctx->launches++;
    fill_normal_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dst_dev, n, seed, mean, sd);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}

extern "C" int srgp_timer_start(srgp_ctx *ctx)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaEventRecord(ctx->tim0, ctx->stream));
    return SRGP_OK;
}

extern "C" int srgp_timer_stop_ms(srgp_ctx *ctx, double *ms)
{
    SRGP_TRY(use_device(ctx));
    SRGP_CUDA(cudaEventRecord(ctx->tim1, ctx->stream));
    SRGP_CUDA(cudaEventSynchronize(ctx->tim1));
    float f = 0.f;
    SRGP_CUDA(cudaEventElapsedTime(&f, ctx->tim0, ctx->tim1));
    if (ms) *ms = (double)f;
    return SRGP_OK;
}

extern "C" int srgp_prof_enable(srgp_ctx *ctx, int on)
{
    if (!ctx) return SRGP_ERR_ARG;
    ctx->prof_on = on != 0;
    return SRGP_OK;
}

// SRGP_TIMELINE=<file> (diagnostic): every accounted kernel bracket is appended as "class start_ms end_ms" relative to the
// last srgp_prof_reset -- where the time of an evaluation goes between the passes (tools/timeline.py).
static int prof_drain(srgp_ctx *ctx, int id)
{
    auto &slot = ctx->prof[id];
    static const char *tl_path = getenv("SRGP_TIMELINE");
    FILE *tl = (tl_path && ctx->tl_ref && !slot.pending.empty()) ? fopen(tl_path, "a") : nullptr;
    size_t k = 0;
    for (auto &pr : slot.pending) {
        SRGP_CUDA(cudaEventSynchronize(pr.second));
        float f = 0.f;
        SRGP_CUDA(cudaEventElapsedTime(&f, pr.first, pr.second));
        if (tl) {
            float t0 = 0.f;
            const int tag = k < slot.tags.size() ? slot.tags[k] : 0;
            if (cudaEventElapsedTime(&t0, ctx->tl_ref, pr.first) == cudaSuccess)
                fprintf(tl, "%d %.4f %.4f %d\n", id, t0, t0 + f, tag);
            else cudaGetLastError();
        }
        k++;
        slot.ms += (double)f;
        ctx->ev_pool.push_back(pr.first);
        ctx->ev_pool.push_back(pr.second);
    }
    slot.pending.clear();
    slot.tags.clear();
    if (tl) fclose(tl);
    return SRGP_OK;
}

extern "C" int srgp_prof_reset(srgp_ctx *ctx)
{
    SRGP_TRY(use_device(ctx));
    for (int id = 0; id < SRGP_PROF_COUNT; id++) {
        SRGP_TRY(prof_drain(ctx, id));
        ctx->prof[id].launches = 0;
        ctx->prof[id].ms = 0.0;
    }
    if (getenv("SRGP_TIMELINE")) {
        if (!ctx->tl_ref) SRGP_CUDA(cudaEventCreate(&ctx->tl_ref));
        SRGP_CUDA(cudaEventRecord(ctx->tl_ref, ctx->stream));
        if (FILE *tl = fopen(getenv("SRGP_TIMELINE"), "a")) {
            fprintf(tl, "# reset\n");
            fclose(tl);
        }
    }
    return SRGP_OK;
}

extern "C" int srgp_prof_get(srgp_ctx *ctx, int id, int64_t *launches, double *ms)
{
    SRGP_TRY(use_device(ctx));
    if (id < 0 || id >= SRGP_PROF_COUNT) {
        set_error("bad profile id %d", id);
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(prof_drain(ctx, id));
    if (launches) *launches = ctx->prof[id].launches;
    if (ms) *ms = ctx->prof[id].ms;
    return SRGP_OK;
}

extern "C" int64_t srgp_launch_count(srgp_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int srgp_flush_l2(srgp_ctx *ctx)
{
    SRGP_TRY(use_device(ctx));
    const size_t bytes = size_t(256) << 20;  // 256 MiB > 126 MB L2
    SRGP_TRY(ctx->flush.reserve(bytes));
    ctx->launches++;
    fill_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->flush.d(), (int64_t)(bytes / 8), 0.0);
    SRGP_LAUNCH_CHECK();
    return SRGP_OK;
}
