// common.cuh -- context, error plumbing, scratch buffers, per-kernel accounting.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/srgp.h"

namespace srgp {

void set_error(const char *fmt, ...);

#define SRGP_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            srgp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                 \
                            cudaGetErrorString(e__));                                            \
            return SRGP_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

#define SRGP_TRY(call)                                                                           \
    do {                                                                                         \
        int s__ = (call);                                                                        \
        if (s__ != SRGP_OK) return s__;                                                          \
    } while (0)

#define SRGP_LAUNCH_CHECK()                                                                      \
    do {                                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) {                                                                \
            srgp::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,             \
                            cudaGetErrorString(e__));                                            \
            return SRGP_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return SRGP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        // round up so small growth steps do not thrash the allocator
        size_t want = (bytes + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            p = nullptr;
            return SRGP_ERR_CUDA;
        }
        cap = want;
        return SRGP_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    double *d() const { return static_cast<double *>(p); }
};

// releases a set of local DevBufs on every exit path of a function
struct DevBufScope {
    DevBuf *bufs[8];
    int n = 0;
    void own(DevBuf &b) { bufs[n++] = &b; }
    ~DevBufScope()
    {
        for (int i = 0; i < n; i++) bufs[i]->release();
    }
};

struct ProfSlot {
    int64_t launches = 0;
    double ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<int> tags;           // stream of each pending bracket (1 = main, 2, 3, 4), for SRGP_TIMELINE
};

}  // namespace srgp

struct srgp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // main stream: passes, dense chain
    cudaStream_t stream2 = nullptr;  // side stream: work independent of the main chain
    cudaStream_t stream3 = nullptr;  // generator stream: K chunk c+1 is generated while the DMMA kernel eats chunk c
    cudaStream_t stream4 = nullptr;  // lowest priority: pass 2's K generated ahead, under the m x m stage (gauss_pregen_k2)
    cudaEvent_t ev_gen[4] = {}, ev_used[4] = {};   // per chunk buffer: generated / consumed (pass 1 of the INT8 engine uses 4)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_aux = nullptr;    // side stream -> main: an intermediate result is ready (gauss_vi.cu: S before its inverse)
    cudaEvent_t tim0 = nullptr, tim1 = nullptr;
    cudaEvent_t tl_ref = nullptr;    // SRGP_TIMELINE: reference event of the last srgp_prof_reset

    // API-parity scratch (K1/K2/K5 with host pointers)
    srgp::DevBuf in_x, in_xp, out_mat, tmp0, tmp1;
    srgp::DevBuf flush;

    // resident data shard
    srgp::DevBuf X, y, mu;  // X: n x d column-major, y/mu: n
    const double *Xp = nullptr, *yp = nullptr, *mup = nullptr;
    int64_t n = 0;
    int d = 0;
    bool have_data = false;
    uint64_t data_version = 0;   // bumped whenever the rows behind Xp may have changed (keys the resident K image, gauss_i8.cu)

    // workspace of the fused pipeline (owned by gauss.cu / laplace.cu)
    void *ws = nullptr;
    void (*ws_free)(void *) = nullptr;

    // communicator (comm.cpp)
    void *comm = nullptr;
    int world = 1, rank = 0;

    // accounting
    bool prof_on = false;
    srgp::ProfSlot prof[SRGP_PROF_COUNT];
    std::vector<cudaEvent_t> ev_pool;
    int64_t launches = 0;
};

namespace srgp {

// RAII bracket: counts the launch and, when profiling is on, times it with CUDA events on `s`.
struct KernelScope {
    srgp_ctx *ctx;
    int id;
    cudaStream_t s;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    KernelScope(srgp_ctx *c, int id_, cudaStream_t s_, int n_launches = 1) : ctx(c), id(id_), s(s_)
    {
        ctx->launches += n_launches;
        ctx->prof[id].launches += n_launches;
        if (ctx->prof_on) {
            e0 = take();
            e1 = take();
            cudaEventRecord(e0, s);
        }
    }
    ~KernelScope()
    {
        if (e0) {
            cudaEventRecord(e1, s);
            ctx->prof[id].pending.emplace_back(e0, e1);
            ctx->prof[id].tags.push_back(s == ctx->stream ? 1 : s == ctx->stream2 ? 2 : s == ctx->stream3 ? 3 : 4);
        }
    }
    cudaEvent_t take()
    {
        if (!ctx->ev_pool.empty()) {
            cudaEvent_t e = ctx->ev_pool.back();
            ctx->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

// cudaFuncSetAttribute applies to the CURRENT device only, and one process may hold contexts on several GPUs:
// a call site keeps one of these (static) and configures its kernels once per device.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool need(int device)
    {
        if (device < 0 || device >= 64) return true;
        const unsigned long long bit = 1ull << device;
        return (mask.fetch_or(bit) & bit) == 0;
    }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int use_device(srgp_ctx *ctx);

}  // namespace srgp
