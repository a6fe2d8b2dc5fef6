// comm.cpp -- C1: row sharding over ranks with one NCCL sum-allreduce per pass (m x m Gram + m-vector +
// scalars; gradient partials).  libnccl is loaded with dlopen so that single-GPU use has no NCCL dependency
// and a process that already carries torch's bundled NCCL reuses that copy (same SONAME).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return SRGP_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        srgp::set_error("cannot load libnccl.so.2: %s", dlerror());
        return SRGP_ERR_COMM;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        srgp::set_error("libnccl is missing a required symbol");
        return SRGP_ERR_COMM;
    }
    g_nccl.handle = h;
    return SRGP_OK;
}

int nccl_fail(const char *what, ncclResult_t r)
{
    srgp::set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    return SRGP_ERR_COMM;
}

}  // namespace

namespace srgp {

// In-place sum over ranks on stream `s` (no host synchronisation).  No-op for a single rank.
int comm_allreduce(srgp_ctx *ctx, double *buf, size_t count, cudaStream_t s)
{
    if (ctx->world <= 1 || !ctx->comm) return SRGP_OK;
    KernelScope ks(ctx, SRGP_PROF_COMM, s);
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, s);
    if (r != 0) return nccl_fail("ncclAllReduce", r);
    return SRGP_OK;
}

}  // namespace srgp

extern "C" int srgp_comm_unique_id(char id[SRGP_UNIQUE_ID_BYTES])
{
    if (!id) return SRGP_ERR_ARG;
    SRGP_TRY(load_nccl());
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if (r != 0) return nccl_fail("ncclGetUniqueId", r);
    static_assert(sizeof(uid) == SRGP_UNIQUE_ID_BYTES, "unique id size");
    memcpy(id, &uid, sizeof(uid));
    return SRGP_OK;
}

extern "C" int srgp_comm_init(srgp_ctx *ctx, int world, int rank, const char id[SRGP_UNIQUE_ID_BYTES])
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) {
        srgp::set_error("bad argument");
        return SRGP_ERR_ARG;
    }
    SRGP_TRY(srgp::use_device(ctx));
    if (ctx->comm) srgp_comm_destroy(ctx);
    ctx->world = world;
    ctx->rank = rank;
    if (world == 1) return SRGP_OK;
    SRGP_TRY(load_nccl());
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (r != 0) return nccl_fail("ncclCommInitRank", r);
    ctx->comm = comm;
    return SRGP_OK;
}

extern "C" int srgp_comm_destroy(srgp_ctx *ctx)
{
    if (!ctx) return SRGP_ERR_ARG;
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return SRGP_OK;
}
