// rm_comm.cu — the multi-GPU exchange step behind the C ABI: one NCCL communicator per context and the three-step
// frame reduction (include/raym0nade_b200.h, "Multi-GPU reduction").
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the one already in the process when the host is PyTorch), so the
// library loads and renders on one GPU without it; rm_comm_* fail loudly when it is missing.  One process per GPU,
// one context per process: the host hands the unique id from rank 0 to the other ranks (MPI, torch.distributed, a file).
#include <dlfcn.h>

#include "rm_context.cuh"

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI; enums as in nccl.h)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0, ncclMax = 2 };

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl &nccl() {
    static Nccl N;
    if (N.lib) return N;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        N.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (N.lib) break;
    }
    if (!N.lib) return N;
    auto sym = [&](const char *s) { return dlsym(N.lib, s); };
    N.GetUniqueId = reinterpret_cast<decltype(N.GetUniqueId)>(sym("ncclGetUniqueId"));
    N.CommInitRank = reinterpret_cast<decltype(N.CommInitRank)>(sym("ncclCommInitRank"));
    N.CommDestroy = reinterpret_cast<decltype(N.CommDestroy)>(sym("ncclCommDestroy"));
    N.AllReduce = reinterpret_cast<decltype(N.AllReduce)>(sym("ncclAllReduce"));
    N.Reduce = reinterpret_cast<decltype(N.Reduce)>(sym("ncclReduce"));
    N.ReduceScatter = reinterpret_cast<decltype(N.ReduceScatter)>(sym("ncclReduceScatter"));
    N.GroupStart = reinterpret_cast<decltype(N.GroupStart)>(sym("ncclGroupStart"));
    N.GroupEnd = reinterpret_cast<decltype(N.GroupEnd)>(sym("ncclGroupEnd"));
    N.GetErrorString = reinterpret_cast<decltype(N.GetErrorString)>(sym("ncclGetErrorString"));
    N.ok = N.GetUniqueId && N.CommInitRank && N.CommDestroy && N.AllReduce && N.Reduce && N.ReduceScatter && N.GroupStart && N.GroupEnd && N.GetErrorString;
    return N;
}

int need_nccl(const char *who) {
    if (!nccl().ok) return rm_fail(RM_ERR_STATE, "%s: NCCL is not available in this process (libnccl.so.2 could not be loaded)", who);
    return RM_OK;
}

#define RM_NCCL(call)                                                                                    \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != ncclSuccess) return rm_fail(RM_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r_)); \
    } while (0)

} // namespace

void rm_comm_state_free(RmContext *ctx) {
    if (ctx->comm && nccl().ok) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
}

extern "C" {

int rm_comm_unique_id(uint8_t id[128]) {
    if (!id) return rm_fail(RM_ERR_INVALID, "rm_comm_unique_id: id is NULL");
    int rc = need_nccl("rm_comm_unique_id");
    if (rc) return rc;
    ncclUniqueId u;
    RM_NCCL(nccl().GetUniqueId(&u));
    std::memcpy(id, u.internal, 128);
    return RM_OK;
}

int rm_comm_init(RmContext *ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    if (!ctx || !id) return rm_fail(RM_ERR_INVALID, "rm_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return rm_fail(RM_ERR_INVALID, "rm_comm_init: rank %d of %d", rank, world);
    int rc = need_nccl("rm_comm_init");
    if (rc) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    rm_comm_state_free(ctx);
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    ncclComm_t comm = nullptr;
    RM_NCCL(nccl().CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    return RM_OK;
}

int rm_comm_destroy(RmContext *ctx) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    rm_comm_state_free(ctx);
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return RM_OK;
}

// The frame reduction, enqueued on the context's stream: SUM of {sum of sample luminances, sample count} and MAX of the
// held-back luminance over all ranks -> every rank commits its held-back sample against the global totals -> SUM of the
// 16 radiance floats per pixel to rank `root`, which then resolves.
int rm_reduce(RmContext *ctx, int32_t root) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    if (!ctx->comm) return rm_fail(RM_ERR_STATE, "rm_reduce: call rm_comm_init first");
    if (root < 0 || root >= ctx->comm_world) return rm_fail(RM_ERR_INVALID, "rm_reduce: root %d of %d", root, ctx->comm_world);
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    float *d_sum = nullptr, *d_max = nullptr, *d_rad = nullptr;
    int64_t n_sum = 0, n_max = 0, n_rad = 0;
    int rc;
    if ((rc = rm_accum_view(ctx, &d_sum, &n_sum, &d_max, &n_max))) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    RM_NCCL(nccl().GroupStart());
    RM_NCCL(nccl().AllReduce(d_sum, d_sum, size_t(n_sum), ncclFloat32, ncclSum, comm, ctx->stream));
    RM_NCCL(nccl().AllReduce(d_max, d_max, size_t(n_max), ncclFloat32, ncclMax, comm, ctx->stream));
    RM_NCCL(nccl().GroupEnd());
    if ((rc = rm_accum_after_reduce(ctx, ctx->comm_rank, ctx->comm_world))) return rc;
    if ((rc = rm_accum_radiance(ctx, &d_rad, &n_rad))) return rc;
    RM_NCCL(nccl().Reduce(d_rad, d_rad, size_t(n_rad), ncclFloat32, ncclSum, root, comm, ctx->stream));
    return RM_OK;
}

// The same exchange with the last step scattered: after it rank r holds the summed radiance of ITS slice of the frame -
// pixels [r * per, (r + 1) * per), per = ceil(npix / world) - and nothing of the others'.  Every rank then finalises and
// downloads its own slice (rm_resolve_slice) over its own PCIe link, instead of rank `root` finalising and downloading
// the whole frame while the others wait.
int rm_reduce_scatter(RmContext *ctx) {
    if (!ctx) return rm_fail(RM_ERR_INVALID, "context is NULL");
    if (!ctx->comm) return rm_fail(RM_ERR_STATE, "rm_reduce_scatter: call rm_comm_init first");
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    float *d_sum = nullptr, *d_max = nullptr, *d_rad = nullptr;
    int64_t n_sum = 0, n_max = 0, n_rad = 0;
    int rc;
    if ((rc = rm_accum_view(ctx, &d_sum, &n_sum, &d_max, &n_max))) return rc;
    RM_CUDA(cudaSetDevice(ctx->device));
    RM_NCCL(nccl().GroupStart());
    RM_NCCL(nccl().AllReduce(d_sum, d_sum, size_t(n_sum), ncclFloat32, ncclSum, comm, ctx->stream));
    RM_NCCL(nccl().AllReduce(d_max, d_max, size_t(n_max), ncclFloat32, ncclMax, comm, ctx->stream));
    RM_NCCL(nccl().GroupEnd());
    if ((rc = rm_accum_after_reduce(ctx, ctx->comm_rank, ctx->comm_world))) return rc;
    if ((rc = rm_accum_radiance(ctx, &d_rad, &n_rad))) return rc;
    // in place: a rank's result lands where its slice already lies (the accumulators are padded to world * per pixels)
    const int64_t per = (n_rad / 16 + ctx->comm_world - 1) / ctx->comm_world;
    RM_NCCL(nccl().ReduceScatter(d_rad, d_rad + size_t(ctx->comm_rank) * size_t(per) * 16, size_t(per) * 16, ncclFloat32, ncclSum, comm, ctx->stream));
    return rm_accum_mark_slice(ctx, ctx->comm_rank * per, per);
}

} // extern "C"
