// exchange.cu -- the path's one exchange step: variable-length all-gather of the M_r x 7 point rows
// over an NCCL communicator owned by the host (one rank per GPU).  The reference is single-process and simply
// appends every main frame's rows (recon.cpp:115-116); with main frames sharded in contiguous blocks, the rows of
// all ranks concatenated in rank order reproduce that append order.
//
// NCCL is resolved at run time -- from the symbols the host process already has (it created the communicator, so
// its NCCL is the one that must be called), else from libnccl.so.2 -- so libmeshrecon_b200.so itself has no link
// dependency on NCCL and loads on single-GPU boxes without it.
#include <dlfcn.h>

#include <vector>

#include "common.cuh"

namespace {

// stable NCCL ABI (nccl.h): opaque communicator, int result (0 = ncclSuccess), datatype enum values
typedef void *nccl_comm_t;
typedef int nccl_result_t;
enum { NCCL_INT32 = 2, NCCL_FLOAT32 = 7 };

struct NcclApi {
    nccl_result_t (*CommCount)(nccl_comm_t, int *) = nullptr;
    nccl_result_t (*CommUserRank)(nccl_comm_t, int *) = nullptr;
    nccl_result_t (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = RTLD_DEFAULT;
    if (!dlsym(RTLD_DEFAULT, "ncclAllGather")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by the host (e.g. torch's copy)?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            api.why = "NCCL not found: no ncclAllGather in the process and libnccl.so.2 cannot be loaded";
            return api;
        }
    }
    auto sym = [&](const char *n) { return dlsym(h, n); };
    api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.CommCount && api.CommUserRank && api.AllGather && api.Broadcast && api.GroupStart && api.GroupEnd;
    if (!api.ok) api.why = "NCCL library lacks a required symbol";
    return api;
}

}  // namespace

#define MR_NCCL(ctx, api, call)                                                                                  \
    do {                                                                                                         \
        nccl_result_t r__ = (call);                                                                              \
        if (r__ != 0) return mr_fail(ctx, MR_ECUDA, #call, (api).GetErrorString ? (api).GetErrorString(r__) : "NCCL error"); \
    } while (0)

extern "C" int mr_allgather_points(mr_context *ctx, void *nccl_comm, const float *rows, int count, float *out_rows,
                                   size_t out_capacity_rows, int *out_counts, int *out_total)
{
    if (!ctx) return MR_EINVAL;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nccl_comm || count < 0 || (count > 0 && !rows) || !out_rows) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "bad argument");
    if (!mr_is_device_ptr(out_rows) || (count > 0 && !mr_is_device_ptr(rows)))
        return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "rows and out_rows must be device memory");
    NcclApi &api = nccl_api();
    if (!api.ok) return mr_fail(ctx, MR_ENODEVICE, "mr_allgather_points", api.why.c_str());
    int world = 0, rank = -1;
    MR_NCCL(ctx, api, api.CommCount(nccl_comm, &world));
    MR_NCCL(ctx, api, api.CommUserRank(nccl_comm, &rank));
    if (world < 1 || world > 4096) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "bad communicator");
    // 1. counts: one int per rank (device all-gather, read back through the pinned scratch)
    int *d_cnt = mr_buf<int>(ctx, "xchg_counts", (size_t)world + 1);
    int *h_cnt = nullptr;
    if (!d_cnt) return mr_fail(ctx, MR_ENOMEM, "mr_allgather_points", "alloc");
    if (ctx->h_xchg_cap < world + 1) {
        if (ctx->h_xchg) cudaFreeHost(ctx->h_xchg);
        ctx->h_xchg = nullptr;
        MR_CUDA(ctx, cudaMallocHost(&ctx->h_xchg, sizeof(int) * (size_t)(world + 1)));
        ctx->h_xchg_cap = world + 1;
    }
    h_cnt = ctx->h_xchg;
    h_cnt[world] = count;
    MR_CUDA(ctx, cudaMemcpyAsync(d_cnt + world, h_cnt + world, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    MR_NCCL(ctx, api, api.AllGather(d_cnt + world, d_cnt, 1, NCCL_INT32, nccl_comm, ctx->stream));
    MR_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t total = 0;
    std::vector<size_t> off((size_t)world);
    for (int r = 0; r < world; r++) {
        if (h_cnt[r] < 0) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "negative count received");
        off[r] = total;
        total += (size_t)h_cnt[r];
        if (out_counts) out_counts[r] = h_cnt[r];
    }
    if (out_total) *out_total = (int)total;
    if (total > out_capacity_rows) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "out_rows is too small for the gathered rows");
    // 2. rows: one broadcast per rank with its exact count, grouped into a single NCCL operation (no padding to the
    //    largest count, rows land at their final offset: rank-order concatenation == the reference's append order)
    MR_NCCL(ctx, api, api.GroupStart());
    for (int r = 0; r < world; r++) {
        if (h_cnt[r] == 0) continue;
        float *dst = out_rows + off[r] * 7;
        const void *src = (r == rank) ? (const void *)rows : (const void *)dst;
        nccl_result_t rr = api.Broadcast(src, dst, (size_t)h_cnt[r] * 7, NCCL_FLOAT32, r, nccl_comm, ctx->stream);
        if (rr != 0) {
            api.GroupEnd();
            return mr_fail(ctx, MR_ECUDA, "ncclBroadcast", api.GetErrorString ? api.GetErrorString(rr) : "NCCL error");
        }
    }
    MR_NCCL(ctx, api, api.GroupEnd());
    return MR_OK;   // stream-ordered: out_rows is complete after mr_synchronize(ctx) (or any later work on mr_stream)
}
