// exchange.cu -- the path's one exchange step: variable-length all-gather of the M_r x 7 point rows
// over an NCCL communicator owned by the host (one rank per GPU).  The reference is single-process and simply
// appends every main frame's rows (recon.cpp:115-116); with main frames sharded in contiguous blocks, the rows of
// all ranks concatenated in rank order reproduce that append order.
//
// NCCL is resolved at run time -- from the symbols the host process already has (it created the communicator, so
// its NCCL is the one that must be called), else from libnccl.so.2 -- so libmeshrecon_b200.so itself has no link
// dependency on NCCL and loads on single-GPU boxes without it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include <vector>

#include "common.cuh"

namespace {

// stable NCCL ABI (nccl.h): opaque communicator, int result (0 = ncclSuccess), datatype enum values
typedef void *nccl_comm_t;
typedef int nccl_result_t;
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT32 = 7 };

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

void nccl_api_init(NcclApi &api);

NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;          // contexts may be driven from different host threads
    std::call_once(once, []() { nccl_api_init(api); });
    return api;
}

void nccl_api_init(NcclApi &api)
{
    void *h = RTLD_DEFAULT;
    if (!dlsym(RTLD_DEFAULT, "ncclAllGather")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by the host (e.g. torch's copy)?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            api.why = "NCCL not found: no ncclAllGather in the process and libnccl.so.2 cannot be loaded";
            return;
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
}

}  // namespace

#define MR_NCCL(ctx, api, call)                                                                                  \
    do {                                                                                                         \
        nccl_result_t r__ = (call);                                                                              \
        if (r__ != 0) return mr_fail(ctx, MR_ECUDA, #call, (api).GetErrorString ? (api).GetErrorString(r__) : "NCCL error"); \
    } while (0)

extern "C" int mr_allgather_points(mr_context *ctx, void *nccl_comm, const float *rows, int count, float *out_rows,
                                   size_t out_capacity_rows, int *out_counts, long long *out_total)
{
    if (!ctx) return MR_EINVAL;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!nccl_comm) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "null communicator");
    NcclApi &api = nccl_api();
    if (!api.ok) return mr_fail(ctx, MR_ENODEVICE, "mr_allgather_points", api.why.c_str());
    int world = 0, rank = -1;
    MR_NCCL(ctx, api, api.CommCount(nccl_comm, &world));
    MR_NCCL(ctx, api, api.CommUserRank(nccl_comm, &rank));
    if (world < 1 || world > 4096) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "bad communicator");
    // A rank whose own arguments are bad still takes part in the first collective (with count = -1) so that EVERY rank
    // sees the failure and none is left waiting inside the row broadcasts.
    const bool local_bad = count < 0 || (count > 0 && !rows) || !out_rows || !mr_is_device_ptr(out_rows) || (count > 0 && !mr_is_device_ptr(rows));
    // 1. (count, capacity) of every rank: two int64 per rank (device all-gather, read back through the pinned scratch)
    long long *d_cnt = mr_buf<long long>(ctx, "xchg_counts", 2 * ((size_t)world + 1));
    if (!d_cnt) return mr_fail(ctx, MR_ENOMEM, "mr_allgather_points", "alloc");
    const int need = 4 * (world + 1);                       // pinned scratch is counted in ints
    if (ctx->h_xchg_cap < need) {
        if (ctx->h_xchg) cudaFreeHost(ctx->h_xchg);
        ctx->h_xchg = nullptr;
        ctx->h_xchg_cap = 0;
        MR_CUDA(ctx, cudaMallocHost(&ctx->h_xchg, sizeof(int) * (size_t)need));
        ctx->h_xchg_cap = need;
    }
    long long *h_cnt = reinterpret_cast<long long *>(ctx->h_xchg);
    h_cnt[2 * world] = local_bad ? -1 : (long long)count;
    h_cnt[2 * world + 1] = (long long)out_capacity_rows;
    MR_CUDA(ctx, cudaMemcpyAsync(d_cnt + 2 * world, h_cnt + 2 * world, 2 * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    MR_NCCL(ctx, api, api.AllGather(d_cnt + 2 * world, d_cnt, 2, NCCL_INT64, nccl_comm, ctx->stream));
    MR_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 2 * sizeof(long long) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t total = 0;
    long long min_cap = -1;
    bool any_bad = false;
    std::vector<size_t> off((size_t)world);
    for (int r = 0; r < world; r++) {
        const long long c = h_cnt[2 * r], cap = h_cnt[2 * r + 1];
        if (c < 0 || c > 0x7fffffffLL) any_bad = true;
        off[r] = total;
        if (c > 0) total += (size_t)c;
        if (min_cap < 0 || cap < min_cap) min_cap = cap;
        if (out_counts) out_counts[r] = (int)c;
    }
    if (out_total) *out_total = (long long)total;
    // every rank evaluates the same two conditions on the same gathered data: the abort is collective
    if (local_bad) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "bad argument (rows / out_rows must be device memory, count >= 0)");
    if (any_bad) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "a rank reported a bad argument; no rows were exchanged");
    if ((long long)total > min_cap) return mr_fail(ctx, MR_EINVAL, "mr_allgather_points", "out_rows of some rank is too small for the gathered rows; no rows were exchanged");
    // 2. rows: one broadcast per rank with its exact count, grouped into a single NCCL operation (no padding to the
    //    largest count, rows land at their final offset: rank-order concatenation == the reference's append order)
    MR_NCCL(ctx, api, api.GroupStart());
    for (int r = 0; r < world; r++) {
        const size_t c = (size_t)h_cnt[2 * r];
        if (c == 0) continue;
        float *dst = out_rows + off[r] * 7;
        const void *src = (r == rank) ? (const void *)rows : (const void *)dst;
        nccl_result_t rr = api.Broadcast(src, dst, c * 7, NCCL_FLOAT32, r, nccl_comm, ctx->stream);
        if (rr != 0) {
            api.GroupEnd();
            return mr_fail(ctx, MR_ECUDA, "ncclBroadcast", api.GetErrorString ? api.GetErrorString(rr) : "NCCL error");
        }
    }
    MR_NCCL(ctx, api, api.GroupEnd());
    return MR_OK;   // stream-ordered: out_rows is complete after mr_synchronize(ctx) (or any later work on mr_stream)
}

// ---------------------------------------------------------------------------------------------------------------
// Peer-memory exchange (one process per GPU on an NVLink / NVSwitch node): every rank owns a receive buffer with one
// slot per rank, exports it over CUDA IPC, and PUSHES its rows into its slot of every peer's buffer with copy-engine
// DMAs over NVLink -- no SMs, so the exchange cannot take SMs (or whole-SM CTAs) away from the path's kernels the way
// a kernel-based collective does, and the rows never pass through a send buffer (the normals kernel writes them into
// the rank's own slot).  Measured: 740-755 GB/s per GPU on idle GPUs with one copy stream
// (NCCL all-gather: 443-584 GB/s); weak scaling 98.7 % at 2 GPUs, 90.5 % at 4; at 8 GPUs the 7 x 464 MB per step
// exceed what the copy engines get while the path's kernels run (~280 GB/s) and the exchange sets the step time,
// as it does with NCCL.  Completion is signalled by the host's own barrier, entered stream-ordered after
// mr_xchg_stream (e.g. a 4-byte all-reduce): once it completes everywhere, every slot has landed.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mr_xchg_alloc(mr_context *ctx, size_t bytes, void **dev_ptr, unsigned char ipc_handle[64])
{
    if (!ctx) return MR_EINVAL;
    if (!dev_ptr || !ipc_handle || bytes == 0) return mr_fail(ctx, MR_EINVAL, "mr_xchg_alloc", "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return mr_fail(ctx, MR_ENOMEM, "mr_xchg_alloc", "cudaMalloc failed");
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return mr_fail(ctx, MR_ECUDA, "cudaIpcGetMemHandle", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, 64);
    *dev_ptr = p;
    return MR_OK;
}

extern "C" int mr_xchg_free(mr_context *ctx, void *dev_ptr)
{
    if (!ctx) return MR_EINVAL;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    MR_CUDA(ctx, cudaDeviceSynchronize());
    if (dev_ptr) MR_CUDA(ctx, cudaFree(dev_ptr));
    return MR_OK;
}

extern "C" int mr_xchg_open(mr_context *ctx, const unsigned char ipc_handle[64], void **peer_ptr)
{
    if (!ctx) return MR_EINVAL;
    if (!ipc_handle || !peer_ptr) return mr_fail(ctx, MR_EINVAL, "mr_xchg_open", "bad argument");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    MR_CUDA(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));   // maps the peer's buffer, enables P2P
    return MR_OK;
}

extern "C" int mr_xchg_close(mr_context *ctx, void *peer_ptr)
{
    if (!ctx) return MR_EINVAL;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < mr_context::N_PUSH; i++)
        if (ctx->push_stream[i]) MR_CUDA(ctx, cudaStreamSynchronize(ctx->push_stream[i]));
    if (peer_ptr) MR_CUDA(ctx, cudaIpcCloseMemHandle(peer_ptr));
    return MR_OK;
}

static int ensure_push_streams(mr_context *ctx)
{
    if (!ctx->push_stream[0]) {
        int lo = 0, hi = 0;
        MR_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < mr_context::N_PUSH; i++) {
            MR_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->push_stream[i], cudaStreamNonBlocking, hi));   // ahead of the compute streams
            MR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_push[i], cudaEventDisableTiming));
        }
    }
    return MR_OK;
}

// Stream on which to order the completion signal: everything enqueued on it after this call runs after every push
// issued so far (the pushes themselves are spread over several streams so that several copy engines work at once).
extern "C" void *mr_xchg_stream(mr_context *ctx)
{
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess || ensure_push_streams(ctx) != MR_OK) return nullptr;
    for (int i = 1; i < mr_context::N_PUSH; i++) {
        if (cudaEventRecord(ctx->ev_push[i], ctx->push_stream[i]) != cudaSuccess) return nullptr;
        if (cudaStreamWaitEvent(ctx->push_stream[0], ctx->ev_push[i], 0) != cudaSuccess) return nullptr;
    }
    return (void *)ctx->push_stream[0];
}

// DMA `bytes` from src (this GPU) to dst (normally a peer's buffer opened with mr_xchg_open) on one of the context's
// push streams, ordered after everything queued so far on mr_stream(ctx).
extern "C" int mr_xchg_push(mr_context *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return MR_EINVAL;
    if (!dst || !src) return mr_fail(ctx, MR_EINVAL, "mr_xchg_push", "bad argument");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_push_streams(ctx);
    if (rc) return rc;
    static const int n_streams = []() {                      // MR_XCHG_STREAMS=1..4 (tuning knob; default 4)
        const char *e = getenv("MR_XCHG_STREAMS");
        int n = e ? atoi(e) : mr_context::N_PUSH;
        return n < 1 ? 1 : (n > mr_context::N_PUSH ? mr_context::N_PUSH : n);
    }();
    const int i = ctx->push_next % n_streams;
    ctx->push_next = (i + 1) % n_streams;
    MR_CUDA(ctx, cudaEventRecord(ctx->ev_push[0], ctx->stream));
    MR_CUDA(ctx, cudaStreamWaitEvent(ctx->push_stream[i], ctx->ev_push[0], 0));
    if (bytes) MR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->push_stream[i]));
    return MR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Multicast push (NVSwitch): `mc_dst` is the multicast mapping of the ranks' receive buffers (cuMulticastBindMem; one
// store to it is replicated by the switch into the same offset of EVERY rank's buffer), so a rank sends its rows ONCE
// instead of world - 1 times.  The copy engines reach such an address at 330 GB/s on idle GPUs but only ~35 GB/s
// while the path's kernels run (profiles/mcast_push_r2.txt), so the write is done by a few CTAs of 128-bit multimem.st
// instead, on the high-priority push streams: they take the first SM slots that come free and leave them within
// microseconds.  Ordered after everything queued on mr_stream(ctx); completion is signalled like mr_xchg_push's.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mcast_store_kernel(const uint4 *__restrict__ src, uint4 *mc_dst, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int U = 4;                                   // four 16-byte loads in flight per thread
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n16; i0 += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (i0 + u * stride < n16) v[u] = __ldcs(src + i0 + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (i0 + u * stride < n16)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i0 + u * stride),
                             "f"(__uint_as_float(v[u].x)), "f"(__uint_as_float(v[u].y)), "f"(__uint_as_float(v[u].z)), "f"(__uint_as_float(v[u].w))
                             : "memory");
    }
    __threadfence_system();
}

extern "C" int mr_xchg_push_mcast(mr_context *ctx, void *mc_dst, const void *src, size_t bytes)
{
    if (!ctx) return MR_EINVAL;
    if (!mc_dst || !src || (bytes & 15) || ((uintptr_t)mc_dst & 15) || ((uintptr_t)src & 15))
        return mr_fail(ctx, MR_EINVAL, "mr_xchg_push_mcast", "null or not 16-byte aligned");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_push_streams(ctx);
    if (rc) return rc;
    static const int n_ctas = []() {                         // MR_MCAST_CTAS (tuning knob)
        const char *e = getenv("MR_MCAST_CTAS");
        int n = e ? atoi(e) : 32;
        return n < 1 ? 1 : (n > 1184 ? 1184 : n);
    }();
    MR_CUDA(ctx, cudaEventRecord(ctx->ev_push[0], ctx->stream));
    MR_CUDA(ctx, cudaStreamWaitEvent(ctx->push_stream[0], ctx->ev_push[0], 0));
    if (bytes) {
        mcast_store_kernel<<<n_ctas, 256, 0, ctx->push_stream[0]>>>((const uint4 *)src, (uint4 *)mc_dst, bytes / 16);
        MR_LAUNCH_CHECK(ctx, "mcast_store_kernel");
    }
    return MR_OK;
}
