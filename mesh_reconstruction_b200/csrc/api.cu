// api.cu -- extern "C" entry points of libmeshrecon_b200.so (include/meshrecon_b200.h).
// Host-side plumbing only: argument checks, host<->device staging, stage sequencing.
// There is no CPU implementation of any stage behind these calls.
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "jacobi3.cuh"

thread_local std::string g_mr_create_error;
// process-wide debug knobs (mr_set_vr_impl): atomics, so that contexts driven from different threads never race on them
std::atomic<int> g_mr_vr_impl{1};
extern std::atomic<int> g_mr_vr_tma;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel, shared by every context of this
// process on that device: one registry per (device, kernel), only ever raised, guarded by a mutex because contexts are
// independent and may be driven by one host thread each (SURVEY 8b "re-entrancy across different ctx").
cudaError_t mr_ensure_smem_raw(int device, const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> reg;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = reg[std::make_pair(device, kernel)];
    if (bytes <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

int mr_fail(mr_context *ctx, int code, const char *what, const char *detail)
{
    std::string msg = std::string(what ? what : "") + ": " + (detail ? detail : "");
    if (ctx) ctx->err = msg;
    else g_mr_create_error = msg;
    return code;
}

void *mr_buf_raw(mr_context *ctx, const char *name, size_t bytes)
{
    DevBuf &b = ctx->bufs[name];
    if (b.p && bytes <= b.bytes) return b.p;
    if (bytes == 0) return b.p;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t alloc = (bytes + 255) & ~(size_t)255;
    if (cudaMalloc(&b.p, alloc) != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return nullptr;
    }
    b.bytes = alloc;
    return b.p;
}

bool mr_is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

const void *mr_in(mr_context *ctx, const void *p, size_t bytes, const char *staging)
{
    if (mr_is_device_ptr(p)) return p;
    void *d = mr_buf_raw(ctx, staging, bytes);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return nullptr;
    return d;
}

int mr_out(mr_context *ctx, void *dst, const void *src_dev, size_t bytes)
{
    if (dst == src_dev) return MR_OK;
    MR_CUDA(ctx, cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDefault, ctx->stream));
    return MR_OK;
}

static Mat4 to_mat4(const float *m)
{
    Mat4 r;
    memcpy(r.m, m, sizeof(r.m));
    return r;
}

#define CHECK_CTX(ctx) \
    if (!(ctx)) return MR_EINVAL
#define CHECK_ARG(ctx, cond, msg) \
    if (!(cond)) return mr_fail(ctx, MR_EINVAL, __func__, msg)
#define SET_DEVICE(ctx) MR_CUDA(ctx, cudaSetDevice((ctx)->device))
#define RC(x)                 \
    do {                      \
        int rc__ = (x);       \
        if (rc__) return rc__; \
    } while (0)

extern "C" {

int mr_version(void) { return 100; }

int mr_create(mr_context **out, int device, int width, int height)
{
    if (!out) return MR_EINVAL;
    *out = nullptr;
    if (width <= 0 || height <= 0 || (size_t)width * height > ((size_t)1 << 30)) return mr_fail(nullptr, MR_EINVAL, "mr_create", "bad size");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return mr_fail(nullptr, MR_ENODEVICE, "mr_create", "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return mr_fail(nullptr, MR_ENODEVICE, "mr_create", "device index out of range");
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return mr_fail(nullptr, MR_ECUDA, "mr_create", "cudaSetDevice failed");
    if (prop.major != 10) return mr_fail(nullptr, MR_ENODEVICE, "mr_create", "device is not sm_100 (B200); kernels are built for sm_100a only");
    mr_context *ctx = new mr_context();
    ctx->device = device;
    ctx->W = width;
    ctx->H = height;
    ctx->N = (size_t)width * height;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->h_count, 64) != cudaSuccess) {
        delete ctx;
        return mr_fail(nullptr, MR_ECUDA, "mr_create", "stream / pinned alloc failed");
    }
    // pyramid geometry of compare() (util.cpp:335-351)
    int size = width < height ? width : height, w = width, h = height, L = 0;
    size_t off = 0;
    for (;;) {
        ctx->lw[L] = w; ctx->lh[L] = h; ctx->loff[L] = off;
        off += (size_t)w * h;
        L++;
        if (size <= 2 || L >= MR_MAX_LEVELS) break;
        w = (w + 1) / 2; h = (h + 1) / 2; size /= 2;
    }
    ctx->n_levels = L;
    ctx->pyr_total = off;
    int rc = mr_flow_init_tables(ctx);
    if (rc) {
        g_mr_create_error = ctx->err;
        mr_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return MR_OK;
}

void mr_destroy(mr_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->bufs)
        if (kv.second.p) cudaFree(kv.second.p);
    if (ctx->h_count) cudaFreeHost(ctx->h_count);
    if (ctx->h_xchg) cudaFreeHost(ctx->h_xchg);
    for (int i = 0; i < mr_context::N_PUSH; i++) {
        if (ctx->push_stream[i]) { cudaStreamSynchronize(ctx->push_stream[i]); cudaStreamDestroy(ctx->push_stream[i]); }
        if (ctx->ev_push[i]) cudaEventDestroy(ctx->ev_push[i]);
    }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    for (int i = 0; i < 2; i++) if (ctx->ev_copy_done[i]) cudaEventDestroy(ctx->ev_copy_done[i]);
    for (int i = 0; i < 2; i++) if (ctx->ev_rows_done[i]) cudaEventDestroy(ctx->ev_rows_done[i]);
    for (auto e : ctx->ev_copy_ring) if (e) cudaEventDestroy(e);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    for (auto &r : ctx->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *mr_last_error(const mr_context *ctx) { return ctx ? ctx->err.c_str() : g_mr_create_error.c_str(); }
void *mr_stream(mr_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t mr_launch_count(const mr_context *ctx) { return ctx ? ctx->launches : 0; }

int mr_synchronize(mr_context *ctx)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) MR_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending[0] = ctx->copy_pending[1] = false;
    return MR_OK;
}

int mr_wait_copies(mr_context *ctx)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    if (ctx->copy_stream) MR_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending[0] = ctx->copy_pending[1] = false;
    return MR_OK;
}

int mr_wait_copies_until(mr_context *ctx, int max_in_flight)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, max_in_flight >= 0, "max_in_flight must be >= 0");
    if (!ctx->copy_stream || ctx->copy_seq <= (unsigned long long)max_in_flight) return MR_OK;
    // copies complete in submission order: waiting for number `last` covers every earlier one.  A target that has
    // already left the ring is covered by the oldest event still in it.
    unsigned long long last = ctx->copy_seq - 1 - (unsigned long long)max_in_flight;
    if (ctx->copy_seq - last > (unsigned long long)mr_context::COPY_RING) last = ctx->copy_seq - mr_context::COPY_RING;
    MR_CUDA(ctx, cudaEventSynchronize(ctx->ev_copy_ring[last % mr_context::COPY_RING]));
    return MR_OK;
}

// Configuration::useFarneback (configuration.cpp:26, -f) for the fused main-frame call
int mr_set_use_farneback(mr_context *ctx, int on)
{
    CHECK_CTX(ctx);
    ctx->use_farneback = on != 0;
    return MR_OK;
}

// debug / benchmarking knob: 0 = plane-per-stage VR kernels, 1 = fused tile kernel with TMA staging
// (default), 2 = fused tile kernel with plain loads
int mr_set_vr_impl(int impl)
{
    g_mr_vr_impl.store(impl ? 1 : 0);
    g_mr_vr_tma.store(impl == 2 ? 0 : 1);
    return MR_OK;
}

int mr_set_use_graphs(mr_context *ctx, int on)
{
    CHECK_CTX(ctx);
    CHECK_ARG(ctx, on >= 0 && on <= 2, "mode must be 0 (never), 1 (host rows) or 2 (always)");
    ctx->graphs_mode = on;
    return MR_OK;
}

uint64_t mr_graph_launch_count(const mr_context *ctx) { return ctx ? ctx->graph_launches : 0; }

int mr_load_mesh(mr_context *ctx, const float *vertices_xyzw, int n_vertices, const int32_t *faces, int n_faces)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    ctx->graph_warm = false;   // per-face buffers may grow: the next submission runs plainly (and allocates) again
    CHECK_ARG(ctx, n_vertices >= 0 && n_faces >= 0, "negative count");
    CHECK_ARG(ctx, n_faces == 0 || (vertices_xyzw && faces), "null mesh pointers");
    ctx->mesh_loaded = true;
    if (n_faces == 0) return k_load_mesh(ctx, nullptr, 0, nullptr, 0, nullptr);
    CHECK_ARG(ctx, n_vertices > 0, "faces without vertices");
    const float *dv = (const float *)mr_in(ctx, vertices_xyzw, (size_t)n_vertices * 4 * sizeof(float), "in_vtx");
    const int32_t *df = (const int32_t *)mr_in(ctx, faces, (size_t)n_faces * 3 * sizeof(int32_t), "in_faces");
    int *d_bad = mr_buf<int>(ctx, "mesh_bad", 1);
    if (!dv || !df || !d_bad) return mr_fail(ctx, MR_ENOMEM, "mr_load_mesh", "staging");
    MR_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    RC(k_load_mesh(ctx, dv, n_vertices, df, n_faces, d_bad));
    MR_CUDA(ctx, cudaMemcpyAsync(ctx->h_count + 8, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_count[8]) {
        // the offending faces were neutralised on the device (NaN corners never rasterise); unload the mesh so that a
        // caller that ignores the error gets MR_ENOMESH instead of a silently different scene
        ctx->F = 0;
        ctx->mesh_loaded = false;
        return mr_fail(ctx, MR_EINVAL, "mr_load_mesh", "a face references a vertex index outside [0, n_vertices)");
    }
    return MR_OK;
}

int mr_depth(mr_context *ctx, const float camera[16], float *out_depth)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, camera && out_depth, "null argument");
    if (!ctx->mesh_loaded) return mr_fail(ctx, MR_ENOMESH, "mr_depth", "loadMesh has not been called");
    unsigned long long *vis = mr_buf<unsigned long long>(ctx, "vis_main", ctx->N);
    bool dev_out = mr_is_device_ptr(out_depth);
    float *d = dev_out ? out_depth : mr_buf<float>(ctx, "depth", ctx->N);
    if (!vis || !d) return mr_fail(ctx, MR_ENOMEM, "mr_depth", "alloc");
    RC(k_raster(ctx, to_mat4(camera), vis));
    RC(k_resolve_depth(ctx, vis, d));
    if (!dev_out) {
        RC(mr_out(ctx, out_depth, d, ctx->N * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

#define MR_QUERY_MAX_PER_CAMERA 4096
int mr_depth_samples(mr_context *ctx, const float *cameras, int n_cameras, const int32_t *rows, const int32_t *cols, int n_per_camera,
                     float *out)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, cameras && rows && cols && out, "null argument");
    CHECK_ARG(ctx, n_cameras >= 0 && n_per_camera >= 0, "negative count");
    if (!ctx->mesh_loaded) return mr_fail(ctx, MR_ENOMESH, "mr_depth_samples", "loadMesh has not been called");
    size_t total = (size_t)n_cameras * n_per_camera;
    if (total == 0) return MR_OK;
    unsigned long long *vis = mr_buf<unsigned long long>(ctx, "vis_main", ctx->N);
    const int32_t *d_rows = (const int32_t *)mr_in(ctx, rows, total * sizeof(int32_t), "q_rows");
    const int32_t *d_cols = (const int32_t *)mr_in(ctx, cols, total * sizeof(int32_t), "q_cols");
    bool dev_out = mr_is_device_ptr(out);
    float *d_out = dev_out ? out : mr_buf<float>(ctx, "q_out", total);
    if (!vis || !d_rows || !d_cols || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_depth_samples", "alloc");
    if (n_per_camera <= MR_QUERY_MAX_PER_CAMERA) {
        // ray queries: the rasteriser's pixel test on the n query pixels of every viewer, no maps rendered
        const float *d_cams = (const float *)mr_in(ctx, cameras, (size_t)n_cameras * 16 * sizeof(float), "q_cams");
        if (!d_cams) return mr_fail(ctx, MR_ENOMEM, "mr_depth_samples", "alloc");
        RC(k_depth_query(ctx, d_cams, n_cameras, d_rows, d_cols, n_per_camera, d_out));
    } else {
        std::vector<float> h_cams;                 // the raster launches take the matrix by value: it has to be readable here
        const float *cams_host = cameras;
        if (mr_is_device_ptr(cameras)) {
            h_cams.resize((size_t)n_cameras * 16);
            MR_CUDA(ctx, cudaMemcpy(h_cams.data(), cameras, h_cams.size() * sizeof(float), cudaMemcpyDeviceToHost));
            cams_host = h_cams.data();
        }
        for (int i = 0; i < n_cameras; i++) {      // many queries per viewer: render each map, index it on the device
            RC(k_raster(ctx, to_mat4(cams_host + 16 * i), vis));
            RC(k_depth_samples(ctx, vis, d_rows + (size_t)i * n_per_camera, d_cols + (size_t)i * n_per_camera, n_per_camera,
                               d_out + (size_t)i * n_per_camera));
        }
    }
    if (!dev_out) {
        RC(mr_out(ctx, out, d_out, total * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

// shadow pass + dilation for `projector`, leaves the dilated map in "shadow_dil"
static int shadow_pass(mr_context *ctx, const float *projector, float **out)
{
    unsigned long long *vis = mr_buf<unsigned long long>(ctx, "vis_side", ctx->N);
    float *sh = mr_buf<float>(ctx, "shadow", ctx->N);
    float *shd = mr_buf<float>(ctx, "shadow_dil", ctx->N);
    if (!vis || !sh || !shd) return mr_fail(ctx, MR_ENOMEM, "shadow_pass", "alloc");
    {
        StageScope sc(ctx, ST_RASTER);
        RC(k_raster(ctx, to_mat4(projector), vis));
        RC(k_resolve_depth(ctx, vis, sh));
        RC(k_dilate_shadow(ctx, sh, shd));
    }
    *out = shd;
    return MR_OK;
}

int mr_projected(mr_context *ctx, const float camera[16], const uint8_t *frame, const float projector[16], uint8_t *out_rgb)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, camera && frame && projector && out_rgb, "null argument");
    if (!ctx->mesh_loaded) return mr_fail(ctx, MR_ENOMESH, "mr_projected", "loadMesh has not been called");
    const uint8_t *d_frame = (const uint8_t *)mr_in(ctx, frame, ctx->N, "in_side");
    unsigned long long *vis = mr_buf<unsigned long long>(ctx, "vis_main", ctx->N);
    bool dev_out = mr_is_device_ptr(out_rgb);
    uint8_t *rgb = dev_out ? out_rgb : mr_buf<uint8_t>(ctx, "rgb", ctx->N * 3);
    if (!d_frame || !vis || !rgb) return mr_fail(ctx, MR_ENOMEM, "mr_projected", "alloc");
    float *shd = nullptr;
    RC(shadow_pass(ctx, projector, &shd));
    RC(k_raster(ctx, to_mat4(camera), vis));
    RC(k_shade(ctx, vis, to_mat4(camera), to_mat4(projector), d_frame, shd, rgb, nullptr, nullptr, nullptr));
    if (!dev_out) {
        RC(mr_out(ctx, out_rgb, rgb, ctx->N * 3));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_mix_background(mr_context *ctx, const uint8_t *image_rgb, const uint8_t *background, float *depth_inout, uint8_t *out_mixed)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, image_rgb && background && depth_inout && out_mixed, "null argument");
    const uint8_t *d_rgb = (const uint8_t *)mr_in(ctx, image_rgb, ctx->N * 3, "in_rgb");
    const uint8_t *d_bg = (const uint8_t *)mr_in(ctx, background, ctx->N, "in_main");
    bool dev_depth = mr_is_device_ptr(depth_inout), dev_out = mr_is_device_ptr(out_mixed);
    float *d_depth = dev_depth ? depth_inout : (float *)mr_in(ctx, depth_inout, ctx->N * sizeof(float), "depth");
    uint8_t *d_out = dev_out ? out_mixed : mr_buf<uint8_t>(ctx, "mixed0", ctx->N);
    if (!d_rgb || !d_bg || !d_depth || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_mix_background", "alloc");
    RC(k_mix_background(ctx, d_rgb, d_bg, d_depth, d_out));
    if (!dev_depth) RC(mr_out(ctx, depth_inout, d_depth, ctx->N * sizeof(float)));
    if (!dev_out) RC(mr_out(ctx, out_mixed, d_out, ctx->N));
    if (!dev_depth || !dev_out) MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MR_OK;
}

// device-side calculateFlow: VR -> remap -> compare -> pack (flow.cpp:29-40)
static int calculate_flow_dev(mr_context *ctx, const uint8_t *d_prev, const uint8_t *d_next, float *d_flow4, bool farneback)
{
    uint8_t *remapped = mr_buf<uint8_t>(ctx, "remapped", ctx->N);
    if (!remapped) return mr_fail(ctx, MR_ENOMEM, "calculate_flow", "alloc");
    {
        StageScope sc(ctx, ST_VR);
        if (farneback) RC(k_farneback(ctx, d_prev, d_next, d_flow4));                // flow.cpp:22-26
        else RC(k_variational_refinement(ctx, d_prev, d_next, d_flow4));               // flow.cpp:29; both write (u, v, 0, 0)
    }
    {
        StageScope sc(ctx, ST_REMAP);
        RC(k_flow_remap(ctx, d_flow4, 4, d_next, remapped));
    }
    {
        StageScope sc(ctx, ST_COMPARE);
        RC(k_compare(ctx, d_prev, remapped, d_flow4, 4, 2));         // variance -> channel 2
    }
    return MR_OK;
}

int mr_calculate_flow(mr_context *ctx, const uint8_t *prev, const uint8_t *next, int use_farneback, float *out_flow4)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, prev && next && out_flow4, "null argument");
    const uint8_t *d_prev = (const uint8_t *)mr_in(ctx, prev, ctx->N, "in_main");
    const uint8_t *d_next = (const uint8_t *)mr_in(ctx, next, ctx->N, "in_side");
    bool dev_out = mr_is_device_ptr(out_flow4);
    float *d_flow = dev_out ? out_flow4 : mr_buf<float>(ctx, "flow0", ctx->N * 4);
    if (!d_prev || !d_next || !d_flow) return mr_fail(ctx, MR_ENOMEM, "mr_calculate_flow", "alloc");
    RC(calculate_flow_dev(ctx, d_prev, d_next, d_flow, use_farneback != 0));
    if (!dev_out) {
        RC(mr_out(ctx, out_flow4, d_flow, ctx->N * 4 * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_flow_remap(mr_context *ctx, const float *flow, int stride_floats, const uint8_t *image, uint8_t *out)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, flow && image && out, "null argument");
    CHECK_ARG(ctx, stride_floats >= 2 && stride_floats <= 4, "stride_floats must be 2..4");
    const float *d_flow = (const float *)mr_in(ctx, flow, ctx->N * stride_floats * sizeof(float), "flow0");
    const uint8_t *d_img = (const uint8_t *)mr_in(ctx, image, ctx->N, "in_side");
    bool dev_out = mr_is_device_ptr(out);
    uint8_t *d_out = dev_out ? out : mr_buf<uint8_t>(ctx, "remapped", ctx->N);
    if (!d_flow || !d_img || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_flow_remap", "alloc");
    RC(k_flow_remap(ctx, d_flow, stride_floats, d_img, d_out));
    if (!dev_out) {
        RC(mr_out(ctx, out, d_out, ctx->N));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_compare(mr_context *ctx, const uint8_t *prev, const uint8_t *next, float *out)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, prev && next && out, "null argument");
    const uint8_t *d_prev = (const uint8_t *)mr_in(ctx, prev, ctx->N, "in_main");
    const uint8_t *d_next = (const uint8_t *)mr_in(ctx, next, ctx->N, "in_side");
    bool dev_out = mr_is_device_ptr(out);
    float *d_out = dev_out ? out : mr_buf<float>(ctx, "variance", ctx->N);
    if (!d_prev || !d_next || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_compare", "alloc");
    RC(k_compare(ctx, d_prev, d_next, d_out, 1, 0));
    if (!dev_out) {
        RC(mr_out(ctx, out, d_out, ctx->N * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_image_gradient(mr_context *ctx, const float *image, float *out_grad2)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, image && out_grad2, "null argument");
    const float *d_img = (const float *)mr_in(ctx, image, ctx->N * sizeof(float), "depth");
    bool dev_out = mr_is_device_ptr(out_grad2);
    float *d_out = dev_out ? out_grad2 : mr_buf<float>(ctx, "grad2", ctx->N * 2);
    if (!d_img || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_image_gradient", "alloc");
    RC(k_image_gradient(ctx, d_img, d_out));
    if (!dev_out) {
        RC(mr_out(ctx, out_grad2, d_out, ctx->N * 2 * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

static const char *flow_name(int i)
{
    static const char *names[MR_MAX_SIDE] = {"flow0", "flow1", "flow2", "flow3", "flow4", "flow5", "flow6", "flow7",
                                              "flow8", "flow9", "flow10", "flow11", "flow12", "flow13", "flow14", "flow15"};
    return names[i];
}
static const char *mixed_name(int i)
{
    static const char *names[MR_MAX_SIDE] = {"mixed0", "mixed1", "mixed2", "mixed3", "mixed4", "mixed5", "mixed6", "mixed7",
                                              "mixed8", "mixed9", "mixed10", "mixed11", "mixed12", "mixed13", "mixed14", "mixed15"};
    return names[i];
}

int mr_triangulate_pixels(mr_context *ctx, const float *const *flows, int n_side, const float main_camera[16], const float *cameras,
                          const float *depth, float *out_points, int *out_count)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, flows && main_camera && cameras && depth && out_count, "null argument");
    CHECK_ARG(ctx, n_side >= 1 && n_side <= MR_MAX_SIDE, "n_side must be in 1..MR_MAX_SIDE");
    const float *d_flows[MR_MAX_SIDE];
    for (int i = 0; i < n_side; i++) {
        CHECK_ARG(ctx, flows[i], "null flow pointer");
        d_flows[i] = (const float *)mr_in(ctx, flows[i], ctx->N * 4 * sizeof(float), flow_name(i));
        if (!d_flows[i]) return mr_fail(ctx, MR_ENOMEM, "mr_triangulate_pixels", "staging");
    }
    const float *d_depth = (const float *)mr_in(ctx, depth, ctx->N * sizeof(float), "depth");
    bool dev_out = out_points && mr_is_device_ptr(out_points);
    float *d_out = dev_out ? out_points : mr_buf<float>(ctx, "points", ctx->N * 7);
    if (!d_depth || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_triangulate_pixels", "alloc");
    if (!dev_out && ctx->copy_pending[0]) MR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[0], 0));
    RC(k_triangulate(ctx, d_flows, n_side, main_camera, cameras, d_depth, d_out, out_count));
    ctx->last_S = n_side;
    ctx->last_rows = d_out;
    if (out_points && !dev_out && *out_count > 0) {
        RC(mr_out(ctx, out_points, d_out, (size_t)*out_count * 7 * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_extract_camera_center(const float camera[16], float out_center3[3])
{
    if (!camera || !out_center3) return MR_EINVAL;
    mr_camera_center(camera, out_center3);
    return MR_OK;
}

enum { PMF_SYNC = 0, PMF_ASYNC_COPY = 1, PMF_SUBMIT = 2 };

static int ensure_copy_stream(mr_context *ctx)
{
    if (!ctx->copy_stream) {
        // highest priority: the few CTAs of the device-side row copy must not queue behind full-grid compute kernels
        int lo = 0, hi = 0;
        MR_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MR_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 2; i++) {
            MR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_done[i], cudaEventDisableTiming));
            MR_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rows_done[i], cudaEventDisableTiming));
        }
        for (auto &e : ctx->ev_copy_ring) MR_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return MR_OK;
}

// bookkeeping for mr_wait_copies_until: the copy just queued on the copy stream is number copy_seq
static int note_copy_queued(mr_context *ctx)
{
    MR_CUDA(ctx, cudaEventRecord(ctx->ev_copy_ring[ctx->copy_seq % mr_context::COPY_RING], ctx->copy_stream));
    ctx->copy_seq++;
    return MR_OK;
}

// device-visible alias of a user pointer: device memory as is, pinned (mapped) host memory through UVA
static void *device_alias(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return const_cast<void *>(p);
    if (a.type == cudaMemoryTypeHost) return a.devicePointer;
    return nullptr;
}

// Enqueues steps recon.cpp:70-89 (depth, then projected -> mixBackground -> calculateFlow per side camera) on the
// context stream.  Device work only -- no allocation after the first call of a given shape, no synchronisation --
// so the sequence can be stream-captured into a CUDA graph by mr_submit_main_frame.
static int enqueue_flows(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                         const uint8_t *const *side_frames, const float *side_cameras, const float **d_flows, float **depth_out)
{
    size_t N = ctx->N;
    const uint8_t *d_main = (const uint8_t *)mr_in(ctx, main_frame, N, "in_main");
    unsigned long long *vis_main = mr_buf<unsigned long long>(ctx, "vis_main", N);
    float *depth = mr_buf<float>(ctx, "depth", N);
    if (!d_main || !vis_main || !depth) return mr_fail(ctx, MR_ENOMEM, "mr_process_main_frame", "alloc");
    Mat4 Pm = to_mat4(main_camera);
    {
        StageScope sc(ctx, ST_RASTER);
        RC(k_raster(ctx, Pm, vis_main));            // recon.cpp:70  depth = render->depth(camera(fa))
        RC(k_resolve_depth(ctx, vis_main, depth));
    }
    for (int i = 0; i < n_side; i++) {           // recon.cpp:81
        const uint8_t *d_side = (const uint8_t *)mr_in(ctx, side_frames[i], N, "in_side");
        float *flow = mr_buf<float>(ctx, flow_name(i), N * 4);
        uint8_t *mixed = mr_buf<uint8_t>(ctx, mixed_name(i), N);
        if (!d_side || !flow || !mixed) return mr_fail(ctx, MR_ENOMEM, "mr_process_main_frame", "alloc");
        float *shd = nullptr;
        RC(shadow_pass(ctx, side_cameras + 16 * i, &shd));                       // render_glx.cpp:272-329
        {
            StageScope sc(ctx, ST_SHADE);
            RC(k_shade(ctx, vis_main, Pm, to_mat4(side_cameras + 16 * i), d_side, shd, nullptr, d_main, depth,
                       mixed));                                                  // recon.cpp:85-86
        }
        RC(calculate_flow_dev(ctx, d_main, mixed, flow, ctx->use_farneback));                        // recon.cpp:89
        d_flows[i] = flow;
    }
    *depth_out = depth;
    return MR_OK;
}

static bool is_pinned_or_device(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeHost;
}

// mr_submit_main_frame: the whole device sequence of one main frame (~35 launches at S = 1).  After a first plain
// run of a given shape (which allocates every buffer), later submissions are stream-captured and replayed as ONE
// CUDA graph (cudaGraphExecUpdate re-targets the pointers and per-camera constants of the instantiated graph):
// the front end then fetches a single launch instead of ~35 command packets over PCIe.  That matters exactly when
// the link is busy with the previous frame's 58 MB row DMA -- measured: stream launches slow the small kernels at
// the head of a frame by up to 2x under that traffic, graph launches do not (scripts/graph_contention_micro.py,
// scripts/dma_contention_probe.py).
static int submit_enqueue(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                          const uint8_t *const *side_frames, const float *side_cameras, float *d_rows, int *d_cnt, bool host_rows)
{
    // shape of the launch sequence: anything that changes the graph's topology
    unsigned long long key = (unsigned long long)n_side | ((unsigned long long)ctx->use_farneback << 8) | ((unsigned long long)(g_mr_vr_impl.load() & 0xff) << 9) |
                             ((unsigned long long)(g_mr_vr_tma.load() & 1) << 17) | ((unsigned long long)(mr_is_device_ptr(main_frame) ? 1 : 0) << 18);
    // Graph replay is used when the rows go to the HOST (graphs_mode 1, the default) or always (2).  With rows left in
    // HBM there is no PCIe traffic to hide from, and with two contexts per GPU plain stream launches interleave the
    // two frames' kernels slightly better (measured at 1080p, 2 contexts: 1.24 vs 1.27 ms/pair resident, but
    // 1.38 vs 1.30 ms/pair with host rows).
    // The Farneback branch (-f) is ~360 short launches per pair: always worth a graph.
    bool capturable = (ctx->graphs_mode == 2 || (ctx->graphs_mode == 1 && (host_rows || ctx->use_farneback))) && !ctx->profile && is_pinned_or_device(main_frame);
    for (int i = 0; i < n_side; i++) {
        key |= (unsigned long long)(mr_is_device_ptr(side_frames[i]) ? 1 : 0) << (19 + i);
        capturable = capturable && is_pinned_or_device(side_frames[i]);
    }
    const float *d_flows[MR_MAX_SIDE];
    float *depth = nullptr;
    if (!capturable || !ctx->graph_warm || ctx->graph_key != key) {
        RC(enqueue_flows(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, d_flows, &depth));
        RC(k_triangulate(ctx, d_flows, n_side, main_camera, side_cameras, depth, d_rows, nullptr, d_cnt));
        if (capturable) {
            if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
            ctx->graph_warm = true;
            ctx->graph_key = key;
        }
        return MR_OK;
    }
    const uint64_t launches0 = ctx->launches;
    MR_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    int rc = enqueue_flows(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, d_flows, &depth);
    if (rc == MR_OK) rc = k_triangulate(ctx, d_flows, n_side, main_camera, side_cameras, depth, d_rows, nullptr, d_cnt);
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
    if (rc != MR_OK || ce != cudaSuccess || !g) {
        // capture refused something: drop graphs for this context and run the frame the plain way
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        ctx->graphs_mode = 0;
        ctx->launches = launches0;
        RC(enqueue_flows(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, d_flows, &depth));
        return k_triangulate(ctx, d_flows, n_side, main_camera, side_cameras, depth, d_rows, nullptr, d_cnt);
    }
    if (ctx->graph_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(ctx->graph_exec, g, &info) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphExecDestroy(ctx->graph_exec);
            ctx->graph_exec = nullptr;
        }
    }
    if (!ctx->graph_exec) {
        ce = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
        if (ce != cudaSuccess) {
            cudaGraphDestroy(g);
            return mr_fail(ctx, MR_ECUDA, "cudaGraphInstantiate", cudaGetErrorString(ce));
        }
    }
    cudaGraphDestroy(g);
    MR_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    ctx->graph_launches++;
    return MR_OK;
}

static int process_main_frame_impl(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                                   const uint8_t *const *side_frames, const float *side_cameras, float *out_points, int *out_count,
                                   int mode)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    const bool async_copy = mode == PMF_ASYNC_COPY;
    CHECK_ARG(ctx, main_frame && main_camera && side_frames && side_cameras && (out_count || mode == PMF_SUBMIT), "null argument");
    CHECK_ARG(ctx, n_side >= 1 && n_side <= MR_MAX_SIDE, "n_side must be in 1..MR_MAX_SIDE");
    for (int i = 0; i < n_side; i++) CHECK_ARG(ctx, side_frames[i], "null side frame");
    if (!ctx->mesh_loaded) return mr_fail(ctx, MR_ENOMESH, "mr_process_main_frame", "loadMesh has not been called");
    size_t N = ctx->N;
    bool dev_out = out_points && mr_is_device_ptr(out_points);
    if (mode == PMF_SUBMIT) {
        // fully asynchronous: nothing below waits for the GPU; rows and count are delivered by device-side stores / DMA
        CHECK_ARG(ctx, out_points, "mr_submit_main_frame needs an output buffer");
        float *out_alias = (float *)device_alias(out_points);
        int *cnt_alias = out_count ? (int *)device_alias(out_count) : nullptr;
        CHECK_ARG(ctx, out_alias && (!out_count || cnt_alias), "mr_submit_main_frame: out_points / out_count must be device or pinned host memory");
        CHECK_ARG(ctx, !out_count || mr_is_device_ptr(out_points) == mr_is_device_ptr(out_count), "mr_submit_main_frame: out_points and out_count must live in the same memory space");
        RC(ensure_copy_stream(ctx));
        ctx->last_S = n_side;
        ctx->last_count = -1;       // not known on the host without a synchronisation (mr_points_device reports -1)
        if (dev_out) {
            int *d_cnt = cnt_alias ? cnt_alias : mr_buf<int>(ctx, "count", 1);
            if (!d_cnt) return mr_fail(ctx, MR_ENOMEM, "mr_submit_main_frame", "alloc");
            ctx->last_rows = out_points;
            return submit_enqueue(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, out_points, d_cnt, false);
        }
        const int slot = ctx->rows_cur;
        float *d_rows = mr_buf<float>(ctx, slot ? "points1" : "points", N * 7);
        int *d_cnt = mr_buf<int>(ctx, slot ? "count1" : "count0", 1);
        if (!d_rows || !d_cnt) return mr_fail(ctx, MR_ENOMEM, "mr_submit_main_frame", "alloc");
        ctx->last_rows = d_rows;
        // this slot's previous rows (two submissions ago) must have left for the host before they are overwritten
        if (ctx->copy_pending[slot]) MR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[slot], 0));
        RC(submit_enqueue(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, d_rows, d_cnt, true));
        MR_CUDA(ctx, cudaEventRecord(ctx->ev_rows_done[slot], ctx->stream));
        MR_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rows_done[slot], 0));
        // DMA (copy engine, no SMs) of the row buffer at full capacity -- the row count is not known on the host without
        // a synchronisation, and a device-side copy kernel over PCIe measured only ~24 GB/s against ~57 GB/s for the
        // DMA; only the first *out_count rows are meaningful.  58 MB at 1080p: hidden under the next frame's compute.
        MR_CUDA(ctx, cudaMemcpyAsync(out_points, d_rows, N * 7 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (out_count) MR_CUDA(ctx, cudaMemcpyAsync(out_count, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
        MR_CUDA(ctx, cudaEventRecord(ctx->ev_copy_done[slot], ctx->copy_stream));
        RC(note_copy_queued(ctx));
        ctx->copy_pending[slot] = true;
        ctx->rows_cur = slot ^ 1;
        return MR_OK;
    }
    const float *d_flows[MR_MAX_SIDE];
    float *depth = nullptr;
    RC(enqueue_flows(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, d_flows, &depth));
    const bool pipelined = async_copy && out_points && !dev_out;
    float *d_out;
    int slot = 0;
    if (dev_out) d_out = out_points;
    else if (pipelined) {
        // ping-pong row buffers: the copy of slot s may still be in flight while slot s^1 is computed
        RC(ensure_copy_stream(ctx));
        slot = ctx->rows_cur;
        d_out = mr_buf<float>(ctx, slot ? "points1" : "points", N * 7);
        if (d_out && ctx->copy_pending[slot]) MR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[slot], 0));
    } else {
        d_out = mr_buf<float>(ctx, "points", N * 7);
        // "points" doubles as row slot 0 of the pipelined calls: a copy of it may still be in flight
        if (d_out && ctx->copy_pending[0]) MR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy_done[0], 0));
    }
    if (!d_out) return mr_fail(ctx, MR_ENOMEM, "mr_process_main_frame", "alloc");
    RC(k_triangulate(ctx, d_flows, n_side, main_camera, side_cameras, depth, d_out, out_count));   // recon.cpp:114 (syncs the stream)
    ctx->last_S = n_side;
    ctx->last_rows = d_out;
    if (pipelined) {
        if (*out_count > 0)
            MR_CUDA(ctx, cudaMemcpyAsync(out_points, d_out, (size_t)*out_count * 7 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
        MR_CUDA(ctx, cudaEventRecord(ctx->ev_copy_done[slot], ctx->copy_stream));
        RC(note_copy_queued(ctx));
        ctx->copy_pending[slot] = true;
        ctx->rows_cur = slot ^ 1;
    } else if (out_points && !dev_out && *out_count > 0) {
        RC(mr_out(ctx, out_points, d_out, (size_t)*out_count * 7 * sizeof(float)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_process_main_frame(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                          const uint8_t *const *side_frames, const float *side_cameras, float *out_points, int *out_count)
{
    return process_main_frame_impl(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, out_points, out_count, PMF_SYNC);
}

int mr_submit_main_frame(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                         const uint8_t *const *side_frames, const float *side_cameras, float *out_points, int *out_count)
{
    return process_main_frame_impl(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, out_points, out_count, PMF_SUBMIT);
}

int mr_process_main_frame_async(mr_context *ctx, const uint8_t *main_frame, const float main_camera[16], int n_side,
                                const uint8_t *const *side_frames, const float *side_cameras, float *out_points, int *out_count)
{
    return process_main_frame_impl(ctx, main_frame, main_camera, n_side, side_frames, side_cameras, out_points, out_count, PMF_ASYNC_COPY);
}

int mr_profile_enable(mr_context *ctx, int on)
{
    CHECK_CTX(ctx);
    ctx->profile = on != 0;
    return MR_OK;
}

int mr_profile_read(mr_context *ctx, double *ms_by_stage, uint64_t *launches_by_stage, int n)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        ctx->prof_ms[r.stage] += ms;
        ctx->prof_launches[r.stage] += r.launches;
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
    for (int i = 0; i < n && i < ST_COUNT; i++) {
        if (ms_by_stage) ms_by_stage[i] = ctx->prof_ms[i];
        if (launches_by_stage) launches_by_stage[i] = ctx->prof_launches[i];
        ctx->prof_ms[i] = 0;
        ctx->prof_launches[i] = 0;
    }
    return ST_COUNT;
}

const char *mr_stage_name(int stage)
{
    static const char *names[ST_COUNT] = {"raster", "shade_mix", "variational_refinement", "cubic_remap", "pyramid_compare",
                                          "triangulate", "normals"};
    return (stage >= 0 && stage < ST_COUNT) ? names[stage] : "";
}

// test hook: the host instantiation of the device's 3x3 Jacobi (same source, jacobi3.cuh)
int mr_debug_jacobi3(const float cov6[6], float evals3[3], float evecs9[9])
{
    if (!cov6 || !evals3 || !evecs9) return MR_EINVAL;
    mr_jacobi3(cov6, evals3, evecs9);
    return MR_OK;
}

// Frame ingest of Configuration::Configuration (configuration.cpp:226-245): BGR frame -> (INTER_AREA shrink) -> gray.
static int ingest_impl(mr_context *ctx, const uint8_t *bgr, int src_width, int src_height, const float *exposure, uint8_t *out_gray)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, bgr && out_gray, "null argument");
    CHECK_ARG(ctx, src_width >= ctx->W && src_height >= ctx->H && src_width <= 64 * ctx->W && src_height <= 64 * ctx->H,
              "the frame must be at least the render size in both directions (INTER_AREA shrinks, configuration.cpp:232-233)");
    const size_t in_bytes = (size_t)src_width * src_height * 3;
    const uint8_t *d_in = (const uint8_t *)mr_in(ctx, bgr, in_bytes, "in_bgr");
    const bool dev_out = mr_is_device_ptr(out_gray);
    uint8_t *d_out = dev_out ? out_gray : mr_buf<uint8_t>(ctx, "ingest_gray", ctx->N);
    if (!d_in || !d_out) return mr_fail(ctx, MR_ENOMEM, "mr_ingest_frame", "alloc");
    RC(k_ingest(ctx, d_in, src_width, src_height, d_out, exposure));
    if (!dev_out) {
        RC(mr_out(ctx, out_gray, d_out, ctx->N));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_ingest_frame(mr_context *ctx, const uint8_t *bgr, int src_width, int src_height, uint8_t *out_gray)
{
    return ingest_impl(ctx, bgr, src_width, src_height, nullptr, out_gray);
}

// The same with estimateExposure's per-frame channel weights instead of BGR2GRAY (configuration.cpp:417-425).
int mr_ingest_frame_exposure(mr_context *ctx, const uint8_t *bgr, int src_width, int src_height, const float exposure_bgr[3], uint8_t *out_gray)
{
    CHECK_CTX(ctx);
    CHECK_ARG(ctx, exposure_bgr, "null exposure");
    return ingest_impl(ctx, bgr, src_width, src_height, exposure_bgr, out_gray);
}

int mr_set_gray_shift(mr_context *ctx, int shift)
{
    CHECK_CTX(ctx);
    CHECK_ARG(ctx, shift == 14 || shift == 15, "shift must be 14 (OpenCV 3.0-3.4.5) or 15 (OpenCV >= 3.4.6 / 4.x)");
    ctx->gray_shift = shift;
    return MR_OK;
}

// Heuristic::filterPoints (heuristic.cpp:55-176) on the device.  points: n x 4 (stride ps floats), normals: n x 3 (stride ns) or
// null; host or device memory (host inputs are staged).  Outputs likewise; out_points / out_normals / out_keep may be null.
static int filter_impl(mr_context *ctx, const float *points, int ps, const float *normals, int ns, size_t n, float radius, float *out_points,
                       int ops, float *out_normals, int ons, int32_t *out_keep, size_t *out_count, bool rows7)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, out_count, "null out_count");
    *out_count = 0;
    CHECK_ARG(ctx, n <= 0x7fffffffu, "too many points");
    CHECK_ARG(ctx, n == 0 || points, "null points");
    CHECK_ARG(ctx, radius == radius, "radius is NaN");
    ctx->filter_info[0] = ctx->filter_info[1] = ctx->filter_info[2] = 0;
    ctx->filter_n = (int)n;
    if (n == 0) return MR_OK;
    const bool in_dev = mr_is_device_ptr(points);
    CHECK_ARG(ctx, !normals || rows7 || mr_is_device_ptr(normals) == in_dev, "points and normals must live in the same memory space");
    const float *d_pts = (const float *)mr_in(ctx, points, n * (size_t)ps * sizeof(float), "fl_in_pts");
    const float *d_nrm = nullptr;
    if (rows7) d_nrm = d_pts + 4;
    else if (normals) d_nrm = (const float *)mr_in(ctx, normals, n * (size_t)ns * sizeof(float), "fl_in_nrm");
    if (!d_pts || (normals && !d_nrm)) return mr_fail(ctx, MR_ENOMEM, "mr_filter_points", "staging");
    const bool out_dev_p = out_points && mr_is_device_ptr(out_points), out_dev_k = out_keep && mr_is_device_ptr(out_keep);
    const bool out_dev_n = out_normals && mr_is_device_ptr(out_normals);
    CHECK_ARG(ctx, !(out_dev_p && out_points == points), "in-place filtering of device buffers is not supported");
    float *d_op = out_points ? (out_dev_p ? out_points : mr_buf<float>(ctx, "fl_out_pts", n * (size_t)ops)) : nullptr;
    float *d_on = nullptr;
    if (rows7) d_on = d_op ? d_op + 4 : nullptr;
    else if (out_normals) d_on = out_dev_n ? out_normals : mr_buf<float>(ctx, "fl_out_nrm", n * (size_t)ons);
    int *d_ok = out_keep ? (out_dev_k ? out_keep : mr_buf<int>(ctx, "fl_out_keep", n)) : nullptr;
    if ((out_points && !d_op) || (out_normals && !rows7 && !d_on) || (out_keep && !d_ok)) return mr_fail(ctx, MR_ENOMEM, "mr_filter_points", "alloc");
    int m = 0;
    RC(k_filter_points(ctx, d_pts, ps, d_nrm, ns, (int)n, radius, d_op, ops, d_on, ons, d_ok, &m, ctx->filter_info));
    *out_count = (size_t)m;
    if (m > 0) {
        if (out_points && !out_dev_p) RC(mr_out(ctx, out_points, d_op, (size_t)m * ops * sizeof(float)));
        if (out_normals && !rows7 && !out_dev_n) RC(mr_out(ctx, out_normals, d_on, (size_t)m * ons * sizeof(float)));
        if (out_keep && !out_dev_k) RC(mr_out(ctx, out_keep, d_ok, (size_t)m * sizeof(int)));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MR_OK;
}

int mr_filter_points(mr_context *ctx, const float *points_xyzw, const float *normals_xyz, size_t n, float radius, float *out_points_xyzw,
                     float *out_normals_xyz, int32_t *out_keep, size_t *out_count)
{
    return filter_impl(ctx, points_xyzw, 4, normals_xyz, 3, n, radius, out_points_xyzw, 4, out_normals_xyz, 3, out_keep, out_count, false);
}

int mr_filter_rows(mr_context *ctx, const float *rows7, size_t n, float radius, float *out_rows7, int32_t *out_keep, size_t *out_count)
{
    return filter_impl(ctx, rows7, 7, rows7 ? rows7 + 4 : nullptr, 7, n, radius, out_rows7, 7, nullptr, 7, out_keep, out_count, true);
}

int mr_filter_info(mr_context *ctx, long long info3[3], float *density, float *score)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    if (info3) for (int i = 0; i < 3; i++) info3[i] = ctx->filter_info[i];
    const size_t n = (size_t)ctx->filter_n;
    if (n && density) MR_CUDA(ctx, cudaMemcpy(density, mr_buf_raw(ctx, "fl_density", 0), n * sizeof(float), cudaMemcpyDefault));
    if (n && score) MR_CUDA(ctx, cudaMemcpy(score, mr_buf_raw(ctx, "fl_score", 0), n * sizeof(float), cudaMemcpyDefault));
    return MR_OK;
}

// test hook: the exact emulation of a left-to-right double accumulation of non-negative float terms (filter.cu: seqsum)
int mr_debug_seqsum(mr_context *ctx, const float *terms, size_t n, double *out)
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, out && (terms || n == 0), "null argument");
    const float *d = n ? (const float *)mr_in(ctx, terms, n * sizeof(float), "fl_dbg_terms") : nullptr;
    if (n && !d) return mr_fail(ctx, MR_ENOMEM, "mr_debug_seqsum", "staging");
    return k_seqsum(ctx, d, (long long)n, out);
}

// Counters of the window-PCA covariance kernel since the context was created (normals.cu: normals_cov_kernel)
int mr_normals_stats(mr_context *ctx, uint64_t out5[5])
{
    CHECK_CTX(ctx);
    SET_DEVICE(ctx);
    CHECK_ARG(ctx, out5, "null argument");
    for (int i = 0; i < 5; i++) out5[i] = 0;
    if (!ctx->bufs.count("nrm_stats")) return MR_OK;
    unsigned long long h[8];
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MR_CUDA(ctx, cudaMemcpy(h, mr_buf_raw(ctx, "nrm_stats", 0), sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; i++) out5[i] = h[i];
    return MR_OK;
}

const float *mr_points_device(mr_context *ctx, int *out_count)
{
    if (!ctx) return nullptr;
    if (out_count) *out_count = ctx->last_count;
    return ctx->last_rows;
}
const float *mr_last_depth_device(mr_context *ctx) { return ctx ? (const float *)mr_buf_raw(ctx, "depth", 0) : nullptr; }
const float *mr_last_flow_device(mr_context *ctx, int side)
{
    if (!ctx || side < 0 || side >= MR_MAX_SIDE) return nullptr;
    return (const float *)mr_buf_raw(ctx, flow_name(side), 0);
}
const uint8_t *mr_last_mixed_device(mr_context *ctx, int side)
{
    if (!ctx || side < 0 || side >= MR_MAX_SIDE) return nullptr;
    return (const uint8_t *)mr_buf_raw(ctx, mixed_name(side), 0);
}

}  // extern "C"
