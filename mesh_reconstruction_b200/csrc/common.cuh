// common.cuh -- context, scratch arena and helpers shared by the kernels of
// libmeshrecon_b200.so.  Everything here is sm_100a-only; there is no CPU fallback.
//
// Numerical contract: all kernels are compiled with -fmad=false and use IEEE
// division / sqrt, and are written in the SAME operation order as the CPU oracle
// (oracle/recon_oracle.c, oracle/cvprims.py), so integer/byte outputs and most
// float outputs are bit-identical to it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/meshrecon_b200.h"

#define MR_MAX_LEVELS 16

struct Mat4 {
    float m[16];
};

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

// Per-main-camera constants of triangulatePixels (util.cpp:85-89,99,174,209,219),
// computed on the host exactly like the oracle's tri_ctx_init.
struct TriConst {
    float Pinv[16];
    float M[MR_MAX_SIDE][16];
    float B[MR_MAX_SIDE][6];
    float pd[MR_MAX_SIDE][2];
    float pw[MR_MAX_SIDE][4];
    float centers[(MR_MAX_SIDE + 1) * 3];
    int S;
};

// per-stage device timing (bench.py's roofline / stage breakdown); off by default
enum MrStage { ST_RASTER = 0, ST_SHADE, ST_VR, ST_REMAP, ST_COMPARE, ST_TRI, ST_NORMALS, ST_COUNT };
struct ProfRec {
    int stage;
    cudaEvent_t a, b;
    uint64_t launches;
};

struct mr_context {
    int device = 0, W = 0, H = 0;
    size_t N = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    std::map<std::string, DevBuf> bufs;
    // mesh (Render::loadMesh)
    int F = 0;
    bool mesh_loaded = false;                    // a successful mr_load_mesh happened (render calls give MR_ENOMESH otherwise)
    int gray_shift = 15;                         // mr_set_gray_shift: BGR2GRAY fixed-point coefficients (15: OpenCV >= 3.4.6 / 4.x, 14: older 3.x)
    bool use_farneback = false;   // mr_set_use_farneback: the reference's -f switch for mr_process_main_frame
    // last results
    int last_count = 0;
    int last_S = 0;
    const float *last_rows = nullptr;            // device buffer holding the last main frame's rows (mr_points_device)
    long long filter_info[3] = {0, 0, 0};        // last mr_filter_points: neighbour pairs, power iterations, thinning rounds
    int filter_n = 0;
    // pinned host scratch for small readbacks
    int *h_count = nullptr;
    int *h_xchg = nullptr;                        // pinned per-rank counts of mr_allgather_points (exchange.cu)
    int h_xchg_cap = 0;
    static constexpr int N_PUSH = 4;              // peer-memory pushes of mr_xchg_push go round-robin over these (exchange.cu)
    cudaStream_t push_stream[N_PUSH] = {};
    cudaEvent_t ev_push[N_PUSH] = {};
    int push_next = 0;
    // pyramid geometry of compare()
    int n_levels = 0;
    int lw[MR_MAX_LEVELS], lh[MR_MAX_LEVELS];
    size_t loff[MR_MAX_LEVELS];
    size_t pyr_total = 0;
    // pipelined device->host copies of point rows (mr_process_main_frame_async)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_rows_done[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    static constexpr int COPY_RING = 32;          // completion events of the last row copies (mr_wait_copies_until)
    cudaEvent_t ev_copy_ring[COPY_RING] = {};
    unsigned long long copy_seq = 0;              // row copies queued so far by the async / submit calls
    int rows_cur = 0;
    // mr_submit_main_frame: the per-frame launch sequence replayed as one CUDA graph (api.cu: submit_enqueue)
    int graphs_mode = 1;                          // mr_set_use_graphs: 0 never, 1 when the rows go to the host, 2 always; zeroed if a capture is ever refused
    bool graph_warm = false;                      // a plain run of shape graph_key has allocated every buffer
    unsigned long long graph_key = 0;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_launches = 0;
    // profiling
    bool profile = false;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[ST_COUNT] = {0};
    uint64_t prof_launches[ST_COUNT] = {0};
};

// RAII: brackets a stage with CUDA events on the context stream when profiling is on
struct StageScope {
    mr_context *ctx;
    ProfRec r;
    bool on;
    StageScope(mr_context *c, int stage) : ctx(c), on(c->profile)
    {
        if (!on) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        r.stage = stage; r.a = get(); r.b = get(); r.launches = ctx->launches;
        cudaEventRecord(r.a, ctx->stream);
    }
    void end()
    {
        if (!on) return;
        on = false;
        cudaEventRecord(r.b, ctx->stream);
        r.launches = ctx->launches - r.launches;
        ctx->prof_pending.push_back(r);
    }
    ~StageScope() { end(); }
};

extern thread_local std::string g_mr_create_error;

int mr_fail(mr_context *ctx, int code, const char *what, const char *detail);

#define MR_CUDA(ctx, call)                                                          \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) return mr_fail(ctx, MR_ECUDA, #call, cudaGetErrorString(e__)); \
    } while (0)

#define MR_LAUNCH_CHECK(ctx, name)                                                  \
    do {                                                                            \
        (ctx)->launches++;                                                          \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) return mr_fail(ctx, MR_ECUDA, name, cudaGetErrorString(e__)); \
    } while (0)

// scratch arena: named device buffers, grown on demand, owned by the context
void *mr_buf_raw(mr_context *ctx, const char *name, size_t bytes);
template <class T>
static inline T *mr_buf(mr_context *ctx, const char *name, size_t count)
{
    return (T *)mr_buf_raw(ctx, name, count * sizeof(T));
}

bool mr_is_device_ptr(const void *p);
// Returns a device pointer holding `bytes` of `p`: `p` itself if it already lives on the
// device, otherwise a staging buffer filled by an async H2D copy on the context stream.
const void *mr_in(mr_context *ctx, const void *p, size_t bytes, const char *staging);
// Copies a device result to the user's buffer (host: async D2H + stream sync by the caller).
int mr_out(mr_context *ctx, void *dst, const void *src_dev, size_t bytes);

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel (shared by every context of
// this process on that device): keep one registry per (device, kernel) and only ever raise the limit.
cudaError_t mr_ensure_smem_raw(int device, const void *kernel, size_t bytes);   // api.cu (mutex-guarded)
template <class K>
static inline cudaError_t mr_ensure_smem(mr_context *ctx, K kernel, size_t bytes)
{
    return mr_ensure_smem_raw(ctx->device, (const void *)kernel, bytes);
}

// ---- stage launchers (all enqueue on ctx->stream; device pointers only) ------------
// raster.cu
int k_load_mesh(mr_context *ctx, const float *d_vtx, int V, const int32_t *d_faces, int F, int *d_bad);
int k_raster(mr_context *ctx, const Mat4 &P, unsigned long long *d_vis);
int k_resolve_depth(mr_context *ctx, const unsigned long long *d_vis, float *d_depth);
int k_dilate_shadow(mr_context *ctx, const float *d_depth_td, float *d_out_td);
int k_shade(mr_context *ctx, const unsigned long long *d_vis_main, const Mat4 &Pmain, const Mat4 &Pside,
            const uint8_t *d_side_frame, const float *d_shadow_td, uint8_t *d_rgb /*or null*/,
            const uint8_t *d_main_frame /*or null*/, float *d_depth_inout /*or null*/, uint8_t *d_mixed /*or null*/);
int k_depth_samples(mr_context *ctx, const unsigned long long *d_vis, const int32_t *d_rows, const int32_t *d_cols, int n, float *d_out);
int k_depth_query(mr_context *ctx, const float *d_cams, int n_cameras, const int32_t *d_rows, const int32_t *d_cols, int n, float *d_out);
int k_mix_background(mr_context *ctx, const uint8_t *d_rgb, const uint8_t *d_bg, float *d_depth, uint8_t *d_out);
// flow.cu
int k_variational_refinement(mr_context *ctx, const uint8_t *d_i0, const uint8_t *d_i1, float *d_flow4);
int k_flow_remap(mr_context *ctx, const float *d_flow, int stride_floats, const uint8_t *d_img, uint8_t *d_out);
int k_compare(mr_context *ctx, const uint8_t *d_prev, const uint8_t *d_next, float *d_out, int out_stride, int out_off);
int k_farneback(mr_context *ctx, const uint8_t *d_prev, const uint8_t *d_next, float *d_flow4);   // farneback.cu
int mr_flow_init_tables(mr_context *ctx);
// ingest.cu
int k_ingest(mr_context *ctx, const uint8_t *d_bgr, int src_w, int src_h, uint8_t *d_gray, const float *exposure = nullptr);
// filter.cu
int k_filter_points(mr_context *ctx, const float *d_pts, int pstride, const float *d_nrm, int nstride, int n, float radius, float *d_out_pts,
                    int opstride, float *d_out_nrm, int onstride, int *d_out_keep, int *h_count, long long *info);
int k_seqsum(mr_context *ctx, const float *d_terms, long long n, double *h_out);
// tri.cu
int k_image_gradient(mr_context *ctx, const float *d_img, float *d_grad2);
int k_triangulate(mr_context *ctx, const float *const *d_flows_host_array, int S, const float *Pmain, const float *cams,
                  const float *d_depth, float *d_out7, int *out_count, int *d_count_out = nullptr);
void mr_tri_const_init(TriConst *c, const float *Pmain, const float *cams, int S);
void mr_camera_center(const float *P, float *c3);
