// raster.cu -- CUDA replacement of the reference's OpenGL stage:
//   Render::loadMesh  render_glx.cpp:230-258      -> k_load_mesh
//   Render::depth     render_glx.cpp:369-397      -> k_raster + k_resolve_depth
//   Render::projected render_glx.cpp:261-367      -> k_raster (side) + k_dilate_shadow + k_shade
//   shader.vert:9-13 / shader.frag:11-25          -> shade_pixel()
//   mixBackground     util.cpp:366-387            -> fused into k_shade, or k_mix_background
//
// Rasterisation is triangle-parallel with a 64-bit visibility buffer:
//   key = (orderable(z_ndc) << 32) | triangle_index,  atomicMin per covered pixel.
// GL_LESS + draw order == smallest z, ties to the lowest triangle index, so the result
// is order-independent and identical to the sequential oracle (orc_raster).
// Edge functions are homogeneous (no clipping needed for triangles crossing the camera
// plane) and exactly negation-symmetric for shared edges => watertight.
#include "common.cuh"

#define BG_KEY 0xFFFFFFFFFFFFFFFFull
#define RASTER_CHUNKS 8

struct TriSetup {
    float e[3][3];
    float zA, zB, zC;
    int valid, x0, x1, y0, y1;
};

__device__ __forceinline__ void clip_vertex(const float *P, const float *v, float *c)
{
#pragma unroll
    for (int r = 0; r < 4; r++) c[r] = ((P[r * 4 + 0] * v[0] + P[r * 4 + 1] * v[1]) + P[r * 4 + 2] * v[2]) + P[r * 4 + 3];
}

__device__ __forceinline__ unsigned int z_to_ord(float z)
{
    unsigned int u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_z(unsigned int u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- loadMesh: de-index + dehomogenise into a triangle soup ---------------------------
// A vertex index outside [0, V) (the reference's readMesh stores -1 for an `f` line it cannot parse, util.cpp)
// raises *bad and yields a NaN corner, which tri_setup_kernel rejects: never an out-of-bounds read.
__global__ void load_mesh_kernel(const float *__restrict__ vtx, int V, const int32_t *__restrict__ faces, int F, float *__restrict__ soup,
                                 int *__restrict__ bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 3) return;
    const int vi = faces[i];
    if (vi < 0 || vi >= V) {
        *bad = 1;
        const float nan = __int_as_float(0x7fc00000);
        soup[3 * i + 0] = soup[3 * i + 1] = soup[3 * i + 2] = nan;
        return;
    }
    const float *p = vtx + 4 * (size_t)vi;
    soup[3 * i + 0] = p[0] / p[3];
    soup[3 * i + 1] = p[1] / p[3];
    soup[3 * i + 2] = p[2] / p[3];
}

int k_load_mesh(mr_context *ctx, const float *d_vtx, int V, const int32_t *d_faces, int F, int *d_bad)
{
    float *soup = mr_buf<float>(ctx, "soup", (size_t)(F > 0 ? F : 1) * 9);
    if (!soup) return mr_fail(ctx, MR_ENOMEM, "soup", "alloc");
    mr_buf<TriSetup>(ctx, "setup", (size_t)(F > 0 ? F : 1));
    ctx->F = F;
    if (F == 0) return MR_OK;
    load_mesh_kernel<<<cdiv(F * 3, 256), 256, 0, ctx->stream>>>(d_vtx, V, d_faces, F, soup, d_bad);
    MR_LAUNCH_CHECK(ctx, "load_mesh_kernel");
    return MR_OK;
}

// ---- triangle setup -------------------------------------------------------------------
__global__ void tri_setup_kernel(const float *__restrict__ soup, int F, Mat4 P, int W, int H, TriSetup *__restrict__ out)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float *tri9 = soup + 9 * f;
    float c[3][4];
#pragma unroll
    for (int i = 0; i < 3; i++) clip_vertex(P.m, tri9 + 3 * i, c[i]);
    float e[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float *a = c[(i + 1) % 3], *b = c[(i + 2) % 3];
        e[i][0] = a[1] * b[3] - b[1] * a[3];
        e[i][1] = b[0] * a[3] - a[0] * b[3];
        e[i][2] = a[0] * b[1] - b[0] * a[1];
    }
    float det = (c[0][0] * e[0][0] + c[0][1] * e[0][1]) + c[0][3] * e[0][2];
    TriSetup t;
    t.valid = 0;
    t.x0 = 0; t.x1 = -1; t.y0 = 0; t.y1 = -1;
    t.zA = t.zB = t.zC = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) t.e[i][k] = 0.f;
    if (det != 0.f && isfinite(det)) {
        float s = det > 0.f ? 1.f : -1.f;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int k = 0; k < 3; k++) t.e[i][k] = s * e[i][k];
        float adet = s * det;
        t.zA = ((t.e[0][0] * c[0][2] + t.e[1][0] * c[1][2]) + t.e[2][0] * c[2][2]) / adet;
        t.zB = ((t.e[0][1] * c[0][2] + t.e[1][1] * c[1][2]) + t.e[2][1] * c[2][2]) / adet;
        t.zC = ((t.e[0][2] * c[0][2] + t.e[1][2] * c[1][2]) + t.e[2][2] * c[2][2]) / adet;
        t.valid = 1;
        t.x0 = 0; t.x1 = W - 1; t.y0 = 0; t.y1 = H - 1;
        if (c[0][3] > 0.f && c[1][3] > 0.f && c[2][3] > 0.f) {
            float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
            bool bad = false;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float x = c[i][0] / c[i][3], y = c[i][1] / c[i][3];
                if (!(x == x) || !(y == y)) bad = true;
                xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
            }
            if (!bad) {
                float fx0 = (xmin + 1.f) * 0.5f * (float)W - 2.f, fx1 = (xmax + 1.f) * 0.5f * (float)W + 2.f;
                float fy0 = (1.f - ymax) * 0.5f * (float)H - 2.f, fy1 = (1.f - ymin) * 0.5f * (float)H + 2.f;
                if (fx0 > 0.f) t.x0 = fx0 < (float)W ? (int)fx0 : W;
                if (fx1 < (float)(W - 1)) t.x1 = fx1 > -1.f ? (int)fx1 : -1;
                if (fy0 > 0.f) t.y0 = fy0 < (float)H ? (int)fy0 : H;
                if (fy1 < (float)(H - 1)) t.y1 = fy1 > -1.f ? (int)fy1 : -1;
            }
        }
    }
    out[f] = t;
}

__device__ __forceinline__ bool edge_inside(const float *e, float X, float Y)
{
    float l = (e[0] * X + e[1] * Y) + e[2];
    if (l > 0.f) return true;
    if (l < 0.f) return false;
    if (!(l == 0.f)) return false;
    return e[0] > 0.f || (e[0] == 0.f && e[1] > 0.f);
}

// grid = (F, RASTER_CHUNKS): each block walks a horizontal band of one triangle's bounding box.
__global__ void __launch_bounds__(128) raster_kernel(const TriSetup *__restrict__ setups, int W, int H,
                                                     unsigned long long *__restrict__ vis)
{
    __shared__ TriSetup t;
    int f = blockIdx.x;
    if (threadIdx.x < sizeof(TriSetup) / 4) ((int *)&t)[threadIdx.x] = ((const int *)&setups[f])[threadIdx.x];
    __syncthreads();
    if (!t.valid) return;
    int bw = t.x1 - t.x0 + 1, bh = t.y1 - t.y0 + 1;
    if (bw <= 0 || bh <= 0) return;
    int rows_per = (bh + gridDim.y - 1) / gridDim.y;
    int r0 = t.y0 + blockIdx.y * rows_per;
    int r1 = min(r0 + rows_per, t.y1 + 1);
    if (r0 >= r1) return;
    float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    int total = bw * (r1 - r0);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int row = r0 + i / bw, col = t.x0 + i % bw;
        float X = ((float)col + 0.5f) * sx - 1.0f;
        float Y = 1.0f - ((float)row + 0.5f) * sy;
        if (!edge_inside(t.e[0], X, Y) || !edge_inside(t.e[1], X, Y) || !edge_inside(t.e[2], X, Y)) continue;
        float z = ((t.zA * X + t.zB * Y) + t.zC) + 0.0f;
        if (!(z >= -1.0f && z < 1.0f)) continue;
        unsigned long long key = ((unsigned long long)z_to_ord(z) << 32) | (unsigned int)f;
        atomicMin(&vis[(size_t)row * W + col], key);
    }
}

int k_raster(mr_context *ctx, const Mat4 &P, unsigned long long *d_vis)
{
    MR_CUDA(ctx, cudaMemsetAsync(d_vis, 0xFF, ctx->N * sizeof(unsigned long long), ctx->stream));
    if (ctx->F == 0) return MR_OK;
    float *soup = mr_buf<float>(ctx, "soup", 0);
    TriSetup *setups = mr_buf<TriSetup>(ctx, "setup", 0);
    tri_setup_kernel<<<cdiv(ctx->F, 128), 128, 0, ctx->stream>>>(soup, ctx->F, P, ctx->W, ctx->H, setups);
    MR_LAUNCH_CHECK(ctx, "tri_setup_kernel");
    dim3 grid(ctx->F, RASTER_CHUNKS);
    raster_kernel<<<grid, 128, 0, ctx->stream>>>(setups, ctx->W, ctx->H, d_vis);
    MR_LAUNCH_CHECK(ctx, "raster_kernel");
    return MR_OK;
}

__global__ void resolve_depth_kernel(const unsigned long long *__restrict__ vis, size_t N, float *__restrict__ depth)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    unsigned long long k = vis[i];
    depth[i] = (k == BG_KEY) ? MR_BACKGROUND_DEPTH : ord_to_z((unsigned int)(k >> 32));
}

int k_resolve_depth(mr_context *ctx, const unsigned long long *d_vis, float *d_depth)
{
    resolve_depth_kernel<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(d_vis, ctx->N, d_depth);
    MR_LAUNCH_CHECK(ctx, "resolve_depth_kernel");
    return MR_OK;
}

// ---- shadow-map dilation (render_glx.cpp:287-314) in closed, parallel form -------------
// With o = undilated map in GL (bottom-up) row order:
//   HF[0][j]  = min(o[0][0..j+1])                     (the reference's in-place row-0 pass is a prefix MIN)
//   HF[i][j]  = max(o[i][j-1], o[i][j], o[i][j+1])    i >= 1
//   out[i][j] = max(HF[i-1][j], HF[i][j], HF[i+1][j]) (rows clipped to [0, H-1]); columns 0 and W-1 untouched.
// Equivalence with the sequential loop is tested in tests/test_oracle_path.py.
__global__ void __launch_bounds__(1024) row_prefix_min_kernel(const float *__restrict__ row, int W, float *__restrict__ out)
{
    // single block, inclusive prefix-min over W <= 1024 * items values: warp shuffles + one smem pass
    __shared__ float warp_min[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float carry = __int_as_float(0x7f800000);  // +inf
    for (int base = 0; base < W; base += blockDim.x) {
        int i = base + threadIdx.x;
        float v = (i < W) ? row[i] : __int_as_float(0x7f800000);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d && t < v) v = t;
        }
        if (lane == 31) warp_min[wid] = v;
        __syncthreads();
        float pre = carry;
        for (int w = 0; w < wid; w++) pre = fminf(pre, warp_min[w]);
        if (pre < v) v = pre;
        if (i < W) out[i] = v;
        float tot = carry;
        for (int w = 0; w < nw; w++) tot = fminf(tot, warp_min[w]);
        __syncthreads();
        carry = tot;
    }
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float m = b;
    if (a > m) m = a;
    if (c > m) m = c;
    return m;
}

__global__ void dilate_kernel(const float *__restrict__ d, const float *__restrict__ pm, int W, int H, float *__restrict__ out)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    int row = blockIdx.y * blockDim.y + threadIdx.y;  // top-down
    if (col >= W || row >= H) return;
    size_t idx = (size_t)row * W + col;
    if (col == 0 || col == W - 1) { out[idx] = d[idx]; return; }
    int i = H - 1 - row;  // GL row
    float best = 0.f;
    bool first = true;
#pragma unroll
    for (int di = -1; di <= 1; di++) {
        int ii = i + di;
        if (ii < 0 || ii > H - 1) continue;
        float hf;
        if (ii == 0) hf = pm[col + 1];
        else {
            const float *r = d + (size_t)(H - 1 - ii) * W;
            hf = max3(r[col - 1], r[col], r[col + 1]);
        }
        if (first || hf > best) best = hf;
        first = false;
    }
    out[idx] = best;
}

int k_dilate_shadow(mr_context *ctx, const float *d_depth_td, float *d_out_td)
{
    int W = ctx->W, H = ctx->H;
    float *pm = mr_buf<float>(ctx, "rowmin", (size_t)W);
    row_prefix_min_kernel<<<1, 1024, 0, ctx->stream>>>(d_depth_td + (size_t)(H - 1) * W, W, pm);
    MR_LAUNCH_CHECK(ctx, "row_prefix_min_kernel");
    dim3 b(32, 8), g(cdiv(W, 32), cdiv(H, 8));
    dilate_kernel<<<g, b, 0, ctx->stream>>>(d_depth_td, pm, W, H, d_out_td);
    MR_LAUNCH_CHECK(ctx, "dilate_kernel");
    return MR_OK;
}

// ---- fragment shading (shader.frag) + optional fused mixBackground --------------------
__device__ __forceinline__ int wrapi(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// returns true if visible && inframe; *r = predicted gray value
__device__ __forceinline__ bool shade_pixel(const float *__restrict__ soup, const Mat4 &Pmain, const Mat4 &Pside,
                                            const uint8_t *__restrict__ frame, const float *__restrict__ shadow_td,
                                            int W, int H, int row, int col, int tri, uint8_t *r)
{
    *r = 0;
    const float *tv = soup + 9 * (size_t)tri;
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = tv[i];
    float c[3][4];
#pragma unroll
    for (int i = 0; i < 3; i++) clip_vertex(Pmain.m, v + 3 * i, c[i]);
    float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    float X = ((float)col + 0.5f) * sx - 1.0f, Y = 1.0f - ((float)row + 0.5f) * sy;
    float l[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float *a = c[(i + 1) % 3], *b = c[(i + 2) % 3];
        float e0 = a[1] * b[3] - b[1] * a[3];
        float e1 = b[0] * a[3] - a[0] * b[3];
        float e2 = a[0] * b[1] - b[0] * a[1];
        l[i] = (e0 * X + e1 * Y) + e2;
    }
    float L = (l[0] + l[1]) + l[2];
    float pos[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = ((l[0] * v[0 + k] + l[1] * v[3 + k]) + l[2] * v[6 + k]) / L;
    float sc[4];
    clip_vertex(Pside.m, pos, sc);
    float nx = sc[0] / sc[3], ny = sc[1] / sc[3], nz = sc[2] / sc[3];
    bool inframe = nx > -1.0f && nx < 1.0f && ny > -1.0f && ny < 1.0f;
    if (!inframe) return false;
    float u = nx * 0.5f + 0.5f, vv = ny * 0.5f + 0.5f;
    int si = (int)floorf(u * (float)W), sj = (int)floorf(vv * (float)H);
    si = si < 0 ? 0 : (si > W - 1 ? W - 1 : si);
    sj = sj < 0 ? 0 : (sj > H - 1 ? H - 1 : sj);
    float shadowDepth = shadow_td[(size_t)(H - 1 - sj) * W + si];
    bool visible = shadowDepth + 0.01f > nz;
    if (!visible) return false;
    float tx = u * (float)W - 0.5f, ty = vv * (float)H - 0.5f;
    float fx0 = floorf(tx), fy0 = floorf(ty);
    float ax = tx - fx0, ay = ty - fy0;
    int i0 = wrapi((int)fx0, W), i1 = wrapi((int)fx0 + 1, W);
    int j0 = wrapi((int)fy0, H), j1 = wrapi((int)fy0 + 1, H);
    float t00 = frame[(size_t)(H - 1 - j0) * W + i0], t10 = frame[(size_t)(H - 1 - j0) * W + i1];
    float t01 = frame[(size_t)(H - 1 - j1) * W + i0], t11 = frame[(size_t)(H - 1 - j1) * W + i1];
    float bx = 1.0f - ax, by = 1.0f - ay;
    float val = by * (bx * t00 + ax * t10) + ay * (bx * t01 + ax * t11);
    float rr = rintf(val);
    *r = (uint8_t)(rr < 0.f ? 0.f : (rr > 255.f ? 255.f : rr));
    return true;
}

// MODE 0: write RGB (Render::projected).  MODE 1: fused mixBackground (util.cpp:366-387):
// mixed = masked ? main_frame : R ; depth = masked ? 1.0 : depth.
template <int MODE>
__global__ void __launch_bounds__(256) shade_kernel(const unsigned long long *__restrict__ vis, const float *__restrict__ soup,
                                                    Mat4 Pmain, Mat4 Pside, const uint8_t *__restrict__ frame,
                                                    const float *__restrict__ shadow_td, int W, int H,
                                                    uint8_t *__restrict__ rgb, const uint8_t *__restrict__ main_frame,
                                                    float *__restrict__ depth, uint8_t *__restrict__ mixed)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    int row = blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= W || row >= H) return;
    size_t idx = (size_t)row * W + col;
    unsigned long long k = vis[idx];
    uint8_t r = 0;
    bool ok = false;
    if (MODE == 1) {
        // pixels already masked by an earlier side camera stay masked (quirk C10) -- skip the shading work
        if (depth[idx] == MR_BACKGROUND_DEPTH) { mixed[idx] = main_frame[idx]; return; }
    }
    if (k != BG_KEY) ok = shade_pixel(soup, Pmain, Pside, frame, shadow_td, W, H, row, col, (int)(unsigned int)k, &r);
    if (MODE == 0) {
        rgb[3 * idx + 0] = ok ? r : 0;
        rgb[3 * idx + 1] = ok ? 255 : 0;
        rgb[3 * idx + 2] = ok ? 255 : 0;
    } else {
        if (!ok) {
            mixed[idx] = main_frame[idx];
            depth[idx] = MR_BACKGROUND_DEPTH;
        } else
            mixed[idx] = r;
    }
}

int k_shade(mr_context *ctx, const unsigned long long *d_vis_main, const Mat4 &Pmain, const Mat4 &Pside,
            const uint8_t *d_side_frame, const float *d_shadow_td, uint8_t *d_rgb, const uint8_t *d_main_frame,
            float *d_depth_inout, uint8_t *d_mixed)
{
    int W = ctx->W, H = ctx->H;
    float *soup = mr_buf<float>(ctx, "soup", 0);
    dim3 b(32, 8), g(cdiv(W, 32), cdiv(H, 8));
    if (d_rgb)
        shade_kernel<0><<<g, b, 0, ctx->stream>>>(d_vis_main, soup, Pmain, Pside, d_side_frame, d_shadow_td, W, H, d_rgb,
                                                  nullptr, nullptr, nullptr);
    else
        shade_kernel<1><<<g, b, 0, ctx->stream>>>(d_vis_main, soup, Pmain, Pside, d_side_frame, d_shadow_td, W, H, nullptr,
                                                  d_main_frame, d_depth_inout, d_mixed);
    MR_LAUNCH_CHECK(ctx, "shade_kernel");
    return MR_OK;
}

__global__ void mix_background_kernel(const uint8_t *__restrict__ rgb, const uint8_t *__restrict__ bg, float *__restrict__ depth,
                                      size_t N, uint8_t *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (depth[i] == MR_BACKGROUND_DEPTH || !rgb[3 * i + 1]) {
        out[i] = bg[i];
        depth[i] = MR_BACKGROUND_DEPTH;
    } else
        out[i] = rgb[3 * i];
}

int k_mix_background(mr_context *ctx, const uint8_t *d_rgb, const uint8_t *d_bg, float *d_depth, uint8_t *d_out)
{
    mix_background_kernel<<<(unsigned)((ctx->N + 255) / 256), 256, 0, ctx->stream>>>(d_rgb, d_bg, d_depth, ctx->N, d_out);
    MR_LAUNCH_CHECK(ctx, "mix_background_kernel");
    return MR_OK;
}

// ---- single-pixel depth queries (heuristic.cpp:285-341,456) ------------------------------------------------
// Heuristic::chooseCameras renders a full depth map from a synthetic "viewer" camera for each of its 200
// surface shots and then reads ONE pixel per candidate camera (filterCameras, heuristic.cpp:306-311).
// The queries are answered on the device straight from the visibility buffer, so only n floats per shot
// cross PCIe instead of a W x H map.  Addressing follows the reference's depth.at<float>(row, col) on a
// continuous Mat (its `col > depth.cols` test lets col == W through, which reads the next row's first
// pixel); rows outside [0, H) or cols outside [0, W] give the background depth.
__global__ void depth_samples_kernel(const unsigned long long *__restrict__ vis, int W, int H, const int32_t *__restrict__ rows,
                                     const int32_t *__restrict__ cols, int n, float *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = rows[i], c = cols[i];
    float d = MR_BACKGROUND_DEPTH;
    if (r >= 0 && r < H && c >= 0 && c <= W) {
        long long idx = (long long)r * W + c;
        long long last = (long long)W * H - 1;
        if (idx > last) idx = last;
        unsigned long long k = vis[idx];
        if (k != BG_KEY) d = ord_to_z((unsigned int)(k >> 32));
    }
    out[i] = d;
}

int k_depth_samples(mr_context *ctx, const unsigned long long *d_vis, const int32_t *d_rows, const int32_t *d_cols, int n, float *d_out)
{
    if (n <= 0) return MR_OK;
    depth_samples_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(d_vis, ctx->W, ctx->H, d_rows, d_cols, n, d_out);
    MR_LAUNCH_CHECK(ctx, "depth_samples_kernel");
    return MR_OK;
}
