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
#include <algorithm>

#include "common.cuh"

#define BG_KEY 0xFFFFFFFFFFFFFFFFull

struct alignas(16) TriSetup {
    float e[3][3];
    float zA, zB, zC;
    int valid, x0, x1, y0, y1;
    int pad[3];            // 80 bytes: whole-struct 128-bit loads
};
static_assert(sizeof(TriSetup) == 80, "TriSetup layout");

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
// The pixel tests below (edge_inside x 3, z range) are what DEFINES coverage; the bounding box only has to contain every
// accepted pixel.  Triangles with all w > 0 get the box of their projected corners (+ 2 pixels), like the oracle.  A
// triangle that touches or crosses the camera plane has no such box; the oracle walks the whole image for it.  Here
// its box is that of the convex region  { l0, l1, l2 >= -eps, -1 - eps <= z <= 1 + eps }  intersected with the image
// rectangle (Sutherland-Hodgman in double on the five affine functions the pixel test itself evaluates, each relaxed by
// 16 x the worst-case rounding error of its float evaluation, + 2 pixels): a superset of the accepted pixels, so the
// result is unchanged, but a Render::depth from a viewer sitting ON the mesh (faceCamera, heuristic.cpp:193-247) no longer
// costs W x H tests per face.
struct HalfPlane {
    double a, b, c;        // a X + b Y + c >= 0 (already relaxed)
};

__device__ __forceinline__ HalfPlane relaxed_plane(float a, float b, float c, float shift, float sign)
{
    HalfPlane h;
    h.a = (double)sign * (double)a;
    h.b = (double)sign * (double)b;
    h.c = (double)sign * (double)c + (double)shift;
    h.c += ((double)fabsf(a) + (double)fabsf(b) + (double)fabsf(c) + 1.0) * (1.0 / 1048576.0) + 1e-30;
    return h;
}

// clips the polygon (px, py, n) by h; returns the new vertex count (<= n + 1)
__device__ int clip_poly(double *px, double *py, int n, const HalfPlane &h)
{
    double qx[12], qy[12];
    int m = 0;
    for (int i = 0; i < n; i++) {
        const int j = i + 1 == n ? 0 : i + 1;
        const double di = h.a * px[i] + h.b * py[i] + h.c, dj = h.a * px[j] + h.b * py[j] + h.c;
        if (di >= 0.0) { qx[m] = px[i]; qy[m] = py[i]; m++; }
        if ((di >= 0.0) != (dj >= 0.0)) {
            const double t = di / (di - dj);
            qx[m] = px[i] + (px[j] - px[i]) * t;
            qy[m] = py[i] + (py[j] - py[i]) * t;
            m++;
        }
    }
    for (int i = 0; i < m; i++) { px[i] = qx[i]; py[i] = qy[i]; }
    return m;
}

__device__ void tighten_bbox(TriSetup &t, int W, int H)
{
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) amax = fmaxf(amax, fabsf(t.e[i][k]));
    amax = fmaxf(amax, fmaxf(fabsf(t.zA), fmaxf(fabsf(t.zB), fabsf(t.zC))));
    if (!(amax < 1e30f)) return;                       // NaN / inf / huge coefficients: keep the whole image
    double px[12] = {-1.0, 1.0, 1.0, -1.0}, py[12] = {-1.0, -1.0, 1.0, 1.0};
    int n = 4;
#pragma unroll 1
    for (int k = 0; k < 5 && n > 0; k++) {
        HalfPlane h;
        if (k < 3) h = relaxed_plane(t.e[k][0], t.e[k][1], t.e[k][2], 0.f, 1.f);
        else if (k == 3) h = relaxed_plane(t.zA, t.zB, t.zC, 1.f, 1.f);        // z >= -1
        else h = relaxed_plane(t.zA, t.zB, t.zC, 1.f, -1.f);                   // z <= 1
        n = clip_poly(px, py, n, h);
    }
    if (n == 0) { t.x0 = 0; t.x1 = -1; t.y0 = 0; t.y1 = -1; return; }
    double xmin = 2.0, xmax = -2.0, ymin = 2.0, ymax = -2.0;
    for (int i = 0; i < n; i++) {
        xmin = fmin(xmin, px[i]); xmax = fmax(xmax, px[i]);
        ymin = fmin(ymin, py[i]); ymax = fmax(ymax, py[i]);
    }
    if (!(xmin >= -1.5 && xmax <= 1.5 && ymin >= -1.5 && ymax <= 1.5)) return;
    // pixel centre of column c is X = (c + 0.5) * 2 / W - 1, of row r  Y = 1 - (r + 0.5) * 2 / H
    const double c0 = (xmin + 1.0) * 0.5 * W - 0.5 - 2.0, c1 = (xmax + 1.0) * 0.5 * W - 0.5 + 2.0;
    const double r0 = (1.0 - ymax) * 0.5 * H - 0.5 - 2.0, r1 = (1.0 - ymin) * 0.5 * H - 0.5 + 2.0;
    t.x0 = max(0, (int)floor(c0)); t.x1 = min(W - 1, (int)ceil(c1));
    t.y0 = max(0, (int)floor(r0)); t.y1 = min(H - 1, (int)ceil(r1));
}

__device__ __forceinline__ void compute_setup(const float *__restrict__ tri9, const Mat4 &P, int W, int H, TriSetup &t)
{
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
        bool boxed = false;
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
                boxed = true;
            }
        }
        if (!boxed) tighten_bbox(t, W, H);
    }
}

// Work split by box size: up to RASTER_SMALL_MAX pixels -> ONE WARP of raster_small_kernel walks the box; above that the
// triangle is appended to the medium or the large list, whose (triangle, row band) items the persistent warps of
// raster_big_kernel consume (16 bands for a medium triangle, walked like a small box; 64 for a large one, one row at a time, columns restricted
// to the row's span of the relaxed half-planes).
#define RASTER_SMALL_MAX 1024
#define RASTER_MEDIUM_MAX 65536
#define RASTER_BANDS_MEDIUM 16
#define RASTER_BANDS_LARGE 64

// lists: cnt[0] = medium count, cnt[1] = large count; medium indices at list[0 .. F), large ones at list[F .. 2F)
__global__ void tri_setup_kernel(const float *__restrict__ soup, int F, Mat4 P, int W, int H, TriSetup *__restrict__ out,
                                 int *__restrict__ list, int *__restrict__ cnt, int direct)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    TriSetup t;
    compute_setup(soup + 9 * (size_t)f, P, W, H, t);
    out[f] = t;
    const long long area = (long long)max(0, t.x1 - t.x0 + 1) * (long long)max(0, t.y1 - t.y0 + 1);
    if (!direct && t.valid && area > RASTER_SMALL_MAX) {
        if (area <= RASTER_MEDIUM_MAX) list[atomicAdd(cnt, 1)] = f;
        else list[F + atomicAdd(cnt + 1, 1)] = f;
    }
}

__device__ __forceinline__ bool edge_inside(const float *e, float X, float Y)
{
    float l = (e[0] * X + e[1] * Y) + e[2];
    if (l > 0.f) return true;
    if (l < 0.f) return false;
    if (!(l == 0.f)) return false;
    return e[0] > 0.f || (e[0] == 0.f && e[1] > 0.f);
}

// the pixel test of the oracle's orc_raster: true + z if triangle t covers the centre of pixel (row, col)
__device__ __forceinline__ bool pixel_test(const TriSetup &t, int row, int col, float sx, float sy, float *z_out)
{
    float X = ((float)col + 0.5f) * sx - 1.0f;
    float Y = 1.0f - ((float)row + 0.5f) * sy;
    if (!edge_inside(t.e[0], X, Y) || !edge_inside(t.e[1], X, Y) || !edge_inside(t.e[2], X, Y)) return false;
    float z = ((t.zA * X + t.zB * Y) + t.zC) + 0.0f;
    if (!(z >= -1.0f && z < 1.0f)) return false;
    *z_out = z;
    return true;
}

__device__ __forceinline__ void load_setup_warp(const TriSetup *__restrict__ src, TriSetup &t)
{
    // every lane reads the same 80 bytes: broadcast loads served by one L1 line pair
    const int4 *p = reinterpret_cast<const int4 *>(src);
    int4 *q = reinterpret_cast<int4 *>(&t);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(TriSetup) / 16); i++) q[i] = __ldg(p + i);
}

// i -> (i / bw, i % bw) for 0 <= i < 2^17, 1 <= bw < 2^13 without the integer-division sequence: float reciprocal + fix-up
__device__ __forceinline__ void divmod_small(int i, int bw, float inv_bw, int *q, int *r)
{
    int qq = __float2int_rz(((float)i + 0.5f) * inv_bw);
    int rr = i - qq * bw;
    if (rr < 0) { qq--; rr += bw; }
    else if (rr >= bw) { qq++; rr -= bw; }
    *q = qq; *r = rr;
}

// one warp walks `rows` rows of the triangle's box from row r0 on
__device__ __forceinline__ void walk_box(const TriSetup &t, int f, int r0, int rows, int lane, int W, float sx, float sy,
                                         unsigned long long *__restrict__ vis)
{
    const int bw = t.x1 - t.x0 + 1, total = bw * rows;
    const float inv_bw = 1.0f / (float)bw;
    for (int i = lane; i < total; i += 32) {
        int dr, dc;
        divmod_small(i, bw, inv_bw, &dr, &dc);
        const int row = r0 + dr, col = t.x0 + dc;
        float z;
        if (!pixel_test(t, row, col, sx, sy, &z)) continue;
        atomicMin(&vis[(size_t)row * W + col], ((unsigned long long)z_to_ord(z) << 32) | (unsigned int)f);
    }
}

// small boxes: one warp per triangle (its own kernel: 34 registers, so that a 10^6-face mesh runs at full occupancy)
__global__ void __launch_bounds__(256) raster_small_kernel(const TriSetup *__restrict__ setups, int F, int W, int H,
                                                           unsigned long long *__restrict__ vis)
{
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= F) return;
    TriSetup t;
    load_setup_warp(setups + f, t);
    if (!t.valid) return;
    const int bw = t.x1 - t.x0 + 1, bh = t.y1 - t.y0 + 1;
    if (bw <= 0 || bh <= 0 || (long long)bw * bh > RASTER_SMALL_MAX) return;
    walk_box(t, f, t.y0, bh, lane, W, 2.0f / (float)W, 2.0f / (float)H, vis);
}

// Rows r0, r0 + step, ... < r1 of a LARGE box, one warp: per row only the columns that can pass the five relaxed half-planes
//   a X + (b Y + c) >= 0   <=>   X >= -(b Y + c) / a  (a > 0)   or   X <= -(b Y + c) / a  (a < 0)
// -- a superset of the accepted pixels (+ 2 columns for the rounding of the reciprocal and of the pixel centre).
__device__ __forceinline__ void walk_spans(const TriSetup &t, int f, int r0, int r1, int step, int lane, int W, float sx, float sy,
                                           unsigned long long *__restrict__ vis)
{
    double hb[5], hc[5], hinv[5];          // hinv = -1 / a, 0 when a == 0
    int hsgn[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        HalfPlane h;
        if (k < 3) h = relaxed_plane(t.e[k][0], t.e[k][1], t.e[k][2], 0.f, 1.f);
        else if (k == 3) h = relaxed_plane(t.zA, t.zB, t.zC, 1.f, 1.f);
        else h = relaxed_plane(t.zA, t.zB, t.zC, 1.f, -1.f);
        hb[k] = h.b; hc[k] = h.c;
        hsgn[k] = h.a > 0.0 ? 1 : (h.a < 0.0 ? -1 : 0);
        hinv[k] = hsgn[k] ? -1.0 / h.a : 0.0;
    }
    for (int row = r0; row < r1; row += step) {
        const double Y = (double)(1.0f - ((float)row + 0.5f) * sy);
        double lo = -2.0, hi = 2.0;
        bool empty = false;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double r = hb[k] * Y + hc[k];
            const double x = r * hinv[k];
            if (hsgn[k] > 0) lo = fmax(lo, x);
            else if (hsgn[k] < 0) hi = fmin(hi, x);
            else if (r < 0.0) empty = true;         // (NaN coefficients fall through: whole box)
        }
        if (empty || lo > hi + 1e-6) continue;
        int c_lo = t.x0, c_hi = t.x1;
        if (lo > -1.5) c_lo = max(c_lo, (int)floor((lo + 1.0) * 0.5 * W - 0.5) - 2);
        if (hi < 1.5) c_hi = min(c_hi, (int)ceil((hi + 1.0) * 0.5 * W - 0.5) + 2);
        for (int col = c_lo + lane; col <= c_hi; col += 32) {
            float z;
            if (!pixel_test(t, row, col, sx, sy, &z)) continue;
            atomicMin(&vis[(size_t)row * W + col], ((unsigned long long)z_to_ord(z) << 32) | (unsigned int)f);
        }
    }
}

// Few triangles (F <= RASTER_DIRECT_MAX_F, e.g. the alpha-shape proxy of the first outer iteration): small AND medium boxes
// are walked by a (F, 8) grid of 128-thread CTAs, one row band each -- thousands of CTAs of a few pixels per thread hide
// the latency that a warp-per-item walk of so few items cannot; a large box is walked by the same CTAs row by row with spans.
// No lists, no second launch.
#define RASTER_DIRECT_MAX_F 4096
#define RASTER_DIRECT_BANDS 8
__global__ void __launch_bounds__(128) raster_direct_kernel(const TriSetup *__restrict__ setups, int W, int H, unsigned long long *__restrict__ vis)
{
    __shared__ TriSetup t;
    const int f = blockIdx.x;
    if (threadIdx.x < sizeof(TriSetup) / 4) ((int *)&t)[threadIdx.x] = ((const int *)&setups[f])[threadIdx.x];
    __syncthreads();
    if (!t.valid) return;
    const int bw = t.x1 - t.x0 + 1, bh = t.y1 - t.y0 + 1;
    if (bw <= 0 || bh <= 0) return;
    const int rows_per = (bh + gridDim.y - 1) / gridDim.y;
    const int r0 = t.y0 + blockIdx.y * rows_per, r1 = min(r0 + rows_per, t.y1 + 1);
    if (r0 >= r1) return;
    const float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    if ((long long)bw * bh > RASTER_MEDIUM_MAX) {
        walk_spans(t, f, r0 + (threadIdx.x >> 5), r1, blockDim.x >> 5, threadIdx.x & 31, W, sx, sy, vis);
        return;
    }
    const int total = bw * (r1 - r0);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int row = r0 + i / bw, col = t.x0 + i % bw;
        float z;
        if (!pixel_test(t, row, col, sx, sy, &z)) continue;
        atomicMin(&vis[(size_t)row * W + col], ((unsigned long long)z_to_ord(z) << 32) | (unsigned int)f);
    }
}

// One (triangle, row band) item per warp.  Per row the columns that can pass the five relaxed half-planes
//   a X + (b Y + c) >= 0   <=>   X >= -(b Y + c) / a  (a > 0)   or   X <= -(b Y + c) / a  (a < 0)
// are a superset of the accepted pixels (+ 2 columns for the rounding of the reciprocal and of the pixel centre).
// cnt[0], cnt[1]: list lengths (written by tri_setup_kernel), cnt[2]: CTAs that have finished; the last one to finish
// zeroes all three for the next call, so no memset has to precede tri_setup_kernel.
__global__ void __launch_bounds__(256) raster_big_kernel(const TriSetup *__restrict__ setups, int F, const int *__restrict__ list,
                                                         int *__restrict__ cnt, int W, int H, unsigned long long *__restrict__ vis)
{
    const int lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int bid = blockIdx.x, nbig = gridDim.x;
    const long long warp0 = (long long)bid * nwarp + (threadIdx.x >> 5), nwarps = (long long)nbig * nwarp;
    const long long n_med = (long long)cnt[0] * RASTER_BANDS_MEDIUM, items = n_med + (long long)cnt[1] * RASTER_BANDS_LARGE;
    const float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    __syncthreads();                                   // every thread has read the counts
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(cnt + 2, 1) == nbig - 1) { cnt[0] = 0; cnt[1] = 0; cnt[2] = 0; }
    }
    for (long long it = warp0; it < items; it += nwarps) {
        int f, band, nb;
        if (it < n_med) { f = list[it / RASTER_BANDS_MEDIUM]; band = (int)(it % RASTER_BANDS_MEDIUM); nb = RASTER_BANDS_MEDIUM; }
        else { const long long j = it - n_med; f = list[F + j / RASTER_BANDS_LARGE]; band = (int)(j % RASTER_BANDS_LARGE); nb = RASTER_BANDS_LARGE; }
        TriSetup t;
        load_setup_warp(setups + f, t);
        const int bh = t.y1 - t.y0 + 1;
        const int rows_per = (bh + nb - 1) / nb;
        const int r0 = t.y0 + band * rows_per, r1 = min(r0 + rows_per, t.y1 + 1);
        if (r0 >= r1) continue;
        if (nb == RASTER_BANDS_MEDIUM) {
            walk_box(t, f, r0, r1 - r0, lane, W, sx, sy, vis);      // medium box: plain walk of the band
            continue;
        }
        walk_spans(t, f, r0, r1, 1, lane, W, sx, sy, vis);
    }
}

int k_raster(mr_context *ctx, const Mat4 &P, unsigned long long *d_vis)
{
    MR_CUDA(ctx, cudaMemsetAsync(d_vis, 0xFF, ctx->N * sizeof(unsigned long long), ctx->stream));
    if (ctx->F == 0) return MR_OK;
    const int F = ctx->F;
    float *soup = mr_buf<float>(ctx, "soup", 0);
    TriSetup *setups = mr_buf<TriSetup>(ctx, "setup", 0);
    const bool fresh = !ctx->bufs.count("raster_lists") || ctx->bufs["raster_lists"].bytes < (2 * (size_t)F + 4) * sizeof(int);
    int *big = mr_buf<int>(ctx, "raster_lists", 2 * (size_t)F + 4);      // [0..2]: counters (self-resetting), [4..]: the two lists
    if (!big) return mr_fail(ctx, MR_ENOMEM, "raster_lists", "alloc");
    if (fresh) MR_CUDA(ctx, cudaMemsetAsync(big, 0, 4 * sizeof(int), ctx->stream));
    const bool direct = F <= RASTER_DIRECT_MAX_F;
    tri_setup_kernel<<<cdiv(F, 128), 128, 0, ctx->stream>>>(soup, F, P, ctx->W, ctx->H, setups, big + 4, big, direct ? 1 : 0);
    MR_LAUNCH_CHECK(ctx, "tri_setup_kernel");
    if (direct) {
        raster_direct_kernel<<<dim3(F, RASTER_DIRECT_BANDS), 128, 0, ctx->stream>>>(setups, ctx->W, ctx->H, d_vis);
        MR_LAUNCH_CHECK(ctx, "raster_direct_kernel");
        return MR_OK;
    }
    raster_small_kernel<<<cdiv(F, 8), 256, 0, ctx->stream>>>(setups, F, ctx->W, ctx->H, d_vis);
    MR_LAUNCH_CHECK(ctx, "raster_small_kernel");
    const int ctas = (int)std::min<long long>(((long long)F * RASTER_BANDS_LARGE + 7) / 8, 148 * 16);
    raster_big_kernel<<<ctas, 256, 0, ctx->stream>>>(setups, F, big + 4, big, ctx->W, ctx->H, d_vis);
    MR_LAUNCH_CHECK(ctx, "raster_big_kernel");
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

// The same queries WITHOUT rendering the maps ("200 depth() renders -> 200 single-pixel ray queries"):
// one thread per (viewer, triangle) builds the triangle's setup for that viewer and runs the rasteriser's own pixel test
// on the viewer's n query pixels only; the nearest hit per query is an atomicMin on the same 64-bit key.  Same setup,
// same pixel test, same tie rule as k_raster, so the answers are bit-identical to indexing the full maps
// (tests/test_gpu_edge_and_fullsize.py), at F x n tests per viewer instead of a W x H raster.
#define QUERY_CHUNK 256
__global__ void __launch_bounds__(128) depth_query_kernel(const float *__restrict__ soup, int F, const float *__restrict__ cams, int W, int H,
                                                          const int32_t *__restrict__ rows, const int32_t *__restrict__ cols, int n,
                                                          unsigned long long *__restrict__ keys)
{
    __shared__ Mat4 P;
    __shared__ int2 q[QUERY_CHUNK];
    const int cam = blockIdx.y;
    if (threadIdx.x < 16) P.m[threadIdx.x] = cams[16 * (size_t)cam + threadIdx.x];
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    TriSetup t;
    t.valid = 0;
    if (f < F) compute_setup(soup + 9 * (size_t)f, P, W, H, t);
    const float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    const long long last = (long long)W * H - 1;
    for (int base = 0; base < n; base += QUERY_CHUNK) {
        __syncthreads();
        for (int j = threadIdx.x; j < QUERY_CHUNK; j += blockDim.x) {
            int2 v = make_int2(-1, -1);
            if (base + j < n) {
                const int r = rows[(size_t)cam * n + base + j], c = cols[(size_t)cam * n + base + j];
                if (r >= 0 && r < H && c >= 0 && c <= W) {
                    long long idx = (long long)r * W + c;          // depth.at<float>(row, col) on a continuous Mat
                    if (idx > last) idx = last;
                    v = make_int2((int)(idx / W), (int)(idx % W));
                }
            }
            q[j] = v;
        }
        __syncthreads();
        if (!t.valid) continue;
        const int m = min(QUERY_CHUNK, n - base);
        for (int j = 0; j < m; j++) {
            const int2 v = q[j];
            if (v.x < t.y0 || v.x > t.y1 || v.y < t.x0 || v.y > t.x1) continue;
            float z;
            if (!pixel_test(t, v.x, v.y, sx, sy, &z)) continue;
            atomicMin(&keys[(size_t)cam * n + base + j], ((unsigned long long)z_to_ord(z) << 32) | (unsigned int)f);
        }
    }
}

__global__ void resolve_query_kernel(const unsigned long long *__restrict__ keys, size_t total, float *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    unsigned long long k = keys[i];
    out[i] = (k == BG_KEY) ? MR_BACKGROUND_DEPTH : ord_to_z((unsigned int)(k >> 32));
}

int k_depth_query(mr_context *ctx, const float *d_cams, int n_cameras, const int32_t *d_rows, const int32_t *d_cols, int n, float *d_out)
{
    const size_t total = (size_t)n_cameras * n;
    if (total == 0) return MR_OK;
    unsigned long long *keys = mr_buf<unsigned long long>(ctx, "q_keys", total);
    if (!keys) return mr_fail(ctx, MR_ENOMEM, "q_keys", "alloc");
    MR_CUDA(ctx, cudaMemsetAsync(keys, 0xFF, total * sizeof(unsigned long long), ctx->stream));
    if (ctx->F > 0) {
        float *soup = mr_buf<float>(ctx, "soup", 0);
        for (int c0 = 0; c0 < n_cameras; c0 += 32768) {            // gridDim.y limit
            const int nc = std::min(32768, n_cameras - c0);
            dim3 grid(cdiv(ctx->F, 128), nc);
            depth_query_kernel<<<grid, 128, 0, ctx->stream>>>(soup, ctx->F, d_cams + 16 * (size_t)c0, ctx->W, ctx->H, d_rows + (size_t)c0 * n,
                                                              d_cols + (size_t)c0 * n, n, keys + (size_t)c0 * n);
            MR_LAUNCH_CHECK(ctx, "depth_query_kernel");
        }
    }
    resolve_query_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(keys, total, d_out);
    MR_LAUNCH_CHECK(ctx, "resolve_query_kernel");
    return MR_OK;
}

int k_depth_samples(mr_context *ctx, const unsigned long long *d_vis, const int32_t *d_rows, const int32_t *d_cols, int n, float *d_out)
{
    if (n <= 0) return MR_OK;
    depth_samples_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(d_vis, ctx->W, ctx->H, d_rows, d_cols, n, d_out);
    MR_LAUNCH_CHECK(ctx, "depth_samples_kernel");
    return MR_OK;
}
