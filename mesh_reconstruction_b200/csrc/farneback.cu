// farneback.cu -- the `-f` branch of calculateFlow (flow.cpp:22-26): OpenCV's FarnebackOpticalFlow with the
// reference's parameters
//     levels 10, pyr_scale 0.8, fastPyramids false, winsize (H+W)/100, iterations 7,
//     poly_n (sigma < 1.5 ? 5 : 7), poly_sigma (H+W)/1000, flags 0 (box filter)
// restated stage by stage from OpenCV's published algorithm (modules/video/src/optflowgf.cpp: polynomial
// expansion, UpdateMatrices, UpdateFlow_Blur, per-level Gaussian blur + bilinear resize).  The NumPy
// restatement in oracle/farneback_np.py agrees with the cv2 binary to ~2e-6 px; the kernels here follow the
// same operation order (float where OpenCV uses float, double where it uses double).  Box sums are taken
// in double like OpenCV's (vertical: sliding per column; horizontal: 2m+1 taps from shared memory), so the
// only differences are summation-order roundings far below the 0.01 px parity bound.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace {

constexpr int FB_MAXK = 64;   // max (half) kernel taps uploaded per launch

struct KernF {
    float k[2 * FB_MAXK + 1];
    int r;
};

__device__ __forceinline__ int refl101f(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// cv::GaussianBlur on CV_32F (separable, BORDER_REFLECT_101): row filter then column filter,
// symmetric accumulation  k0*c + sum_i k_i*(a_i + b_i).
template <class T>
__global__ void __launch_bounds__(256) fb_blur_h_kernel(const T *__restrict__ src, int W, int H, KernF kf, float *__restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const T *row = src + (size_t)y * W;
    float acc = (float)row[x] * kf.k[kf.r];
    for (int i = 1; i <= kf.r; i++) acc = acc + ((float)row[refl101f(x - i, W)] + (float)row[refl101f(x + i, W)]) * kf.k[kf.r + i];
    dst[(size_t)y * W + x] = acc;
}
__global__ void __launch_bounds__(256) fb_blur_v_kernel(const float *__restrict__ src, int W, int H, KernF kf, float *__restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float acc = src[(size_t)y * W + x] * kf.k[kf.r];
    for (int i = 1; i <= kf.r; i++)
        acc = acc + (src[(size_t)refl101f(y - i, H) * W + x] + src[(size_t)refl101f(y + i, H) * W + x]) * kf.k[kf.r + i];
    dst[(size_t)y * W + x] = acc;
}

// cv::resize(INTER_LINEAR) on CV_32FC<C>; optional multiplication of the result (flow *= 1/pyr_scale)
__device__ __forceinline__ void lin_coeff(int d, double scale, int n_src, int &s, float &f)
{
    f = (float)((d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
}
template <int C>
__global__ void __launch_bounds__(256) fb_resize_kernel(const float *__restrict__ src, int w, int h, float *__restrict__ dst, int W, int H,
                                                        double sx, double sy, float mul)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int x0, y0;
    float fx, fy;
    lin_coeff(x, sx, w, x0, fx);
    lin_coeff(y, sy, h, y0, fy);
    int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
#pragma unroll
    for (int c = 0; c < C; c++) {
        float a = src[((size_t)y0 * w + x0) * C + c] * (1.f - fx) + src[((size_t)y0 * w + x1) * C + c] * fx;
        float b = src[((size_t)y1 * w + x0) * C + c] * (1.f - fx) + src[((size_t)y1 * w + x1) * C + c] * fx;
        float v = a * (1.f - fy) + b * fy;
        dst[((size_t)y * W + x) * C + c] = (mul != 1.f) ? v * mul : v;
    }
}

// FarnebackPolyExp: vertical pass -> 3 floats per pixel, horizontal pass (double accumulators) -> 5 floats.
struct PolyK {
    float g[FB_MAXK + 1], xg[FB_MAXK + 1], xxg[FB_MAXK + 1];   // index k = 0..n
    double ig11, ig03, ig33, ig55;
    int n;
};
__global__ void __launch_bounds__(256) fb_poly_v_kernel(const float *__restrict__ src, int W, int H, PolyK pk, float *__restrict__ row3)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float t0 = src[(size_t)y * W + x] * pk.g[0], t1 = 0.f, t2 = 0.f;
    for (int k = 1; k <= pk.n; k++) {
        float s0 = src[(size_t)max(y - k, 0) * W + x], s1 = src[(size_t)min(y + k, H - 1) * W + x];
        float p = s0 + s1;
        t0 = t0 + pk.g[k] * p;
        t1 = t1 + pk.xg[k] * (s1 - s0);
        t2 = t2 + pk.xxg[k] * p;
    }
    float *o = row3 + ((size_t)y * W + x) * 3;
    o[0] = t0; o[1] = t1; o[2] = t2;
}
__global__ void __launch_bounds__(256) fb_poly_h_kernel(const float *__restrict__ row3, int W, int H, PolyK pk, float *__restrict__ R)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float *row = row3 + (size_t)y * W * 3;
    auto at = [&](int xx, int c) { return row[min(max(xx, 0), W - 1) * 3 + c]; };   // replicated borders
    // mixed precision exactly as in OpenCV: products of two floats are float products, products with
    // the double `tg` are double products; the accumulators are double
    float g0 = pk.g[0];
    double b1 = at(x, 0) * g0, b2 = 0, b3 = at(x, 1) * g0, b4 = 0, b5 = at(x, 2) * g0, b6 = 0;
    for (int k = 1; k <= pk.n; k++) {
        float p0 = at(x + k, 0), m0 = at(x - k, 0), p1 = at(x + k, 1), m1 = at(x - k, 1), p2 = at(x + k, 2), m2 = at(x - k, 2);
        double tg = (double)(p0 + m0);
        g0 = pk.g[k];
        b1 += tg * (double)g0;
        b4 += tg * (double)pk.xxg[k];
        b2 += (double)((p0 - m0) * pk.xg[k]);
        b3 += (double)((p1 + m1) * g0);
        b6 += (double)((p1 - m1) * pk.xg[k]);
        b5 += (double)((p2 + m2) * g0);
    }
    float *d = R + ((size_t)y * W + x) * 5;
    d[1] = (float)(b2 * pk.ig11);
    d[0] = (float)(b3 * pk.ig11);
    d[3] = (float)(b1 * pk.ig03 + b4 * pk.ig33);
    d[2] = (float)(b1 * pk.ig03 + b5 * pk.ig33);
    d[4] = (float)(b6 * pk.ig55);
}

// FarnebackUpdateMatrices
__global__ void __launch_bounds__(256) fb_update_matrices_kernel(const float *__restrict__ R0, const float *__restrict__ R1,
                                                                 const float *__restrict__ flow, int W, int H, float *__restrict__ M)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    size_t i = (size_t)y * W + x;
    float dx = flow[i * 2], dy = flow[i * 2 + 1];
    float fx = (float)x + dx, fy = (float)y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    const float *r0 = R0 + i * 5;
    float r2, r3, r4, r5, r6;
    fx -= (float)x1;
    fy -= (float)y1;
    if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
        const float *p = R1 + ((size_t)y1 * W + x1) * 5;
        const size_t st = (size_t)W * 5;
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * p[0] + a01 * p[5] + a10 * p[st] + a11 * p[st + 5];
        r3 = a00 * p[1] + a01 * p[6] + a10 * p[st + 1] + a11 * p[st + 6];
        r4 = a00 * p[2] + a01 * p[7] + a10 * p[st + 2] + a11 * p[st + 7];
        r5 = a00 * p[3] + a01 * p[8] + a10 * p[st + 3] + a11 * p[st + 8];
        r6 = a00 * p[4] + a01 * p[9] + a10 * p[st + 4] + a11 * p[st + 9];
        r4 = (r0[2] + r4) * 0.5f;
        r5 = (r0[3] + r5) * 0.5f;
        r6 = (r0[4] + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = r0[2];
        r5 = r0[3];
        r6 = r0[4] * 0.5f;
    }
    r2 = (r0[0] - r2) * 0.5f;
    r3 = (r0[1] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(W - 10) || (unsigned)(y - 5) >= (unsigned)(H - 10)) {
        float scale = (x < 5 ? border[x] : 1.f) * (x >= W - 5 ? border[W - x - 1] : 1.f) * (y < 5 ? border[y] : 1.f) *
                      (y >= H - 5 ? border[H - y - 1] : 1.f);
        r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
    }
    float *m = M + i * 5;
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

// FarnebackUpdateFlow_Blur, vertical part: box sums over rows y-m .. y+m (rows clamped) in double,
// one thread per (column, channel, 32-row segment) sliding down its segment like OpenCV's vsum.
#define FB_VSEG 32
__global__ void __launch_bounds__(128) fb_box_v_kernel(const float *__restrict__ M, int W, int H, int m, double *__restrict__ vs)
{
    int xc = blockIdx.x * blockDim.x + threadIdx.x;   // x*5 + c
    if (xc >= W * 5) return;
    const size_t st = (size_t)W * 5;
    const int y0 = blockIdx.y * FB_VSEG, y1 = min(y0 + FB_VSEG, H);
    // window of row y0: rows y0-m .. y0+m, clamped (for y0 = 0 this is OpenCV's (m+2)*row0 + ... initialisation
    // followed by its first slide); then one add and one subtract per row like OpenCV's vsum
    double s = 0;
    for (int j = -m; j <= m; j++) s += (double)M[(size_t)min(max(y0 + j, 0), H - 1) * st + xc];
    vs[(size_t)y0 * st + xc] = s;
    for (int y = y0 + 1; y < y1; y++) {
        s += (double)M[(size_t)min(y + m, H - 1) * st + xc] - (double)M[(size_t)max(y - m - 1, 0) * st + xc];
        vs[(size_t)y * st + xc] = s;
    }
}
// horizontal part + 2x2 solve: one block per row segment of FB_SEG * FB_RUN pixels staged in shared memory; every thread
// owns FB_RUN consecutive pixels: 2m+1 taps for the first, then one add and one subtract per pixel (OpenCV's own
// horizontal pass slides the same way; sums in double)
#define FB_SEG 128
#define FB_RUN 4
__global__ void __launch_bounds__(FB_SEG) fb_box_h_solve_kernel(const double *__restrict__ vs, int W, int H, int m, int block_size,
                                                                float *__restrict__ flow)
{
    extern __shared__ double seg[];   // (FB_SEG * FB_RUN + 2m) * 5
    const int y = blockIdx.y, x0 = blockIdx.x * FB_SEG * FB_RUN;
    const double *row = vs + (size_t)y * W * 5;
    const int n = (FB_SEG * FB_RUN + 2 * m) * 5;
    for (int i = threadIdx.x; i < n; i += FB_SEG) {
        int xx = x0 - m + i / 5, c = i % 5;
        xx = min(max(xx, 0), W - 1);                                                         // replicated borders
        seg[i] = row[xx * 5 + c];
    }
    __syncthreads();
    const int xl = threadIdx.x * FB_RUN;                 // first pixel of the run, segment-local
    if (x0 + xl >= W) return;
    double g[5] = {0, 0, 0, 0, 0};
    const double *p = seg + xl * 5;
    for (int t = 0; t <= 2 * m; t++) {
#pragma unroll
        for (int c = 0; c < 5; c++) g[c] += p[t * 5 + c];
    }
    const double scale = 1. / ((double)block_size * block_size);
#pragma unroll
    for (int j = 0; j < FB_RUN; j++) {
        const int x = x0 + xl + j;
        if (x >= W) break;
        if (j > 0) {
#pragma unroll
            for (int c = 0; c < 5; c++) g[c] += p[(j + 2 * m) * 5 + c] - p[(j - 1) * 5 + c];
        }
        double g11_ = g[0] * scale, g12_ = g[1] * scale, g22_ = g[2] * scale, h1_ = g[3] * scale, h2_ = g[4] * scale;
        double idet = 1. / (g11_ * g22_ - g12_ * g12_ + 1e-3);
        size_t i = (size_t)y * W + x;
        flow[i * 2] = (float)((g11_ * h2_ - g12_ * h1_) * idet);
        flow[i * 2 + 1] = (float)((g22_ * h1_ - g12_ * h2_) * idet);
    }
}

__global__ void fb_pack_kernel(const float *__restrict__ flow2, size_t N, float4 *__restrict__ flow4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flow4[i] = make_float4(flow2[2 * i], flow2[2 * i + 1], 0.f, 0.f);
}

// ---- host-side constants ------------------------------------------------------------------------------------
int cv_round(double v) { return (int)std::nearbyint(v); }

// cv::getGaussianKernel(n, sigma, CV_32F)
void gaussian_kernel(int n, double sigma, KernF &kf)
{
    kf.r = n / 2;
    static const float small3[] = {0.25f, 0.5f, 0.25f};
    if (n == 3 && sigma <= 0) {
        for (int i = 0; i < 3; i++) kf.k[i] = small3[i];
        return;
    }
    double sigmaX = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2X = -0.5 / (sigmaX * sigmaX), sum = 0;
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        kf.k[i] = (float)std::exp(scale2X * x * x);
        sum += kf.k[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) kf.k[i] = (float)(kf.k[i] * sum);
}

// FarnebackPrepareGaussian
void prepare_poly(int n, double sigma, PolyK &pk)
{
    if (sigma < 1.1920929e-07) sigma = n * 0.3;
    std::vector<float> g(2 * n + 1), xg(2 * n + 1), xxg(2 * n + 1);
    double s = 0;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)std::exp(-x * x / (2 * sigma * sigma));
        s += g[x + n];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            G[0][0] += g[y + n] * g[x + n];
            G[1][1] += g[y + n] * g[x + n] * x * x;
            G[3][3] += g[y + n] * g[x + n] * x * x * x * x;
            G[5][5] += g[y + n] * g[x + n] * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    // inverse by Gauss-Jordan with partial pivoting (OpenCV: Cholesky; the matrix is SPD and tiny)
    double A[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 12; j++) A[i][j] = j < 6 ? G[i][j] : (j - 6 == i ? 1. : 0.);
    for (int c = 0; c < 6; c++) {
        int piv = c;
        for (int r = c + 1; r < 6; r++)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        for (int j = 0; j < 12; j++) std::swap(A[c][j], A[piv][j]);
        double d = 1. / A[c][c];
        for (int j = 0; j < 12; j++) A[c][j] *= d;
        for (int r = 0; r < 6; r++)
            if (r != c) {
                double f = A[r][c];
                for (int j = 0; j < 12; j++) A[r][j] -= f * A[c][j];
            }
    }
    pk.n = n;
    for (int k = 0; k <= n; k++) { pk.g[k] = g[n + k]; pk.xg[k] = xg[n + k]; pk.xxg[k] = xxg[n + k]; }
    pk.ig11 = A[1][7]; pk.ig03 = A[0][9]; pk.ig33 = A[3][9]; pk.ig55 = A[5][11];
}

}  // namespace

// calculateFlow(prev, next, use_farneback = true), flow part: writes (u, v, 0, 0) into d_flow4.
int k_farneback(mr_context *ctx, const uint8_t *d_prev, const uint8_t *d_next, float *d_flow4)
{
    const int W0 = ctx->W, H0 = ctx->H;
    const size_t N = ctx->N;
    const double pyr_scale = 0.8, poly_sigma = (H0 + W0) / 1000.0;
    const int winsize = (H0 + W0) / 100, iterations = 7, poly_n = poly_sigma < 1.5 ? 5 : 7, min_size = 32;
    int levels = 10, k;
    double scale = 1;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (W0 * scale < min_size || H0 * scale < min_size) break;
    }
    levels = k;
    if (winsize / 2 > FB_MAXK || poly_n > FB_MAXK) return mr_fail(ctx, MR_EINVAL, "k_farneback", "window too large");
    float *fimg = mr_buf<float>(ctx, "fb_fimg", N), *tmp = mr_buf<float>(ctx, "fb_tmp", N), *I = mr_buf<float>(ctx, "fb_I", N);
    float *row3 = mr_buf<float>(ctx, "fb_row3", N * 3), *R0 = mr_buf<float>(ctx, "fb_R0", N * 5), *R1 = mr_buf<float>(ctx, "fb_R1", N * 5);
    float *M = mr_buf<float>(ctx, "fb_M", N * 5), *fA = mr_buf<float>(ctx, "fb_flowA", N * 2), *fB = mr_buf<float>(ctx, "fb_flowB", N * 2);
    double *vs = mr_buf<double>(ctx, "fb_vs", N * 5);
    if (!fimg || !tmp || !I || !row3 || !R0 || !R1 || !M || !fA || !fB || !vs) return mr_fail(ctx, MR_ENOMEM, "k_farneback", "alloc");
    PolyK pk;
    prepare_poly(poly_n, poly_sigma, pk);
    float *flow = fA, *prev_flow = nullptr;
    int pw = 0, ph = 0;
    for (k = levels; k >= 0; k--) {
        scale = 1;
        for (int i = 0; i < k; i++) scale *= pyr_scale;
        double sigma = (1. / scale - 1) * 0.5;
        int smooth_sz = cv_round(sigma * 5) | 1;
        if (smooth_sz < 3) smooth_sz = 3;
        if (smooth_sz / 2 > FB_MAXK) return mr_fail(ctx, MR_EINVAL, "k_farneback", "blur kernel too large");
        const int w = cv_round(W0 * scale), h = cv_round(H0 * scale);
        flow = (prev_flow == fA) ? fB : fA;
        dim3 b(256), g(cdiv(w, 256), h), g0(cdiv(W0, 256), H0);
        if (!prev_flow) MR_CUDA(ctx, cudaMemsetAsync(flow, 0, (size_t)w * h * 2 * sizeof(float), ctx->stream));
        else {
            fb_resize_kernel<2><<<g, b, 0, ctx->stream>>>(prev_flow, pw, ph, flow, w, h, (double)pw / w, (double)ph / h, (float)(1. / pyr_scale));
            MR_LAUNCH_CHECK(ctx, "fb_resize_kernel");
        }
        KernF kf;
        gaussian_kernel(smooth_sz, sigma, kf);
        for (int i = 0; i < 2; i++) {
            fb_blur_h_kernel<uint8_t><<<g0, b, 0, ctx->stream>>>(i == 0 ? d_prev : d_next, W0, H0, kf, tmp);
            MR_LAUNCH_CHECK(ctx, "fb_blur_h_kernel");
            fb_blur_v_kernel<<<g0, b, 0, ctx->stream>>>(tmp, W0, H0, kf, fimg);
            MR_LAUNCH_CHECK(ctx, "fb_blur_v_kernel");
            fb_resize_kernel<1><<<g, b, 0, ctx->stream>>>(fimg, W0, H0, I, w, h, (double)W0 / w, (double)H0 / h, 1.f);
            MR_LAUNCH_CHECK(ctx, "fb_resize_kernel");
            fb_poly_v_kernel<<<g, b, 0, ctx->stream>>>(I, w, h, pk, row3);
            MR_LAUNCH_CHECK(ctx, "fb_poly_v_kernel");
            fb_poly_h_kernel<<<g, b, 0, ctx->stream>>>(row3, w, h, pk, i == 0 ? R0 : R1);
            MR_LAUNCH_CHECK(ctx, "fb_poly_h_kernel");
        }
        fb_update_matrices_kernel<<<g, b, 0, ctx->stream>>>(R0, R1, flow, w, h, M);
        MR_LAUNCH_CHECK(ctx, "fb_update_matrices_kernel");
        const int m = winsize / 2;
        for (int it = 0; it < iterations; it++) {
            fb_box_v_kernel<<<dim3(cdiv(w * 5, 128), cdiv(h, FB_VSEG)), 128, 0, ctx->stream>>>(M, w, h, m, vs);
            MR_LAUNCH_CHECK(ctx, "fb_box_v_kernel");
            fb_box_h_solve_kernel<<<dim3(cdiv(w, FB_SEG * FB_RUN), h), FB_SEG, (FB_SEG * FB_RUN + 2 * m) * 5 * sizeof(double), ctx->stream>>>(vs, w, h, m, winsize, flow);
            MR_LAUNCH_CHECK(ctx, "fb_box_h_solve_kernel");
            if (it < iterations - 1) {
                fb_update_matrices_kernel<<<g, b, 0, ctx->stream>>>(R0, R1, flow, w, h, M);
                MR_LAUNCH_CHECK(ctx, "fb_update_matrices_kernel");
            }
        }
        prev_flow = flow;
        pw = w;
        ph = h;
    }
    fb_pack_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(flow, N, (float4 *)d_flow4);
    MR_LAUNCH_CHECK(ctx, "fb_pack_kernel");
    return MR_OK;
}
