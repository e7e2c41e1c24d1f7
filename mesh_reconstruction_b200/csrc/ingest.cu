// ingest.cu -- frame ingest of Configuration::Configuration (configuration.cpp:226-245), the step right before the hot
// path: every decoded BGR frame is (optionally) shrunk to the render size with
// cv::resize(..., CV_INTER_AREA) and turned into the gray frame the path works on (cv::cvtColor CV_BGR2GRAY).  Video
// DECODING stays on the host (cv::VideoCapture / FFmpeg, no NVDEC library in this image); what moves to the GPU is the
// per-pixel work, so that a decoded frame crosses PCIe once (3 B/px, or 3 f^2 B per output pixel when shrinking) and
// the gray frame is born in HBM, where mr_process_main_frame wants it.
//
// Arithmetic = OpenCV's, checked against the cv2 binary (tests/test_gpu_ingest.py):
//   * INTER_AREA with integer factors fx, fy is OpenCV's resizeAreaFast_: integer box sum, then  fx == fy == 2: (sum + 2) >> 2 ;
//     otherwise saturate_cast<uchar>(sum * (1.f / (fx * fy))) (float product, round half to even).
//   * INTER_AREA with any other shrink factor (the reference only warns when the frame size is not divisible by -s,
//     configuration.cpp:149-151; e.g. 1920x1080 with -s 1.5) is OpenCV's general ResizeArea_Invoker<uchar, float>: per axis a
//     table of (source index, float weight) -- a partial first cell, whole cells of weight (float)(1 / cellWidth), a partial
//     last cell, computed in double from scale = ssize / dsize -- then per source row  buf = sum_k S[k] * alpha_k  (float, table
//     order) and  sum = beta * buf  /  sum += beta * buf  over the rows, saturate_cast<uchar> (round half to even) at the end.
//     One thread per output pixel rebuilds its table entries on the fly with the same double arithmetic.
//   * BGR2GRAY 8U: (B * BY + G * GY + R * RY + half) >> shift with the 15-bit coefficients (3735, 19235, 9798) of
//     OpenCV >= 3.4.6 / 4.x (the cv2 binary the oracle is pinned to) or the 14-bit ones (1868, 9617, 4899) of
//     OpenCV 3.0 - 3.4.5 (mr_set_gray_shift).
//   * with estimateExposure on (configuration.cpp:417-425) the gray conversion is replaced by the exposure mix
//         frame = zeros(8UC1);  for c in B, G, R:  frame += channel[c] * exposure[c]
//     which cv::Mat evaluates as  t_c = saturate_cast<uchar>(rint((float)channel * (float)exposure[c]))  (convertTo with a
//     float scale, round half to even) followed by SATURATING 8-bit additions, in channel order.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r, int shift)
{
    if (shift == 15) return (b * 3735u + g * 19235u + r * 9798u + (1u << 14)) >> 15;
    return (b * 1868u + g * 9617u + r * 4899u + (1u << 13)) >> 14;
}

struct Expo {
    float e[3];
    int on;
};

__device__ __forceinline__ unsigned expo_of(unsigned b, unsigned g, unsigned r, const Expo &x)
{
    const unsigned ch[3] = {b, g, r};
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float v = (float)ch[c] * x.e[c];
        int t = 0;
        // saturate_cast<uchar>(cvRound(v)): cvRound of NaN or of |v| >= 2^31 is the integer indefinite INT_MIN -> 0
        if (!(v < 2147483648.f)) t = 0;
        else if (v >= 255.5f) t = 255;
        else if (v > 0.f) t = __float2int_rn(v);
        acc = min(acc + (unsigned)t, 255u);                // cv::add on CV_8U saturates
    }
    return acc;
}

__device__ __forceinline__ unsigned out_of(unsigned b, unsigned g, unsigned r, int shift, const Expo &x)
{
    return x.on ? expo_of(b, g, r, x) : gray_of(b, g, r, shift);
}

// factor 1: four pixels per thread (12 bytes in as three 32-bit words when aligned, one 32-bit word out)
__global__ void __launch_bounds__(256) gray_kernel(const uint8_t *__restrict__ bgr, size_t n, int shift, Expo ex, uint8_t *__restrict__ out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
    const size_t p0 = q * 4;
    if (p0 >= n) return;
    if (p0 + 4 <= n && ((uintptr_t)bgr & 3) == 0 && ((uintptr_t)out & 3) == 0) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(bgr) + q * 3;
        const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
        // bytes: a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3
        const unsigned g0 = out_of(a & 255, (a >> 8) & 255, (a >> 16) & 255, shift, ex);
        const unsigned g1 = out_of(a >> 24, b & 255, (b >> 8) & 255, shift, ex);
        const unsigned g2 = out_of((b >> 16) & 255, b >> 24, c & 255, shift, ex);
        const unsigned g3 = out_of((c >> 8) & 255, (c >> 16) & 255, c >> 24, shift, ex);
        reinterpret_cast<uint32_t *>(out)[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        return;
    }
    for (size_t p = p0; p < n && p < p0 + 4; p++) out[p] = (uint8_t)out_of(bgr[3 * p], bgr[3 * p + 1], bgr[3 * p + 2], shift, ex);
}

// factor f >= 2: one thread per output pixel, box sums of the three channels, OpenCV's rounding, then gray
__global__ void __launch_bounds__(256) area_gray_kernel(const uint8_t *__restrict__ bgr, int sw, int W, int H, int f, int fy, int shift, Expo ex, uint8_t *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    unsigned s0 = 0, s1 = 0, s2 = 0;
    for (int dy = 0; dy < fy; dy++) {
        const uint8_t *row = bgr + ((size_t)(y * fy + dy) * sw + (size_t)x * f) * 3;
        for (int dx = 0; dx < f; dx++) { s0 += row[3 * dx]; s1 += row[3 * dx + 1]; s2 += row[3 * dx + 2]; }
    }
    unsigned b, g, r;
    if (f == 2 && fy == 2) { b = (s0 + 2) >> 2; g = (s1 + 2) >> 2; r = (s2 + 2) >> 2; }
    else {
        const float scale = 1.f / (float)(f * fy);
        b = (unsigned)__float2int_rn((float)s0 * scale); g = (unsigned)__float2int_rn((float)s1 * scale); r = (unsigned)__float2int_rn((float)s2 * scale);
        b = min(b, 255u); g = min(g, 255u); r = min(r, 255u);
    }
    out[(size_t)y * W + x] = (uint8_t)out_of(b, g, r, shift, ex);
}

// one axis of OpenCV's computeResizeAreaTab for destination index d: source range and the weights of its first and last
// partial cells (full cells weigh `full`); cells are visited as  [first partial] sx1 .. sx2-1 [last partial]
struct AreaCell {
    int sx1, sx2;          // whole cells: sx1 <= sx < sx2
    float a_first, a_full, a_last;
    bool has_first, has_last;
};
__device__ __forceinline__ AreaCell area_cell(int d, int ssize, double scale)
{
    AreaCell c;
    const double fsx1 = d * scale, fsx2 = fsx1 + scale;
    const double cell = fmin(scale, (double)ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = min(sx2, ssize - 1);
    sx1 = min(sx1, sx2);
    c.sx1 = sx1; c.sx2 = sx2;
    c.has_first = sx1 - fsx1 > 1e-3;
    c.a_first = (float)((sx1 - fsx1) / cell);
    c.a_full = (float)(1.0 / cell);
    c.has_last = fsx2 - sx2 > 1e-3;
    c.a_last = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
    return c;
}

__global__ void __launch_bounds__(256) area_general_kernel(const uint8_t *__restrict__ bgr, int sw, int sh, int W, int H, double scale_x,
                                                           double scale_y, int shift, Expo ex, uint8_t *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const AreaCell cx = area_cell(x, sw, scale_x), cy = area_cell(y, sh, scale_y);
    float sum[3] = {0.f, 0.f, 0.f};
    bool first_row = true;
    auto add_row = [&](int sy, float beta) {
        const uint8_t *row = bgr + (size_t)sy * sw * 3;
        float buf[3] = {0.f, 0.f, 0.f};
        auto tap = [&](int sx, float a) {
#pragma unroll
            for (int c = 0; c < 3; c++) buf[c] = buf[c] + (float)row[3 * sx + c] * a;
        };
        if (cx.has_first) tap(cx.sx1 - 1, cx.a_first);
        for (int sx = cx.sx1; sx < cx.sx2; sx++) tap(sx, cx.a_full);
        if (cx.has_last) tap(cx.sx2, cx.a_last);
#pragma unroll
        for (int c = 0; c < 3; c++) sum[c] = first_row ? beta * buf[c] : sum[c] + beta * buf[c];
        first_row = false;
    };
    if (cy.has_first) add_row(cy.sx1 - 1, cy.a_first);
    for (int sy = cy.sx1; sy < cy.sx2; sy++) add_row(sy, cy.a_full);
    if (cy.has_last) add_row(cy.sx2, cy.a_last);
    unsigned v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = (unsigned)min(max(__float2int_rn(sum[c]), 0), 255);
    out[(size_t)y * W + x] = (uint8_t)out_of(v[0], v[1], v[2], shift, ex);
}

}  // namespace

int k_ingest(mr_context *ctx, const uint8_t *d_bgr, int src_w, int src_h, uint8_t *d_gray, const float *exposure)
{
    Expo ex;
    ex.on = exposure ? 1 : 0;
    for (int c = 0; c < 3; c++) ex.e[c] = exposure ? exposure[c] : 0.f;
    const int W = ctx->W, H = ctx->H;
    if (src_w == W && src_h == H) {
        const size_t groups = (ctx->N + 3) / 4;
        gray_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, ctx->stream>>>(d_bgr, ctx->N, ctx->gray_shift, ex, d_gray);
        MR_LAUNCH_CHECK(ctx, "gray_kernel");
        return MR_OK;
    }
    const int f = src_w / W;
    dim3 b(32, 8), g(cdiv(W, 32), cdiv(H, 8));
    if (src_w % W != 0 || src_h % H != 0) {                 // integer factors (possibly different in x and y) take OpenCV's fast path
        area_general_kernel<<<g, b, 0, ctx->stream>>>(d_bgr, src_w, src_h, W, H, (double)src_w / W, (double)src_h / H, ctx->gray_shift, ex, d_gray);
        MR_LAUNCH_CHECK(ctx, "area_general_kernel");
        return MR_OK;
    }
    area_gray_kernel<<<g, b, 0, ctx->stream>>>(d_bgr, src_w, W, H, f, src_h / H, ctx->gray_shift, ex, d_gray);
    MR_LAUNCH_CHECK(ctx, "area_gray_kernel");
    return MR_OK;
}
