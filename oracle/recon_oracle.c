/*
 * recon_oracle.c -- CPU restatement (plain C) of the reference-owned arithmetic on
 * the dense-correspondence hot path of addam/mesh-reconstruction.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mesh_reconstruction_b200/ may link,
 * load or call this file; it is the checker for the CUDA path (tests/, smoke(),
 * bench.py's cpu_baseline / --impl reference legs).
 *
 * Every function cites the reference file:line it follows (paths into the
 * upstream tree).  cv::Mat arithmetic is restated with the exact evaluation rules
 * of OpenCV (verified against cv2 4.13 in tests/test_oracle_cv.py):
 *   - Mat*Mat (cv::gemm): 2<=K<=4 and K==rows or K==cols of the result, flags==0
 *         -> sequential float32 accumulation; otherwise double accumulation;
 *   - Mat /= s, Mat / s, expr / s  -> multiply by (float)(1.0/(double)s);
 *   - Mat::dot, cv::determinant(2x2), cv::norm -> double;
 *   - Mat::inv(): 4x4 float LU with partial pivoting; 2x2 closed form with the
 *         determinant in double and its reciprocal cast to float.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction), see oracle/Makefile.
 *
 * Parity status: the reference's own tests pin no numeric result on this path
 * ("parity unpinned" by upstream tests); the OpenCV-owned
 * pieces are pinned against the cv2 binary, the GL semantics are DEFINED here
 * (quirk table in DESIGN.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BACKGROUND_DEPTH 1.0f /* recon.hpp:30 */

/* ------------------------------------------------------------------------ */
/* small linear algebra                                                      */
/* ------------------------------------------------------------------------ */

/* Mat::inv() DECOMP_LU for 4x4 CV_32F (util.cpp:174): OpenCV hal::LU32f. */
int orc_lu_inv4(const float *m, float *out)
{
    float A[16], b[16];
    int i, j, k;
    memcpy(A, m, sizeof(A));
    for (i = 0; i < 16; i++) b[i] = 0.f;
    for (i = 0; i < 4; i++) b[i * 4 + i] = 1.f;
    for (i = 0; i < 4; i++) {
        k = i;
        for (j = i + 1; j < 4; j++)
            if (fabsf(A[j * 4 + i]) > fabsf(A[k * 4 + i])) k = j;
        if (fabsf(A[k * 4 + i]) < 1.1920929e-07f) { /* FLT_EPSILON */
            memset(out, 0, 16 * sizeof(float));
            return 0;
        }
        if (k != i) {
            for (j = i; j < 4; j++) { float t = A[i * 4 + j]; A[i * 4 + j] = A[k * 4 + j]; A[k * 4 + j] = t; }
            for (j = 0; j < 4; j++) { float t = b[i * 4 + j]; b[i * 4 + j] = b[k * 4 + j]; b[k * 4 + j] = t; }
        }
        float d = -1.f / A[i * 4 + i];
        for (j = i + 1; j < 4; j++) {
            float alpha = A[j * 4 + i] * d;
            for (k = i + 1; k < 4; k++) A[j * 4 + k] += alpha * A[i * 4 + k];
            for (k = 0; k < 4; k++) b[j * 4 + k] += alpha * b[i * 4 + k];
        }
    }
    for (i = 3; i >= 0; i--)
        for (j = 0; j < 4; j++) {
            float s = b[i * 4 + j];
            for (k = i + 1; k < 4; k++) s -= A[i * 4 + k] * b[k * 4 + j];
            b[i * 4 + j] = s / A[i * 4 + i];
        }
    memcpy(out, b, sizeof(b));
    return 1;
}

/* 4x4 * 4x4, cv::gemm small-matrix path: sequential float accumulation. */
static void mul44(const float *a, const float *b, float *c)
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            c[i * 4 + j] = ((a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j]) + a[i * 4 + 2] * b[2 * 4 + j]) + a[i * 4 + 3] * b[3 * 4 + j];
}
/* 4x4 * 4x1, sequential float accumulation. */
static void mul41(const float *a, const float *v, float *o)
{
    for (int i = 0; i < 4; i++)
        o[i] = ((a[i * 4 + 0] * v[0] + a[i * 4 + 1] * v[1]) + a[i * 4 + 2] * v[2]) + a[i * 4 + 3] * v[3];
}
/* x * (float)(1.0/(double)s): what `Mat /= s` and `Mat / s` evaluate to. */
static inline float rcpf_d(float s) { return (float)(1.0 / (double)s); }

/* ------------------------------------------------------------------------ */
/* Render::loadMesh  render_glx.cpp:230-258                                  */
/* ------------------------------------------------------------------------ */
void orc_load_mesh(const float *vtx_xyzw, int V, const int32_t *faces, int F, float *soup /* F*9 */)
{
    (void)V;
    for (int i = 0; i < F; i++)
        for (int j = 0; j < 3; j++) {
            const float *p = vtx_xyzw + 4 * faces[3 * i + j];
            soup[9 * i + 3 * j + 0] = p[0] / p[3];
            soup[9 * i + 3 * j + 1] = p[1] / p[3];
            soup[9 * i + 3 * j + 2] = p[2] / p[3];
        }
}

/* ------------------------------------------------------------------------ */
/* Rasteriser: DEFINED semantics for the GL pipeline the reference drives     */
/* (render_glx.cpp:276-280, 350, 369-397; shader.vert:9-13).                  */
/*   - clip = P * (v,1), sequential float accumulation per row;               */
/*   - homogeneous (clip-less) edge functions, pixel centres, exact-negation  */
/*     symmetric edges + top-left style tie rule -> watertight;               */
/*   - depth = NDC z = affine plane in screen space, float32; accepted iff    */
/*     -1 <= z < 1 (near/far clip + GL_LESS against the cleared 1.0);         */
/*   - nearest fragment wins, ties -> lowest triangle index (draw order).     */
/* Output: depth (H*W, NDC z, background exactly 1.0, rows top-down) and the  */
/* winning triangle index (-1 = background).                                  */
/* ------------------------------------------------------------------------ */
typedef struct {
    float e[3][3]; /* edge functions, already multiplied by sign(det) */
    float zA, zB, zC;
    int valid, x0, x1, y0, y1;
} TriSetup;

static void clip_vertex(const float *P, const float *v, float *c)
{
    for (int r = 0; r < 4; r++)
        c[r] = ((P[r * 4 + 0] * v[0] + P[r * 4 + 1] * v[1]) + P[r * 4 + 2] * v[2]) + P[r * 4 + 3];
}

static void tri_setup(const float *P, const float *tri9, int W, int H, TriSetup *t)
{
    float c[3][4];
    for (int i = 0; i < 3; i++) clip_vertex(P, tri9 + 3 * i, c[i]);
    float e[3][3];
    for (int i = 0; i < 3; i++) {
        const float *a = c[(i + 1) % 3], *b = c[(i + 2) % 3];
        e[i][0] = a[1] * b[3] - b[1] * a[3];
        e[i][1] = b[0] * a[3] - a[0] * b[3];
        e[i][2] = a[0] * b[1] - b[0] * a[1];
    }
    float det = (c[0][0] * e[0][0] + c[0][1] * e[0][1]) + c[0][3] * e[0][2];
    t->valid = 0;
    if (!(det != 0.f) || !isfinite(det)) return;
    float s = det > 0.f ? 1.f : -1.f;
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) t->e[i][k] = s * e[i][k];
    float adet = s * det;
    /* z_ndc(X,Y) = sum_i lambda_i z_i / det  ->  plane coefficients */
    t->zA = ((t->e[0][0] * c[0][2] + t->e[1][0] * c[1][2]) + t->e[2][0] * c[2][2]) / adet;
    t->zB = ((t->e[0][1] * c[0][2] + t->e[1][1] * c[1][2]) + t->e[2][1] * c[2][2]) / adet;
    t->zC = ((t->e[0][2] * c[0][2] + t->e[1][2] * c[1][2]) + t->e[2][2] * c[2][2]) / adet;
    t->valid = 1;
    /* conservative pixel bounding box (only a search bound; never changes results) */
    t->x0 = 0; t->x1 = W - 1; t->y0 = 0; t->y1 = H - 1;
    if (c[0][3] > 0.f && c[1][3] > 0.f && c[2][3] > 0.f) {
        float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
        for (int i = 0; i < 3; i++) {
            float x = c[i][0] / c[i][3], y = c[i][1] / c[i][3];
            if (!(x == x) || !(y == y)) { xmin = ymin = -1e30f; xmax = ymax = 1e30f; break; }
            xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
        }
        float fx0 = (xmin + 1.f) * 0.5f * (float)W - 2.f, fx1 = (xmax + 1.f) * 0.5f * (float)W + 2.f;
        float fy0 = (1.f - ymax) * 0.5f * (float)H - 2.f, fy1 = (1.f - ymin) * 0.5f * (float)H + 2.f;
        if (fx0 > 0.f) t->x0 = fx0 < (float)W ? (int)fx0 : W;
        if (fx1 < (float)(W - 1)) t->x1 = fx1 > -1.f ? (int)fx1 : -1;
        if (fy0 > 0.f) t->y0 = fy0 < (float)H ? (int)fy0 : H;
        if (fy1 < (float)(H - 1)) t->y1 = fy1 > -1.f ? (int)fy1 : -1;
    }
}

static inline int edge_inside(const float *e, float X, float Y)
{
    float l = (e[0] * X + e[1] * Y) + e[2];
    if (l > 0.f) return 1;
    if (l < 0.f) return 0;
    if (!(l == 0.f)) return 0; /* NaN */
    return e[0] > 0.f || (e[0] == 0.f && e[1] > 0.f);
}

void orc_raster(const float *soup, int F, const float *P, int W, int H, float *depth, int32_t *tri_idx)
{
    for (int i = 0; i < W * H; i++) { depth[i] = BACKGROUND_DEPTH; tri_idx[i] = -1; }
    float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    for (int f = 0; f < F; f++) {
        TriSetup t;
        tri_setup(P, soup + 9 * f, W, H, &t);
        if (!t.valid) continue;
        for (int row = t.y0; row <= t.y1; row++) {
            float Y = 1.0f - ((float)row + 0.5f) * sy;
            for (int col = t.x0; col <= t.x1; col++) {
                float X = ((float)col + 0.5f) * sx - 1.0f;
                if (!edge_inside(t.e[0], X, Y) || !edge_inside(t.e[1], X, Y) || !edge_inside(t.e[2], X, Y)) continue;
                float z = ((t.zA * X + t.zB * Y) + t.zC) + 0.0f; /* -0 -> +0 */
                if (!(z >= -1.0f && z < 1.0f)) continue;
                int idx = row * W + col;
                if (z < depth[idx]) { depth[idx] = z; tri_idx[idx] = f; } /* GL_LESS, draw order */
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Shadow-map dilation, render_glx.cpp:287-314, verbatim semantics.           */
/* The reference runs it on the GL (bottom-up) image; `shadow_gl` must be in   */
/* that orientation. In place.                                                */
/* ------------------------------------------------------------------------ */
void orc_dilate_shadow_gl(float *shadow, int W, int H)
{
    float *prevRowHF = (float *)calloc((size_t)W, sizeof(float));
    float *curRow = shadow, *prevRow;
    for (int j = 1; j < W - 1; j++) {
        prevRowHF[j] = curRow[j];
        if (curRow[j - 1] < prevRowHF[j]) prevRowHF[j] = curRow[j - 1];
        if (curRow[j + 1] < prevRowHF[j]) prevRowHF[j] = curRow[j + 1];
        curRow[j] = prevRowHF[j];
    }
    for (int i = 1; i < H; i++) {
        prevRow = curRow;
        curRow = shadow + (size_t)i * W;
        float prevVal = curRow[0];
        for (int j = 1; j < W - 1; j++) {
            float val = curRow[j];
            if (prevVal > curRow[j]) curRow[j] = prevVal;
            if (curRow[j + 1] > curRow[j]) curRow[j] = curRow[j + 1];
            float valHF = curRow[j];
            if (prevRowHF[j] > curRow[j]) curRow[j] = prevRowHF[j];
            if (curRow[j] > prevRow[j]) prevRow[j] = curRow[j];
            prevRowHF[j] = valHF;
            prevVal = val;
        }
    }
    free(prevRowHF);
}

/* ------------------------------------------------------------------------ */
/* Render::projected  render_glx.cpp:261-367 + shader.frag:11-25              */
/* frame: H*W 8UC1 top-down (side camera's frame).  out: H*W*3 8UC3 top-down.  */
/* shadow_td: dilated side-camera depth (NDC z), rows TOP-DOWN.                */
/* ------------------------------------------------------------------------ */
static inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }

void orc_shade_pixel(const float *soup, const float *Pmain, const float *Pside, const uint8_t *frame,
                     const float *shadow_td, int W, int H, int row, int col, int tri, uint8_t *rgb)
{
    rgb[0] = rgb[1] = rgb[2] = 0;
    if (tri < 0) return;
    float c[3][4];
    const float *tv = soup + 9 * tri;
    for (int i = 0; i < 3; i++) clip_vertex(Pmain, tv + 3 * i, c[i]);
    float sx = 2.0f / (float)W, sy = 2.0f / (float)H;
    float X = ((float)col + 0.5f) * sx - 1.0f, Y = 1.0f - ((float)row + 0.5f) * sy;
    float l[3];
    for (int i = 0; i < 3; i++) {
        const float *a = c[(i + 1) % 3], *b = c[(i + 2) % 3];
        float e0 = a[1] * b[3] - b[1] * a[3];
        float e1 = b[0] * a[3] - a[0] * b[3];
        float e2 = a[0] * b[1] - b[0] * a[1];
        l[i] = (e0 * X + e1 * Y) + e2;
    }
    float L = (l[0] + l[1]) + l[2];
    float pos[3];
    for (int k = 0; k < 3; k++) /* perspective-correct interpolation of the world position (shader.vert:12) */
        pos[k] = ((l[0] * tv[0 + k] + l[1] * tv[3 + k]) + l[2] * tv[6 + k]) / L;
    float sc[4];
    clip_vertex(Pside, pos, sc); /* shader.frag:15 */
    float nx = sc[0] / sc[3], ny = sc[1] / sc[3], nz = sc[2] / sc[3];
    int inframe = nx > -1.0f && nx < 1.0f && ny > -1.0f && ny < 1.0f; /* shader.frag:19 */
    if (!inframe) return;
    float u = nx * 0.5f + 0.5f, v = ny * 0.5f + 0.5f; /* shader.frag:17,22 under GL_REPEAT */
    /* shadow map: GL_NEAREST */
    int si = (int)floorf(u * (float)W), sj = (int)floorf(v * (float)H);
    si = si < 0 ? 0 : (si > W - 1 ? W - 1 : si);
    sj = sj < 0 ? 0 : (sj > H - 1 ? H - 1 : sj);
    float shadowDepth = shadow_td[(size_t)(H - 1 - sj) * W + si];
    int visible = shadowDepth + 0.01f > nz; /* shader.frag:18 */
    if (!visible) return;
    /* side frame: GL_LINEAR on level 0, GL_REPEAT (quirk C17) */
    float tx = u * (float)W - 0.5f, ty = v * (float)H - 0.5f;
    float fx0 = floorf(tx), fy0 = floorf(ty);
    float ax = tx - fx0, ay = ty - fy0;
    int i0 = wrapi((int)fx0, W), i1 = wrapi((int)fx0 + 1, W);
    int j0 = wrapi((int)fy0, H), j1 = wrapi((int)fy0 + 1, H);
    float t00 = frame[(size_t)(H - 1 - j0) * W + i0], t10 = frame[(size_t)(H - 1 - j0) * W + i1];
    float t01 = frame[(size_t)(H - 1 - j1) * W + i0], t11 = frame[(size_t)(H - 1 - j1) * W + i1];
    float bx = 1.0f - ax, by = 1.0f - ay;
    float val = by * (bx * t00 + ax * t10) + ay * (bx * t01 + ax * t11);
    float r = rintf(val);
    rgb[0] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
    rgb[1] = rgb[2] = 255;
}

void orc_projected(const float *soup, int F, const float *Pmain, const uint8_t *frame, const float *Pside,
                   int W, int H, uint8_t *out_rgb)
{
    size_t N = (size_t)W * H;
    float *shadow = (float *)malloc(N * sizeof(float));
    float *tmp = (float *)malloc(N * sizeof(float));
    int32_t *tri = (int32_t *)malloc(N * sizeof(int32_t));
    /* shadow pass from the side camera (render_glx.cpp:265-280) */
    orc_raster(soup, F, Pside, W, H, tmp, tri);
    for (int r = 0; r < H; r++) memcpy(shadow + (size_t)r * W, tmp + (size_t)(H - 1 - r) * W, W * sizeof(float)); /* GL orientation */
    orc_dilate_shadow_gl(shadow, W, H);
    for (int r = 0; r < H; r++) memcpy(tmp + (size_t)r * W, shadow + (size_t)(H - 1 - r) * W, W * sizeof(float)); /* back to top-down */
    /* colour pass from the main camera (render_glx.cpp:336-350) */
    orc_raster(soup, F, Pmain, W, H, shadow, tri);
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++)
            orc_shade_pixel(soup, Pmain, Pside, frame, tmp, W, H, row, col, tri[(size_t)row * W + col], out_rgb + 3 * ((size_t)row * W + col));
    free(shadow); free(tmp); free(tri);
}

/* ------------------------------------------------------------------------ */
/* mixBackground  util.cpp:366-387                                            */
/* ------------------------------------------------------------------------ */
void orc_mix_background(const uint8_t *rgb, const uint8_t *bg, float *depth, int W, int H, uint8_t *out)
{
    for (size_t i = 0; i < (size_t)W * H; i++) {
        if (depth[i] == BACKGROUND_DEPTH || !rgb[3 * i + 1]) {
            out[i] = bg[i];
            depth[i] = BACKGROUND_DEPTH;
        } else
            out[i] = rgb[3 * i];
    }
}

/* ------------------------------------------------------------------------ */
/* goodSample util.cpp:44-53 ; sampleImage<T> util.cpp:438-461                */
/* ------------------------------------------------------------------------ */
static int good_sample(const float *img, int W, int H, float x, float y)
{
    int ix = (int)x, iy = (int)y;
    if (ix <= 0 || ix >= W - 1 || iy <= 0 || iy >= H - 1) return 0;
    return img[iy * W + ix] != BACKGROUND_DEPTH && img[iy * W + ix + 1] != BACKGROUND_DEPTH &&
           img[(iy + 1) * W + ix] != BACKGROUND_DEPTH && img[(iy + 1) * W + ix + 1] != BACKGROUND_DEPTH;
}

/* linear element index of at<T>(y, x) on a continuous Mat, clamped to the buffer
 * (quirk C5: the reference reads past the last row for border pixels; reads that
 * stay inside the buffer wrap to the next row exactly as continuous memory does,
 * reads past the end are clamped to the last element). */
static inline size_t at_index(int W, int H, float y, float x)
{
    long idx = (long)(int)y * W + (long)(int)x;
    long last = (long)W * H - 1;
    if (idx < 0) idx = 0;
    if (idx > last) idx = last;
    return (size_t)idx;
}

/* sampleImage<float>: NB the weights are swapped w.r.t. true bilinear (quirk C3). */
static float sample_float(const float *img, int W, int H, float x, float y)
{
    float lw = (float)fmod((double)x, 1.0), rw = 1 - lw, tw = (float)fmod((double)y, 1.0), bw = 1 - tw;
    float a = img[at_index(W, H, y, x)], b = img[at_index(W, H, y, x + 1)];
    float c = img[at_index(W, H, y + 1, x)], d = img[at_index(W, H, y + 1, x + 1)];
    return (a * lw + b * rw) * tw + (c * lw + d * rw) * bw;
}

/* saturate_cast<int>(float) == cvRound: round-half-even, x86 "integer indefinite" on overflow */
static inline int32_t cv_round_f(float v)
{
    if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT32_MIN;
    return (int32_t)lrintf(v);
}
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* sampleImage<cv::Point> applied to the CV_32FC2 gradient (quirk C4): the float
 * bit patterns are read as ints, scaled by float weights with saturate_cast<int>,
 * added as ints, and the resulting ints are stored back into float slots. */
static void sample_point_bits(const float *grad2, int W, int H, float x, float y, float *gx, float *gy)
{
    float lw = (float)fmod((double)x, 1.0), rw = 1 - lw, tw = (float)fmod((double)y, 1.0), bw = 1 - tw;
    const int32_t *g = (const int32_t *)grad2;
    size_t ia = at_index(W, H, y, x), ib = at_index(W, H, y, x + 1), ic = at_index(W, H, y + 1, x), id = at_index(W, H, y + 1, x + 1);
    int32_t out[2];
    for (int k = 0; k < 2; k++) {
        int32_t a = g[2 * ia + k], b = g[2 * ib + k], c = g[2 * ic + k], d = g[2 * id + k];
        int32_t top = wrap_add(cv_round_f((float)a * lw), cv_round_f((float)b * rw));
        int32_t bot = wrap_add(cv_round_f((float)c * lw), cv_round_f((float)d * rw));
        out[k] = wrap_add(cv_round_f((float)top * tw), cv_round_f((float)bot * bw));
    }
    memcpy(gx, &out[0], 4);
    memcpy(gy, &out[1], 4);
}

/* ------------------------------------------------------------------------ */
/* triangulatePixels pass 1 (util.cpp:167-248) + triangulatePixel (62-164)    */
/* flows: S pointers to H*W*4 float (u, v, variance, 0). cams: S*16.          */
/* gradient: H*W*2 float = imageGradient(depth) (util.cpp:176, cv::Sobel).     */
/* dense5: H*W*5 out (X,Y,Z,W,pdf) ; valid: H*W out (0/1).                     */
/* ------------------------------------------------------------------------ */
#define MAX_SIDE 16

typedef struct {
    float Pinv[16];
    float M[MAX_SIDE][16];   /* camera * mainCameraInv                       util.cpp:99,209 */
    float B[MAX_SIDE][6];    /* camera[0:2,0:3] * mainCameraInv[0:3,0:3]      util.cpp:219    */
    float pd[MAX_SIDE][2];   /* projectionDerivatives column                  util.cpp:86     */
    float pw[MAX_SIDE][4];   /* projectionW row                               util.cpp:87,89  */
    int S;
} TriCtx;

static void tri_ctx_init(TriCtx *c, const float *Pmain, const float *cams, int S)
{
    c->S = S;
    orc_lu_inv4(Pmain, c->Pinv);
    for (int i = 0; i < S; i++) {
        const float *P = cams + 16 * i;
        mul44(P, c->Pinv, c->M[i]);
        for (int r = 0; r < 2; r++)
            for (int k = 0; k < 3; k++) /* K=3 == result width -> float sequential */
                c->B[i][r * 3 + k] = (P[r * 4 + 0] * c->Pinv[0 * 4 + k] + P[r * 4 + 1] * c->Pinv[1 * 4 + k]) + P[r * 4 + 2] * c->Pinv[2 * 4 + k];
        for (int r = 0; r < 2; r++) { /* 2x4 * 4x1 -> generic gemm, double accumulation */
            double s = 0;
            for (int k = 0; k < 4; k++) s += (double)P[r * 4 + k] * (double)c->Pinv[k * 4 + 2];
            c->pd[i][r] = (float)s;
        }
        for (int k = 0; k < 4; k++) /* Sx4 * 4x4: K=4 == result width -> float sequential */
            c->pw[i][k] = ((P[12] * c->Pinv[0 * 4 + k] + P[13] * c->Pinv[1 * 4 + k]) + P[14] * c->Pinv[2 * 4 + k]) + P[15] * c->Pinv[3 * 4 + k];
    }
}

/* diagnostics: if non-NULL, receives (first iteration index at which z repeats an earlier value, period) */
static __thread int *g_cycle_out = 0;
/* returns number of Newton iterations performed */
static int triangulate_pixel(const TriCtx *c, float x, float y, const float *meas /*S*2*/, const float *icov /*S*4*/,
                             float depth, float *out4, float *pdf_out)
{
    int S = c->S;
    float k[4] = {x, y, depth, 1.f};
    float p[MAX_SIDE][2], dp[MAX_SIDE][2], diff[MAX_SIDE][2];
    int iter;
    float zhist[64];
    int rep_at = -1, rep_period = 0;
    for (iter = 0;; iter++) {
        if (g_cycle_out && rep_at < 0) {
            for (int j = iter - 1; j >= 0; j--)
                if (zhist[j] == k[2]) { rep_at = iter; rep_period = iter - j; break; }
        }
        zhist[iter] = k[2];
        for (int i = 0; i < S; i++) {
            float est[4];
            mul41(c->M[i], k, est);
            float sc = rcpf_d(est[3]);
            p[i][0] = est[0] * sc;
            p[i][1] = est[1] * sc;
        }
        for (int i = 0; i < S; i++) {
            float w;
            if (S == 4) /* Sx4 * 4x1 with K == result height -> float sequential */
                w = ((c->pw[i][0] * k[0] + c->pw[i][1] * k[1]) + c->pw[i][2] * k[2]) + c->pw[i][3] * k[3];
            else {
                double s = 0;
                for (int q = 0; q < 4; q++) s += (double)c->pw[i][q] * (double)k[q];
                w = (float)s;
            }
            dp[i][0] = c->pd[i][0] / w;
            dp[i][1] = c->pd[i][1] / w;
        }
        double firstDz = 0, secondDz = 0;
        for (int i = 0; i < S; i++) {
            diff[i][0] = p[i][0] - meas[2 * i];
            diff[i][1] = p[i][1] - meas[2 * i + 1];
            const float *ic = icov + 4 * i;
            float t0 = ic[0] * dp[i][0] + ic[1] * dp[i][1];
            float t1 = ic[2] * dp[i][0] + ic[3] * dp[i][1];
            firstDz += (double)diff[i][0] * (double)t0 + (double)diff[i][1] * (double)t1;
            secondDz += (double)dp[i][0] * (double)t0 + (double)dp[i][1] * (double)t1;
        }
        double delta_z = -firstDz / secondDz, eps = 1e-7;
        if (iter >= 50 || (delta_z < eps && delta_z > -eps)) {
            double exponent = 0, product_ivar = 1;
            for (int i = 0; i < S; i++) {
                const float *ic = icov + 4 * i;
                float t0 = ic[0] * diff[i][0] + ic[1] * diff[i][1];
                float t1 = ic[2] * diff[i][0] + ic[3] * diff[i][1];
                exponent -= (double)diff[i][0] * (double)t0 + (double)diff[i][1] * (double)t1;
                product_ivar *= (double)ic[0] * (double)ic[3] - (double)ic[1] * (double)ic[2];
            }
            *pdf_out = (float)(0.159 * product_ivar * exp(0.5 * exponent));
            break;
        }
        k[2] = (float)((double)k[2] + delta_z);
    }
    mul41(c->Pinv, k, out4);
    if (g_cycle_out) { g_cycle_out[0] = rep_at; g_cycle_out[1] = rep_period; }
    return iter;
}

/* one pixel of pass 1; returns 1 if a point was produced */
static int triangulate_at(const TriCtx *c, const float *const *flows, const float *depth, const float *grad2,
                          int W, int H, int row, int col, float *out5, int *iters)
{
    float d0 = depth[(size_t)row * W + col];
    if (d0 == BACKGROUND_DEPTH) return 0;
    int S = c->S;
    float centerX = (float)(W / 2.0), centerY = (float)(H / 2.0);
    float scaleX = (float)(2.0 / W), scaleY = (float)(2.0 / H);
    float x = ((float)col - centerX) * scaleX, y = (centerY - (float)row) * scaleY;
    float meas[MAX_SIDE * 2], icov[MAX_SIDE * 4];
    for (int i = 0; i < S; i++) {
        const float *fl = flows[i] + 4 * ((size_t)row * W + col);
        float flx = fl[0], fly = fl[1], variance = fl[2];
        float sxp = (float)col + flx, syp = (float)row + fly;
        int good = good_sample(depth, W, H, sxp, syp);
        float z = good ? sample_float(depth, W, H, sxp, syp) : d0;
        float vec[4] = {x + flx * scaleX, y + fly * scaleY, z, 1.f}; /* quirk C6: +fly */
        float m[4];
        mul41(c->M[i], vec, m);
        float gx, gy;
        if (good) sample_point_bits(grad2, W, H, sxp, syp, &gx, &gy);
        else sample_point_bits(grad2, W, H, (float)col, (float)row, &gx, &gy);
        /* A = B * D, D = [[1,0],[0,1],[gx,gy]] : 2x3 * 3x2 -> generic gemm (double accumulation) */
        const float *B = c->B[i];
        float A[4];
        for (int r = 0; r < 2; r++) {
            A[r * 2 + 0] = (float)(((double)B[r * 3 + 0] * 1.0 + (double)B[r * 3 + 1] * 0.0) + (double)B[r * 3 + 2] * (double)gx);
            A[r * 2 + 1] = (float)(((double)B[r * 3 + 0] * 0.0 + (double)B[r * 3 + 1] * 1.0) + (double)B[r * 3 + 2] * (double)gy);
        }
        float sw = rcpf_d(m[3]);
        for (int q = 0; q < 4; q++) A[q] = A[q] * sw;
        /* A * A^T (GEMM_2_T -> double accumulation) */
        float C[4];
        for (int r = 0; r < 2; r++)
            for (int q = 0; q < 2; q++)
                C[r * 2 + q] = (float)((double)A[r * 2] * (double)A[q * 2] + (double)A[r * 2 + 1] * (double)A[q * 2 + 1]);
        /* 2x2 inverse */
        double det = (double)C[0] * (double)C[3] - (double)C[1] * (double)C[2];
        float inv[4] = {0, 0, 0, 0};
        if (det != 0.0) {
            float d = (float)(1.0 / det);
            inv[0] = C[3] * d; inv[1] = C[1] * (-d); inv[2] = C[2] * (-d); inv[3] = C[0] * d;
        }
        float sv = rcpf_d(variance);
        for (int q = 0; q < 4; q++) icov[4 * i + q] = inv[q] * sv;
        for (int q = 0; q < 4; q++) m[q] = m[q] * sw;
        if (m[2] < -1.f) return 0;
        meas[2 * i] = m[0];
        meas[2 * i + 1] = m[1];
    }
    int it = triangulate_pixel(c, x, y, meas, icov, d0, out5, out5 + 4);
    if (iters) *iters = it;
    return 1;
}

void orc_triangulate_dense_ex(const float *const *flows, int S, const float *Pmain, const float *cams,
                              const float *depth, const float *grad2, int W, int H,
                              float *dense5, uint8_t *valid, int32_t *iters_out, int32_t *cycle_out /* H*W*2 or NULL */);
void orc_triangulate_dense(const float *const *flows, int S, const float *Pmain, const float *cams,
                           const float *depth, const float *grad2, int W, int H,
                           float *dense5, uint8_t *valid, int32_t *iters_out)
{
    orc_triangulate_dense_ex(flows, S, Pmain, cams, depth, grad2, W, H, dense5, valid, iters_out, 0);
}
void orc_triangulate_dense_ex(const float *const *flows, int S, const float *Pmain, const float *cams,
                              const float *depth, const float *grad2, int W, int H,
                              float *dense5, uint8_t *valid, int32_t *iters_out, int32_t *cycle_out)
{
    TriCtx c;
    tri_ctx_init(&c, Pmain, cams, S);
#pragma omp parallel for schedule(dynamic, 4)
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            size_t i = (size_t)row * W + col;
            int it = 0;
            int cyc[2] = {-1, 0};
            g_cycle_out = cycle_out ? cyc : 0;
            valid[i] = (uint8_t)triangulate_at(&c, flows, depth, grad2, W, H, row, col, dense5 + 5 * i, &it);
            if (cycle_out) { cycle_out[2 * i] = cyc[0]; cycle_out[2 * i + 1] = cyc[1]; }
            if (iters_out) iters_out[i] = valid[i] ? it : -1;
            if (!valid[i]) for (int q = 0; q < 5; q++) dense5[5 * i + q] = 0.f;
        }
}

/* ------------------------------------------------------------------------ */
/* extractCameraCenter util.cpp:33-41 : homogeneous null vector of rows 0,1,3 */
/* (cv::decomposeProjectionMatrix works in double and returns it as float).    */
/* ------------------------------------------------------------------------ */
static double det3(const double *a, const double *b, const double *c)
{
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
void orc_camera_center(const float *P, float *c3)
{
    double r[3][4];
    const int rows[3] = {0, 1, 3};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) r[i][j] = P[rows[i] * 4 + j];
    double col[4][3];
    for (int j = 0; j < 4; j++) for (int i = 0; i < 3; i++) col[j][i] = r[i][j];
    double X = det3(col[1], col[2], col[3]), Y = -det3(col[0], col[2], col[3]);
    double Z = det3(col[0], col[1], col[3]), T = -det3(col[0], col[1], col[2]);
    /* normalise like an SVD null vector (unit length) before the float cast */
    double n = sqrt(X * X + Y * Y + Z * Z + T * T);
    float fx = (float)(X / n), fy = (float)(Y / n), fz = (float)(Z / n), ft = (float)(T / n);
    float s = rcpf_d(ft);
    c3[0] = fx * s; c3[1] = fy * s; c3[2] = fz * s;
}

/* ------------------------------------------------------------------------ */
/* cv::eigen on a symmetric 3x3 CV_32F matrix (Jacobi, OpenCV JacobiImpl_).   */
/* W: eigenvalues descending, V: eigenvectors in rows.                        */
/* ------------------------------------------------------------------------ */
/* OpenCV's own hypot of lapack.cpp (JacobiImpl_ does not call libm's): float arithmetic, scaled by the larger operand */
static float cv_hypotf(float a, float b)
{
    a = fabsf(a); b = fabsf(b);
    if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
    return 0;
}

void orc_jacobi3(float *A, float *W, float *V)
{
    const int n = 3;
    const float eps = 1.1920929e-07f;
    int i, j, k, m, indR[3], indC[3];
    float mv = 0;
    for (i = 0; i < n; i++) { for (j = 0; j < n; j++) V[i * n + j] = 0; V[i * n + i] = 1; }
    for (k = 0; k < n; k++) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabsf(A[n * k + m]), i = k + 2; i < n; i++) {
                float val = fabsf(A[n * k + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabsf(A[k]), i = 1; i < k; i++) {
                float val = fabsf(A[n * i + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }
    for (int iters = 0; iters < n * n * 30; iters++) {
        for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < n - 1; i++) {
            float val = fabsf(A[n * i + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        int l = indR[k];
        for (i = 1; i < n; i++) {
            float val = fabsf(A[n * indC[i] + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }
        float p = A[n * k + l];
        if (fabsf(p) <= eps) break;
        float y = (float)((W[l] - W[k]) * 0.5);
        float t = fabsf(y) + cv_hypotf(p, y);
        float s = cv_hypotf(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[n * k + l] = 0;
        W[k] -= t;
        W[l] += t;
        float a0, b0;
#define ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
        for (i = 0; i < k; i++) ROT(A[n * i + k], A[n * i + l]);
        for (i = k + 1; i < l; i++) ROT(A[n * k + i], A[n * i + l]);
        for (i = l + 1; i < n; i++) ROT(A[n * k + i], A[n * l + i]);
        for (i = 0; i < n; i++) ROT(V[n * k + i], V[n * l + i]);
#undef ROT
        for (j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = fabsf(A[n * idx + m]), i = idx + 2; i < n; i++) {
                    float val = fabsf(A[n * idx + i]);
                    if (mv < val) mv = val, m = i;
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; i++) {
                    float val = fabsf(A[n * i + idx]);
                    if (mv < val) mv = val, m = i;
                }
                indC[idx] = m;
            }
        }
    }
    for (k = 0; k < n - 1; k++) {
        m = k;
        for (i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
        if (k != m) {
            float t = W[m]; W[m] = W[k]; W[k] = t;
            for (i = 0; i < n; i++) { t = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = t; }
        }
    }
}

/* cv::PCA(data Kx3 CV_32F, noArray, DATA_AS_ROW) -> smallest eigenvector.
 * mean: float column sums (sequential) * (float)(1/K); covariance: double
 * accumulation of float-centred products * (1/K) -> float; eigen: Jacobi. */
void orc_pca_normal(const float *pts, int K, float *normal3, float *evals3)
{
    float mean[3] = {pts[0], pts[1], pts[2]};
    for (int i = 1; i < K; i++) for (int q = 0; q < 3; q++) mean[q] = mean[q] + pts[3 * i + q];
    float sK = (float)(1.0 / (double)K);
    for (int q = 0; q < 3; q++) mean[q] = mean[q] * sK;
    float cov[9];
    double scale = 1.0 / (double)K;
    for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) {
            double s = 0;
            for (int i = 0; i < K; i++) s += (double)(float)(pts[3 * i + a] - mean[a]) * (double)(float)(pts[3 * i + b] - mean[b]);
            cov[a * 3 + b] = cov[b * 3 + a] = (float)(s * scale);
        }
    float Wv[3], V[9];
    orc_jacobi3(cov, Wv, V);
    normal3[0] = V[6]; normal3[1] = V[7]; normal3[2] = V[8];
    if (evals3) { evals3[0] = Wv[0]; evals3[1] = Wv[1]; evals3[2] = Wv[2]; }
}

/* ------------------------------------------------------------------------ */
/* triangulatePixels pass 2: normals  util.cpp:250-326                        */
/* dense5/valid from pass 1. Writes compacted rows (x,y,z,w,nx,ny,nz) in      */
/* row-major pixel order; returns the number of rows M.                       */
/* Quirk decisions: `dot` starts at 0 (C8); neighbourhood cleared for every    */
/* pixel (C9).                                                                */
/* ------------------------------------------------------------------------ */
int orc_normals_compact(const float *dense5, const uint8_t *valid, int W, int H, const float *Pmain,
                        const float *cams, int S, float *out7, float *evals_out /* M*4 (l0,l1,l2,K) or NULL: test diagnostics */)
{
    const int radius = 10;
    int nc = S + 1;
    float centers[(MAX_SIDE + 1) * 3];
    orc_camera_center(Pmain, centers);
    for (int i = 0; i < S; i++) orc_camera_center(cams + 16 * i, centers + 3 * (i + 1));
    /* row-major exclusive scan = pixelIndices (util.cpp:241) */
    int32_t *pid = (int32_t *)malloc((size_t)W * H * sizeof(int32_t));
    int M = 0;
    for (size_t i = 0; i < (size_t)W * H; i++) pid[i] = valid[i] ? M++ : -1;
    /* dehomogenised points: row[0:3] * (float)(1/w)  (util.cpp:290) */
    float *deh = (float *)malloc((size_t)W * H * 3 * sizeof(float));
    for (size_t i = 0; i < (size_t)W * H; i++) {
        if (!valid[i]) continue;
        float s = rcpf_d(dense5[5 * i + 3]);
        for (int q = 0; q < 3; q++) deh[3 * i + q] = dense5[5 * i + q] * s;
    }
#pragma omp parallel
    {
        float *nb = (float *)malloc(441 * 3 * sizeof(float));
#pragma omp for schedule(dynamic, 2)
        for (int row = 0; row < H; row++)
            for (int col = 0; col < W; col++) {
                size_t i = (size_t)row * W + col;
                if (pid[i] < 0) continue;
                float pdf = dense5[5 * i + 4];
                if (S > 1) pdf = (float)pow((double)pdf, 1.0 / S);
                int K = 0;
                for (int ny = row - radius; ny <= row + radius; ny++) {
                    if (ny < 0 || ny >= H) continue;
                    for (int nx = col - radius; nx <= col + radius; nx++) {
                        if (nx < 0 || nx >= W) continue;
                        size_t j = (size_t)ny * W + nx;
                        if (!valid[j]) continue;
                        nb[3 * K] = deh[3 * j]; nb[3 * K + 1] = deh[3 * j + 1]; nb[3 * K + 2] = deh[3 * j + 2];
                        K++;
                    }
                }
                float n[3];
                float ev[3] = {0.f, 0.f, 0.f};
                if (K >= 3) {
                    orc_pca_normal(nb, K, n, ev);
                    float dot = 0.f;
                    for (int c = 0; c < nc; c++) {
                        double d = 0;
                        for (int q = 0; q < 3; q++) d += (double)n[q] * (double)(float)(centers[3 * c + q] - deh[3 * i + q]);
                        dot = (float)((double)dot + 1.0 / d);
                    }
                    if (dot < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
                } else {
                    n[0] = n[1] = n[2] = 0.f;
                    for (int c = 0; c < nc; c++) {
                        float v[3];
                        double vv = 0;
                        for (int q = 0; q < 3; q++) { v[q] = centers[3 * c + q] - dense5[5 * i + q]; vv += (double)v[q] * (double)v[q]; }
                        float s = (float)(1.0 / vv);
                        for (int q = 0; q < 3; q++) n[q] = n[q] + v[q] * s;
                    }
                }
                double nn = sqrt((double)n[0] * n[0] + (double)n[1] * n[1] + (double)n[2] * n[2]);
                float sc = (float)((double)pdf * (1.0 / nn));
                float *o = out7 + 7 * (size_t)pid[i];
                if (evals_out) {
                    float *e = evals_out + 4 * (size_t)pid[i];
                    e[0] = ev[0]; e[1] = ev[1]; e[2] = ev[2]; e[3] = (float)K;
                }
                for (int q = 0; q < 4; q++) o[q] = dense5[5 * i + q];
                for (int q = 0; q < 3; q++) o[4 + q] = n[q] * sc;
            }
        free(nb);
    }
    free(pid); free(deh);
    return M;
}
