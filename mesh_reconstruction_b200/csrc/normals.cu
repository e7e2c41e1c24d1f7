// normals.cu -- window-PCA normals, stage 1 (util.cpp:282-301): the 3x3 covariance cv::PCA builds from the valid points
// of the 21x21 window of every pixel, BIT-IDENTICAL to the reference's evaluation:
//     mean_q = (float sum of the K samples, in window scan order) * (float)(1/K)                    [cv::reduce, CV_32F]
//     cov_ab = (float)( (double sum over the samples of fl32(x_a - mean_a) * fl32(x_b - mean_b)) * (1.0 / K) )   [cv::mulTransposed]
// (oracle/recon_oracle.c orc_pca_normal).  The float sum of the mean is order dependent and is evaluated as such: 441
// dependent float adds per pixel and coordinate, four adjacent pixels per thread sharing their loads.
//
// The covariance sums are taken sample by sample only where they have to be.  Per tile (32 x 24 output pixels + 10 pixel
// halo) and coordinate, let [a, b] be the range of |x| over the valid points.  A coordinate is EXACT on the tile if all
// its values have one sign and b <= 1.999 a; then
//   * every fl32(x - mean) is exact (Sterbenz: the float mean lies within [a (1 - 3e-5), b (1 + 3e-5)]),
//   * with u = 2^(floor(log2 a) - 24) every x, the mean and the anchor c (the sample of smallest magnitude) are integer
//     multiples of u, and n = (x - c) / u is an integer with |n| <= 2^25.
// For two exact coordinates  sum_i d_a d_b = u_a u_b T_ab  with  T_ab = S2_ab - m_b S1_a - m_a S1_b + K m_a m_b  a 64-bit
// INTEGER built from the window moments K, S1_a = sum n_a, S2_ab = sum n_a n_b (integer box sums are associative: they
// are separable and sliding) and m = (mean - c) / u.  Whenever T_aa < 2^53 for the exact coordinates, every partial sum
// of the reference's double accumulation is an integer below 2^53 in units of u_a u_b, i.e. that accumulation is exact
// too and equals u_a u_b T_ab: the entry is (float)(T_ab * 2^(e_a + e_b) * (1.0 / K)), bit for bit, without visiting
// the samples.  Entries that involve a coordinate that is NOT exact on the tile (it crosses zero, spans more than a
// factor of two, or holds NaN / inf) are accumulated like the reference does: float-centred samples, one double
// accumulator per entry, window order, three vertically adjacent pixels per thread sharing their loads (products of two
// floats are exact in double, so fma(d_a, d_b, s) rounds exactly like s + d_a * d_b; for an exact coordinate the centred
// sample is taken as (double)x - (double)mean, which saves the float -> double conversion on the quarter-rate pipe).
// The rare pixels with T_aa >= 2^53 run the plain six-accumulator loop.
#include "normals.cuh"

namespace {

constexpr int R = 10, TX = 32, TY = 24, VR = 3, NT = TX * TY / VR;       // 256 threads, 3 vertically adjacent pixels each
constexpr int TW = TX + 2 * R, TH = TY + 2 * R, TP = TW + 1, HP = TX + 1, RUN = 8, MRUN = 4, WIN = 2 * R + 1;
constexpr int NPL = 4;                                                   // planes of horizontal box sums
static_assert(TH * (TX / RUN) <= NT && TY * (TX / MRUN) <= NT, "one horizontal run / mean run per thread");

struct TileInfo {
    float c[3], scale[3], unit[3];   // per SLOT (after the permutation): anchor, 1/u, u
    int e[3];                        // log2 u
    int perm[3];                     // perm[k] = original coordinate held by slot k; exact coordinates first
    int ns;                          // number of slots that are not exact: slots 3-ns .. 2
    int allvalid;                    // every tile entry inside the image is a valid point
};

struct Smem {
    float4 tile[TH][TP];             // (slot0, slot1, slot2, valid); exact slots become integers n late in the kernel
    union {
        long long hs[NPL][TH][HP];   // horizontal box sums
        double dbl[2][TH][TP];       // (double) of the exact slots, for the sample loop
    };
    float4 mean[TY][HP];             // float sums of the window, then the mean
    TileInfo ti;
    float lo[NT / 32][3], hi[NT / 32][3];
    int flags[NT / 32];
};

__device__ __forceinline__ double pow2_d(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }
__device__ __forceinline__ float slot(const float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__host__ __device__ constexpr int pidx(int a, int b) { return a == 0 ? b : (a == 1 ? 2 + b : 5); }   // (a <= b) -> 0..5

// separable sliding integer box sums of NM per-entry moments over the 21x21 windows of the block's pixels
template <int NM, class F>
__device__ __forceinline__ void box_pass(Smem &s, const int tid, F f, long long (*out)[NM])
{
    if (tid < TH * (TX / RUN)) {
        const int r = tid % TH, c0 = (tid / TH) * RUN;
        long long acc[NM], m[NM], m2[NM];
#pragma unroll
        for (int q = 0; q < NM; q++) acc[q] = 0;
#pragma unroll
        for (int t = 0; t < WIN; t++) {
            f(r, c0 + t, m);
#pragma unroll
            for (int q = 0; q < NM; q++) acc[q] += m[q];
        }
#pragma unroll
        for (int q = 0; q < NM; q++) s.hs[q][r][c0] = acc[q];
#pragma unroll
        for (int j = 1; j < RUN; j++) {
            f(r, c0 + j + 2 * R, m);
            f(r, c0 + j - 1, m2);
#pragma unroll
            for (int q = 0; q < NM; q++) { acc[q] += m[q] - m2[q]; s.hs[q][r][c0 + j] = acc[q]; }
        }
    }
    __syncthreads();
    const int vx = tid % TX, r0 = VR * (tid / TX);
#pragma unroll
    for (int q = 0; q < NM; q++) {
        long long a = 0;
#pragma unroll
        for (int t = 0; t < WIN; t++) a += s.hs[q][r0 + t][vx];
        out[0][q] = a;
#pragma unroll
        for (int j = 1; j < VR; j++) {
            a += s.hs[q][r0 + j + 2 * R][vx] - s.hs[q][r0 + j - 1][vx];
            out[j][q] = a;
        }
    }
    __syncthreads();
}

// One window row of the sample loop for the pixels j of a thread selected by JMASK (compile time).
template <int NS, bool ALLVALID, int JMASK>
__device__ __forceinline__ void sample_row(const Smem &s, const int row, const int lx, const float (*mf)[3], const double (*M)[2], double (*acc)[6])
{
    constexpr int NF = 3 - NS;
#pragma unroll 7
    for (int dx = 0; dx < WIN; dx++) {
        const float4 v = s.tile[row][lx + dx];
        double X[2] = {0.0, 0.0};
        if (NF >= 1) X[0] = s.dbl[0][row][lx + dx];
        if (NF >= 2) X[1] = s.dbl[1][row][lx + dx];
        const bool valid = ALLVALID || v.w != 0.f;
#pragma unroll
        for (int j = 0; j < VR; j++) {
            if (!(JMASK & (1 << j))) continue;
            double D[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k < NF) D[k] = X[k] - M[j][k];              // exact coordinate: x - mean is exact in float AND double
                else {
                    float f = slot(v, k) - mf[j][k];            // the reference's float-centred sample
                    if (!ALLVALID && !valid) f = 0.f;           // invalid entries add +0 to every accumulator
                    D[k] = (double)f;
                }
            }
#pragma unroll
            for (int b = NF; b < 3; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) acc[j][pidx(a, b)] = __fma_rn(D[a], D[b], acc[j][pidx(a, b)]);
        }
    }
}

// The reference's loop for the entries that involve one of the NS inexact slots (slots 3-NS .. 2), for the three
// vertically adjacent pixels (ly0 + j, lx) of a thread: window row sy of the union serves pixel j as its row sy - j.
// acc[j][pidx(a, b)] is only touched for b >= 3 - NS.
template <int NS, bool ALLVALID>
__device__ __forceinline__ void sample_loop(const Smem &s, const int ly0, const int lx, const float (*mf)[3], double (*acc)[6])
{
    constexpr int NF = 3 - NS;
    static_assert(VR == 3, "row schedule below is written for three pixels per thread");
    double M[VR][2];
#pragma unroll
    for (int j = 0; j < VR; j++)
#pragma unroll
        for (int k = 0; k < 2; k++) M[j][k] = k < NF ? (double)mf[j][k] : 0.0;
    sample_row<NS, ALLVALID, 1>(s, ly0 + 0, lx, mf, M, acc);
    sample_row<NS, ALLVALID, 3>(s, ly0 + 1, lx, mf, M, acc);
#pragma unroll 1
    for (int sy = 2; sy < WIN; sy++) sample_row<NS, ALLVALID, 7>(s, ly0 + sy, lx, mf, M, acc);
    sample_row<NS, ALLVALID, 6>(s, ly0 + WIN, lx, mf, M, acc);
    sample_row<NS, ALLVALID, 4>(s, ly0 + WIN + 1, lx, mf, M, acc);
}

// plain six-accumulator loop for ONE pixel (window origin ly, lx), the mean being known; slots < n_int hold integers
__device__ __forceinline__ void residual_cov(const Smem &s, const TileInfo &ti, const int n_int, const int ly, const int lx,
                                             const float *mf, double *acc6)
{
#pragma unroll
    for (int q = 0; q < 6; q++) acc6[q] = 0.0;
#pragma unroll 1
    for (int dy = 0; dy < WIN; dy++)
#pragma unroll 1
        for (int dx = 0; dx < WIN; dx++) {
            const float4 v = s.tile[ly + dy][lx + dx];
            const bool valid = v.w != 0.f;
            double D[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float x = slot(v, k);
                if (k < n_int) x = ti.c[k] + (float)__float_as_int(x) * ti.unit[k];     // exact: c + (x - c)
                float f = x - mf[k];
                if (!valid) f = 0.f;
                D[k] = (double)f;
            }
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) acc6[pidx(a, b)] = __fma_rn(D[a], D[b], acc6[pidx(a, b)]);
        }
}

__device__ __forceinline__ void store_covk(CovK *out, const TileInfo &ti, const float *covp, int K)
{
    // slots -> original coordinates
    int sl[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int o = 0; o < 3; o++)
            if (ti.perm[k] == o) sl[o] = k;
    float c[6];
#pragma unroll
    for (int oa = 0; oa < 3; oa++)
#pragma unroll
        for (int ob = oa; ob < 3; ob++) {
            const int a = min(sl[oa], sl[ob]), b = max(sl[oa], sl[ob]);
            float v = covp[0];
#pragma unroll
            for (int q = 1; q < 6; q++)
                if (pidx(a, b) == q) v = covp[q];
            c[pidx(oa, ob)] = v;
        }
    float4 *p = reinterpret_cast<float4 *>(out);
    p[0] = make_float4(c[0], c[1], c[2], c[3]);
    p[1] = make_float4(c[4], c[5], __int_as_float(K), 0.f);
}

// Everything after the mean for one class of tile (NS = number of inexact slots).
template <int NS>
__device__ __forceinline__ void finish_tile(Smem &s, const TileInfo &ti, const int tid, const int bx, const int by, const int W, const int H,
                                            const long long (*A)[3], CovK *__restrict__ out, unsigned long long *__restrict__ stats)
{
    constexpr int NF = 3 - NS;
    const int vx = tid % TX, ly0 = VR * (tid / TX);
    float mf[VR][3];
    int K[VR];
    bool live[VR];
#pragma unroll
    for (int j = 0; j < VR; j++) {
        K[j] = (int)A[j][0];
        live[j] = bx + vx < W && by + ly0 + j < H && s.tile[ly0 + j + R][vx + R].w != 0.f;
        const float4 ms = s.mean[ly0 + j][vx];
        const float sK = rcpf_d((float)(K[j] > 0 ? K[j] : 1));
        mf[j][0] = ms.x * sK; mf[j][1] = ms.y * sK; mf[j][2] = ms.z * sK;
    }
    double acc[VR][6];
#pragma unroll
    for (int j = 0; j < VR; j++)
#pragma unroll
        for (int q = 0; q < 6; q++) acc[j][q] = 0.0;
    if (NS >= 1) {
        if (NF >= 1) {
            for (int i = tid; i < TW * TH; i += NT) {
                const int ty = i / TW, tx = i % TW;
                const float4 v = s.tile[ty][tx];
                s.dbl[0][ty][tx] = (double)v.x;
                if (NF >= 2) s.dbl[1][ty][tx] = (double)v.y;
            }
            __syncthreads();
        }
        if (live[0] || live[1] || live[2]) {
            if (ti.allvalid) sample_loop<NS, true>(s, ly0, vx, mf, acc);
            else sample_loop<NS, false>(s, ly0, vx, mf, acc);
        }
        if (NF >= 1) __syncthreads();           // dbl aliases the box-sum planes
    }
    long long B[VR][3], Cm[VR][4];
    if (NF >= 2) {
        // exact slots -> integers n = (x - c) / u, in place
        for (int i = tid; i < TW * TH; i += NT) {
            const int ty = i / TW, tx = i % TW;
            float4 v = s.tile[ty][tx];
            const bool valid = v.w != 0.f;
            v.x = __int_as_float(valid ? __float2int_rn((v.x - ti.c[0]) * ti.scale[0]) : 0);
            v.y = __int_as_float(valid ? __float2int_rn((v.y - ti.c[1]) * ti.scale[1]) : 0);
            if (NF >= 3) v.z = __int_as_float(valid ? __float2int_rn((v.z - ti.c[2]) * ti.scale[2]) : 0);
            s.tile[ty][tx] = v;
        }
        __syncthreads();
        box_pass<3>(s, tid, [&](int r, int c, long long *m) {
            const float4 v = s.tile[r][c];
            const long long n0 = __float_as_int(v.x), n1 = __float_as_int(v.y);
            m[0] = n1; m[1] = n0 * n1; m[2] = n1 * n1; }, B);
        if (NF >= 3)
            box_pass<4>(s, tid, [&](int r, int c, long long *m) {
                const float4 v = s.tile[r][c];
                const long long n0 = __float_as_int(v.x), n1 = __float_as_int(v.y), n2 = __float_as_int(v.z);
                m[0] = n2; m[1] = n0 * n2; m[2] = n1 * n2; m[3] = n2 * n2; }, Cm);
    }
    constexpr int n_int = NF >= 2 ? NF : 0;      // slots that hold integers by now
    unsigned residual = 0;
#pragma unroll
    for (int j = 0; j < VR; j++) {
        if (!live[j]) continue;
        float covp[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (K[j] >= 3) {
            const double scale = 1.0 / (double)K[j];
            bool ok = true;
            long long m[3] = {0, 0, 0}, S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < NF; q++) {
                const float t = (mf[j][q] - ti.c[q]) * ti.scale[q];
                const int mi = __float2int_rn(t);
                ok = ok && (float)mi == t && fabsf(t) <= 33554432.f;      // the mean is a multiple of u (checked, not assumed)
                m[q] = mi;
            }
            if (NF >= 1) { S1[0] = A[j][1]; S2[pidx(0, 0)] = A[j][2]; }
            if (NF >= 2) { S1[1] = B[j][0]; S2[pidx(0, 1)] = B[j][1]; S2[pidx(1, 1)] = B[j][2]; }
            if (NF >= 3) { S1[2] = Cm[j][0]; S2[pidx(0, 2)] = Cm[j][1]; S2[pidx(1, 2)] = Cm[j][2]; S2[pidx(2, 2)] = Cm[j][3]; }
            long long T[6] = {0, 0, 0, 0, 0, 0};
            const long long Kl = K[j], lim = 1ll << 53;
#pragma unroll
            for (int b = 0; b < NF; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) T[pidx(a, b)] = S2[pidx(a, b)] - m[b] * S1[a] - m[a] * S1[b] + Kl * m[a] * m[b];
#pragma unroll
            for (int q = 0; q < NF; q++) ok = ok && T[pidx(q, q)] >= 0 && T[pidx(q, q)] < lim;
            if (!ok) { residual |= 1u << j; continue; }
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) {
                    const int q = pidx(a, b);
                    if (b < NF) covp[q] = (float)(((double)T[q] * pow2_d(ti.e[a] + ti.e[b])) * scale);
                    else covp[q] = (float)(acc[j][q] * scale);
                }
        }
        store_covk(out + (size_t)(by + ly0 + j) * W + bx + vx, ti, covp, K[j]);
    }
    if (NF >= 1) {
        // Rare: a partial sum of the reference's double accumulation is not exactly representable -> take its own loop.
        // The block's residual pixels are gathered into one list first, so that they fill whole warps instead of
        // keeping every warp that owns one of them busy for a full window walk.
        __shared__ int s_nres;
        __shared__ unsigned s_res[TX * TY];          // ly << 8 | lx
        if (tid == 0) s_nres = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < VR; j++)
            if (residual & (1u << j)) s_res[atomicAdd(&s_nres, 1)] = (unsigned)((ly0 + j) << 8) | (unsigned)vx;
        __syncthreads();
        const int nres = s_nres;
        for (int r = tid; r < nres; r += NT) {
            const int ly = (int)(s_res[r] >> 8), lx = (int)(s_res[r] & 255u);
            // K of that pixel again (cheap next to the window walk) from the valid flags; its window sums from s.mean
            int Kj = 0;
            for (int dy = 0; dy < WIN; dy++)
                for (int dx = 0; dx < WIN; dx++) Kj += s.tile[ly + dy][lx + dx].w != 0.f;
            const float sK = rcpf_d((float)Kj);
            const float4 ms = s.mean[ly][lx];
            const float mj[3] = {ms.x * sK, ms.y * sK, ms.z * sK};
            double a6[6];
            residual_cov(s, ti, n_int, ly, lx, mj, a6);
            const double scale = 1.0 / (double)Kj;
            float covp[6];
#pragma unroll
            for (int q = 0; q < 6; q++) covp[q] = (float)(a6[q] * scale);
            store_covk(out + (size_t)(by + ly) * W + bx + lx, ti, covp, Kj);
        }
        if (tid == 0 && nres && stats) atomicAdd(stats + 4, (unsigned long long)nres);
    }
}

// stats: [0] tiles holding valid output pixels, [1], [2], [3] of them with 1, 2, 3 inexact coordinates, [4] residual pixels
__global__ void __launch_bounds__(NT, 2) normals_cov_kernel(const float4 *__restrict__ deh, int W, int H, CovK *__restrict__ out,
                                                            unsigned long long *__restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char nrm_smem[];
    Smem &s = *reinterpret_cast<Smem *>(nrm_smem);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int bx = blockIdx.x * TX, by = blockIdx.y * TY;
    constexpr int NLD = (TW * TH + NT - 1) / NT;
    float4 ld[NLD];
    {
        // load the tile; meanwhile find the signed range of every coordinate over its valid points
        const float inf = __int_as_float(0x7f800000);
        float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
        int flags = 0;       // 1: a valid point, 2: NaN / inf, 4: a valid OUTPUT pixel, 8: an invalid entry inside the image
#pragma unroll
        for (int k = 0; k < NLD; k++) {                 // all loads of a thread in flight together
            const int i = tid + k * NT;
            const int ty = i / TW, tx = i % TW;
            const int gx = bx + tx - R, gy = by + ty - R;
            ld[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < TW * TH && gx >= 0 && gx < W && gy >= 0 && gy < H) ld[k] = __ldg(deh + (size_t)gy * W + gx);
        }
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const int i = tid + k * NT;
            if (i >= TW * TH) break;
            const int ty = i / TW, tx = i % TW;
            const float4 v = ld[k];
            if (v.w != 0.f) {
                flags |= 1;
                if (ty >= R && ty < R + TY && tx >= R && tx < R + TX) flags |= 4;
                const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fabsf(v.z));
                if (!(m < 3.0e38f)) flags |= 2;
                lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
                lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
                lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
            } else
                flags |= 8;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            flags |= __shfl_xor_sync(0xffffffffu, flags, d);
#pragma unroll
            for (int q = 0; q < 3; q++) {
                lo[q] = fminf(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], d));
                hi[q] = fmaxf(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], d));
            }
        }
        if (lane == 0) {
            s.flags[wid] = flags;
#pragma unroll
            for (int q = 0; q < 3; q++) { s.lo[wid][q] = lo[q]; s.hi[wid][q] = hi[q]; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int flags = 0;
        float lo[3], hi[3];
        for (int q = 0; q < 3; q++) { lo[q] = s.lo[0][q]; hi[q] = s.hi[0][q]; }
        for (int w = 0; w < NT / 32; w++) {
            flags |= s.flags[w];
            for (int q = 0; q < 3; q++) { lo[q] = fminf(lo[q], s.lo[w][q]); hi[q] = fmaxf(hi[q], s.hi[w][q]); }
        }
        TileInfo ti;
        bool exact[3];
        int ea[3];
        for (int q = 0; q < 3; q++) {
            const float a = fminf(fabsf(lo[q]), fabsf(hi[q])), b = fmaxf(fabsf(lo[q]), fabsf(hi[q]));
            const bool same_sign = (lo[q] > 0.f && hi[q] > 0.f) || (lo[q] < 0.f && hi[q] < 0.f);
            ea[q] = (int)((__float_as_uint(a) >> 23) & 0xff) - 127;
            exact[q] = (flags & 1) && !(flags & 2) && same_sign && b <= 1.999f * a && ea[q] >= -60 && ea[q] <= 60;
        }
        int k = 0;
        for (int q = 0; q < 3; q++) if (exact[q]) ti.perm[k++] = q;
        ti.ns = 3 - k;
        for (int q = 0; q < 3; q++) if (!exact[q]) ti.perm[k++] = q;
        for (int sl = 0; sl < 3; sl++) {
            const int q = ti.perm[sl];
            if (sl < 3 - ti.ns) {
                ti.c[sl] = fabsf(lo[q]) <= fabsf(hi[q]) ? lo[q] : hi[q];
                ti.e[sl] = ea[q] - 24;
                ti.scale[sl] = __uint_as_float((unsigned)(127 - ti.e[sl]) << 23);
                ti.unit[sl] = __uint_as_float((unsigned)(127 + ti.e[sl]) << 23);
            } else { ti.c[sl] = 0.f; ti.e[sl] = 0; ti.scale[sl] = 1.f; ti.unit[sl] = 1.f; }
        }
        // entries outside the image are zeros as well, but nothing distinguishes them from invalid ones in the loops
        ti.allvalid = !(flags & 8);
        s.ti = ti;
        s.flags[0] = flags;
        if ((flags & 4) && stats) {
            atomicAdd(stats + 0, 1ull);
            if (ti.ns) atomicAdd(stats + ti.ns, 1ull);
        }
    }
    __syncthreads();
    if (!(s.flags[0] & 4)) return;                  // no valid output pixel in this block: nothing reads its CovK
    const TileInfo ti = s.ti;
#pragma unroll
    for (int k = 0; k < NLD; k++) {                 // the tile, coordinates permuted (exact ones first)
        const int i = tid + k * NT;
        if (i >= TW * TH) break;
        const float4 v = ld[k];
        s.tile[i / TW][i % TW] = make_float4(slot(v, ti.perm[0]), slot(v, ti.perm[1]), slot(v, ti.perm[2]), v.w);
    }
    __syncthreads();
    // ---- K (and the moments of slot 0 if it is exact) ------------------------------------------------------------------
    long long A[VR][3];
    {
        const bool f0 = ti.ns <= 2;
        const float c0 = ti.c[0], sc0 = ti.scale[0];
        box_pass<3>(s, tid, [&](int r, int c, long long *m) {
            const float4 v = s.tile[r][c];
            const bool valid = v.w != 0.f;
            const long long n0 = (valid && f0) ? __float2int_rn((v.x - c0) * sc0) : 0;
            m[0] = valid ? 1 : 0; m[1] = n0; m[2] = n0 * n0; }, A);
    }
    // ---- mean: float sums in window scan order (row-major inside the window), MRUN adjacent pixels per thread.  Packed
    // f32x2 adds (two IEEE round-to-nearest additions per instruction): (x, y) of one pixel, z of two adjacent pixels;
    // where only one of the pair uses the sample the other adds 0, which changes nothing ------------------------------
    if (tid < TY * (TX / MRUN)) {
        static_assert(MRUN == 4, "z chains are packed in pixel pairs");
        const int r = tid % TY, c0 = (tid / TY) * MRUN;
        float2 sxy[MRUN], sz[MRUN / 2];
#pragma unroll
        for (int j = 0; j < MRUN; j++) sxy[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < MRUN / 2; j++) sz[j] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int dy = 0; dy < WIN; dy++) {
            const float4 *trow = &s.tile[r + dy][c0];
#pragma unroll
            for (int dx = 0; dx < WIN + MRUN - 1; dx++) {
                const float4 v = trow[dx];           // invalid entries are all-zero: x + 0 == x
#pragma unroll
                for (int j = 0; j < MRUN; j++)
                    if (dx - j >= 0 && dx - j < WIN) sxy[j] = __fadd2_rn(sxy[j], make_float2(v.x, v.y));
#pragma unroll
                for (int j = 0; j < MRUN; j += 2) {
                    const bool u0 = dx - j >= 0 && dx - j < WIN, u1 = dx - j - 1 >= 0 && dx - j - 1 < WIN;
                    if (u0 || u1) sz[j / 2] = __fadd2_rn(sz[j / 2], make_float2(u0 ? v.z : 0.f, u1 ? v.z : 0.f));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MRUN; j++) s.mean[r][c0 + j] = make_float4(sxy[j].x, sxy[j].y, (j & 1) ? sz[j / 2].y : sz[j / 2].x, 0.f);
    }
    __syncthreads();
    switch (ti.ns) {
    case 0: finish_tile<0>(s, ti, tid, bx, by, W, H, A, out, stats); break;
    case 1: finish_tile<1>(s, ti, tid, bx, by, W, H, A, out, stats); break;
    case 2: finish_tile<2>(s, ti, tid, bx, by, W, H, A, out, stats); break;
    default: finish_tile<3>(s, ti, tid, bx, by, W, H, A, out, stats); break;
    }
}

}  // namespace

int k_normals_cov(mr_context *ctx, const float4 *d_deh, CovK *d_covk)
{
    const bool stats_new = !ctx->bufs.count("nrm_stats");
    unsigned long long *stats = mr_buf<unsigned long long>(ctx, "nrm_stats", 8);
    if (!stats) return mr_fail(ctx, MR_ENOMEM, "nrm_stats", "alloc");
    if (stats_new) MR_CUDA(ctx, cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    MR_CUDA(ctx, mr_ensure_smem(ctx, normals_cov_kernel, sizeof(Smem)));
    dim3 ng(cdiv(ctx->W, TX), cdiv(ctx->H, TY));
    normals_cov_kernel<<<ng, NT, sizeof(Smem), ctx->stream>>>(d_deh, ctx->W, ctx->H, d_covk, stats);
    MR_LAUNCH_CHECK(ctx, "normals_cov_kernel");
    return MR_OK;
}
