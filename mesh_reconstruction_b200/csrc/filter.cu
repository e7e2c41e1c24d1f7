// filter.cu -- CUDA replacement of Heuristic::filterPoints (heuristic.cpp:55-176), the step that consumes the
// all-gathered point cloud right after the hot path: outlier / redundancy filter by local density.
//
//   dehomogenize (util.cpp:16-29)                                 -> dehom_kernel
//   neighbour table, j < i, FLANN result order (heuristic.cpp:70-98) -> uniform grid (cell sort) + count / fill kernels +
//                                                                    per-point (distance, index) sort; the table is also
//                                                                    built transposed (i > k, ascending i) so that every
//                                                                    score can be GATHERED in the reference's order
//   clamped power iteration (heuristic.cpp:104-138)               -> spmv_kernel + seqsum + normalize_kernel per iteration
//   greedy thinning (heuristic.cpp:140-163)                       -> radix sort of (density desc, index desc) + interval
//                                                                    rounds (thin_round_kernel)
//   compaction (heuristic.cpp:165-175)                            -> scan + gather
//
// Everything is bit-identical to oracle/filter_oracle.cpp, including the two accumulations the reference does in
// double over ALL edges / points in sequence (`sum`, `change`): seqsum() reproduces a left-to-right double accumulation
// of non-negative float terms exactly, in parallel -- while the running sum stays inside one binade, adding x is
// "add round(x / ulp) with ties decided by the parity of the running mantissa", an associative map on (parity ->
// increment) that block reductions compose; the few chunks in which the sum crosses a power of two are added term by
// term.
//
// Definitions the reference leaves to its libraries: the neighbour set is the EXACT
// radius set (the reference's FLANN search is randomised and approximate); equal densities are visited in descending
// index order (cv::sortIdx leaves their order to std::sort).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int CELL_BIAS = 1 << 20;
constexpr int CH = 2048;          // terms per chunk of seqsum
constexpr int CHT = 256;          // threads per chunk block (8 terms each)

// ---------------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------------
__global__ void dehom_kernel(const float *__restrict__ pts, int stride, int n, float4 *__restrict__ p3)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = pts + (size_t)i * stride;
    const float w = p[3];
    p3[i] = make_float4(p[0] / w, p[1] / w, p[2] / w, 0.f);      // util.cpp:24-26
}

__device__ __forceinline__ bool cell_of(const float4 &p, double inv_cell, int *c)
{
    const double v[3] = {floor((double)p.x * inv_cell), floor((double)p.y * inv_cell), floor((double)p.z * inv_cell)};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!(fabs(v[k]) < 1.0e6)) return false;                 // NaN / inf / absurdly far: the point has no neighbours
        c[k] = (int)v[k] + CELL_BIAS;
    }
    return true;
}
__device__ __forceinline__ u64 cell_key(int cx, int cy, int cz) { return ((u64)cx << 42) | ((u64)cy << 21) | (u64)cz; }

__global__ void cell_key_kernel(const float4 *__restrict__ p3, int n, double inv_cell, u64 *__restrict__ keys, int *__restrict__ idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
    keys[i] = cell_of(p3[i], inv_cell, c) ? cell_key(c[0], c[1], c[2]) : ~0ull;
    idx[i] = i;
}

// FLANN's L2_Simple<float>: float accumulation of the squared distance, x then y then z
__device__ __forceinline__ float l2_simple(const float4 &a, const float4 &b)
{
    float result = 0.f, diff;
    diff = a.x - b.x; result = result + diff * diff;
    diff = a.y - b.y; result = result + diff * diff;
    diff = a.z - b.z; result = result + diff * diff;
    return result;
}

__device__ __forceinline__ int lower_bound_u64(const u64 *__restrict__ a, int n, u64 key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One thread per point (in cell order, so that neighbouring threads walk the same cells).  FILL == false: count the
// neighbours with a smaller / larger index.  FILL == true: write their sort keys at the point's offsets:
//   lower block: (distance bits << 32) | j   -> ascending (distance, index), FLANN's result order
//   upper block: (j << 32) | distance bits   -> ascending index
template <bool FILL>
__global__ void __launch_bounds__(128) neighbour_kernel(const float4 *__restrict__ p3, const u64 *__restrict__ skeys, const int *__restrict__ sidx,
                                                        int n, double inv_cell, float radius, int *__restrict__ cntL, int *__restrict__ cntU,
                                                        const long long *__restrict__ offL, const long long *__restrict__ offU,
                                                        u64 *__restrict__ keyL, u64 *__restrict__ keyU)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = sidx[t];
    const float4 pi = p3[i];
    int c[3];
    int nl = 0, nu = 0;
    long long ol = 0, ou = 0;
    if (FILL) { ol = offL[i]; ou = offU[i]; }
    if (cell_of(pi, inv_cell, c)) {
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++) {
                // the three cells (cz - 1 .. cz + 1) of one (cx, cy) column are adjacent key values: one search, one run
                const u64 k0 = cell_key(c[0] + dx, c[1] + dy, c[2] - 1), k1 = cell_key(c[0] + dx, c[1] + dy, c[2] + 1);
                for (int s = lower_bound_u64(skeys, n, k0); s < n && skeys[s] <= k1; s++) {
                    const int j = sidx[s];
                    if (j == i) continue;
                    const float d = l2_simple(pi, p3[j]);
                    if (!(d <= radius)) continue;                       // heuristic.cpp:90 (and FLANN's own test)
                    if (j < i) {
                        if (FILL) keyL[ol + nl] = ((u64)__float_as_uint(d) << 32) | (unsigned)j;
                        nl++;
                    } else {
                        if (FILL) keyU[ou + nu] = ((u64)(unsigned)j << 32) | __float_as_uint(d);
                        nu++;
                    }
                }
            }
    }
    if (!FILL) { cntL[i] = nl; cntU[i] = nu; }
}

// densityFn(dist, radius) = (float)(1. - dist / radius)   heuristic.cpp:49-52
__global__ void unpack_edges_kernel(const u64 *__restrict__ keyL, const u64 *__restrict__ keyU, long long E, float radius,
                                    int *__restrict__ nbL, float *__restrict__ wL, int *__restrict__ nbU, float *__restrict__ wU)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const u64 a = keyL[e], b = keyU[e];
    const float da = __uint_as_float((unsigned)(a >> 32)), db = __uint_as_float((unsigned)b);
    nbL[e] = (int)(unsigned)a;
    wL[e] = (float)(1.0 - (double)(da / radius));
    nbU[e] = (int)(unsigned)(b >> 32);
    wU[e] = (float)(1.0 - (double)(db / radius));
}

// ---------------------------------------------------------------------------------------------------------------------
// power iteration
// ---------------------------------------------------------------------------------------------------------------------
struct IterState {
    double sum, change;
    int done;            // the reference's loop has ended: later iterations that were already queued do nothing
    int iters;
    int nonfinite;
    int slow_chunks;     // diagnostics: chunks seqsum added term by term
};

// score[i] exactly as the reference's scatter loop leaves it (heuristic.cpp:110-124): first the gathered densityTemp of
// the point's own block, then the contributions of the later points that have it as a neighbour, in ascending order of
// those points.  terms[e] = the float added to the double `sum` for edge e (global edge order == the reference's).
__global__ void __launch_bounds__(128) spmv_kernel(int n, const long long *__restrict__ offL, const int *__restrict__ nbL, const float *__restrict__ wL,
                                                   const long long *__restrict__ offU, const int *__restrict__ nbU, const float *__restrict__ wU,
                                                   const float *__restrict__ density, float *__restrict__ score, float *__restrict__ terms,
                                                   const IterState *__restrict__ st)
{
    if (st->done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float di = density[i];
    float t = 0.f;
    for (long long e = offL[i], e1 = offL[i + 1]; e < e1; e++) {
        const float dn = density[nbL[e]], w = wL[e];
        t = t + dn * w;
        terms[e] = (di + dn) * w;
    }
    float s = 0.f + t;
    for (long long e = offU[i], e1 = offU[i + 1]; e < e1; e++) s = s + density[nbU[e]] * wU[e];
    score[i] = s;
}

__global__ void normalize_kernel(int n, const float *__restrict__ score, float *__restrict__ density, float *__restrict__ cterms,
                                 const IterState *__restrict__ st)
{
    if (st->done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float normalizer = (float)((double)n / st->sum);                 // float normalizer = pointCount / sum;
    float nd = score[i] * normalizer;
    if ((double)nd > 2.) nd = 2.f;
    const float diff = density[i] - nd;
    cterms[i] = diff * diff;                                                // pow2()
    density[i] = nd;
}

__global__ void iter_end_kernel(int n, int max_iters, IterState *st)
{
    if (st->done) return;
    const double change = st->change / (double)n;
    st->iters += 1;
    if (!(change > 1e-6 && st->iters < max_iters)) st->done = 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// seqsum: left-to-right double accumulation of non-negative float terms, exactly
// ---------------------------------------------------------------------------------------------------------------------
struct IncMap { long long i0, i1; };     // increment of the running mantissa for even / odd incoming parity

__device__ __forceinline__ IncMap compose(const IncMap &a, const IncMap &b)      // a first, then b
{
    IncMap r;
    r.i0 = a.i0 + ((a.i0 & 1) ? b.i1 : b.i0);
    r.i1 = a.i1 + (((1 + a.i1) & 1) ? b.i1 : b.i0);
    return r;
}

// map of adding the float x to a double whose exponent is e (unit u = 2^(e - 52)); requires x < 2^(e + 1)
__device__ __forceinline__ IncMap term_map(float x, int e)
{
    IncMap r;
    r.i0 = r.i1 = 0;
    const unsigned b = __float_as_uint(x);
    const int ex = (int)((b >> 23) & 0xff);
    long long m = b & 0x7fffff;
    int lsb;                                   // x = m * 2^lsb
    if (ex) { m |= 0x800000; lsb = ex - 150; } else lsb = -149;
    if (m == 0) return r;
    const int sh = lsb - (e - 52);
    if (sh >= 0) { r.i0 = r.i1 = m << sh; return r; }
    const int t = -sh;
    if (t > 25) return r;                      // x < u / 4
    const long long k = m >> t, rem = m & ((1ll << t) - 1), half = 1ll << (t - 1);
    if (rem > half) r.i0 = r.i1 = k + 1;
    else if (rem < half) r.i0 = r.i1 = k;
    else { r.i0 = k + (k & 1); r.i1 = k + ((k + 1) & 1); }        // tie: to the even mantissa
    return r;
}

__global__ void __launch_bounds__(CHT) chunk_sum_kernel(const float *__restrict__ terms, long long n, double *__restrict__ csum,
                                                        const IterState *__restrict__ st, int *__restrict__ nonfinite)
{
    if (st && st->done) return;
    __shared__ double sh[CHT / 32];
    const long long base = (long long)blockIdx.x * CH;
    double s = 0;
    bool bad = false;
    for (int k = threadIdx.x; k < CH; k += CHT) {
        const long long e = base + k;
        if (e < n) {
            const float x = terms[e];
            if (!(x >= 0.f && x < 3.0e38f)) bad = true;
            s += (double)x;
        }
    }
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    if (bad) *nonfinite = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < CHT / 32; w++) t += sh[w];
        csum[blockIdx.x] = t;
    }
}

// approximate prefix of the chunk sums (one block) and the exponent every chunk may assume for the running sum, or
// INT_MIN where the sum may cross a power of two inside the chunk (any summation order of n <= 2^31 non-negative terms
// is within 2^-22 relative of any other: the margin used is 2^-20)
__global__ void __launch_bounds__(1024) chunk_class_kernel(const double *__restrict__ csum, int nchunks, int *__restrict__ cexp,
                                                           const IterState *__restrict__ st)
{
    if (st && st->done) return;
    __shared__ double wsum[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int c = base + threadIdx.x;
        const double v = c < nchunks ? csum[c] : 0.0;
        double incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        double pre = carry_s;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) pre += wsum[w];
        const double start = pre + (incl - v), end = pre + incl;
        if (c < nchunks) {
            const double lo = start * (1.0 - 9.5367431640625e-07), hi = end * (1.0 + 9.5367431640625e-07);
            int e = INT_MIN;
            if (lo > 0.0) {
                const int elo = (int)((__double_as_longlong(lo) >> 52) & 0x7ff), ehi = (int)((__double_as_longlong(hi) >> 52) & 0x7ff);
                if (elo == ehi && elo > 0 && elo < 0x7ff) e = elo - 1023;
            }
            cexp[c] = e;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = end;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CHT) chunk_map_kernel(const float *__restrict__ terms, long long n, const int *__restrict__ cexp,
                                                        IncMap *__restrict__ cmap, const IterState *__restrict__ st)
{
    if (st && st->done) return;
    const int e = cexp[blockIdx.x];
    if (e == INT_MIN) return;
    __shared__ IncMap sh[CHT / 32];
    const long long base = (long long)blockIdx.x * CH + (long long)threadIdx.x * (CH / CHT);
    IncMap m;
    m.i0 = m.i1 = 0;
#pragma unroll
    for (int k = 0; k < CH / CHT; k++) {
        const long long idx = base + k;
        if (idx < n) m = compose(m, term_map(terms[idx], e));
    }
    for (int d = 1; d < 32; d <<= 1) {            // ordered tree: lane L takes the composition of the next d lanes
        IncMap r;
        r.i0 = __shfl_down_sync(0xffffffffu, m.i0, d);
        r.i1 = __shfl_down_sync(0xffffffffu, m.i1, d);
        if (((threadIdx.x & 31) % (2 * d)) == 0) m = compose(m, r);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        IncMap t = sh[0];
        for (int w = 1; w < CHT / 32; w++) t = compose(t, sh[w]);
        cmap[blockIdx.x] = t;
    }
}

// one warp walks the chunks in order with the exact running sum
__global__ void __launch_bounds__(32) chunk_walk_kernel(const float *__restrict__ terms, long long n, int nchunks, const int *__restrict__ cexp,
                                                        const IncMap *__restrict__ cmap, const int *__restrict__ nonfinite, double *__restrict__ out,
                                                        const IterState *__restrict__ st)
{
    if (st && st->done) return;
    __shared__ float buf[CH + 32];          // 32 sub-chunks of 64 terms at a pitch of 65 (conflict-free lane-private walks)
    const int lane = threadIdx.x;
    double s = 0.0;
    if (*nonfinite) {
        // a NaN / inf / negative term (never produced by finite clouds): plain sequential addition
        if (lane == 0) {
            for (long long e = 0; e < n; e++) s += (double)terms[e];
            *out = s;
        }
        return;
    }
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
        const int c = c0 + lane;
        const int e_l = c < nchunks ? cexp[c] : 0;
        IncMap m_l;
        m_l.i0 = m_l.i1 = 0;
        if (c < nchunks && e_l != INT_MIN) m_l = cmap[c];
        {
            // common case: the 32 chunks of this group all sit in the binade of the running sum -> compose their maps
            // with an ordered shuffle tree and apply the composite in one step
            const long long bits = __double_as_longlong(s);
            const int es = (int)((bits >> 52) & 0x7ff) - 1023;
            const int e_mine = c < nchunks ? e_l : es;          // lanes past the end: identity map in the same binade
            if (s > 0.0 && __all_sync(0xffffffffu, e_mine == es && e_mine != INT_MIN)) {
                IncMap t = m_l;
                for (int d = 1; d < 32; d <<= 1) {
                    IncMap r;
                    r.i0 = __shfl_down_sync(0xffffffffu, t.i0, d);
                    r.i1 = __shfl_down_sync(0xffffffffu, t.i1, d);
                    if ((lane % (2 * d)) == 0) t = compose(t, r);
                }
                const long long g0 = __shfl_sync(0xffffffffu, t.i0, 0), g1 = __shfl_sync(0xffffffffu, t.i1, 0);
                const long long mant = (bits & 0xfffffffffffffll) | (1ll << 52);
                const long long m2 = mant + ((mant & 1) ? g1 : g0);
                if (m2 < (1ll << 53)) {
                    s = __longlong_as_double((bits & ~0xfffffffffffffll) | (m2 & 0xfffffffffffffll));
                    continue;
                }
            }
        }
        for (int k = 0; k < 32 && c0 + k < nchunks; k++) {
            const int e = __shfl_sync(0xffffffffu, e_l, k);
            const long long i0 = __shfl_sync(0xffffffffu, m_l.i0, k), i1 = __shfl_sync(0xffffffffu, m_l.i1, k);
            const long long bits = __double_as_longlong(s);
            const int es = (int)((bits >> 52) & 0x7ff) - 1023;
            bool fast = false;
            if (e != INT_MIN && es == e && s > 0.0) {
                const long long mant = (bits & 0xfffffffffffffll) | (1ll << 52);
                const long long inc = (mant & 1) ? i1 : i0;
                const long long m2 = mant + inc;
                if (m2 < (1ll << 53)) {                 // stays in the binade (guaranteed by the classification; checked)
                    s = __longlong_as_double((bits & ~0xfffffffffffffll) | (m2 & 0xfffffffffffffll));
                    fast = true;
                }
            }
            if (!fast) {
                // The running sum may cross a power of two inside this chunk.  Same idea one level down: 32 sub-chunks of
                // 64 terms, one per lane -- approximate prefix inside the chunk, exponent per sub-chunk, increment maps for
                // the stable ones (computed by all lanes at once), then the lanes are visited in order and only the
                // sub-chunk(s) that really contain the crossing are added term by term.
                if (lane == 0 && st) atomicAdd((int *)&st->slow_chunks, 1);
                constexpr int SUB = CH / 32;
                const long long base = (long long)(c0 + k) * CH;
                for (int q = lane; q < CH; q += 32) buf[(q / SUB) * (SUB + 1) + q % SUB] = base + q < n ? terms[base + q] : 0.f;
                __syncwarp();
                const float *mine = buf + lane * (SUB + 1);
                double part = 0.0;
                for (int q = 0; q < SUB; q++) part += (double)mine[q];
                double incl = part;
                for (int d = 1; d < 32; d <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                const double lo = (s + (incl - part)) * (1.0 - 9.5367431640625e-07), hi = (s + incl) * (1.0 + 9.5367431640625e-07);
                int e2 = INT_MIN;
                if (lo > 0.0) {
                    const int elo = (int)((__double_as_longlong(lo) >> 52) & 0x7ff), ehi = (int)((__double_as_longlong(hi) >> 52) & 0x7ff);
                    if (elo == ehi && elo > 0 && elo < 0x7ff) e2 = elo - 1023;
                }
                IncMap m2;
                m2.i0 = m2.i1 = 0;
                if (e2 != INT_MIN)
                    for (int q = 0; q < SUB; q++) m2 = compose(m2, term_map(mine[q], e2));
                for (int l = 0; l < 32; l++) {
                    const int el = __shfl_sync(0xffffffffu, e2, l);
                    const long long j0 = __shfl_sync(0xffffffffu, m2.i0, l), j1 = __shfl_sync(0xffffffffu, m2.i1, l);
                    const long long b2 = __double_as_longlong(s);
                    const int es2 = (int)((b2 >> 52) & 0x7ff) - 1023;
                    bool ok2 = false;
                    if (el != INT_MIN && es2 == el && s > 0.0) {
                        const long long mant = (b2 & 0xfffffffffffffll) | (1ll << 52);
                        const long long mm = mant + ((mant & 1) ? j1 : j0);
                        if (mm < (1ll << 53)) {
                            s = __longlong_as_double((b2 & ~0xfffffffffffffll) | (mm & 0xfffffffffffffll));
                            ok2 = true;
                        }
                    }
                    if (!ok2) {
                        double a = s;
                        if (lane == l)
                            for (int q = 0; q < SUB; q++) a += (double)mine[q];
                        s = __shfl_sync(0xffffffffu, a, l);
                    }
                }
                __syncwarp();
            }
        }
    }
    if (lane == 0) *out = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// thinning
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned orderable(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void rank_key_kernel(int n, const float *__restrict__ density, u64 *__restrict__ keys, int *__restrict__ idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((u64)(0xffffffffu - orderable(density[i])) << 32) | (u64)(0xffffffffu - (unsigned)i);   // density desc, index desc
    idx[i] = i;
}
__global__ void rank_scatter_kernel(int n, const int *__restrict__ order, int *__restrict__ rank)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) rank[order[r]] = r;
}
// upper edges keyed by the rank of the later point, so that every point can replay the subtractions it receives in the
// order the reference applies them
__global__ void upper_rank_key_kernel(long long E, const int *__restrict__ nbU, const int *__restrict__ rank, u64 *__restrict__ keys)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) keys[e] = ((u64)(unsigned)rank[nbU[e]] << 32) | (u64)(unsigned)e;
}

enum { UNDECIDED = 0, ACCEPTED = 1, REJECTED = 2 };

// One interval round.  For an undecided point p the final score is score[p] minus (density[q] * w) for every ACCEPTED
// later-index neighbour q that is visited before p, applied in visiting order as (float)((double)score - product).
// Subtracting more can only lower the result (float subtraction is monotone), so replaying the predecessors that are
// not known to be rejected gives a lower bound L, replaying only the accepted ones an upper bound U:
//   !(L < 0.7) -> accepted whatever the undecided ones turn out to be;  U < 0.7 -> rejected;  else wait.
// With every predecessor decided L == U is the reference's value.  The first undecided point in visiting order has no
// undecided predecessor, so every round decides at least one point.
__global__ void __launch_bounds__(128) thin_round_kernel(int n_work, const int *__restrict__ work, const long long *__restrict__ offU,
                                                         const u64 *__restrict__ rkeys, const int *__restrict__ nbU, const float *__restrict__ wU,
                                                         const int *__restrict__ rank, const float *__restrict__ density, const float *__restrict__ score,
                                                         int *state, int *__restrict__ next_work, int *__restrict__ n_next)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_work) return;
    const int p = work[t];
    const unsigned rp = (unsigned)rank[p];
    float L = score[p], U = L;
    for (long long e = offU[p], e1 = offU[p + 1]; e < e1; e++) {
        const u64 k = rkeys[e];
        if ((unsigned)(k >> 32) >= rp) break;                   // sorted by rank: the rest is visited after p
        const long long src = (long long)(unsigned)k;
        const int q = nbU[src];
        const int sq = ((volatile int *)state)[q];
        if (sq == REJECTED) continue;
        const double prod = (double)density[q] * (double)wU[src];     // localDensity * neighbors[j].second
        L = (float)((double)L - prod);
        if (sq == ACCEPTED) U = (float)((double)U - prod);
    }
    const float limit = 0.7f;
    int s = UNDECIDED;
    if (!(L < limit)) s = ACCEPTED;
    else if (U < limit) s = REJECTED;
    if (s != UNDECIDED) state[p] = s;
    else next_work[atomicAdd(n_next, 1)] = p;
}

__global__ void fill_kernel(int n, float v, float *a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
__global__ void iota_kernel(int n, int *a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void keep_flag_kernel(int n, const int *__restrict__ state, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = state[i] == ACCEPTED;
}
__global__ void gather_kernel(int n, const int *__restrict__ flag, const int *__restrict__ pos, const float *__restrict__ pts, int pstride,
                              const float *__restrict__ nrm, int nstride, float *__restrict__ out_pts, int opstride, float *__restrict__ out_nrm,
                              int onstride, int *__restrict__ out_keep, int *__restrict__ out_count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) *out_count = pos[i] + flag[i];
    if (!flag[i]) return;
    const int o = pos[i];
    if (out_keep) out_keep[o] = i;
    if (out_pts)
        for (int k = 0; k < 4; k++) out_pts[(size_t)o * opstride + k] = pts[(size_t)i * pstride + k];
    if (out_nrm && nrm)
        for (int k = 0; k < 3; k++) out_nrm[(size_t)o * onstride + k] = nrm[(size_t)i * nstride + k];
}

#define FL_LAUNCH(ctx, name) MR_LAUNCH_CHECK(ctx, name)

// MR_FILTER_TIMING=1: phase times on stderr (synchronises; diagnostics only)
struct PhaseTimer {
    mr_context *ctx;
    bool on;
    cudaEvent_t a, b;
    PhaseTimer(mr_context *c) : ctx(c), on(getenv("MR_FILTER_TIMING") != nullptr)
    {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream); }
    }
    void lap(const char *what)
    {
        if (!on) return;
        cudaEventRecord(b, ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        fprintf(stderr, "[mr_filter] %-28s %9.3f ms\n", what, ms);
        cudaEventRecord(a, ctx->stream);
    }
    ~PhaseTimer() { if (on) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};

int seqsum(mr_context *ctx, const float *terms, long long n, double *d_out, const IterState *st, double *csum, int *cexp, IncMap *cmap, int *nonfinite)
{
    const int nchunks = (int)((n + CH - 1) / CH);
    MR_CUDA(ctx, cudaMemsetAsync(nonfinite, 0, sizeof(int), ctx->stream));
    if (nchunks == 0) {
        MR_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream));
        return MR_OK;
    }
    chunk_sum_kernel<<<nchunks, CHT, 0, ctx->stream>>>(terms, n, csum, st, nonfinite);
    FL_LAUNCH(ctx, "chunk_sum_kernel");
    chunk_class_kernel<<<1, 1024, 0, ctx->stream>>>(csum, nchunks, cexp, st);
    FL_LAUNCH(ctx, "chunk_class_kernel");
    chunk_map_kernel<<<nchunks, CHT, 0, ctx->stream>>>(terms, n, cexp, cmap, st);
    FL_LAUNCH(ctx, "chunk_map_kernel");
    chunk_walk_kernel<<<1, 32, 0, ctx->stream>>>(terms, n, nchunks, cexp, cmap, nonfinite, d_out, st);
    FL_LAUNCH(ctx, "chunk_walk_kernel");
    return MR_OK;
}

}  // namespace

// test hook: exact sequential double sum of non-negative float terms (device memory)
int k_seqsum(mr_context *ctx, const float *d_terms, long long n, double *h_out)
{
    const int nchunks = (int)((n + CH - 1) / CH) + 1;
    double *csum = mr_buf<double>(ctx, "fl_csum", (size_t)nchunks + 2);
    int *cexp = mr_buf<int>(ctx, "fl_cexp", (size_t)nchunks + 2);
    IncMap *cmap = mr_buf<IncMap>(ctx, "fl_cmap", (size_t)nchunks);
    if (!csum || !cexp || !cmap) return mr_fail(ctx, MR_ENOMEM, "seqsum", "alloc");
    int rc = seqsum(ctx, d_terms, n, csum + nchunks, nullptr, csum, cexp, cmap, cexp + nchunks);
    if (rc) return rc;
    MR_CUDA(ctx, cudaMemcpyAsync(h_out, csum + nchunks, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MR_OK;
}

struct IntToI64 {
    __host__ __device__ long long operator()(int v) const { return (long long)v; }
};

#define CUB_CALL(ctx, expr)                                                         \
    do {                                                                            \
        size_t bytes__ = 0;                                                         \
        void *tmp__ = nullptr;                                                      \
        MR_CUDA(ctx, (expr));                                                       \
        tmp__ = mr_buf_raw(ctx, "fl_cub_tmp", bytes__ ? bytes__ : 1);               \
        if (!tmp__) return mr_fail(ctx, MR_ENOMEM, "fl_cub_tmp", "alloc");          \
        MR_CUDA(ctx, (expr));                                                       \
        ctx->launches++;                                                            \
    } while (0)

// d_pts: n rows of >= 4 floats (stride pstride), d_nrm: optional n rows of >= 3 floats (stride nstride); outputs on the
// device (any may be null except d_count).  info: [0] edges, [1] power iterations, [2] thinning rounds.
int k_filter_points(mr_context *ctx, const float *d_pts, int pstride, const float *d_nrm, int nstride, int n, float radius, float *d_out_pts,
                    int opstride, float *d_out_nrm, int onstride, int *d_out_keep, int *h_count, long long *info)
{
    cudaStream_t st = ctx->stream;
    if (info) info[0] = info[1] = info[2] = 0;
    *h_count = 0;
    if (n == 0) return MR_OK;
    const int T = 256, G = cdiv(n, T);
    float4 *p3 = mr_buf<float4>(ctx, "fl_p3", (size_t)n);
    u64 *ckey = mr_buf<u64>(ctx, "fl_ckey", 2 * (size_t)n);
    int *cidx = mr_buf<int>(ctx, "fl_cidx", 2 * (size_t)n);
    int *cnt = mr_buf<int>(ctx, "fl_cnt", 2 * (size_t)n);
    long long *off = mr_buf<long long>(ctx, "fl_off", 2 * ((size_t)n + 1));
    float *density = mr_buf<float>(ctx, "fl_density", (size_t)n), *score = mr_buf<float>(ctx, "fl_score", (size_t)n);
    float *cterms = mr_buf<float>(ctx, "fl_cterms", (size_t)n);
    IterState *its = mr_buf<IterState>(ctx, "fl_iter", 1);
    if (!p3 || !ckey || !cidx || !cnt || !off || !density || !score || !cterms || !its) return mr_fail(ctx, MR_ENOMEM, "filter", "alloc");
    u64 *skey = ckey + n;
    int *sidx = cidx + n, *cntL = cnt, *cntU = cnt + n;
    long long *offL = off, *offU = off + (n + 1);

    PhaseTimer tm(ctx);
    dehom_kernel<<<G, T, 0, st>>>(d_pts, pstride, n, p3);
    FL_LAUNCH(ctx, "dehom_kernel");
    // cell edge a little above the Euclidean radius (`radius` bounds SQUARED distances); the grid only narrows the
    // candidates, acceptance is the float comparison of the reference
    const double cell = std::sqrt((double)radius) * 1.0001 + 1e-30;
    const double inv_cell = (radius >= 0.f && std::isfinite(cell)) ? 1.0 / cell : 0.0;
    cell_key_kernel<<<G, T, 0, st>>>(p3, n, inv_cell, ckey, cidx);
    FL_LAUNCH(ctx, "cell_key_kernel");
    CUB_CALL(ctx, cub::DeviceRadixSort::SortPairs(tmp__, bytes__, ckey, skey, cidx, sidx, n, 0, 63, st));
    MR_CUDA(ctx, cudaMemsetAsync(cnt, 0, 2 * (size_t)n * sizeof(int), st));
    if (radius >= 0.f && inv_cell > 0.0) {
        neighbour_kernel<false><<<cdiv(n, 128), 128, 0, st>>>(p3, skey, sidx, n, inv_cell, radius, cntL, cntU, nullptr, nullptr, nullptr, nullptr);
        FL_LAUNCH(ctx, "neighbour_kernel<count>");
    }
    tm.lap("cells + neighbour count");
    MR_CUDA(ctx, cudaMemsetAsync(off, 0, 2 * ((size_t)n + 1) * sizeof(long long), st));
    // the counts are scanned as 64-bit values: CUB accumulates in the INPUT type, and a radius that is too large for the cloud
    // (more than 2^31 pairs) must come out as a number that the check below can refuse, not as a wrapped int
    cub::TransformInputIterator<long long, IntToI64, const int *> cntL64(cntL, IntToI64()), cntU64(cntU, IntToI64());
    CUB_CALL(ctx, cub::DeviceScan::InclusiveSum(tmp__, bytes__, cntL64, offL + 1, n, st));
    CUB_CALL(ctx, cub::DeviceScan::InclusiveSum(tmp__, bytes__, cntU64, offU + 1, n, st));
    long long hE[2] = {0, 0};
    MR_CUDA(ctx, cudaMemcpyAsync(&hE[0], offL + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
    MR_CUDA(ctx, cudaMemcpyAsync(&hE[1], offU + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
    MR_CUDA(ctx, cudaStreamSynchronize(st));
    const long long E = hE[0];
    if (hE[1] != E) return mr_fail(ctx, MR_ECUDA, "filter", "neighbour table is not symmetric");
    if (E > 0x7fffffffll) return mr_fail(ctx, MR_EINVAL, "mr_filter_points", "more than 2^31 neighbour pairs (the reference's int table overflows too): use a smaller radius");
    if (info) info[0] = E;
    const size_t Ea = (size_t)(E > 0 ? E : 1);
    u64 *keyL = mr_buf<u64>(ctx, "fl_keyL", 2 * Ea), *keyU = mr_buf<u64>(ctx, "fl_keyU", 2 * Ea);
    int *nbL = mr_buf<int>(ctx, "fl_nbL", Ea), *nbU = mr_buf<int>(ctx, "fl_nbU", Ea);
    float *wL = mr_buf<float>(ctx, "fl_wL", Ea), *wU = mr_buf<float>(ctx, "fl_wU", Ea), *terms = mr_buf<float>(ctx, "fl_terms", Ea);
    const int nchunks = (int)((Ea + CH - 1) / CH) + 1;
    double *csum = mr_buf<double>(ctx, "fl_csum", (size_t)nchunks + 2);
    int *cexp = mr_buf<int>(ctx, "fl_cexp", (size_t)nchunks + 2);
    IncMap *cmap = mr_buf<IncMap>(ctx, "fl_cmap", (size_t)nchunks);
    if (!keyL || !keyU || !nbL || !nbU || !wL || !wU || !terms || !csum || !cexp || !cmap) return mr_fail(ctx, MR_ENOMEM, "filter", "alloc (edges)");
    if (E > 0) {
        neighbour_kernel<true><<<cdiv(n, 128), 128, 0, st>>>(p3, skey, sidx, n, inv_cell, radius, nullptr, nullptr, offL, offU, keyL, keyU);
        FL_LAUNCH(ctx, "neighbour_kernel<fill>");
        CUB_CALL(ctx, cub::DeviceSegmentedSort::SortKeys(tmp__, bytes__, keyL, keyL + Ea, (int)E, n, offL, offL + 1, st));
        CUB_CALL(ctx, cub::DeviceSegmentedSort::SortKeys(tmp__, bytes__, keyU, keyU + Ea, (int)E, n, offU, offU + 1, st));
        unpack_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(keyL + Ea, keyU + Ea, E, radius, nbL, wL, nbU, wU);
        FL_LAUNCH(ctx, "unpack_edges_kernel");
    }
    tm.lap("neighbour table (fill + sort)");
    // ---- power iteration ------------------------------------------------------------------------------------------
    fill_kernel<<<G, T, 0, st>>>(n, 1.f, density);                            // std::vector<float> density(pointCount, 1.)
    FL_LAUNCH(ctx, "fill_kernel");
    MR_CUDA(ctx, cudaMemsetAsync(its, 0, sizeof(IterState), st));
    const int max_iters = 200, batch = 8;
    int h_done = 0, h_iters = 0;
    int *nonfinite = cexp + nchunks;
    for (int it = 0; it < max_iters && !h_done; it += batch) {
        for (int b = 0; b < batch && it + b < max_iters; b++) {
            spmv_kernel<<<cdiv(n, 128), 128, 0, st>>>(n, offL, nbL, wL, offU, nbU, wU, density, score, terms, its);
            FL_LAUNCH(ctx, "spmv_kernel");
            int rc = seqsum(ctx, terms, E, &its->sum, its, csum, cexp, cmap, nonfinite);
            if (rc) return rc;
            normalize_kernel<<<G, T, 0, st>>>(n, score, density, cterms, its);
            FL_LAUNCH(ctx, "normalize_kernel");
            rc = seqsum(ctx, cterms, n, &its->change, its, csum, cexp, cmap, nonfinite);
            if (rc) return rc;
            iter_end_kernel<<<1, 1, 0, st>>>(n, max_iters, its);
            FL_LAUNCH(ctx, "iter_end_kernel");
        }
        IterState h;
        MR_CUDA(ctx, cudaMemcpyAsync(&h, its, sizeof(h), cudaMemcpyDeviceToHost, st));
        MR_CUDA(ctx, cudaStreamSynchronize(st));
        h_done = h.done;
        h_iters = h.iters;
        if (tm.on) fprintf(stderr, "[mr_filter] iterations %d, slow chunks so far %d\n", h.iters, h.slow_chunks);
    }
    if (info) info[1] = h_iters;
    tm.lap("power iteration");
    // ---- thinning -----------------------------------------------------------------------------------------------------
    int *rank = mr_buf<int>(ctx, "fl_rank", (size_t)n), *state = mr_buf<int>(ctx, "fl_state", (size_t)n);
    int *work = mr_buf<int>(ctx, "fl_work", 2 * (size_t)n), *nwork = mr_buf<int>(ctx, "fl_nwork", 2);
    if (!rank || !state || !work || !nwork) return mr_fail(ctx, MR_ENOMEM, "filter", "alloc (thinning)");
    rank_key_kernel<<<G, T, 0, st>>>(n, density, ckey, cidx);
    FL_LAUNCH(ctx, "rank_key_kernel");
    CUB_CALL(ctx, cub::DeviceRadixSort::SortPairs(tmp__, bytes__, ckey, skey, cidx, sidx, n, 0, 64, st));
    rank_scatter_kernel<<<G, T, 0, st>>>(n, sidx, rank);
    FL_LAUNCH(ctx, "rank_scatter_kernel");
    u64 *rkeys = keyU + Ea;
    if (E > 0) {
        upper_rank_key_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(E, nbU, rank, keyU);
        FL_LAUNCH(ctx, "upper_rank_key_kernel");
        CUB_CALL(ctx, cub::DeviceSegmentedSort::SortKeys(tmp__, bytes__, keyU, rkeys, (int)E, n, offU, offU + 1, st));
    }
    MR_CUDA(ctx, cudaMemsetAsync(state, 0, (size_t)n * sizeof(int), st));
    iota_kernel<<<G, T, 0, st>>>(n, work);
    FL_LAUNCH(ctx, "iota_kernel");
    int n_work = n, cur = 0, rounds = 0;
    while (n_work > 0) {
        MR_CUDA(ctx, cudaMemsetAsync(nwork, 0, sizeof(int), st));
        thin_round_kernel<<<cdiv(n_work, 128), 128, 0, st>>>(n_work, work + (size_t)cur * n, offU, rkeys, nbU, wU, rank, density, score, state,
                                                             work + (size_t)(cur ^ 1) * n, nwork);
        FL_LAUNCH(ctx, "thin_round_kernel");
        int h_next = 0;
        MR_CUDA(ctx, cudaMemcpyAsync(&h_next, nwork, sizeof(int), cudaMemcpyDeviceToHost, st));
        MR_CUDA(ctx, cudaStreamSynchronize(st));
        rounds++;
        if (h_next >= n_work) return mr_fail(ctx, MR_ECUDA, "mr_filter_points", "thinning made no progress (internal error)");
        n_work = h_next;
        cur ^= 1;
    }
    if (info) info[2] = rounds;
    tm.lap("thinning");
    // ---- compaction, ascending index (heuristic.cpp:165-175) -------------------------------------------------------------
    int *flag = cntL, *pos = cntU;
    keep_flag_kernel<<<G, T, 0, st>>>(n, state, flag);
    FL_LAUNCH(ctx, "keep_flag_kernel");
    CUB_CALL(ctx, cub::DeviceScan::ExclusiveSum(tmp__, bytes__, flag, pos, n, st));
    gather_kernel<<<G, T, 0, st>>>(n, flag, pos, d_pts, pstride, d_nrm, nstride, d_out_pts, opstride, d_out_nrm, onstride, d_out_keep, nwork);
    FL_LAUNCH(ctx, "gather_kernel");
    MR_CUDA(ctx, cudaMemcpyAsync(h_count, nwork, sizeof(int), cudaMemcpyDeviceToHost, st));
    MR_CUDA(ctx, cudaStreamSynchronize(st));
    tm.lap("compaction");
    return MR_OK;
}
