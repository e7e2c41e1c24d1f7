// vr_fused.cu -- fused shared-memory tile kernel for OpenCV's VariationalRefinement (flow.cpp:29,32).
//
// One launch per fixed-point iteration.  A CTA owns an OTW x OTH output tile and evaluates the whole
// iteration (derivatives -> robust data term -> smoothness weights -> 5 red-black SOR sweeps) on the
// tile plus a 10-pixel halo, so that nothing but du/dv crosses HBM between iterations:
//
//   region I (tile + 11/16): the two 8-bit frames  -> shared memory, staged by TMA (cp.async.bulk.tensor.2d
//                                                     + mbarrier).  The out-of-image halo is zero-filled by
//                                                     the TMA unit and never read (taps are clamped).
//   region D (tile + 10)   : Ix, Iy, Iz, du, dv, w -> shared memory, red/black SPLIT layout: each colour of
//                                                     the checkerboard is a dense array, so every neighbour
//                                                     access of a warp is unit-stride (no bank conflicts)
//   region C (tile +  9)   : A11, A12, A22, b1, b2 -> REGISTERS of the owning thread (6 column pairs each)
//
// Dependency radius of one iteration is 10 (one pixel per SOR half-sweep), +1 for the weights, +2 for
// the derivatives; halo pixels keep being updated with stale neighbours after their values stop
// mattering, and that contamination travels inwards one pixel per half-sweep, i.e. it never reaches
// the tile.  Arithmetic and operation order are those of vr_math.cuh (bit-exact vs cv2), -fmad=false.
//
// Thread layout: 42 x 14 (588 of the 608 threads; 6 rows each keep the kernel at 96 registers, i.e. 58K of the SM's 64K:
// the 7K left are one 128-thread CTA of the Newton kernel of another context -- with 42 x 12 / 7 rows / 128 registers the VR CTA
// owned the whole register file).  Thread (tx, ty) owns column pair tx (D-local columns 2tx, 2tx+1) of rows
// 1 + ty + 14k, k = 0..5; the row parity -- hence which column of the pair is red -- is fixed per
// thread, and every shared-memory address is `base + k * const`.
#include <cstddef>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "vr_math.cuh"

namespace {

constexpr int HALO = 10;
constexpr int OTW = 64, OTH = 64;                          // output tile
constexpr int DW = OTW + 2 * HALO, DH = OTH + 2 * HALO;    // region D (origin = tile - 10)
constexpr int NP = DW / 2;                                 // column pairs per row == entries per colour per row
constexpr int IX_OFF = 16, IY_OFF = HALO + 1;              // image tile origin = tile - (16, 11): x must be 16-byte aligned for TMA
constexpr int IPITCH = 96, IH = DH + 2;                    // image tile: 96 bytes x 86 rows
constexpr int TY = 14;                                     // thread rows
constexpr int NT = 608;                                    // threads per CTA (NP * TY = 588 active in the pair phases)
constexpr int SLOTS = (DH - 2 + TY - 1) / TY;              // rows owned per thread in the pair phases (6)
constexpr int RSLOTS = DH / TY;                            // rows owned per thread in the staging phases (6)
constexpr int PLANE = DH * NP;                             // floats per colour per plane
static_assert(DW % 2 == 0 && NP * TY <= NT && (TY % 2) == 0 && (HALO % 2) == 0 && (OTW % 2) == 0, "thread layout");
static_assert(IX_OFF + OTW + HALO + 1 <= IPITCH, "image tile width");

struct Smem {
    alignas(128) uint8_t i0[IH * IPITCH];
    alignas(128) uint8_t i1[IH * IPITCH];
    float Ix[2][PLANE], Iy[2][PLANE], Iz[2][PLANE];        // [colour][row * NP + idx]; colour 0 = red = (x + y) even
                                                           // (Ix + Iy are reused as float2 diag[2][PLANE] by the SOR sweeps)
    float2 uv[2][PLANE];                                   // (du, dv) interleaved: one 64-bit access per neighbour
    float ws[2][PLANE];
    alignas(8) unsigned long long bar;
};

static_assert(offsetof(Smem, Iy) == offsetof(Smem, Ix) + sizeof(float) * 2 * PLANE && offsetof(Smem, Ix) % 8 == 0,
              "Ix + Iy must be one contiguous 8-byte aligned block (reused as the float2 diagonal plane)");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// du / dv of region D from the previous iteration (zero on the first one and outside the image), then the first
// derivatives and the smoothness weights on D.  Same 42 x 14 layout as the pair phases: thread (tx, ty) owns
// column pair tx of rows ty + 14k, k = 0..5 (DH == 6 * 14), so the colour of the pair's first column is ty & 1
// for every slot and all 14 global loads of a thread are in flight together.
template <bool INTERIOR, bool FIRST, bool USE_TMA>
__device__ __forceinline__ void stage_phases(Smem &s, const int tid, const int dx0, const int dy0, const int ix0, const int iy0,
                                             const int W, const int H, const float *__restrict__ du_in,
                                             const float *__restrict__ dv_in)
{
    static_assert(DH == RSLOTS * TY, "row slots");
    const int tx = tid % NP, ty = tid / NP;
    const bool tact = tid < NP * TY;
    const int pr = ty & 1;                                  // colour of column 2tx in every owned row
    const int x0 = dx0 + 2 * tx;
    if (tact) {
        float2 u[RSLOTS], v[RSLOTS];
        if (!FIRST) {
            if ((W & 1) == 0) {                             // x0 is even: 8-byte aligned pairs when W is even
                const bool xin = INTERIOR || (x0 >= 0 && x0 < W);
#pragma unroll
                for (int k = 0; k < RSLOTS; k++) {
                    const int y = dy0 + ty + k * TY;
                    u[k] = v[k] = make_float2(0.f, 0.f);
                    if (xin && (INTERIOR || (y >= 0 && y < H))) {
                        const size_t g = (size_t)y * W + x0;
                        u[k] = __ldg(reinterpret_cast<const float2 *>(du_in + g));
                        v[k] = __ldg(reinterpret_cast<const float2 *>(dv_in + g));
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < RSLOTS; k++) {
                    const int y = dy0 + ty + k * TY;
                    u[k] = v[k] = make_float2(0.f, 0.f);
                    if (y >= 0 && y < H) {
                        const size_t g = (size_t)y * W + x0;
                        if (x0 >= 0 && x0 < W) { u[k].x = __ldg(du_in + g); v[k].x = __ldg(dv_in + g); }
                        if (x0 + 1 >= 0 && x0 + 1 < W) { u[k].y = __ldg(du_in + g + 1); v[k].y = __ldg(dv_in + g + 1); }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RSLOTS; k++) {
            const int idx = (ty + k * TY) * NP + tx;
            if (FIRST) u[k] = v[k] = make_float2(0.f, 0.f);
            s.uv[pr][idx] = make_float2(u[k].x, v[k].x);
            s.uv[pr ^ 1][idx] = make_float2(u[k].y, v[k].y);
        }
    }
    if (USE_TMA) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&s.bar)) : "memory");
    }
    __syncthreads();

    if (tact) {
        const int lx = x0 - ix0;                            // I-local column of the pair (even)
#pragma unroll
        for (int k = 0; k < RSLOTS; k++) {
            const int r = ty + k * TY;
            const int y = dy0 + r;
            const int idx = r * NP + tx;
            float ix[2] = {0.f, 0.f}, iy[2] = {0.f, 0.f}, iz[2] = {0.f, 0.f}, w[2] = {0.f, 0.f};
            if (INTERIOR) {
                // rows y-1, y, y+1 of both frames: the pair as one 16-bit load, plus x-1 and x+2 on the centre row
                const int ic = (y - iy0) * IPITCH + lx;
                float a[4], zc[2], au[2], ad[2];
#pragma unroll
                for (int t = 0; t < 4; t++) a[t] = 0.5f * (float)s.i0[ic - 1 + t] + 0.5f * (float)s.i1[ic - 1 + t];
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    zc[t] = (float)s.i1[ic + t] - (float)s.i0[ic + t];
                    au[t] = 0.5f * (float)s.i0[ic - IPITCH + t] + 0.5f * (float)s.i1[ic - IPITCH + t];
                    ad[t] = 0.5f * (float)s.i0[ic + IPITCH + t] + 0.5f * (float)s.i1[ic + IPITCH + t];
                }
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    ix[t] = a[t + 2] - a[t];
                    iy[t] = ad[t] - au[t];
                    iz[t] = zc[t];
                }
                if (r < DH - 1) {
                    const float2 p0 = s.uv[pr][idx], p1 = s.uv[pr ^ 1][idx], d0 = s.uv[pr ^ 1][idx + NP];
                    w[0] = vr_smooth_weight(p1.x - p0.x, p1.y - p0.y, d0.x - p0.x, d0.y - p0.y);
                    if (tx < NP - 1) {
                        const float2 r1 = s.uv[pr][idx + 1], d1 = s.uv[pr][idx + NP];
                        w[1] = vr_smooth_weight(r1.x - p1.x, r1.y - p1.y, d1.x - p1.x, d1.y - p1.y);
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int x = x0 + t, c = 2 * tx + t;
                    const int col = pr ^ t, o = col ^ 1;
                    if (x >= 0 && x < W && y >= 0 && y < H) {
                        const int px = x - ix0, py = y - iy0;
                        const int lxl = vr_clampi(x - 1, W) - ix0, lxr = vr_clampi(x + 1, W) - ix0;
                        const int lyu = vr_clampi(y - 1, H) - iy0, lyd = vr_clampi(y + 1, H) - iy0;
                        auto A = [&](int xx, int yy) { int i = yy * IPITCH + xx; return 0.5f * (float)s.i0[i] + 0.5f * (float)s.i1[i]; };
                        ix[t] = A(lxr, py) - A(lxl, py);
                        iy[t] = A(px, lyd) - A(px, lyu);
                        const int i = py * IPITCH + px;
                        iz[t] = (float)s.i1[i] - (float)s.i0[i];
                        if (r < DH - 1 && c < DW - 1) {
                            const float2 pc = s.uv[col][idx];
                            float ux = 0.f, vx = 0.f, uy = 0.f, vy = 0.f;
                            const int ir = idx + t, id = idx + NP;
                            if (x < W - 1) { const float2 q = s.uv[o][ir]; ux = q.x - pc.x; vx = q.y - pc.y; }
                            if (y < H - 1) { const float2 q = s.uv[o][id]; uy = q.x - pc.x; vy = q.y - pc.y; }
                            w[t] = vr_smooth_weight(ux, vx, uy, vy);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                s.Ix[pr ^ t][idx] = ix[t];
                s.Iy[pr ^ t][idx] = iy[t];
                s.Iz[pr ^ t][idx] = iz[t];
                s.ws[pr ^ t][idx] = w[t];
            }
        }
    }
    __syncthreads();
}

// Data term + red-black SOR for the pixels a thread owns.  INTERIOR tiles (tile + halo entirely inside
// the image) drop every border predicate at compile time.
template <bool INTERIOR>
__device__ __forceinline__ void pair_phases(Smem &s, const int tid, const int dx0, const int dy0, const int W, const int H)
{
    // ---- per-thread constants of the pair phases ------------------------------------------------------------------
    const int tx = tid % NP, ty = tid / NP;
    const bool tact = tid < NP * TY;
    const int par = (1 + ty) & 1;                           // parity of every row this thread owns
    const int base = (1 + ty) * NP + tx;                    // split index of slot 0 (both colours)
    const int ystart = dy0 + 1 + ty;
    int nl[2], xs[2];                                       // per colour: index of the left neighbour (other colour), image x
    bool cact[2], hasL[2], hasR[2];
#pragma unroll
    for (int col = 0; col < 2; col++) {
        int c = 2 * tx + (col == 0 ? par : 1 - par);
        xs[col] = dx0 + c;
        nl[col] = (col == 0) ? tx + par - 1 : tx - par;     // relative to the row start
        nl[col] -= tx;                                      // -> offset from the own index (-1, 0)
        cact[col] = tact && c >= 1 && c <= DW - 2 && (INTERIOR || (xs[col] >= 0 && xs[col] < W));
        hasL[col] = INTERIOR || xs[col] > 0;
        hasR[col] = INTERIOR || xs[col] < W - 1;
    }

    // ---- data term of the owned pixels -> registers ---------------------------------------------------------------
    float A11[SLOTS][2], A12[SLOTS][2], A22[SLOTS][2], B1[SLOTS][2], B2[SLOTS][2];
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        const int idx = base + k * TY * NP;
        const int y = ystart + k * TY;
        const bool ract = (1 + ty + k * TY) <= DH - 2 && (INTERIOR || (y >= 0 && y < H));
        const bool hasU = INTERIOR || y > 0, hasD = INTERIOR || y < H - 1;
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const int o = col ^ 1;
            VrLin l;
            l.A11 = l.A22 = 1.f; l.A12 = l.b1 = l.b2 = 0.f;
            if (ract && cact[col]) {
                const int iL = idx + nl[col], iR = iL + 1, iU = idx - NP, iD = idx + NP;
                VrDeriv d;
                d.Ix = s.Ix[col][idx];
                d.Iy = s.Iy[col][idx];
                d.Iz = s.Iz[col][idx];
                // second derivatives: central differences of the first ones, border-replicated
                float ixl = hasL[col] ? s.Ix[o][iL] : d.Ix, ixr = hasR[col] ? s.Ix[o][iR] : d.Ix;
                float ixu = hasU ? s.Ix[o][iU] : d.Ix, ixd = hasD ? s.Ix[o][iD] : d.Ix;
                float iyu = hasU ? s.Iy[o][iU] : d.Iy, iyd = hasD ? s.Iy[o][iD] : d.Iy;
                float izl = hasL[col] ? s.Iz[o][iL] : d.Iz, izr = hasR[col] ? s.Iz[o][iR] : d.Iz;
                float izu = hasU ? s.Iz[o][iU] : d.Iz, izd = hasD ? s.Iz[o][iD] : d.Iz;
                d.Ixx = ixr - ixl;
                d.Ixy = ixd - ixu;
                d.Iyy = iyd - iyu;
                d.Ixz = izr - izl;
                d.Iyz = izd - izu;
                const float2 pc = s.uv[col][idx];
                l = vr_data_term(d, pc.x, pc.y);
                // link weights -> diagonal (colour-dependent accumulation order)
                float wsP = s.ws[col][idx];
                float sR = hasR[col] ? wsP : 0.f, sD = hasD ? wsP : 0.f;
                float sL = hasL[col] ? s.ws[o][iL] : 0.f, sU = hasU ? s.ws[o][iU] : 0.f;
                l.A11 = vr_add_links(l.A11, sR, sL, sD, sU, col == 0);
                l.A22 = vr_add_links(l.A22, sR, sL, sD, sU, col == 0);
            }
            A11[k][col] = l.A11; A12[k][col] = l.A12; A22[k][col] = l.A22; B1[k][col] = l.b1; B2[k][col] = l.b2;
        }
    }

    // ---- red-black SOR ---------------------------------------------------------------------------------------------
    // The derivative planes are dead from here on: park the diagonals (A11, A22) in their place (own entries only,
    // one 64-bit load per update) and keep their refined reciprocals in the registers instead.
    __syncthreads();
    float2 (*diag)[PLANE] = reinterpret_cast<float2 (*)[PLANE]>(&s.Ix[0][0]);
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
        const int idx = base + k * TY * NP;
        if (tact && (1 + ty + k * TY) <= DH - 2) {
#pragma unroll
            for (int col = 0; col < 2; col++) {
                diag[col][idx] = make_float2(A11[k][col], A22[k][col]);
                A11[k][col] = vr_rcp_refined(A11[k][col]);
                A22[k][col] = vr_rcp_refined(A22[k][col]);
            }
        }
    }

    // Half-sweep hs (0 .. 2 * VR_SOR - 1) has to be right on tile + (2 * VR_SOR - 1 - hs) only: D-row r < HALO is
    // needed while hs < r, row r > HALO + OTH - 1 while hs < DH - 1 - r.  (Skipping is as good as updating with
    // stale neighbours: either way the row's values are never read by a pixel that still matters.)
    static_assert(HALO == 2 * VR_SOR, "halo == dependency radius of the SOR sweeps");
    const int row_first = 1 + ty, row_last = 1 + ty + (SLOTS - 1) * TY;
    const int live_top = row_first < HALO ? row_first : 2 * VR_SOR;
    const int live_bot = row_last > HALO + OTH - 1 ? DH - 1 - row_last : 2 * VR_SOR;
#pragma unroll 1
    for (int sweep = 0; sweep < VR_SOR; sweep++) {
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const int o = col ^ 1;
            const int hs = 2 * sweep + col;
            if (cact[col]) {
#pragma unroll
                for (int k = 0; k < SLOTS; k++) {
                    const int idx = base + k * TY * NP;
                    const int y = ystart + k * TY;
                    bool ract = (1 + ty + k * TY) <= DH - 2 && (INTERIOR || (y >= 0 && y < H));
                    // halo rows stop mattering once the half-sweeps left cannot carry them into the tile
                    if (k == 0) ract = ract && hs < live_top;
                    if (k == SLOTS - 1) ract = ract && hs < live_bot;
                    if (ract) {
                        const int iL = idx + nl[col], iR = iL + 1, iU = idx - NP, iD = idx + NP;
                        float wsP = s.ws[col][idx];
                        float sR = hasR[col] ? wsP : 0.f, sD = (INTERIOR || y < H - 1) ? wsP : 0.f;
                        float sL = hasL[col] ? s.ws[o][iL] : 0.f, sU = (INTERIOR || y > 0) ? s.ws[o][iU] : 0.f;
                        float2 p = s.uv[col][idx];
                        const float2 pL = s.uv[o][iL], pR = s.uv[o][iR], pU = s.uv[o][iU], pD = s.uv[o][iD];
                        const float2 dg = diag[col][idx];
                        vr_sor_update_r(p.x, p.y, sL, sR, sU, sD, pL.x, pR.x, pU.x, pD.x, pL.y, pR.y, pU.y, pD.y, B1[k][col], B2[k][col],
                                        A12[k][col], dg.x, dg.y, A11[k][col], A22[k][col]);
                        s.uv[col][idx] = p;
                    }
                }
            }
            __syncthreads();
        }
    }

}

template <bool FIRST, bool LAST, bool USE_TMA>
__global__ void __launch_bounds__(NT, 1) vr_fused_kernel(const uint8_t *__restrict__ g0, const uint8_t *__restrict__ g1,
                                                         const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                                                         const float *__restrict__ du_in, const float *__restrict__ dv_in,
                                                         float *__restrict__ du_out, float *__restrict__ dv_out,
                                                         float4 *__restrict__ flow4, int W, int H)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;
    const int gx0 = blockIdx.x * OTW, gy0 = blockIdx.y * OTH;
    const int dx0 = gx0 - HALO, dy0 = gy0 - HALO;          // image coords of D-local (0, 0): both even
    const int ix0 = gx0 - IX_OFF, iy0 = gy0 - IY_OFF;      // image coords of I-local (0, 0)

    // ---- stage the two frames ----------------------------------------------------------------------------
    if (USE_TMA) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s.bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s.bar)), "r"(2 * IH * IPITCH) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s.i0)), "l"(&tm0), "r"(ix0), "r"(iy0), "r"(smem_u32(&s.bar)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(s.i1)), "l"(&tm1), "r"(ix0), "r"(iy0), "r"(smem_u32(&s.bar)) : "memory");
        }
    } else {
        for (int e = tid; e < IH * IPITCH; e += NT) {
            int r = e / IPITCH, c = e % IPITCH;
            int x = ix0 + c, y = iy0 + r;
            uint8_t a = 0, b = 0;
            if (x >= 0 && x < W && y >= 0 && y < H) { a = g0[(size_t)y * W + x]; b = g1[(size_t)y * W + x]; }
            s.i0[e] = a;
            s.i1[e] = b;
        }
    }
    // tile + halo + derivative taps inside the image?  (uniform per CTA)
    const bool interior = ix0 >= 0 && iy0 >= 0 && dx0 + DW + 1 <= W && dy0 + DH + 1 <= H;
    if (interior) stage_phases<true, FIRST, USE_TMA>(s, tid, dx0, dy0, ix0, iy0, W, H, du_in, dv_in);
    else stage_phases<false, FIRST, USE_TMA>(s, tid, dx0, dy0, ix0, iy0, W, H, du_in, dv_in);

    if (interior) pair_phases<true>(s, tid, dx0, dy0, W, H);
    else pair_phases<false>(s, tid, dx0, dy0, W, H);

    // ---- write the tile (column pairs; the pair's first column is even, so its colour is the row parity) -----
    const bool vecw = (W & 1) == 0;
    for (int e = tid; e < OTH * (OTW / 2); e += NT) {
        const int tyy = e / (OTW / 2), p = e % (OTW / 2);
        const int x = gx0 + 2 * p, y = gy0 + tyy;
        if (x < W && y < H) {
            const int r = tyy + HALO;
            const int col = r & 1, idx = r * NP + p + HALO / 2;
            const float2 q0 = s.uv[col][idx], q1 = s.uv[col ^ 1][idx];
            const float u0 = q0.x, v0 = q0.y, u1 = q1.x, v1 = q1.y;
            const size_t g = (size_t)y * W + x;
            const bool second = x + 1 < W;
            if (LAST) {
                flow4[g] = make_float4(u0, v0, 0.f, 0.f);
                if (second) flow4[g + 1] = make_float4(u1, v1, 0.f, 0.f);
            } else if (vecw) {
                *reinterpret_cast<float2 *>(du_out + g) = make_float2(u0, u1);
                *reinterpret_cast<float2 *>(dv_out + g) = make_float2(v0, v1);
            } else {
                du_out[g] = u0; dv_out[g] = v0;
                if (second) { du_out[g + 1] = u1; dv_out[g + 1] = v1; }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;          // contexts may be driven from different host threads
    std::call_once(once, []() {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    });
    return fn;
}

// 2-D uint8 tensor map with a (96 x 86) box.  TMA needs: base and row pitch multiples of 16 bytes, and
// (measured on B200: "illegal instruction" otherwise) the box's x origin a multiple of 16 bytes, which
// the kernel guarantees (tile origin - 16).
bool make_u8_map(CUtensorMap *tm, const uint8_t *base, int W, int H)
{
    EncodeTiledFn enc = get_encode();
    if (!enc || (W % 16) != 0 || ((uintptr_t)base % 16) != 0) return false;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {(cuuint32_t)IPITCH, (cuuint32_t)IH};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool FIRST, bool LAST, bool USE_TMA>
int launch(mr_context *ctx, const uint8_t *g0, const uint8_t *g1, const CUtensorMap &tm0, const CUtensorMap &tm1, const float *du_in,
           const float *dv_in, float *du_out, float *dv_out, float4 *flow4)
{
    auto kern = vr_fused_kernel<FIRST, LAST, USE_TMA>;
    MR_CUDA(ctx, mr_ensure_smem(ctx, kern, sizeof(Smem)));
    dim3 grid(cdiv(ctx->W, OTW), cdiv(ctx->H, OTH));
    kern<<<grid, NT, sizeof(Smem), ctx->stream>>>(g0, g1, tm0, tm1, du_in, dv_in, du_out, dv_out, flow4, ctx->W, ctx->H);
    MR_LAUNCH_CHECK(ctx, "vr_fused_kernel");
    return MR_OK;
}

template <bool USE_TMA>
int run(mr_context *ctx, const uint8_t *g0, const uint8_t *g1, const CUtensorMap &tm0, const CUtensorMap &tm1, float *d_flow4)
{
    size_t N = ctx->N;
    float *buf = mr_buf<float>(ctx, "vr_pingpong", N * 4);
    if (!buf) return mr_fail(ctx, MR_ENOMEM, "vr_pingpong", "alloc");
    float *duA = buf, *dvA = buf + N, *duB = buf + 2 * N, *dvB = buf + 3 * N;
    float4 *f4 = (float4 *)d_flow4;
    int rc = launch<true, false, USE_TMA>(ctx, g0, g1, tm0, tm1, nullptr, nullptr, duA, dvA, f4);
    if (rc) return rc;
    for (int it = 1; it < VR_FIXED_POINT - 1; it++) {
        rc = launch<false, false, USE_TMA>(ctx, g0, g1, tm0, tm1, duA, dvA, duB, dvB, f4);
        if (rc) return rc;
        float *t = duA; duA = duB; duB = t;
        t = dvA; dvA = dvB; dvB = t;
    }
    return launch<false, true, USE_TMA>(ctx, g0, g1, tm0, tm1, duA, dvA, nullptr, nullptr, f4);
}

}  // namespace

std::atomic<int> g_mr_vr_tma{1};   // 1 = stage the frames with TMA when the layout allows it, 0 = plain loads

int k_vr_fused(mr_context *ctx, const uint8_t *d_i0, const uint8_t *d_i1, float *d_flow4)
{
    static_assert(VR_FIXED_POINT >= 2, "fixed point iterations");
    CUtensorMap tm0, tm1;
    memset(&tm0, 0, sizeof(tm0));
    memset(&tm1, 0, sizeof(tm1));
    bool tma = g_mr_vr_tma.load() && make_u8_map(&tm0, d_i0, ctx->W, ctx->H) && make_u8_map(&tm1, d_i1, ctx->W, ctx->H);
    if (tma) return run<true>(ctx, d_i0, d_i1, tm0, tm1, d_flow4);
    return run<false>(ctx, d_i0, d_i1, tm0, tm1, d_flow4);
}
