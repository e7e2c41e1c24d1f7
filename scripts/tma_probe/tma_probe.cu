// minimal TMA 2D u8 tile load probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define BOXW 96
#define BOXH 86
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VARIANT>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int y0, uint8_t *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *bar = (unsigned long long *)(smem + BOXW * BOXH);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        if (VARIANT == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BOXW * BOXH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem)), "l"(&tm), "r"(x0), "r"(y0), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = smem[i];
}
int main()
{
    int W = 160, H = 128;
    std::vector<uint8_t> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (uint8_t)(i * 7 + i / W);
    uint8_t *d, *o;
    cudaMalloc(&d, W * H); cudaMalloc(&o, BOXW * BOXH);
    cudaMemcpy(d, h.data(), W * H, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {BOXW, BOXH}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d\n", (int)r, (int)q);
    for (int variant = 0; variant < 2; variant++) {
        const int xs[5] = {16, 16, -16, 5, -11}, ys[5] = {8, -11, -11, 8, 3};
        for (int t = 0; t < 5; t++) {
            int x0 = xs[t], y0 = ys[t];
            cudaMemset(o, 0xEE, BOXW * BOXH);
            if (variant == 0) probe<0><<<1, 128, BOXW * BOXH + 16>>>(tm, x0, y0, o);
            else probe<1><<<1, 128, BOXW * BOXH + 16>>>(tm, x0, y0, o);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint8_t> g(BOXW * BOXH);
            cudaMemcpy(g.data(), o, BOXW * BOXH, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int yy = 0; yy < BOXH; yy++) for (int xx = 0; xx < BOXW; xx++) {
                int gx = x0 + xx, gy = y0 + yy;
                uint8_t exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0;
                bad += g[yy * BOXW + xx] != exp;
            }
            printf("variant %d origin (%d,%d): %s, mismatches %d\n", variant, x0, y0, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) { printf("stop\n"); return 1; }
        }
    }
    return 0;
}
