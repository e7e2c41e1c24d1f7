// pipes.cu -- B200 pipe micro-benchmarks behind the design of the normals / Newton kernels:
// throughput (independent chains) of F2F.F64.F32, F2F.F32.F64, integer widening, DFMA, DADD, FFMA, I2F.F64, MUFU.RCP,
// per SM and cycle.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

__device__ __forceinline__ double widen_int(float f)
{
    // exact float -> double for normal floats, integer ALU only
    const unsigned b = __float_as_uint(f);
    const unsigned t = b >> 3;
    const unsigned hi = (t & 0x0fffffffu) + 0x38000000u + 7u * (t & 0x10000000u);
    const unsigned lo = b << 29;
    return __hiloint2double((int)hi, (int)lo);
}

template <int OP>
__global__ void k(float *out, float seed, long long *cycles)
{
    float f[CHAINS];
    double d[CHAINS];
    int n[CHAINS];
    for (int i = 0; i < CHAINS; i++) { f[i] = seed + i * 0.37f + threadIdx.x * 1e-3f; d[i] = f[i]; n[i] = (int)(f[i] * 1000.f); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) { d[i] = d[i] + (double)f[i]; f[i] = f[i] + 1.0f; }                 // F2F.F64.F32 + DADD + FADD
            else if (OP == 1) { d[i] = __fma_rn(d[i], 1.0000001, 0.5); }                      // DFMA
            else if (OP == 2) { d[i] = d[i] + 0.5; }                                           // DADD
            else if (OP == 3) { f[i] = __fmaf_rn(f[i], 1.0000001f, 0.5f); }                   // FFMA
            else if (OP == 4) { d[i] = d[i] + widen_int(f[i]); f[i] = f[i] + 1.0f; }         // int widening + DADD + FADD
            else if (OP == 5) { f[i] = f[i] + (float)d[i]; d[i] = d[i] + 1.0; }               // F2F.F32.F64 + FADD + DADD
            else if (OP == 6) { d[i] = d[i] + (double)n[i]; n[i] += 3; }                      // I2F.F64.S32 + DADD + IADD
            else if (OP == 7) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[i])); f[i] = r + 1.0f; }  // MUFU.RCP + FADD
            else if (OP == 8) { f[i] = f[i] + 1.0f; }                                          // FADD only (baseline for 0/4)
            else if (OP == 9) { n[i] = __float2int_rn(f[i]) + n[i]; f[i] = f[i] + 1.0f; }     // F2I + IADD + FADD
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += f[i] + (float)d[i] + (float)n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_iter)
{
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<OP><<<148, warps * 32>>>(out, 1.5f, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_clk = (double)warps * 32 * CHAINS * ITERS / (double)h;   // chain-steps per clock per SM
        printf("%-44s warps/SM %2d  %8.2f steps/clk/SM  (%d ops per step)\n", name, warps, per_clk, ops_per_iter);
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<8>("FADD", 1);
    run<3>("FFMA", 1);
    run<2>("DADD", 1);
    run<1>("DFMA", 1);
    run<0>("F2F.F64.F32 + DADD + FADD", 3);
    run<4>("int-widen(5 ops) + DADD + FADD", 7);
    run<5>("F2F.F32.F64 + FADD + DADD", 3);
    run<6>("I2F.F64.S32 + DADD + IADD", 3);
    run<9>("F2I.S32.F32 + IADD + FADD", 3);
    run<7>("MUFU.RCP + FADD", 2);
    return 0;
}
