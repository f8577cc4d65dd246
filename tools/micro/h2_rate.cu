// Issue rate of the half2 instructions a packed (two shots per 32-bit slot) min-sum would use, against their f32 twins.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o h2_rate h2_rate.cu && ./h2_rate
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t hmin_xs(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax_xs(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float fmin_xs(float a, float b) { float d; asm volatile("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t hadd(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t heqm(uint32_t a, uint32_t b) { uint32_t d; asm volatile("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE>
__global__ void k(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x3C003C00u;
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(x[i] & 0x3FFFFFFFu);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = hmin_xs(x[i], x[(i + 1) & 7]);
                if (MODE == 1) x[i] = hmax_xs(x[i], x[(i + 1) & 7]);
                if (MODE == 2) f[i] = fmin_xs(f[i], f[(i + 1) & 7]);
                if (MODE == 3) x[i] = hadd(x[i], x[(i + 1) & 7]);
                if (MODE == 4) x[i] = heqm(x[i], x[(i + 1) & 7]);
                if (MODE == 5) f[i] = f[i] + f[(i + 1) & 7];
                if (MODE == 6) x[i] = (x[i] ^ x[(i + 1) & 7]) & 0x7FFF7FFFu;
            }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) acc ^= x[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char *name)
{
    uint32_t *out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4096;
    k<MODE><<<148, 1024>>>(out, 64, 12345u);
    cudaEventRecord(a); k<MODE><<<148, 1024>>>(out, iters, 12345u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_instr = (double)iters * 32 * 32;      // per SM: 32 warps x 32 ops per iteration
    printf("%-28s %.3f ms  %.2f cycles per warp instruction per SM (1965 MHz)\n", name, ms, ms * 1e-3 * 1.965e9 / warp_instr);
    cudaFree(out);
}

int main()
{
    run<0>("min.xorsign.abs.f16x2"); run<1>("max.xorsign.abs.f16x2"); run<2>("min.xorsign.abs.f32");
    run<3>("add.rn.f16x2"); run<4>("set.eq.u32.f16x2"); run<5>("add.f32"); run<6>("lop3 (xor+and)");
    return 0;
}
