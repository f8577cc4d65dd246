// Microbenchmark: shared-memory instruction throughput per SM for the access shapes of minsum_edge_kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float lds(uint32_t a) { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v)); }
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float *out, long long *cyc, int iters)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32768; i += 1024) sm[i] = (float)i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    // scattered conflict-free: lane l reads word (row r)*32 + ((l*5 + r) & 31) for pseudo-random rows r
    uint32_t a[8];
    for (int j = 0; j < 8; ++j) {
        const uint32_t row = (uint32_t)((warp * 131 + j * 977 + lane * 37) % 1000);
        a[j] = base + (row * 32 + ((lane * 5 + j + warp) & 31)) * 4;
        if (MODE == 0) a[j] = base + ((warp * 8 + j) * 32 + lane) * 4;               // consecutive
    }
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += lds(a[j]);
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sts(a[j], acc);
        } else if (MODE == 3) {   // gather + scatter mix like phase B
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += lds(a[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) sts(a[j], acc);
        } else if (MODE == 4) {   // LDS.128 consecutive
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + ((warp * 8 + j) * 32 + lane) * 16 % 131072));
                acc += v.x + v.w;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 1024 + tid] = acc;
}
int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    const char *names[5] = {"LDS.32 consecutive", "LDS.32 scattered conflict-free", "STS.32 scattered", "4 LDS + 4 STS scattered", "LDS.128 consecutive"};
    for (int mode = 0; mode < 5; ++mode) {
        void (*fn)(float *, long long *, int) = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : mode == 3 ? k<3> : k<4>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fn<<<148, 1024, 131072>>>(out, cyc, iters);
        cudaEventRecord(e0);
        fn<<<148, 1024, 131072>>>(out, cyc, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double c = ms * 1e-3 * 1.965e9;
        const double per = (double)c / ((double)iters * 8 * 32);   // cycles per warp-level instruction per SM (at 1965 MHz)
        printf("%-34s %.2f cycles per warp instruction per SM (%s)\n", names[mode], per, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
