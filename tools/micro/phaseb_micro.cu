// Microbenchmark of phase B of minsum_edge_kernel on synthetic conflict-free tables (same device code).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../qldpc-branched-off_b200/csrc/minsum_edge.cu"
using namespace qb;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) kb(float *out, int iters, int tasks_per_warp, int D, int a_warps, int a_tasks, const __grid_constant__ EdgePriors pri)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int e_words = (MODE >= 2) ? 34000 : 24000;
    float *E = reinterpret_cast<float *>(smem_raw);
    uint32_t *idx = reinterpret_cast<uint32_t *>(E + e_words);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_min_sync(0xFFFFFFFFu, tid >> 5);
    const int nwarps = blockDim.x >> 5;
    const int H = (D + 1) / 2;
    const int n_csl = (nwarps - a_warps) * tasks_per_warp;
    uint32_t *cmeta = idx + n_csl * H * 32;
    uint8_t *csig = reinterpret_cast<uint8_t *>(cmeta + n_csl);
    const uint32_t e_word = (uint32_t)__cvta_generic_to_shared(E) >> 2;
    for (int i = tid; i < e_words; i += blockDim.x) E[i] = 0.001f * (i & 1023);
    for (int t = 0; t < n_csl; ++t)
        for (int i = tid; i < H * 32; i += blockDim.x) {
            const int kk = i >> 5, l = i & 31;
            uint32_t s[2];
            for (int h = 0; h < 2; ++h) {
                const int k = 2 * kk + h;
                const uint32_t row = (uint32_t)((t * 131 + k * 977 + l * 37 + 11) % 740);
                s[h] = row * 32 + ((l * 5 + k + t) & 31) + e_word;
            }
            idx[t * H * 32 + edge_idx_off(H, kk, l)] = s[0] | (s[1] << 16);
        }
    for (int i = tid; i < n_csl; i += blockDim.x) cmeta[i] = 0;
    for (int i = tid; i < n_csl * 32; i += blockDim.x) csig[i] = (uint8_t)i;
    __syncthreads();
    const uint32_t idx_addr = (uint32_t)__cvta_generic_to_shared(idx);
    const int c0 = warp * tasks_per_warp;
    uint4 cls = make_uint4(0, 0, 0, 0);
    if (D == 2) cls.x = tasks_per_warp << 16;
    if (D == 3) cls.x = tasks_per_warp << 24;
    if (D == 4) cls.y = tasks_per_warp;
    if (D == 5) cls.y = tasks_per_warp << 8;
    if (D == 6) cls.y = tasks_per_warp << 16;
    uint32_t fpacc = 0;
    if (MODE >= 2) {
        // overlap experiment: warps [0, a_warps) run a_tasks row tasks (K = 9) per iteration, the others run phase B
        const int bw = warp - a_warps;
        const int cb = bw * tasks_per_warp;
        for (int it = 0; it < iters; ++it) {
            if (warp < a_warps) {
                if (MODE != 4)
                    for (int t = 0; t < a_tasks; ++t) {
                        const int base_unit = ((warp * a_tasks + t) * 297) % 8000;
                        row_task<9, false>(E, nullptr, base_unit, 33, lane, 0u, 0.75f, 20.f, make_uint2(0xFFFF0000u + (uint32_t)(base_unit * 4 + lane), 0xFFFFFFFFu));
                    }
            } else if (MODE != 3) {
                ColCtx c;
                c.ix = idx_addr + cb * H * 128; c.lane4 = lane * 4; c.lane8 = lane * 8;
                c.sg = (uint32_t)__cvta_generic_to_shared(csig + cb * 32 + lane);
                c.fp = 0u; c.myhw = 0u; c.t4 = 4u * cb; c.lane_t4 = 4u * (cb + lane); c.lane = lane; c.vid = nullptr; c.post = nullptr;
                phase_b<false>(c, cls, cb + tasks_per_warp, cmeta, nullptr, pri);
                fpacc ^= c.fp ^ c.myhw;
            }
            __syncthreads();
        }
        out[blockIdx.x * blockDim.x + tid] = E[tid] + fpacc;
        return;
    }
    for (int it = 0; it < iters; ++it) {
        ColCtx c;
        c.ix = idx_addr + c0 * H * 128; c.lane4 = lane * 4; c.lane8 = lane * 8;
        c.sg = (uint32_t)__cvta_generic_to_shared(csig + c0 * 32 + lane);
        c.fp = 0u; c.myhw = 0u; c.t4 = 4u * c0; c.lane_t4 = 4u * (c0 + lane); c.lane = lane; c.vid = nullptr; c.post = nullptr;
        phase_b<false>(c, cls, c0 + tasks_per_warp, cmeta, nullptr, pri);
        fpacc ^= c.fp ^ c.myhw;
        if (MODE == 0) __syncthreads();
    }
    out[blockIdx.x * blockDim.x + tid] = E[tid] + fpacc;
}

int main()
{
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    EdgePriors pri; for (int i = 0; i < EDGE_MAX_CSL; ++i) pri.bits[i] = 0x40400000u;   // 3.0f
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {1024, 512})
            for (int D : {3, 6}) {
                const int tpw = (threads == 1024 ? 9 : 18);
                auto fn = mode == 0 ? kb<0> : kb<1>;
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                fn<<<148, threads, 220000>>>(out, 10, tpw, D, 0, 0, pri);
                cudaEventRecord(e0);
                fn<<<148, threads, 220000>>>(out, iters, tpw, D, 0, 0, pri);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double cyc = ms * 1e-3 * 1.965e9 / iters;
                const int tasks = (threads / 32) * tpw;
                printf("%s threads=%d D=%d tasks/SM=%d: %.0f cycles per phase B, %.1f cycles/task/SM, LSU instr/clk %.2f (%s)\n", mode == 0 ? "barrier" : "no-barrier", threads, D, tasks, cyc,
                       cyc / tasks, tasks * (2.0 * D + (D + 1) / 2 + 2) / cyc, cudaGetErrorString(cudaGetLastError()));
            }
    // overlap: 8 A-warps x 5 row tasks (40 K=9 row slices), 24 B-warps x 12 D=3 tasks (288 column slices)
    for (int mode = 2; mode <= 4; ++mode) {
        auto fn = mode == 2 ? kb<2> : mode == 3 ? kb<3> : kb<4>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fn<<<148, 1024, 225000>>>(out, 10, 12, 3, 8, 5, pri);
        cudaEventRecord(e0);
        fn<<<148, 1024, 225000>>>(out, iters, 12, 3, 8, 5, pri);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.0f cycles per iteration (%s)\n", mode == 2 ? "A (8 warps x 5 row tasks) || B (24 warps x 12 tasks)" : mode == 3 ? "A only (8 warps x 5 row tasks)" : "B only (24 warps x 12 tasks)",
               ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    // check-row phase alone in the real kernel's shape: 25 warps x 1 row task (K = 9), 7 idle warps
    for (int aw : {25, 28, 32}) {
        auto fn = kb<3>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fn<<<148, 1024, 225000>>>(out, 10, 1, 3, aw, 1, pri);
        cudaEventRecord(e0);
        fn<<<148, 1024, 225000>>>(out, iters, 1, 3, aw, 1, pri);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("A only: %d warps x 1 row task (K=9): %.0f cycles per iteration (%s)\n", aw, ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
namespace qb { int cuda_fail(cudaError_t, const char *, const char *, int) { return -2; } void set_error(const std::string &) {} }
