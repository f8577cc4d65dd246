// Device-side view of the per-edge min-sum plan and helpers shared by the float32 kernel (minsum_edge.cu) and the
// packed two-shots-per-slot kernel (minsum_edge_h2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "common.cuh"
#include "edge_layout.h"

namespace qb {

struct EdgeDev {
    int n_rsl, n_csl, e_words, e_dummy, idx_words, nw, n, mw;
    const float4 *E0;            // [e_words/4] prior per slot, +inf in unused slots
    const uint32_t *col_idx;     // [idx_words]
    const uint32_t *col_rowpos;  // [idx_words]
    const uint2 *rtask;          // [n_rsl]
    const uint2 *ctask;          // [n_csl]
    const uint16_t *row_id;      // [n_rsl*32]
    const uint2 *row_pads;       // [n_rsl*32] 4 x u16
    const uint16_t *var_id;      // [n_csl*32]
    const float *lane_prior;     // [n_csl*32] or nullptr (uniform priors in ctask)
    const int32_t *wr_ptr, *wc_ptr;
    const uint4 *wc_cls;         // [nwarps] 16 x u8 slices per class
    const uint32_t *row_mask;    // [n_rsl*32]
    const uint32_t *col_sig;     // [n_csl*32]
    // byte offsets of the float32 kernel's shared-memory arrays behind E (minsum_edge.cu; precomputed so that the kernel
    // does not re-derive a chain of eight array sizes every time it needs one of the addresses)
    uint32_t o_idx, o_rtask, o_syn, o_par, o_hperm, o_hnat, o_cmeta, o_csig;
};

// ---- phase B ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v)); }
__device__ __forceinline__ uint32_t lds_u32v(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v)); }
// descriptors and slot indices are written once before the first barrier: plain (movable) loads
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) { uint2 v; asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }

// per-slice priors travel in the kernel parameters (constant bank): a warp-uniform index makes them uniform
// loads, off the shared-memory pipe that bounds phase B
constexpr int EDGE_MAX_CSL = 768;
struct EdgePriors { uint32_t bits[EDGE_MAX_CSL]; };

// weight of the residual syndrome par ^ syn (0 = converged); resets par.  Called by warp 0 only.
__device__ __forceinline__ int residual_weight(uint32_t *par, const uint32_t *syn, int n_rsl, int lane)
{
    int wt = 0;
    for (int w = lane; w < n_rsl; w += 32) { wt += __popc(par[w] ^ syn[w]); par[w] = 0u; }
    return __reduce_add_sync(0xFFFFFFFFu, wt);
}

// exact H.hard for the hard decision in hperm: par ^= rows of every variable whose bit is set (all threads; par must
// be 0; contains barriers).  The set bits are first compacted into a list so that every (variable, edge) pair gets
// its own thread: the row positions come from global memory and the loads of one thread would otherwise serialise.
constexpr int PAR_LIST_CAP = 512;
__device__ __forceinline__ void parity_of_hard(const EdgeDev &eg, const uint32_t *hperm, const uint32_t *cmeta, uint32_t *par,
                                               uint16_t *list, int *list_count, int tid, int nthreads)
{
    if (tid == 0) *list_count = 0;
    __syncthreads();
    for (int t = tid; t < eg.n_csl; t += nthreads) {
        uint32_t bits = hperm[t];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int slot = atomicAdd(list_count, 1);
            if (slot < PAR_LIST_CAP) list[slot] = (uint16_t)(t * 32 + b);
        }
    }
    __syncthreads();
    const int cnt = *list_count;
    if (cnt <= PAR_LIST_CAP) {
        for (int i = tid; i < cnt * 8; i += nthreads) {              // degree <= 16: two passes of 8 edges
            const int e = list[i >> 3], t = e >> 5, b = e & 31;
            const uint32_t dx = cmeta[t];
            const int D = (dx >> 16) & 63, H = (D + 1) >> 1;
            const uint32_t *rp = eg.col_rowpos + (dx & 0xFFFFu) * 32;
            for (int k = i & 7; k < D; k += 8) {
                const uint32_t w = __ldg(&rp[edge_idx_off(H, k >> 1, b)]);
                const uint32_t pos = (k & 1) ? (w >> 16) : (w & 0xFFFFu);
                atomicXor(&par[pos >> 5], 1u << (pos & 31));
            }
        }
    } else {
        for (int t = tid; t < eg.n_csl; t += nthreads) {
            uint32_t bits = hperm[t];
            if (!bits) continue;
            const uint32_t dx = cmeta[t];
            const int D = (dx >> 16) & 63, H = (D + 1) >> 1;
            const uint32_t *rp = eg.col_rowpos + (dx & 0xFFFFu) * 32;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                for (int k = 0; k < D; ++k) {
                    const uint32_t w = __ldg(&rp[edge_idx_off(H, k >> 1, b)]);
                    const uint32_t pos = (k & 1) ? (w >> 16) : (w & 0xFFFFu);
                    atomicXor(&par[pos >> 5], 1u << (pos & 31));
                }
            }
        }
    }
}

struct EdgePlanH2;   // packed two-shots-per-slot mode (minsum_edge_h2.cu), built on first use
int edge_plan_h2_create(const struct EdgePlan *ep, int nw, const float *prior_h, EdgePlanH2 **out);
void edge_plan_h2_destroy(EdgePlanH2 *p);
bool edge_h2_fits(const qb_decoder *dec, const struct EdgePlan *ep, const EdgePlanH2 *hp);
int launch_minsum_edge_h2(qb_decoder *dec, struct EdgePlan *ep, EdgePlanH2 *hp, const MinsumLaunch &a, cudaStream_t st);

struct EdgePlan {
    EdgeLayout L;
    EdgeDev dev{};
    std::vector<void *> owned;
    float *d_E0 = nullptr;
    float *d_lane_prior = nullptr;
    int *d_counter = nullptr;
    int threads = 0, ctas_per_sm = 0;
    size_t smem = 0;
    EdgePriors pri{};
};


}  // namespace qb
