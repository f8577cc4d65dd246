// K3 (fast path): flooding min-sum with one float per Tanner-graph edge resident in shared memory.
//
// Reference semantics: src/decoding/kernels.py:235-366 (minsum_decoder_full) with damping == 1, which is
// every call the engine makes (engine.py:84-88).  One persistent CTA per SM decodes one shot at a time:
//
//   E[slot]  one float per edge, in the conflict-free row-major layout built by edge_layout.cu.
//            Between phase B and phase A it holds the variable-to-check message Q (kernels.py:323-345),
//            between phase A and phase B the check-to-variable message R (kernels.py:283-316).
//   phase A  one lane per check row: stream the row's Q with LDS.128, keep them in registers, reduce
//            (min1 with the sign product in ONE instruction: min.xorsign.abs; min2 with FMNMX + a 3-input FMNMX3
//            per two edges), then R = sign * alpha * (|Q| == min1 ? min2 : min1) written back in place (STS.128).
//            The value test replaces the reference's argmin position: when two edges tie for min1, min2 == min1.
//            The clip of kernels.py:330-333 is applied here, once per row, to min1 / min2 (E holds unclipped Q).
//   phase B  one lane per variable: gather its <= 16 R through precomputed 16-bit slot indices (bank
//            conflict free by construction), sum in row order + prior (kernels.py:316-320), hard decision,
//            Q = v - R (NaN -> 0 next to degree-1 rows, kernels.py:327-329) scattered back to the same slots.
//            Convergence (H.hard == syndrome, kernels.py:352-364) is first tested on an 8-bit linear fingerprint
//            (XOR of per-column random signatures over the variables whose hard decision is 1 against the
//            XOR of per-row masks over the syndrome) and confirmed exactly only when the fingerprints agree,
//            so the common non-converged iteration never walks the graph a second time.
// Iteration 0 reads Q = prior (unclipped, kernels.py:263-265) straight from a global image of E with LDG.128 (a TMA bulk
// copy of the image into E between two shots, cp.async.bulk + mbarrier, was measured 3 % slower: E is only free for
// ~3 k cycles and the per-SM bulk-copy rate does not cover 130 KB in that window, while the LDG path overlaps the
// copy with the check-row arithmetic of iteration 0).
// No fast-math: IEEE inf/NaN conventions are the reference's (degree-1 rows send +-inf).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "edge_dev.cuh"
#include "edge_layout.h"
#include "edge_phases.cuh"

namespace qb {

#ifdef QB_EDGE_PROFILE
__device__ unsigned long long g_edge_prof[256 * 32 * 4];   // [cta][warp]{phase A work, wait 1, phase B work, wait 2} cycles
#define PROF_T(x) long long x; asm volatile("mov.u64 %0, %%clock64;" : "=l"(x) :: "memory")
// BAR.SYNC does not block at issue: a read of barrier-protected shared memory and a branch on its value do
#define PROF_BLOCK() do { unsigned d_; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d_) : "r"((unsigned)__cvta_generic_to_shared(&s_wt)) : "memory"); \
                          if (d_ == 0x7FFFFFF1u) prof[0] += 1; } while (0)
#define PROF_ADD(i, d) prof[i] += (unsigned long long)(d)
#else
#define PROF_T(x)
#define PROF_ADD(i, d)
#define PROF_BLOCK()
#endif

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
minsum_edge_kernel(const __grid_constant__ EdgeDev eg, const __grid_constant__ MinsumLaunch a, int *shot_counter,
                   const __grid_constant__ EdgePriors pri)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // (array offsets: edge_smem_offsets() on the host)
    float *E = reinterpret_cast<float *>(smem_raw);                                   // [e_words]
    uint32_t *idx = reinterpret_cast<uint32_t *>(smem_raw + eg.o_idx);                // [idx_words]
    uint2 *rtask = reinterpret_cast<uint2 *>(smem_raw + eg.o_rtask);                  // [n_rsl]
    uint32_t *syn = reinterpret_cast<uint32_t *>(smem_raw + eg.o_syn);                // [n_rsl] permuted syndrome bits
    uint32_t *par = reinterpret_cast<uint32_t *>(smem_raw + eg.o_par);                // [n_rsl] parity of the hard decision
    uint32_t *hperm = reinterpret_cast<uint32_t *>(smem_raw + eg.o_hperm);            // [n_csl] hard decision word per column slice
    uint32_t *hnat = reinterpret_cast<uint32_t *>(smem_raw + eg.o_hnat);              // [nw] hard decision, natural order
    uint32_t *cmeta = reinterpret_cast<uint32_t *>(smem_raw + eg.o_cmeta);            // [n_csl] degree / lanes of every column slice
    uint8_t *csig = smem_raw + eg.o_csig;                                             // [n_csl*32] 8-bit fingerprint per variable
    __shared__ int s_wt, s_next, s_pcount;
    __shared__ uint16_t s_plist[PAR_LIST_CAP];
    __shared__ float s_alpha[128];                                                    // alpha schedule (first 128 iterations)
    __shared__ uint32_t s_fp[2], s_target;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_min_sync(0xFFFFFFFFu, tid >> 5);                       // provably warp-uniform
    const uint32_t idx_addr = (uint32_t)__cvta_generic_to_shared(idx);
    // slot indices become absolute shared word addresses (E base folded in), so that a gather address is one
    // shift + one mask away from the packed pair
    const uint32_t e_word = (uint32_t)__cvta_generic_to_shared(E) >> 2;
    // (an upper half EDGE_SIG_TAG | f is the fingerprint f of an odd-degree variable, not a slot,
    // and its lower half becomes a byte address: byte address | fingerprint << 24)
    for (int i = tid; i < eg.idx_words; i += THREADS) {
        const uint32_t w = eg.col_idx[i];
        idx[i] = (w >> 24) == (EDGE_SIG_TAG >> 8) ? ((((w & 0xFFFFu) + e_word) << 2) | (w >> 16 << 24)) : w + (e_word | (e_word << 16));
    }
    for (int i = tid; i < eg.n_csl; i += THREADS) {
        const uint2 d = eg.ctask[i];
        cmeta[i] = d.x;
        hperm[i] = 0u;
    }
    for (int i = tid; i < eg.n_rsl; i += THREADS) rtask[i] = eg.rtask[i];
    for (int i = tid; i < 128 && i < a.max_iter; i += THREADS) s_alpha[i] = a.alpha_d[i];
    for (int i = tid; i < eg.n_csl * 32; i += THREADS) csig[i] = (uint8_t)eg.col_sig[i];
    if (tid < 32) E[eg.e_dummy + tid] = 0.f;                                         // dummy lanes of partial column slices
    const int r0 = eg.wr_ptr[warp], r1 = eg.wr_ptr[warp + 1];
    const int c0 = eg.wc_ptr[warp];
    const uint4 cls = eg.wc_cls[warp];

    const int c1 = eg.wc_ptr[warp + 1];
    // shared address of the index block of the warp's first column slice
    const uint32_t ix0 = idx_addr + (c0 < eg.n_csl ? (eg.ctask[c0].x & 0xFFFFu) * 128u : 0u);
    if (tid == 0) { s_next = atomicAdd(shot_counter, 1); s_fp[0] = 0u; s_fp[1] = 0u; s_target = 0u; }
    __syncthreads();
    int shot = s_next;
    const bool api = !a.post_failed_only;
#ifdef QB_EDGE_PROFILE
    unsigned long long prof[4] = {0, 0, 0, 0};
#endif

    // row id and fingerprint mask of the lane's row in the warp's first two row slices (static: kept in registers so
    // that the per-shot syndrome load is one global round trip)
    uint32_t rid_c[2], msk_c[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int t = warp + u * (THREADS / 32);
        rid_c[u] = t < eg.n_rsl ? eg.row_id[t * 32 + lane] : 0xFFFFu;
        msk_c[u] = t < eg.n_rsl ? (eg.row_mask[t * 32 + lane] & 0xFFu) : 0u;
    }

    while (shot < a.B) {                                                              // uniform
        // ---- load: permuted syndrome words and their fingerprint, parity = 0 ----------------------------------
        {
            uint32_t tg = 0u;
            int u = 0;
            for (int t = warp; t < eg.n_rsl; t += THREADS / 32, ++u) {
                const uint32_t rid = u == 0 ? rid_c[0] : (u == 1 ? rid_c[1] : (uint32_t)eg.row_id[t * 32 + lane]);
                const bool bit = rid != 0xFFFFu && ((a.syn_bits[(size_t)shot * eg.mw + (rid >> 5)] >> (rid & 31)) & 1u);
                if (bit) tg ^= u == 0 ? msk_c[0] : (u == 1 ? msk_c[1] : (eg.row_mask[t * 32 + lane] & 0xFFu));
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
                if (lane == 0) { syn[t] = word; par[t] = 0u; }
            }
            tg = __reduce_xor_sync(0xFFFFFFFFu, tg);
            if (lane == 0 && tg) atomicXor(&s_target, tg);
        }
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(shot_counter, 1);                            // prefetch the next shot id
        const uint32_t target = s_target;

        bool conv = false;
        int fin = a.max_iter - 1;
        // hard decision of the last variable phase: bit j of lane l = variable l of the warp's j-th column slice.  The
        // words of hperm are formed (one ballot per slice) only when they are read: on a fingerprint match and at the end
        uint32_t negbits = 0u;
        auto publish_hperm = [&]() {
            uint32_t mine = 0u;
            for (int s = 0; s < c1 - c0; ++s) {
                const uint32_t hw = __ballot_sync(0xFFFFFFFFu, (negbits >> s) & 1u);
                if (lane == s) mine = hw;
            }
            if (lane < c1 - c0) hperm[c0 + lane] = mine;
        };
        for (int it = 0; it < a.max_iter; ++it) {
            // ---- phase A --------------------------------------------------------------------------------
            const float alpha = it < 128 ? s_alpha[it] : a.alpha_d[it];
            PROF_T(t0);
#ifndef QB_EDGE_SKIP_A
            for (int t = r0; t < r1; ++t) {
                const uint2 d = rtask[t];
                const int K = d.y & 255, nl = (d.y >> 8) & 255, stride = d.y >> 16;
                if (K == 0 || lane >= nl) continue;
                const uint32_t synsign = ((syn[t] >> lane) & 1u) << 31;
                if (d.x & 1u) {                                                        // exactly one unused slot per row (K <= 9)
                    const uint32_t pad = __ldg(reinterpret_cast<const uint32_t *>(&eg.row_pads[t * 32 + lane]));
                    if (it == 0) row_dispatch_onepad<true>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, INFINITY, pad);
                    else row_dispatch_onepad<false>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, a.clip, pad);
                    continue;
                }
                const uint2 pads = __ldg(&eg.row_pads[t * 32 + lane]);
                if (it == 0) row_dispatch<true>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, INFINITY, pads);
                else row_dispatch<false>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, a.clip, pads);
            }
#endif
            if (tid == 0) s_fp[it & 1] = 0u;                                           // fingerprint accumulator of this iteration
            PROF_T(t1);
            __syncthreads();
            PROF_BLOCK();
            PROF_T(t2);
            // ---- phase B --------------------------------------------------------------------------------
            const bool write_v = a.post && (api || it == a.max_iter - 1);
            ColCtx c;
            c.ix = ix0;
            c.lane4 = lane * 4; c.lane8 = lane * 8;
            c.sg = (uint32_t)__cvta_generic_to_shared(csig + lane);
            c.fp = 0u; c.fpw = 0u; c.myhw = 0u; c.negbits = 0u; c.ubit = 1u;
            c.t4 = 4u * (uint32_t)c0; c.lane_t4 = 4u * (uint32_t)(c0 + lane); c.lane = lane;
            c.win = 0u;
#ifndef QB_EDGE_SKIP_B
            if (write_v) {                                                             // (posterior output: last iteration only in the pipeline)
                c.vid = eg.var_id + c0 * 32 + lane;
                c.vid_next = __ldg(c.vid);
                c.post = a.post + (size_t)shot * eg.n;
                phase_b<true, true>(c, cls, c1, cmeta, eg.lane_prior, pri);
            } else {
                c.vid = nullptr; c.vid_next = 0u; c.post = nullptr;
                phase_b<false, true>(c, cls, c1, cmeta, eg.lane_prior, pri);
            }
#endif
            negbits = c.negbits;
            const uint32_t fp = __reduce_xor_sync(0xFFFFFFFFu, c.fp ^ (c.fpw >> 24)) & 0xFFu;
            if (lane == 0 && fp) atomicXor(&s_fp[it & 1], fp);
            PROF_T(t3);
            __syncthreads();
            PROF_BLOCK();
            PROF_T(t4);
            PROF_ADD(0, t1 - t0); PROF_ADD(1, t2 - t1); PROF_ADD(2, t3 - t2); PROF_ADD(3, t4 - t3);
            // ---- convergence: fingerprint of H.hard against the syndrome's, exact test only on a match -----
            if (s_fp[it & 1] == target) {                                              // uniform
                publish_hperm();
                parity_of_hard(eg, hperm, cmeta, par, s_plist, &s_pcount, tid, THREADS);
                __syncthreads();
                if (warp == 0) { const int w = residual_weight(par, syn, eg.n_rsl, lane); if (lane == 0) s_wt = w; }
                __syncthreads();
                if (s_wt == 0) { conv = true; fin = it; break; }                       // kernels.py:352-364
            }
        }
        // ---- end of the shot: hard decision back to natural column order and, for a non-converged shot, the weight
        // of its residual syndrome (OSD scheduling hint).  Both walk the variables whose hard decision is 1: their
        // (slice, lane) pairs are listed once, then every (variable, edge) pair and every variable id gets a thread.
        const bool need_wt = !conv && a.max_iter > 0 && a.fail_wt != nullptr;
        publish_hperm();
        for (int w = tid; w < eg.nw; w += THREADS) hnat[w] = 0u;
        if (tid == 0) { s_target = 0u; s_pcount = 0; }
        __syncthreads();
        for (int t = tid; t < eg.n_csl; t += THREADS) {
            uint32_t bits = hperm[t];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                const int slot = atomicAdd(&s_pcount, 1);
                if (slot < PAR_LIST_CAP) s_plist[slot] = (uint16_t)(t * 32 + b);
            }
        }
        __syncthreads();
        const int n_set = s_pcount;
        if (n_set <= PAR_LIST_CAP) {
            for (int i = tid; i < n_set * 8; i += THREADS) {
                const int e = s_plist[i >> 3], t = e >> 5, b = e & 31, k0 = i & 7;
                if (k0 == 7) {                                                           // one thread of the eight: the variable id
                    const uint32_t vid = eg.var_id[e];
                    atomicOr(&hnat[vid >> 5], 1u << (vid & 31));
                }
                if (need_wt) {
                    const uint32_t dx = cmeta[t];
                    const int D = (dx >> 16) & 63, H = (D + 1) >> 1;
                    const uint32_t *rp = eg.col_rowpos + (dx & 0xFFFFu) * 32;
                    for (int k = k0; k < D; k += 8) {
                        const uint32_t w = __ldg(&rp[edge_idx_off(H, k >> 1, b)]);
                        const uint32_t pos = (k & 1) ? (w >> 16) : (w & 0xFFFFu);
                        atomicXor(&par[pos >> 5], 1u << (pos & 31));
                    }
                }
            }
        } else {                                                                         // very heavy hard decision: per-slice walk
            for (int t = tid; t < eg.n_csl; t += THREADS) {
                uint32_t bits = hperm[t];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t vid = eg.var_id[t * 32 + b];
                    atomicOr(&hnat[vid >> 5], 1u << (vid & 31));
                }
            }
            if (need_wt) parity_of_hard(eg, hperm, cmeta, par, s_plist, &s_pcount, tid, THREADS);
        }
        __syncthreads();
        if (need_wt && warp == 0) { const int w = residual_weight(par, syn, eg.n_rsl, lane); if (lane == 0) s_wt = w; }
        for (int w = tid; w < eg.nw; w += THREADS) a.hard_bits[(size_t)shot * eg.nw + w] = hnat[w];
        const int done_shot = shot;
        shot = s_next;
        __syncthreads();
        const int s_wt_done = need_wt ? s_wt : 0;                                       // (next written after two more barriers)
        // bookkeeping of the finished shot after the barrier: the round trip of the queue atomic overlaps the next
        // shot's syndrome load instead of holding all warps at the barrier
        if (tid == THREADS - 1) {
            a.converged[done_shot] = conv ? 1 : 0;
            a.final_iter[done_shot] = fin;
            if (!conv && a.fail_count) {
                const int slot = atomicAdd(a.fail_count, 1);
                a.fail_idx[slot] = done_shot;
                if (a.fail_wt) a.fail_wt[slot] = s_wt_done;
            }
        }
    }
#ifdef QB_EDGE_PROFILE
    if (lane == 0 && blockIdx.x < 256)
        for (int i = 0; i < 4; ++i) g_edge_prof[(blockIdx.x * 32 + warp) * 4 + i] = prof[i];
#endif
}

// ------------------------------------------------------------------------------------------------------------
static size_t edge_smem_bytes(const EdgeLayout &L, int nw)
{
    return (size_t)L.e_words * 4 + (size_t)L.idx_words * 4 + (size_t)L.n_csl * 8 + (size_t)L.n_rsl * 8 +
           (size_t)L.n_rsl * 8 + (size_t)L.n_csl * 4 + (size_t)nw * 4 + (size_t)L.n_csl * 32 + 64;
}

template <class T>
static int up(EdgePlan *p, const std::vector<T> &h, const T **out)
{
    T *d = nullptr;
    QB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d), sizeof(T) * std::max<size_t>(1, h.size())));
    p->owned.push_back(d);
    if (!h.empty()) QB_CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    *out = d;
    return QB_OK;
}

void edge_plan_destroy(EdgePlan *p)
{
    if (!p) return;
    for (void *q : p->owned) cudaFree(q);
    delete p;
}

static std::vector<float> edge_E0(const EdgeLayout &L, const float *prior)
{
    std::vector<float> e0(L.e_words, INFINITY);
    for (int i = 0; i < L.e_words; ++i) {
        if (L.slot_var[i] >= 0) e0[i] = prior[L.slot_var[i]] + 0.0f;   // -0.0 -> +0.0
        else if (L.slot_var[i] == -2) e0[i] = 0.f;
    }
    return e0;
}

// Build (or rebuild after a prior change) the per-edge plan of a decoder.  *out = nullptr when the graph
// does not fit the kernel (the caller falls back to the compressed-state kernel).
int edge_plan_create(const qb_decoder *dec, const float *prior_h, EdgePlan **out)
{
    *out = nullptr;
    if (getenv("QLDPC_B200_NO_EDGE")) return QB_OK;
    const GraphDev &g = dec->g;
    if (g.m <= 0 || g.n <= 0 || g.nnz <= 0) return QB_OK;
    const size_t limit = (size_t)dec->max_smem_optin;
    int nwarps = 32;
    EdgeLayout L = build_edge_layout(g.m, g.n, dec->h_indptr.data(), dec->h_indices.data(), prior_h, nwarps);
    if (!L.ok || L.n_csl > EDGE_MAX_CSL) return QB_OK;
    size_t smem = edge_smem_bytes(L, g.nw);
    if (smem + 2048 > limit) return QB_OK;             // 2 KB: the kernel's static shared memory
    // several CTAs per SM for small codes: fewer warps each
    int ctas = (int)std::min<size_t>(4, (limit + 1024) / (smem + 2048 + 1024));
    if (const char *e = getenv("QLDPC_B200_EDGE_CTAS")) ctas = std::max(1, std::min(ctas, atoi(e)));
    int want = ctas >= 3 ? 8 : (ctas >= 2 ? 16 : 32);     // 3-4 CTAs of 8 warps, 2 of 16, 1 of 32 per SM (64 registers per thread)
    if (const char *e = getenv("QLDPC_B200_EDGE_WARPS")) { const int w = atoi(e); if (w == 8 || w == 16 || w == 32) want = w; }
    if (want != nwarps) {
        // (a layout can fail with few warps -- more than 32 column slices per warp -- and still work with 32)
        EdgeLayout L2 = build_edge_layout(g.m, g.n, dec->h_indptr.data(), dec->h_indices.data(), prior_h, want);
        if (L2.ok && edge_smem_bytes(L2, g.nw) + 2048 <= limit) {
            nwarps = want;
            smem = edge_smem_bytes(L2, g.nw);
            L = std::move(L2);
        } else {
            ctas = 1;
        }
    }
    EdgePlan *p = new EdgePlan();
    p->threads = nwarps * 32; p->ctas_per_sm = ctas; p->smem = smem;
    EdgeDev &d = p->dev;
    d.n_rsl = L.n_rsl; d.n_csl = L.n_csl; d.e_words = L.e_words; d.e_dummy = L.e_dummy; d.idx_words = L.idx_words;
    d.nw = g.nw; d.n = g.n; d.mw = g.mw;
    d.o_idx = (uint32_t)L.e_words * 4; d.o_rtask = d.o_idx + (uint32_t)L.idx_words * 4; d.o_syn = d.o_rtask + (uint32_t)L.n_rsl * 8;
    d.o_par = d.o_syn + (uint32_t)L.n_rsl * 4; d.o_hperm = d.o_par + (uint32_t)L.n_rsl * 4; d.o_hnat = d.o_hperm + (uint32_t)L.n_csl * 4;
    d.o_cmeta = d.o_hnat + (uint32_t)g.nw * 4; d.o_csig = d.o_cmeta + (uint32_t)L.n_csl * 4;       // (the order edge_smem_bytes() adds up)
    int rc = QB_OK;
    const std::vector<float> e0 = edge_E0(L, prior_h);
    const float *pf = nullptr; const uint32_t *pu = nullptr; const uint16_t *ph = nullptr; const int32_t *pi = nullptr;
    if (!rc) { rc = up(p, e0, &pf); d.E0 = reinterpret_cast<const float4 *>(pf); p->d_E0 = const_cast<float *>(pf); }
    if (!rc) { rc = up(p, L.col_idx, &pu); d.col_idx = pu; }
    if (!rc) { rc = up(p, L.col_rowpos, &pu); d.col_rowpos = pu; }
    if (!rc) { rc = up(p, L.rtask, &pu); d.rtask = reinterpret_cast<const uint2 *>(pu); }
    if (!rc) { rc = up(p, L.ctask, &pu); d.ctask = reinterpret_cast<const uint2 *>(pu); }
    if (!rc) { rc = up(p, L.row_id, &ph); d.row_id = ph; }
    if (!rc) { rc = up(p, L.row_pads, &ph); d.row_pads = reinterpret_cast<const uint2 *>(ph); }
    if (!rc) { rc = up(p, L.var_id, &ph); d.var_id = ph; }
    d.lane_prior = nullptr;
    if (!rc && !L.uniform_prior) { rc = up(p, L.lane_prior, &pf); d.lane_prior = pf; }
    if (!rc) { rc = up(p, L.wr_ptr, &pi); d.wr_ptr = pi; }
    if (!rc) { rc = up(p, L.wc_ptr, &pi); d.wc_ptr = pi; }
    { const uint8_t *p8 = nullptr; if (!rc) { rc = up(p, L.wc_cls, &p8); d.wc_cls = reinterpret_cast<const uint4 *>(p8); } }
    if (!rc) { rc = up(p, L.row_mask, &pu); d.row_mask = pu; }
    if (!rc) { rc = up(p, L.col_sig, &pu); d.col_sig = pu; }
    if (!rc) { std::vector<int> z(1, 0); const int *pc = nullptr; rc = up(p, z, &pc); p->d_counter = const_cast<int *>(pc); }
    if (rc) { edge_plan_destroy(p); return rc; }
    for (int t = 0; t < L.n_csl; ++t) p->pri.bits[t] = L.ctask[2 * t + 1];
    p->L = std::move(L);
    *out = p;
    return QB_OK;
}

template <int THREADS, int MINB>
static int launch_edge_t(EdgePlan *p, const MinsumLaunch &a, int grid, cudaStream_t st)
{
    QB_CUDA(cudaFuncSetAttribute(minsum_edge_kernel<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    minsum_edge_kernel<THREADS, MINB><<<grid, THREADS, p->smem, st>>>(p->dev, a, p->d_counter, p->pri);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

#ifdef QB_EDGE_PROFILE
extern "C" int qb_debug_edge_profile(unsigned long long *out_h)
{
    return cudaMemcpyFromSymbol(out_h, g_edge_prof, sizeof(g_edge_prof)) == cudaSuccess ? 0 : -2;
}
#endif

int launch_minsum_edge(qb_decoder *dec, EdgePlan *p, const MinsumLaunch &a, cudaStream_t st)
{
    QB_CUDA(cudaMemsetAsync(p->d_counter, 0, sizeof(int), st));
    const int grid = std::max(1, std::min(a.B, dec->sm_count * p->ctas_per_sm));
    if (p->threads == 1024) return launch_edge_t<1024, 1>(p, a, grid, st);
    if (p->threads == 512 && p->ctas_per_sm == 1) return launch_edge_t<512, 1>(p, a, grid, st);
    if (p->threads == 512) return launch_edge_t<512, 2>(p, a, grid, st);
    return launch_edge_t<256, 4>(p, a, grid, st);
}

}  // namespace qb
