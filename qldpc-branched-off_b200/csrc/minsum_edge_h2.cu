// K3, opt-in packed mode: flooding min-sum with TWO shots per 32-bit shared-memory slot (half2).
//
// Same plan, same slot layout, same phases as minsum_edge.cu (reference recurrence: src/decoding/kernels.py:235-366 with
// damping == 1); slot s holds { shot A, shot B } as two IEEE half floats, so every LDS / STS / index word / address
// computation and every packed arithmetic instruction (HMNMX2.XORSIGN, HSET2, HADD2) serves two shots.  Measured issue
// rates on B200 (tools/micro/h2_rate.cu): HMNMX2 / HSET2 / HADD2 run at the rate of FMNMX (2 warp instructions per clock
// per SM), i.e. twice the shots per alu-pipe cycle in the check rows, and the variable phase -- bound by shared-memory
// instruction issue -- needs half the instructions per shot.
//
// This is NOT the reference's arithmetic: messages carry 11 significant bits (clip 20 -> steps of 1/64 at the top of
// the range), alpha_it = 1 - 2^-(it+1) is exact only up to it = 10 (1.0 afterwards), sums of up to 16 messages are
// rounded to half after every addition.  What is kept in float32: the posteriors written for OSD (accumulated in
// float32 from the half messages in the last iteration) and the hard decision of that iteration.  The mode is selected
// explicitly (qb_decode_config.precision = QB_PRECISION_HALF2 / qb_decoder_set_precision); the default and every
// parity claim of the package are float32.  DESIGN.md has the measured agreement with the float64 recurrence and the
// logical error rates.
//
//   check rows     per slot: t = max.xorsign.abs(v, m1s); m2 = min.xorsign.abs(m2, t); m1s = min.xorsign.abs(m1s, v)
//                  (three HMNMX2 per two edge-messages: |m1s| = running minimum, sign = running sign product, |m2| = second
//                  minimum), then R = sign * alpha * (|Q| == min1 ? min2 : min1) with one HSET2 mask + two LOP3 per slot.
//   variables      gather, HADD2 in row order + prior, Q = v - R (HADD2 with negation), one packed sign test per slice
//                  and two ballots (one hard-decision word per shot).
//   convergence    per shot exactly as in the float32 kernel (8-bit fingerprint, exact parity on a match); a shot that
//                  converges is written out at once and its half of the slots keeps iterating harmlessly until the
//                  partner is done.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "edge_dev.cuh"
#include "edge_layout.h"

namespace qb {

constexpr uint32_t H2_INF = 0x7C007C00u, H2_ABS = 0x7FFF7FFFu, H2_SIGN = 0x80008000u;

__device__ __forceinline__ uint32_t h2_min_xs(uint32_t a, uint32_t b) { uint32_t d; asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_max_xs(uint32_t a, uint32_t b) { uint32_t d; asm("max.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b) { uint32_t d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) { uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { uint32_t d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_eq_mask(uint32_t a, uint32_t b) { uint32_t d; asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_lt_mask(uint32_t a, uint32_t b) { uint32_t d; asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2_nan_mask(uint32_t a) { uint32_t d; asm("set.nan.u32.f16x2 %0, %1, %1;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ float2 h2_to_f2(uint32_t a) { return __half22float2(*reinterpret_cast<const __half2 *>(&a)); }
__device__ __forceinline__ uint32_t f_to_h2(float a) { const __half2 h = __float2half2_rn(a); return *reinterpret_cast<const uint32_t *>(&h); }

// ---- check rows: one lane per row, K chunks of 4 slots in registers ------------------------------------------------
template <int K, bool FIRST>
__device__ __forceinline__ void row_task_h2(uint32_t *E, const uint4 *E0, int base_unit, int stride, int lane,
                                            uint32_t synsign2, uint32_t alpha2, uint32_t clip2, uint2 pads)
{
    uint4 q[K];
    uint4 *e4 = reinterpret_cast<uint4 *>(E) + base_unit + lane;
    if constexpr (FIRST) {
        const uint4 *g4 = E0 + base_unit + lane;
#pragma unroll
        for (int c = 0; c < K; ++c) q[c] = __ldg(g4 + c * stride);
    } else {
#pragma unroll
        for (int c = 0; c < K; ++c) q[c] = e4[c * stride];
    }
    uint32_t m1s = H2_INF, m2 = H2_INF;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const uint32_t v[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t t = h2_max_xs(v[i], m1s);
            m2 = h2_min_xs(m2, t);
            m1s = h2_min_xs(m1s, v[i]);
        }
    }
    const uint32_t m1 = m1s & H2_ABS;
    const uint32_t tot = (m1s & H2_SIGN) ^ synsign2;                               // kernels.py:289-298, per shot
    uint32_t clipB = clip2;
    if constexpr (K == 1) clipB = ((pads.y & 0xFFFFu) != 0xFFFFu) ? H2_INF : clip2;     // degree-1 row: min2 stays +inf
    uint32_t a1 = h2_mul(alpha2, h2_min(m1, clip2)) ^ tot, a2 = h2_mul(alpha2, h2_min(m2 & H2_ABS, clipB)) ^ tot;
    uint32_t d = a1 ^ a2;
    asm volatile("" : "+r"(a1), "+r"(d));
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const uint32_t v[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t mask = h2_eq_mask(v[i] & H2_ABS, m1);                     // 0xFFFF in the half that holds the minimum
            r[i] = (a1 ^ (d & mask)) ^ (v[i] & H2_SIGN);
        }
        e4[c * stride] = make_uint4(r[0], r[1], r[2], r[3]);
    }
    E[pads.x & 0xFFFFu] = H2_INF;
    if ((pads.x >> 16) != 0xFFFFu) E[pads.x >> 16] = H2_INF;
    if ((pads.y & 0xFFFFu) != 0xFFFFu) E[pads.y & 0xFFFFu] = H2_INF;
    if ((pads.y >> 16) != 0xFFFFu) E[pads.y >> 16] = H2_INF;
}

template <bool FIRST>
__device__ __noinline__ void row_task_loop_h2(uint32_t *E, const uint4 *E0, int base_unit, int stride, int lane, int K,
                                              uint32_t synsign2, uint32_t alpha2, uint32_t clip2, uint2 pads)
{
    uint4 *e4 = reinterpret_cast<uint4 *>(E) + base_unit + lane;
    const uint4 *g4 = E0 + base_unit + lane;
    uint32_t m1s = H2_INF, m2 = H2_INF;
    for (int c = 0; c < K; ++c) {
        const uint4 qq = FIRST ? __ldg(g4 + c * stride) : e4[c * stride];
        const uint32_t v[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint32_t t = h2_max_xs(v[i], m1s); m2 = h2_min_xs(m2, t); m1s = h2_min_xs(m1s, v[i]); }
    }
    const uint32_t m1 = m1s & H2_ABS;
    const uint32_t tot = (m1s & H2_SIGN) ^ synsign2;
    const uint32_t a1 = h2_mul(alpha2, h2_min(m1, clip2)) ^ tot, a2 = h2_mul(alpha2, h2_min(m2 & H2_ABS, clip2)) ^ tot;
    const uint32_t d = a1 ^ a2;
    for (int c = 0; c < K; ++c) {
        const uint4 qq = FIRST ? __ldg(g4 + c * stride) : e4[c * stride];
        const uint32_t v[4] = {qq.x, qq.y, qq.z, qq.w};
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = (a1 ^ (d & h2_eq_mask(v[i] & H2_ABS, m1))) ^ (v[i] & H2_SIGN);
        e4[c * stride] = make_uint4(r[0], r[1], r[2], r[3]);
    }
    E[pads.x & 0xFFFFu] = H2_INF;
    if ((pads.x >> 16) != 0xFFFFu) E[pads.x >> 16] = H2_INF;
    if ((pads.y & 0xFFFFu) != 0xFFFFu) E[pads.y & 0xFFFFu] = H2_INF;
    if ((pads.y >> 16) != 0xFFFFu) E[pads.y >> 16] = H2_INF;
}

template <bool FIRST>
__device__ __forceinline__ void row_dispatch_h2(uint32_t *E, const uint4 *E0, int base_unit, int stride, int lane, int K,
                                                uint32_t synsign2, uint32_t alpha2, uint32_t clip2, uint2 pads)
{
    switch (K) {
    case 1: row_task_h2<1, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 2: row_task_h2<2, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 3: row_task_h2<3, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 4: row_task_h2<4, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 5: row_task_h2<5, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 6: row_task_h2<6, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 7: row_task_h2<7, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 8: row_task_h2<8, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    case 9: row_task_h2<9, FIRST>(E, E0, base_unit, stride, lane, synsign2, alpha2, clip2, pads); break;
    default: row_task_loop_h2<FIRST>(E, E0, base_unit, stride, lane, K, synsign2, alpha2, clip2, pads); break;
    }
}

// ---- variables -----------------------------------------------------------------------------------------------------
struct ColCtxH2 {
    uint32_t ix, lane4, lane8, sg;
    uint32_t fp2;               // XOR of the fingerprints of the variables whose hard decision is 1: shot A low half, shot B high half
    uint32_t hwA, hwB;          // lane j keeps the hard-decision words of the warp's j-th task
    uint32_t t4, lane_t4;
    int lane;
    const uint16_t *vid;
    float *postA, *postB;       // posterior rows (nullptr: shot absent / not wanted)
};

template <int D>
__device__ __forceinline__ void load_idx_words_h2(const ColCtxH2 &c, uint32_t (&w)[(D + 1) / 2 + 1])
{
    constexpr int H = (D + 1) / 2;
#pragma unroll
    for (int u = 0; u < H / 2; ++u) {
        const uint2 p = lds_u64(c.ix + u * 256 + c.lane8);
        w[2 * u] = p.x; w[2 * u + 1] = p.y;
    }
    if constexpr (H & 1) w[H - 1] = lds_u32(c.ix + (H / 2) * 256 + c.lane4);
}

// one full slice of degree D with a uniform prior; WRITE_V: float32 posteriors + hard decision from them (last iteration)
template <int D, bool EXACT, bool WRITE_V>
__device__ __forceinline__ void col_task_h2(ColCtxH2 &c, const EdgePriors &pri, const EdgePriors &pri_h2)
{
    uint32_t w[(D + 1) / 2 + 1];
    uint32_t addr[D + 1], r[D + 1];
    load_idx_words_h2<D>(c, w);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        addr[k] = (k & 1) ? ((w[k >> 1] >> 14) & 0x3FFFCu) : ((w[k >> 1] << 2) & 0x3FFFCu);
        r[k] = lds_u32v(addr[k]);
    }
    uint32_t acc = D > 0 ? r[0] : 0u;
#pragma unroll
    for (int k = 1; k < D; ++k) acc = h2_add(acc, r[k]);
    const uint32_t v = h2_add(acc, *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(pri_h2.bits) + c.t4));
#pragma unroll
    for (int k = 0; k < D; ++k) {
        uint32_t q = h2_sub(v, r[k]);
        if constexpr (EXACT) q &= ~h2_nan_mask(q);                               // kernels.py:328-329
        sts_u32(addr[k], q);
    }
    uint32_t lt;                                                                  // 0xFFFF per half whose hard decision is 1
    if constexpr (WRITE_V) {
        float2 s = D > 0 ? h2_to_f2(r[0]) : make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 1; k < D; ++k) { const float2 x = h2_to_f2(r[k]); s.x += x.x; s.y += x.y; }
        const float pf = __uint_as_float(*reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(pri.bits) + c.t4));
        s.x += pf; s.y += pf;
        lt = (s.x < 0.f ? 0xFFFFu : 0u) | (s.y < 0.f ? 0xFFFF0000u : 0u);
        const uint32_t vid = __ldg(c.vid);
        if (vid != 0xFFFFu) { if (c.postA) c.postA[vid] = s.x; if (c.postB) c.postB[vid] = s.y; }
    } else {
        lt = h2_lt_mask(v, 0u);
    }
    const uint32_t sig = lds_u8(c.sg);
    c.fp2 ^= __byte_perm(sig, 0u, 0x4040) & lt;                                   // sig in both halves
    const uint32_t ha = __ballot_sync(0xFFFFFFFFu, (lt & 0xFFFFu) != 0u), hb = __ballot_sync(0xFFFFFFFFu, (lt >> 16) != 0u);
    if (c.lane_t4 == c.t4) { c.hwA = ha; c.hwB = hb; }
    c.vid += 32;
    c.ix += ((D + 1) / 2) * 128; c.sg += 32; c.t4 += 4;
}

// any slice: partial with a negative prior, per-lane priors, large degree
template <bool WRITE_V>
__device__ __forceinline__ void col_task_generic_h2(ColCtxH2 &c, uint32_t meta, const float *lane_prior, const EdgePriors &pri)
{
    const int D = (meta >> 16) & 63, nl = (meta >> 22) & 63, H = (D + 1) >> 1;
    bool negA = false, negB = false;
    if (c.lane < nl) {
        const float pf = lane_prior ? __ldg(lane_prior) : __uint_as_float(pri.bits[c.t4 >> 2]);
        uint32_t acc = 0u;
        float2 s = make_float2(0.f, 0.f);
        for (int k = 0; k < D; ++k) {
            const uint32_t w = lds_u32(c.ix + edge_idx_off(H, k >> 1, c.lane) * 4);
            const uint32_t rr = lds_u32v((k & 1) ? ((w >> 14) & 0x3FFFCu) : ((w << 2) & 0x3FFFCu));
            acc = k == 0 ? rr : h2_add(acc, rr);
            if (WRITE_V) { const float2 x = h2_to_f2(rr); s.x += x.x; s.y += x.y; }
        }
        const uint32_t v = D > 0 ? h2_add(acc, f_to_h2(pf)) : f_to_h2(pf);
        for (int k = 0; k < D; ++k) {
            const uint32_t w = lds_u32(c.ix + edge_idx_off(H, k >> 1, c.lane) * 4);
            const uint32_t ad = (k & 1) ? ((w >> 14) & 0x3FFFCu) : ((w << 2) & 0x3FFFCu);
            uint32_t q = h2_sub(v, lds_u32v(ad));
            q &= ~h2_nan_mask(q);
            sts_u32(ad, q);
        }
        if (WRITE_V) {
            s.x += pf; s.y += pf;
            negA = s.x < 0.f; negB = s.y < 0.f;
            const uint32_t vid = __ldg(c.vid);
            if (c.postA) c.postA[vid] = s.x;
            if (c.postB) c.postB[vid] = s.y;
        } else {
            const uint32_t lt = h2_lt_mask(v, 0u);
            negA = (lt & 0xFFFFu) != 0u; negB = (lt >> 16) != 0u;
        }
        const uint32_t sig = lds_u8(c.sg);
        if (negA) c.fp2 ^= sig;
        if (negB) c.fp2 ^= sig << 16;
    }
    const uint32_t ha = __ballot_sync(0xFFFFFFFFu, negA), hb = __ballot_sync(0xFFFFFFFFu, negB);
    if (c.lane_t4 == c.t4) { c.hwA = ha; c.hwB = hb; }
    c.vid += 32;
    c.ix += H * 128; c.sg += 32; c.t4 += 4;
}

template <int D, bool EXACT, bool WRITE_V>
__device__ __forceinline__ void col_class_h2(ColCtxH2 &c, int cnt, const EdgePriors &pri, const EdgePriors &pri_h2)
{
    const uint32_t t4_end = c.t4 + 4u * (uint32_t)cnt;
#pragma unroll 1
    while (c.t4 != t4_end) col_task_h2<D, EXACT, WRITE_V>(c, pri, pri_h2);
}

template <bool WRITE_V>
__device__ __forceinline__ void phase_b_h2(ColCtxH2 &c, uint4 cls, int t_end, const uint32_t *cmeta, const float *lane_prior,
                                           const EdgePriors &pri, const EdgePriors &pri_h2)
{
    col_class_h2<0, false, WRITE_V>(c, cls.x & 255, pri, pri_h2);
    col_class_h2<1, false, WRITE_V>(c, (cls.x >> 8) & 255, pri, pri_h2);
    col_class_h2<2, false, WRITE_V>(c, (cls.x >> 16) & 255, pri, pri_h2);
    col_class_h2<3, false, WRITE_V>(c, cls.x >> 24, pri, pri_h2);
    col_class_h2<4, false, WRITE_V>(c, cls.y & 255, pri, pri_h2);
    col_class_h2<5, false, WRITE_V>(c, (cls.y >> 8) & 255, pri, pri_h2);
    col_class_h2<6, false, WRITE_V>(c, (cls.y >> 16) & 255, pri, pri_h2);
    if (c.t4 >= 4u * (uint32_t)t_end) return;
    col_class_h2<7, false, WRITE_V>(c, cls.y >> 24, pri, pri_h2);
    col_class_h2<8, false, WRITE_V>(c, cls.z & 255, pri, pri_h2);
    col_class_h2<1, true, WRITE_V>(c, (cls.z >> 8) & 255, pri, pri_h2);
    col_class_h2<2, true, WRITE_V>(c, (cls.z >> 16) & 255, pri, pri_h2);
    col_class_h2<3, true, WRITE_V>(c, cls.z >> 24, pri, pri_h2);
    col_class_h2<4, true, WRITE_V>(c, cls.w & 255, pri, pri_h2);
    col_class_h2<5, true, WRITE_V>(c, (cls.w >> 8) & 255, pri, pri_h2);
    col_class_h2<6, true, WRITE_V>(c, (cls.w >> 16) & 255, pri, pri_h2);
    const int ngen = cls.w >> 24;
    for (int i = 0; i < ngen; ++i)
        col_task_generic_h2<WRITE_V>(c, cmeta[c.t4 >> 2], lane_prior ? lane_prior + (c.t4 >> 2) * 32 + c.lane : nullptr, pri);
}

// write out one shot of the pair: hard decision to natural order, flags, failure queue (uniform call: contains barriers)
template <int THREADS>
__device__ __forceinline__ void finish_shot_h2(const EdgeDev &eg, const MinsumLaunch &a, const uint32_t *hperm, const uint32_t *syn,
                                               uint32_t *par, uint32_t *hnat, const uint32_t *cmeta, uint16_t *s_plist, int *s_pcount,
                                               int *s_wt, int shot, bool converged, int it_fin, int tid, int warp, int lane)
{
    const bool need_wt = !converged && a.max_iter > 0 && a.fail_wt != nullptr;
    for (int w = tid; w < eg.nw; w += THREADS) hnat[w] = 0u;
    __syncthreads();
    for (int t = tid; t < eg.n_csl; t += THREADS) {
        uint32_t bits = hperm[t];
        while (bits) {
            const int b = __ffs(bits) - 1; bits &= bits - 1;
            const uint32_t vid = eg.var_id[t * 32 + b];
            atomicOr(&hnat[vid >> 5], 1u << (vid & 31));
        }
    }
    if (need_wt) {
        parity_of_hard(eg, hperm, cmeta, par, s_plist, s_pcount, tid, THREADS);
        __syncthreads();
        if (warp == 0) { const int w = residual_weight(par, syn, eg.n_rsl, lane); if (lane == 0) *s_wt = w; }
    }
    __syncthreads();
    for (int w = tid; w < eg.nw; w += THREADS) a.hard_bits[(size_t)shot * eg.nw + w] = hnat[w];
    if (tid == 0) {
        a.converged[shot] = converged ? 1 : 0;
        a.final_iter[shot] = it_fin;
        if (!converged && a.fail_count) {
            const int slot = atomicAdd(a.fail_count, 1);
            a.fail_idx[slot] = shot;
            if (a.fail_wt) a.fail_wt[slot] = need_wt ? *s_wt : 0;
        }
    }
    __syncthreads();
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
minsum_edge_h2_kernel(const __grid_constant__ EdgeDev eg, const __grid_constant__ MinsumLaunch a, int *pair_counter,
                      const __grid_constant__ EdgePriors pri, const __grid_constant__ EdgePriors pri_h2, const uint4 *E0h)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *E = reinterpret_cast<uint32_t *>(smem_raw);                             // [e_words] half2 per slot
    uint32_t *idx = E + eg.e_words;                                                   // [idx_words]
    uint2 *rtask = reinterpret_cast<uint2 *>(idx + eg.idx_words);                     // [n_rsl]
    uint32_t *synA = reinterpret_cast<uint32_t *>(rtask + eg.n_rsl);                  // [n_rsl] permuted syndrome bits, shot A
    uint32_t *synB = synA + eg.n_rsl;
    uint32_t *par = synB + eg.n_rsl;                                                  // [n_rsl]
    uint32_t *hpermA = par + eg.n_rsl;                                                // [n_csl] hard decision word per column slice
    uint32_t *hpermB = hpermA + eg.n_csl;
    uint32_t *hnat = hpermB + eg.n_csl;                                               // [nw]
    uint32_t *cmeta = hnat + eg.nw;                                                   // [n_csl]
    uint8_t *csig = reinterpret_cast<uint8_t *>(cmeta + eg.n_csl);                    // [n_csl*32]
    __shared__ int s_wt, s_next, s_pcount;
    __shared__ uint16_t s_plist[PAR_LIST_CAP];
    __shared__ uint32_t s_alpha2[128];
    __shared__ uint32_t s_fp[2], s_target[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_min_sync(0xFFFFFFFFu, tid >> 5);
    const uint32_t idx_addr = (uint32_t)__cvta_generic_to_shared(idx);
    const uint32_t e_word = (uint32_t)__cvta_generic_to_shared(E) >> 2;
    for (int i = tid; i < eg.idx_words; i += THREADS) idx[i] = eg.col_idx[i] + (e_word | (e_word << 16));
    for (int i = tid; i < eg.n_csl; i += THREADS) { cmeta[i] = eg.ctask[i].x; hpermA[i] = 0u; hpermB[i] = 0u; }
    for (int i = tid; i < eg.n_rsl; i += THREADS) rtask[i] = eg.rtask[i];
    for (int i = tid; i < 128 && i < a.max_iter; i += THREADS) s_alpha2[i] = f_to_h2(a.alpha_d[i]);
    for (int i = tid; i < eg.n_csl * 32; i += THREADS) csig[i] = (uint8_t)eg.col_sig[i];
    if (tid < 32) E[eg.e_dummy + tid] = 0u;
    const int r0 = eg.wr_ptr[warp], r1 = eg.wr_ptr[warp + 1];
    const int c0 = eg.wc_ptr[warp], c1 = eg.wc_ptr[warp + 1];
    const uint4 cls = eg.wc_cls[warp];
    const uint32_t ix0 = idx_addr + (c0 < eg.n_csl ? (eg.ctask[c0].x & 0xFFFFu) * 128u : 0u);
    const int n_pairs = (a.B + 1) >> 1;
    if (tid == 0) { s_next = atomicAdd(pair_counter, 1); s_target[0] = 0u; s_target[1] = 0u; }
    __syncthreads();
    int pair = s_next;
    const bool api = !a.post_failed_only;
    const uint32_t clip2 = f_to_h2(a.clip);

    while (pair < n_pairs) {
        const int shotA = 2 * pair, shotB = 2 * pair + 1;
        const bool presentB = shotB < a.B;
        // ---- load both syndromes (permuted) and their fingerprints ----
        for (int t = warp; t < eg.n_rsl; t += THREADS / 32) {
            const uint32_t rid = eg.row_id[t * 32 + lane];
            const uint32_t msk = eg.row_mask[t * 32 + lane] & 0xFFu;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int sh = s ? shotB : shotA;
                const bool bit = (s == 0 || presentB) && rid != 0xFFFFu && ((a.syn_bits[(size_t)sh * eg.mw + (rid >> 5)] >> (rid & 31)) & 1u);
                const uint32_t tg = __reduce_xor_sync(0xFFFFFFFFu, bit ? msk : 0u);
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
                if (lane == 0) { (s ? synB : synA)[t] = word; if (tg) atomicXor(&s_target[s], tg); }
            }
            if (lane == 0) par[t] = 0u;
        }
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(pair_counter, 1);
        const uint32_t targetA = s_target[0], targetB = s_target[1];
        bool doneA = false, doneB = !presentB;
        for (int it = 0; it < a.max_iter; ++it) {
            const uint32_t alpha2 = it < 128 ? s_alpha2[it] : f_to_h2(a.alpha_d[it]);
            for (int t = r0; t < r1; ++t) {
                const uint2 d = rtask[t];
                const int K = d.y & 255, nl = (d.y >> 8) & 255, stride = d.y >> 16;
                if (K == 0 || lane >= nl) continue;
                const uint32_t synsign2 = (((synA[t] >> lane) & 1u) << 15) | (((synB[t] >> lane) & 1u) << 31);
                const uint2 pads = __ldg(&eg.row_pads[t * 32 + lane]);
                if (it == 0) row_dispatch_h2<true>(E, E0h, (int)(d.x >> 2), stride, lane, K, synsign2, alpha2, H2_INF, pads);
                else row_dispatch_h2<false>(E, E0h, (int)(d.x >> 2), stride, lane, K, synsign2, alpha2, clip2, pads);
            }
            if (tid == 0) s_fp[it & 1] = 0u;
            __syncthreads();
            const bool write_v = a.post && (api || it == a.max_iter - 1);
            ColCtxH2 c;
            c.ix = ix0; c.lane4 = lane * 4; c.lane8 = lane * 8;
            c.sg = (uint32_t)__cvta_generic_to_shared(csig + c0 * 32 + lane);
            c.fp2 = 0u; c.hwA = 0u; c.hwB = 0u;
            c.t4 = 4u * (uint32_t)c0; c.lane_t4 = 4u * (uint32_t)(c0 + lane); c.lane = lane;
            c.vid = eg.var_id + c0 * 32 + lane;
            c.postA = (a.post && !doneA) ? a.post + (size_t)shotA * eg.n : nullptr;
            c.postB = (a.post && !doneB) ? a.post + (size_t)shotB * eg.n : nullptr;
            if (write_v) phase_b_h2<true>(c, cls, c1, cmeta, eg.lane_prior, pri, pri_h2);
            else phase_b_h2<false>(c, cls, c1, cmeta, eg.lane_prior, pri, pri_h2);
            if (lane < c1 - c0) { if (!doneA) hpermA[c0 + lane] = c.hwA; if (!doneB) hpermB[c0 + lane] = c.hwB; }
            const uint32_t f2 = __reduce_xor_sync(0xFFFFFFFFu, c.fp2);
            if (lane == 0 && f2) atomicXor(&s_fp[it & 1], f2);
            __syncthreads();
            const uint32_t fpw = s_fp[it & 1];
            if (!doneA && (fpw & 0xFFFFu) == targetA) {                                   // uniform
                parity_of_hard(eg, hpermA, cmeta, par, s_plist, &s_pcount, tid, THREADS);
                __syncthreads();
                if (warp == 0) { const int w = residual_weight(par, synA, eg.n_rsl, lane); if (lane == 0) s_wt = w; }
                __syncthreads();
                if (s_wt == 0) {                                                          // kernels.py:352-364
                    doneA = true;
                    finish_shot_h2<THREADS>(eg, a, hpermA, synA, par, hnat, cmeta, s_plist, &s_pcount, &s_wt, shotA, true, it, tid, warp, lane);
                }
            }
            if (!doneB && (fpw >> 16) == targetB) {
                parity_of_hard(eg, hpermB, cmeta, par, s_plist, &s_pcount, tid, THREADS);
                __syncthreads();
                if (warp == 0) { const int w = residual_weight(par, synB, eg.n_rsl, lane); if (lane == 0) s_wt = w; }
                __syncthreads();
                if (s_wt == 0) {
                    doneB = true;
                    finish_shot_h2<THREADS>(eg, a, hpermB, synB, par, hnat, cmeta, s_plist, &s_pcount, &s_wt, shotB, true, it, tid, warp, lane);
                }
            }
            if (doneA && doneB) break;
        }
        if (!doneA) finish_shot_h2<THREADS>(eg, a, hpermA, synA, par, hnat, cmeta, s_plist, &s_pcount, &s_wt, shotA, false, a.max_iter - 1, tid, warp, lane);
        if (!doneB) finish_shot_h2<THREADS>(eg, a, hpermB, synB, par, hnat, cmeta, s_plist, &s_pcount, &s_wt, shotB, false, a.max_iter - 1, tid, warp, lane);
        if (tid == 0) { s_target[0] = 0u; s_target[1] = 0u; }
        pair = s_next;
        __syncthreads();
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
struct EdgePlanH2 {
    uint4 *d_E0h = nullptr;
    int *d_counter = nullptr;
    EdgePriors pri_h2{};
    size_t smem = 0;
};

static float __uint_as_float_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static uint16_t host_f2h(float f)
{
    const __half h = __float2half_rn(f);
    return *reinterpret_cast<const uint16_t *>(&h);
}

void edge_plan_h2_destroy(EdgePlanH2 *p)
{
    if (!p) return;
    if (p->d_E0h) cudaFree(p->d_E0h);
    if (p->d_counter) cudaFree(p->d_counter);
    delete p;
}

int edge_plan_h2_create(const EdgePlan *ep, int nw, const float *prior_h, EdgePlanH2 **out)
{
    *out = nullptr;
    const EdgeLayout &L = ep->L;
    std::vector<uint32_t> e0(L.e_words, H2_INF);
    for (int i = 0; i < L.e_words; ++i) {
        if (L.slot_var[i] >= 0) { const uint32_t h = host_f2h(prior_h[L.slot_var[i]] + 0.0f); e0[i] = h | (h << 16); }
        else if (L.slot_var[i] == -2) e0[i] = 0u;
    }
    EdgePlanH2 *p = new EdgePlanH2();
    if (cudaMalloc(reinterpret_cast<void **>(&p->d_E0h), sizeof(uint32_t) * e0.size()) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&p->d_counter), sizeof(int)) != cudaSuccess ||
        cudaMemcpy(p->d_E0h, e0.data(), sizeof(uint32_t) * e0.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        edge_plan_h2_destroy(p);
        return cuda_fail(cudaGetLastError(), "packed min-sum plan", __FILE__, __LINE__);
    }
    for (int t = 0; t < L.n_csl; ++t) {
        const uint32_t h = host_f2h(__uint_as_float_host(L.ctask[2 * t + 1]));
        p->pri_h2.bits[t] = h | (h << 16);
    }
    // the float32 kernel's shared memory + a second syndrome / hard-decision set
    p->smem = ep->smem + (size_t)L.n_rsl * 4 + (size_t)L.n_csl * 4;
    (void)nw;
    *out = p;
    return QB_OK;
}

template <int THREADS, int MINB>
static int launch_h2_t(EdgePlan *ep, EdgePlanH2 *hp, const MinsumLaunch &a, int grid, cudaStream_t st)
{
    QB_CUDA(cudaFuncSetAttribute(minsum_edge_h2_kernel<THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hp->smem));
    minsum_edge_h2_kernel<THREADS, MINB><<<grid, THREADS, hp->smem, st>>>(ep->dev, a, hp->d_counter, ep->pri, hp->pri_h2, hp->d_E0h);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

bool edge_h2_fits(const qb_decoder *dec, const EdgePlan *ep, const EdgePlanH2 *hp)
{
    return hp && hp->smem + 2560 <= (size_t)dec->max_smem_optin && ep->L.uniform_prior;
}

int launch_minsum_edge_h2(qb_decoder *dec, EdgePlan *ep, EdgePlanH2 *hp, const MinsumLaunch &a, cudaStream_t st)
{
    QB_CUDA(cudaMemsetAsync(hp->d_counter, 0, sizeof(int), st));
    const int pairs = (a.B + 1) / 2;
    const int grid = std::max(1, std::min(pairs, dec->sm_count * ep->ctas_per_sm));
    if (ep->threads == 1024) return launch_h2_t<1024, 1>(ep, hp, a, grid, st);
    if (ep->threads == 512 && ep->ctas_per_sm == 1) return launch_h2_t<512, 1>(ep, hp, a, grid, st);
    if (ep->threads == 512) return launch_h2_t<512, 2>(ep, hp, a, grid, st);
    return launch_h2_t<256, 4>(ep, hp, a, grid, st);
}

}  // namespace qb
