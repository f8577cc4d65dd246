// The two phases of the per-edge min-sum iteration as device functions (shared by minsum_edge.cu, one CTA per shot, and
// minsum_edge_cluster.cu, one thread-block cluster per shot).  See minsum_edge.cu for the description of the phases.
#pragma once
#include <math.h>

#include "edge_dev.cuh"

namespace qb {

__device__ __forceinline__ float min_xorsign_abs(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// if (v < 0) { a ^= x; b |= y; } as two predicated instructions (the compiler's own form is a select + a logic op each)
__device__ __forceinline__ void xor_or_if_negative(float v, uint32_t &a, uint32_t x, uint32_t &b, uint32_t y)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %2, 0f00000000;\n\t@p xor.b32 %0, %0, %3;\n\t@p or.b32 %1, %1, %4;\n\t}"
        : "+r"(a), "+r"(b) : "f"(v), "r"(x), "r"(y));
}

__device__ __forceinline__ void or_if_negative(float v, uint32_t &b, uint32_t y)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(b) : "f"(v), "r"(y));
}

// ---- phase A: one row slice, K chunks of 4 edges per lane held in registers -------------------------------
// ONEPAD: every row of the slice has exactly one unused slot (degree 4K - 1; rtask.x bit 0), so only pads.x's lower half is set
template <int K, bool FIRST, bool ONEPAD = false>
__device__ __forceinline__ void row_task(float *E, const float4 *E0, int base_unit, int stride, int lane,
                                         uint32_t synsign, float alpha, float clip, uint2 pads)
{
    float4 q[K];
    float4 *e4 = reinterpret_cast<float4 *>(E) + base_unit + lane;
    if constexpr (FIRST) {
        const float4 *g4 = E0 + base_unit + lane;
#pragma unroll
        for (int c = 0; c < K; ++c) q[c] = __ldg(g4 + c * stride);
    } else {
#pragma unroll
        for (int c = 0; c < K; ++c) q[c] = e4[c * stride];
    }
    float m1s = INFINITY, m2 = INFINITY;                   // |m1s| = running minimum, sign(m1s) = running sign product
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const float v[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m2 = fminf(m2, fmaxf(fabsf(v[i]), fabsf(m1s)));
            m1s = min_xorsign_abs(m1s, v[i]);
        }
    }
    // E holds the unclipped v - R; clip (kernels.py:330-333) is monotone in |Q| and keeps the sign, so the two
    // smallest |clip(Q)| are min(., clip) of the two smallest |Q|: one clamp per row instead of one per edge
    // (iteration 0 uses the priors unclipped, kernels.py:263-265: clip = +inf there)
    const float m1 = fabsf(m1s);
    const uint32_t tot = (__float_as_uint(m1s) & 0x80000000u) ^ synsign;      // kernels.py:289-298
    // a row of degree 1 (three unused slots in a one-chunk row) has no second edge: its min2 stays +inf
    float clip2 = clip;
    if constexpr (K == 1 && !ONEPAD) clip2 = ((pads.y & 0xFFFFu) != 0xFFFFu) ? INFINITY : clip;
    const float A1 = alpha * fminf(m1, clip), A2 = alpha * fminf(m2, clip2);  // kernels.py:309-314, A1 <= A2
    // Measured alternatives that lost (B200): forming the magnitude on the idle fma pipe (t = |Q| - min1 scaled to
    // -inf unless 0, max(A2 + t, A1)): 6 % slower, the extra issue slots cost more than the alu-pipe relief; moving
    // sign(Q) into bit 31 with IMAD.HI + IMAD instead of LOP3: 3 % slower; a two-pass loop over chunks instead of
    // the fully unrolled row: 5 % slower.
    uint32_t a1 = __float_as_uint(A1) ^ tot, a2 = __float_as_uint(A2) ^ tot;
    asm volatile("" : "+r"(a1), "+r"(a2));                  // keep the multiplications out of the per-edge code
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const float v[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t sel = (fabsf(v[i]) == m1) ? a2 : a1;
            r[i] = __uint_as_float(sel ^ (__float_as_uint(v[i]) & 0x80000000u));
        }
        e4[c * stride] = make_float4(r[0], r[1], r[2], r[3]);
    }
    // unused slots back to +inf (every row has at least one)
    E[pads.x & 0xFFFFu] = INFINITY;
    if constexpr (!ONEPAD) {
        if ((pads.x >> 16) != 0xFFFFu) E[pads.x >> 16] = INFINITY;
        if ((pads.y & 0xFFFFu) != 0xFFFFu) E[pads.y & 0xFFFFu] = INFINITY;
        if ((pads.y >> 16) != 0xFFFFu) E[pads.y >> 16] = INFINITY;
    }
}

// generic row (K > 9): two passes over shared memory
template <bool FIRST>
__device__ __noinline__ void row_task_loop(float *E, const float4 *E0, int base_unit, int stride, int lane, int K,
                                           uint32_t synsign, float alpha, float clip, uint2 pads)
{
    float4 *e4 = reinterpret_cast<float4 *>(E) + base_unit + lane;
    const float4 *g4 = E0 + base_unit + lane;
    float m1s = INFINITY, m2 = INFINITY;
    for (int c = 0; c < K; ++c) {
        const float4 qq = FIRST ? __ldg(g4 + c * stride) : e4[c * stride];
        const float v[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m2 = fminf(m2, fmaxf(fabsf(v[i]), fabsf(m1s)));
            m1s = min_xorsign_abs(m1s, v[i]);
        }
    }
    const float m1 = fabsf(m1s);
    const uint32_t tot = (__float_as_uint(m1s) & 0x80000000u) ^ synsign;
    uint32_t a1 = __float_as_uint(alpha * fminf(m1, clip)) ^ tot, a2 = __float_as_uint(alpha * fminf(m2, clip)) ^ tot;
    asm volatile("" : "+r"(a1), "+r"(a2));
    for (int c = 0; c < K; ++c) {
        const float4 qq = FIRST ? __ldg(g4 + c * stride) : e4[c * stride];
        const float v[4] = {qq.x, qq.y, qq.z, qq.w};
        float r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t sel = (fabsf(v[i]) == m1) ? a2 : a1;
            r[i] = __uint_as_float(sel ^ (__float_as_uint(v[i]) & 0x80000000u));
        }
        e4[c * stride] = make_float4(r[0], r[1], r[2], r[3]);
    }
    E[pads.x & 0xFFFFu] = INFINITY;
    if ((pads.x >> 16) != 0xFFFFu) E[pads.x >> 16] = INFINITY;
    if ((pads.y & 0xFFFFu) != 0xFFFFu) E[pads.y & 0xFFFFu] = INFINITY;
    if ((pads.y >> 16) != 0xFFFFu) E[pads.y >> 16] = INFINITY;
}

template <bool FIRST>
__device__ __forceinline__ void row_dispatch(float *E, const float4 *E0, int base_unit, int stride, int lane, int K,
                                             uint32_t synsign, float alpha, float clip, uint2 pads)
{
    switch (K) {
    case 1: row_task<1, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 2: row_task<2, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 3: row_task<3, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 4: row_task<4, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 5: row_task<5, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 6: row_task<6, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 7: row_task<7, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 8: row_task<8, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 9: row_task<9, FIRST>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    default: row_task_loop<FIRST>(E, E0, base_unit, stride, lane, K, synsign, alpha, clip, pads); break;
    }
}

// slices whose rows all have exactly one unused slot (most rows of a regular code: the gross code's 792 rows of degree 35)
template <bool FIRST>
__device__ __forceinline__ void row_dispatch_onepad(float *E, const float4 *E0, int base_unit, int stride, int lane, int K,
                                                    uint32_t synsign, float alpha, float clip, uint32_t pad)
{
    const uint2 pads = make_uint2(pad, 0xFFFFFFFFu);
    switch (K) {
    case 1: row_task<1, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 2: row_task<2, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 3: row_task<3, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 4: row_task<4, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 5: row_task<5, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 6: row_task<6, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 7: row_task<7, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    case 8: row_task<8, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    default: row_task<9, FIRST, true>(E, E0, base_unit, stride, lane, synsign, alpha, clip, pads); break;
    }
}

// ---- phase B ---------------------------------------------------------------------------------------------
// state shared by the column tasks of one warp during one phase B (the index blocks, fingerprints and priors of
// consecutive tasks are consecutive in memory)
struct ColCtx {
    uint32_t ix;            // shared address of the index block of the next task
    uint32_t lane4, lane8;
    uint32_t sg;            // shared address of the 8-bit fingerprint of the lane's variable in slice 0 (slice t: + 32 t = + 8 t4)
    uint32_t fp;            // XOR of the fingerprints of the variables whose hard decision is 1
    uint32_t fpw;           // same for the variables whose fingerprint sits in the upper half of an index word (SIGW): bits 16-23
    uint32_t myhw;          // lane j keeps the hard-decision word of the warp's j-th task
    uint32_t negbits;       // SIGW kernels instead: bit j = hard decision of the lane's variable in the warp's j-th task (the words
    uint32_t ubit;          // are formed by ballots only when somebody reads them); ubit = 1 << (next task - first task), warp-uniform
    uint32_t t4;            // 4 * next task: byte offset of its prior, compared with lane_t4 to pick the lane that keeps its hard-decision word
    uint32_t lane_t4;       // 4 * (first task of the warp + lane)
    int lane;
    const uint16_t *vid;    // next task's variable id of the lane (posterior output)
    uint32_t vid_next;      // its value, loaded one task ahead
    float *post;            // posterior row of the shot
    uint32_t win;           // bits 24+ of the CTA's shared window (rank of the CTA in its cluster; 0 without clusters): the
                            // 16-bit slot indices are word offsets inside the window
};

// index words of task (t + j) of a run of degree-D slices starting at c.ix: two words per LDS.64
template <int D>
__device__ __forceinline__ void load_idx_words(const ColCtx &c, int j, uint32_t (&w)[(D + 1) / 2 + 1])
{
    constexpr int H = (D + 1) / 2;
    const uint32_t base = c.ix + j * H * 128;
#pragma unroll
    for (int u = 0; u < H / 2; ++u) {
        const uint2 p = lds_u64(base + u * 256 + c.lane8);
        w[2 * u] = p.x; w[2 * u + 1] = p.y;
    }
    if constexpr (H & 1) w[H - 1] = lds_u32(base + (H / 2) * 256 + c.lane4);
}

// gather addresses and gathered messages of a group of N slices
template <int D, int N>
struct ColGroup {
    uint32_t addr[N][D + 1];
    float r[N][D + 1];
};

template <int D, int N>
__device__ __forceinline__ void group_load_idx(const ColCtx &c, int g, uint32_t (&w)[N][(D + 1) / 2 + 1])
{
#pragma unroll
    for (int j = 0; j < N; ++j) load_idx_words<D>(c, g * N + j, w[j]);
}

template <int D, int N, bool SIGW>
__device__ __forceinline__ void group_gather(const uint32_t (&w)[N][(D + 1) / 2 + 1], ColGroup<D, N> &G, uint32_t win)
{
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (SIGW && (D & 1) && k == D - 1)       // tagged last word of an odd degree: byte address | fingerprint << 24
                G.addr[j][k] = w[j][k >> 1] & 0x3FFFCu;
            else
                G.addr[j][k] = ((k & 1) ? ((w[j][k >> 1] >> 14) & 0x3FFFCu) : ((w[j][k >> 1] << 2) & 0x3FFFCu)) | win;
            G.r[j][k] = lds_f32(G.addr[j][k]);
        }
}

// E holds R on entry and the unclipped v - R on exit (the clamp is applied per row in phase A)
// SIGW: the layout keeps the fingerprint of an odd-degree variable in the free upper half of its last index word
// (edge_layout.h, EDGE_SIG_TAG), which saves the byte load; c.fp then carries tag bits above bit 7 that the caller masks
template <int D, bool EXACT, bool WRITE_V, int N, bool SIGW>
__device__ __forceinline__ void group_finish(ColCtx &c, const ColGroup<D, N> &G, const EdgePriors &pri, const uint32_t (&w)[N][(D + 1) / 2 + 1])
{
    float v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float acc = D > 0 ? G.r[j][0] : 0.f;               // kernels.py:316 (row order)
#pragma unroll
        for (int k = 1; k < D; ++k) acc += G.r[j][k];
        v[j] = acc + __uint_as_float(*reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(pri.bits) + c.t4 + 4 * j));   // kernels.py:320
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float q = v[j] - G.r[j][k];                    // kernels.py:326
            if constexpr (EXACT) q = (q != q) ? 0.f : q;   // kernels.py:328-329
            sts_f32(G.addr[j][k], q);
        }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const bool neg = v[j] < 0.f;                       // kernels.py:349
        if constexpr (SIGW && (D & 1)) xor_or_if_negative(v[j], c.fpw, w[j][(D + 1) / 2 - 1], c.negbits, c.ubit << j);   // (fingerprint: bits 24-31)
        else { if (neg) c.fp ^= lds_u8(c.sg + 8 * c.t4 + 32 * j); }
        if constexpr (SIGW) {
            if constexpr (!(D & 1)) or_if_negative(v[j], c.negbits, c.ubit << j);
        } else {
            const uint32_t hw = __ballot_sync(0xFFFFFFFFu, neg);
            if (c.lane_t4 == c.t4 + 4 * j) c.myhw = hw;
        }
        if constexpr (WRITE_V) {
            static_assert(N == 1, "posterior output assumes one slice per group");
            const uint32_t vid = c.vid_next;                // loaded one task ahead: the store does not wait on it
            c.vid += 32;
            c.vid_next = __ldg(c.vid);                      // (the table is padded by one slice)
            if (vid != 0xFFFFu) c.post[vid] = v[j];
        }
    }
    c.ix += N * ((D + 1) / 2) * 128; c.t4 += 4 * N;
    if constexpr (SIGW) c.ubit <<= N;
}

// any slice: partial with a negative prior, per-lane priors, large degree (meta = degree << 16 | lanes << 22)
template <bool WRITE_V>
__device__ __forceinline__ void col_task_generic(ColCtx &c, uint32_t meta, const float *lane_prior, const EdgePriors &pri)
{
    const int D = (meta >> 16) & 63, nl = (meta >> 22) & 63, H = (D + 1) >> 1;
    bool neg = false;
    if (c.lane < nl) {
        float acc = 0.f;
        for (int k = 0; k < D; ++k) {
            const uint32_t w = lds_u32(c.ix + edge_idx_off(H, k >> 1, c.lane) * 4);
            acc += lds_f32(((k & 1) ? ((w >> 14) & 0x3FFFCu) : ((w << 2) & 0x3FFFCu)) | c.win);
        }
        const float v = acc + (lane_prior ? __ldg(lane_prior) : __uint_as_float(pri.bits[c.t4 >> 2]));
        for (int k = 0; k < D; ++k) {
            const uint32_t w = lds_u32(c.ix + edge_idx_off(H, k >> 1, c.lane) * 4);
            const uint32_t addr = ((k & 1) ? ((w >> 14) & 0x3FFFCu) : ((w << 2) & 0x3FFFCu)) | c.win;
            const float q = v - lds_f32(addr);
            sts_f32(addr, (q != q) ? 0.f : q);
        }
        neg = v < 0.f;
        if (neg) c.fp ^= lds_u8(c.sg + 8 * c.t4);
        if (WRITE_V) c.post[c.vid_next] = v;
    }
    const uint32_t hw = __ballot_sync(0xFFFFFFFFu, neg);
    if (c.lane_t4 == c.t4) c.myhw = hw;
    if (neg) c.negbits |= c.ubit;
    c.ubit <<= 1;
    if (WRITE_V) { c.vid += 32; c.vid_next = __ldg(c.vid); }
    c.ix += H * 128; c.t4 += 4;
}

// Groups of N consecutive slices of one class (two slices at a time double the independent gathers in flight).
// A software-pipelined version (gathers of the next group issued before the current one is scattered) was measured
// slower on B200: the larger unrolled code costs more in instruction fetch than the extra overlap gains.
#ifndef QB_EDGE_GROUP
#define QB_EDGE_GROUP 1
#endif
template <int D, bool EXACT, bool WRITE_V, bool SIGW>
__device__ __forceinline__ void col_class(ColCtx &c, int cnt, const EdgePriors &pri)
{
    constexpr int N = (D <= 4 && !WRITE_V) ? QB_EDGE_GROUP : 1;
    const uint32_t t4_end = c.t4 + 4u * (uint32_t)(cnt / N * N);
#pragma unroll 1
    while (c.t4 != t4_end) {
        uint32_t w[N][(D + 1) / 2 + 1];
        ColGroup<D, N> G;
        group_load_idx<D, N>(c, 0, w);
        group_gather<D, N, SIGW>(w, G, c.win);
        group_finish<D, EXACT, WRITE_V, N, SIGW>(c, G, pri, w);
    }
    if constexpr (N == 2) {
        if (cnt & 1) {
            uint32_t w[1][(D + 1) / 2 + 1];
            ColGroup<D, 1> G;
            group_load_idx<D, 1>(c, 0, w);
            group_gather<D, 1, SIGW>(w, G, c.win);
            group_finish<D, EXACT, WRITE_V, 1, SIGW>(c, G, pri, w);
        }
    }
}

// The warp's column slices are sorted by class; cls holds the number of slices per class (16 x u8).  (A jump-table
// dispatch over a per-warp class program and 2-slice / software-pipelined groups were all measured slower: this
// phase is sensitive to instruction-fetch stalls, the smallest code wins.)
template <bool WRITE_V, bool SIGW = false>
__device__ __forceinline__ void phase_b(ColCtx &c, uint4 cls, int t_end, const uint32_t *cmeta, const float *lane_prior, const EdgePriors &pri)
{
    col_class<0, false, WRITE_V, SIGW>(c, cls.x & 255, pri);
    col_class<1, false, WRITE_V, SIGW>(c, (cls.x >> 8) & 255, pri);
    col_class<2, false, WRITE_V, SIGW>(c, (cls.x >> 16) & 255, pri);
    col_class<3, false, WRITE_V, SIGW>(c, cls.x >> 24, pri);
    col_class<4, false, WRITE_V, SIGW>(c, cls.y & 255, pri);
    col_class<5, false, WRITE_V, SIGW>(c, (cls.y >> 8) & 255, pri);
    col_class<6, false, WRITE_V, SIGW>(c, (cls.y >> 16) & 255, pri);
    if (c.t4 >= 4u * (uint32_t)t_end) return;              // the rare classes follow: skip their tests (far jumps)
    col_class<7, false, WRITE_V, SIGW>(c, cls.y >> 24, pri);
    col_class<8, false, WRITE_V, SIGW>(c, cls.z & 255, pri);
    col_class<1, true, WRITE_V, SIGW>(c, (cls.z >> 8) & 255, pri);
    col_class<2, true, WRITE_V, SIGW>(c, (cls.z >> 16) & 255, pri);
    col_class<3, true, WRITE_V, SIGW>(c, cls.z >> 24, pri);
    col_class<4, true, WRITE_V, SIGW>(c, cls.w & 255, pri);
    col_class<5, true, WRITE_V, SIGW>(c, (cls.w >> 8) & 255, pri);
    col_class<6, true, WRITE_V, SIGW>(c, (cls.w >> 16) & 255, pri);
    const int ngen = cls.w >> 24;
    for (int i = 0; i < ngen; ++i)
        col_task_generic<WRITE_V>(c, cmeta[c.t4 >> 2], lane_prior ? lane_prior + (c.t4 >> 2) * 32 + c.lane : nullptr, pri);
}


}  // namespace qb
