// K4+K5 (pipeline path): OSD-0 on the sides min-sum did not converge on, as two kernels:
//
//   osd_select_kernel   one CTA per failed side: residual syndrome s = syndrome ^ H.hard (osd.py:7-9) and the first
//                       ~1000 columns in ascending |posterior| order, ties by column index (the stable form of
//                       osd.py:11-12), via a 2048-bin histogram of the float bit patterns + exact ranking inside bins.
//   osd_free_kernel     one WARP per failed side: GF(2) elimination in the reference's column order
//                       (gf2_elimination_packed_core, src/decoding/kernels.py:49-96) restricted to what OSD-0 needs.
//
// Free-row elimination.  The reference sweeps the dense, column-permuted m x n matrix.  For a consistent syndrome
// (every simulated shot) the pivot COLUMNS are the greedy independent set in reliability order whatever pivot rows are
// picked, the solution on them is unique, and all pivots found after s entered their span get e = 0 -- so the sweep
// can stop there (gross code: ~150 pivots instead of rank(H) ~ 930-1000).  What the kernel keeps per side:
//   * rows are renumbered on first touch ("compact rows" x = 0, 1, ...: first the support of s, then the rows a
//     candidate brings in; ~1.1 new rows per pivot, 170 on average for the gross code at p = 0.005);
//   * only the FREE part of the row transform: for every compact row x the vector T.e_x restricted to the rows that
//     are not pivots yet.  A pivot consumes one free row and brings ~1.1 new ones, so at any time only ~residual
//     weight + a few rows are free (gross code: mean 46, 99.5 % of the sides <= 128).  Free rows live in SLOTS that
//     are recycled when their row becomes a pivot, so a vector is FW = 128 * Q bits (one uint4 for Q = 1) whatever
//     m is: the whole elimination state of a side is 16 * Q bytes per touched row (2.7 KB on average instead of the
//     20 KB of the full-width kernel), which is what lets ~20 independent one-warp sides share an SM;
//   * candidate c: v = XOR of the vectors of its <= 8 rows (one shared-memory gather + 3 shuffle steps); it is a
//     pivot iff v != 0; any set slot b serves as the pivot row.  Update: every vector with bit b gets ^= v (this also
//     clears bit b everywhere, so the slot is free again), the transformed syndrome likewise;
//   * pivot parts are never formed.  Instead each pivot records the row of the transform it froze (R_t = the set of
//     compact rows whose vector had bit b: exactly the ballots of the update scan) and sigma_t = bit b of the
//     transformed syndrome.  The frozen rows form an upper-triangular system, solved backwards at the end:
//         e_t = sigma_t ^ parity(R_t & y),   y ^= (compact support of column c_t) when e_t = 1
//     -- equal to s_reduced[pivot_row] of the reference's Gauss-Jordan sweep (osd.py:19-21) bit for bit.
// Sides that do not fit (more than FW free rows, RCAP touched rows, the candidate window or the record buffer) are
// appended to an overflow queue untouched and solved by the full-width kernel of osd.cu (~1 % of the gross-code sides).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace qb {

constexpr int SELF_THREADS = 256;
constexpr int SELF_BINS = 2048;            // 64 per octave over 2^-25 .. 2^7, clamped (monotone in the key)
constexpr int SELF_SHIFT = 17;
constexpr int SELF_BASE = (127 - 25) << 6;
constexpr int FREE_WARPS = 4;              // independent sides per CTA of the elimination kernel (fewer when a side's state is large)

struct OsdFreeArgs {
    GraphDev g;
    OsdLaunch a;
    int F;                       // queue length bound (n_fail_d gives the exact count when set)
    int cap, sel_min;            // candidates materialised per side: window closed at >= sel_min, never above cap
    int rcap;                    // compact rows per side (T capacity)
    int fw_bits;                 // free slots per side = 128 * Q
    int rec_cap;                 // record words per warp slot
    uint16_t *cand;              // [F][cap] column ids, ascending (|posterior|, index)
    int32_t *ncand;              // [F] candidates materialised; -1: hand the side to the full-width kernel
    uint32_t *res;               // [F][mw] residual syndrome
    uint32_t *rec;               // [slots][rec_cap]  frozen transform rows, appended per pivot
    uint32_t *meta;              // [slots][rcap]     column | sigma << 16 | words << 17 per pivot
    int32_t *counters;           // [0] select queue, [1] elimination queue, [2] overflow count
    int32_t *ovf_idx;            // overflow queue (shot indices)
};

__device__ __forceinline__ int self_bin(uint32_t key) { return min(max((int)(key >> SELF_SHIFT) - SELF_BASE, 0), SELF_BINS - 1); }

// ---- selection ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SELF_THREADS) osd_select_kernel(const __grid_constant__ OsdFreeArgs P)
{
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);                 // [SELF_BINS] count; later offset << 16 | count
    uint32_t *listK = hist + SELF_BINS;                                       // [cap]
    uint16_t *listI = reinterpret_cast<uint16_t *>(listK + P.cap);            // [cap]
    uint16_t *ord = listI + P.cap;                                            // [cap]
    uint32_t *sv = reinterpret_cast<uint32_t *>(ord + P.cap);                 // [mw]
    __shared__ int s_q, s_b1, s_b2, s_wsum[SELF_THREADS / 32], s_wt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n, mw = g.mw;
    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.F) : P.F;

    while (true) {
        if (tid == 0) { s_q = atomicAdd(&P.counters[0], 1); s_b1 = SELF_BINS - 1; s_b2 = -1; s_wt = 0; }
        __syncthreads();
        const int q = s_q;
        if (q >= F) break;
        const int shot = P.a.fail_idx ? P.a.fail_idx[q] : q;
        const uint32_t *hard = P.a.hard_bits + (size_t)shot * g.nw;
        const float *post = P.a.post + (size_t)shot * n;
        // ---- residual syndrome (osd.py:7-9) ----
        for (int w = tid; w < mw; w += SELF_THREADS) sv[w] = P.a.syn_bits[(size_t)shot * mw + w];
        for (int b = tid; b < SELF_BINS; b += SELF_THREADS) hist[b] = 0u;
        __syncthreads();
        for (int w = tid; w < g.nw; w += SELF_THREADS) {
            uint32_t bits = hard[w];
            while (bits) {
                const int b = __ffs(bits) - 1; bits &= bits - 1;
                const int j = w * 32 + b;
                if (j < n) {
                    const uint4 sg = g.colsig[j];
                    const uint32_t rr[8] = {sg.x & 0xFFFFu, sg.x >> 16, sg.y & 0xFFFFu, sg.y >> 16, sg.z & 0xFFFFu, sg.z >> 16, sg.w & 0xFFFFu, sg.w >> 16};
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (rr[k] != 0xFFFFu) atomicXor(&sv[rr[k] >> 5], 1u << (rr[k] & 31));
                }
            }
        }
        // ---- histogram of the |posterior| bit patterns ----
        for (int j0 = tid; j0 < n; j0 += 8 * SELF_THREADS) {
            uint32_t kb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + u * SELF_THREADS; kb[u] = j < n ? (uint32_t)self_bin(__float_as_uint(fabsf(post[j]))) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (kb[u] != 0xFFFFFFFFu) atomicAdd(&hist[kb[u]], 1u);
        }
        __syncthreads();
        {   // weight of the residual: more free rows than the elimination kernel has slots -> full-width kernel
            int wt = 0;
            for (int w = tid; w < mw; w += SELF_THREADS) { const uint32_t x = sv[w]; wt += __popc(x); P.res[(size_t)q * mw + w] = x; }
            wt = __reduce_add_sync(0xFFFFFFFFu, wt);
            if (lane == 0 && wt) atomicAdd(&s_wt, wt);
        }
        // ---- window [0, hi]: closed at >= sel_min candidates, never above cap (block-wide scan over the bins) ----
        constexpr int BPT = SELF_BINS / SELF_THREADS;
        uint32_t c[BPT];
        uint32_t local = 0;
#pragma unroll
        for (int i = 0; i < BPT; ++i) { c[i] = hist[tid * BPT + i]; local += c[i]; }
        uint32_t inc = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_wsum[warp] = (int)inc;
        __syncthreads();
        uint32_t run = inc - local;
        for (int w = 0; w < warp; ++w) run += (uint32_t)s_wsum[w];
        int b1 = SELF_BINS - 1, b2 = -1;                       // first bin reaching sel_min / last bin within cap
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int b = tid * BPT + i;
            hist[b] = (min(run, 65535u) << 16) | min(c[i], 65535u);
            run += c[i];
            if ((int)run >= P.sel_min) b1 = min(b1, b);
            if ((int)run <= P.cap) b2 = b;
        }
        b1 = __reduce_min_sync(0xFFFFFFFFu, b1); b2 = __reduce_max_sync(0xFFFFFFFFu, b2);
        if (lane == 0) { atomicMin(&s_b1, b1); atomicMax(&s_b2, b2); }
        __syncthreads();
        const int hi = min(s_b1, s_b2);
        int M = 0;
        if (hi >= 0) { const uint32_t h = hist[hi]; M = (int)(h >> 16) + (int)(h & 0xFFFFu); }
        const bool too_heavy = s_wt > P.fw_bits;
        if (M == 0 || too_heavy) {                            // first bin alone exceeds the window (mass ties), or too many free rows
            if (tid == 0) P.ncand[q] = -1;
            __syncthreads();
            continue;
        }
        // ---- scatter (unordered inside a bin), then rank inside each bin by (key, index) ----
        for (int j0 = tid; j0 < n; j0 += 8 * SELF_THREADS) {
            uint32_t kb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + u * SELF_THREADS; kb[u] = j < n ? __float_as_uint(fabsf(post[j])) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t key = kb[u];
                const int b = self_bin(key);
                if (key != 0xFFFFFFFFu && b <= hi) {
                    const uint32_t old = atomicSub(&hist[b], 1u);
                    const int slot = (int)(old >> 16) + (int)(old & 0xFFFFu) - 1;
                    listK[slot] = key; listI[slot] = (uint16_t)(j0 + u * SELF_THREADS);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < M; i += SELF_THREADS) {
            const uint32_t key = listK[i];
            const uint16_t id = listI[i];
            const int b = self_bin(key);
            const int lo = (int)(hist[b] >> 16);
            const int hi2 = b < hi ? (int)(hist[b + 1] >> 16) : M;
            int rank = lo;
            for (int k = lo; k < hi2; ++k) {
                const uint32_t kk = listK[k];
                rank += (kk < key) || (kk == key && listI[k] < id);
            }
            ord[rank] = id;
        }
        __syncthreads();
        uint16_t *out = P.cand + (size_t)q * P.cap;
        for (int i = tid; i < M; i += SELF_THREADS) out[i] = ord[i];
        if (tid == 0) P.ncand[q] = M;
        __syncthreads();
    }
}

// ---- elimination ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t u4_word(const uint4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// Q = uint4 per vector (128 * Q free slots)
template <int Q>
__global__ void __launch_bounds__(FREE_WARPS * 32) osd_free_kernel(const __grid_constant__ OsdFreeArgs P)
{
    constexpr int WV = 4 * Q;                 // words per vector
    constexpr int LV = WV;                    // lanes per vector in the gather layout (Q = 1, 2 -> 4, 8)
    constexpr int KP = 32 / LV;               // rows gathered per pass
    constexpr int PASSES = 8 / KP;
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_pad16 = (g.m + 7) & ~7;       // rowmap entries, multiple of 8 (16-byte fills)
    const size_t per_warp = ((size_t)m_pad16 * 2 + (size_t)P.rcap * 16 * Q + 512 + 64 + (size_t)(P.rcap / 32) * 4 + 15) & ~(size_t)15;
    unsigned char *base = smem_raw + per_warp * warp;
    uint4 *T4 = reinterpret_cast<uint4 *>(base);                                       // [rcap][Q]
    uint32_t *Tw = reinterpret_cast<uint32_t *>(base);
    uint4 *sigbuf = reinterpret_cast<uint4 *>(base + (size_t)P.rcap * 16 * Q);         // [32] signatures of the current batch
    const uint16_t *sig16 = reinterpret_cast<const uint16_t *>(sigbuf);
    uint16_t *rowmap = reinterpret_cast<uint16_t *>(sigbuf + 32);                      // [m_pad16] original row -> compact row
    uint32_t *ybits = reinterpret_cast<uint32_t *>(rowmap + m_pad16);                  // [rcap / 32]
    const int slot_id = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *rec = P.rec + (size_t)slot_id * P.rec_cap;
    uint32_t *meta = P.meta + (size_t)slot_id * P.rcap;
    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.F) : P.F;
    const int mw = g.mw;
    const int k_of_lane = lane / LV, w_of_lane = lane % LV;

    while (true) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&P.counters[1], 1);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        if (q >= F) break;
        const int shot = P.a.fail_idx ? P.a.fail_idx[q] : q;
        const int M = P.ncand[q];
        bool overflow = M < 0;
        int R = 0, t = 0, off = 0;
        uint32_t used[WV], s[WV];
#pragma unroll
        for (int i = 0; i < WV; ++i) { used[i] = 0u; s[i] = 0u; }
        bool finished = false;
        if (!overflow) {
            // ---- reset the row map, then compact rows / slots 0 .. wt-1 for the support of the residual syndrome ----
            {
                uint4 *rm4 = reinterpret_cast<uint4 *>(rowmap);
                const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                for (int i = lane; i < m_pad16 / 8; i += 32) rm4[i] = ff;
            }
            __syncwarp();
            for (int w0 = 0; w0 < mw; w0 += 32) {
                const int w = w0 + lane;
                uint32_t word = w < mw ? P.res[(size_t)q * mw + w] : 0u;
                const int cnt = __popc(word);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
                int x = R + inc - cnt;
                R += __shfl_sync(0xFFFFFFFFu, inc, 31);
                while (word) {                                   // (the select kernel guarantees wt <= 128 * Q <= rcap)
                    const int b = __ffs(word) - 1; word &= word - 1;
                    rowmap[w * 32 + b] = (uint16_t)x;
#pragma unroll
                    for (int qq = 0; qq < Q; ++qq) {
                        uint4 e = make_uint4(0u, 0u, 0u, 0u);
                        if ((x >> 7) == qq) {
                            const uint32_t bit = 1u << (x & 31);
                            const int wi = (x >> 5) & 3;
                            e.x = wi == 0 ? bit : 0u; e.y = wi == 1 ? bit : 0u; e.z = wi == 2 ? bit : 0u; e.w = wi == 3 ? bit : 0u;
                        }
                        T4[x * Q + qq] = e;
                    }
                    ++x;
                }
            }
#pragma unroll
            for (int i = 0; i < WV; ++i) {
                const int lo = 32 * i;
                const uint32_t msk = R >= lo + 32 ? 0xFFFFFFFFu : (R > lo ? (1u << (R - lo)) - 1u : 0u);
                used[i] = msk; s[i] = msk;
            }
            finished = R == 0;
            __syncwarp();
        }

        // ---- candidates in reliability order, 32 at a time (ids and signatures prefetched one batch ahead) ----
        const uint16_t *cand = P.cand + (size_t)q * P.cap;
        uint32_t idx_cur = 0xFFFFu, idx_nxt = 0xFFFFu;
        uint4 sig_cur = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), sig_nxt = sig_cur;
        if (!overflow && !finished) {
            if (lane < M) { idx_cur = cand[lane]; sig_cur = g.colsig[idx_cur]; }
            if (32 + lane < M) idx_nxt = cand[32 + lane];
        }
        for (int c0 = 0; c0 < M && !overflow && !finished; c0 += 32) {
            sigbuf[lane] = sig_cur;
            const uint32_t idx_batch = idx_cur;
            __syncwarp();
            // prefetch the next batch
            if (c0 + 32 + lane < M) sig_nxt = g.colsig[idx_nxt];
            idx_cur = idx_nxt;
            idx_nxt = (c0 + 64 + lane < M) ? (uint32_t)cand[c0 + 64 + lane] : 0xFFFFu;
            const int cnt = min(32, M - c0);
            for (int i = 0; i < cnt; ++i) {
                // ---- v = XOR of the vectors of the candidate's rows (gather layout: k = row slot, w = word) ----
                uint32_t val = 0u;
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int k = ps * KP + k_of_lane;
                    const uint32_t r = sig16[i * 8 + k];
                    const bool valid = r != 0xFFFFu;
                    uint32_t x = valid ? (uint32_t)rowmap[r] : 0xFFFFu;
                    uint32_t fresh = __ballot_sync(0xFFFFFFFFu, valid && x == 0xFFFFu && w_of_lane == 0);
                    if (fresh) {                                 // rows seen for the first time: next compact row, lowest unused slot
                        while (fresh) {
                            const int src = __ffs(fresh) - 1; fresh &= fresh - 1;
                            const uint32_t rr = __shfl_sync(0xFFFFFFFFu, r, src);
                            int f = -1;
#pragma unroll
                            for (int j = WV - 1; j >= 0; --j) if (used[j] != 0xFFFFFFFFu) f = 32 * j + __ffs(~used[j]) - 1;
                            if (f < 0 || R >= P.rcap) { overflow = true; break; }
#pragma unroll
                            for (int j = 0; j < WV; ++j) if (j == (f >> 5)) used[j] |= 1u << (f & 31);
                            if (lane == 0) rowmap[rr] = (uint16_t)R;
                            if (lane < WV) Tw[R * WV + lane] = lane == (f >> 5) ? 1u << (f & 31) : 0u;
                            ++R;
                        }
                        if (overflow) break;
                        __syncwarp();
                        if (valid && x == 0xFFFFu) x = rowmap[r];
                    }
                    if (valid) val ^= Tw[x * WV + w_of_lane];
                }
                if (overflow) break;
#pragma unroll
                for (int o = LV; o < 32; o <<= 1) val ^= __shfl_xor_sync(0xFFFFFFFFu, val, o);
                const uint32_t nz = __ballot_sync(0xFFFFFFFFu, val != 0u) & ((1u << WV) - 1u);
                if (nz == 0u) continue;                          // dependent on the pivots so far
                // ---- pivot: slot b = lowest set bit of v ----
                const int wsel = __ffs(nz) - 1;
                const uint32_t vword = __shfl_sync(0xFFFFFFFFu, val, wsel);
                const uint32_t bmask = vword & (0u - vword);
                uint4 v4[Q];
#pragma unroll
                for (int qq = 0; qq < Q; ++qq) {
                    v4[qq].x = __shfl_sync(0xFFFFFFFFu, val, 4 * qq); v4[qq].y = __shfl_sync(0xFFFFFFFFu, val, 4 * qq + 1);
                    v4[qq].z = __shfl_sync(0xFFFFFFFFu, val, 4 * qq + 2); v4[qq].w = __shfl_sync(0xFFFFFFFFu, val, 4 * qq + 3);
                }
                const int nwr = (R + 31) >> 5;
                if (off + nwr > P.rec_cap || t >= P.rcap) { overflow = true; break; }
                // every vector with bit b: ^= v (clears bit b: the slot is free again); the ballots are the frozen row
                for (int blk = 0; blk < nwr; ++blk) {
                    const int x = blk * 32 + lane;
                    bool has = false;
                    if (x < R) {
                        if constexpr (Q == 1) {
                            uint4 col = T4[x];
                            has = (u4_word(col, wsel) & bmask) != 0u;
                            if (has) { col.x ^= v4[0].x; col.y ^= v4[0].y; col.z ^= v4[0].z; col.w ^= v4[0].w; T4[x] = col; }
                        } else {
                            has = (Tw[x * WV + wsel] & bmask) != 0u;
                            if (has) {
#pragma unroll
                                for (int qq = 0; qq < Q; ++qq) {
                                    uint4 col = T4[x * Q + qq];
                                    col.x ^= v4[qq].x; col.y ^= v4[qq].y; col.z ^= v4[qq].z; col.w ^= v4[qq].w;
                                    T4[x * Q + qq] = col;
                                }
                            }
                        }
                    }
                    const uint32_t flags = __ballot_sync(0xFFFFFFFFu, has);
                    if (lane == 0) rec[off + blk] = flags;
                }
                uint32_t sw = 0u;
#pragma unroll
                for (int j = 0; j < WV; ++j) if (j == wsel) sw = s[j];
                const uint32_t sigma = (sw & bmask) ? 1u : 0u;
                uint32_t any = 0u;
#pragma unroll
                for (int j = 0; j < WV; ++j) {
                    if (sigma) s[j] ^= u4_word(v4[j >> 2], j & 3);
                    if (j == wsel) used[j] &= ~bmask;
                    any |= s[j];
                }
                const uint32_t col_id = __shfl_sync(0xFFFFFFFFu, idx_batch, i);
                if (lane == 0) meta[t] = col_id | (sigma << 16) | ((uint32_t)nwr << 17);
                off += nwr; ++t;
                __syncwarp();
                if (any == 0u) { finished = true; break; }
            }
            sig_cur = sig_nxt;
            __syncwarp();
        }
        if (!finished) overflow = true;                          // window exhausted (or nothing materialised)
        if (overflow) {
            if (lane == 0) { const int o = atomicAdd(&P.counters[2], 1); P.ovf_idx[o] = shot; }
            continue;
        }
        // ---- back substitution over the frozen rows, last pivot first ----
        for (int w = lane; w < (R + 31) >> 5; w += 32) ybits[w] = 0u;
        __syncwarp();
        uint32_t *hard_rw = P.a.hard_bits + (size_t)shot * g.nw;
        uint32_t mt_nxt = t > 0 ? ldcg_u32(&meta[t - 1]) : 0u;
        for (int tt = t - 1; tt >= 0; --tt) {
            const uint32_t mt = mt_nxt;
            if (tt > 0) mt_nxt = ldcg_u32(&meta[tt - 1]);
            const int nwr = (int)(mt >> 17), col = (int)(mt & 0xFFFFu);
            off -= nwr;
            // signature of the column (needed only when e_t = 1, but loaded ahead of the parity to hide its latency)
            const uint32_t r = lane < 8 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(g.colsig) + (size_t)col * 8 + lane) : 0xFFFFu;
            uint32_t par = 0u;
            for (int w = lane; w < nwr; w += 32) par ^= (uint32_t)__popc(ldcg_u32(&rec[off + w]) & ybits[w]);
            par = __reduce_xor_sync(0xFFFFFFFFu, par) & 1u;
            if ((((mt >> 16) & 1u) ^ par) != 0u) {
                if (lane == 0) atomicXor(&hard_rw[col >> 5], 1u << (col & 31));
                if (r != 0xFFFFu) { const uint32_t x = rowmap[r]; atomicXor(&ybits[x >> 5], 1u << (x & 31)); }
                __syncwarp();
            }
        }
        if (P.a.rank_out && lane == 0) P.a.rank_out[shot] = t | (1 << 16);
        __syncwarp();
    }
}

// ---- launcher ---------------------------------------------------------------------------------------------------
struct FreePlan { int Q, rcap, cap, sel_min, rec_cap, ctas_per_sm, warps; size_t smem_sel, smem_free; };

static bool free_plan(const qb_decoder *dec, FreePlan &pl)
{
    const GraphDev &g = dec->g;
    if (!g.colsig || g.n > 65535 || g.m > 65535 || g.m <= 0 || g.n <= 0) return false;
    if (getenv("QLDPC_B200_OSD_FULLWIDTH")) return false;
    pl.Q = g.m <= 1536 ? 1 : 2;
    pl.rcap = g.m <= 1536 ? 512 : 2048;
    if (const char *e = getenv("QLDPC_B200_OSD_RCAP")) { const int v = atoi(e); if (v >= 128 && v <= 4096) pl.rcap = v & ~31; }
    pl.rcap = std::min(pl.rcap, (g.m + 31) & ~31);
    pl.cap = g.m <= 1536 ? 1024 : 8192;
    if (const char *e = getenv("QLDPC_B200_OSD_CAP")) { const int v = atoi(e); if (v >= 64 && v <= 16384) pl.cap = v & ~31; }
    pl.cap = std::min(pl.cap, (g.n + 31) & ~31);
    pl.sel_min = std::max(32, pl.cap - pl.cap / 8);
    pl.rec_cap = pl.rcap * std::max(4, pl.rcap * 3 / 128);                  // 3/4 of rcap * rcap / 32 (the records are triangular)
    const int m_pad16 = (g.m + 7) & ~7;
    const size_t per_warp = ((size_t)m_pad16 * 2 + (size_t)pl.rcap * 16 * pl.Q + 512 + 64 + (size_t)(pl.rcap / 32) * 4 + 15) & ~(size_t)15;
    const size_t limit = (size_t)dec->max_smem_optin;
    pl.warps = (int)std::min<size_t>(FREE_WARPS, (limit - 1024) / per_warp);
    if (pl.warps < 1) return false;
    pl.smem_free = per_warp * pl.warps;
    pl.smem_sel = sizeof(uint32_t) * SELF_BINS + (size_t)pl.cap * (4 + 2 + 2) + sizeof(uint32_t) * (size_t)g.mw + 16;
    if (pl.smem_free + 1024 > limit || pl.smem_sel + 1024 > limit) return false;
    pl.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (limit + 1024) / (pl.smem_free + 1024)));
    return true;
}

bool osd_free_applicable(const qb_decoder *dec) { FreePlan pl; return free_plan(dec, pl); }

// Selection + free-row elimination for the queue of `a`; sides that do not fit end in the overflow queue
// (*ovf_count_d / *ovf_idx_d, device pointers valid until the decoder's next OSD launch).
int launch_osd0_free(qb_decoder *dec, const OsdLaunch &a, int32_t **ovf_count_d, int32_t **ovf_idx_d, cudaStream_t st)
{
    FreePlan pl;
    if (!free_plan(dec, pl)) { set_error("free-row OSD kernel not applicable"); return QB_ERR_UNSUPPORTED; }
    const GraphDev &g = dec->g;
    OsdFreeArgs P{};
    P.g = g; P.a = a; P.F = a.F;
    P.cap = pl.cap; P.sel_min = pl.sel_min; P.rcap = pl.rcap; P.fw_bits = 128 * pl.Q; P.rec_cap = pl.rec_cap;
    const int grid_free = std::max(1, std::min(ceil_div(a.F, pl.warps), dec->sm_count * pl.ctas_per_sm));
    const int sel_ctas = (int)std::max<size_t>(1, std::min<size_t>(2048 / SELF_THREADS, ((size_t)dec->max_smem_optin + 1024) / (pl.smem_sel + 1024)));
    const int grid_sel = std::max(1, std::min(a.F, dec->sm_count * sel_ctas));
    const size_t slots = (size_t)grid_free * pl.warps;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t F = (size_t)a.F;
    const size_t need = 256 + al(F * pl.cap * 2) + al(F * 4) + al(F * g.mw * 4) + al(slots * pl.rec_cap * 4) + al(slots * pl.rcap * 4) + al(F * 4);
    if (int rc = dec->ovf.ensure(need)) return rc;
    unsigned char *p = dec->ovf.as<unsigned char>();
    P.counters = reinterpret_cast<int32_t *>(p); p += 256;
    P.cand = reinterpret_cast<uint16_t *>(p); p += al(F * pl.cap * 2);
    P.ncand = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    P.res = reinterpret_cast<uint32_t *>(p); p += al(F * g.mw * 4);
    P.rec = reinterpret_cast<uint32_t *>(p); p += al(slots * pl.rec_cap * 4);
    P.meta = reinterpret_cast<uint32_t *>(p); p += al(slots * pl.rcap * 4);
    P.ovf_idx = reinterpret_cast<int32_t *>(p);
    QB_CUDA(cudaMemsetAsync(P.counters, 0, 16, st));
    QB_CUDA(cudaFuncSetAttribute(osd_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_sel));
    osd_select_kernel<<<grid_sel, SELF_THREADS, pl.smem_sel, st>>>(P);
    QB_CUDA(cudaGetLastError());
    if (pl.Q == 1) {
        QB_CUDA(cudaFuncSetAttribute(osd_free_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_free));
        osd_free_kernel<1><<<grid_free, pl.warps * 32, pl.smem_free, st>>>(P);
    } else {
        QB_CUDA(cudaFuncSetAttribute(osd_free_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_free));
        osd_free_kernel<2><<<grid_free, pl.warps * 32, pl.smem_free, st>>>(P);
    }
    QB_CUDA(cudaGetLastError());
    *ovf_count_d = P.counters + 2;
    *ovf_idx_d = P.ovf_idx;
    return QB_OK;
}

}  // namespace qb
