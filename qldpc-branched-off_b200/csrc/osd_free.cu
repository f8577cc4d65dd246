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
constexpr int FREE_WARPS_MAX = 8;          // upper bound of sides per CTA (launch bound)
constexpr int FREE_WARPS = 4;              // independent sides per CTA of the elimination kernel (fewer when a side's state is large)

// counters[]: 0 select queue, 1 tier-A queue, 2 tier-B queue, 3 tier-A overflow count (= tier-B input), 4 tier-B overflow
// count (= input of the full-width kernel), 5 second-selection queue, 6 tier-A window exits (= its input), 8.. statistics: 8 window exhausted, 9 touched rows, 10 free slots, 11 records,
// 12 not materialised
enum { CNT_SEL = 0, CNT_A = 1, CNT_B = 2, CNT_OVF_A = 3, CNT_OVF_B = 4, CNT_SEL2 = 5, CNT_WIN_A = 6, CNT_STAT = 8, CNT_STAT_B = 16, CNT_WORDS = 24 };

struct OsdFreeArgs {
    GraphDev g;
    OsdLaunch a;
    int F;                       // queue length bound (n_fail_d gives the exact count when set)
    int cap;                     // candidates materialised per side at most (= stride of cand)
    int cap_per_wt;              // a side of residual weight wt gets cap_per_wt * wt candidates (256 .. cap, multiple of 64)
    int max_wt;                  // heavier residuals are not materialised (more free rows than any tier has slots)
    uint16_t *cand;              // [F][cap] column ids, ascending (|posterior|, index)
    int32_t *ncand;              // [F] candidates materialised; -1: hand the side to the full-width kernel
    uint32_t *res;               // [F][mw] residual syndrome
    int32_t *counters;           // [CNT_WORDS]
    // second selection pass (sides whose first window was exhausted in tier A, or that could not be cut into a window):
    // the positions listed in sel_q get ALL columns in order; entry i of the list owns cand[i][cap] / ncand[i] of this
    // pass's buffers and win_slot[position] = i tells tier B where to read
    const int32_t *sel_count;    // nullptr: first pass over the whole failure queue
    const int32_t *sel_q;
    int sel_counter;             // index into counters[] of the pass's work counter
    int sel_max;                 // entries of sel_q the second pass has buffers for
    int32_t *win_slot;           // [F] -1, or the side's entry in the second pass's buffers
    int32_t *win_list;           // positions wanting the second pass (appended by pass 1 and by tier A; counters[CNT_WIN_A])
    const uint16_t *cand2; const int32_t *ncand2; int cap2;     // second-pass buffers as seen by the elimination tiers
};

// one tier of the elimination kernel: sizes of a side's state and where its sides come from / overflow to
struct OsdFreeTier {
    int rcap;                    // compact rows per side (T capacity)
    int rec_cap;                 // record words per warp slot
    uint32_t *rec;               // [slots][rec_cap]  frozen transform rows, appended per pivot
    uint32_t *meta;              // [slots][rcap]     column | sigma << 16 | words << 17 per pivot
    int queue_counter;           // index into counters[] of this tier's work counter
    const int32_t *in_count;     // nullptr: the failure queue itself (positions 0 .. F-1); else *in_count entries of in_q
    const int32_t *in_q;         // queue positions handed over by the previous tier
    int out_counter;             // index into counters[] of this tier's overflow count
    int32_t *out_list;           // overflow: queue positions (out_shots == 0) or shot indices (out_shots == 1)
    int out_shots;
    int32_t *win_list;           // nullable: positions whose window was exhausted are also listed here (counters[CNT_WIN_A])
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
    uint32_t *sv = reinterpret_cast<uint32_t *>(listI + P.cap);               // [mw]
    __shared__ int s_q, s_b1, s_b2, s_wsum[SELF_THREADS / 32], s_wt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n, mw = g.mw;
    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.F) : P.F;
    const int n_in = P.sel_count ? min(min(*P.sel_count, F), P.sel_max) : F;
    const bool second = P.sel_q != nullptr;
    const bool packed_ok = n <= (1 << 14);                  // column ids fit the 14 bits beside the 17 low key bits

    while (true) {
        if (tid == 0) { s_q = atomicAdd(&P.counters[P.sel_counter], 1); s_b1 = SELF_BINS - 1; s_b2 = -1; s_wt = 0; }
        __syncthreads();
        const int qi = s_q;
        if (qi >= n_in) break;
#ifdef QB_OSD_STATS
        const long long sp0 = clock64();
#endif
        const int q = P.sel_q ? P.sel_q[qi] : qi;
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
#ifdef QB_OSD_STATS
        const long long sp1 = clock64();
#endif
        {   // weight of the residual: sizes the window (the candidates an elimination examines grow with it)
            int wt = 0;
            for (int w = tid; w < mw; w += SELF_THREADS) { const uint32_t x = sv[w]; wt += __popc(x); P.res[(size_t)q * mw + w] = x; }
            wt = __reduce_add_sync(0xFFFFFFFFu, wt);
            if (lane == 0 && wt) atomicAdd(&s_wt, wt);
        }
        // ---- window [0, hi]: closed at >= sel_min candidates, never above the side's cap (block-wide scan over the bins) ----
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
        const int wt_side = s_wt;
        // first pass: cap_per_wt candidates per unit of residual weight (measured on the gross code: sides of weight
        // 20-40 examine <= 592 candidates in 99 % of the cases, 60-80: <= 1061, 100+: <= 3015), closed 64 below the cap
        int cap_side = min(P.cap, max(256, (P.cap_per_wt * wt_side + 63) & ~63));
        if (second) cap_side = P.cap;
        const int sel_min = second ? cap_side : cap_side - 64;                  // second pass: every column
        uint32_t run = inc - local;
        for (int w = 0; w < warp; ++w) run += (uint32_t)s_wsum[w];
        int b1 = SELF_BINS - 1, b2 = -1;                       // first bin reaching sel_min / last bin within the cap
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int b = tid * BPT + i;
            hist[b] = (min(run, 65535u) << 16) | min(c[i], 65535u);
            run += c[i];
            if ((int)run >= sel_min) b1 = min(b1, b);
            if ((int)run <= cap_side) b2 = b;
        }
        b1 = __reduce_min_sync(0xFFFFFFFFu, b1); b2 = __reduce_max_sync(0xFFFFFFFFu, b2);
        if (lane == 0) { atomicMin(&s_b1, b1); atomicMax(&s_b2, b2); }
        __syncthreads();
        const int hi = min(s_b1, s_b2);
        int M = 0;
        if (hi >= 0) { const uint32_t h = hist[hi]; M = (int)(h >> 16) + (int)(h & 0xFFFFu); }
        const int oi = second ? qi : q;                       // entry of this pass's output buffers
        if (M == 0 || wt_side > P.max_wt) {                   // first bin alone exceeds the window (mass ties), or too many free rows
            if (tid == 0) {
                P.ncand[oi] = -1;
                if (M == 0 && !second && wt_side <= P.max_wt) P.win_list[atomicAdd(&P.counters[CNT_WIN_A], 1)] = q;   // full order in pass 2
            }
            __syncthreads();
            continue;
        }
#ifdef QB_OSD_STATS
        const long long sp2 = clock64();
#endif
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
                    const uint32_t j = (uint32_t)(j0 + u * SELF_THREADS);
                    // inside an unclamped bin the upper 15 bits of the keys agree: (low 17 key bits, 14-bit column) is one
                    // 31-bit word whose order is the (key, column) order, and the bin number takes the column's place
                    if (packed_ok && b > 0 && b < SELF_BINS - 1) { listK[slot] = ((key & ((1u << SELF_SHIFT) - 1u)) << 14) | j; listI[slot] = (uint16_t)b; }
                    else { listK[slot] = key; listI[slot] = (uint16_t)j; }
                }
            }
        }
        __syncthreads();
#ifdef QB_OSD_STATS
        const long long sp3 = clock64();
#endif
        uint16_t *out = P.cand + (size_t)oi * P.cap;
        // the clamped bins sit at the two ends of the list: [0, p_lo) is bin 0, [p_hi, M) the last bin; packed words between
        const int p_lo = packed_ok ? (hi >= 1 ? (int)(hist[1] >> 16) : M) : M;
        const int p_hi = hi == SELF_BINS - 1 ? (int)(hist[SELF_BINS - 1] >> 16) : M;
        for (int i = tid; i < M; i += SELF_THREADS) {
            if (i >= p_lo && i < p_hi) {
                const uint32_t mine = listK[i];
                const int b = listI[i];
                const int lo = (int)(hist[b] >> 16);
                const int hi2 = b < hi ? (int)(hist[b + 1] >> 16) : M;
                int rank = lo;
                for (int k = lo; k < hi2; ++k) rank += listK[k] < mine;
                out[rank] = (uint16_t)(mine & 0x3FFFu);
                continue;
            }
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
            out[rank] = id;
        }
        if (tid == 0) { P.ncand[oi] = M; if (second) P.win_slot[q] = qi; }
        __syncthreads();
#ifdef QB_OSD_STATS
        if (tid == 0) {
            const long long sp4 = clock64();
            int *c = &P.counters[second ? 58 : 52];
            atomicAdd(&c[0], 1); atomicAdd(&c[1], (int)((sp1 - sp0) >> 8)); atomicAdd(&c[2], (int)((sp2 - sp1) >> 8));
            atomicAdd(&c[3], (int)((sp3 - sp2) >> 8)); atomicAdd(&c[4], (int)((sp4 - sp3) >> 8)); atomicAdd(&c[5], M);
        }
#endif
    }
}

// ---- elimination ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

__host__ __device__ inline size_t free_per_warp_bytes(int m, int rcap, int Q)
{
    const int m_pad16 = (m + 7) & ~7;
    return ((size_t)rcap * 16 * Q + 512 + (size_t)m_pad16 * 2 + (size_t)((rcap + 31) / 32) * 4 + 64 + 15) & ~(size_t)15;
}

// Q = uint4 per vector (128 * Q free slots).  Lane roles: in the gather layout lane = (row slot k, word w) of the candidate's
// signature; lanes 0 .. 4Q-1 additionally own word `lane` of the transformed syndrome (sl) and of the slot-in-use mask (ul).
// (Q = 3: 384 slots, 12 words per vector in a 16-lane group)
template <int Q>
__global__ void __launch_bounds__(FREE_WARPS_MAX * 32) osd_free_kernel(const __grid_constant__ OsdFreeArgs P, const __grid_constant__ OsdFreeTier Tr)
{
    constexpr int WV = 4 * Q;                 // words per vector
    constexpr int LV = Q == 1 ? 4 : (Q == 2 ? 8 : 16);   // lanes per vector in the gather layout (power of two >= WV)
    constexpr int KP = 32 / LV;               // rows gathered per pass
    constexpr int PASSES = 8 / KP;
    constexpr uint32_t WVMASK = (1u << WV) - 1u;
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_pad16 = (g.m + 7) & ~7;       // rowmap entries, multiple of 8 (16-byte fills)
    unsigned char *base = smem_raw + free_per_warp_bytes(g.m, Tr.rcap, Q) * warp;
    uint4 *T4 = reinterpret_cast<uint4 *>(base);                                       // [rcap][Q]
    uint32_t *Tw = reinterpret_cast<uint32_t *>(base);
    uint4 *sigbuf = reinterpret_cast<uint4 *>(base + (size_t)Tr.rcap * 16 * Q);        // [32] signatures of the current batch
    const uint16_t *sig16 = reinterpret_cast<const uint16_t *>(sigbuf);
    uint16_t *rowmap = reinterpret_cast<uint16_t *>(sigbuf + 32);                      // [m_pad16] original row -> compact row
    uint32_t *ybits = reinterpret_cast<uint32_t *>(rowmap + m_pad16);                  // [ceil(rcap / 32)]
    const int slot_id = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *rec = Tr.rec + (size_t)slot_id * Tr.rec_cap;
    uint32_t *meta = Tr.meta + (size_t)slot_id * Tr.rcap;
    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.F) : P.F;
    const int n_in = Tr.in_count ? min(*Tr.in_count, F) : F;
    const int mw = g.mw;
    const int k_of_lane = lane / LV, w_of_lane = lane % LV;

    while (true) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(&P.counters[Tr.queue_counter], 1);
        qi = __shfl_sync(0xFFFFFFFFu, qi, 0);
        if (qi >= n_in) break;
        const int q = Tr.in_q ? Tr.in_q[qi] : qi;
        const int shot = P.a.fail_idx ? P.a.fail_idx[q] : q;
        const int ws = P.win_slot[q];                            // >= 0: the second selection pass re-listed this side
        const int M = ws >= 0 ? P.ncand2[ws] : P.ncand[q];
        int why = M < 0 ? 4 : -1;                                // >= 0: overflow, CNT_STAT + why counts the reason
        int R = 0, t = 0, off = 0;
#ifdef QB_OSD_STATS
        int st_blocks = 0, st_hitblocks = 0, st_rows = 0, st_cands = 0;
        const long long st_t0 = clock64();
#endif
        uint32_t ul = 0u, sl = 0u;                               // lanes < WV: slots in use / transformed syndrome, word `lane`
        // lane l: OR of the vectors of compact rows 32 l .. 32 l + 31 (a superset: bits cancelled by an update stay until
        // their slot is the pivot).  Measured on the gross code: a pivot's slot is present in 1.5 of the 5.4 blocks of 32
        // rows a side has touched, so the update sweep visits only the blocks whose summary has the bit.
        uint32_t sum[WV];
#pragma unroll
        for (int j = 0; j < WV; ++j) sum[j] = 0u;
        bool finished = false;
        if (why < 0) {
            // ---- reset the row map, then compact rows / slots 0 .. wt-1 for the support of the residual syndrome ----
            {
                uint4 *rm4 = reinterpret_cast<uint4 *>(rowmap);
                const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                for (int i = lane; i < m_pad16 / 8; i += 32) rm4[i] = ff;
            }
            __syncwarp();
            for (int w0 = 0; w0 < mw && why < 0; w0 += 32) {
                const int w = w0 + lane;
                uint32_t word = w < mw ? P.res[(size_t)q * mw + w] : 0u;
                const int cnt = __popc(word);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
                int x = R + inc - cnt;
                R += __shfl_sync(0xFFFFFFFFu, inc, 31);
                if (R > 128 * Q || R > Tr.rcap) { why = 2; break; }
                while (word) {
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
            {
                const int lo = 32 * lane;
                const uint32_t msk = R >= lo + 32 ? 0xFFFFFFFFu : (R > lo ? (1u << (R - lo)) - 1u : 0u);
                ul = lane < WV ? msk : 0xFFFFFFFFu; sl = lane < WV ? msk : 0u;
                // rows 0 .. R-1 start as unit vectors on slots 0 .. R-1: block `lane` uses exactly word `lane`
#pragma unroll
                for (int j = 0; j < WV; ++j) sum[j] = j == lane ? msk : 0u;
            }
            finished = R == 0;
            __syncwarp();
        }

#ifdef QB_OSD_STATS
        const long long st_t1 = clock64();
#endif
        // ---- candidates in reliability order, 32 at a time (ids and signatures prefetched one batch ahead) ----
        const uint16_t *cand = ws >= 0 ? P.cand2 + (size_t)ws * P.cap2 : P.cand + (size_t)q * P.cap;
        uint32_t idx_cur = 0xFFFFu, idx_nxt = 0xFFFFu;
        uint4 sig_cur = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), sig_nxt = sig_cur;
        if (why < 0 && !finished) {
            if (lane < M) { idx_cur = cand[lane]; sig_cur = g.colsig[idx_cur]; }
            if (32 + lane < M) idx_nxt = cand[32 + lane];
        }
        for (int c0 = 0; c0 < M && why < 0 && !finished; c0 += 32) {
            sigbuf[lane] = sig_cur;
            const uint32_t idx_batch = idx_cur;
            __syncwarp();
            if (c0 + 32 + lane < M) sig_nxt = g.colsig[idx_nxt];             // prefetch the next batch
            idx_cur = idx_nxt;
            idx_nxt = (c0 + 64 + lane < M) ? (uint32_t)cand[c0 + 64 + lane] : 0xFFFFu;
            const int cnt = min(32, M - c0);
            const uint16_t *sigp = sig16 + k_of_lane;
#pragma unroll 1
            for (int i = 0; i < cnt; ++i, sigp += 8) {
#ifdef QB_OSD_STATS
                ++st_cands;
#endif
                // ---- v = XOR of the vectors of the candidate's rows (gather layout: k = row slot, w = word) ----
                // A row seen for the first time would get the next compact row and a unit vector on a new slot.  The FIRST
                // such row of a candidate needs no slot: the candidate is then certainly a pivot (nothing else has that
                // bit), the new row itself can serve as the pivot row, and eliminating with it changes no other vector --
                // the new compact row simply starts with v (the XOR of the candidate's other rows) and is the frozen row.
                uint32_t val = 0u, vpr = 0xFFFFFFFFu;
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const uint32_t r = sigp[ps * KP];
                    const bool valid = r != 0xFFFFu;
                    uint32_t x = valid ? (uint32_t)rowmap[r] : 0xFFFFu;
                    uint32_t fresh = __ballot_sync(0xFFFFFFFFu, valid && x == 0xFFFFu && w_of_lane == 0);
                    if (fresh) {
                        if (vpr == 0xFFFFFFFFu) { const int src = __ffs(fresh) - 1; fresh &= fresh - 1; vpr = __shfl_sync(0xFFFFFFFFu, r, src); }
                        while (fresh) {                          // further new rows: next compact row, lowest unused slot
                            const int src = __ffs(fresh) - 1; fresh &= fresh - 1;
                            const uint32_t rr = __shfl_sync(0xFFFFFFFFu, r, src);
                            const uint32_t um = __ballot_sync(0xFFFFFFFFu, ul != 0xFFFFFFFFu);
                            if (um == 0u) { why = 2; break; }
                            if (R >= Tr.rcap) { why = 1; break; }
                            const int wj = __ffs(um) - 1;
                            const uint32_t uw = __shfl_sync(0xFFFFFFFFu, ul, wj);
                            const uint32_t fbit = ~uw & (uw + 1u);               // lowest zero bit
                            if (lane == wj) ul |= fbit;
                            if (lane == 0) rowmap[rr] = (uint16_t)R;
                            if (lane < WV) Tw[R * WV + lane] = lane == wj ? fbit : 0u;
#pragma unroll
                            for (int j = 0; j < WV; ++j) if (j == wj && lane == (R >> 5)) sum[j] |= fbit;
                            ++R;
                        }
                        if (why >= 0) break;
                        __syncwarp();
                        if (valid && x == 0xFFFFu) x = rowmap[r];                 // (stays 0xFFFF for the pivot row)
                    }
                    if (x != 0xFFFFu && w_of_lane < WV) val ^= Tw[x * WV + w_of_lane];
                }
                if (why >= 0) break;
#pragma unroll
                for (int o = LV; o < 32; o <<= 1) val ^= __shfl_xor_sync(0xFFFFFFFFu, val, o);
                if (vpr != 0xFFFFFFFFu) {
                    if (R >= Tr.rcap) { why = 1; break; }
                    if (off + 1 > Tr.rec_cap || t >= Tr.rcap) { why = 3; break; }
                    if (lane < WV) Tw[R * WV + lane] = val;
#pragma unroll
                    for (int j = 0; j < WV; ++j) { const uint32_t wjv = __shfl_sync(0xFFFFFFFFu, val, j); if (lane == (R >> 5)) sum[j] |= wjv; }
                    const uint32_t col_id = __shfl_sync(0xFFFFFFFFu, idx_batch, i);
                    if (lane == 0) { rowmap[vpr] = (uint16_t)R; rec[off] = (uint32_t)R; meta[t] = col_id; }   // unit record: sigma = 0, word count 0
                    off += 1; ++t; ++R;
                    __syncwarp();
                    continue;
                }
                const uint32_t nz = __ballot_sync(0xFFFFFFFFu, val != 0u) & WVMASK;
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
                if (off + nwr > Tr.rec_cap || t >= Tr.rcap) { why = 3; break; }
                // every vector with bit b: ^= v (clears bit b: the slot is free again); the ballots are the frozen row.
                // Branch-free: every lane loads, masks and stores its row (rows >= R lie inside the T buffer and are dead).
                uint32_t myflags = 0u;
                uint32_t bm[WV];
#pragma unroll
                for (int j = 0; j < WV; ++j) bm[j] = j == wsel ? bmask : 0u;
                uint4 *tp = T4 + (size_t)lane * Q;
                const int nblk = min(nwr, Tr.rcap >> 5);
                auto scan_block = [&](int blk) -> uint32_t {
                    uint4 col[Q];
                    uint32_t hit = 0u;
#pragma unroll
                    for (int qq = 0; qq < Q; ++qq) {
                        col[qq] = tp[qq];
                        hit |= (col[qq].x & bm[4 * qq]) | (col[qq].y & bm[4 * qq + 1]) | (col[qq].z & bm[4 * qq + 2]) | (col[qq].w & bm[4 * qq + 3]);
                    }
                    const bool has = hit != 0u && blk * 32 + lane < R;
                    const uint32_t mk = has ? 0xFFFFFFFFu : 0u;
#pragma unroll
                    for (int qq = 0; qq < Q; ++qq) {
                        col[qq].x ^= v4[qq].x & mk; col[qq].y ^= v4[qq].y & mk; col[qq].z ^= v4[qq].z & mk; col[qq].w ^= v4[qq].w & mk;
                        tp[qq] = col[qq];
                    }
                    tp += 32 * Q;
                    return __ballot_sync(0xFFFFFFFFu, has);
                };
                const int nb1 = min(nblk, 32);
                uint32_t sw = 0u;
#pragma unroll
                for (int j = 0; j < WV; ++j) sw = j == wsel ? sum[j] : sw;
                uint32_t need = __ballot_sync(0xFFFFFFFFu, (sw & bmask) != 0u && lane < nb1);
                while (need) {
                    const int blk = __ffs(need) - 1; need &= need - 1;
                    tp = T4 + (size_t)(blk * 32 + lane) * Q;
                    const uint32_t flags = scan_block(blk);
                    if (blk == lane) {
                        myflags = flags;
                        if (flags) {
#pragma unroll
                            for (int qq = 0; qq < Q; ++qq) { sum[4 * qq] |= v4[qq].x; sum[4 * qq + 1] |= v4[qq].y; sum[4 * qq + 2] |= v4[qq].z; sum[4 * qq + 3] |= v4[qq].w; }
                        }
                    }
#ifdef QB_OSD_STATS
                    ++st_blocks; st_hitblocks += flags != 0u; st_rows += __popc(flags);
#endif
                }
#pragma unroll
                for (int j = 0; j < WV; ++j) sum[j] &= ~bm[j];                   // bit b is clear in every row now
                tp = T4 + (size_t)(32 * 32 + lane) * Q;
                for (int blk = 32; blk < nblk; ++blk) {                      // (more than 1024 touched rows: always visited)
                    const uint32_t flags = scan_block(blk);
                    if (lane == 0) rec[off + blk] = flags;
                }
                if (lane < nwr) rec[off + lane] = myflags;
                const uint32_t sigma = __ballot_sync(0xFFFFFFFFu, lane == wsel && (sl & bmask) != 0u) ? 1u : 0u;
                if (sigma && lane < WV) sl ^= val;
                if (lane == wsel) ul &= ~bmask;
                const uint32_t col_id = __shfl_sync(0xFFFFFFFFu, idx_batch, i);
                if (lane == 0) meta[t] = col_id | (sigma << 16) | ((uint32_t)nwr << 17);
                off += nwr; ++t;
                __syncwarp();
                if (__ballot_sync(0xFFFFFFFFu, lane < WV && sl != 0u) == 0u) { finished = true; break; }
            }
            sig_cur = sig_nxt;
            __syncwarp();
        }
        if (why < 0 && !finished) why = 0;                       // window exhausted
        if (why >= 0) {
            if (lane == 0) {
                const int o = atomicAdd(&P.counters[Tr.out_counter], 1);
                Tr.out_list[o] = Tr.out_shots ? shot : q;
                // whatever made the side leave this tier, it is a heavy one: the next tier gets all columns in order
                if (why != 4 && Tr.win_list && ws < 0) Tr.win_list[atomicAdd(&P.counters[CNT_WIN_A], 1)] = q;
                atomicAdd(&P.counters[(Tr.out_shots ? CNT_STAT_B : CNT_STAT) + why], 1);
            }
            continue;
        }
#ifdef QB_OSD_STATS
        const long long st_t2 = clock64();
#endif
        // ---- back substitution over the frozen rows, last pivot first ----
        // 32 pivots at a time: their meta words, records and column signatures are fetched from L2 with three rounds of
        // coalesced loads (the dependent chain e_t -> y -> e_(t-1) then runs from shared memory: one L2 round trip per pivot
        // was 30 % of a side's time).  The records are staged in the T buffer, which is dead by now; a chunk holds at most
        // 32 * rcap / 32 = rcap words, T has 4 Q rcap.
        for (int w = lane; w < (R + 31) >> 5; w += 32) ybits[w] = 0u;
        uint32_t *hard_rw = P.a.hard_bits + (size_t)shot * g.nw;
        uint32_t *stage = Tw;
        int off_hi = off;
        for (int hi_t = t; hi_t > 0; hi_t -= 32) {
            const int nch = min(32, hi_t);
            const uint32_t mt_l = lane < nch ? ldcg_u32(&meta[hi_t - 1 - lane]) : 0u;          // lane j: pivot hi_t - 1 - j
            // (a pivot made by a new row has a unit record: word count 0 in meta, one record word = the compact row)
            const int cnt_l = lane < nch ? max((int)(mt_l >> 17), 1) : 0;
            int inc = cnt_l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
            const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            const int off_lo = off_hi - total;
            const int start_l = total - inc;                                                     // offset of the lane's record inside the chunk
            __syncwarp();
            for (int w = lane; w < total; w += 32) stage[w] = ldcg_u32(&rec[off_lo + w]);
            sigbuf[lane] = lane < nch ? g.colsig[mt_l & 0xFFFFu] : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            __syncwarp();
            // Lane j evaluates e_j = sigma_j ^ parity(R_j & y) of ITS pivot against the current y; the first pivot (in solve
            // order) with e = 1 changes y, everything before it is final with e = 0, everything after it is evaluated
            // again.  Rounds = solution bits in the chunk + 1 (about a quarter of the pivots) instead of one dependent
            // step per pivot.
            const int nwr_l = (int)(mt_l >> 17);
            int p0 = 0;                                                                          // pivots before p0 are settled
            while (p0 < nch) {
                uint32_t par = 0u;
                if (lane >= p0 && lane < nch) {
                    if (nwr_l == 0) {
                        const uint32_t x = stage[start_l];
                        par = (ybits[x >> 5] >> (x & 31)) & 1u;
                    } else {
                        for (int w = 0; w < nwr_l; ++w) par ^= (uint32_t)__popc(stage[start_l + w] & ybits[w]);
                        par &= 1u;
                    }
                    par ^= (mt_l >> 16) & 1u;
                }
                const uint32_t ones = __ballot_sync(0xFFFFFFFFu, par != 0u);
                if (ones == 0u) break;
                const int j = __ffs(ones) - 1;
                const uint32_t col = __shfl_sync(0xFFFFFFFFu, mt_l, j) & 0xFFFFu;
                if (lane == 0) atomicXor(&hard_rw[col >> 5], 1u << (col & 31));
                const uint32_t r = lane < 8 ? (uint32_t)sig16[j * 8 + lane] : 0xFFFFu;
                if (r != 0xFFFFu) { const uint32_t x = rowmap[r]; atomicXor(&ybits[x >> 5], 1u << (x & 31)); }
                __syncwarp();
                p0 = j + 1;
            }
            off_hi = off_lo;
        }
        if (P.a.rank_out && lane == 0) P.a.rank_out[shot] = t | (1 << 16);
#ifdef QB_OSD_STATS
        if (lane == 0) {
            atomicAdd(&P.counters[32], 1); atomicAdd(&P.counters[33], t); atomicAdd(&P.counters[34], st_blocks);
            atomicAdd(&P.counters[35], st_hitblocks); atomicAdd(&P.counters[36], st_rows); atomicAdd(&P.counters[37], st_cands);
            atomicAdd(&P.counters[38], R);
            // slowest side of the tier: cycles (in units of 256), its pivots / candidates / rows
            const int cyc = (int)((clock64() - st_t0) >> 8);
            int *mx = &P.counters[Tr.out_shots ? 44 : 40];
            if (atomicMax(&mx[0], cyc) < cyc) { mx[1] = t; mx[2] = st_cands; mx[3] = R; }
            atomicAdd(&P.counters[Tr.out_shots ? 49 : 48], 1);
            atomicAdd(&P.counters[Tr.out_shots ? 51 : 50], cyc);
            if (!Tr.out_shots) { atomicAdd(&P.counters[39], (int)((st_t1 - st_t0) >> 8)); atomicAdd(&P.counters[64], (int)((st_t2 - st_t1) >> 8)); }
        }
#endif
        __syncwarp();
    }
}

// ---- launcher ---------------------------------------------------------------------------------------------------
struct FreeTierPlan { int Q, rcap, rec_cap, warps, ctas_per_sm; size_t smem; };
struct FreePlan { int cap, cap2, cap_per_wt, max_wt; size_t smem_sel, smem_sel2; FreeTierPlan A, B; };

static bool tier_plan(const qb_decoder *dec, int Q, int rcap, int max_warps, FreeTierPlan &tp)
{
    const GraphDev &g = dec->g;
    tp.Q = Q;
    tp.rcap = std::max(128 * Q, std::min(rcap, (g.m + 31) & ~31)) & ~31;
    tp.rec_cap = tp.rcap * std::max(4, tp.rcap * 3 / 128);                  // 3/4 of rcap * rcap / 32 (the records are triangular)
    const size_t per_warp = free_per_warp_bytes(g.m, tp.rcap, Q);
    const size_t limit = (size_t)dec->max_smem_optin;
    if (per_warp + 1024 > limit) return false;
    tp.warps = (int)std::min<size_t>(max_warps, (limit - 1024) / per_warp);
    tp.smem = per_warp * tp.warps;
    tp.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (limit + 1024) / (tp.smem + 1024)));
    return true;
}

static bool free_plan(const qb_decoder *dec, FreePlan &pl)
{
    const GraphDev &g = dec->g;
    if (!g.colsig || g.n > 65535 || g.m > 65535 || g.m <= 0 || g.n <= 0) return false;
    if (getenv("QLDPC_B200_OSD_FULLWIDTH")) return false;
    const bool big = g.m > 1536;
    int rcapA = big ? 1536 : 512, rcapB = big ? 3072 : 1024;
    if (const char *e = getenv("QLDPC_B200_OSD_RCAP")) { const int v = atoi(e); if (v >= 64 && v <= 4096) { rcapA = v; rcapB = 2 * v; } }
    if (const char *e = getenv("QLDPC_B200_OSD_RCAP_B")) { const int v = atoi(e); if (v >= 64 && v <= 8192) rcapB = v; }
    pl.cap = big ? 8192 : 2048;
    if (const char *e = getenv("QLDPC_B200_OSD_CAP")) { const int v = atoi(e); if (v >= 64 && v <= 16384) pl.cap = v & ~31; }
    pl.cap = std::max(32, std::min(pl.cap, (g.n + 31) & ~31));
    pl.cap_per_wt = 32;
    if (const char *e = getenv("QLDPC_B200_OSD_CAP_PER_WT")) { const int v = atoi(e); if (v >= 1 && v <= 256) pl.cap_per_wt = v; }
    int warpsA = FREE_WARPS, warpsB = 6;
    if (const char *e = getenv("QLDPC_B200_OSD_WARPS_A")) { const int v = atoi(e); if (v >= 1 && v <= FREE_WARPS_MAX) warpsA = v; }
    if (const char *e = getenv("QLDPC_B200_OSD_WARPS_B")) { const int v = atoi(e); if (v >= 1 && v <= FREE_WARPS_MAX) warpsB = v; }
    if (!tier_plan(dec, big ? 2 : 1, rcapA, warpsA, pl.A)) return false;
    if (!tier_plan(dec, big ? 3 : 2, rcapB, warpsB, pl.B)) return false;
    pl.max_wt = 128 * pl.B.Q;
    pl.cap2 = (g.n + 31) & ~31;                                              // second pass: all columns
    if (const char *e = getenv("QLDPC_B200_OSD_CAP2")) { const int v = atoi(e); if (v >= 256) pl.cap2 = std::min(pl.cap2, v & ~31); }
    pl.smem_sel = sizeof(uint32_t) * SELF_BINS + (size_t)pl.cap * (4 + 2) + sizeof(uint32_t) * (size_t)g.mw + 16;
    pl.smem_sel2 = sizeof(uint32_t) * SELF_BINS + (size_t)pl.cap2 * (4 + 2) + sizeof(uint32_t) * (size_t)g.mw + 16;
    return pl.smem_sel + 1024 <= (size_t)dec->max_smem_optin && pl.smem_sel2 + 1024 <= (size_t)dec->max_smem_optin;
}

bool osd_free_applicable(const qb_decoder *dec) { FreePlan pl; return free_plan(dec, pl); }

template <int Q>
static int launch_tier(const OsdFreeArgs &P, const OsdFreeTier &T, const FreeTierPlan &tp, int grid, cudaStream_t st)
{
    QB_CUDA(cudaFuncSetAttribute(osd_free_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem));
    osd_free_kernel<Q><<<grid, tp.warps * 32, tp.smem, st>>>(P, T);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}
static int launch_tier_q(const OsdFreeArgs &P, const OsdFreeTier &T, const FreeTierPlan &tp, int grid, cudaStream_t st)
{
    switch (tp.Q) {
        case 1: return launch_tier<1>(P, T, tp, grid, st);
        case 2: return launch_tier<2>(P, T, tp, grid, st);
        default: return launch_tier<3>(P, T, tp, grid, st);
    }
}

// Selection + two tiers of free-row elimination for the queue of `a`; what fits neither tier ends in the overflow queue
// (*ovf_count_d / *ovf_idx_d: shot indices; device pointers valid until the decoder's next OSD launch).
int launch_osd0_free(qb_decoder *dec, const OsdLaunch &a, int32_t **ovf_count_d, int32_t **ovf_idx_d, cudaStream_t st)
{
    FreePlan pl;
    if (!free_plan(dec, pl)) { set_error("free-row OSD kernel not applicable"); return QB_ERR_UNSUPPORTED; }
    const GraphDev &g = dec->g;
    OsdFreeArgs P{};
    P.g = g; P.a = a; P.F = a.F;
    P.cap = pl.cap; P.cap_per_wt = pl.cap_per_wt; P.max_wt = pl.max_wt;
    const int gridA = std::max(1, std::min(ceil_div(a.F, pl.A.warps), dec->sm_count * pl.A.ctas_per_sm));
    int gridB = std::max(1, std::min(ceil_div(a.F, pl.B.warps), dec->sm_count * pl.B.ctas_per_sm));
    if (const char *e = getenv("QLDPC_B200_OSD_GRID_B")) { const int v = atoi(e); if (v >= 1) gridB = std::min(gridB, v); }
    const int sel_ctas = (int)std::max<size_t>(1, std::min<size_t>(2048 / SELF_THREADS, ((size_t)dec->max_smem_optin + 1024) / (pl.smem_sel + 1024)));
    const int grid_sel = std::max(1, std::min(a.F, dec->sm_count * sel_ctas));
    const size_t slotsA = (size_t)gridA * pl.A.warps, slotsB = (size_t)gridB * pl.B.warps;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t F = (size_t)a.F;
    // sides the second selection pass has buffers for: 1/8 of the queue where 1-4 % leave tier A (m <= 1536; gross code
    // 1.1 % at p = 0.005, 3.6 % at p = 0.006), 1/4 for the large codes (the [[288,12,18]] code at p = 0.006: 12 %)
    const size_t F2 = std::max<size_t>(64, g.m <= 1536 ? F / 8 : F / 4);
    const size_t need = 512 + al(F * pl.cap * 2) + al(F * 4) + al(F * g.mw * 4) + al(slotsA * pl.A.rec_cap * 4) + al(slotsA * pl.A.rcap * 4) +
                        al(slotsB * pl.B.rec_cap * 4) + al(slotsB * pl.B.rcap * 4) + 4 * al(F * 4) + al(F2 * pl.cap2 * 2) + al(F2 * 4);
    const bool grown = dec->ovf.cap < need;
    if (int rc = dec->ovf.ensure(need)) return rc;
    unsigned char *p = dec->ovf.as<unsigned char>();
    if (grown) QB_CUDA(cudaMemsetAsync(p, 0, 512, st));
    P.counters = reinterpret_cast<int32_t *>(p); p += 512;
    P.cand = reinterpret_cast<uint16_t *>(p); p += al(F * pl.cap * 2);
    P.ncand = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    P.res = reinterpret_cast<uint32_t *>(p); p += al(F * g.mw * 4);
    OsdFreeTier TA{}, TB{};
    TA.rcap = pl.A.rcap; TA.rec_cap = pl.A.rec_cap; TB.rcap = pl.B.rcap; TB.rec_cap = pl.B.rec_cap;
    TA.rec = reinterpret_cast<uint32_t *>(p); p += al(slotsA * pl.A.rec_cap * 4);
    TA.meta = reinterpret_cast<uint32_t *>(p); p += al(slotsA * pl.A.rcap * 4);
    TB.rec = reinterpret_cast<uint32_t *>(p); p += al(slotsB * pl.B.rec_cap * 4);
    TB.meta = reinterpret_cast<uint32_t *>(p); p += al(slotsB * pl.B.rcap * 4);
    int32_t *listA = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    int32_t *listB = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    int32_t *listW = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    P.win_slot = reinterpret_cast<int32_t *>(p); p += al(F * 4);
    uint16_t *cand2 = reinterpret_cast<uint16_t *>(p); p += al(F2 * pl.cap2 * 2);
    int32_t *ncand2 = reinterpret_cast<int32_t *>(p);
    P.win_list = listW; P.cand2 = cand2; P.ncand2 = ncand2; P.cap2 = pl.cap2; P.sel_max = (int)F2;
    TA.queue_counter = CNT_A; TA.in_count = nullptr; TA.in_q = nullptr; TA.out_counter = CNT_OVF_A; TA.out_list = listA; TA.out_shots = 0; TA.win_list = listW;
    TB.queue_counter = CNT_B; TB.in_count = P.counters + CNT_OVF_A; TB.in_q = listA; TB.out_counter = CNT_OVF_B; TB.out_list = listB; TB.out_shots = 1; TB.win_list = nullptr;
    static const bool dbg = getenv("QLDPC_B200_DEBUG_SYNC") != nullptr;
    // the statistics words accumulate over launches (qb_debug_osd_free_stats reads and clears them)
    QB_CUDA(cudaMemsetAsync(P.counters, 0, CNT_STAT * sizeof(int32_t), st));
    QB_CUDA(cudaMemsetAsync(P.win_slot, 0xFF, F * sizeof(int32_t), st));
    QB_CUDA(cudaFuncSetAttribute(osd_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(pl.smem_sel, pl.smem_sel2)));
    P.sel_count = nullptr; P.sel_q = nullptr; P.sel_counter = CNT_SEL;
    osd_select_kernel<<<grid_sel, SELF_THREADS, pl.smem_sel, st>>>(P);
    QB_CUDA(cudaGetLastError());
    if (dbg) QB_CUDA(cudaStreamSynchronize(st));
    if (int rc = launch_tier_q(P, TA, pl.A, gridA, st)) return rc;
    if (dbg) QB_CUDA(cudaStreamSynchronize(st));
    {   // sides that ran out of candidates get the full window before tier B looks at them
        OsdFreeArgs P2 = P;
        P2.sel_count = P.counters + CNT_WIN_A; P2.sel_q = listW; P2.sel_counter = CNT_SEL2;
        P2.cand = cand2; P2.ncand = ncand2; P2.cap = pl.cap2;
        QB_CUDA(cudaFuncSetAttribute(osd_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(pl.smem_sel, pl.smem_sel2)));
        osd_select_kernel<<<(int)std::min<size_t>(F2, (size_t)dec->sm_count * 2), SELF_THREADS, pl.smem_sel2, st>>>(P2);
        QB_CUDA(cudaGetLastError());
        if (dbg) QB_CUDA(cudaStreamSynchronize(st));
    }
    if (int rc = launch_tier_q(P, TB, pl.B, gridB, st)) return rc;
    if (dbg) QB_CUDA(cudaStreamSynchronize(st));
    *ovf_count_d = P.counters + CNT_OVF_B;
    *ovf_idx_d = listB;
    return QB_OK;
}

// statistics of the free-row path since the last call: sides that left tier A (out[0..4]) / tier B (out[5..9], these go to
// the full-width kernel) because of {window exhausted, touched rows, free slots, record buffer, not materialised}
#ifdef QB_OSD_STATS
extern "C" int qb_debug_osd_work(qb_decoder *dec, int32_t *out8)
{
    if (!dec->ovf.ptr) return -1;
    cudaDeviceSynchronize();
    int32_t *c = dec->ovf.as<int32_t>() + 32;
    if (cudaMemcpy(out8, c, 40 * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    cudaMemset(c, 0, 40 * sizeof(int32_t));
    return 0;
}
#endif

int osd_free_stats(qb_decoder *dec, int32_t *out10)
{
    for (int i = 0; i < 10; ++i) out10[i] = 0;
    if (!dec->ovf.ptr) return QB_OK;
    QB_CUDA(cudaDeviceSynchronize());
    int32_t *c = dec->ovf.as<int32_t>() + CNT_STAT;
    int32_t h[16];
    QB_CUDA(cudaMemcpy(h, c, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemset(c, 0, 16 * sizeof(int32_t)));
    for (int i = 0; i < 5; ++i) { out10[i] = h[i]; out10[5 + i] = h[8 + i]; }
    return QB_OK;
}

}  // namespace qb
