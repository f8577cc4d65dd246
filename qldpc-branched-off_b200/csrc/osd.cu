// K4+K5: batched OSD-0 on the sides min-sum did not converge on, and a dense GF(2) Gauss-Jordan.
//
// Reference semantics: performOSD_enhanced order 0 (src/decoding/osd.py:5-29) on top of
// gf2_elimination_packed_core (src/decoding/kernels.py:49-96).
//
// The reference permutes the dense m x n matrix by reliability and runs a full Gauss-Jordan sweep
// (1008 x 8785 bits per side for the gross code).  Here one CTA handles one failed side and
//   1. forms the residual syndrome  s = syndrome ^ H.hard   (osd.py:7-9);
//   2. sorts the columns by |posterior| ascending with a stable in-CTA LSD radix sort on the float
//      bit patterns (ties by column index; osd.py:11-12 uses an unstable argsort, so callers that
//      need the reference's exact tie order pass the ordering in);
//   3. eliminates *without materialising the permuted matrix*: H is fixed and column-sparse, so it
//      keeps only the row transform T restricted to the columns that belong to pivot rows
//      (T.e_r = e_r for every non-pivot row r).  A candidate column c is reduced as
//      v = XOR_{r in supp(h_c)} T.e_r  (<= 6 shared-memory vectors), it pivots iff v has a bit on a
//      non-pivot row, and the pivot row is the one the reference would pick (first row at or below
//      the current one in its swapped row order, kernels.py:71-82), tracked with a position table
//      instead of physically swapping;
//   4. stops as soon as the transformed syndrome has no bit left on non-pivot rows: all remaining
//      pivots would get e = 0 (their s_reduced entries can no longer change), so the result equals
//      the full sweep bit for bit while typically needing ~150 instead of ~930 pivots;
//   5. flips hard[ordering[pivot_col]] where s_reduced[pivot_row] = 1   (osd.py:19-25).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace qb {

constexpr int OSD_CTAS_PER_SM = 9;
constexpr int OSD_NW = OSD_THREADS / 32;   // warps per CTA = candidates reduced per round
constexpr int SEL_BINS = 2048;             // histogram bins: 64 per octave over 2^-25 .. 2^7, clamped (monotone in the key)
constexpr int SEL_SHIFT = 17;
constexpr int SEL_BASE = (127 - 25) << 6;
constexpr int SEL_CAP = 1024;              // candidates materialised per selection window
constexpr int SEL_MIN = 640;               // a window is closed once it holds at least this many

struct OsdArgs {
    GraphDev g;
    OsdLaunch a;
    int tcap;             // T columns resident in shared memory; the rest spills to gT
    int sel_min;          // a selection window is closed once it holds at least this many candidates
    int cstride;          // words per T column (mw + 1 when mw is even: conflict-free column-parallel reads)
    int rank_cap;         // min(m, n)
    uint32_t *gT;         // [grid][(rank_cap - tcap) * cstride] spill
    // per-CTA global scratch for the rare paths: later selection windows and the full-sort fallback
    uint32_t *g_hist;     // [grid][SEL_BINS]
    uint32_t *g_listK;    // [grid][SEL_CAP]
    uint16_t *g_listI;    // [grid][SEL_CAP]
    uint32_t *g_keys;     // [grid][n]          (full sort)
    uint16_t *g_idx;      // [grid][2][n_pad2]  (full sort)
    uint32_t *g_cnt;      // [grid][256 * OSD_NW]
    int32_t *work_counter; // device counter, zero at launch
    uint16_t *g_pivpos;   // [grid][rank_cap] pivot positions in the ordering (only filled when pivots_out is set)
    uint32_t *g_piv;      // [grid][rank_cap] pivot row | column << 16 (written once per pivot, read once at the end)
};

__device__ __forceinline__ int sel_bin(uint32_t key) { return min(max((int)(key >> SEL_SHIFT) - SEL_BASE, 0), SEL_BINS - 1); }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }

// Stable LSD radix sort (4 x 8 bit) of idx[] by keys[idx]; warps own contiguous segments so the
// (digit, warp) counter order is the element order.  Result ends in idx0.  Fallback path only.
__device__ void full_radix_sort(const uint32_t *keys, uint16_t *idx0, uint16_t *idx1, uint32_t *cnt, int n)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    __shared__ uint32_t wsum[32];
    const int seg = ((n + NW - 1) / NW + 31) & ~31;
    const int s0 = min(n, warp * seg), s1 = min(n, s0 + seg);
    uint16_t *src = idx0, *dst = idx1;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        for (int i = tid; i < 256 * NW; i += blockDim.x) cnt[i] = 0u;
        __syncthreads();
        for (int i0 = s0; i0 < s1; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < s1;
            const uint32_t d = valid ? ((keys[src[i]] >> shift) & 255u) : (256u + lane);
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
            if (valid && (peers & lanemask_lt()) == 0) cnt[d * NW + warp] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        {
            const int total = 256 * NW, per = (total + blockDim.x - 1) / blockDim.x;
            const int b0 = tid * per;
            uint32_t local = 0;
            for (int i = b0; i < min(total, b0 + per); ++i) local += cnt[i];
            uint32_t inc = local;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                uint32_t x = lane < NW ? wsum[lane] : 0u, xi = x;
                for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (lane >= o) xi += y; }
                wsum[lane] = xi - x;
            }
            __syncthreads();
            uint32_t run = wsum[warp] + inc - local;
            for (int i = b0; i < min(total, b0 + per); ++i) { const uint32_t c = cnt[i]; cnt[i] = run; run += c; }
        }
        __syncthreads();
        for (int i0 = s0; i0 < s1; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < s1;
            const uint16_t id = valid ? src[i] : (uint16_t)0;
            const uint32_t d = valid ? ((keys[id] >> shift) & 255u) : (256u + lane);
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
            if (valid) dst[cnt[d * NW + warp] + __popc(peers & lanemask_lt())] = id;
            __syncwarp();
            if (valid && (peers & lanemask_lt()) == 0) cnt[d * NW + warp] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        uint16_t *tmp = src; src = dst; dst = tmp;
    }
}

#ifdef QB_OSD_PROFILE
// instrumented build (tools/osd_side_profile.py): per side {pivots, candidates examined, cycles, start time}
constexpr int OSD_PROF_CAP = 1 << 18;
__device__ unsigned long long g_osd_prof[OSD_PROF_CAP * 8];   // + cycles of: setup/residual, histogram, windows, writeback
__device__ int g_osd_prof_n;
#endif

// EXACTROWS: choose each pivot row exactly like the reference's swapped row order (needed only when the
// syndrome may be inconsistent, i.e. outside the column space of H, where the result depends on it);
// for consistent syndromes -- every simulated shot -- the solution is unique and the lowest free row is used.
template <int WPL, bool EXACTROWS>
__global__ void __launch_bounds__(OSD_THREADS, OSD_CTAS_PER_SM) osd0_kernel(OsdArgs P)
{
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = OSD_NW;
    const int mw = g.mw, n = g.n, m = g.m, cs = P.cstride;
    const int n_pad2 = (n + 1) & ~1;

    // ---- shared memory carve-up -------------------------------------------------------------------
    unsigned char *sp = smem_raw;
    uint16_t *pos_of_row = nullptr, *row_at_pos = nullptr;
    if (EXACTROWS) {
        pos_of_row = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * g.m_pad;
        row_at_pos = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * g.m_pad;
    }
    int16_t *pivcol_of_row = reinterpret_cast<int16_t *>(sp); sp += sizeof(int16_t) * g.m_pad;
    uint32_t *npmask = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    uint32_t *sv = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    uint32_t *pv = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    uint16_t *ord = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * SEL_CAP;   // current window, sorted
    uint32_t *regionX = reinterpret_cast<uint32_t *>(sp);
    // selection scratch of the first window overlays T (T is empty until the first pivot)
    uint32_t *s_hist = regionX;                 // per bin: low 16 bits = count still to place, high 16 = window offset
    uint32_t *s_listK = s_hist + SEL_BINS;
    uint16_t *s_listI = reinterpret_cast<uint16_t *>(s_listK + SEL_CAP);
    uint32_t *Tsm = regionX;
    uint32_t *Tgl = P.gT ? P.gT + (size_t)blockIdx.x * (size_t)(P.rank_cap - P.tcap) * cs : nullptr;
    uint16_t *piv_pos = P.g_pivpos + (size_t)blockIdx.x * P.rank_cap;
    uint32_t *piv_rc = P.g_piv + (size_t)blockIdx.x * P.rank_cap;
    __shared__ int s_flags[NW];
    __shared__ int s_rho, s_binhi, s_wincount;

    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.a.F) : P.a.F;

    __shared__ int s_qi;
    while (true) {
        // dynamic work distribution: per-side cost varies by more than an order of magnitude
        if (tid == 0) s_qi = atomicAdd(P.work_counter, 1);
        __syncthreads();
        const int qi = s_qi;
        __syncthreads();
        if (qi >= F) break;
        const int shot = P.a.fail_idx ? P.a.fail_idx[qi] : qi;
#ifdef QB_OSD_PROFILE
        const long long prof_t0 = clock64();
        long long prof_win = 0, prof_t1 = 0, prof_t2 = 0, prof_t3 = 0;
#endif
        const uint32_t *hard = P.a.hard_bits + (size_t)shot * g.nw;
        const int32_t *ext_order = P.a.ordering ? P.a.ordering + (size_t)shot * n : nullptr;
        const float *post = P.a.post ? P.a.post + (size_t)shot * n : nullptr;

        // ---- 1. residual syndrome, bookkeeping ----------------------------------------------------
        for (int w = tid; w < 32 * WPL; w += blockDim.x) {
            sv[w] = w < mw ? P.a.syn_bits[(size_t)shot * mw + w] : 0u;
            uint32_t full = 0u;
            if (w * 32 + 32 <= m) full = 0xFFFFFFFFu;
            else if (w * 32 < m) full = (1u << (m - w * 32)) - 1u;
            npmask[w] = full;
        }
        for (int r = tid; r < g.m_pad; r += blockDim.x) { if (EXACTROWS) { pos_of_row[r] = (uint16_t)r; row_at_pos[r] = (uint16_t)r; } pivcol_of_row[r] = -1; }
        if (!ext_order) for (int b = tid; b < SEL_BINS; b += blockDim.x) s_hist[b] = 0u;
        __syncthreads();
        for (int w = tid; w < g.nw; w += blockDim.x) {
            uint32_t bits = hard[w];
            while (bits) {
                const int b = __ffs(bits) - 1; bits &= bits - 1;
                const int j = w * 32 + b;
                if (j < n)
                    for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) { const int r = g.rowidx[p]; atomicXor(&sv[r >> 5], 1u << (r & 31)); }
            }
        }
#ifdef QB_OSD_PROFILE
        prof_t1 = clock64();
#endif
        // ---- 2a. histogram of |posterior| bit patterns (selection pass 1) ---------------------------
        if (!ext_order) {
            for (int j0 = tid; j0 < n; j0 += 8 * OSD_THREADS) {      // 8 independent loads in flight per thread
                uint32_t kb[8];
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) { const int j = j0 + u8 * OSD_THREADS; kb[u8] = j < n ? (uint32_t)sel_bin(__float_as_uint(fabsf(post[j]))) : 0xFFFFFFFFu; }
#pragma unroll
                for (int u8 = 0; u8 < 8; ++u8) if (kb[u8] != 0xFFFFFFFFu) atomicAdd(&s_hist[kb[u8]], 1u);
            }
        }
        __syncthreads();

#ifdef QB_OSD_PROFILE
        prof_t2 = clock64();
#endif
        // ordering modes: 0 = caller-supplied, 1 = selection windows, 2 = full sort in global scratch
        int mode = ext_order ? 0 : 1;
        int bin_next = 0;                 // first histogram bin not yet consumed
        int win_start = 0, win_end = ext_order ? n : 0;
        uint32_t *hist = s_hist; uint32_t *listK = s_listK; uint16_t *listI = s_listI;
        const uint16_t *gsorted = nullptr;

        // T column x, word w: shared memory for x < tcap (32-bit addressing), global spill beyond
        auto ldT = [&](int x, int w) -> uint32_t { return x < P.tcap ? Tsm[x * cs + w] : Tgl[(size_t)(x - P.tcap) * cs + w]; };
        auto stT = [&](int x, int w, uint32_t val) { if (x < P.tcap) Tsm[x * cs + w] = val; else Tgl[(size_t)(x - P.tcap) * cs + w] = val; };
        auto xorT = [&](int x, int w, uint32_t val) { if (x < P.tcap) Tsm[x * cs + w] ^= val; else Tgl[(size_t)(x - P.tcap) * cs + w] ^= val; };
        auto order_at = [&](int c) -> int {
            if (mode == 0) return ext_order[c];
            if (mode == 2) return (int)gsorted[c];
            return (int)ord[c - win_start];
        };
        // every warp keeps its own register copy of the transformed syndrome and of the non-pivot-row
        // mask (lane = word) and updates them identically, so the termination test needs no barrier
        uint32_t svr[WPL], npr[WPL];
#pragma unroll
        for (int i = 0; i < WPL; ++i) { svr[i] = sv[lane + 32 * i]; npr[i] = npmask[lane + 32 * i]; }
        auto unresolved = [&]() -> bool {
            bool any = false;
#pragma unroll
            for (int i = 0; i < WPL; ++i) any |= (svr[i] & npr[i]) != 0u;
            return __any_sync(0xFFFFFFFFu, any);
        };
        // reduce column j against the current T (lane = word)
        auto reduce_column = [&](int j, uint32_t (&v)[WPL]) {
#pragma unroll
            for (int i = 0; i < WPL; ++i) v[i] = 0u;
            auto add_row = [&](int r) {
                const int pc = pivcol_of_row[r];
                if (pc >= 0) {
#pragma unroll
                    for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; if (w < mw) v[i] ^= ldT(pc, w); }
                } else {
#pragma unroll
                    for (int i = 0; i < WPL; ++i) if ((r >> 5) == lane + 32 * i) v[i] ^= 1u << (r & 31);
                }
            };
            if (g.colsig) {
                const uint4 sg = g.colsig[j];
                const uint32_t rr[8] = {sg.x & 0xFFFFu, sg.x >> 16, sg.y & 0xFFFFu, sg.y >> 16, sg.z & 0xFFFFu, sg.z >> 16, sg.w & 0xFFFFu, sg.w >> 16};
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) if (rr[k2] != 0xFFFFu) add_row((int)rr[k2]);
            } else {
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) add_row(g.rowidx[p]);
            }
        };

        int t = 0;
        bool done = !unresolved();
        int c0 = 0;
        while (!done && c0 < n && t < P.rank_cap) {
            // ---- 2b. materialise the next window of candidates, sorted by (|posterior|, index) --------
#ifdef QB_OSD_PROFILE
            const long long prof_w0 = clock64();
#endif
            if (c0 >= win_end && mode == 1) {
                if (t > 0) {       // T now lives in regionX: later windows use the global scratch
                    hist = P.g_hist + (size_t)blockIdx.x * SEL_BINS;
                    listK = P.g_listK + (size_t)blockIdx.x * SEL_CAP; listI = P.g_listI + (size_t)blockIdx.x * SEL_CAP;
                    for (int b = tid; b < SEL_BINS; b += blockDim.x) hist[b] = 0u;
                    __syncthreads();
                    for (int j = tid; j < n; j += blockDim.x) {
                        const int b = sel_bin(__float_as_uint(fabsf(post[j])));
                        if (b >= bin_next) atomicAdd(&hist[b], 1u);
                    }
                    __syncthreads();
                }
                if (warp == 0) {   // choose [bin_next, bin_hi]: close the window at >= SEL_MIN, never exceed SEL_CAP
                    int cum = 0, hi = bin_next - 1;
                    bool stop = false;
                    for (int b0 = bin_next; b0 < SEL_BINS && !stop; b0 += 32) {
                        const int b = b0 + lane;
                        const int cnt = b < SEL_BINS ? (int)(hist[b] & 0xFFFFu) : 0;
                        int inc = cnt;
                        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
                        if (b < SEL_BINS) hist[b] = ((uint32_t)min(cum + inc - cnt, 65535) << 16) | (uint32_t)cnt;
                        const uint32_t over = __ballot_sync(0xFFFFFFFFu, cum + inc > SEL_CAP);
                        const uint32_t enough = __ballot_sync(0xFFFFFFFFu, cum + inc >= P.sel_min);
                        int last = 31;
                        if (over | enough) {
                            const int fo = over ? __ffs(over) - 1 : 32, fe = enough ? __ffs(enough) - 1 : 32;
                            last = fe < fo ? fe : fo - 1;          // stop before a bin that would overflow the window
                            stop = true;
                        }
                        if (last >= 0) { cum += __shfl_sync(0xFFFFFFFFu, inc, last) - 0; hi = b0 + last; }
                    }
                    if (!stop) hi = SEL_BINS - 1;
                    if (lane == 0) { s_binhi = hi; s_wincount = cum; }
                }
                __syncthreads();
                const int bin_hi = s_binhi, M = s_wincount;
                __syncthreads();
                if (M == 0) {
                    if (bin_hi >= SEL_BINS - 1) break;                 // no candidates left
                    mode = 2;                                          // a single bin exceeds SEL_CAP: full sort
                } else {
                    // scatter (unordered inside a bin), then rank inside each bin by (key, index)
                    for (int j0 = tid; j0 < n; j0 += 8 * OSD_THREADS) {
                        uint32_t kb[8];
#pragma unroll
                        for (int u8 = 0; u8 < 8; ++u8) { const int j = j0 + u8 * OSD_THREADS; kb[u8] = j < n ? __float_as_uint(fabsf(post[j])) : 0xFFFFFFFFu; }
#pragma unroll
                        for (int u8 = 0; u8 < 8; ++u8) {
                            const uint32_t key = kb[u8];
                            const int b = sel_bin(key);
                            if (key != 0xFFFFFFFFu && b >= bin_next && b <= bin_hi) {
                                const uint32_t old = atomicSub(&hist[b], 1u);       // count down: consumed bins end at count 0
                                const int slot = (int)(old >> 16) + (int)(old & 0xFFFFu) - 1;
                                listK[slot] = key; listI[slot] = (uint16_t)(j0 + u8 * OSD_THREADS);
                            }
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < M; i += blockDim.x) {
                        const uint32_t key = listK[i];
                        const uint16_t id = listI[i];
                        const int b = sel_bin(key);
                        const int lo = (int)(hist[b] >> 16);
                        int hi2 = M;
                        if (b < bin_hi) hi2 = (int)(hist[b + 1] >> 16);       // offsets are non-decreasing over the window
                        int rank = lo;
                        for (int k2 = lo; k2 < hi2; ++k2) {
                            const uint32_t kk = listK[k2];
                            rank += (kk < key) || (kk == key && listI[k2] < id);
                        }
                        ord[rank] = id;
                    }
                    win_start = win_end; win_end += M; bin_next = bin_hi + 1;
                    __syncthreads();
                }
            }
            if (mode == 2 && gsorted == nullptr) {
                uint32_t *keys = P.g_keys + (size_t)blockIdx.x * n;
                uint16_t *idx0 = P.g_idx + (size_t)blockIdx.x * 2 * n_pad2, *idx1 = idx0 + n_pad2;
                for (int j = tid; j < n; j += blockDim.x) { keys[j] = __float_as_uint(fabsf(post[j])); idx0[j] = (uint16_t)j; }
                __syncthreads();
                full_radix_sort(keys, idx0, idx1, P.g_cnt + (size_t)blockIdx.x * 256 * NW, n);
                gsorted = idx0; win_start = 0; win_end = n;
                __syncthreads();
            }

#ifdef QB_OSD_PROFILE
            prof_win += clock64() - prof_w0;
#endif
            // ---- 3. reduce NW candidates against the current T -----------------------------------------
            const int c = c0 + warp;
            const bool have = c < win_end && c < n;
            uint32_t v[WPL];
#pragma unroll
            for (int i = 0; i < WPL; ++i) v[i] = 0u;
            int myj = 0;
            if (have) { myj = order_at(c); reduce_column(myj, v); }
            while (true) {
                bool f_ = false;
#pragma unroll
                for (int i = 0; i < WPL; ++i) f_ |= (v[i] & npr[i]) != 0u;
                const bool flag = __any_sync(0xFFFFFFFFu, f_);
                if (lane == 0) s_flags[warp] = flag ? 1 : 0;
                __syncthreads();
                const uint32_t fb = __ballot_sync(0xFFFFFFFFu, lane < NW && s_flags[lane] != 0);
                if (fb == 0u) break;
                const int f = __ffs(fb) - 1;
                if (warp == f) {
                    // pivot row: EXACTROWS = first row, in the reference's current (swapped) row order, with the
                    // bit set; otherwise the lowest free row with the bit set
                    uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) {
                        const int w = lane + 32 * i;
                        uint32_t bits = v[i] & npr[i];
                        if (EXACTROWS) {
                            while (bits) {
                                const int b = __ffs(bits) - 1; bits &= bits - 1;
                                const int r = w * 32 + b;
                                best = min(best, ((uint32_t)pos_of_row[r] << 16) | (uint32_t)r);
                            }
                        } else if (bits) {
                            best = min(best, (uint32_t)(w * 32 + __ffs(bits) - 1));
                        }
                    }
                    best = __reduce_min_sync(0xFFFFFFFFu, best);
                    const int q = EXACTROWS ? (int)(best >> 16) : 0, rho = best & 0xFFFF;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) {
                        const int w = lane + 32 * i;
                        if (w < mw) {
                            stT(t, w, v[i]);                                  // T.e_rho after this step
                            pv[w] = ((rho >> 5) == w) ? (v[i] & ~(1u << (rho & 31))) : v[i];
                        }
                        v[i] = 0u;
                    }
                    if (lane == 0) {
                        if (EXACTROWS) {
                            const int rt = row_at_pos[t];
                            row_at_pos[t] = (uint16_t)rho; row_at_pos[q] = (uint16_t)rt;
                            pos_of_row[rt] = (uint16_t)q; pos_of_row[rho] = (uint16_t)t;
                        }
                        pivcol_of_row[rho] = (int16_t)t;
                        piv_rc[t] = (uint32_t)rho | ((uint32_t)myj << 16);
                        if (P.a.pivots_out) piv_pos[t] = (uint16_t)c;
                        s_rho = rho;
                    }
                } else if (warp < f) {
#pragma unroll
                    for (int i = 0; i < WPL; ++i) v[i] = 0u;                  // dependent on earlier columns
                }
                __syncthreads();
                const int rho = s_rho;
                const int rw = rho >> 5, rl = rw & 31, ri = rw >> 5;
                const uint32_t rbit = 1u << (rho & 31);
                uint32_t u[WPL];
#pragma unroll
                for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; u[i] = w < mw ? pv[w] : 0u; }
                // (a) the candidates still held in registers
                if (warp > f) {
                    uint32_t mine = 0u;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) if (i == ri) mine = v[i];
                    if (__shfl_sync(0xFFFFFFFFu, mine, rl) & rbit) {
#pragma unroll
                        for (int i = 0; i < WPL; ++i) v[i] ^= u[i];
                    }
                }
                // (b) the transformed syndrome and the non-pivot mask (register copies, all warps)
                {
                    uint32_t mine = 0u;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) if (i == ri) mine = svr[i];
                    if (__shfl_sync(0xFFFFFFFFu, mine, rl) & rbit) {
#pragma unroll
                        for (int i = 0; i < WPL; ++i) svr[i] ^= u[i];
                    }
#pragma unroll
                    for (int i = 0; i < WPL; ++i) if (i == ri && lane == rl) npr[i] &= ~rbit;
                }
                // (c) stored columns of earlier pivots; column x is only ever updated by warp x % NW, so successive
                // pivots need no barrier between their updates, and the work is spread over the warps for any t.
                // 32 columns are tested per read.
                for (int i0 = 0; i0 * NW + warp < t; i0 += 32) {
                    const int x = (i0 + lane) * NW + warp;
                    const bool has = x < t && (ldT(x, rw) & rbit) != 0u;
                    uint32_t msk = __ballot_sync(0xFFFFFFFFu, has);
                    while (msk) {
                        const int b = __ffs(msk) - 1; msk &= msk - 1;
#pragma unroll
                        for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; if (w < mw) xorT((i0 + b) * NW + warp, w, u[i]); }
                    }
                }
                ++t;
                done = !unresolved();
                if (done || t >= P.rank_cap) break;
            }
            __syncthreads();
            c0 += NW;
            if (mode == 1 && c0 > win_end) c0 = win_end;   // do not skip candidates of the next window
        }

#ifdef QB_OSD_PROFILE
        prof_t3 = clock64();
#endif
        // ---- 5. solution = hard ^ e, e[ordering[pivot_col]] = s_reduced[pivot_row] ---------------
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < WPL; ++i) sv[lane + 32 * i] = svr[i];
        }
        __syncthreads();
        uint32_t *hard_rw = P.a.hard_bits + (size_t)shot * g.nw;
        for (int i = tid; i < t; i += blockDim.x) {
            const uint32_t prc = piv_rc[i];
            const int rho = prc & 0xFFFFu;
            if ((sv[rho >> 5] >> (rho & 31)) & 1u) {
                const int j = (int)(prc >> 16);
                atomicXor(&hard_rw[j >> 5], 1u << (j & 31));
            }
        }
        if (P.a.pivots_out)
            for (int i = tid; i < P.rank_cap; i += blockDim.x)
                P.a.pivots_out[(size_t)shot * P.rank_cap + i] = i < t ? (int)piv_pos[i] : -1;
        if (P.a.rank_out && tid == 0) P.a.rank_out[shot] = t | P.a.rank_tag;
#ifdef QB_OSD_PROFILE
        if (tid == 0) {
            const int slot = atomicAdd(&g_osd_prof_n, 1);
            if (slot < OSD_PROF_CAP) {
                g_osd_prof[slot * 8 + 0] = (unsigned long long)t; g_osd_prof[slot * 8 + 1] = (unsigned long long)c0;
                g_osd_prof[slot * 8 + 2] = (unsigned long long)(clock64() - prof_t0);
                g_osd_prof[slot * 8 + 4] = (unsigned long long)(prof_t1 - prof_t0); g_osd_prof[slot * 8 + 5] = (unsigned long long)(prof_t2 - prof_t1);
                g_osd_prof[slot * 8 + 6] = (unsigned long long)prof_win; g_osd_prof[slot * 8 + 7] = (unsigned long long)(clock64() - prof_t3);
                unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_osd_prof[slot * 8 + 3] = gt;
            }
        }
#endif
        __syncthreads();
    }
}

template <int WPL, bool EXACTROWS>
static int launch_osd_wpl(qb_decoder *dec, const OsdLaunch &a, cudaStream_t st)
{
    const GraphDev &g = dec->g;
    OsdArgs P{};
    P.g = g; P.a = a;
    P.rank_cap = std::min(g.m, g.n);
    P.cstride = (g.mw & 1) ? g.mw : g.mw + 1;
    P.sel_min = SEL_MIN;
    if (const char *e = getenv("QLDPC_B200_OSD_SEL_MIN")) { const int v = atoi(e); if (v >= 32 && v <= SEL_CAP) P.sel_min = v; }
    const int n_pad2 = (g.n + 1) & ~1;
    const size_t fixed = sizeof(uint16_t) * (EXACTROWS ? 3 : 1) * (size_t)g.m_pad + sizeof(uint32_t) * 32 * WPL * 3 + sizeof(uint16_t) * SEL_CAP;
    const size_t sel_b = sizeof(uint32_t) * SEL_BINS + sizeof(uint32_t) * SEL_CAP + sizeof(uint16_t) * SEL_CAP + 16;
    const size_t budget = (size_t)dec->max_smem_optin - 1024;
    const size_t colb = sizeof(uint32_t) * (size_t)P.cstride;
    const size_t want = colb * (size_t)P.rank_cap;
    // one warp per side, OSD_CTAS_PER_SM sides in flight per SM (1 KB of shared memory per CTA is reserved by
    // the system); T columns beyond the shared-memory share spill to global memory
    int target = OSD_CTAS_PER_SM;
    if (const char *e = getenv("QLDPC_B200_OSD_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= OSD_CTAS_PER_SM) target = v; }
    const size_t per_cta = ((size_t)dec->max_smem_optin + 1024) / target - 1024 - 256;
    size_t share = per_cta > fixed ? per_cta - fixed : 0;
    size_t regionX = std::max(sel_b, std::min(want, share));
    size_t smem = fixed + regionX;
    QB_REQUIRE(smem <= budget, "OSD: problem too large for shared memory");
    P.tcap = (int)std::min<size_t>(P.rank_cap, regionX / colb);
    const int ctas_per_sm = std::max(1, std::min(target, (int)(((size_t)dec->max_smem_optin + 1024) / (smem + 1024))));
    const int grid = std::max(1, std::min(a.F, dec->sm_count * ctas_per_sm));
    const size_t spill = (size_t)(P.rank_cap - P.tcap) * P.cstride * sizeof(uint32_t);
    const size_t b_hist = sizeof(uint32_t) * SEL_BINS;
    const size_t b_lk = sizeof(uint32_t) * SEL_CAP, b_li = sizeof(uint16_t) * SEL_CAP;
    const size_t b_keys = sizeof(uint32_t) * (size_t)g.n, b_idx = sizeof(uint16_t) * 2 * (size_t)n_pad2, b_cnt = sizeof(uint32_t) * 256 * OSD_NW;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t G = (size_t)grid;
    const size_t b_pp = sizeof(uint16_t) * (size_t)std::max(1, P.rank_cap);
    const size_t b_pr = sizeof(uint32_t) * (size_t)std::max(1, P.rank_cap);
    const size_t need = 256 + al(G * b_pp) + al(G * b_pr) + al(G * spill) + al(G * b_hist) + al(G * b_lk) + al(G * b_li) + al(G * b_keys) + al(G * b_idx) + al(G * b_cnt) + 256;
    if (int rc = dec->work.ensure(need)) return rc;
    unsigned char *p = dec->work.as<unsigned char>();
    P.work_counter = reinterpret_cast<int32_t *>(p); p += 256;
    QB_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(int32_t), st));
    P.gT = spill ? reinterpret_cast<uint32_t *>(p) : nullptr; p += al(G * spill);
    P.g_hist = reinterpret_cast<uint32_t *>(p); p += al(G * b_hist);
    P.g_listK = reinterpret_cast<uint32_t *>(p); p += al(G * b_lk);
    P.g_listI = reinterpret_cast<uint16_t *>(p); p += al(G * b_li);
    P.g_keys = reinterpret_cast<uint32_t *>(p); p += al(G * b_keys);
    P.g_idx = reinterpret_cast<uint16_t *>(p); p += al(G * b_idx);
    P.g_cnt = reinterpret_cast<uint32_t *>(p); p += al(G * b_cnt);
    P.g_pivpos = reinterpret_cast<uint16_t *>(p); p += al(G * b_pp);
    P.g_piv = reinterpret_cast<uint32_t *>(p);
    QB_CUDA(cudaFuncSetAttribute(osd0_kernel<WPL, EXACTROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    osd0_kernel<WPL, EXACTROWS><<<grid, OSD_THREADS, smem, st>>>(P);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

// Counting sort of the failure queue by residual-syndrome weight, heaviest first: the number of pivots an
// OSD-0 needs grows with that weight, so the dynamic work distribution starts the long eliminations first.
__global__ void __launch_bounds__(1024) sort_failures_kernel(const int32_t *fail_idx, const int32_t *fail_wt, const int32_t *n_fail,
                                                             int32_t *sorted_idx)
{
    __shared__ int hist[256], start[256];
    const int tid = threadIdx.x, F = *n_fail;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < F; i += blockDim.x) atomicAdd(&hist[255 - min(255, max(0, fail_wt[i] >> 1))], 1);
    __syncthreads();
    if (tid == 0) { int run = 0; for (int b = 0; b < 256; ++b) { start[b] = run; run += hist[b]; } }
    __syncthreads();
    for (int i = tid; i < F; i += blockDim.x) sorted_idx[atomicAdd(&start[255 - min(255, max(0, fail_wt[i] >> 1))], 1)] = fail_idx[i];
}

int launch_sort_failures(const int32_t *fail_idx, const int32_t *fail_wt, const int32_t *n_fail_d, int32_t *sorted_idx, cudaStream_t st)
{
    sort_failures_kernel<<<1, 1024, 0, st>>>(fail_idx, fail_wt, n_fail_d, sorted_idx);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int osd_launches_per_call(const qb_decoder *dec) { return osd_free_applicable(dec) ? 5 : 1; }   // select, tier A, select 2, tier B, full-width

int launch_osd0(qb_decoder *dec, const OsdLaunch &a, cudaStream_t st)
{
    if (a.F <= 0) return QB_OK;
    const GraphDev &g = dec->g;
    if (g.n > 65535 || g.m > 32 * 32 * OSD_MAX_WPL) {
        set_error("OSD-0 kernel supports n <= 65535 columns and m <= 4096 rows");
        return QB_ERR_UNSUPPORTED;
    }
    if (!a.ordering && !a.post) { set_error("OSD-0 needs posteriors or an ordering"); return QB_ERR_ARG; }
    const int wpl = ceil_div(g.mw, 32);
    if (!a.exact_rows && !a.ordering && osd_free_applicable(dec)) {
        // pipeline path: selection + free-row elimination (osd_free.cu); the full-width kernel below takes what that
        // hands back (more free rows / touched rows / candidates than it provides for), normally ~1 % of the sides
        int32_t *ovf_count = nullptr, *ovf_idx = nullptr;
        if (int rc = launch_osd0_free(dec, a, &ovf_count, &ovf_idx, st)) return rc;
        OsdLaunch b = a;
        b.fail_idx = ovf_idx; b.n_fail_d = ovf_count;
        switch (wpl) {
            case 1: return launch_osd_wpl<1, false>(dec, b, st);
            case 2: return launch_osd_wpl<2, false>(dec, b, st);
            case 3: return launch_osd_wpl<3, false>(dec, b, st);
            default: return launch_osd_wpl<4, false>(dec, b, st);
        }
    }
    if (a.exact_rows) {
        switch (wpl) {
            case 1: return launch_osd_wpl<1, true>(dec, a, st);
            case 2: return launch_osd_wpl<2, true>(dec, a, st);
            case 3: return launch_osd_wpl<3, true>(dec, a, st);
            default: return launch_osd_wpl<4, true>(dec, a, st);
        }
    }
    switch (wpl) {
        case 1: return launch_osd_wpl<1, false>(dec, a, st);
        case 2: return launch_osd_wpl<2, false>(dec, a, st);
        case 3: return launch_osd_wpl<3, false>(dec, a, st);
        default: return launch_osd_wpl<4, false>(dec, a, st);
    }
}

// ------------------------------------------------------------------------------------------------
// Dense GF(2) Gauss-Jordan on a bit-packed (uint32 words) m x n matrix with rhs, one CTA.
// Same sweep as gf2_elimination / gf2_elimination_packed_core (kernels.py:6-34, :49-96): column by
// column, pivot = first row >= current with the bit set, swap into place, clear the column elsewhere.
__global__ void __launch_bounds__(1024) gf2_dense_kernel(uint32_t *A, uint32_t *b, int m, int n, int nw,
                                                         int32_t *pivot_rows, int32_t *pivot_cols, int32_t *num_pivots)
{
    __shared__ int s_piv;
    const int tid = threadIdx.x;
    int row = 0, np = 0;
    for (int col = 0; col < n && row < m; ++col) {
        const int w = col >> 5;
        const uint32_t bit = 1u << (col & 31);
        if (tid == 0) s_piv = 0x7FFFFFFF;
        __syncthreads();
        for (int r = row + tid; r < m; r += blockDim.x)
            if (A[(size_t)r * nw + w] & bit) { atomicMin(&s_piv, r); break; }
        __syncthreads();
        const int pr = s_piv;
        __syncthreads();
        if (pr == 0x7FFFFFFF) continue;
        if (pr != row) {
            for (int k = tid; k < nw; k += blockDim.x) {
                const uint32_t x = A[(size_t)row * nw + k]; A[(size_t)row * nw + k] = A[(size_t)pr * nw + k]; A[(size_t)pr * nw + k] = x;
            }
            if (tid == 0) {
                const uint32_t br = (b[row >> 5] >> (row & 31)) & 1u, bp = (b[pr >> 5] >> (pr & 31)) & 1u;
                if (br != bp) { b[row >> 5] ^= 1u << (row & 31); b[pr >> 5] ^= 1u << (pr & 31); }
            }
            __syncthreads();
        }
        if (tid == 0) { pivot_rows[np] = row; pivot_cols[np] = col; }
        ++np;
        const uint32_t brow = (b[row >> 5] >> (row & 31)) & 1u;
        __syncthreads();
        // every warp clears the pivot column in a strided set of rows
        const int lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
        for (int r = warp; r < m; r += NW) {
            if (r == row) continue;
            if (A[(size_t)r * nw + w] & bit) {
                __syncwarp();
                for (int k = lane; k < nw; k += 32) A[(size_t)r * nw + k] ^= A[(size_t)row * nw + k];
                if (lane == 0 && brow) atomicXor(&b[r >> 5], 1u << (r & 31));
            }
        }
        ++row;
        __syncthreads();
    }
    if (tid == 0) *num_pivots = np;
}

int launch_gf2_dense(uint32_t *A, uint32_t *b, int m, int n, int nw, int32_t *pr, int32_t *pc, int32_t *np,
                     cudaStream_t st)
{
    gf2_dense_kernel<<<1, 1024, 0, st>>>(A, b, m, n, nw, pr, pc, np);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb

#ifdef QB_OSD_PROFILE
// copies the per-side records to the host and resets the counter; returns the number of records
extern "C" int qb_debug_osd_profile(unsigned long long *out_h, int max_records)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, qb::g_osd_prof_n, sizeof(int));
    n = std::min(std::min(n, qb::OSD_PROF_CAP), max_records);
    if (n > 0) cudaMemcpyFromSymbol(out_h, qb::g_osd_prof, sizeof(unsigned long long) * 8 * (size_t)n);
    const int zero = 0;
    cudaMemcpyToSymbol(qb::g_osd_prof_n, &zero, sizeof(int));
    return n;
}
#endif
