// K4+K5 (pipeline fast path): OSD-0 with ONE WARP per failed side.
//
// Same algorithm as osd.cu (reference: performOSD_enhanced order 0, src/decoding/osd.py:5-29, on top of
// gf2_elimination_packed_core, src/decoding/kernels.py:49-96; row transform T restricted to pivot rows, exact early
// termination), for the shapes of the Monte-Carlo pipeline: m <= 1024 check rows, so that a GF(2) vector over the
// rows is exactly one 32-bit register per lane (lane = word) and every vector operation is one warp instruction.
// The four-warp kernel of osd.cu spends ~35x more instructions on its per-pivot protocol (flags, barriers, ballots
// across warps) than on the bit-vector work itself; here a side is a strictly sequential program of one warp --
// no barrier, no cross-warp state -- and the SM is filled with independent sides (one 32-thread CTA each).
//
//   1. residual syndrome s = syndrome ^ H.hard                                     (osd.py:7-9)
//   2. candidates in ascending |posterior| (ties by column index): 256-bin histogram of the float bit patterns,
//      a window of the <= 512 least reliable columns, bitonic sort of (key << 16 | column) in shared memory;
//      further windows only if the elimination is not finished                       (osd.py:11-12)
//   3. per candidate: v = XOR of the stored columns of its pivot rows / unit vectors of its free rows; pivot iff v
//      has a bit on a free row (lowest such row); stored columns with that bit are updated      (kernels.py:60-96)
//   4. stop when the transformed syndrome has no bit on a free row; solution bits = transformed syndrome on the pivot
//      rows, applied to hard                                                        (osd.py:19-25)
// Sides whose first histogram bin alone exceeds the window (never seen on the BB codes) are handed to osd.cu.
//
// STATUS: opt-in (QLDPC_B200_OSD_WARP=1), not the default.  Measured on B200, gross code, 65536 shots/step (124k sides):
// 41-65 ms per step against 35.4 ms for the four-warp kernel, although it executes 36 % fewer instructions
// (1.75e9 vs 2.72e9 per 16384 shots).  A side is a chain of dependent instructions (one issue per ~16-26 cycles per
// warp: shared-memory loads, shuffles, branches, instruction fetch -- 23 % of stall samples are no_inst because every
// warp of the SM sits in a different part of the 3.8k-instruction program); hiding that needs ~40 warps per SM, and
// the 18-20 KB of stored columns per side allow 10 (or 16-24 with most columns spilled to global memory, which
// then shows up as long-scoreboard stalls).  profiles/r1b_microbenchmarks.txt has the numbers.
// Way out: an elimination touches only ~1.1 rows per pivot (141 of 1008 on average), so with rows renumbered on
// first touch the stored columns of a side take ~2-8 KB instead of 18-20 KB and 40+ sides fit an SM.
#include <algorithm>

#include "common.cuh"

namespace qb {

constexpr int OW_BINS = 256;        // (key >> 21) - OW_BASE clamped: 4 bins per octave, 2^-57 .. 2^6
constexpr int OW_BASE = 277;
constexpr int OW_WIN = 512;         // candidates per window (bitonic sort size)
constexpr int OW_MIN = 320;         // a window is closed once it holds at least this many
constexpr int OW_G = 4;             // candidates reduced together

struct OsdWarpArgs {
    GraphDev g;
    const uint32_t *syn_bits;    // [B][mw]
    uint32_t *hard_bits;         // [B][nw] in/out
    const float *post;           // [B][n]
    const int32_t *fail_idx;     // failure queue (heaviest first)
    const int32_t *n_fail_d;
    int F_max;
    int tcap;                    // T columns resident in shared memory per side
    uint32_t *gT;                // [grid][(rank_cap - tcap) * 32] spill
    uint32_t *g_piv;             // [grid][rank_cap]  pivot row | column << 16
    unsigned long long *g_list;  // [grid][OW_WIN] sort scratch of later windows
    int32_t *work_counter;       // zero at launch
    int32_t *overflow_count;     // sides handed to the four-warp kernel
    int32_t *overflow_idx;
    int rank_cap;
};

__device__ __forceinline__ uint32_t ow_bin(uint32_t key)
{
    const int b = (int)(key >> 21) - OW_BASE;
    return (uint32_t)min(max(b, 0), OW_BINS - 1);
}

// bitonic sort of OW_WIN 64-bit keys by one warp (ascending)
__device__ __forceinline__ void ow_sort(unsigned long long *a, int lane)
{
    for (int k = 2; k <= OW_WIN; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < OW_WIN; i += 32) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncwarp();
        }
}

__global__ void __launch_bounds__(32) osd0_warp_kernel(OsdWarpArgs P)
{
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int n = g.n, mw = g.mw;
    // shared memory: pivot column of every row, current window (sorted columns), histogram, T columns (stride 33)
    int16_t *pivcol = reinterpret_cast<int16_t *>(smem_raw);                       // [1024]
    uint16_t *ord = reinterpret_cast<uint16_t *>(pivcol + 1024);                   // [OW_WIN]
    uint32_t *hist = reinterpret_cast<uint32_t *>(ord + OW_WIN);                   // [OW_BINS]
    uint32_t *sv_sm = hist + OW_BINS;                                              // [32]
    uint32_t *Tsm = sv_sm + 32;                                                    // [tcap][33]; first window: sort scratch
    uint32_t *Tgl = P.gT + (size_t)blockIdx.x * (size_t)max(0, P.rank_cap - P.tcap) * 32;
    uint32_t *piv = P.g_piv + (size_t)blockIdx.x * P.rank_cap;
    const int F = min(*P.n_fail_d, P.F_max);

    while (true) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(P.work_counter, 1);
        qi = __shfl_sync(0xFFFFFFFFu, qi, 0);
        if (qi >= F) break;
        const int shot = P.fail_idx[qi];
        uint32_t *hard = P.hard_bits + (size_t)shot * g.nw;
        const float *post = P.post + (size_t)shot * n;

        // ---- 1. residual syndrome (lane = word), bookkeeping ------------------------------------------------
        uint32_t sv = lane < mw ? P.syn_bits[(size_t)shot * mw + lane] : 0u;
        uint32_t np;                                                               // free (non-pivot) rows
        if (lane * 32 + 32 <= g.m) np = 0xFFFFFFFFu; else if (lane * 32 < g.m) np = (1u << (g.m - lane * 32)) - 1u; else np = 0u;
        for (int r = lane; r < 1024; r += 32) pivcol[r] = -1;
        for (int b = lane; b < OW_BINS; b += 32) hist[b] = 0u;
        sv_sm[lane] = 0u;
        __syncwarp();
        for (int w0 = 0; w0 < g.nw; w0 += 32) {                          // every lane walks the set bits of one word of hard
            const int w = w0 + lane;
            uint32_t bits = w < g.nw ? hard[w] : 0u;
            while (__any_sync(0xFFFFFFFFu, bits != 0u)) {
                int j = -1;
                if (bits) { j = w * 32 + __ffs(bits) - 1; bits &= bits - 1; }
                if (j >= 0 && j < n) {
                    const uint4 sg = g.colsig[j];
                    const uint32_t rr[8] = {sg.x & 0xFFFFu, sg.x >> 16, sg.y & 0xFFFFu, sg.y >> 16, sg.z & 0xFFFFu, sg.z >> 16, sg.w & 0xFFFFu, sg.w >> 16};
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (rr[k] != 0xFFFFu) atomicXor(&sv_sm[rr[k] >> 5], 1u << (rr[k] & 31));
                }
            }
        }
        __syncwarp();
        sv ^= sv_sm[lane];
        __syncwarp();
        // ---- 2a. histogram of the |posterior| bit patterns ----------------------------------------------------
        for (int j0 = lane; j0 < n; j0 += 8 * 32) {
            uint32_t kb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32; kb[u] = j < n ? __float_as_uint(fabsf(post[j])) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (kb[u] != 0xFFFFFFFFu) atomicAdd(&hist[ow_bin(kb[u])], 1u);
        }
        __syncwarp();

        int t = 0, bin_next = 0;
        bool done = !__any_sync(0xFFFFFFFFu, (sv & np) != 0u);
        bool overflow = false;
        while (!done && t < P.rank_cap && bin_next < OW_BINS) {
            // ---- 2b. next window: bins [bin_next, bin_hi] with >= OW_MIN and <= OW_WIN columns ------------------
            int cum = 0, bin_hi = bin_next - 1;
            {
                bool stop = false;
                for (int b0 = bin_next; b0 < OW_BINS && !stop; b0 += 32) {
                    const int b = b0 + lane;
                    const int cnt = b < OW_BINS ? (int)hist[b] : 0;
                    int inc = cnt;
                    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
                    const uint32_t over = __ballot_sync(0xFFFFFFFFu, cum + inc > OW_WIN);
                    const uint32_t enough = __ballot_sync(0xFFFFFFFFu, cum + inc >= OW_MIN);
                    int last = 31;
                    if (over | enough) {
                        const int fo = over ? __ffs(over) - 1 : 32, fe = enough ? __ffs(enough) - 1 : 32;
                        last = fe < fo ? fe : fo - 1;
                        stop = true;
                    }
                    if (last >= 0) { cum += __shfl_sync(0xFFFFFFFFu, inc, last); bin_hi = min(OW_BINS - 1, b0 + last); }
                }
                if (!stop) bin_hi = OW_BINS - 1;
            }
            if (cum == 0) {
                if (bin_hi >= OW_BINS - 1) break;                       // no candidates left
                overflow = true; break;                                 // one bin alone exceeds the window
            }
            const int M = cum;
            unsigned long long *list = t == 0 ? reinterpret_cast<unsigned long long *>(Tsm)
                                              : P.g_list + (size_t)blockIdx.x * OW_WIN;
            for (int i = lane; i < OW_WIN; i += 32) list[i] = ~0ull;
            __syncwarp();
            int base = 0;
            for (int j0 = 0; j0 < n; j0 += 8 * 32) {
                uint32_t kb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; kb[u] = j < n ? __float_as_uint(fabsf(post[j])) : 0xFFFFFFFFu; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u * 32 + lane;
                    const int b = (int)ow_bin(kb[u]);
                    const bool in = j < n && b >= bin_next && b <= bin_hi;
                    const uint32_t msk = __ballot_sync(0xFFFFFFFFu, in);
                    if (in) list[base + __popc(msk & ((1u << lane) - 1u))] = ((unsigned long long)kb[u] << 16) | (unsigned long long)j;
                    base += __popc(msk);
                }
            }
            __syncwarp();
            ow_sort(list, lane);
            for (int i = lane; i < M; i += 32) ord[i] = (uint16_t)(list[i] & 0xFFFFull);
            __syncwarp();
            bin_next = bin_hi + 1;

            // ---- 3. candidates of the window, in order, OW_G at a time: the reductions of a group are independent
            // (OW_G x 6 stored-column loads in flight), then its pivots are resolved one after the other with the later
            // members of the group corrected in registers (a row that became a pivot row meanwhile shows up as a set bit)
            for (int c0 = 0; c0 < M && !done && t < P.rank_cap; c0 += OW_G) {
                uint32_t v[OW_G];
                int jj[OW_G];
#pragma unroll
                for (int q = 0; q < OW_G; ++q) {
                    v[q] = 0u; jj[q] = -1;
                    if (c0 + q < M) {
                        jj[q] = ord[c0 + q];
                        const uint4 sg = g.colsig[jj[q]];
                        const uint32_t rr[8] = {sg.x & 0xFFFFu, sg.x >> 16, sg.y & 0xFFFFu, sg.y >> 16, sg.z & 0xFFFFu, sg.z >> 16, sg.w & 0xFFFFu, sg.w >> 16};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (rr[k] == 0xFFFFu) continue;                     // uniform
                            const int pc = pivcol[rr[k]];                       // uniform (broadcast)
                            if (pc >= 0) v[q] ^= pc < P.tcap ? Tsm[pc * 33 + lane] : Tgl[(size_t)(pc - P.tcap) * 32 + lane];
                            else if ((int)(rr[k] >> 5) == lane) v[q] ^= 1u << (rr[k] & 31);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < OW_G; ++q) {
                    if (jj[q] < 0 || done || t >= P.rank_cap) continue;         // uniform
                    const uint32_t fb = v[q] & np;
                    uint32_t best = fb ? (uint32_t)(lane * 32 + __ffs(fb) - 1) : 0xFFFFFFFFu;
                    best = __reduce_min_sync(0xFFFFFFFFu, best);
                    if (best == 0xFFFFFFFFu) continue;                          // dependent on earlier columns
                    const int rho = (int)best, rl = rho >> 5;
                    const uint32_t rbit = 1u << (rho & 31);
                    if (t < P.tcap) Tsm[t * 33 + lane] = v[q]; else Tgl[(size_t)(t - P.tcap) * 32 + lane] = v[q];
                    const uint32_t u = lane == rl ? (v[q] & ~rbit) : v[q];
                    if (__shfl_sync(0xFFFFFFFFu, sv, rl) & rbit) sv ^= u;
                    if (lane == rl) np &= ~rbit;
                    if (lane == 0) { pivcol[rho] = (int16_t)t; piv[t] = (uint32_t)rho | ((uint32_t)jj[q] << 16); }
#pragma unroll
                    for (int q2 = q + 1; q2 < OW_G; ++q2)
                        if (__shfl_sync(0xFFFFFFFFu, v[q2], rl) & rbit) v[q2] ^= u;
                    // stored columns of earlier pivots with a bit on row rho
                    for (int x0 = 0; x0 < t; x0 += 32) {
                        const int x = x0 + lane;
                        uint32_t wv = 0u;
                        if (x < t) wv = x < P.tcap ? Tsm[x * 33 + rl] : Tgl[(size_t)(x - P.tcap) * 32 + rl];
                        uint32_t msk = __ballot_sync(0xFFFFFFFFu, (wv & rbit) != 0u);
                        while (msk) {                                           // four columns per step: independent read-modify-writes
                            int xb[4];
                            uint32_t val[4];
#pragma unroll
                            for (int q3 = 0; q3 < 4; ++q3) { xb[q3] = msk ? x0 + __ffs(msk) - 1 : -1; msk &= msk - 1; }
#pragma unroll
                            for (int q3 = 0; q3 < 4; ++q3)
                                if (xb[q3] >= 0) val[q3] = xb[q3] < P.tcap ? Tsm[xb[q3] * 33 + lane] : Tgl[(size_t)(xb[q3] - P.tcap) * 32 + lane];
#pragma unroll
                            for (int q3 = 0; q3 < 4; ++q3)
                                if (xb[q3] >= 0) { if (xb[q3] < P.tcap) Tsm[xb[q3] * 33 + lane] = val[q3] ^ u; else Tgl[(size_t)(xb[q3] - P.tcap) * 32 + lane] = val[q3] ^ u; }
                        }
                    }
                    __syncwarp();
                    ++t;
                    done = !__any_sync(0xFFFFFFFFu, (sv & np) != 0u);
                }
            }
        }
        if (overflow) {                                                 // (never on the BB codes) leave it to the four-warp kernel
            if (lane == 0) { const int s = atomicAdd(P.overflow_count, 1); P.overflow_idx[s] = shot; }
            continue;
        }
        // ---- 4. solution = hard ^ e, e[pivot column] = transformed syndrome on the pivot row ------------------------
        sv_sm[lane] = sv;
        __syncwarp();
        for (int i = lane; i < t; i += 32) {
            const uint32_t pr = piv[i];
            const int rho = pr & 0xFFFF, j = pr >> 16;
            if ((sv_sm[rho >> 5] >> (rho & 31)) & 1u) atomicXor(&hard[j >> 5], 1u << (j & 31));
        }
        __syncwarp();
    }
}

// Returns QB_OK and sets *used = 1 when the warp kernel was launched (sides it could not take are listed in
// overflow_idx_d / overflow_count_d for the four-warp kernel).
int launch_osd0_warp(qb_decoder *dec, const OsdLaunch &a, int32_t *overflow_count_d, int32_t *overflow_idx_d, int *used, cudaStream_t st)
{
    *used = 0;
    const GraphDev &g = dec->g;
    if (a.exact_rows || a.ordering || !a.post || !a.fail_idx || !a.n_fail_d || a.rank_out || a.pivots_out) return QB_OK;
    if (!g.colsig || g.m > 1024 || g.n > 65535 || g.m < 1) return QB_OK;
    const char *opt = getenv("QLDPC_B200_OSD_WARP");
    if (!opt || opt[0] != '1') return QB_OK;
    OsdWarpArgs P{};
    P.g = g; P.syn_bits = a.syn_bits; P.hard_bits = a.hard_bits; P.post = a.post; P.fail_idx = a.fail_idx; P.n_fail_d = a.n_fail_d;
    P.F_max = a.F;
    P.rank_cap = std::min(g.m, g.n);
    int sides_per_sm = 24;
    if (const char *e = getenv("QLDPC_B200_OSD_WARP_SIDES")) { const int v = atoi(e); if (v >= 1 && v <= 32) sides_per_sm = v; }
    const size_t fixed = 1024 * 2 + OW_WIN * 2 + OW_BINS * 4 + 32 * 4;
    const size_t per = ((size_t)dec->max_smem_optin + 1024) / sides_per_sm - 1024;
    if (per < fixed + (size_t)OW_WIN * 8) return QB_OK;
    P.tcap = (int)std::min<size_t>(P.rank_cap, (per - fixed) / (33 * 4));
    const size_t smem = fixed + std::max((size_t)P.tcap * 33 * 4, (size_t)OW_WIN * 8);
    const int grid = std::max(1, std::min(a.F, dec->sm_count * sides_per_sm));
    const size_t spill = (size_t)std::max(0, P.rank_cap - P.tcap) * 32 * 4;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t G = (size_t)grid;
    const size_t need = 256 + al(G * spill) + al(G * (size_t)P.rank_cap * 4) + al(G * OW_WIN * 8) + 256;
    if (int rc = dec->work.ensure(need)) return rc;
    unsigned char *p = dec->work.as<unsigned char>();
    P.work_counter = reinterpret_cast<int32_t *>(p); p += 256;
    QB_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(int32_t), st));
    QB_CUDA(cudaMemsetAsync(overflow_count_d, 0, sizeof(int32_t), st));
    P.gT = reinterpret_cast<uint32_t *>(p); p += al(G * spill);
    P.g_piv = reinterpret_cast<uint32_t *>(p); p += al(G * (size_t)P.rank_cap * 4);
    P.g_list = reinterpret_cast<unsigned long long *>(p);
    P.overflow_count = overflow_count_d; P.overflow_idx = overflow_idx_d;
    QB_CUDA(cudaFuncSetAttribute(osd0_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    osd0_warp_kernel<<<grid, 32, smem, st>>>(P);
    QB_CUDA(cudaGetLastError());
    *used = 1;
    return QB_OK;
}

}  // namespace qb
