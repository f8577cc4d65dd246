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
#include <algorithm>

#include "common.cuh"

namespace qb {

struct OsdArgs {
    GraphDev g;
    OsdLaunch a;
    int sort_in_smem;     // keys / index ping-pong buffers in shared memory
    int tcap;             // T columns resident in shared memory; the rest spills to gT
    uint32_t *gT;         // [grid][(min(m,n) - tcap) * mw] spill
    uint32_t *gkeys;      // [grid][n] when !sort_in_smem
    uint16_t *gidx;       // [grid][2][n_pad] when !sort_in_smem
    int rank_cap;         // min(m, n)
};

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }

template <int WPL>
__global__ void __launch_bounds__(OSD_THREADS, 2) osd0_kernel(OsdArgs P)
{
    const GraphDev &g = P.g;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int mw = g.mw, n = g.n, m = g.m;
    const int n_pad2 = (n + 1) & ~1;

    // ---- shared memory carve-up -------------------------------------------------------------------
    unsigned char *sp = smem_raw;
    uint16_t *pos_of_row = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * g.m_pad;
    uint16_t *row_at_pos = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * g.m_pad;
    int16_t *pivcol_of_row = reinterpret_cast<int16_t *>(sp); sp += sizeof(int16_t) * g.m_pad;
    uint16_t *piv_row = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * g.m_pad;
    int32_t *piv_cand = reinterpret_cast<int32_t *>(sp); sp += sizeof(int32_t) * g.m_pad;
    uint32_t *npmask = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    uint32_t *sv = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    uint32_t *pv = reinterpret_cast<uint32_t *>(sp); sp += sizeof(uint32_t) * 32 * WPL;
    int *flags = reinterpret_cast<int *>(sp); sp += sizeof(int) * 32;
    uint16_t *idx0 = nullptr, *idx1 = nullptr;
    uint32_t *keys = nullptr, *cnt = nullptr;
    unsigned char *regionX;
    if (P.sort_in_smem) {
        idx0 = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * n_pad2;
        regionX = sp;                       // sort scratch, later overlaid by T
        keys = reinterpret_cast<uint32_t *>(regionX);
        idx1 = reinterpret_cast<uint16_t *>(regionX + sizeof(uint32_t) * n);
        cnt = reinterpret_cast<uint32_t *>(regionX + sizeof(uint32_t) * n + sizeof(uint16_t) * n_pad2);
    } else {
        regionX = sp;
        cnt = reinterpret_cast<uint32_t *>(regionX);     // 256*NW counters, overlaid by T afterwards
        keys = P.gkeys + (size_t)blockIdx.x * n;
        idx0 = P.gidx + (size_t)blockIdx.x * 2 * n_pad2;
        idx1 = idx0 + n_pad2;
    }
    uint32_t *Tsm = reinterpret_cast<uint32_t *>(regionX);
    uint32_t *Tgl = P.gT ? P.gT + (size_t)blockIdx.x * (size_t)(P.rank_cap - P.tcap) * mw : nullptr;
    __shared__ int s_rho;

    const int F = P.a.n_fail_d ? min(*P.a.n_fail_d, P.a.F) : P.a.F;

    for (int qi = blockIdx.x; qi < F; qi += gridDim.x) {
        const int shot = P.a.fail_idx ? P.a.fail_idx[qi] : qi;
        const uint32_t *hard = P.a.hard_bits + (size_t)shot * g.nw;
        const int32_t *ext_order = P.a.ordering ? P.a.ordering + (size_t)shot * n : nullptr;

        // ---- 1. residual syndrome, bookkeeping ----------------------------------------------------
        for (int w = tid; w < 32 * WPL; w += blockDim.x) {
            sv[w] = w < mw ? P.a.syn_bits[(size_t)shot * mw + w] : 0u;
            uint32_t full = 0u;
            if (w * 32 + 32 <= m) full = 0xFFFFFFFFu;
            else if (w * 32 < m) full = (1u << (m - w * 32)) - 1u;
            npmask[w] = full;
        }
        for (int r = tid; r < g.m_pad; r += blockDim.x) { pos_of_row[r] = (uint16_t)r; row_at_pos[r] = (uint16_t)r; pivcol_of_row[r] = -1; }
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

        // ---- 2. stable sort of columns by |posterior| -----------------------------------------------
        if (!ext_order) {
            const float *post = P.a.post + (size_t)shot * n;
            for (int j = tid; j < n; j += blockDim.x) { keys[j] = __float_as_uint(fabsf(post[j])); idx0[j] = (uint16_t)j; }
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
                {   // exclusive scan of cnt[256*NW] in (digit, warp) order
                    const int total = 256 * NW, per = (total + blockDim.x - 1) / blockDim.x;
                    const int b0 = tid * per;
                    uint32_t local = 0;
                    for (int i = b0; i < min(total, b0 + per); ++i) local += cnt[i];
                    uint32_t inc = local;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += y; }
                    __shared__ uint32_t wsum[32];
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
                    if (valid) {
                        const uint32_t base = cnt[d * NW + warp];
                        dst[base + __popc(peers & lanemask_lt())] = id;
                    }
                    __syncwarp();
                    if (valid && (peers & lanemask_lt()) == 0) cnt[d * NW + warp] += __popc(peers);
                    __syncwarp();
                }
                __syncthreads();
                uint16_t *tmp = src; src = dst; dst = tmp;
            }
            // after 4 passes the result is back in idx0
        }
        __syncthreads();

        // ---- 3. elimination ---------------------------------------------------------------------------
        auto order_at = [&](int c) -> int { return ext_order ? ext_order[c] : (int)idx0[c]; };
        auto Tcol = [&](int x) -> uint32_t * { return x < P.tcap ? Tsm + (size_t)x * mw : Tgl + (size_t)(x - P.tcap) * mw; };
        auto unresolved = [&]() -> bool {
            bool any = false;
#pragma unroll
            for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; any |= (sv[w] & npmask[w]) != 0u; }
            return __any_sync(0xFFFFFFFFu, any);
        };
        int t = 0;
        bool done = !unresolved();
        for (int c0 = 0; c0 < n && !done && t < P.rank_cap; c0 += NW) {
            const int c = c0 + warp;
            uint32_t v[WPL];
#pragma unroll
            for (int i = 0; i < WPL; ++i) v[i] = 0u;
            if (c < n) {
                const int j = order_at(c);
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) {
                    const int r = g.rowidx[p];
                    const int pc = pivcol_of_row[r];
                    if (pc >= 0) {
                        const uint32_t *col = Tcol(pc);
#pragma unroll
                        for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; if (w < mw) v[i] ^= col[w]; }
                    } else {
#pragma unroll
                        for (int i = 0; i < WPL; ++i) if ((r >> 5) == lane + 32 * i) v[i] ^= 1u << (r & 31);
                    }
                }
            }
            while (true) {
                bool f_ = false;
#pragma unroll
                for (int i = 0; i < WPL; ++i) f_ |= (v[i] & npmask[lane + 32 * i]) != 0u;
                const bool flag = __any_sync(0xFFFFFFFFu, f_);
                if (lane == 0) flags[warp] = flag ? 1 : 0;
                __syncthreads();
                const uint32_t fb = __ballot_sync(0xFFFFFFFFu, lane < NW && flags[lane] != 0);
                if (fb == 0u) break;
                const int f = __ffs(fb) - 1;
                if (warp == f) {
                    // pivot row = first row, in the reference's current (swapped) row order, with the bit set
                    uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) {
                        const int w = lane + 32 * i;
                        uint32_t bits = v[i] & npmask[w];
                        while (bits) {
                            const int b = __ffs(bits) - 1; bits &= bits - 1;
                            const int r = w * 32 + b;
                            best = min(best, ((uint32_t)pos_of_row[r] << 16) | (uint32_t)r);
                        }
                    }
                    for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
                    const int q = best >> 16, rho = best & 0xFFFF;
                    uint32_t *col = Tcol(t);
#pragma unroll
                    for (int i = 0; i < WPL; ++i) {
                        const int w = lane + 32 * i;
                        if (w < mw) {
                            col[w] = v[i];                                    // T.e_rho after this step
                            pv[w] = ((rho >> 5) == w) ? (v[i] & ~(1u << (rho & 31))) : v[i];
                        }
                        v[i] = 0u;
                    }
                    if (lane == 0) {
                        const int rt = row_at_pos[t];
                        row_at_pos[t] = (uint16_t)rho; row_at_pos[q] = (uint16_t)rt;
                        pos_of_row[rt] = (uint16_t)q; pos_of_row[rho] = (uint16_t)t;
                        pivcol_of_row[rho] = (int16_t)t;
                        piv_row[t] = (uint16_t)rho; piv_cand[t] = c;
                        npmask[rho >> 5] &= ~(1u << (rho & 31));
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
                    const uint32_t has = __shfl_sync(0xFFFFFFFFu, mine, rl) & rbit;
                    if (has) {
#pragma unroll
                        for (int i = 0; i < WPL; ++i) v[i] ^= u[i];
                    }
                }
                // (b) stored columns of earlier pivots and (c) the transformed syndrome (slot t)
                for (int x = warp; x <= t; x += NW) {
                    uint32_t *col = (x == t) ? sv : Tcol(x);
                    uint32_t cw[WPL];
#pragma unroll
                    for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; cw[i] = (w < mw) ? col[w] : 0u; }
                    uint32_t mine = 0u;
#pragma unroll
                    for (int i = 0; i < WPL; ++i) if (i == ri) mine = cw[i];
                    const uint32_t has = __shfl_sync(0xFFFFFFFFu, mine, rl) & rbit;
                    if (has) {
#pragma unroll
                        for (int i = 0; i < WPL; ++i) { const int w = lane + 32 * i; if (w < mw) col[w] = cw[i] ^ u[i]; }
                    }
                }
                ++t;
                __syncthreads();
                done = !unresolved();
                if (done || t >= P.rank_cap) break;
            }
            __syncthreads();
        }

        // ---- 5. solution = hard ^ e, e[ordering[pivot_col]] = s_reduced[pivot_row] ---------------
        uint32_t *hard_rw = P.a.hard_bits + (size_t)shot * g.nw;
        for (int i = tid; i < t; i += blockDim.x) {
            const int rho = piv_row[i];
            if ((sv[rho >> 5] >> (rho & 31)) & 1u) {
                const int j = order_at(piv_cand[i]);
                atomicXor(&hard_rw[j >> 5], 1u << (j & 31));
            }
        }
        if (P.a.pivots_out)
            for (int i = tid; i < P.rank_cap; i += blockDim.x)
                P.a.pivots_out[(size_t)shot * P.rank_cap + i] = i < t ? piv_cand[i] : -1;
        if (P.a.rank_out && tid == 0) P.a.rank_out[shot] = t;
        __syncthreads();
    }
}

template <int WPL>
static int launch_osd_wpl(qb_decoder *dec, const OsdLaunch &a, cudaStream_t st)
{
    const GraphDev &g = dec->g;
    OsdArgs P{};
    P.g = g; P.a = a;
    P.rank_cap = std::min(g.m, g.n);
    const int NW = OSD_THREADS / 32;
    const int n_pad2 = (g.n + 1) & ~1;
    const size_t fixed = sizeof(uint16_t) * 4 * (size_t)g.m_pad + sizeof(int32_t) * (size_t)g.m_pad +
                         sizeof(uint32_t) * 32 * WPL * 3 + sizeof(int) * 32 + 64;
    const size_t budget = (size_t)dec->max_smem_optin - 2048;   // static __shared__ + slack
    const size_t sort_x = sizeof(uint32_t) * (size_t)g.n + sizeof(uint16_t) * n_pad2 + sizeof(uint32_t) * 256 * NW;
    const size_t idx0_b = sizeof(uint16_t) * (size_t)n_pad2;
    size_t regionX, smem;
    const size_t colb = sizeof(uint32_t) * (size_t)g.mw;
    const size_t want = colb * (size_t)P.rank_cap;      // T with every possible pivot resident
    if (fixed + idx0_b + sort_x <= budget) {
        P.sort_in_smem = 1;
        const size_t avail1 = budget - fixed - idx0_b;                                   // 1 CTA / SM
        const size_t avail2 = budget / 2 > fixed + idx0_b ? budget / 2 - fixed - idx0_b : 0;   // 2 CTAs / SM
        regionX = std::max(sort_x, std::min(want, sort_x <= avail2 ? avail2 : avail1));
        smem = fixed + idx0_b + regionX;
    } else {
        P.sort_in_smem = 0;
        const size_t cnt_b = sizeof(uint32_t) * 256 * NW;
        regionX = std::max(cnt_b, std::min(want, (size_t)96 * 1024));
        smem = fixed + regionX;
    }
    QB_REQUIRE(smem <= budget + 2048, "OSD: problem too large for shared memory");
    P.tcap = (int)std::min<size_t>(P.rank_cap, regionX / colb);
    int grid = std::max(1, std::min(a.F, dec->sm_count * (smem <= budget / 2 ? 2 : 1)));
    size_t need = 0;
    const size_t spill = (size_t)(P.rank_cap - P.tcap) * g.mw * sizeof(uint32_t);
    const size_t gk = P.sort_in_smem ? 0 : sizeof(uint32_t) * (size_t)g.n;
    const size_t gi = P.sort_in_smem ? 0 : sizeof(uint16_t) * 2 * (size_t)n_pad2;
    need = (size_t)grid * (spill + gk + gi) + 256;
    if (int rc = dec->work.ensure(need)) return rc;
    unsigned char *base = dec->work.as<unsigned char>();
    P.gT = spill ? reinterpret_cast<uint32_t *>(base) : nullptr;
    P.gkeys = reinterpret_cast<uint32_t *>(base + (size_t)grid * spill);
    P.gidx = reinterpret_cast<uint16_t *>(base + (size_t)grid * (spill + gk));
    QB_CUDA(cudaFuncSetAttribute(osd0_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    osd0_kernel<WPL><<<grid, OSD_THREADS, smem, st>>>(P);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_osd0(qb_decoder *dec, const OsdLaunch &a, cudaStream_t st)
{
    if (a.F <= 0) return QB_OK;
    const GraphDev &g = dec->g;
    if (g.n > 65535 || g.m > 32 * 32 * OSD_MAX_WPL) {
        set_error("OSD-0 kernel supports n <= 65535 columns and m <= 4096 rows");
        return QB_ERR_UNSUPPORTED;
    }
    const int wpl = ceil_div(g.mw, 32);
    switch (wpl) {
        case 1: return launch_osd_wpl<1>(dec, a, st);
        case 2: return launch_osd_wpl<2>(dec, a, st);
        case 3: return launch_osd_wpl<3>(dec, a, st);
        default: return launch_osd_wpl<4>(dec, a, st);
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
