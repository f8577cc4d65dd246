// K3: batched flooding min-sum (and the rarely used general / tanh-BP / single-pass variants).
//
// Reference semantics: src/decoding/kernels.py:235-366 (minsum_decoder_full), :370-485
// (autoregressive alpha), :139-169 (minsum_core_sparse), :172-193 + dense.py:75-96 (tanh BP).
//
// Fast kernel (damping == 1, which is every call the engine makes, engine.py:84-88):
//   * one CTA owns S shots of one side; all per-shot state lives in shared memory:
//       posterior values  float [n_pad][S]           (shot-interleaved -> one LDS.128 per edge for S=4)
//       check state       uint4 [S][m_pad]           {alpha*min1, alpha*min2, signs 0..31,
//                                                      signs 32..56 | argmin<<25 | total sign<<31}
//     Because damping == 1 the variable-to-check message is  clip(value[j] - R_old(edge)) and
//     R_old is recomputed from the compressed check state, so no per-edge message is stored.
//   * phase A: one thread per check row walks the row's sliced-ELL column list (coalesced uint16
//     loads, the graph is shared by the S shots) and produces the new compressed state;
//   * phase B: one thread per variable gathers its <=6 check states in row order (same summation
//     order as the reference), adds the prior, and XORs its checks' parity bits when the hard
//     decision is 1, which yields the convergence test without another pass over the graph.
// IEEE inf/NaN behaviour is kept (no fast-math): degree-1 checks produce +-inf messages and
// inf-inf = NaN -> 0 exactly like kernels.py:327-329.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "edge_dev.cuh"

namespace qb {

// ------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void load_vals(const float *vals, uint32_t j, float (&v)[S])
{
    if constexpr (S == 4) {
        const float4 f = *reinterpret_cast<const float4 *>(vals + j * 4);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else if constexpr (S == 2) {
        const float2 f = *reinterpret_cast<const float2 *>(vals + j * 2);
        v[0] = f.x; v[1] = f.y;
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s) v[s] = vals[j * S + s];
    }
}

// Shared-memory extents: one extra "dummy" variable (index n_pad, value +inf) and one extra dummy check
// (index m_pad, all-zero state) absorb the ELL padding, so the inner loops carry no validity tests:
// a +inf variable never lowers a minimum and has sign '+', a zero check state contributes R = +0.
__host__ __device__ inline int vals_rows(const GraphDev &g) { return g.n_pad + 4; }
__host__ __device__ inline int chk_rows(const GraphDev &g) { return g.m_pad + 4; }

// Check state (uint4 per check and shot):
//   x = alpha*min1, y = alpha*min2 (float bits)
//   z, w = sign of R per row position (already multiplied by the row's total sign); position p lives at bit
//          p ^ 7 of the 64-bit word w:z (byte = chunk of 8 edges, newest edge in the low bit: the bits are
//          collected with one funnel shift per edge); w bits 24..29 = argmin position, so row degree <= 56
//
// One row slice (32 check rows, one per lane) for S shots.  EXACT = the slice can see inf/NaN (a row of
// degree 1 gives min2 = inf -> +-inf messages, and inf - inf = NaN -> 0, kernels.py:327-329; or non-finite
// priors): that variant keeps the explicit NaN test, compares instead of using sign bits and skips padding
// explicitly.  Everywhere else |q| <= clip is finite, q is never NaN or -0.0, sign tests reduce to moving the
// IEEE sign bit, and padding entries point at the +inf dummy variable.
template <int S, bool EXACT>
__device__ __forceinline__ void row_slice(const float *vals, uint4 *chk, const uint32_t *syn, const GraphDev &g,
                                          const uint4 *ell, uint4 cur, int deg, int r, int it, float alpha, float clip)
{
    const int crow = chk_rows(g);
    const uint32_t dummy = (uint32_t)g.n_pad;
    uint32_t o1[S], o2[S], oz[S], ow[S];
    int oam[S];
    float mn1[S], mn2[S];
    uint32_t nz[S], nw_[S];
    int am[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        mn1[s] = INFINITY; mn2[s] = INFINITY; nz[s] = 0u; nw_[s] = 0u; am[s] = 0;
        if (it > 0) {
            const uint4 st = chk[s * crow + r];
            o1[s] = st.x; o2[s] = st.y; oz[s] = st.z; ow[s] = st.w; oam[s] = (int)(st.w >> 24);
        } else {
            o1[s] = 0u; o2[s] = 0u; oz[s] = 0u; ow[s] = 0u; oam[s] = -1;
        }
    }
    const int nch = (deg + 7) >> 3;
    for (int c = 0; c < nch; ++c) {
        const uint4 nxt = ell[(c + 1 < nch ? c + 1 : c) * 32];         // prefetch next chunk
        const uint32_t idx[8] = {cur.x & 0xFFFFu, cur.x >> 16, cur.y & 0xFFFFu, cur.y >> 16,
                                 cur.z & 0xFFFFu, cur.z >> 16, cur.w & 0xFFFFu, cur.w >> 16};
        uint32_t owc[S], nsc[S], pm[S];
        int dlt[S];
        const int sh = (c & 3) * 8;
        const int nvalid = min(8, deg - c * 8);                          // uniform over the warp
#pragma unroll
        for (int s = 0; s < S; ++s) {
            owc[s] = (c < 4 ? oz[s] : ow[s]) >> sh;
            dlt[s] = oam[s] - c * 8;
            nsc[s] = 0u; pm[s] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nvalid) {
                const uint32_t j = idx[i];
                if (!EXACT || j != dummy) {
                    float v[S];
                    load_vals<S>(vals, j, v);
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const uint32_t omag = (dlt[s] == i) ? o2[s] : o1[s];
                        const float rold = __uint_as_float(omag ^ ((owc[s] << (24 + i)) & 0x80000000u));
                        float q = v[s] - rold;
                        uint32_t qb;
                        if constexpr (EXACT) {
                            q = (q != q) ? 0.f : q;                            // kernels.py:328-329
                            qb = (q < 0.f) ? 0x80000000u : 0u;                 // val >= 0 -> '+', kernels.py:296
                        } else {
                            qb = __float_as_uint(q);
                        }
                        // clip(q) (kernels.py:330-333) is monotone in |q| and keeps the sign, so the two smallest
                        // |clip(q)| are min(., clip) of the two smallest |q|: the clamp is applied once per row below
                        // (rows of degree 1 keep min2 = +inf, so the EXACT variant clamps per edge like the reference)
                        const float ab = EXACT ? fminf(fabsf(q), clip) : fabsf(q);
                        nsc[s] = __funnelshift_l(qb, nsc[s], 1);
                        uint32_t lt;                                           // sign bit set iff ab < min1 (strict: first minimum wins)
                        if constexpr (EXACT) lt = (ab < mn1[s]) ? 0x80000000u : 0u;
                        else lt = __float_as_uint(ab - mn1[s]);
                        pm[s] = __funnelshift_l(lt, pm[s], 1);
                        mn2[s] = fminf(mn2[s], fmaxf(ab, mn1[s]));
                        mn1[s] = fminf(mn1[s], ab);
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) { nsc[s] <<= 1; pm[s] <<= 1; }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            nsc[s] <<= (8 - nvalid); pm[s] <<= (8 - nvalid);            // entry i sits at bit 7 - i of the chunk byte
            if (pm[s]) am[s] = c * 8 + 8 - __ffs(pm[s]);               // last edge of the chunk that lowered min1
            if (c < 4) nz[s] |= nsc[s] << sh; else nw_[s] |= nsc[s] << sh;
        }
        cur = nxt;
    }
    if (r < g.m) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const uint32_t sbit = (syn[s * g.mw + (r >> 5)] >> (r & 31)) & 1u;
            const uint32_t tot = (sbit ^ (uint32_t)(__popc(nz[s]) + __popc(nw_[s]))) & 1u;
            const uint32_t tmask = 0u - tot;
            uint4 st;
            st.x = __float_as_uint(alpha * (EXACT ? mn1[s] : fminf(mn1[s], clip)));
            st.y = __float_as_uint(alpha * (EXACT ? mn2[s] : fminf(mn2[s], clip)));
            st.z = nz[s] ^ tmask;
            st.w = ((nw_[s] ^ tmask) & 0x00FFFFFFu) | ((uint32_t)am[s] << 24);
            chk[s * crow + r] = st;
        }
    }
}

// R messages of one chunk (4 column entries) of the lane's variable, added in row order (kernels.py:316).
template <int S>
__device__ __forceinline__ void col_chunk(const uint4 *chk, int crow, uint4 e4, int nvalid, float (&acc)[S])
{
    const uint32_t ent[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < nvalid) {                                                // uniform over the warp
            const uint32_t e = ent[i];
            const uint32_t cidx = e >> 8, pos = e & 63u;
            const bool hi = pos >= 32u;
            const uint32_t shl = 31u - ((pos ^ 7u) & 31u);               // position p -> bit p ^ 7 (see row_slice)
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const uint4 st = chk[s * crow + cidx];
                const uint32_t mag = ((st.w >> 24) == pos) ? st.y : st.x;
                const uint32_t word = hi ? st.w : st.z;
                acc[s] += __uint_as_float(mag ^ ((word << shl) & 0x80000000u));
            }
        }
    }
}

template <int S>
__global__ void __launch_bounds__((S == 1 ? 1024 : MS_THREADS), (S == 2 ? 2 : 1))
minsum_fast_kernel(GraphDev g, MinsumLaunch a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int vrow = vals_rows(g), crow = chk_rows(g);
    float *vals = reinterpret_cast<float *>(smem_raw);                               // [vrow][S]
    uint4 *chk = reinterpret_cast<uint4 *>(smem_raw + (size_t)vrow * S * sizeof(float));     // [S][crow]
    uint32_t *syn = reinterpret_cast<uint32_t *>(chk + (size_t)S * crow);           // [S][mw]
    uint32_t *par = syn + S * g.mw;                                                  // [S][mw]
    int *s_rptr = reinterpret_cast<int *>(par + S * g.mw);                           // slice tables, staged once
    int *s_cptr = s_rptr + g.n_rslices;
    int *s_rdeg = s_cptr + g.n_cslices;                                              // degree | exact-path flag << 8
    int *s_cdeg = s_rdeg + g.n_rslices;
    __shared__ int s_unsat[S];
    __shared__ int s_wt[S];
    __shared__ int s_active[S];
    __shared__ int s_nactive;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n_tiles = (a.B + S - 1) / S;
    for (int i = tid; i < g.n_rslices; i += blockDim.x) {
        s_rptr[i] = g.rslice_ptr[i];
        s_rdeg[i] = (int)g.rslice_deg[i] | ((g.nan_anywhere | (int)g.rslice_exact[i]) << 8);
    }
    for (int i = tid; i < g.n_cslices; i += blockDim.x) { s_cptr[i] = g.cslice_ptr[i]; s_cdeg[i] = (int)g.cslice_deg[i]; }
    // dummy variable (+inf) and dummy check (zero state)
    if (tid < S) { vals[g.n_pad * S + tid] = INFINITY; chk[tid * crow + g.m_pad] = make_uint4(0u, 0u, 0u, 0u); }

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int shot0 = tile * S;
        // ---- load: values = prior, syndrome words, parity = 0 ---------------------------------
        for (int j = tid; j < g.n_pad; j += blockDim.x) {
            float p = j < g.n ? g.prior[j] : 0.f;
#pragma unroll
            for (int s = 0; s < S; ++s) vals[j * S + s] = p;
        }
        for (int i = tid; i < S * g.mw; i += blockDim.x) {
            int s = i / g.mw, w = i - s * g.mw;
            syn[i] = (shot0 + s < a.B) ? a.syn_bits[(size_t)(shot0 + s) * g.mw + w] : 0u;
            par[i] = 0u;
        }
        if (tid < S) { s_active[tid] = (shot0 + tid < a.B); s_unsat[tid] = 0; s_wt[tid] = 0; }
        if (tid == 0) s_nactive = min(S, a.B - shot0);
        __syncthreads();
        if (a.max_iter <= 0) {
            // reference returns all-zero candidate, converged False, final_iter -1 (kernels.py:267)
            for (int s = 0; s < S; ++s) {
                if (shot0 + s >= a.B) break;
                size_t shot = shot0 + s;
                for (int w = tid; w < g.nw; w += blockDim.x) a.hard_bits[shot * g.nw + w] = 0u;
                if (a.post) for (int j = tid; j < g.n; j += blockDim.x) a.post[shot * g.n + j] = 0.f;
                if (tid == 0) {
                    a.converged[shot] = 0; a.final_iter[shot] = -1;
                    if (a.fail_count) { const int slot = atomicAdd(a.fail_count, 1); a.fail_idx[slot] = (int)shot; if (a.fail_wt) a.fail_wt[slot] = 0; }
                }
            }
            __syncthreads();
            continue;
        }

        for (int it = 0; it < a.max_iter; ++it) {
            const float alpha = a.alpha_d[it];
            // iteration 0 uses Q = prior unclipped (kernels.py:263-265): old state = 0, clip = inf
            const float clip = it > 0 ? a.clip : INFINITY;

            // ---- phase A: check rows ----------------------------------------------------------
            uint4 rfirst = make_uint4(0u, 0u, 0u, 0u);
            if (warp < g.n_rslices && (s_rdeg[warp] & 255)) rfirst = g.row_ell4[s_rptr[warp] + lane];
            for (int rs = warp; rs < g.n_rslices; rs += nwarps) {
                const int base = s_rptr[rs];
                const int dg = s_rdeg[rs];
                const uint4 cur0 = rfirst;
                {   // prefetch the first chunk of this warp's next slice
                    const int nrs = rs + nwarps;
                    if (nrs < g.n_rslices && (s_rdeg[nrs] & 255)) rfirst = g.row_ell4[s_rptr[nrs] + lane];
                }
                if ((dg & 255) == 0) continue;
                const uint4 *ell = g.row_ell4 + base + lane;
                if (dg >> 8) row_slice<S, true>(vals, chk, syn, g, ell, cur0, dg & 255, rs * 32 + lane, it, alpha, clip);
                else row_slice<S, false>(vals, chk, syn, g, ell, cur0, dg & 255, rs * 32 + lane, it, alpha, clip);
            }
            __syncthreads();

            // ---- phase B: variables ------------------------------------------------------------
            // the first two chunks (8 entries) and the priors of a slice are fetched one slice ahead
            uint4 cf0 = make_uint4(0u, 0u, 0u, 0u), cf1 = cf0;
            float pfirst = 0.f;
            if (warp < g.n_cslices) {
                const int d0 = s_cdeg[warp];
                if (d0 > 0) cf0 = g.col_ell4[s_cptr[warp] + lane];
                if (d0 > 4) cf1 = g.col_ell4[s_cptr[warp] + 32 + lane];
                if (warp * 32 + lane < g.n) pfirst = g.prior[warp * 32 + lane];
            }
            for (int cs = warp; cs < g.n_cslices; cs += nwarps) {
                const int base = s_cptr[cs];
                const int deg = s_cdeg[cs];
                const uint4 *ell = g.col_ell4 + base + lane;
                const int j = cs * 32 + lane;
                const uint4 e0 = cf0, e1 = cf1;
                const float pr = pfirst;
                {
                    const int ncs = cs + nwarps;
                    if (ncs < g.n_cslices) {
                        const int d1 = s_cdeg[ncs];
                        if (d1 > 0) cf0 = g.col_ell4[s_cptr[ncs] + lane];
                        if (d1 > 4) cf1 = g.col_ell4[s_cptr[ncs] + 32 + lane];
                        pfirst = (ncs * 32 + lane < g.n) ? g.prior[ncs * 32 + lane] : 0.f;
                    }
                }
                float acc[S];
#pragma unroll
                for (int s = 0; s < S; ++s) acc[s] = 0.f;
                if (deg > 0) col_chunk<S>(chk, crow, e0, min(4, deg), acc);
                if (deg > 4) col_chunk<S>(chk, crow, e1, min(4, deg - 4), acc);
                for (int c = 2; c * 4 < deg; ++c) col_chunk<S>(chk, crow, ell[c * 32], min(4, deg - c * 4), acc);
                if (j < g.n) {
                    uint32_t negmask = 0u;
                    float v[S];
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        v[s] = acc[s] + pr;                        // kernels.py:320
                        if (v[s] < 0.f) negmask |= 1u << s;        // kernels.py:349
                    }
                    if constexpr (S == 4) *reinterpret_cast<float4 *>(vals + j * 4) = make_float4(v[0], v[1], v[2], v[3]);
                    else if constexpr (S == 2) *reinterpret_cast<float2 *>(vals + j * 2) = make_float2(v[0], v[1]);
                    else vals[j] = v[0];
                    if (negmask) {                                 // hard decision 1: flip the parity of its checks
                        for (int c = 0; c * 4 < deg; ++c) {
                            const uint4 q4 = c == 0 ? e0 : (c == 1 ? e1 : ell[c * 32]);
                            const uint32_t ent[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t cidx = ent[i] >> 8;
                                if (c * 4 + i >= deg || cidx >= (uint32_t)g.m) continue;
#pragma unroll
                                for (int s = 0; s < S; ++s)
                                    if (negmask & (1u << s)) atomicXor(&par[s * g.mw + (cidx >> 5)], 1u << (cidx & 31));
                            }
                        }
                    }
                }
            }
            __syncthreads();

            // ---- convergence: H.hard == syndrome  (kernels.py:352-364) ---------------------------
            for (int i = tid; i < S * g.mw; i += blockDim.x) {
                const uint32_t d = par[i] ^ syn[i];
                if (d) { s_unsat[i / g.mw] = 1; atomicAdd(&s_wt[i / g.mw], __popc(d)); }    // weight of the residual syndrome
                par[i] = 0u;
            }
            __syncthreads();
            const bool last = (it == a.max_iter - 1);
            bool any_done = false;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                if (!s_active[s]) continue;
                const bool conv = (s_unsat[s] == 0);
                if (conv || last) {
                    any_done = true;
                    const size_t shot = shot0 + s;
                    for (int j = tid; j < g.n_pad; j += blockDim.x) {
                        const float v = vals[j * S + s];
                        const uint32_t word = __ballot_sync(0xFFFFFFFFu, j < g.n && v < 0.f);
                        if (lane == 0 && (j >> 5) < g.nw) a.hard_bits[shot * g.nw + (j >> 5)] = word;
                        if (a.post && j < g.n && !(a.post_failed_only && conv)) a.post[shot * g.n + j] = v;
                    }
                    if (tid == 0) {
                        a.converged[shot] = conv ? 1 : 0;
                        a.final_iter[shot] = it;
                        if (!conv && a.fail_count) {
                            const int slot = atomicAdd(a.fail_count, 1);
                            a.fail_idx[slot] = (int)shot;
                            if (a.fail_wt) a.fail_wt[slot] = s_wt[s];
                        }
                    }
                }
            }
            if (any_done || last) {                       // uniform across the CTA
                __syncthreads();
                if (tid == 0) {
                    int na = 0;
                    for (int s = 0; s < S; ++s) {
                        if (s_active[s] && (s_unsat[s] == 0 || last)) s_active[s] = 0;
                        na += s_active[s];
                    }
                    s_nactive = na;
                }
            }
            __syncthreads();
            if (tid < S) { s_unsat[tid] = 0; s_wt[tid] = 0; }
            if (s_nactive == 0) break;
            // a finished shot keeps being iterated (its outputs are already written and never
            // overwritten: s_active gates the output stage), which keeps the inner loops branch-free.
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// General kernel: per-edge messages in global memory, any degree, damping != 1, dense.py variant.
// One CTA per shot slot; ws holds per-slot Q[nnz], R[nnz], values[n] floats.
__global__ void __launch_bounds__(256)
minsum_general_kernel(GraphDev g, MinsumLaunch a, float *ws)
{
    extern __shared__ uint32_t sm_words[];
    uint32_t *syn = sm_words, *par = sm_words + g.mw;
    __shared__ int s_unsat;
    const int tid = threadIdx.x;
    float *Q = ws + (size_t)blockIdx.x * (2 * (size_t)g.nnz + g.n);
    float *R = Q + g.nnz;
    float *vals = R + g.nnz;
    const float clip = a.clip, damping = a.damping;

    for (int shot = blockIdx.x; shot < a.B; shot += gridDim.x) {
        for (int e = tid; e < g.nnz; e += blockDim.x) Q[e] = g.prior[g.indices[e]];
        for (int j = tid; j < g.n; j += blockDim.x) vals[j] = 0.f;
        for (int w = tid; w < g.mw; w += blockDim.x) { syn[w] = a.syn_bits[(size_t)shot * g.mw + w]; par[w] = 0u; }
        if (tid == 0) s_unsat = 0;
        __syncthreads();
        int fin = a.max_iter - 1;
        bool conv = false;
        for (int it = 0; it < a.max_iter; ++it) {
            const float alpha = a.alpha_d[it];
            for (int r = tid; r < g.m; r += blockDim.x) {
                const int rs = g.indptr[r], re = g.indptr[r + 1];
                if (rs == re) continue;
                uint32_t tot = (syn[r >> 5] >> (r & 31)) & 1u;
                float mn1 = INFINITY, mn2 = INFINITY;
                int mp = -1;
                for (int e = rs; e < re; ++e) {
                    const float q = Q[e];
                    tot ^= (q >= 0.f) ? 0u : 1u;
                    const float ab = fabsf(q);
                    if (ab < mn1) { mn2 = mn1; mn1 = ab; mp = e; } else if (ab < mn2) mn2 = ab;
                }
                const float a1 = alpha * mn1, a2 = alpha * mn2;
                for (int e = rs; e < re; ++e) {
                    const uint32_t sg = tot ^ ((Q[e] >= 0.f) ? 0u : 1u);
                    const float mag = (e == mp) ? a2 : a1;
                    R[e] = sg ? -mag : mag;
                }
            }
            __syncthreads();
            for (int j = tid; j < g.n; j += blockDim.x) {
                float acc = 0.f;
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) acc += R[g.csc_edge[p]];
                const float v = acc + g.prior[j];
                vals[j] = v;
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) {
                    const int e = g.csc_edge[p];
                    float q = v - R[e];
                    if (q != q) q = 0.f;
                    else if (a.dense_variant) { if (isinf(q)) q = q > 0.f ? clip : -clip; }
                    else q = fminf(fmaxf(q, -clip), clip);
                    float qd = damping * q + (1.0f - damping) * Q[e];
                    qd = fminf(fmaxf(qd, -clip), clip);
                    Q[e] = qd;
                    if (v < 0.f) { const int c = g.rowidx[p]; atomicXor(&par[c >> 5], 1u << (c & 31)); }
                }
            }
            __syncthreads();
            for (int w = tid; w < g.mw; w += blockDim.x) { if (par[w] != syn[w]) s_unsat = 1; par[w] = 0u; }
            __syncthreads();
            const bool ok = (s_unsat == 0);
            __syncthreads();
            if (tid == 0) s_unsat = 0;
            if (ok) { fin = it; conv = true; break; }
        }
        __syncthreads();
        for (int j0 = 0; j0 < g.n_pad; j0 += blockDim.x) {
            const int j = j0 + tid;
            const float v = j < g.n ? vals[j] : 0.f;
            const bool neg = (a.max_iter > 0) && j < g.n && v < 0.f;
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, neg);
            if ((tid & 31) == 0 && (j >> 5) < g.nw) a.hard_bits[(size_t)shot * g.nw + (j >> 5)] = word;
            if (a.post && j < g.n && !(a.post_failed_only && conv)) a.post[(size_t)shot * g.n + j] = v;
        }
        if (tid == 0) {
            a.converged[shot] = conv ? 1 : 0;
            a.final_iter[shot] = fin;
            if (!conv && a.fail_count) { const int slot = atomicAdd(a.fail_count, 1); a.fail_idx[slot] = shot; if (a.fail_wt) a.fail_wt[slot] = 0; }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// minsum_core_sparse (kernels.py:139-169), double precision (used by the alpha estimators).
__global__ void minsum_core_kernel(GraphDev g, const double *Q, const double *ssign, int B, double alpha,
                                   double *R, double *Rsum)
{
    const int shot = blockIdx.x;
    const double *q = Q + (size_t)shot * g.nnz;
    double *r_ = R + (size_t)shot * g.nnz;
    for (int r = threadIdx.x; r < g.m; r += blockDim.x) {
        const int rs = g.indptr[r], re = g.indptr[r + 1];
        if (rs == re) continue;
        double sp = ssign[(size_t)shot * g.m + r], mn1 = INFINITY, mn2 = INFINITY;
        int mp = -1;
        for (int e = rs; e < re; ++e) {
            const double v = q[e];
            sp *= (v >= 0) ? 1.0 : -1.0;
            const double ab = fabs(v);
            if (ab < mn1) { mn2 = mn1; mn1 = ab; mp = e; } else if (ab < mn2) mn2 = ab;
        }
        for (int e = rs; e < re; ++e) {
            const double sj = (q[e] >= 0) ? 1.0 : -1.0;
            r_[e] = alpha * (sp * sj) * ((e == mp) ? mn2 : mn1);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < g.n; j += blockDim.x) {
        double acc = 0.0;
        for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) acc += r_[g.csc_edge[p]];
        Rsum[(size_t)shot * g.n + j] = acc;
    }
}

// Messages for the Alvarado alpha estimators (src/decoding/alpha.py:122-137 and :206-253): advance the
// decoder n_prev iterations with the given alphas (no convergence stop, damping / clip like the decoder,
// alpha.py:217-243), then one unscaled check pass (alpha = 1) whose messages R are returned.
// Double precision like the reference's pre-pass.  ws per slot: Q[nnz], R[nnz] doubles.
__global__ void __launch_bounds__(256)
alpha_messages_kernel(GraphDev g, const int8_t *syndrome, int B, int n_prev, const double *alpha_prev,
                      double damping, double clip, const double *prior64, double *R_out, double *ws)
{
    const int tid = threadIdx.x;
    double *Q = ws + (size_t)blockIdx.x * 2 * (size_t)g.nnz;
    double *R = Q + g.nnz;
    for (int shot = blockIdx.x; shot < B; shot += gridDim.x) {
        const int8_t *syn = syndrome + (size_t)shot * g.m;
        for (int e = tid; e < g.nnz; e += blockDim.x) Q[e] = prior64[g.indices[e]];
        __syncthreads();
        for (int it = 0; it <= n_prev; ++it) {
            const bool final_pass = (it == n_prev);
            const double alpha = final_pass ? 1.0 : alpha_prev[it];
            double *Rdst = final_pass ? R_out + (size_t)shot * g.nnz : R;
            for (int r = tid; r < g.m; r += blockDim.x) {
                const int rs = g.indptr[r], re = g.indptr[r + 1];
                if (rs == re) continue;
                double sp = 1.0 - 2.0 * (double)syn[r], mn1 = INFINITY, mn2 = INFINITY;
                int mp = -1;
                for (int e = rs; e < re; ++e) {
                    const double v = Q[e];
                    sp *= (v >= 0) ? 1.0 : -1.0;
                    const double ab = fabs(v);
                    if (ab < mn1) { mn2 = mn1; mn1 = ab; mp = e; } else if (ab < mn2) mn2 = ab;
                }
                for (int e = rs; e < re; ++e) {
                    const double sj = (Q[e] >= 0) ? 1.0 : -1.0;
                    Rdst[e] = alpha * (sp * sj) * ((e == mp) ? mn2 : mn1);
                }
            }
            __syncthreads();
            if (final_pass) break;
            for (int j = tid; j < g.n; j += blockDim.x) {
                double acc = 0.0;
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) acc += R[g.csc_edge[p]];
                const double v = acc + prior64[j];
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) {
                    const int e = g.csc_edge[p];
                    double q = v - R[e];
                    if (q != q) q = 0.0; else if (q > clip) q = clip; else if (q < -clip) q = -clip;
                    double qd = damping * q + (1.0 - damping) * Q[e];
                    if (qd > clip) qd = clip; else if (qd < -clip) qd = -clip;
                    Q[e] = qd;
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// tanh/atanh BP (dense.py:75-96, kernels.py:172-193), double precision; ws per slot: Q[nnz], R[nnz], vals[n]
__global__ void __launch_bounds__(256)
bp_kernel(GraphDev g, const uint32_t *syn_bits, int B, int max_iter, uint32_t *hard_bits, uint8_t *converged,
          int32_t *final_iter, double *post, double *ws)
{
    extern __shared__ uint32_t sm_words[];
    uint32_t *syn = sm_words, *par = sm_words + g.mw;
    __shared__ int s_unsat;
    const int tid = threadIdx.x;
    double *Q = ws + (size_t)blockIdx.x * (2 * (size_t)g.nnz + g.n);
    double *R = Q + g.nnz;
    double *vals = R + g.nnz;
    const double CLIP = 0.9999999;
    for (int shot = blockIdx.x; shot < B; shot += gridDim.x) {
        for (int e = tid; e < g.nnz; e += blockDim.x) Q[e] = (double)g.prior[g.indices[e]];
        for (int j = tid; j < g.n; j += blockDim.x) vals[j] = 0.0;
        for (int w = tid; w < g.mw; w += blockDim.x) { syn[w] = syn_bits[(size_t)shot * g.mw + w]; par[w] = 0u; }
        if (tid == 0) s_unsat = 0;
        __syncthreads();
        int fin = max_iter - 1;
        bool conv = false;
        for (int it = 0; it < max_iter; ++it) {
            for (int r = tid; r < g.m; r += blockDim.x) {
                const double ss = ((syn[r >> 5] >> (r & 31)) & 1u) ? -1.0 : 1.0;
                double prod = 1.0;
                for (int e = g.indptr[r]; e < g.indptr[r + 1]; ++e) {
                    double t = tanh(Q[e] * 0.5);
                    if (fabs(t) < 1e-15) t = (t >= 0) ? 1e-15 : -1e-15;
                    prod *= t;
                }
                for (int e = g.indptr[r]; e < g.indptr[r + 1]; ++e) {
                    double t = tanh(Q[e] * 0.5);
                    if (fabs(t) < 1e-15) t = (t >= 0) ? 1e-15 : -1e-15;
                    double po = prod / t * ss;
                    po = fmin(fmax(po, -CLIP), CLIP);
                    R[e] = 2.0 * atanh(po);
                }
            }
            __syncthreads();
            for (int j = tid; j < g.n; j += blockDim.x) {
                double acc = 0.0;
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) acc += R[g.csc_edge[p]];
                const double v = acc + (double)g.prior[j];
                vals[j] = v;
                for (int p = g.colptr[j]; p < g.colptr[j + 1]; ++p) {
                    const int e = g.csc_edge[p];
                    Q[e] = v - R[e];
                    if (v < 0) { const int c = g.rowidx[p]; atomicXor(&par[c >> 5], 1u << (c & 31)); }
                }
            }
            __syncthreads();
            for (int w = tid; w < g.mw; w += blockDim.x) { if (par[w] != syn[w]) s_unsat = 1; par[w] = 0u; }
            __syncthreads();
            const bool ok = (s_unsat == 0);
            __syncthreads();
            if (tid == 0) s_unsat = 0;
            if (ok) { fin = it; conv = true; break; }
        }
        __syncthreads();
        for (int j0 = 0; j0 < g.n_pad; j0 += blockDim.x) {
            const int j = j0 + tid;
            const double v = j < g.n ? vals[j] : 0.0;
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, max_iter > 0 && j < g.n && v < 0);
            if ((tid & 31) == 0 && (j >> 5) < g.nw) hard_bits[(size_t)shot * g.nw + (j >> 5)] = word;
            if (post && j < g.n) post[(size_t)shot * g.n + j] = v;
        }
        if (tid == 0) { converged[shot] = conv ? 1 : 0; final_iter[shot] = fin; }
        __syncthreads();
    }
}

// H.candidate mod 2 (syndrome_check, kernels.py:223-231): one warp per (shot, 32 rows)
__global__ void syndrome_check_kernel(GraphDev g, const uint32_t *cand_bits, int B, uint32_t *syn_bits)
{
    const int shot = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = 0;
    if (r < g.m)
        for (int e = g.indptr[r]; e < g.indptr[r + 1]; ++e) {
            const int j = g.indices[e];
            s ^= (cand_bits[(size_t)shot * g.nw + (j >> 5)] >> (j & 31)) & 1u;
        }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, s != 0);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < g.mw) syn_bits[(size_t)shot * g.mw + (r >> 5)] = word;
}

// ------------------------------------------------------------------------------------------------
static size_t fast_smem_bytes(const GraphDev &g, int S)
{
    return (size_t)vals_rows(g) * S * 4 + (size_t)S * chk_rows(g) * 16 + (size_t)2 * S * g.mw * 4 + (size_t)(2 * g.n_rslices + 2 * g.n_cslices) * 4;
}

int fast_shots_per_cta(const qb_decoder *dec)
{
    if (!dec->fast_ok) return 0;
    const int cand[3] = {4, 2, 1};
    if (const char *e = getenv("QLDPC_B200_MS_S")) {      // tuning override
        const int S = atoi(e);
        if ((S == 1 || S == 2 || S == 4) && fast_smem_bytes(dec->g, S) + 1024 <= (size_t)dec->max_smem_optin) return S;
    }
    for (int S : cand)
        if (fast_smem_bytes(dec->g, S) + 1024 <= (size_t)dec->max_smem_optin) return S;
    return 0;
}

template <int S>
static int launch_fast(qb_decoder *dec, const MinsumLaunch &a, cudaStream_t st)
{
    const size_t smem = fast_smem_bytes(dec->g, S);
    QB_CUDA(cudaFuncSetAttribute(minsum_fast_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (a.B + S - 1) / S;
    int ctas_per_sm = std::max(1, std::min(8, (int)(((size_t)dec->max_smem_optin + 1024) / (smem + 1024))));
    // thread count: 512 while at most two CTAs fit an SM, fewer for small codes so that several CTAs share it
    int threads = MS_THREADS;
    if (ctas_per_sm > 2) threads = 256;
    if (S == 1 && ctas_per_sm == 1) threads = 1024;       // one large shot per SM (e.g. [[288,12,18]]): all 32 warps on it
    if (const char *e = getenv("QLDPC_B200_MS_THREADS")) { const int t = atoi(e); if (t >= 64 && t <= (S == 1 ? 1024 : MS_THREADS) && t % 32 == 0) threads = t; }
    ctas_per_sm = std::min(ctas_per_sm, 2048 / threads);
    const int grid = std::max(1, std::min(tiles, dec->sm_count * ctas_per_sm));
    minsum_fast_kernel<S><<<grid, threads, smem, st>>>(dec->g, a);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_minsum(qb_decoder *dec, const MinsumLaunch &a, cudaStream_t st)
{
    if (a.B <= 0) return QB_OK;
    if (a.precision == QB_PRECISION_HALF2) {
        // opt-in packed mode: only on the per-edge plan (damping 1, uniform priors per slice); no silent downgrade
        if (!(a.damping == 1.0f && dec->edge && a.max_iter > 0)) { set_error("packed half2 min-sum needs the per-edge plan (damping = 1, max_iter > 0, graph fits one SM)"); return QB_ERR_UNSUPPORTED; }
        if (!dec->edge_h2) if (int rc = edge_plan_h2_create(dec->edge, dec->g.nw, dec->h_prior.data(), &dec->edge_h2)) return rc;
        if (!edge_h2_fits(dec, dec->edge, dec->edge_h2)) { set_error("packed half2 min-sum: plan does not fit (shared memory / per-lane priors)"); return QB_ERR_UNSUPPORTED; }
        return launch_minsum_edge_h2(dec, dec->edge, dec->edge_h2, a, st);
    }
    if (a.damping == 1.0f && dec->edge && a.max_iter > 0) return launch_minsum_edge(dec, dec->edge, a, st);
    if (a.damping == 1.0f && dec->cluster && a.max_iter > 0) return launch_minsum_cluster(dec, dec->cluster, a, st);
    const int S = (a.damping == 1.0f) ? fast_shots_per_cta(dec) : 0;
    if (S == 4) return launch_fast<4>(dec, a, st);
    if (S == 2) return launch_fast<2>(dec, a, st);
    if (S == 1) return launch_fast<1>(dec, a, st);
    // general path
    const GraphDev &g = dec->g;
    const int slots = std::max(1, std::min(a.B, dec->sm_count * 4));
    const size_t per = (2 * (size_t)g.nnz + g.n) * sizeof(float);
    if (int rc = dec->work.ensure(per * slots)) return rc;
    minsum_general_kernel<<<slots, 256, 2 * g.mw * sizeof(uint32_t), st>>>(g, a, dec->work.as<float>());
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_minsum_core(qb_decoder *dec, const double *Q, const double *ssign, int B, double alpha, double *R,
                       double *Rsum, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    minsum_core_kernel<<<B, 256, 0, st>>>(dec->g, Q, ssign, B, alpha, R, Rsum);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_alpha_messages(qb_decoder *dec, const int8_t *syn, int B, int n_prev, const double *alpha_prev_d, double damping,
                          double clip, const double *prior64_d, double *R_out, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    const GraphDev &g = dec->g;
    const int slots = std::max(1, std::min(B, dec->sm_count * 8));
    if (int rc = dec->work.ensure((size_t)slots * 2 * (size_t)std::max(1, g.nnz) * sizeof(double))) return rc;
    alpha_messages_kernel<<<slots, 256, 0, st>>>(g, syn, B, n_prev, alpha_prev_d, damping, clip, prior64_d, R_out, dec->work.as<double>());
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_bp(qb_decoder *dec, const uint32_t *syn_bits, int B, int max_iter, uint32_t *hard_bits,
              uint8_t *converged, int32_t *final_iter, double *post, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    const GraphDev &g = dec->g;
    const int slots = std::max(1, std::min(B, dec->sm_count * 4));
    const size_t per = (2 * (size_t)g.nnz + g.n) * sizeof(double);
    if (int rc = dec->work.ensure(per * slots)) return rc;
    bp_kernel<<<slots, 256, 2 * g.mw * sizeof(uint32_t), st>>>(g, syn_bits, B, max_iter, hard_bits, converged,
                                                               final_iter, post, dec->work.as<double>());
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_syndrome_check(qb_decoder *dec, const uint32_t *cand_bits, int B, uint32_t *syn_bits, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    dim3 grid(ceil_div(dec->g.m_pad, 128), B);
    syndrome_check_kernel<<<grid, 128, 0, st>>>(dec->g, cand_bits, B, syn_bits);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int upload_alpha(qb_decoder *dec, int max_iter, int alpha_mode, double alpha, const double *seq, int len,
                 cudaStream_t st)
{
    const int n = std::max(1, max_iter);
    if (n > dec->alpha_cap) {
        if (dec->d_alpha) cudaFree(dec->d_alpha);
        dec->d_alpha = nullptr;
        QB_CUDA(cudaMalloc(&dec->d_alpha, sizeof(float) * n));
        dec->alpha_cap = n;
    }
    std::vector<float> h(n);
    for (int it = 0; it < n; ++it) {
        double v;
        if (alpha_mode == QB_ALPHA_DYNAMIC) v = 1.0 - pow(2.0, -(double)(it + 1));       // kernels.py:273
        else if (alpha_mode == QB_ALPHA_SEQUENCE) v = seq[it < len ? it : len - 1];        // kernels.py:402-405
        else v = alpha;
        h[it] = (float)v;
    }
    QB_CUDA(cudaMemcpyAsync(dec->d_alpha, h.data(), sizeof(float) * n, cudaMemcpyHostToDevice, st));
    QB_CUDA(cudaStreamSynchronize(st));   // h goes out of scope
    return QB_OK;
}

}  // namespace qb
