// K3 for graphs whose edge array does not fit one SM (the [[288,12,18]] code: 376 KB of messages + 183 KB of slot
// indices per side): the per-edge min-sum of minsum_edge.cu on a THREAD-BLOCK CLUSTER, one shot per cluster.
//
// Reference semantics: src/decoding/kernels.py:235-366 with damping == 1, exactly as minsum_edge.cu (the two phases are
// the same device functions, edge_phases.cuh).  The check rows are cut into NC contiguous slabs (syndrome bit index =
// round * n2 + check, so a slab is a block of rounds); CTA c of the cluster keeps the messages of its slab's rows in its
// shared memory and owns every variable whose first row lies in the slab.  A fault touches at most two consecutive rounds,
// so a variable either lives entirely in its owner's slab (~85 % for four slabs of five rounds: these go through the
// unchanged conflict-free column path) or straddles into the next slab: the owner then gathers / scatters the far
// messages through distributed shared memory (cluster.map_shared_rank), from a plain per-variable list.  Rows of the next
// slab reserve their slots for such edges as "phantom" columns in the layout builder.
//   per iteration: check rows (local) | cluster barrier | variables (local + DSMEM for straddlers), fingerprint into
//   CTA 0 | cluster barrier.  Convergence: the 8-bit fingerprint of H.hard is accumulated in CTA 0's shared memory by all
//   CTAs; on a match the exact parity is formed slab by slab (straddlers XOR into the next CTA's parity words) and the
//   residual weights are added up in CTA 0.
// Outputs: hard decisions with atomicOr into the (pre-cleared) global words, posteriors per owner, flags by CTA 0.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "common.cuh"
#include "edge_dev.cuh"
#include "edge_layout.h"
#include "edge_phases.cuh"

namespace qb {

constexpr int CL_MAXC = 8;              // largest cluster tried (portable limit)
constexpr int CL_DMAX = 8;              // straddling variables: at most 8 edges
constexpr uint32_t CL_REMOTE = 0x80000000u;

// tables of one CTA rank
struct ClusterRankDev {
    EdgeDev eg;                  // its slab as a self-contained per-edge plan (row_id / var_id hold GLOBAL ids; n, nw, mw global)
    const EdgePriors *pri;       // per-slice priors (global memory; the single-CTA kernel keeps them in the constant bank)
    int n_str, n_strp;           // straddling variables owned by this rank; rounded up to a multiple of 32
    // [6][n_strp], copied to shared memory: rows 0-3 = word of E of edges (2q, 2q+1) as a uint16 pair, edges in row order (the
    // edges in the lower slab first); row 4 = global variable id | degree << 16 | edges in the lower slab << 20 |
    // (owner is the upper slab) << 23 | fingerprint << 24; row 5 = prior bits.  An edge is in the OTHER rank's E (the
    // next rank when the owner is the lower slab, the previous one otherwise) when it is not in the owner's slab.
    const uint32_t *str_tab;
    const uint32_t *str_rowpos;  // [CL_DMAX][n_str] permuted row position of each edge's row (| CL_REMOTE: in the other rank)
};

// barrier.cluster is split: arrive (release: this CTA's shared-memory writes become visible), later wait (acquire).  Every
// arrive is matched by one wait before the next arrive.  Both are CTA-wide as well (all threads of all CTAs take part).
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() { cl_arrive(); cl_wait(); }
// shared::cluster address of the same variable in CTA `rank`, and accesses through it
__device__ __forceinline__ uint32_t cl_map(const void *p, unsigned rank)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cl_ld_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float cl_ld_f32(uint32_t a) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void cl_st_f32(uint32_t a, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void cl_st_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_xor_u32(uint32_t a, uint32_t v) { asm volatile("red.relaxed.cluster.shared::cluster.xor.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_add_u32(uint32_t a, uint32_t v) { asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

#ifdef QB_CLUSTER_PROFILE
__device__ unsigned long long g_cluster_prof[256 * 32 * 8];   // [cta][warp]{rows, local sync, columns, wait A, straddlers, sync B, check, iterations}
#define CPROF_T(x) long long x; asm volatile("mov.u64 %0, %%clock64;" : "=l"(x) :: "memory")
#define CPROF_ADD(i, d) prof[i] += (unsigned long long)(d)
#else
#define CPROF_T(x)
#define CPROF_ADD(i, d)
#endif

struct ClusterWarpTask { uint4 cls; int r0, r1, c0, c1; uint32_t ix0, pad[3]; };

struct ClusterShared {                   // same static layout in every CTA of the cluster
    ClusterWarpTask task[32];            // per warp: its row slices, column slices and their classes (re-read every phase)
    ClusterRankDev R;
    EdgePriors pri;
    float alpha[128];
    uint16_t plist[PAR_LIST_CAP];
    uint32_t fp[2];                      // fingerprint of H.hard (pushed by the warps of all CTAs), per iteration parity
    uint32_t target;                     // fingerprint of the syndrome (same)
    uint32_t par_off;                    // byte offset of the parity words in the dynamic part
    uint32_t e_own, e_next, e_prev;      // shared::cluster addresses of the messages of this / the next / the previous rank
    uint32_t par_next, par_prev;         // ... and of their parity words
    int wt, next, pcount;
};

// exact weight of H.hard ^ syndrome over the whole cluster (cluster-uniform call; par must be 0 on entry, is 0 on exit)
template <int THREADS>
__device__ __noinline__ int cluster_residual_weight(ClusterShared *S, const uint32_t *hperm, const uint32_t *hstr, const uint32_t *cmeta,
                                                    uint32_t *par, const uint32_t *syn, const uint32_t *stab, unsigned NC)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const ClusterRankDev &R = S->R;
    parity_of_hard(R.eg, hperm, cmeta, par, S->plist, &S->pcount, tid, THREADS);
    for (int i = tid; i < R.n_str; i += THREADS) {
        if (!((hstr[i >> 5] >> (i & 31)) & 1u)) continue;
        const uint32_t meta = stab[4 * R.n_strp + i];
        const int D = (meta >> 16) & 15;
        const uint32_t par_rem = ((meta >> 23) & 1u) ? S->par_prev : S->par_next;
        for (int k = 0; k < D; ++k) {
            const uint32_t rp = R.str_rowpos[(size_t)k * R.n_str + i];
            const uint32_t pos = rp & ~CL_REMOTE;
            if (rp & CL_REMOTE) cl_xor_u32(par_rem + (pos >> 5) * 4, 1u << (pos & 31));
            else atomicXor(&par[pos >> 5], 1u << (pos & 31));
        }
    }
    cl_sync();
    if (tid < 32) { const int w = residual_weight(par, syn, R.eg.n_rsl, lane); if (lane == 0) S->wt = w; }
    cl_sync();
    int wt = 0;
    for (unsigned r = 0; r < NC; ++r) wt += (int)cl_ld_u32(cl_map(&S->wt, r));
    return wt;              // (S->wt is next written after at least one more cluster barrier)
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
minsum_cluster_kernel(const ClusterRankDev *ranks, const __grid_constant__ MinsumLaunch a, int *shot_counter)
{
    unsigned rank, NC;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(NC));
    __shared__ ClusterShared S;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_min_sync(0xFFFFFFFFu, tid >> 5);
    {   // rank tables: one coalesced copy
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ranks + rank);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&S.R);
        for (int i = tid; i < (int)(sizeof(ClusterRankDev) / 4); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const ClusterRankDev &R = S.R;
    const EdgeDev &eg = R.eg;
    for (int i = tid; i < EDGE_MAX_CSL; i += THREADS) S.pri.bits[i] = R.pri->bits[i];

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *E = reinterpret_cast<float *>(smem_raw);                                   // [e_words]   (same offset in every rank)
    uint32_t *idx = reinterpret_cast<uint32_t *>(E + eg.e_words);
    uint2 *rtask = reinterpret_cast<uint2 *>(idx + eg.idx_words);
    uint32_t *syn = reinterpret_cast<uint32_t *>(rtask + eg.n_rsl);
    uint32_t *par = syn + eg.n_rsl;
    uint32_t *hperm = par + eg.n_rsl;
    uint32_t *hstr = hperm + eg.n_csl;                                                // [ceil(n_str / 32) + 1] hard decisions of the straddlers
    uint32_t *cmeta = hstr + ((R.n_str + 31) >> 5) + 1;
    uint8_t *csig = reinterpret_cast<uint8_t *>(cmeta + eg.n_csl);
    uint32_t *stab = reinterpret_cast<uint32_t *>(csig + eg.n_csl * 32);            // [6][n_strp] straddler tables

    const uint32_t idx_addr = (uint32_t)__cvta_generic_to_shared(idx);
    // a CTA's shared window starts at (rank in the cluster) << 24: slot indices stay 16-bit word offsets inside it
    const uint32_t e_addr = (uint32_t)__cvta_generic_to_shared(E);
    const uint32_t e_word = (e_addr & 0x00FFFFFFu) >> 2;
    const uint32_t win = e_addr & 0xFF000000u;
    for (int i = tid; i < eg.idx_words; i += THREADS) idx[i] = eg.col_idx[i] + (e_word | (e_word << 16));
    for (int i = tid; i < eg.n_csl; i += THREADS) { cmeta[i] = eg.ctask[i].x; hperm[i] = 0u; }
    for (int i = tid; i < eg.n_rsl; i += THREADS) { rtask[i] = eg.rtask[i]; par[i] = 0u; }
    for (int i = tid; i < 128 && i < a.max_iter; i += THREADS) S.alpha[i] = a.alpha_d[i];
    for (int i = tid; i < eg.n_csl * 32; i += THREADS) csig[i] = (uint8_t)eg.col_sig[i];
    for (int i = tid; i < ((R.n_str + 31) >> 5) + 1; i += THREADS) hstr[i] = 0u;
    for (int i = tid; i < 6 * R.n_strp; i += THREADS) stab[i] = R.str_tab[i];
    if (tid < 32) E[eg.e_dummy + tid] = 0.f;
    if (tid < 32 && tid < THREADS / 32) {
        ClusterWarpTask &T = S.task[tid];
        T.r0 = eg.wr_ptr[tid]; T.r1 = eg.wr_ptr[tid + 1];
        T.c0 = eg.wc_ptr[tid]; T.c1 = eg.wc_ptr[tid + 1];
        T.cls = eg.wc_cls[tid];
        T.ix0 = idx_addr + (T.c0 < eg.n_csl ? (eg.ctask[T.c0].x & 0xFFFFu) * 128u : 0u);
    }
    const bool api = !a.post_failed_only;
    if (tid == 0) { S.par_off = (uint32_t)(reinterpret_cast<unsigned char *>(par) - smem_raw); S.fp[0] = 0u; S.fp[1] = 0u; S.target = 0u; S.wt = 0; }
    cl_sync();
    // the next rank's messages and parity words (its carve-up depends on its own table sizes)
    const unsigned nxt = rank + 1 < NC ? rank + 1 : rank;
    // (values used outside the column phase are kept in shared memory and re-read where needed: the column phase has
    //  no register to spare, and a spill would be re-read from L2 after every cluster barrier, which invalidates L1)
    const unsigned prv = rank > 0 ? rank - 1 : rank;
    if (tid == 0) {
        S.e_own = cl_map(E, rank); S.e_next = cl_map(E, nxt); S.e_prev = cl_map(E, prv);
        S.par_next = cl_map(smem_raw, nxt) + cl_ld_u32(cl_map(&S.par_off, nxt));
        S.par_prev = cl_map(smem_raw, prv) + cl_ld_u32(cl_map(&S.par_off, prv));
    }
    __syncthreads();
#ifdef QB_CLUSTER_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif

    while (true) {
        if (rank == 0 && tid == 0) {
            const int sh = atomicAdd(shot_counter, 1);
            for (unsigned r = 0; r < NC; ++r) cl_st_u32(cl_map(&S.next, r), (uint32_t)sh);
        }
        if (tid == 0) { S.target = 0u; S.fp[0] = 0u; S.fp[1] = 0u; }
        cl_sync();
        const int shot = S.next;
        if (shot >= a.B) break;
        // ---- syndrome bits of the slab's rows (row_id holds global rows) and their fingerprint ----
        {
            uint32_t tg = 0u;
            for (int t = warp; t < eg.n_rsl; t += THREADS / 32) {
                const uint32_t rid = eg.row_id[t * 32 + lane];
                const bool bit = rid != 0xFFFFu && ((a.syn_bits[(size_t)shot * eg.mw + (rid >> 5)] >> (rid & 31)) & 1u);
                if (bit) tg ^= eg.row_mask[t * 32 + lane] & 0xFFu;
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
                if (lane == 0) syn[t] = word;
            }
            tg = __reduce_xor_sync(0xFFFFFFFFu, tg);
            if (lane < NC && tg) cl_xor_u32(cl_map(&S.target, lane), tg);          // every CTA accumulates the whole fingerprint
        }
        cl_sync();
        bool conv = false;
        int fin = a.max_iter - 1;
        for (int it = 0; it < a.max_iter; ++it) {
            CPROF_T(t0);
            // ---- check rows of the slab ----
            const float alpha = it < 128 ? S.alpha[it] : a.alpha_d[it];
            const int r0 = S.task[warp].r0, r1 = S.task[warp].r1;
            for (int t = r0; t < r1; ++t) {
                const uint2 d = rtask[t];
                const int K = d.y & 255, nl = (d.y >> 8) & 255, stride = d.y >> 16;
                if (K == 0 || lane >= nl) continue;
                const uint32_t synsign = ((syn[t] >> lane) & 1u) << 31;
                const uint2 pads = __ldg(&eg.row_pads[t * 32 + lane]);
                if (it == 0) row_dispatch<true>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, INFINITY, pads);
                else row_dispatch<false>(E, eg.E0, (int)(d.x >> 2), stride, lane, K, synsign, alpha, a.clip, pads);
            }
            if (tid == 0) S.fp[(it + 1) & 1] = 0u;      // (its last readers passed barrier B of the previous iteration)
            CPROF_T(t1);
            cl_arrive();                     // (A) this slab's rows are done ...
            __syncthreads();
            CPROF_T(t2);
            // ---- variables of this rank that live inside the slab: need the local rows only ----
            const bool write_v = a.post && (api || it == a.max_iter - 1);
            const uint4 cls = S.task[warp].cls;
            const int c0 = S.task[warp].c0, c1 = S.task[warp].c1;
            ColCtx c;
            c.ix = S.task[warp].ix0;
            c.lane4 = lane * 4; c.lane8 = lane * 8;
            c.sg = (uint32_t)__cvta_generic_to_shared(csig + lane);
            c.fp = 0u; c.fpw = 0u; c.myhw = 0u;
            c.t4 = 4u * (uint32_t)c0; c.lane_t4 = 4u * (uint32_t)(c0 + lane); c.lane = lane;
            c.vid = eg.var_id + c0 * 32 + lane;
            c.vid_next = write_v ? __ldg(c.vid) : 0u;
            c.post = a.post ? a.post + (size_t)shot * eg.n : nullptr;
            c.win = win;
            if (write_v) phase_b<true>(c, cls, c1, cmeta, eg.lane_prior, S.pri);
            else phase_b<false>(c, cls, c1, cmeta, eg.lane_prior, S.pri);
            if (lane < c1 - c0) hperm[c0 + lane] = c.myhw;
            uint32_t fp = c.fp;
            CPROF_T(t3);
            cl_wait();                       // ... (A) and so are the next slab's
            CPROF_T(t4);
            // ---- its straddling variables: the messages of the other slab through distributed shared memory ----
            const int n_str = R.n_str, n_strp = R.n_strp;
            const uint32_t e_next = S.e_next, e_prev = S.e_prev;
            float *post_row = c.post;
            for (int i0 = warp * 32; i0 < n_str; i0 += THREADS) {
                const int i = i0 + lane;
                bool neg = false;
                if (i < n_str) {
                    const uint32_t meta = stab[4 * n_strp + i];
                    const int D = (meta >> 16) & 15, nf = (meta >> 20) & 7;
                    const bool rf = (meta >> 23) & 1u;
                    const uint32_t e_rem = rf ? e_prev : e_next;                       // E of the other slab
                    uint32_t w[CL_DMAX / 2];
                    float r[CL_DMAX];
#pragma unroll
                    for (int q = 0; q < CL_DMAX / 2; ++q) w[q] = 2 * q < D ? stab[q * n_strp + i] : 0u;
#pragma unroll
                    for (int k = 0; k < CL_DMAX; ++k) {
                        const uint32_t aw = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                        r[k] = 0.f;
                        if (k < D) r[k] = ((k < nf) != rf) ? E[aw] : cl_ld_f32(e_rem + aw * 4u);   // (scattered DSMEM accesses are slow: only where needed)
                    }
                    float acc = r[0];                                                 // kernels.py:316 (row order)
#pragma unroll
                    for (int k = 1; k < CL_DMAX; ++k) if (k < D) acc += r[k];
                    const float v = acc + __uint_as_float(stab[5 * n_strp + i]);      // kernels.py:320
#pragma unroll
                    for (int k = 0; k < CL_DMAX; ++k) {
                        if (k < D) {
                            const uint32_t aw = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                            float q = v - r[k];
                            q = (q != q) ? 0.f : q;                                   // kernels.py:328-329
                            if ((k < nf) != rf) E[aw] = q; else cl_st_f32(e_rem + aw * 4u, q);
                        }
                    }
                    neg = v < 0.f;
                    if (neg) fp ^= meta >> 24;
                    if (write_v) post_row[meta & 0xFFFFu] = v;
                }
                const uint32_t hw = __ballot_sync(0xFFFFFFFFu, neg);
                if (lane == 0) hstr[i0 >> 5] = hw;
            }
            fp = __reduce_xor_sync(0xFFFFFFFFu, fp);
            if (lane < NC && fp) cl_xor_u32(cl_map(&S.fp[it & 1], lane), fp);        // every CTA accumulates the whole fingerprint
            CPROF_T(t5);
            cl_sync();                       // (B) all variables done, fingerprints complete
            CPROF_T(t6);
            bool done = false;
            if (S.fp[it & 1] == S.target)                                             // same values in every CTA: uniform over the cluster
                done = cluster_residual_weight<THREADS>(&S, hperm, hstr, cmeta, par, syn, stab, NC) == 0;   // kernels.py:352-364
            CPROF_T(t7);
            CPROF_ADD(0, t1 - t0); CPROF_ADD(1, t2 - t1); CPROF_ADD(2, t3 - t2); CPROF_ADD(3, t4 - t3); CPROF_ADD(4, t5 - t4);
            CPROF_ADD(5, t6 - t5); CPROF_ADD(6, t7 - t6); CPROF_ADD(7, 1);
            if (done) { conv = true; fin = it; break; }
        }
        // ---- end of the shot: hard decision into the (pre-cleared) global words, residual weight for the OSD queue ----
        uint32_t *hard_out = a.hard_bits + (size_t)shot * eg.nw;
        for (int t = tid; t < eg.n_csl; t += THREADS) {
            uint32_t bits = hperm[t];
            while (bits) {
                const int b = __ffs(bits) - 1; bits &= bits - 1;
                const uint32_t vid = eg.var_id[t * 32 + b];
                atomicOr(&hard_out[vid >> 5], 1u << (vid & 31));
            }
        }
        for (int i = tid; i < R.n_str; i += THREADS)
            if ((hstr[i >> 5] >> (i & 31)) & 1u) { const uint32_t vid = stab[4 * R.n_strp + i] & 0xFFFFu; atomicOr(&hard_out[vid >> 5], 1u << (vid & 31)); }
        int wt = 0;
        if (!conv && a.max_iter > 0 && a.fail_wt != nullptr) wt = cluster_residual_weight<THREADS>(&S, hperm, hstr, cmeta, par, syn, stab, NC);
        if (rank == 0 && tid == 0) {
            a.converged[shot] = conv ? 1 : 0;
            a.final_iter[shot] = fin;
            if (!conv && a.fail_count) {
                const int slot = atomicAdd(a.fail_count, 1);
                a.fail_idx[slot] = shot;
                if (a.fail_wt) a.fail_wt[slot] = wt;
            }
        }
    }
    cl_sync();          // no CTA leaves while a neighbour may still address its shared memory
#ifdef QB_CLUSTER_PROFILE
    if (lane == 0 && blockIdx.x < 256)
        for (int i = 0; i < 8; ++i) g_cluster_prof[(blockIdx.x * 32 + warp) * 8 + i] = prof[i];
#endif
}

#ifdef QB_CLUSTER_PROFILE
extern "C" int qb_debug_cluster_profile(unsigned long long *out_h)
{
    return cudaMemcpyFromSymbol(out_h, g_cluster_prof, sizeof(g_cluster_prof)) == cudaSuccess ? 0 : -2;
}
#endif

// ---- host side -------------------------------------------------------------------------------------------------------
struct ClusterPlan {
    int nc = 0;
    size_t smem = 0;
    std::vector<void *> owned;
    ClusterRankDev *d_ranks = nullptr;
    int *d_counter = nullptr;
    int n = 0, nw = 0;
    int max_clusters = 0;
};

void cluster_plan_destroy(ClusterPlan *p)
{
    if (!p) return;
    for (void *q : p->owned) cudaFree(q);
    delete p;
}

template <class T>
static int cup(ClusterPlan *p, const std::vector<T> &h, const T **out)
{
    T *d = nullptr;
    QB_CUDA(cudaMalloc(reinterpret_cast<void **>(&d), sizeof(T) * std::max<size_t>(1, h.size())));
    p->owned.push_back(d);
    if (!h.empty()) QB_CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    *out = d;
    return QB_OK;
}

static size_t cluster_smem_bytes(const EdgeLayout &L, int n_str)
{
    return (size_t)L.e_words * 4 + (size_t)L.idx_words * 4 + (size_t)L.n_rsl * 8 + (size_t)L.n_rsl * 8 + (size_t)L.n_csl * 4 +
           ((size_t)(n_str + 31) / 32 + 1) * 4 + (size_t)L.n_csl * 4 + (size_t)L.n_csl * 32 + (size_t)(n_str + 31) / 32 * 32 * 24 + 64;
}

// Build the cluster plan for a graph that does not fit one SM.  *out = nullptr when it cannot be built (a variable spans
// more than two slabs, a slab still does not fit, ...): the caller keeps the compressed-state kernel.
int cluster_plan_create(const qb_decoder *dec, const float *prior, ClusterPlan **out)
{
    *out = nullptr;
    // Opt-in (QLDPC_B200_CLUSTER=1): measured on B200 for the [[288,12,18]] graphs (profiles/r2_cluster_288.txt) the
    // cluster kernel is bit-identical to the compressed-state kernel but 10 % slower (455 against 414 ms per 8192 shots at
    // maxIter 100), so the latter stays the default.
    const char *on = getenv("QLDPC_B200_CLUSTER");
    if (!on || !on[0] || on[0] == '0') return QB_OK;
    const GraphDev &g = dec->g;
    const int m = g.m, n = g.n;
    if (m <= 0 || n <= 0 || n >= 65535 || m >= 65535) return QB_OK;
    const std::vector<int32_t> &indptr = dec->h_indptr, &indices = dec->h_indices, &colptr = dec->h_colptr, &rowidx = dec->h_rowidx;
    const size_t limit = (size_t)dec->max_smem_optin;
    for (int nc : {2, 4, 8}) {
        if (nc > CL_MAXC || m < nc * 32) continue;
        // slabs of equal size (a multiple of 32 rows keeps the row slices full)
        std::vector<int> lo(nc + 1);
        for (int c = 0; c <= nc; ++c) lo[c] = (int)(((long long)m * c / nc + 16) / 32 * 32);
        lo[0] = 0; lo[nc] = m;
        auto slab_of = [&](int r) { int c = 0; while (r >= lo[c + 1]) ++c; return c; };
        // ownership and straddlers
        std::vector<int> owner(n, -1), strad(n, 0);
        bool ok = true;
        for (int j = 0; j < n && ok; ++j) {
            if (colptr[j] == colptr[j + 1]) { owner[j] = 0; continue; }                // (rows ascending inside a column)
            const int cf = slab_of(rowidx[colptr[j]]), cl = slab_of(rowidx[colptr[j + 1] - 1]);
            owner[j] = cf;
            if (cl != cf) { strad[j] = 1; if (cl != cf + 1 || colptr[j + 1] - colptr[j] > CL_DMAX) ok = false; }
        }
        if (!ok) continue;
        {   // a straddler is processed by one of its two slabs: balance the variable phase (measured: a straddler costs about eight local columns)
            std::vector<double> load(nc, 0.0);
            const double str_cost = getenv("QLDPC_B200_CLUSTER_STRCOST") ? atof(getenv("QLDPC_B200_CLUSTER_STRCOST")) : 8.0;
            for (int j = 0; j < n; ++j) if (!strad[j]) load[owner[j]] += 1.0;
            for (int j = 0; j < n; ++j) {
                if (!strad[j]) continue;
                const int cf = owner[j];
                if (load[cf + 1] < load[cf]) owner[j] = cf + 1;
                load[owner[j]] += str_cost;
            }
        }
        // per rank: sub-graph = slab rows x (owned local variables + phantoms for every straddler touching the slab)
        std::vector<EdgeLayout> L(nc);
        std::vector<std::vector<int>> cols(nc);                  // sub-column -> global variable
        std::vector<std::map<long long, int>> edge_of(nc);      // (local row, global var) -> sub edge
        std::vector<std::vector<int>> sub_of(nc);               // global variable -> sub-column or -1
        std::vector<int> nstr(nc, 0);
        size_t smem = 0;
        for (int c = 0; c < nc && ok; ++c) {
            sub_of[c].assign(n, -1);
            const int mc = lo[c + 1] - lo[c];
            std::vector<int32_t> ip(mc + 1, 0), ix;
            std::vector<float> pr;
            std::vector<uint8_t> ph;
            for (int r = lo[c]; r < lo[c + 1]; ++r) {
                for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
                    const int j = indices[e];
                    if (sub_of[c][j] < 0) { sub_of[c][j] = (int)cols[c].size(); cols[c].push_back(j); pr.push_back(prior[j]); ph.push_back((uint8_t)strad[j]); }
                    edge_of[c][(long long)(r - lo[c]) * n + j] = (int)ix.size();
                    ix.push_back(sub_of[c][j]);
                }
                ip[r - lo[c] + 1] = (int)ix.size();
            }
            if (c == 0)      // variables without any row (all-zero columns) belong to rank 0 as ordinary degree-0 columns
                for (int j = 0; j < n; ++j) if (colptr[j] == colptr[j + 1]) { sub_of[c][j] = (int)cols[c].size(); cols[c].push_back(j); pr.push_back(prior[j]); ph.push_back(0); }
            // (24 rows per slice would give every warp of a 736-row slab one slice; measured slower: a row slice is latency
            //  bound, its time does not drop with the lane count)
            const int rps = getenv("QLDPC_B200_CLUSTER_RPS") ? atoi(getenv("QLDPC_B200_CLUSTER_RPS")) : 32;
            L[c] = build_edge_layout(mc, (int)cols[c].size(), ip.data(), ix.data(), pr.data(), 32, 0x9E3779B97F4A7C15ull + c, ph.data(), rps);
            if (!L[c].ok || !L[c].uniform_prior || L[c].n_csl > EDGE_MAX_CSL || L[c].e_words > 65535) { ok = false; break; }
            for (int j = 0; j < n; ++j) if (strad[j] && owner[j] == c) nstr[c]++;
            smem = std::max(smem, cluster_smem_bytes(L[c], nstr[c]));
        }
        if (getenv("QLDPC_B200_CLUSTER_INFO"))
            for (int c = 0; c < nc && ok; ++c)
                fprintf(stderr, "[qldpc_b200] cluster nc=%d rank %d: rows %d, columns %d (+%d phantom), straddlers owned %d, E %d words, idx %d words, row slices %d, column slices %d, smem %zu\n",
                        nc, c, lo[c + 1] - lo[c], L[c].n_csl * 32, (int)cols[c].size(), nstr[c], L[c].e_words, L[c].idx_words, L[c].n_rsl, L[c].n_csl, cluster_smem_bytes(L[c], nstr[c]));
        if (!ok || smem + 8192 > limit) continue;                  // 8 KB: static shared memory of the kernel (tables, priors)
        // ---- device tables ----
        ClusterPlan *p = new ClusterPlan();
        p->nc = nc; p->smem = smem; p->n = n; p->nw = g.nw;
        std::vector<ClusterRankDev> ranks(nc);
        int rc = QB_OK;
        for (int c = 0; c < nc && !rc; ++c) {
            EdgeLayout &Lc = L[c];
            EdgeDev &d = ranks[c].eg;
            d.n_rsl = Lc.n_rsl; d.n_csl = Lc.n_csl; d.e_words = Lc.e_words; d.e_dummy = Lc.e_dummy; d.idx_words = Lc.idx_words;
            d.nw = g.nw; d.n = g.n; d.mw = g.mw;
            std::vector<float> e0(Lc.e_words, INFINITY);
            for (int i = 0; i < Lc.e_words; ++i) {
                if (Lc.slot_var[i] >= 0) e0[i] = prior[cols[c][Lc.slot_var[i]]] + 0.0f;
                else if (Lc.slot_var[i] == -2) e0[i] = 0.f;
            }
            std::vector<uint16_t> row_gid(Lc.row_id.size(), 0xFFFFu), var_gid(Lc.var_id.size(), 0xFFFFu);
            for (size_t i = 0; i < Lc.row_id.size(); ++i) if (Lc.row_id[i] != 0xFFFFu) row_gid[i] = (uint16_t)(lo[c] + Lc.row_id[i]);
            for (size_t i = 0; i < Lc.var_id.size(); ++i) if (Lc.var_id[i] != 0xFFFFu) var_gid[i] = (uint16_t)cols[c][Lc.var_id[i]];
            const float *pf = nullptr; const uint32_t *pu = nullptr; const uint16_t *ph16 = nullptr; const int32_t *pi = nullptr; const uint8_t *p8 = nullptr;
            if (!rc) { rc = cup(p, e0, &pf); d.E0 = reinterpret_cast<const float4 *>(pf); }
            if (!rc) { rc = cup(p, Lc.col_idx, &pu); d.col_idx = pu; }
            if (!rc) { rc = cup(p, Lc.col_rowpos, &pu); d.col_rowpos = pu; }
            if (!rc) { rc = cup(p, Lc.rtask, &pu); d.rtask = reinterpret_cast<const uint2 *>(pu); }
            if (!rc) { rc = cup(p, Lc.ctask, &pu); d.ctask = reinterpret_cast<const uint2 *>(pu); }
            if (!rc) { rc = cup(p, row_gid, &ph16); d.row_id = ph16; }
            if (!rc) { rc = cup(p, Lc.row_pads, &ph16); d.row_pads = reinterpret_cast<const uint2 *>(ph16); }
            if (!rc) { rc = cup(p, var_gid, &ph16); d.var_id = ph16; }
            d.lane_prior = nullptr;
            if (!rc) { rc = cup(p, Lc.wr_ptr, &pi); d.wr_ptr = pi; }
            if (!rc) { rc = cup(p, Lc.wc_ptr, &pi); d.wc_ptr = pi; }
            if (!rc) { rc = cup(p, Lc.wc_cls, &p8); d.wc_cls = reinterpret_cast<const uint4 *>(p8); }
            if (!rc) { rc = cup(p, Lc.row_mask, &pu); d.row_mask = pu; }
            if (!rc) { rc = cup(p, Lc.col_sig, &pu); d.col_sig = pu; }
            std::vector<EdgePriors> prv(1);
            memset(&prv[0], 0, sizeof(EdgePriors));
            for (int t = 0; t < Lc.n_csl; ++t) prv[0].bits[t] = Lc.ctask[2 * t + 1];
            const EdgePriors *pp = nullptr;
            if (!rc) { rc = cup(p, prv, &pp); ranks[c].pri = pp; }
            // straddlers owned by this rank
            const int ns = nstr[c], nsp = (ns + 31) / 32 * 32;
            std::vector<uint32_t> tab((size_t)6 * std::max(1, nsp), 0u), sr((size_t)CL_DMAX * std::max(1, ns), 0u);
            int i = 0;
            for (int j = 0; j < n; ++j) {
                if (!strad[j] || owner[j] != c) continue;
                const int D = colptr[j + 1] - colptr[j];
                const int c_lo = slab_of(rowidx[colptr[j]]);
                uint32_t sig = 0u;
                int nf = 0;
                for (int k = 0; k < D; ++k) {
                    const int r = rowidx[colptr[j] + k];
                    const int cr = slab_of(r);
                    const EdgeLayout &Lr = L[cr];
                    const int e = edge_of[cr].at((long long)(r - lo[cr]) * n + j);
                    if (cr == c_lo) ++nf;
                    tab[(size_t)(k >> 1) * nsp + i] |= (Lr.edge_slot[e] & 0xFFFFu) << (16 * (k & 1));
                    sr[(size_t)k * ns + i] = Lr.row_pos[r - lo[cr]] | (cr == c ? 0u : CL_REMOTE);
                    sig ^= Lr.row_mask[Lr.row_pos[r - lo[cr]]] & 0xFFu;
                }
                tab[(size_t)4 * nsp + i] = (uint32_t)j | ((uint32_t)D << 16) | ((uint32_t)nf << 20) | ((c != c_lo ? 1u : 0u) << 23) | (sig << 24);
                const float pj = prior[j] + 0.0f;
                memcpy(&tab[(size_t)5 * nsp + i], &pj, 4);
                ++i;
            }
            ranks[c].n_str = ns; ranks[c].n_strp = nsp;
            if (!rc) { rc = cup(p, tab, &pu); ranks[c].str_tab = pu; }
            if (!rc) { rc = cup(p, sr, &pu); ranks[c].str_rowpos = pu; }
        }
        const ClusterRankDev *dr = nullptr;
        if (!rc) { rc = cup(p, ranks, &dr); p->d_ranks = const_cast<ClusterRankDev *>(dr); }
        if (!rc) { std::vector<int> z(1, 0); const int *pc = nullptr; rc = cup(p, z, &pc); p->d_counter = const_cast<int *>(pc); }
        if (rc) { cluster_plan_destroy(p); return rc; }
        *out = p;
        return QB_OK;
    }
    return QB_OK;
}

int cluster_plan_size(const ClusterPlan *p) { return p ? p->nc : 0; }

int launch_minsum_cluster(qb_decoder *dec, ClusterPlan *p, const MinsumLaunch &a, cudaStream_t st)
{
    constexpr int THREADS = 1024;
    QB_CUDA(cudaMemsetAsync(p->d_counter, 0, sizeof(int), st));
    QB_CUDA(cudaMemsetAsync(a.hard_bits, 0, (size_t)a.B * p->nw * sizeof(uint32_t), st));
    auto kern = minsum_cluster_kernel<THREADS>;
    QB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(dec->sm_count / p->nc * p->nc), 1, 1);
    cfg.blockDim = dim3(THREADS, 1, 1);
    cfg.dynamicSmemBytes = p->smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p->nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (p->max_clusters == 0) {          // clusters that can be resident at once (a cluster stays inside one GPC)
        int nclu = 0;
        QB_CUDA(cudaOccupancyMaxActiveClusters(&nclu, kern, &cfg));
        p->max_clusters = std::max(1, nclu);
        if (getenv("QLDPC_B200_CLUSTER_INFO")) fprintf(stderr, "[qldpc_b200] cluster plan: %d CTAs per cluster, %zu B shared memory per CTA, %d clusters resident\n", p->nc, p->smem, p->max_clusters);
    }
    const int clusters = std::max(1, std::min(a.B, p->max_clusters));
    cfg.gridDim = dim3((unsigned)(clusters * p->nc), 1, 1);
    const ClusterRankDev *ranks = p->d_ranks;
    int *counter = p->d_counter;
    QB_CUDA(cudaLaunchKernelEx(&cfg, kern, ranks, a, counter));
    return QB_OK;
}

}  // namespace qb
