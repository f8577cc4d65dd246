// Host-only: shared-memory layout of the per-edge min-sum kernel (see edge_layout.h).
#include "edge_layout.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>

namespace qb {

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed ? seed : 1) {}
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

// longest-processing-time schedule: returns for every warp the list of task ids (heaviest first)
std::vector<std::vector<int>> lpt(const std::vector<int> &cost, int nwarps)
{
    std::vector<int> order(cost.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<std::vector<int>> out(nwarps);
    std::vector<long> load(nwarps, 0);
    for (int t : order) {
        int best = 0;
        for (int w = 1; w < nwarps; ++w) if (load[w] < load[best]) best = w;
        out[best].push_back(t);
        load[best] += cost[t];
    }
    return out;
}

}  // namespace

EdgeLayout build_edge_layout(int m, int n, const int32_t *indptr, const int32_t *indices, const float *prior,
                             int nwarps, uint64_t seed, const uint8_t *phantom, int rows_per_slice)
{
    const int rcap = std::max(8, std::min(32, rows_per_slice));
    EdgeLayout L;
    L.m = m; L.n = n; L.nnz = m > 0 ? indptr[m] : 0; L.nwarps = nwarps;
    auto fail = [&](const char *why) { L.ok = false; L.why = why; return L; };
    if (m <= 0 || n <= 0) return fail("empty graph");
    if (n >= 65535 || m >= 65535) return fail("more than 65534 rows or columns");
    for (int j = 0; j < n; ++j) if (!std::isfinite(prior[j])) return fail("non-finite prior");
    const int nnz = L.nnz;

    // ---- column structure (rows ascending inside a column = the reference's summation order) ---------
    std::vector<int> colptr(n + 1, 0), col_edge(nnz);
    for (int e = 0; e < nnz; ++e) colptr[indices[e] + 1]++;
    for (int j = 0; j < n; ++j) colptr[j + 1] += colptr[j];
    {
        std::vector<int> fill(colptr.begin(), colptr.end() - 1);
        for (int r = 0; r < m; ++r) {
            for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
                if (e > indptr[r] && indices[e] == indices[e - 1]) return fail("duplicate entry in a row");
                col_edge[fill[indices[e]]++] = e;
            }
        }
    }
    std::vector<int> edge_row(nnz);
    for (int r = 0; r < m; ++r) for (int e = indptr[r]; e < indptr[r + 1]; ++e) edge_row[e] = r;

    // ---- row slices --------------------------------------------------------------------------------
    struct RSlice { std::vector<int> rows; int K; };
    std::vector<RSlice> rsl;
    {
        std::vector<int> order(m);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return indptr[a + 1] - indptr[a] > indptr[b + 1] - indptr[b]; });
        for (int r : order) {
            const int deg = indptr[r + 1] - indptr[r];
            const int K = deg == 0 ? 0 : (deg + 1 + 3) / 4;
            bool joined = false;
            if (!rsl.empty() && (int)rsl.back().rows.size() < rcap) {
                const int Ks = rsl.back().K;
                if ((Ks == 0 && deg == 0) || (Ks > 0 && deg > 0 && 4 * Ks - deg <= 4)) joined = true;
            }
            if (!joined) rsl.push_back(RSlice{{}, K});
            rsl.back().rows.push_back(r);
        }
    }
    int maxK = 0;
    for (auto &s : rsl) maxK = std::max(maxK, s.K);
    L.max_K = maxK;
    if (maxK > 16) return fail("row degree above 62");
    // ---- column slices -----------------------------------------------------------------------------
    std::vector<uint8_t> exact(n, 0);
    for (int r = 0; r < m; ++r) if (indptr[r + 1] - indptr[r] == 1) exact[indices[indptr[r]]] = 1;
    int max_cdeg = 0;
    for (int j = 0; j < n; ++j) max_cdeg = std::max(max_cdeg, colptr[j + 1] - colptr[j]);
    L.max_cdeg = max_cdeg;
    if (max_cdeg > 16) return fail("column degree above 16");
    {
        std::map<uint32_t, int> distinct;
        for (int j = 0; j < n; ++j) { uint32_t b; memcpy(&b, &prior[j], 4); distinct[b]++; }
        L.uniform_prior = distinct.size() <= 64;
    }
    struct CSlice { std::vector<int> vars; int deg; int exact; float prior; int cls; };
    std::vector<CSlice> csl;
    {
        std::vector<int> order;
        for (int j = 0; j < n; ++j) if (!phantom || !phantom[j]) order.push_back(j);      // phantom columns only reserve slots
        auto key = [&](int j) {
            uint32_t b = 0;
            if (L.uniform_prior) memcpy(&b, &prior[j], 4);
            return std::make_tuple((int)exact[j], -(colptr[j + 1] - colptr[j]), b);
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
        for (size_t i = 0; i < order.size(); ++i) {
            const int j = order[i];
            if (i == 0 || key(j) != key(order[i - 1]) || csl.back().vars.size() == 32)
                csl.push_back(CSlice{{}, colptr[j + 1] - colptr[j], exact[j], prior[j], 15});
            csl.back().vars.push_back(j);
        }
        for (auto &c : csl) {
            c.cls = 15;
            if ((c.vars.size() == 32 || c.prior >= 0.f) && L.uniform_prior) {
                if (!c.exact && c.deg <= 8) c.cls = c.deg;
                else if (c.exact && c.deg >= 1 && c.deg <= 6) c.cls = 8 + c.deg;
            }
        }
    }
    // ---- per-warp schedules, renumber slices in task order ---------------------------------------------
    {
        std::vector<int> cost(rsl.size());
        for (size_t i = 0; i < rsl.size(); ++i) cost[i] = rsl[i].K ? rsl[i].K * 4 + 2 : 0;
        auto sched = lpt(cost, nwarps);
        std::vector<RSlice> re;
        L.wr_ptr.assign(nwarps + 1, 0);
        for (int w = 0; w < nwarps; ++w) { for (int t : sched[w]) re.push_back(rsl[t]); L.wr_ptr[w + 1] = (int)re.size(); }
        rsl.swap(re);
    }
    {
        std::vector<int> cost(csl.size());
        // fast-path cost of a slice of degree D = max(instructions, 4 x shared-memory instructions) of its loop body in
        // the float32 kernel (odd degrees carry the fingerprint in the index word and are cheaper)
        static const int fast_cost[9] = {10, 16, 25, 28, 40, 48, 60, 64, 76};
        for (size_t i = 0; i < csl.size(); ++i)
            cost[i] = csl[i].cls == 15 ? (csl[i].deg * 8 + 8) * 2 : fast_cost[std::min(8, csl[i].deg)] + (csl[i].exact ? csl[i].deg : 0);
        // (cutting the class-sorted list into contiguous equal-cost pieces -- long runs of one class per warp -- was measured
        // 1.5 % slower than the LPT mix)
        auto sched = lpt(cost, nwarps);
        std::vector<CSlice> re;
        L.wc_ptr.assign(nwarps + 1, 0);
        L.wc_cls.assign((size_t)nwarps * 16, 0);
        for (int w = 0; w < nwarps; ++w) {
            std::stable_sort(sched[w].begin(), sched[w].end(), [&](int a, int b) { return csl[a].cls < csl[b].cls; });
            for (int t : sched[w]) {
                re.push_back(csl[t]);
                if (L.wc_cls[(size_t)w * 16 + csl[t].cls] == 255) return fail("too many column slices per warp");
                L.wc_cls[(size_t)w * 16 + csl[t].cls]++;
            }
            L.wc_ptr[w + 1] = (int)re.size();
            if (L.wc_ptr[w + 1] - L.wc_ptr[w] > 32) return fail("more than 32 column slices per warp");
        }
        csl.swap(re);
    }
    L.n_rsl = (int)rsl.size(); L.n_csl = (int)csl.size();
    if (L.n_rsl * 32 >= 65535) return fail("too many row slices");

    // ---- E geometry ------------------------------------------------------------------------------------
    std::vector<int> rs_base(L.n_rsl), rs_stride(L.n_rsl), row_slice(m, -1), row_lane(m, -1);
    int words = 0;
    for (int t = 0; t < L.n_rsl; ++t) {
        const int nl = (int)rsl[t].rows.size();
        rs_stride[t] = 8 * ((nl + 7) / 8) + 1;
        rs_base[t] = words;
        words += rsl[t].K * rs_stride[t] * 4;
        for (int l = 0; l < nl; ++l) { row_slice[rsl[t].rows[l]] = t; row_lane[rsl[t].rows[l]] = l; }
    }
    L.e_dummy = std::max(4, words);
    L.e_words = L.e_dummy + 32;
    if (L.e_words >= 65000) return fail("edge array above 65000 words");
    auto slot_word = [&](int t, int lane, int slot) { return rs_base[t] + ((slot >> 2) * rs_stride[t] + lane) * 4 + (slot & 3); };

    // ---- groups: the k-th edges of the variables of one column slice are gathered by one instruction ----
    std::vector<int> cgroup_base(L.n_csl + 1, 0);
    for (int t = 0; t < L.n_csl; ++t) cgroup_base[t + 1] = cgroup_base[t] + csl[t].deg;
    const int n_groups = cgroup_base[L.n_csl];
    std::vector<int> edge_group(nnz, -1);
    for (int t = 0; t < L.n_csl; ++t)
        for (int j : csl[t].vars)
            for (int p = colptr[j]; p < colptr[j + 1]; ++p) edge_group[col_edge[p]] = cgroup_base[t] + (p - colptr[j]);

    // ---- colouring: slot of every edge inside its row ------------------------------------------------------
    std::vector<int> slot_of(nnz, -1);
    std::vector<int> owner((size_t)m * maxK * 4, -1);                       // [row][slot] -> edge
    std::vector<uint8_t> gcnt((size_t)std::max(1, n_groups) * 32, 0);       // [group][bank]
    auto bank_of = [&](int r, int slot) { return slot_word(row_slice[r], row_lane[r], slot) & 31; };
    Rng rng(seed);
    {   // greedy, group by group
        std::vector<int> by_group(nnz);
        std::iota(by_group.begin(), by_group.end(), 0);
        // (edges of phantom columns belong to no gather group: they take whatever slot is left, after the others)
        std::stable_sort(by_group.begin(), by_group.end(), [&](int a, int b) {
            return (edge_group[a] < 0 ? (1 << 30) : edge_group[a]) < (edge_group[b] < 0 ? (1 << 30) : edge_group[b]); });
        for (int e : by_group) {
            const int r = edge_row[e], g = edge_group[e];
            const int nslots = rsl[row_slice[r]].K * 4;
            int best = -1, best_cost = 1 << 30;
            const int start = rng.below(nslots);
            for (int q = 0; q < nslots; ++q) {
                const int s = (start + q) % nslots;
                if (owner[(size_t)r * maxK * 4 + s] >= 0) continue;
                const int c = g < 0 ? 0 : gcnt[(size_t)g * 32 + bank_of(r, s)];
                if (c < best_cost) { best_cost = c; best = s; if (c == 0) break; }
            }
            if (best < 0) return fail("internal: no free slot");
            slot_of[e] = best; owner[(size_t)r * maxK * 4 + best] = e;
            if (g >= 0) gcnt[(size_t)g * 32 + bank_of(r, best)]++;
        }
    }
    auto n_conflicted = [&]() {
        int c = 0;
        for (int e = 0; e < nnz; ++e) if (edge_group[e] >= 0 && gcnt[(size_t)edge_group[e] * 32 + bank_of(edge_row[e], slot_of[e])] > 1) ++c;
        return c;
    };
    {   // min-conflicts local search: move a conflicted edge to another slot of its row (swapping with the occupant)
        std::vector<int> bad;
        for (int pass = 0; pass < 200; ++pass) {
            bad.clear();
            for (int e = 0; e < nnz; ++e) if (edge_group[e] >= 0 && gcnt[(size_t)edge_group[e] * 32 + bank_of(edge_row[e], slot_of[e])] > 1) bad.push_back(e);
            if (bad.empty()) break;
            for (size_t i = bad.size(); i > 1; --i) std::swap(bad[i - 1], bad[rng.below((uint32_t)i)]);
            for (int e : bad) {
                const int r = edge_row[e], g = edge_group[e], s = slot_of[e], b = bank_of(r, s);
                if (gcnt[(size_t)g * 32 + b] <= 1) continue;                // fixed meanwhile
                const int nslots = rsl[row_slice[r]].K * 4;
                int best = -1, best_delta = 1 << 30, nbest = 0;
                for (int s2 = 0; s2 < nslots; ++s2) {
                    if (s2 == s) continue;
                    const int b2 = bank_of(r, s2);
                    if (b2 == b) continue;
                    const int e2 = owner[(size_t)r * maxK * 4 + s2];
                    int delta = (gcnt[(size_t)g * 32 + b2] >= 1 ? 1 : 0) - 1;
                    if (e2 >= 0 && edge_group[e2] >= 0) {
                        const int g2 = edge_group[e2];
                        if (g2 == g) continue;                               // same multiset of banks
                        delta += (gcnt[(size_t)g2 * 32 + b] >= 1 ? 1 : 0) - (gcnt[(size_t)g2 * 32 + b2] >= 2 ? 1 : 0);
                    }
                    if (delta < best_delta) { best_delta = delta; best = s2; nbest = 1; }
                    else if (delta == best_delta && rng.below(++nbest) == 0) best = s2;
                }
                if (best < 0 || best_delta > 0) continue;
                if (best_delta == 0 && (rng.next() & 3) == 0) continue;     // plateau moves most of the time
                const int s2 = best, b2 = bank_of(r, s2), e2 = owner[(size_t)r * maxK * 4 + s2];
                gcnt[(size_t)g * 32 + b]--; gcnt[(size_t)g * 32 + b2]++;
                slot_of[e] = s2; owner[(size_t)r * maxK * 4 + s2] = e; owner[(size_t)r * maxK * 4 + s] = e2;
                if (e2 >= 0) {
                    const int g2 = edge_group[e2];
                    if (g2 >= 0) { gcnt[(size_t)g2 * 32 + b2]--; gcnt[(size_t)g2 * 32 + b]++; }
                    slot_of[e2] = s;
                }
            }
        }
    }
    L.conflict_pairs = n_conflicted();
    L.gather_groups = n_groups;
    L.gather_wavefronts = 0;
    for (int g = 0; g < n_groups; ++g) {
        int mx = 1;
        for (int b = 0; b < 32; ++b) mx = std::max<int>(mx, gcnt[(size_t)g * 32 + b]);
        L.gather_wavefronts += mx;
    }

    // ---- emit tables -----------------------------------------------------------------------------------
    L.rtask.assign((size_t)L.n_rsl * 2, 0u);
    L.row_id.assign((size_t)L.n_rsl * 32, 0xFFFFu);
    L.row_pads.assign((size_t)L.n_rsl * 32 * 4, 0xFFFFu);
    L.slot_var.assign(L.e_words, -1);
    for (int i = L.e_dummy; i < L.e_words; ++i) L.slot_var[i] = -2;
    for (int t = 0; t < L.n_rsl; ++t) {
        const int nl = (int)rsl[t].rows.size();
        bool onepad = rsl[t].K > 0 && rsl[t].K <= 9;
        for (int r : rsl[t].rows) if (indptr[r + 1] - indptr[r] != 4 * rsl[t].K - 1) onepad = false;
        L.rtask[2 * t] = (uint32_t)rs_base[t] | (onepad ? 1u : 0u);           // (the base is a multiple of 4 words; bit 0: one unused slot per row)
        L.rtask[2 * t + 1] = (uint32_t)rsl[t].K | ((uint32_t)nl << 8) | ((uint32_t)rs_stride[t] << 16);
        for (int l = 0; l < nl; ++l) {
            const int r = rsl[t].rows[l];
            L.row_id[t * 32 + l] = (uint16_t)r;
            int np = 0;
            for (int s = 0; s < rsl[t].K * 4; ++s) {
                const int e = owner[(size_t)r * maxK * 4 + s];
                if (e >= 0) L.slot_var[slot_word(t, l, s)] = indices[e];
                else {
                    if (np >= 4) return fail("internal: more than 4 unused slots in a row");
                    L.row_pads[((size_t)t * 32 + l) * 4 + np++] = (uint16_t)slot_word(t, l, s);
                }
            }
        }
    }
    L.edge_slot.assign(nnz, 0u);
    L.row_pos.assign(m, 0u);
    for (int r = 0; r < m; ++r) L.row_pos[r] = (uint32_t)(row_slice[r] * 32 + row_lane[r]);
    for (int e = 0; e < nnz; ++e) L.edge_slot[e] = (uint32_t)slot_word(row_slice[edge_row[e]], row_lane[edge_row[e]], slot_of[e]);
    L.row_mask.assign((size_t)L.n_rsl * 32, 0u);
    {
        Rng mr(seed ^ 0xD1B54A32D192ED03ull);
        for (int t = 0; t < L.n_rsl; ++t)
            for (size_t l = 0; l < rsl[t].rows.size(); ++l) L.row_mask[t * 32 + l] = mr.next() | 1u;
    }
    L.col_sig.assign((size_t)L.n_csl * 32, 0u);
    L.ctask.assign((size_t)L.n_csl * 2, 0u);
    L.var_id.assign((size_t)(L.n_csl + 1) * 32, 0xFFFFu);      // one slice of padding: the kernel reads one slice ahead
    if (!L.uniform_prior) L.lane_prior.assign((size_t)L.n_csl * 32, 0.f);
    int iw = 0;
    std::vector<int> cs_base(L.n_csl);
    for (int t = 0; t < L.n_csl; ++t) { cs_base[t] = iw; iw += ((csl[t].deg + 1) / 2) * 32; }
    L.idx_words = std::max(32, iw);
    if (L.idx_words / 32 >= 65535) return fail("index table too large");
    L.col_idx.assign(L.idx_words, 0xFFFFFFFFu);
    L.col_rowpos.assign(L.idx_words, 0xFFFFFFFFu);
    for (int t = 0; t < L.n_csl; ++t) {
        const int nl = (int)csl[t].vars.size();
        uint32_t pb; memcpy(&pb, &csl[t].prior, 4);
        L.ctask[2 * t] = (uint32_t)(cs_base[t] / 32) | ((uint32_t)csl[t].deg << 16) | ((uint32_t)nl << 22) | ((uint32_t)csl[t].exact << 28);
        L.ctask[2 * t + 1] = pb;
        for (int l = 0; l < nl; ++l) {
            const int j = csl[t].vars[l];
            L.var_id[t * 32 + l] = (uint16_t)j;
            if (!L.uniform_prior) L.lane_prior[t * 32 + l] = prior[j];
            for (int p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int k = p - colptr[j], e = col_edge[p], r = edge_row[e];
                const uint32_t slot = (uint32_t)slot_word(row_slice[r], row_lane[r], slot_of[e]);
                const uint32_t rpos = (uint32_t)(row_slice[r] * 32 + row_lane[r]);
                L.col_sig[t * 32 + l] ^= L.row_mask[rpos];
                const int H = (csl[t].deg + 1) / 2;
                uint32_t &wi = L.col_idx[cs_base[t] + edge_idx_off(H, k >> 1, l)];
                uint32_t &wr = L.col_rowpos[cs_base[t] + edge_idx_off(H, k >> 1, l)];
                if (k & 1) { wi = (wi & 0x0000FFFFu) | (slot << 16); wr = (wr & 0x0000FFFFu) | (rpos << 16); }
                else { wi = (wi & 0xFFFF0000u) | slot; wr = (wr & 0xFFFF0000u) | rpos; }
            }
        }
        if (csl[t].cls != 15) {          // dummy lanes of a partial slice on the fast path
            const int H = (csl[t].deg + 1) / 2;
            for (int l = nl; l < 32; ++l)
                for (int kk = 0; kk < H; ++kk) {
                    const uint32_t slot = (uint32_t)(L.e_dummy + l);
                    L.col_idx[cs_base[t] + edge_idx_off(H, kk, l)] = slot | (slot << 16);
                }
        }
        // odd degree on the fast path: the upper half of a lane's last index word is free and carries EDGE_SIG_TAG | the
        // variable's 8-bit fingerprint, so the float32 kernel needs no separate fingerprint load
        if ((csl[t].deg & 1) && csl[t].cls != 15) {
            const int H = (csl[t].deg + 1) / 2;
            for (int l = 0; l < 32; ++l) {
                uint32_t &wi = L.col_idx[cs_base[t] + edge_idx_off(H, H - 1, l)];
                wi = (wi & 0x0000FFFFu) | ((EDGE_SIG_TAG | (l < nl ? (L.col_sig[t * 32 + l] & 0xFFu) : 0u)) << 16);
            }
        }
    }
    L.ok = true;
    return L;
}

}  // namespace qb

// ---- C entry point for host-side tests (no CUDA call): statistics of the layout of a graph ---------------
// stats_out[16] = { ok, n_rsl, n_csl, e_words, idx_words, gather_groups, gather_wavefronts, conflict_pairs,
//                   uniform_prior, max_K, max_cdeg, every edge in exactly one slot (1/0), 0... }
extern "C" int qb_edge_layout_probe(int32_t m, int32_t n, const int32_t *indptr, const int32_t *indices,
                                    const double *prior, int32_t nwarps, int64_t *stats_out)
{
    std::vector<float> pf(n > 0 ? n : 0);
    for (int j = 0; j < n; ++j) pf[j] = (float)prior[j];
    const qb::EdgeLayout L = qb::build_edge_layout(m, n, indptr, indices, pf.data(), nwarps);
    for (int i = 0; i < 16; ++i) stats_out[i] = 0;
    stats_out[0] = L.ok;
    if (!L.ok) return 0;
    stats_out[1] = L.n_rsl; stats_out[2] = L.n_csl; stats_out[3] = L.e_words; stats_out[4] = L.idx_words;
    stats_out[5] = L.gather_groups; stats_out[6] = L.gather_wavefronts; stats_out[7] = L.conflict_pairs;
    stats_out[8] = L.uniform_prior; stats_out[9] = L.max_K; stats_out[10] = L.max_cdeg;
    // consistency: every (variable, k) entry of the index table points at a slot owned by that variable, once
    std::vector<int> seen(L.e_words, 0);
    bool good = true;
    long cnt = 0;
    for (int t = 0; t < L.n_csl; ++t) {
        const int base = (int)(L.ctask[2 * t] & 0xFFFFu) * 32, deg = (int)((L.ctask[2 * t] >> 16) & 63u), nl = (int)((L.ctask[2 * t] >> 22) & 63u);
        for (int l = 0; l < nl; ++l)
            for (int k = 0; k < deg; ++k) {
                const uint32_t w = L.col_idx[base + qb::edge_idx_off((deg + 1) / 2, k >> 1, l)];
                const int slot = (k & 1) ? (int)(w >> 16) : (int)(w & 0xFFFFu);
                if (slot >= L.e_words || L.slot_var[slot] != (int)L.var_id[t * 32 + l] || seen[slot]++) good = false;
                ++cnt;
            }
    }
    if (cnt != L.nnz) good = false;
    // odd-degree slices on the fast path carry EDGE_SIG_TAG | fingerprint in the upper half of every lane's last index word
    // (dummy lanes: fingerprint 0); no other word has an upper half in the tag range
    for (int t = 0; t < L.n_csl; ++t) {
        const int base = (int)(L.ctask[2 * t] & 0xFFFFu) * 32, deg = (int)((L.ctask[2 * t] >> 16) & 63u), nl = (int)((L.ctask[2 * t] >> 22) & 63u);
        const int H = (deg + 1) / 2;
        // (fast path = what build_edge_layout classified below class 15: full or non-negative prior, uniform priors, degree
        // <= 8, or <= 6 next to a degree-1 row; recognised here by the tag itself on lane 0)
        const bool tagged = (deg & 1) && ((L.col_idx[base + qb::edge_idx_off(H, H - 1, 0)] >> 16) & 0xFF00u) == qb::EDGE_SIG_TAG;
        for (int l = 0; l < 32; ++l)
            for (int kk = 0; kk < H; ++kk) {
                const uint32_t hi = L.col_idx[base + qb::edge_idx_off(H, kk, l)] >> 16;
                const bool in_range = (hi & 0xFF00u) == qb::EDGE_SIG_TAG;
                if (tagged && kk == H - 1) {
                    if (hi != (qb::EDGE_SIG_TAG | (l < nl ? (L.col_sig[t * 32 + l] & 0xFFu) : 0u))) good = false;
                } else if (in_range) good = false;
            }
        if ((deg & 1) && deg <= 8 && nl == 32 && L.uniform_prior && !tagged && !((L.ctask[2 * t] >> 28) & 1u)) good = false;   // a full plain slice must be tagged
    }
    stats_out[11] = good;
    return 0;
}
