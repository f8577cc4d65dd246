// Host-side layout builder for the per-edge min-sum kernel (minsum_edge.cu).
//
// The kernel keeps one float per Tanner-graph edge in shared memory ("E"), ordered so that the check-node
// phase streams it with conflict-free 128-bit accesses and the variable-node phase gathers / scatters it
// through a precomputed slot index per edge.  Because the graph is fixed per decoder, the slot of every edge
// inside its row is chosen here, once, so that the 32 gathers of a warp instruction fall into 32 different
// shared-memory banks (a constrained bipartite edge colouring, solved greedily + min-conflicts local search).
//
// Layout of E (32-bit words):
//   row slice t (<= 32 check rows, one per lane, sorted by degree):  base[t] + (c * stride + lane) * 4 + i
//   for chunk c < K (K = ceil((max degree + 1) / 4)), i < 4; stride = 8 * ceil(lanes / 8) + 1 sixteen-byte units,
//   so the 8 lanes of a quarter warp cover the 32 banks and the bank of a slot is ((c + lane) * 4 + i) mod 32
//   up to the base: every row sees every bank.  Unused slots (>= 1, <= 4 per row) hold +inf.
// Column slices hold <= 32 variables of equal degree (and equal prior when priors take few distinct values);
// their slot indices are stored as uint16 pairs: idx[base + edge_idx_off(H, kk, lane)] = slot(2kk) | slot(2kk+1) << 16.
// A slice with fewer than 32 variables and a non-negative prior is filled with dummy lanes whose indices all point
// at a private dummy word behind the row slices (value 0 at start, never negative), so that it takes the fast path.
// Slices are numbered in per-warp task order (LPT schedule), warp w owns [wr_ptr[w], wr_ptr[w+1]) and
// [wc_ptr[w], wc_ptr[w+1]).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#if defined(__CUDACC__)
#define QB_HD __host__ __device__
#else
#define QB_HD
#endif

namespace qb {

// Word offset, inside the index block of a column slice with H = ceil(degree / 2) index words per variable, of
// word kk of the variable in `lane`: the words are stored two per lane (one LDS.64 fetches 4 slot indices), an odd
// last word one per lane.
QB_HD inline int edge_idx_off(int H, int kk, int lane)
{
    return (kk >> 1) * 64 + ((kk == H - 1 && (H & 1)) ? lane : lane * 2 + (kk & 1));
}

// Upper half of the last index word of an odd-degree variable in a fast-path slice: EDGE_SIG_TAG | 8-bit fingerprint
// (0xFE00 .. 0xFEFF).  Shared memory has fewer than 0xFE00 words, so no absolute slot address reaches the tag (the
// kernels add their E base to slot halves only); an unused half stays 0xFFFF.
constexpr uint32_t EDGE_SIG_TAG = 0xFE00u;

struct EdgeLayout {
    bool ok = false;
    std::string why;                     // reason when !ok
    int m = 0, n = 0, nnz = 0;
    int nwarps = 0;
    int n_rsl = 0, n_csl = 0;
    int e_words = 0;                     // size of E in words (multiple of 4), including 32 dummy words at the end
    int e_dummy = 0;                     // first dummy word
    int idx_words = 0;                   // size of the column index table in words (multiple of 32)
    int max_K = 0, max_cdeg = 0;
    bool uniform_prior = false;          // every column slice has one prior value (else per-lane priors, lane_prior)
    // row slices (task order)
    std::vector<uint32_t> rtask;         // [n_rsl][2]: {base word, K | lanes << 8 | stride << 16}
    std::vector<uint16_t> row_id;        // [n_rsl*32] original row (0xFFFF = inactive lane)
    std::vector<uint16_t> row_pads;      // [n_rsl*32][4] absolute word index of the row's unused slots (0xFFFF = none)
    // column slices (task order)
    std::vector<uint32_t> ctask;         // [n_csl][2]: {idx base / 32 | deg << 16 | lanes << 22 | exact << 28, prior bits}
    std::vector<uint16_t> var_id;        // [n_csl*32] original variable (0xFFFF = inactive lane)
    std::vector<uint32_t> col_idx;       // [idx_words] slot pairs
    std::vector<uint32_t> col_rowpos;    // [idx_words] permuted row position (slice*32+lane) pairs, same layout
    std::vector<float> lane_prior;       // [n_csl*32] (only when !uniform_prior)
    std::vector<int32_t> slot_var;       // [e_words] variable of each slot, -1 = unused (+inf), -2 = dummy (0)
    std::vector<uint32_t> edge_slot;     // [nnz] word of E that holds the edge (CSR edge order)
    std::vector<uint32_t> row_pos;       // [m] permuted position (slice * 32 + lane) of every row
    std::vector<int32_t> wr_ptr, wc_ptr; // [nwarps+1]
    // column slices of a warp are sorted by class: 0..8 = full slice of degree 0..8, uniform prior;
    // 9..14 = same for variables next to a degree-1 row (NaN handling), degree 1..6; 15 = everything else
    std::vector<uint8_t> wc_cls;         // [nwarps][16] number of slices per class
    // random 32-bit word per row and their XOR over the rows of each column: a linear fingerprint of H.x, so that
    // "H.hard == syndrome" is first tested on 32 bits and only confirmed exactly when the fingerprints agree
    std::vector<uint32_t> row_mask;      // [n_rsl*32]
    std::vector<uint32_t> col_sig;       // [n_csl*32]
    // quality of the colouring: extra shared-memory wavefronts per iteration caused by bank conflicts
    int gather_groups = 0;               // warp-level gather instructions per iteration
    int gather_wavefronts = 0;           // sum over groups of the largest bank multiplicity
    int conflict_pairs = 0;
};

// prior: float priors (finite).  slack: extra free slots per row beyond the mandatory one.
// phantom (nullable, [n]): columns that only reserve a slot per edge in their rows -- their variable is processed by
// another CTA of a cluster or by the cluster kernel's own list (minsum_edge_cluster.cu) -- and get no column slice.
// rows_per_slice (8..32): fewer rows per slice spread a small number of rows over all warps.
EdgeLayout build_edge_layout(int m, int n, const int32_t *indptr, const int32_t *indices, const float *prior,
                             int nwarps, uint64_t seed = 0x9E3779B97F4A7C15ull, const uint8_t *phantom = nullptr,
                             int rows_per_slice = 32);

}  // namespace qb
