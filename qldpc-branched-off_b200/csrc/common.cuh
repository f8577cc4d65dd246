// Internal declarations shared by the .cu files of libqldpc_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/qldpc_b200.h"

namespace qb {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define QB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return qb::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define QB_REQUIRE(cond, msg)                 \
    do {                                      \
        if (!(cond)) {                        \
            qb::set_error(msg);               \
            return QB_ERR_ARG;                \
        }                                     \
    } while (0)

// Grow-only device scratch buffer owned by a handle.
struct Scratch {
    void *ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(ptr); }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int MS_THREADS = 512;      // min-sum CTA size (16 warps)
constexpr int MS_MAX_ROW_DEG = 56;   // 56 sign bits + 6-bit argmin must fit two 32-bit words
constexpr int OSD_THREADS = 128;      // four warps per failed side
constexpr int OSD_MAX_WPL = 4;       // syndrome words per lane -> m <= 4096

// Device view of one decoding side (what kernels receive by value).
struct GraphDev {
    int m, n, nnz, k, mw, nw, m_pad, n_pad;
    int n_rslices, n_cslices;
    // sliced, chunked ELL: 32 rows (columns) per slice; slice s owns uint4 words [ptr[s], ptr[s+1]);
    // word ptr[s] + 32*c + lane holds entries 8c..8c+7 (rows: uint16 column index, padding = n_pad, the
    // dummy +inf variable) resp. 4c..4c+3 (columns: uint32 check << 8 | position in row, padding =
    // m_pad << 8, the dummy all-zero check) of the
    // lane's row / column, so a warp reads one coalesced 512-byte line per chunk.
    const int32_t *rslice_ptr;
    const uint4 *row_ell4;
    const uint8_t *rslice_exact;               // slice contains a row of degree 1 (inf messages possible)
    const uint8_t *rslice_deg, *cslice_deg;    // largest row / column degree inside the slice
    int nan_anywhere;                          // non-finite priors: every slice takes the exact path
    const int32_t *cslice_ptr;
    const uint4 *col_ell4;
    // plain CSR / CSC (general kernels, OSD)
    const int32_t *indptr, *indices;           // CSR
    const int32_t *colptr, *rowidx, *csc_edge; // CSC: row index and CSR edge id of each entry
    const uint4 *colsig;                       // per column: up to 8 row indices as uint16 (0xFFFF = none); NULL if a column has > 8
    const float *prior;
    const uint32_t *logmask;                   // per column: bit b = logical row b contains the column
};

struct EdgePlan;   // per-edge min-sum kernel: layout + device tables (minsum_edge.cu)
struct EdgePlanH2; // its packed half2 companion (minsum_edge_h2.cu)
struct ClusterPlan; // per-edge kernel on a thread-block cluster for graphs larger than one SM (minsum_edge_cluster.cu)
}  // namespace qb

struct qb_decoder {
    int device = 0;
    qb::GraphDev g{};
    int max_row_deg = 0, max_col_deg = 0;
    bool fast_ok = false;       // fast (on-chip message) min-sum kernel applicable
    int graph_nan = 0;          // graph structure alone can produce NaN posteriors
    std::vector<int32_t> h_indptr, h_indices, h_colptr, h_rowidx;
    std::vector<void *> owned;  // device allocations freed on destroy
    float *d_prior = nullptr;
    float *d_alpha = nullptr;   // per-iteration alpha, device
    int alpha_cap = 0;
    qb::Scratch scratch;        // host-API staging
    qb::Scratch work;           // kernel workspaces (general min-sum messages, OSD spill)
    qb::Scratch ovf;            // OSD: workspaces of the free-row path (candidate lists, records, overflow queue)
    int sm_count = 148;
    int max_smem_optin = 0;
    qb::EdgePlan *edge = nullptr;   // nullptr: graph does not fit the per-edge kernel
    qb::EdgePlanH2 *edge_h2 = nullptr;   // packed mode, built on first use
    qb::ClusterPlan *cluster = nullptr;  // only when edge == nullptr: the graph cut into slabs over a cluster of CTAs
    std::vector<float> h_prior;     // float priors as the kernels see them
    int precision = 0;              // QB_PRECISION_* used by the host-buffer entry points (qb_decoder_set_precision)
};

struct qb_sampler {
    int device = 0;
    int L = 0, k = 0;
    int mZ = 0, nZ = 0, mX = 0, nX = 0, mwZ = 0, mwX = 0;
    int8_t *d_kind = nullptr;
    int32_t *d_colZ = nullptr, *d_colX = nullptr;        // [L][4]
    // column signatures: 8 x uint16 per column: up to 7 rows (0xFFFF = none) ... or CSC when wider
    int32_t *d_cpZ = nullptr, *d_rowZ = nullptr, *d_cpX = nullptr, *d_rowX = nullptr;
    uint32_t *d_lmZ = nullptr, *d_lmX = nullptr;
    uint32_t *d_geo = nullptr;          // jump table of the fault sampler for error rate geo_p (sampler.cu)
    std::vector<uint32_t> h_geo;
    double geo_p = -1.0;
    std::vector<void *> owned;
    qb::Scratch scratch;
    int sm_count = 148;
};

namespace qb {

// ---- launchers implemented in the individual .cu files --------------------------------------------
struct MinsumLaunch {
    const uint32_t *syn_bits;
    int B, max_iter;
    const float *alpha_d;       // [max_iter]
    float damping, clip;
    int dense_variant;
    uint32_t *hard_bits;
    uint8_t *converged;
    int32_t *final_iter;
    float *post;                // nullable [B][n]
    int post_failed_only;       // write post only for non-converged shots
    int32_t *fail_count;        // nullable: device counter, non-converged shots are appended
    int32_t *fail_idx;
    int32_t *fail_wt;           // nullable: residual syndrome weight of each appended shot (OSD scheduling hint)
    int precision;              // QB_PRECISION_F32 (default) / QB_PRECISION_HALF2 (packed mode, opt-in)
};
int launch_minsum(qb_decoder *dec, const MinsumLaunch &a, cudaStream_t st);
int edge_plan_create(const qb_decoder *dec, const float *prior_h, EdgePlan **out);
void edge_plan_destroy(EdgePlan *p);
int launch_minsum_edge(qb_decoder *dec, EdgePlan *p, const MinsumLaunch &a, cudaStream_t st);
int cluster_plan_create(const qb_decoder *dec, const float *prior_h, ClusterPlan **out);
void cluster_plan_destroy(ClusterPlan *p);
int cluster_plan_size(const ClusterPlan *p);
int launch_minsum_cluster(qb_decoder *dec, ClusterPlan *p, const MinsumLaunch &a, cudaStream_t st);
int upload_alpha(qb_decoder *dec, int max_iter, int alpha_mode, double alpha, const double *seq, int len,
                 cudaStream_t st);

struct OsdLaunch {
    const uint32_t *syn_bits;   // [B][mw]
    uint32_t *hard_bits;        // [B][nw] in/out
    const float *post;          // [B][n] (used when ordering == nullptr)
    const int32_t *ordering;    // nullable [B][n]: caller-supplied column orders
    const int32_t *fail_idx;    // nullable: indices into the batch; nullptr = all of 0..F-1
    int F;                      // number of sides (upper bound when n_fail_d given)
    const int32_t *n_fail_d;    // nullable device count
    int32_t *rank_out;          // nullable [B]: pivots used | rank_tag
    int rank_tag;               // OR-ed into rank_out (the pipeline marks which kernel solved the side: path << 16)
    int32_t *pivots_out;        // nullable [B][min(m,n)]
    int exact_rows;             // emulate the reference's pivot-row order (inconsistent syndromes)
};
int launch_osd0(qb_decoder *dec, const OsdLaunch &a, cudaStream_t st);
// free-row elimination path (osd_free.cu): applicable to graphs with column signatures (column degree <= 8)
bool osd_free_applicable(const qb_decoder *dec);
int launch_osd0_free(qb_decoder *dec, const OsdLaunch &a, int32_t **ovf_count_d, int32_t **ovf_idx_d, cudaStream_t st);
int osd_free_stats(qb_decoder *dec, int32_t *out10);
// kernels one launch_osd0 call issues for a pipeline queue (gpu_launches accounting)
int osd_launches_per_call(const qb_decoder *dec);
// order the failure queue by descending residual weight (longest elimination first)
int launch_sort_failures(const int32_t *fail_idx, const int32_t *fail_wt, const int32_t *n_fail_d, int32_t *sorted_idx, cudaStream_t st);

int launch_events_syndrome(qb_sampler *s, const int32_t *ev_ptr_d, const uint32_t *events_d, int B,
                           uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX, cudaStream_t st);
int launch_sample_syndrome(qb_sampler *s, uint64_t seed, uint64_t first_shot, int B, double p,
                           uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX,
                           int32_t *nfaults, cudaStream_t st);

void geometric_table(double p, std::vector<uint32_t> &T);
int geometric_table_size();

// bit packing helpers (device kernels, utils.cu)
int launch_pack_bits(const int8_t *src, int B, int len, uint32_t *dst, int words, cudaStream_t st);
int launch_unpack_bits(const uint32_t *src, int B, int len, int words, int8_t *dst, cudaStream_t st);
int launch_logical_check(const qb_decoder *dec, const uint32_t *hard_bits, const uint32_t *true_mask, int B,
                         uint8_t *flags, int flag_bit, int64_t *counts, int count_slot, cudaStream_t st);

}  // namespace qb
