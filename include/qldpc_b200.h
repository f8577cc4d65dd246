/*
 * qldpc_b200.h -- C ABI of libqldpc_b200.so: the B200 (sm_100a) implementation of the
 * Monte-Carlo decoding hot path of michelebanfi/qLDPC-branched-off.
 *
 * The reference is pure Python + numba and has no FFI layer; its boundary for this path is the
 * Python call surface listed in SURVEY.md section 8(b).  Each entry point below names the
 * reference function (file:line, relative to the reference root) whose work it performs.
 * The Python package qldpc-branched-off_b200/ binds these with ctypes
 * and re-exposes the reference's own function names and signatures.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative qb_status and
 *     leaves a message for qb_last_error() (thread-local);
 *   - pointer suffix _h = host memory, _d = device memory on the handle's device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);  *_host entry points
 *     copy in, compute, copy out and synchronise before returning;
 *   - bit-packed vectors are little-endian uint32 words: bit i of a vector = (w[i>>5] >> (i&31)) & 1;
 *     syndromes use ceil(m/32) words, column vectors ceil(n/32) words per shot;
 *   - there is no CPU fallback: with no CUDA device every compute call fails with QB_ERR_CUDA.
 */
#ifndef QLDPC_B200_H
#define QLDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QB_OK = 0,
    QB_ERR_ARG = -1,        /* bad argument (mirrors the reference's ValueError cases) */
    QB_ERR_CUDA = -2,       /* CUDA runtime error / no device */
    QB_ERR_UNSUPPORTED = -3,/* problem shape outside what the kernels handle */
    QB_ERR_NOMEM = -4
} qb_status;

/* alpha schedules of the min-sum decoder (src/decoding/sparse.py:18-29) */
#define QB_ALPHA_FIXED 0     /* alpha_mode None / "alvarado": constant alpha */
#define QB_ALPHA_DYNAMIC 1   /* "dynamical": alpha_it = 1 - 2^-(it+1)  (kernels.py:272-273) */
#define QB_ALPHA_SEQUENCE 2  /* "alvarado-autoregressive": alpha_seq[min(it, len-1)] (kernels.py:402-405) */

/* arithmetic of the batched min-sum kernel.  QB_PRECISION_F32 is the default and the only mode the parity claims refer
 * to.  QB_PRECISION_HALF2 is an explicit opt-in: two shots share every 32-bit shared-memory slot as IEEE half floats
 * (about twice the min-sum throughput, NOT the reference's arithmetic; float32 posteriors for OSD). */
#define QB_PRECISION_F32 0
#define QB_PRECISION_HALF2 1

typedef struct qb_decoder qb_decoder;    /* one decoding side: Tanner graph, priors, logical rows */
typedef struct qb_sampler qb_sampler;    /* circuit fault tables: location -> column signatures */
typedef struct qb_pipeline qb_pipeline;  /* sampler + Z decoder + X decoder + batch workspaces */

typedef struct {
    int32_t max_iter;        /* maxIter of run_simulation (engine.py:194) */
    int32_t alpha_mode;      /* QB_ALPHA_* */
    float alpha_z, alpha_x;  /* fixed alpha per side (engine.py:84-88, 104-108) */
    const float *alpha_seq_z_h, *alpha_seq_x_h;   /* QB_ALPHA_SEQUENCE only */
    int32_t alpha_len_z, alpha_len_x;
    float clip_llr;          /* 20.0 in every engine call (sparse.py:13) */
    int32_t use_osd;         /* 1: OSD-0 on non-converged sides (engine.py:96-97); 0: keep BP output */
    int32_t precision;       /* QB_PRECISION_F32 (0, default) or QB_PRECISION_HALF2 (explicit opt-in) */
} qb_decode_config;

const char *qb_last_error(void);
int qb_device_count(void);
const char *qb_version(void);

/* ---- decoder side handle --------------------------------------------------------------- */
/* Graph of one side in CSR form (what sparse.py:33-34 extracts from the scipy matrix), priors
 * (engine.py:210-212) and the k logical rows in CSR form (engine.py:410-411).  k may be 0. */
int qb_decoder_create(int device, int32_t m, int32_t n, const int32_t *indptr_h, const int32_t *indices_h,
                      const double *prior_h, int32_t k, const int32_t *logical_ptr_h,
                      const int32_t *logical_idx_h, qb_decoder **out);
int qb_decoder_set_prior(qb_decoder *dec, const double *prior_h);
/* Min-sum arithmetic used by qb_minsum_batch / qb_minsum_decode_host on this handle (QB_PRECISION_*; the pipeline takes
 * it from qb_decode_config).  QB_PRECISION_HALF2 fails with QB_ERR_UNSUPPORTED at decode time when the graph has no
 * per-edge plan; it never silently falls back. */
int qb_decoder_set_precision(qb_decoder *dec, int32_t precision);
/* Which float32 min-sum kernel the handle uses for damping == 1 (a property of the graph and the priors): 0 = the
 * compressed-state kernel (csrc/minsum.cu), 1 = one float per edge in the shared memory of one SM
 * (csrc/minsum_edge.cu), N >= 2 = the same on a thread-block cluster of N SMs, check rows cut into N slabs
 * (csrc/minsum_edge_cluster.cu; the [[288,12,18]] decoding graphs).  All three follow src/decoding/kernels.py:235-366. */
int qb_decoder_minsum_path(qb_decoder *dec);

/* Host-only (no CUDA call): statistics of the shared-memory layout the per-edge min-sum kernel would use for this
 * graph (csrc/edge_layout.h): one float per Tanner-graph edge, check rows in conflict-free 128-bit order, the slot of
 * every edge chosen so that the 32 gathers of one warp instruction of the variable phase hit 32 different banks.
 * stats_out[16] = { usable, row slices, column slices, edge words, index words, gather instructions per iteration,
 *   shared-memory wavefronts they need (equal when conflict free), edges still in a bank conflict, priors uniform per
 *   slice, max chunks per row, max column degree, self check (every edge owns exactly one slot; the fingerprint tags
 *   in the free index halves of odd-degree slices are where and what the kernel expects), 0... }.
 * Replaces nothing in the reference (numba walks CSR arrays, src/decoding/kernels.py:283-345); it is the
 * data-layout step of qb_decoder_create, exposed for tests. */
int qb_edge_layout_probe(int32_t m, int32_t n, const int32_t *indptr_h, const int32_t *indices_h,
                         const double *prior_h, int32_t nwarps, int64_t *stats_out);
void qb_decoder_destroy(qb_decoder *dec);

/* ---- K3: batched flooding min-sum ------------------------------------------------------- */
/* Replaces minsum_decoder_full / minsum_decoder_full_autoregressive (src/decoding/kernels.py:235-366,
 * :370-485) for B independent syndromes.  Device-resident form: syn_bits_d [B][ceil(m/32)];
 * outputs hard_bits_d [B][ceil(n/32)], converged_d [B], final_iter_d [B] (the reference's
 * final_iter), post_d [B][n] float posteriors (nullable). */
int qb_minsum_batch(qb_decoder *dec, const uint32_t *syn_bits_d, int32_t B, int32_t max_iter,
                    int32_t alpha_mode, float alpha, const float *alpha_seq_h, int32_t alpha_len,
                    float damping, float clip_llr, uint32_t *hard_bits_d, uint8_t *converged_d,
                    int32_t *final_iter_d, float *post_d, void *stream);

/* Host-buffer form with the reference's array types: the call performMinSum_Symmetric_Sparse
 * (src/decoding/sparse.py:5-55) and performMinSum_Symmetric (src/decoding/dense.py:5-73) make.
 * syndrome_h int8 [B][m]; hard_h int8 [B][n]; values_h double [B][n] (nullable).
 * dense_variant != 0 selects dense.py's pre-damping rule (no clip before damping, dense.py:58-64). */
int qb_minsum_decode_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, int32_t max_iter,
                          int32_t alpha_mode, double alpha, const double *alpha_seq_h, int32_t alpha_len,
                          double damping, double clip_llr, int32_t dense_variant,
                          int8_t *hard_h, uint8_t *converged_h, int32_t *final_iter_h, double *values_h);

/* One check-node pass on flat CSR messages: minsum_core_sparse (src/decoding/kernels.py:139-169).
 * Q_h [B][nnz] -> R_h [B][nnz], Rsum_h [B][n]; syndrome_sign_h [B][m] (+1/-1). */
int qb_minsum_core_host(qb_decoder *dec, const double *Q_h, const double *syndrome_sign_h, int32_t B,
                        double alpha, double *R_h, double *Rsum_h);

/* Check-to-variable messages for the Alvarado alpha estimators (src/decoding/alpha.py:122-137, :206-253):
 * for each of B syndromes (int8 [B][m]) advance the min-sum decoder n_prev iterations with alpha_prev_h[0..n_prev)
 * (no convergence stop; damping and clip as in alpha.py:217-243), then return the unscaled (alpha = 1)
 * messages of the next check pass, R_h double [B][nnz] in CSR edge order.  prior_h double [n]. */
int qb_alpha_messages_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, const double *prior_h, int32_t n_prev,
                           const double *alpha_prev_h, double damping, double clip_llr, double *R_h);

/* tanh/atanh sum-product decoder: performBeliefPropagationFast + bp_core
 * (src/decoding/dense.py:75-96, src/decoding/kernels.py:172-193). */
int qb_bp_decode_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, int32_t max_iter,
                      int8_t *hard_h, uint8_t *converged_h, int32_t *final_iter_h, double *values_h);

/* Sparse syndrome of candidate vectors: syndrome_check (src/decoding/kernels.py:223-231).
 * candidate_h int8 [B][n] -> syndrome_h int8 [B][m]. */
int qb_syndrome_check_host(qb_decoder *dec, const int8_t *candidate_h, int32_t B, int8_t *syndrome_h);

/* ---- K4+K5: OSD-0 on BP failures ----------------------------------------------------------- */
/* performOSD_enhanced with order 0 (src/decoding/osd.py:5-29) for the F sides listed in
 * fail_idx_d (indices into the batch): residual syndrome, stable ascending |posterior| ordering
 * (ties by column index), GF(2) Gauss-Jordan in the reference's pivot order
 * (gf2_elimination_packed_core, src/decoding/kernels.py:49-96), solution = hard ^ e.
 * hard_bits_d is updated in place.  n_fail_d (device int32) gives F when F < 0. */
int qb_osd0_batch(qb_decoder *dec, const uint32_t *syn_bits_d, uint32_t *hard_bits_d, const float *post_d,
                  const int32_t *fail_idx_d, int32_t F, const int32_t *n_fail_d, int32_t max_fail, void *stream);

/* Host form, optionally with caller-supplied column orderings (ordering_h int32 [B][n], nullable)
 * so that elimination can be checked bit-exactly against the reference for the same np.argsort
 * result.  llr_h double [B][n] is used only when ordering_h is NULL.  solution_h int64 [B][n]
 * (osd.py returns int64); rank_h / pivots_h (nullable, [B] / [B][min(m,n)], -1 padded) report the pivots
 * actually needed (elimination stops as soon as the residual syndrome is fully reduced). */
int qb_osd0_host(qb_decoder *dec, const int8_t *syndrome_h, const int8_t *hard_h, const double *llr_h,
                 const int32_t *ordering_h, int32_t B, int64_t *solution_h, int32_t *rank_h, int32_t *pivots_h);

/* The pipeline's own OSD-0 path (the kernels qb_pipeline_run* launch for the non-converged sides of a batch; pivot rows
 * are free to differ from the reference's because simulated syndromes are always consistent) on
 * host-supplied float32 posteriors -- the entry the parity tests use to drive that path with crafted reliabilities
 * (mass ties, more candidates than one selection window).  syndrome_h int8 [B][m] must lie in the column space of H;
 * hard_h int8 [B][n]; post_h float32 [B][n]; solution_h int8 [B][n] = hard ^ e (osd.py:19-25);
 * osd_info_h (nullable) [B] = pivots_used | path << 16 as in qb_pipeline_last_batch_detail. */
int qb_osd0_pipeline_host(qb_decoder *dec, const int8_t *syndrome_h, const int8_t *hard_h, const float *post_h, int32_t B,
                          int8_t *solution_h, int32_t *osd_info_h);

/* Accounting of the pipeline's OSD-0 path on this decoder since the previous call (read and cleared):
 * out10_h[0..4] = sides that left the first tier of the free-row elimination kernel because of { candidate window
 * exhausted, touched-row capacity, free-slot capacity, record buffer, window not materialised }, out10_h[5..9] the
 * same for the second tier (those sides are solved by the full-width kernel).  Replaces nothing in the reference. */
int qb_decoder_osd_stats(qb_decoder *dec, int32_t *out10_h);

/* Dense GF(2) Gauss-Jordan of an arbitrary m x n 0/1 matrix with right-hand side:
 * gf2_elimination (src/decoding/kernels.py:6-34) and gf2_elimination_packed (:98-106).
 * A_h int64 [m][n] and b_h int64 [m] are reduced in place (the reference mutates them);
 * A_packed_h (nullable) receives the uint64-packed reduced matrix of kernels.py:36-46, [m][ceil(n/64)];
 * pivot_rows_h / pivot_cols_h int64 [min(m,n)]; *num_pivots_h the count. */
int qb_gf2_eliminate_host(int device, int64_t *A_h, int64_t *b_h, int32_t m, int32_t n, uint64_t *A_packed_h,
                          int64_t *pivot_rows_h, int64_t *pivot_cols_h, int32_t *num_pivots_h);

/* ---- K1+K2: fault sampling and syndrome extraction --------------------------------------------- */
/* Tables of one (code, circuit): for every fault location (gate index of cycle*num_cycles, the
 * rand_idx of generate_noisy_circuit_jit, src/noise/kernels.py:176-353) its kind (0: Z fault only
 * [MeasX/PrepX], 1: X fault only [MeasZ/PrepZ], 2: IDLE, 3: CNOT) and the decoding-matrix column of
 * each Z / X variant (variant = component on q1 + 2*component on q2; index 0 unused = -1), plus the
 * column signatures of both sides in CSC form (rows < m) with a logical-observable mask per column
 * (the rows >= first_logical_row of HZ_full / HX_full, src/noise/builder.py:115-124). */
int qb_sampler_create(int device, int32_t L, const int32_t *loc_kind_h, const int32_t *loc_colZ_h,
                      const int32_t *loc_colX_h,
                      int32_t mZ, int32_t nZ, const int32_t *colptrZ_h, const int32_t *rowsZ_h, const uint32_t *logmaskZ_h,
                      int32_t mX, int32_t nX, const int32_t *colptrX_h, const int32_t *rowsX_h, const uint32_t *logmaskX_h,
                      int32_t k, qb_sampler **out);
void qb_sampler_destroy(qb_sampler *s);

/* K2: syndromes and true logical flips of B shots from explicit fault events
 * (run_trial_fast minus the RNG draws, src/noise/simulation.py:48-105).  Shot b owns events
 * [ev_ptr[b], ev_ptr[b+1]); an event is  location | outcome << 24  with outcome = random_paulis
 * value (IDLE, 0..2), random_two_qubit value (CNOT, 0..14) or 0. */
int qb_syndrome_from_events(qb_sampler *s, const int32_t *ev_ptr_d, const uint32_t *events_d, int32_t B,
                            uint32_t *synZ_bits_d, uint32_t *trueZ_d, uint32_t *synX_bits_d, uint32_t *trueX_d,
                            void *stream);
int qb_syndrome_from_events_host(qb_sampler *s, const int32_t *ev_ptr_h, const uint32_t *events_h, int32_t B,
                                 int8_t *sparseZ_h, int8_t *trueZ_h, int8_t *sparseX_h, int8_t *trueX_h);

/* K1+K2 fused: Philox4x32-10 keyed by `seed`, counter = (global shot index, lane + 32 * call, 2).  Every location fails
 * independently with probability p and gets a uniform Pauli outcome (the noise model of src/noise/kernels.py:176-353; a
 * different generator than the reference's per-shot MT19937 reseed, engine.py:70, validated statistically).  The faults
 * of a shot are placed by sampling the geometric gaps between them: lane l of the shot's warp owns locations
 * [l*C, (l+1)*C), C = ceil(L/32), and consumes its Philox words in order -- one per jump (inversion against
 * T[k] = floor(2^32 (1-p)^k); a word below T[1023] extends the jump by another word), one per IDLE / CNOT fault for the
 * outcome floor(word * K / 2^32), K = 3 / 15.  Shots first_shot .. first_shot+B-1; results do not depend on B or on the
 * number of GPUs.  nfaults_d (nullable, [B]) receives the number of faults drawn per shot. */
int qb_sample_syndromes(qb_sampler *s, uint64_t seed, uint64_t first_shot, int32_t B, double p,
                        uint32_t *synZ_bits_d, uint32_t *trueZ_d, uint32_t *synX_bits_d, uint32_t *trueX_d,
                        int32_t *nfaults_d, void *stream);

/* The jump table qb_sample_syndromes uses for error rate p: table_h[k] = floor(2^32 (1-p)^k), k = 1..1023 (entry 0
 * unused); returns the number of words written (1024) or a negative status.  Host-only; lets a test re-derive a shot's
 * fault stream word for word. */
int qb_sampler_geometric_table(double p, uint32_t *table_h, int32_t capacity);

/* ---- fused per-shot pipeline ----------------------------------------------------------------- */
/* _run_single_trial_fast for a range of shots (src/simulation/engine.py:68-122): sample, syndromes,
 * min-sum both sides, OSD-0 on failures, logical comparison, counters. */
int qb_pipeline_create(qb_sampler *s, qb_decoder *decZ, qb_decoder *decX, int32_t max_batch, qb_pipeline **out);
void qb_pipeline_destroy(qb_pipeline *p);
/* Issue all pipeline work on the caller's stream (use_external != 0; e.g. torch's current stream, so
 * that the caller's CUDA events bracket it) or back on the pipeline's own stream (use_external == 0). */
int qb_pipeline_set_stream(qb_pipeline *p, void *stream, int use_external);

/* counts_h[8] = { z_errors, x_errors, total_errors, shots, z_nonconverged, x_nonconverged,
 *                 z_iterations, x_iterations } accumulated over the call (iterations = min-sum
 * iterations executed, the edge-message count is iterations * nnz).
 * flags_h (nullable, [n_shots]): bit0 = z_err, bit1 = x_err per shot in shot order, so the host can
 * reproduce the reference's in-order early stop (engine.py:450-464). */
int qb_pipeline_run(qb_pipeline *p, uint64_t seed, uint64_t first_shot, int64_t n_shots, double error_rate,
                    const qb_decode_config *cfg, int64_t *counts_h, uint8_t *flags_h);

/* Same pipeline fed with host fault events instead of the Philox sampler (parity testing against the
 * reference on identical host-sampled error batches; also the `e2e` bench path: host buffers in,
 * flags out).  B may exceed the pipeline's max_batch: the events go up in one copy and the batches are pipelined over
 * the two workspaces.  Optional per-shot outputs (nullable, B <= max_batch only): converged_h [2][B], final_iter_h [2][B]. */
int qb_pipeline_run_events_host(qb_pipeline *p, const int32_t *ev_ptr_h, const uint32_t *events_h, int32_t B,
                                const qb_decode_config *cfg, int64_t *counts_h, uint8_t *flags_h,
                                uint8_t *converged_h, int32_t *final_iter_h);

/* Decode-only pipeline on host syndromes (int8 [B][m] per side, true logical masks uint32 [B]):
 * the part of _run_single_trial_fast after run_trial_fast (engine.py:82-122). */
int qb_pipeline_decode_host(qb_pipeline *p, const int8_t *sparseZ_h, const uint32_t *trueZ_h,
                            const int8_t *sparseX_h, const uint32_t *trueX_h, int32_t B,
                            const qb_decode_config *cfg, int64_t *counts_h, uint8_t *flags_h);

/* Per-shot detail of the LAST batch a qb_pipeline_run* call decoded (parity testing of the pipeline's own OSD-0
 * kernels against the reference's elimination, src/decoding/osd.py:5-29 + kernels.py:49-96, on the very posteriors
 * the pipeline's min-sum produced).  side 0 = Z, 1 = X.  All outputs nullable:
 *   hard_bits_h [B][ceil(n/32)]  final corrections (after OSD-0 on the non-converged sides, engine.py:96-97);
 *   post_h      [B][n]           float32 min-sum posteriors; rows of converged sides are stale / unspecified
 *                                (the pipeline writes posteriors only for sides that go to OSD); the hard decision
 *                                min-sum handed to OSD is post < 0 (kernels.py:349);
 *   osd_info_h  [B]              0 for converged sides, else pivots_used | path << 16 with path 1 = free-row kernel,
 *                                2 = full-width kernel (recorded only after qb_pipeline_enable_detail(p, 1)). */
int qb_pipeline_enable_detail(qb_pipeline *p, int on);
int qb_pipeline_last_batch_detail(qb_pipeline *p, int32_t side, uint32_t *hard_bits_h, float *post_h, int32_t *osd_info_h);

/* timing / accounting of the last qb_pipeline_run* call (device time from CUDA events, ms) */
typedef struct {
    float ms_total, ms_sample, ms_minsum, ms_osd, ms_logical;
    int64_t kernel_launches;
    int64_t edge_messages;   /* sum over sides of nnz * iterations executed */
    int64_t osd_sides;       /* sides that went through OSD-0 */
} qb_pipeline_stats;
int qb_pipeline_last_stats(qb_pipeline *p, qb_pipeline_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_B200_H */
