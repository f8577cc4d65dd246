/*
 * qldpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU, double-precision restatement of the Monte-Carlo decoding hot path of
 * michelebanfi/qLDPC-branched-off (a pure Python + numba code base).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs have
 * something to check the CUDA path against and to time beside it.  Nothing in the product
 * package (qldpc-branched-off_b200/) may import, link or call it.
 *
 * Pinning: every function below is checked against the *real* reference (imported from
 * /root/reference with numba, see tests/golden/make_golden.py) and against the committed
 * golden vectors in tests/golden/ (tests/test_oracle_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to the reference
 * root).  Build: see oracle/build.py  (gcc -O2 -ffp-contract=off, strict IEEE: the reference's
 * inf / NaN conventions matter, see orc_minsum_decode).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* opcodes: src/noise/constants.py:8-31 */
enum { OP_CNOT = 1, OP_PREP_X = 2, OP_PREP_Z = 3, OP_MEAS_X = 4, OP_MEAS_Z = 5, OP_IDLE = 6,
       OP_X = 10, OP_Y = 11, OP_Z = 12,
       OP_XX = 20, OP_XY = 21, OP_XZ = 22, OP_YX = 23, OP_YY = 24, OP_YZ = 25,
       OP_ZX = 26, OP_ZY = 27, OP_ZZ = 28 };

/* ------------------------------------------------------------------------------------------
 * Min-sum decoder.  src/decoding/kernels.py:235-366 (minsum_decoder_full) and :370-485
 * (minsum_decoder_full_autoregressive; identical except for the alpha lookup :402-405).
 *   alpha_mode: 0 = fixed alpha_val, 1 = dynamical 1-2^-(it+1), 2 = alpha_seq[min(it,len-1)]
 * dense_variant = 1 restates the one arithmetic difference of the dense entry point
 * (src/decoding/dense.py:58-64): Q is not clipped before damping, only +-inf -> +-clip.
 * Returns converged flag; hard[n], values[n], *final_iter as the reference's return tuple.
 * ---------------------------------------------------------------------------------------- */
int orc_minsum_decode(const int32_t *H_indices, const int32_t *H_indptr, int m, int n,
                      const int8_t *syndrome, const double *prior, int maxIter,
                      int alpha_mode, double alpha_val, const double *alpha_seq, int alpha_len,
                      double damping, double clip_llr, int dense_variant,
                      int8_t *hard, double *values, int *final_iter)
{
    int nnz = H_indptr[m];
    double *Q = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    double *Qold = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    double *R = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    double *Rsum = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int e = 0; e < nnz; ++e) { Q[e] = prior[H_indices[e]]; Qold[e] = Q[e]; }   /* :263-265 */
    for (int j = 0; j < n; ++j) { hard[j] = 0; values[j] = 0.0; }
    int fin = maxIter - 1, conv = 0;                                               /* :267-268 */
    for (int it = 0; it < maxIter; ++it) {
        double alpha;
        if (alpha_mode == 1) alpha = 1.0 - pow(2.0, -(double)(it + 1));            /* :272-275 */
        else if (alpha_mode == 2) alpha = alpha_seq[it < alpha_len ? it : alpha_len - 1];
        else alpha = alpha_val;
        for (int j = 0; j < n; ++j) Rsum[j] = 0.0;
        for (int i = 0; i < m; ++i) {                                              /* :282-316 */
            int rs = H_indptr[i], re = H_indptr[i + 1];
            if (rs == re) continue;
            double sign_prod = 1.0 - 2.0 * (double)syndrome[i];
            double min1 = INFINITY, min2 = INFINITY;
            int min1_pos = -1;
            for (int p = rs; p < re; ++p) {
                double v = Q[p];
                if (!(v >= 0)) sign_prod = -sign_prod;
                double a = fabs(v);
                if (a < min1) { min2 = min1; min1 = a; min1_pos = p; }
                else if (a < min2) min2 = a;
            }
            for (int p = rs; p < re; ++p) {
                double v = Q[p];
                double sj = (v >= 0) ? 1.0 : -1.0;
                double mag = (p == min1_pos) ? min2 : min1;
                double msg = alpha * (sign_prod * sj) * mag;
                R[p] = msg;
                Rsum[H_indices[p]] += msg;
            }
        }
        for (int j = 0; j < n; ++j) values[j] = Rsum[j] + prior[j];                /* :319-320 */
        for (int e = 0; e < nnz; ++e) {                                            /* :323-345 */
            double q = values[H_indices[e]] - R[e];
            if (q != q) q = 0.0;
            else if (dense_variant) {          /* dense.py:62: nan_to_num maps only +-inf to +-clip */
                if (isinf(q)) q = q > 0 ? clip_llr : -clip_llr;
            }
            else if (q > clip_llr) q = clip_llr;
            else if (q < -clip_llr) q = -clip_llr;
            double qd = damping * q + (1.0 - damping) * Qold[e];
            if (qd > clip_llr) qd = clip_llr;
            else if (qd < -clip_llr) qd = -clip_llr;
            Q[e] = qd; Qold[e] = qd;
        }
        for (int j = 0; j < n; ++j) hard[j] = values[j] < 0 ? 1 : 0;               /* :348-349 */
        int ok = 1;
        for (int i = 0; i < m && ok; ++i) {                                        /* :352-359 */
            int s = 0;
            for (int p = H_indptr[i]; p < H_indptr[i + 1]; ++p) s ^= hard[H_indices[p]];
            if (s != syndrome[i]) ok = 0;
        }
        if (ok) { fin = it; conv = 1; break; }
    }
    *final_iter = fin;
    free(Q); free(Qold); free(R); free(Rsum);
    return conv;
}

/* One check-node pass on flat CSR messages.  src/decoding/kernels.py:139-169 */
void orc_minsum_core_sparse(const int32_t *H_indices, const int32_t *H_indptr, int m, int n,
                            const double *Q, const double *syndrome_sign, double alpha,
                            double *R, double *Rsum)
{
    for (int j = 0; j < n; ++j) Rsum[j] = 0.0;
    for (int e = 0; e < H_indptr[m]; ++e) R[e] = 0.0;
    for (int i = 0; i < m; ++i) {
        int rs = H_indptr[i], re = H_indptr[i + 1];
        if (rs == re) continue;
        double sp = syndrome_sign[i], min1 = INFINITY, min2 = INFINITY;
        int mp = -1;
        for (int p = rs; p < re; ++p) {
            double v = Q[p];
            sp *= (v >= 0) ? 1.0 : -1.0;
            double a = fabs(v);
            if (a < min1) { min2 = min1; min1 = a; mp = p; } else if (a < min2) min2 = a;
        }
        for (int p = rs; p < re; ++p) {
            double v = Q[p];
            double sj = (v >= 0) ? 1.0 : -1.0;
            double msg = alpha * (sp * sj) * ((p == mp) ? min2 : min1);
            R[p] = msg;
            Rsum[H_indices[p]] += msg;
        }
    }
}

/* tanh/atanh sum-product decoder on a CSR graph.  src/decoding/dense.py:75-96 with
 * bp_core src/decoding/kernels.py:172-193 (there on a dense mask; same edge order). */
int orc_bp_decode(const int32_t *H_indices, const int32_t *H_indptr, int m, int n,
                  const int8_t *syndrome, const double *prior, int maxIter,
                  int8_t *hard, double *values, int *final_iter)
{
    const double CLIP = 0.9999999;
    int nnz = H_indptr[m];
    double *Q = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    double *R = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    double *Rsum = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int e = 0; e < nnz; ++e) Q[e] = prior[H_indices[e]];
    for (int j = 0; j < n; ++j) { hard[j] = 0; values[j] = 0.0; }
    int conv = 0, it = 0;
    *final_iter = maxIter - 1;
    for (it = 0; it < maxIter; ++it) {
        for (int j = 0; j < n; ++j) Rsum[j] = 0.0;
        for (int i = 0; i < m; ++i) {
            double ss = 1.0 - 2.0 * (double)syndrome[i];
            double prod = 1.0;
            for (int p = H_indptr[i]; p < H_indptr[i + 1]; ++p) {
                double t = tanh(Q[p] * 0.5);
                if (fabs(t) < 1e-15) t = (t >= 0) ? 1e-15 : -1e-15;
                prod *= t;
            }
            for (int p = H_indptr[i]; p < H_indptr[i + 1]; ++p) {
                double t = tanh(Q[p] * 0.5);
                if (fabs(t) < 1e-15) t = (t >= 0) ? 1e-15 : -1e-15;
                double po = prod / t * ss;
                if (po > CLIP) po = CLIP; else if (po < -CLIP) po = -CLIP;
                R[p] = 2.0 * atanh(po);
                Rsum[H_indices[p]] += R[p];
            }
        }
        for (int j = 0; j < n; ++j) { values[j] = Rsum[j] + prior[j]; hard[j] = values[j] < 0; }
        for (int e = 0; e < nnz; ++e) Q[e] = values[H_indices[e]] - R[e];
        int ok = 1;
        for (int i = 0; i < m && ok; ++i) {
            int s = 0;
            for (int p = H_indptr[i]; p < H_indptr[i + 1]; ++p) s ^= hard[H_indices[p]];
            if (s != syndrome[i]) ok = 0;
        }
        if (ok) { conv = 1; *final_iter = it; break; }
    }
    free(Q); free(R); free(Rsum);
    return conv;
}

/* Byte-per-entry GF(2) Gauss-Jordan, in place.  src/decoding/kernels.py:6-34 */
int orc_gf2_elimination(int64_t *A, int64_t *b, int m, int n, int64_t *pivot_rows, int64_t *pivot_cols)
{
    int np_ = 0, row = 0;
    for (int col = 0; col < n; ++col) {
        if (row >= m) break;
        int pr = -1;
        for (int r = row; r < m; ++r) if (A[(size_t)r * n + col] == 1) { pr = r; break; }
        if (pr == -1) continue;
        if (pr != row) {
            for (int j = 0; j < n; ++j) { int64_t t = A[(size_t)row * n + j]; A[(size_t)row * n + j] = A[(size_t)pr * n + j]; A[(size_t)pr * n + j] = t; }
            int64_t t = b[row]; b[row] = b[pr]; b[pr] = t;
        }
        pivot_rows[np_] = row; pivot_cols[np_] = col; ++np_;
        for (int r = 0; r < m; ++r)
            if (r != row && A[(size_t)r * n + col] == 1) {
                for (int j = 0; j < n; ++j) A[(size_t)r * n + j] ^= A[(size_t)row * n + j];
                b[r] ^= b[row];
            }
        ++row;
    }
    return np_;
}

/* Bit-packed GF(2) Gauss-Jordan on uint64 rows (bit c of word c>>6 = column c), in place.
 * src/decoding/kernels.py:49-96 */
int orc_gf2_elimination_packed_core(uint64_t *A, int64_t *b, int m, int nwords, int n,
                                    int64_t *pivot_rows, int64_t *pivot_cols)
{
    int np_ = 0, row = 0;
    for (int col = 0; col < n; ++col) {
        if (row >= m) break;
        int w = col >> 6;
        uint64_t mask = (uint64_t)1 << (col & 63);
        int pr = -1;
        for (int r = row; r < m; ++r) if (A[(size_t)r * nwords + w] & mask) { pr = r; break; }
        if (pr == -1) continue;
        if (pr != row) {
            for (int k = 0; k < nwords; ++k) { uint64_t t = A[(size_t)row * nwords + k]; A[(size_t)row * nwords + k] = A[(size_t)pr * nwords + k]; A[(size_t)pr * nwords + k] = t; }
            int64_t t = b[row]; b[row] = b[pr]; b[pr] = t;
        }
        pivot_rows[np_] = row; pivot_cols[np_] = col; ++np_;
        for (int r = 0; r < m; ++r)
            if (r != row && (A[(size_t)r * nwords + w] & mask)) {
                for (int k = 0; k < nwords; ++k) A[(size_t)r * nwords + k] ^= A[(size_t)row * nwords + k];
                b[r] ^= b[row];
            }
        ++row;
    }
    return np_;
}

/* OSD-0 for a supplied column ordering.  src/decoding/osd.py:5-29:
 *   residual = syndrome ^ H.hard; H_permuted = H[:, ordering]; Gauss-Jordan (packed);
 *   e_permuted[pivot_col] = s_reduced[pivot_row]; solution = hard ^ unpermute(e).
 * H is given column-wise (CSC: col_ptr[n+1], row_idx[]).  pivots_out (nullable, length >= min(m,n))
 * receives the pivot positions in the permuted order; returns the rank. */
int orc_osd0(const int32_t *col_ptr, const int32_t *row_idx, int m, int n,
             const int8_t *syndrome, const int8_t *hard, const int64_t *ordering,
             int64_t *solution, int64_t *pivots_out)
{
    int nwords = (n + 63) / 64;
    uint64_t *A = (uint64_t *)calloc((size_t)m * nwords + 1, sizeof(uint64_t));
    int64_t *b = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
    int mn = m < n ? m : n;
    int64_t *pr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(mn + 1));
    int64_t *pc = (int64_t *)malloc(sizeof(int64_t) * (size_t)(mn + 1));
    for (int i = 0; i < m; ++i) b[i] = syndrome[i];
    for (int j = 0; j < n; ++j)
        if (hard[j]) for (int p = col_ptr[j]; p < col_ptr[j + 1]; ++p) b[row_idx[p]] ^= 1;
    for (int p = 0; p < n; ++p) {
        int j = (int)ordering[p];
        for (int q = col_ptr[j]; q < col_ptr[j + 1]; ++q)
            A[(size_t)row_idx[q] * nwords + (p >> 6)] |= (uint64_t)1 << (p & 63);
    }
    int r = orc_gf2_elimination_packed_core(A, b, m, nwords, n, pr, pc);
    for (int j = 0; j < n; ++j) solution[j] = hard[j];
    for (int t = 0; t < r; ++t) {
        if (b[pr[t]]) solution[ordering[pc[t]]] ^= 1;
        if (pivots_out) pivots_out[t] = pc[t];
    }
    free(A); free(b); free(pr); free(pc);
    return r;
}

/* Noisy-circuit generation.  src/noise/kernels.py:176-353.  Returns the output length. */
int orc_generate_noisy_circuit(const int32_t *ops, const int32_t *q1, const int32_t *q2, int ngates,
                               double p, const double *rv, const int32_t *rp, const int32_t *r2,
                               int32_t *oo, int32_t *o1, int32_t *o2)
{
    static const int two_op[15] = { OP_X, OP_Y, OP_Z, OP_X, OP_Y, OP_Z, OP_XX, OP_YY, OP_ZZ,
                                    OP_XY, OP_YX, OP_YZ, OP_ZY, OP_XZ, OP_ZX };
    int o = 0, ri = 0;
#define EMIT(op_, a_, b_) do { oo[o] = (op_); o1[o] = (a_); o2[o] = (b_); ++o; } while (0)
    for (int i = 0; i < ngates; ++i) {
        int op = ops[i], a = q1[i], b = q2[i];
        if (op == OP_MEAS_X) { if (rv[ri] < p) EMIT(OP_Z, a, -1); ++ri; EMIT(op, a, b); }
        else if (op == OP_MEAS_Z) { if (rv[ri] < p) EMIT(OP_X, a, -1); ++ri; EMIT(op, a, b); }
        else if (op == OP_PREP_X) { EMIT(op, a, b); if (rv[ri] < p) EMIT(OP_Z, a, -1); ++ri; }
        else if (op == OP_PREP_Z) { EMIT(op, a, b); if (rv[ri] < p) EMIT(OP_X, a, -1); ++ri; }
        else if (op == OP_IDLE) {
            if (rv[ri] < p) { int c = rp[ri]; EMIT(c == 0 ? OP_X : (c == 1 ? OP_Y : OP_Z), a, -1); }
            ++ri;
        } else if (op == OP_CNOT) {
            EMIT(op, a, b);
            if (rv[ri] < p) {
                int t = r2[ri];
                if (t < 0 || t > 14) t = 14;
                if (t < 3) EMIT(two_op[t], a, -1);
                else if (t < 6) EMIT(two_op[t], b, -1);
                else EMIT(two_op[t], a, b);
            }
            ++ri;
        } else EMIT(op, a, b);
    }
#undef EMIT
    return o;
}

/* Pauli-frame propagation.  side 0: Z errors / X checks (src/noise/kernels.py:14-91),
 * side 1: X errors / Z checks (:95-172).  state has total_qubits entries; returns syn_count. */
int orc_simulate_circuit(int side, const int32_t *ops, const int32_t *q1, const int32_t *q2, int ngates,
                         int total_qubits, int8_t *history, int8_t *state)
{
    memset(state, 0, (size_t)total_qubits);
    int syn = 0;
    for (int i = 0; i < ngates; ++i) {
        int op = ops[i], a = q1[i], b = q2[i];
        if (side == 0) {
            if (op == OP_CNOT) state[a] ^= state[b];
            else if (op == OP_PREP_X) state[a] = 0;
            else if (op == OP_MEAS_X) history[syn++] = state[a];
            else if (op == OP_Z || op == OP_Y) state[a] ^= 1;
            else if (op == OP_ZX || op == OP_YX) state[a] ^= 1;
            else if (op == OP_XZ || op == OP_XY) state[b] ^= 1;
            else if (op == OP_ZZ || op == OP_YY || op == OP_YZ || op == OP_ZY) { state[a] ^= 1; state[b] ^= 1; }
        } else {
            if (op == OP_CNOT) state[b] ^= state[a];
            else if (op == OP_PREP_Z) state[a] = 0;
            else if (op == OP_MEAS_Z) history[syn++] = state[a];
            else if (op == OP_X || op == OP_Y) state[a] ^= 1;
            else if (op == OP_XZ || op == OP_YZ) state[a] ^= 1;
            else if (op == OP_ZX || op == OP_ZY) state[b] ^= 1;
            else if (op == OP_XX || op == OP_YY || op == OP_XY || op == OP_YX) { state[a] ^= 1; state[b] ^= 1; }
        }
    }
    return syn;
}

/* Detector differencing.  src/noise/kernels.py:357-380 */
void orc_sparsify(const int8_t *history, int syn_count, const int32_t *pos, const int32_t *ptr,
                  int num_checks, int8_t *out)
{
    memcpy(out, history, (size_t)syn_count);
    for (int c = 0; c < num_checks; ++c)
        for (int i = ptr[c] + 1; i < ptr[c + 1]; ++i)
            if (pos[i] < syn_count && pos[i - 1] < syn_count) out[pos[i]] ^= history[pos[i - 1]];
}

/* run_trial_fast without the RNG draws (the caller passes the three random arrays).
 * src/noise/simulation.py:21-107.  L: k x n_data 0/1 bytes, row-major. */
void orc_run_trial(const int32_t *bops, const int32_t *bq1, const int32_t *bq2, int nbase,
                   const int32_t *sops, const int32_t *sq1, const int32_t *sq2, int nsuf,
                   int total_qubits, double p, const double *rv, const int32_t *rp, const int32_t *r2,
                   const int32_t *xpos, const int32_t *xptr, int nx,
                   const int32_t *zpos, const int32_t *zptr, int nz,
                   const int32_t *data_idx, int ndata, const uint8_t *Lx, const uint8_t *Lz, int k,
                   int8_t *sparse_z, int8_t *true_z, int8_t *sparse_x, int8_t *true_x)
{
    int cap = 2 * nbase + nsuf + 8;
    int32_t *oo = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)cap);
    int32_t *o1 = oo + cap, *o2 = oo + 2 * cap;
    int len = orc_generate_noisy_circuit(bops, bq1, bq2, nbase, p, rv, rp, r2, oo, o1, o2);
    memcpy(oo + len, sops, sizeof(int32_t) * (size_t)nsuf);
    memcpy(o1 + len, sq1, sizeof(int32_t) * (size_t)nsuf);
    memcpy(o2 + len, sq2, sizeof(int32_t) * (size_t)nsuf);
    len += nsuf;
    int8_t *hist = (int8_t *)malloc((size_t)len + 8);
    int8_t *state = (int8_t *)malloc((size_t)total_qubits + 8);
    for (int side = 0; side < 2; ++side) {
        int syn = orc_simulate_circuit(side, oo, o1, o2, len, total_qubits, hist, state);
        const uint8_t *L = side == 0 ? Lx : Lz;
        int8_t *tl = side == 0 ? true_z : true_x;
        for (int b = 0; b < k; ++b) {
            int acc = 0;
            for (int j = 0; j < ndata; ++j) acc += L[(size_t)b * ndata + j] * state[data_idx[j]];
            tl[b] = (int8_t)(acc % 2);
        }
        if (side == 0) orc_sparsify(hist, syn, xpos, xptr, nx, sparse_z);
        else orc_sparsify(hist, syn, zpos, zptr, nz, sparse_x);
    }
    free(oo); free(hist); free(state);
}

/* stable argsort of |llr| ascending (ties by index): the documented tie rule of the GPU path;
 * np.argsort in osd.py:12 is unstable, so tests either supply the ordering or use this rule. */
typedef struct { double key; int64_t idx; } orc_kv;
static int orc_kv_cmp(const void *a, const void *b)
{
    const orc_kv *x = (const orc_kv *)a, *y = (const orc_kv *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
void orc_stable_abs_argsort(const double *llr, int n, int64_t *ordering)
{
    orc_kv *kv = (orc_kv *)malloc(sizeof(orc_kv) * (size_t)(n + 1));
    for (int j = 0; j < n; ++j) { kv[j].key = fabs(llr[j]); kv[j].idx = j; }
    qsort(kv, (size_t)n, sizeof(orc_kv), orc_kv_cmp);
    for (int j = 0; j < n; ++j) ordering[j] = kv[j].idx;
    free(kv);
}

/* One decoding side of _run_single_trial_fast: BP, OSD-0 on failure, logical comparison.
 * src/simulation/engine.py:83-100.  logical rows given as CSR over columns (lptr[k+1], lidx).
 * Returns 1 when the decoded logical differs from true_l.  stats[0]=converged, [1]=iterations. */
int orc_decode_side(const int32_t *H_indices, const int32_t *H_indptr,
                    const int32_t *col_ptr, const int32_t *row_idx, int m, int n,
                    const int32_t *lptr, const int32_t *lidx, int k,
                    const int8_t *syndrome, const int8_t *true_l, const double *prior,
                    int maxIter, int alpha_mode, double alpha_val, const double *alpha_seq, int alpha_len,
                    int32_t *stats)
{
    int8_t *hard = (int8_t *)malloc((size_t)n + 8);
    double *values = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    int64_t *sol = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int fin = 0;
    int conv = orc_minsum_decode(H_indices, H_indptr, m, n, syndrome, prior, maxIter, alpha_mode,
                                 alpha_val, alpha_seq, alpha_len, 1.0, 20.0, 0, hard, values, &fin);
    if (!conv) {
        int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
        orc_stable_abs_argsort(values, n, ord);
        orc_osd0(col_ptr, row_idx, m, n, syndrome, hard, ord, sol, NULL);
        free(ord);
    } else {
        for (int j = 0; j < n; ++j) sol[j] = hard[j];
    }
    int err = 0;
    for (int b = 0; b < k; ++b) {
        int acc = 0;
        for (int p = lptr[b]; p < lptr[b + 1]; ++p) acc ^= (int)(sol[lidx[p]] & 1);
        if (acc != true_l[b]) err = 1;
    }
    if (stats) { stats[0] = conv; stats[1] = fin + 1; }
    free(hard); free(values); free(sol);
    return err;
}
