// C ABI of libqldpc_b200.so: handles, host staging and the fused per-shot pipeline.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "edge_dev.cuh"

namespace qb {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
    return QB_ERR_CUDA;
}

int Scratch::ensure(size_t bytes)
{
    if (bytes <= cap) return QB_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    QB_CUDA(cudaMalloc(&ptr, want));
    cap = want;
    return QB_OK;
}
void Scratch::release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }

// kernels / launchers defined in the other translation units
int launch_minsum_core(qb_decoder *dec, const double *Q, const double *ssign, int B, double alpha, double *R,
                       double *Rsum, cudaStream_t st);
int launch_bp(qb_decoder *dec, const uint32_t *syn_bits, int B, int max_iter, uint32_t *hard_bits,
              uint8_t *converged, int32_t *final_iter, double *post, cudaStream_t st);
int launch_syndrome_check(qb_decoder *dec, const uint32_t *cand_bits, int B, uint32_t *syn_bits, cudaStream_t st);
int launch_gf2_dense(uint32_t *A, uint32_t *b, int m, int n, int nw, int32_t *pr, int32_t *pc, int32_t *np, cudaStream_t st);
int fast_shots_per_cta(const qb_decoder *dec);
int launch_alpha_messages(qb_decoder *dec, const int8_t *syn, int B, int n_prev, const double *alpha_prev_d, double damping,
                          double clip, const double *prior64_d, double *R_out, cudaStream_t st);

template <class T>
static int to_device(std::vector<void *> &owned, const std::vector<T> &h, T **out)
{
    *out = nullptr;
    const size_t bytes = sizeof(T) * std::max<size_t>(1, h.size());
    QB_CUDA(cudaMalloc(reinterpret_cast<void **>(out), bytes));
    owned.push_back(*out);
    if (!h.empty()) QB_CUDA(cudaMemcpy(*out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return QB_OK;
}

// carve aligned sub-buffers out of one allocation
struct Carver {
    unsigned char *base; size_t off = 0;
    explicit Carver(void *p) : base(static_cast<unsigned char *>(p)) {}
    template <class T> T *take(size_t count) { off = (off + 255) & ~(size_t)255; T *r = reinterpret_cast<T *>(base + off); off += sizeof(T) * count; return r; }
};
static size_t carve_size(std::initializer_list<size_t> bytes) { size_t t = 0; for (size_t b : bytes) t = ((t + 255) & ~(size_t)255) + b; return t + 256; }

__global__ void f32_to_f64_kernel(const float *src, double *dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}
__global__ void hard_to_i64_kernel(const uint32_t *bits, int B, int n, int nw, int64_t *dst)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * n) return;
    const int b = (int)(i / n), j = (int)(i % n);
    dst[i] = (bits[(size_t)b * nw + (j >> 5)] >> (j & 31)) & 1u;
}
__global__ void fill_bytes_kernel(uint8_t *p, size_t n, uint8_t v) { const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
// counters of one batch: total errors, non-converged sides, iterations
__global__ void batch_counts_kernel(const uint8_t *flags, const uint8_t *convZ, const uint8_t *convX,
                                    const int32_t *itZ, const int32_t *itX, int B, unsigned long long *counts)
{
    unsigned long long tot = 0, ncz = 0, ncx = 0, iz = 0, ix = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        tot += flags[i] != 0; ncz += convZ[i] == 0; ncx += convX[i] == 0; iz += itZ[i] + 1; ix += itX[i] + 1;
    }
    for (int o = 16; o; o >>= 1) {
        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o); ncz += __shfl_xor_sync(0xFFFFFFFFu, ncz, o);
        ncx += __shfl_xor_sync(0xFFFFFFFFu, ncx, o); iz += __shfl_xor_sync(0xFFFFFFFFu, iz, o); ix += __shfl_xor_sync(0xFFFFFFFFu, ix, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (tot) atomicAdd(&counts[2], tot);
        if (ncz) atomicAdd(&counts[4], ncz);
        if (ncx) atomicAdd(&counts[5], ncx);
        atomicAdd(&counts[6], iz); atomicAdd(&counts[7], ix);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[3], (unsigned long long)B);
    }
}

}  // namespace qb

using namespace qb;

extern "C" {

const char *qb_last_error(void) { return g_err.c_str(); }
const char *qb_version(void) { return "qldpc_b200 0.1.0 (sm_100a)"; }
int qb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int qb_decoder_create(int device, int32_t m, int32_t n, const int32_t *indptr, const int32_t *indices,
                      const double *prior, int32_t k, const int32_t *lptr, const int32_t *lidx, qb_decoder **out)
{
    QB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    QB_REQUIRE(m >= 0 && n >= 0 && indptr && prior, "bad graph arguments");
    QB_REQUIRE(k >= 0 && k <= 32, "at most 32 logical rows are supported");
    QB_REQUIRE(k == 0 || (lptr && lidx), "logical rows are NULL");
    QB_REQUIRE(k == 0 || lptr[0] == 0, "logical_ptr[0] must be 0");
    for (int b = 0; b < k; ++b) QB_REQUIRE(lptr[b + 1] >= lptr[b], "logical_ptr must be non-decreasing");
    QB_REQUIRE(indptr[0] == 0, "indptr[0] must be 0");
    for (int i = 0; i < m; ++i) QB_REQUIRE(indptr[i + 1] >= indptr[i], "indptr must be non-decreasing");
    const int nnz = indptr[m];
    QB_REQUIRE(nnz == 0 || indices, "indices is NULL");
    for (int e = 0; e < nnz; ++e) QB_REQUIRE(indices[e] >= 0 && indices[e] < n, "column index out of range");
    QB_CUDA(cudaSetDevice(device));
    qb_decoder *d = new qb_decoder();
    d->device = device;
    cudaDeviceProp prop;
    QB_CUDA(cudaGetDeviceProperties(&prop, device));
    d->sm_count = prop.multiProcessorCount;
    d->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    GraphDev &g = d->g;
    g.m = m; g.n = n; g.nnz = nnz; g.k = k;
    g.mw = std::max(1, ceil_div(m, 32)); g.nw = std::max(1, ceil_div(n, 32));
    g.m_pad = g.mw * 32; g.n_pad = g.nw * 32;
    d->h_indptr.assign(indptr, indptr + m + 1);
    d->h_indices.assign(indices, indices + nnz);
    // CSC with rows ascending + position of every edge inside its row
    std::vector<int32_t> colptr(n + 1, 0), rowidx(nnz), csc_edge(nnz), pos_in_row(nnz);
    for (int e = 0; e < nnz; ++e) colptr[indices[e] + 1]++;
    for (int j = 0; j < n; ++j) colptr[j + 1] += colptr[j];
    {
        std::vector<int32_t> fill(colptr.begin(), colptr.end() - 1);
        for (int r = 0; r < m; ++r)
            for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
                const int p = fill[indices[e]]++;
                rowidx[p] = r; csc_edge[p] = e; pos_in_row[e] = e - indptr[r];
            }
    }
    d->h_colptr = colptr; d->h_rowidx = rowidx;
    int max_rd = 0, max_cd = 0;
    for (int r = 0; r < m; ++r) max_rd = std::max(max_rd, indptr[r + 1] - indptr[r]);
    for (int j = 0; j < n; ++j) max_cd = std::max(max_cd, colptr[j + 1] - colptr[j]);
    d->max_row_deg = max_rd; d->max_col_deg = max_cd;
    d->fast_ok = (n <= 65000) && (max_rd <= MS_MAX_ROW_DEG) && (max_cd <= 255) && (m < (1 << 23));
    // sliced, chunked ELL (see GraphDev)
    g.n_rslices = g.mw; g.n_cslices = g.nw;
    std::vector<int32_t> rsp(g.n_rslices + 1, 0), csp(g.n_cslices + 1, 0);
    for (int s = 0; s < g.n_rslices; ++s) {
        int deg = 0;
        for (int r = s * 32; r < std::min(m, s * 32 + 32); ++r) deg = std::max(deg, indptr[r + 1] - indptr[r]);
        rsp[s + 1] = rsp[s] + ceil_div(deg, 8) * 32;
    }
    for (int s = 0; s < g.n_cslices; ++s) {
        int deg = 0;
        for (int j = s * 32; j < std::min(n, s * 32 + 32); ++j) deg = std::max(deg, colptr[j + 1] - colptr[j]);
        csp[s + 1] = csp[s] + ceil_div(deg, 4) * 32;
    }
    std::vector<uint8_t> rexact(std::max(1, g.n_rslices), 0);
    for (int r = 0; r < m; ++r) if (indptr[r + 1] - indptr[r] == 1) rexact[r >> 5] = 1;
    g.nan_anywhere = 0;
    for (int j = 0; j < n; ++j) if (!std::isfinite((float)prior[j])) g.nan_anywhere = 1;   // the kernels see float priors
    {   // a variable on two degree-1 rows can receive +inf and -inf -> NaN posterior: exact path everywhere
        std::vector<int> deg1(n, 0);
        for (int r = 0; r < m; ++r) if (indptr[r + 1] - indptr[r] == 1 && ++deg1[indices[indptr[r]]] > 1) d->graph_nan = 1;
        g.nan_anywhere |= d->graph_nan;
    }
    std::vector<uint16_t> row_ell((size_t)rsp.back() * 8 + 8, (uint16_t)g.n_pad);               // dummy +inf variable
    std::vector<uint32_t> col_ell((size_t)csp.back() * 4 + 4, (uint32_t)g.m_pad << 8);          // dummy zero check
    std::vector<uint8_t> rdeg(std::max(1, g.n_rslices), 0), cdeg(std::max(1, g.n_cslices), 0);
    for (int r = 0; r < m; ++r) rdeg[r >> 5] = (uint8_t)std::max<int>(rdeg[r >> 5], std::min(255, indptr[r + 1] - indptr[r]));
    for (int j = 0; j < n; ++j) cdeg[j >> 5] = (uint8_t)std::max<int>(cdeg[j >> 5], std::min(255, colptr[j + 1] - colptr[j]));
    if (d->fast_ok) {
        for (int r = 0; r < m; ++r)
            for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
                const int t = e - indptr[r];
                row_ell[((size_t)rsp[r >> 5] + (t >> 3) * 32 + (r & 31)) * 8 + (t & 7)] = (uint16_t)indices[e];
            }
        for (int j = 0; j < n; ++j)
            for (int p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int t = p - colptr[j];
                col_ell[((size_t)csp[j >> 5] + (t >> 2) * 32 + (j & 31)) * 4 + (t & 3)] =
                    ((uint32_t)rowidx[p] << 8) | (uint32_t)pos_in_row[csc_edge[p]];
            }
    }
    std::vector<uint16_t> colsig;
    if (max_cd <= 8 && m <= 65534) {
        colsig.assign((size_t)std::max(1, n) * 8, 0xFFFFu);
        for (int j = 0; j < n; ++j)
            for (int p = colptr[j]; p < colptr[j + 1]; ++p) colsig[(size_t)j * 8 + (p - colptr[j])] = (uint16_t)rowidx[p];
    }
    std::vector<float> pf(n);
    for (int j = 0; j < n; ++j) pf[j] = (float)prior[j];
    d->h_prior = pf;
    std::vector<uint32_t> logmask(n, 0u);
    for (int b = 0; b < k; ++b)
        for (int p = lptr[b]; p < lptr[b + 1]; ++p) {
            if (lidx[p] < 0 || lidx[p] >= n) { delete d; set_error("logical column index out of range"); return QB_ERR_ARG; }
            logmask[lidx[p]] ^= 1u << b;
        }
    int rc = QB_OK;
    int32_t *p32; uint16_t *p16; uint32_t *pu32; float *pfl;
#define UP(vec, ptr, field) if (!rc) { rc = to_device(d->owned, vec, &ptr); field = ptr; }
    UP(rsp, p32, g.rslice_ptr) UP(csp, p32, g.cslice_ptr)
    if (!rc) { rc = to_device(d->owned, row_ell, &p16); g.row_ell4 = reinterpret_cast<const uint4 *>(p16); }
    if (!rc) { rc = to_device(d->owned, col_ell, &pu32); g.col_ell4 = reinterpret_cast<const uint4 *>(pu32); }
    UP(d->h_indptr, p32, g.indptr) UP(d->h_indices, p32, g.indices) UP(colptr, p32, g.colptr) UP(rowidx, p32, g.rowidx)
    UP(csc_edge, p32, g.csc_edge) UP(logmask, pu32, g.logmask)
    if (!rc) { rc = to_device(d->owned, pf, &pfl); g.prior = pfl; d->d_prior = pfl; }
    g.colsig = nullptr;
    { uint8_t *p8 = nullptr; if (!rc) { rc = to_device(d->owned, rexact, &p8); g.rslice_exact = p8; }
      if (!rc) { rc = to_device(d->owned, rdeg, &p8); g.rslice_deg = p8; }
      if (!rc) { rc = to_device(d->owned, cdeg, &p8); g.cslice_deg = p8; } }
    if (!rc && !colsig.empty()) { rc = to_device(d->owned, colsig, &p16); g.colsig = reinterpret_cast<const uint4 *>(p16); }
#undef UP
    if (!rc && !g.nan_anywhere) rc = edge_plan_create(d, pf.data(), &d->edge);
    if (!rc && !g.nan_anywhere && !d->edge) rc = cluster_plan_create(d, pf.data(), &d->cluster);
    if (rc) { qb_decoder_destroy(d); return rc; }
    *out = d;
    return QB_OK;
}

int qb_decoder_set_prior(qb_decoder *dec, const double *prior)
{
    QB_REQUIRE(dec && prior, "NULL argument");
    QB_CUDA(cudaSetDevice(dec->device));
    std::vector<float> pf(dec->g.n);
    dec->g.nan_anywhere = dec->graph_nan;
    for (int j = 0; j < dec->g.n; ++j) { pf[j] = (float)prior[j]; if (!std::isfinite(pf[j])) dec->g.nan_anywhere = 1; }
    if (dec->g.n) QB_CUDA(cudaMemcpy(dec->d_prior, pf.data(), sizeof(float) * pf.size(), cudaMemcpyHostToDevice));
    // the per-edge layout groups variables by prior value: rebuild it
    QB_CUDA(cudaDeviceSynchronize());
    dec->h_prior = pf;
    edge_plan_h2_destroy(dec->edge_h2);
    dec->edge_h2 = nullptr;
    edge_plan_destroy(dec->edge);
    dec->edge = nullptr;
    cluster_plan_destroy(dec->cluster);
    dec->cluster = nullptr;
    // same predicate as qb_decoder_create: non-finite priors or a graph that can produce NaN posteriors on its own
    // (a variable on two degree-1 rows) stay on the exact compressed-state kernel
    if (!dec->g.nan_anywhere) {
        if (int rc = edge_plan_create(dec, pf.data(), &dec->edge)) return rc;
        if (!dec->edge) return cluster_plan_create(dec, pf.data(), &dec->cluster);
    }
    return QB_OK;
}

int qb_decoder_set_precision(qb_decoder *dec, int32_t precision)
{
    QB_REQUIRE(dec != nullptr, "NULL argument");
    QB_REQUIRE(precision == QB_PRECISION_F32 || precision == QB_PRECISION_HALF2, "unknown precision");
    dec->precision = precision;
    return QB_OK;
}

int qb_decoder_minsum_path(qb_decoder *dec)
{
    if (!dec) return 0;
    if (dec->edge) return 1;
    return cluster_plan_size(dec->cluster);
}

void qb_decoder_destroy(qb_decoder *dec)
{
    if (!dec) return;
    cudaSetDevice(dec->device);
    for (void *p : dec->owned) cudaFree(p);
    edge_plan_h2_destroy(dec->edge_h2);
    edge_plan_destroy(dec->edge);
    cluster_plan_destroy(dec->cluster);
    if (dec->d_alpha) cudaFree(dec->d_alpha);
    dec->scratch.release(); dec->work.release(); dec->ovf.release();
    delete dec;
}

static int check_alpha(int alpha_mode, double alpha, const void *seq, int len)
{
    QB_REQUIRE(alpha_mode >= 0 && alpha_mode <= 2, "Unsupported alpha_mode");
    if (alpha_mode == QB_ALPHA_SEQUENCE) QB_REQUIRE(seq != nullptr && len > 0, "alpha must be a non-empty 1D sequence for alvarado-autoregressive");
    (void)alpha;
    return QB_OK;
}

int qb_minsum_batch(qb_decoder *dec, const uint32_t *syn_bits_d, int32_t B, int32_t max_iter, int32_t alpha_mode,
                    float alpha, const float *alpha_seq_h, int32_t alpha_len, float damping, float clip_llr,
                    uint32_t *hard_bits_d, uint8_t *converged_d, int32_t *final_iter_d, float *post_d, void *stream)
{
    QB_REQUIRE(dec && syn_bits_d && hard_bits_d && converged_d && final_iter_d, "NULL argument");
    if (int rc = check_alpha(alpha_mode, alpha, alpha_seq_h, alpha_len)) return rc;
    QB_CUDA(cudaSetDevice(dec->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<double> seq;
    if (alpha_mode == QB_ALPHA_SEQUENCE) seq.assign(alpha_seq_h, alpha_seq_h + alpha_len);
    if (int rc = upload_alpha(dec, max_iter, alpha_mode, alpha, seq.data(), alpha_len, st)) return rc;
    MinsumLaunch a{};
    a.syn_bits = syn_bits_d; a.B = B; a.max_iter = max_iter; a.alpha_d = dec->d_alpha;
    a.damping = damping; a.clip = clip_llr; a.dense_variant = 0;
    a.hard_bits = hard_bits_d; a.converged = converged_d; a.final_iter = final_iter_d; a.post = post_d;
    a.precision = dec->precision;
    return launch_minsum(dec, a, st);
}

int qb_minsum_decode_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, int32_t max_iter, int32_t alpha_mode,
                          double alpha, const double *alpha_seq_h, int32_t alpha_len, double damping, double clip_llr,
                          int32_t dense_variant, int8_t *hard_h, uint8_t *converged_h, int32_t *final_iter_h,
                          double *values_h)
{
    QB_REQUIRE(dec && syndrome_h && hard_h && converged_h && final_iter_h, "NULL argument");
    QB_REQUIRE(B >= 0, "negative batch");
    if (int rc = check_alpha(alpha_mode, alpha, alpha_seq_h, alpha_len)) return rc;
    if (B == 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    const size_t total = carve_size({sB * g.m, sB * g.mw * 4, sB * g.nw * 4, sB, sB * 4, sB * g.n * 4, sB * g.n * 8, sB * g.n});
    if (int rc = dec->scratch.ensure(total)) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_syn8 = cv.take<int8_t>(sB * g.m);
    uint32_t *d_syn = cv.take<uint32_t>(sB * g.mw);
    uint32_t *d_hard = cv.take<uint32_t>(sB * g.nw);
    uint8_t *d_conv = cv.take<uint8_t>(sB);
    int32_t *d_fin = cv.take<int32_t>(sB);
    float *d_post = cv.take<float>(sB * g.n);
    double *d_post64 = cv.take<double>(sB * g.n);
    int8_t *d_hard8 = cv.take<int8_t>(sB * g.n);
    cudaStream_t st = 0;
    if (g.m) QB_CUDA(cudaMemcpyAsync(d_syn8, syndrome_h, sB * g.m, cudaMemcpyHostToDevice, st));
    if (int rc = launch_pack_bits(d_syn8, B, g.m, d_syn, g.mw, st)) return rc;
    if (int rc = upload_alpha(dec, max_iter, alpha_mode, alpha, alpha_seq_h, alpha_len, st)) return rc;
    MinsumLaunch a{};
    a.syn_bits = d_syn; a.B = B; a.max_iter = max_iter; a.alpha_d = dec->d_alpha;
    a.damping = (float)damping; a.clip = (float)clip_llr; a.dense_variant = dense_variant;
    a.hard_bits = d_hard; a.converged = d_conv; a.final_iter = d_fin; a.post = d_post;
    a.precision = dense_variant ? QB_PRECISION_F32 : dec->precision;
    if (int rc = launch_minsum(dec, a, st)) return rc;
    if (int rc = launch_unpack_bits(d_hard, B, g.n, g.nw, d_hard8, st)) return rc;
    if (g.n) QB_CUDA(cudaMemcpyAsync(hard_h, d_hard8, sB * g.n, cudaMemcpyDeviceToHost, st));
    QB_CUDA(cudaMemcpyAsync(converged_h, d_conv, sB, cudaMemcpyDeviceToHost, st));
    QB_CUDA(cudaMemcpyAsync(final_iter_h, d_fin, sB * 4, cudaMemcpyDeviceToHost, st));
    if (values_h && g.n) {
        const size_t cnt = sB * g.n;
        f32_to_f64_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_post, d_post64, cnt);
        QB_CUDA(cudaGetLastError());
        QB_CUDA(cudaMemcpyAsync(values_h, d_post64, cnt * 8, cudaMemcpyDeviceToHost, st));
    }
    QB_CUDA(cudaStreamSynchronize(st));
    return QB_OK;
}

int qb_minsum_core_host(qb_decoder *dec, const double *Q_h, const double *ssign_h, int32_t B, double alpha,
                        double *R_h, double *Rsum_h)
{
    QB_REQUIRE(dec && Q_h && ssign_h && R_h && Rsum_h, "NULL argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    if (int rc = dec->scratch.ensure(carve_size({sB * g.nnz * 8, sB * g.m * 8, sB * g.nnz * 8, sB * g.n * 8}))) return rc;
    Carver cv(dec->scratch.ptr);
    double *dQ = cv.take<double>(sB * g.nnz), *dS = cv.take<double>(sB * g.m), *dR = cv.take<double>(sB * g.nnz), *dRs = cv.take<double>(sB * g.n);
    if (g.nnz) QB_CUDA(cudaMemcpy(dQ, Q_h, sB * g.nnz * 8, cudaMemcpyHostToDevice));
    if (g.m) QB_CUDA(cudaMemcpy(dS, ssign_h, sB * g.m * 8, cudaMemcpyHostToDevice));
    if (g.nnz) QB_CUDA(cudaMemset(dR, 0, sB * g.nnz * 8));
    if (int rc = launch_minsum_core(dec, dQ, dS, B, alpha, dR, dRs, 0)) return rc;
    if (g.nnz) QB_CUDA(cudaMemcpy(R_h, dR, sB * g.nnz * 8, cudaMemcpyDeviceToHost));
    if (g.n) QB_CUDA(cudaMemcpy(Rsum_h, dRs, sB * g.n * 8, cudaMemcpyDeviceToHost));
    return QB_OK;
}

int qb_alpha_messages_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, const double *prior_h, int32_t n_prev,
                           const double *alpha_prev_h, double damping, double clip_llr, double *R_h)
{
    QB_REQUIRE(dec && syndrome_h && prior_h && R_h && n_prev >= 0 && (n_prev == 0 || alpha_prev_h), "bad argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    if (int rc = dec->scratch.ensure(carve_size({sB * g.m, (size_t)g.n * 8, (size_t)std::max(1, n_prev) * 8, sB * g.nnz * 8}))) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_syn = cv.take<int8_t>(sB * g.m);
    double *d_prior = cv.take<double>(g.n), *d_alpha = cv.take<double>(std::max(1, n_prev)), *d_R = cv.take<double>(sB * g.nnz);
    if (g.m) QB_CUDA(cudaMemcpy(d_syn, syndrome_h, sB * g.m, cudaMemcpyHostToDevice));
    if (g.n) QB_CUDA(cudaMemcpy(d_prior, prior_h, (size_t)g.n * 8, cudaMemcpyHostToDevice));
    if (n_prev) QB_CUDA(cudaMemcpy(d_alpha, alpha_prev_h, (size_t)n_prev * 8, cudaMemcpyHostToDevice));
    if (g.nnz) QB_CUDA(cudaMemset(d_R, 0, sB * g.nnz * 8));
    if (int rc = launch_alpha_messages(dec, d_syn, B, n_prev, d_alpha, damping, clip_llr, d_prior, d_R, 0)) return rc;
    if (g.nnz) QB_CUDA(cudaMemcpy(R_h, d_R, sB * g.nnz * 8, cudaMemcpyDeviceToHost));
    return QB_OK;
}

int qb_bp_decode_host(qb_decoder *dec, const int8_t *syndrome_h, int32_t B, int32_t max_iter, int8_t *hard_h,
                      uint8_t *converged_h, int32_t *final_iter_h, double *values_h)
{
    QB_REQUIRE(dec && syndrome_h && hard_h && converged_h && final_iter_h, "NULL argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    if (int rc = dec->scratch.ensure(carve_size({sB * g.m, sB * g.mw * 4, sB * g.nw * 4, sB, sB * 4, sB * g.n * 8, sB * g.n}))) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_syn8 = cv.take<int8_t>(sB * g.m);
    uint32_t *d_syn = cv.take<uint32_t>(sB * g.mw), *d_hard = cv.take<uint32_t>(sB * g.nw);
    uint8_t *d_conv = cv.take<uint8_t>(sB);
    int32_t *d_fin = cv.take<int32_t>(sB);
    double *d_post = cv.take<double>(sB * g.n);
    int8_t *d_hard8 = cv.take<int8_t>(sB * g.n);
    if (g.m) QB_CUDA(cudaMemcpy(d_syn8, syndrome_h, sB * g.m, cudaMemcpyHostToDevice));
    if (int rc = launch_pack_bits(d_syn8, B, g.m, d_syn, g.mw, 0)) return rc;
    if (int rc = launch_bp(dec, d_syn, B, max_iter, d_hard, d_conv, d_fin, d_post, 0)) return rc;
    if (int rc = launch_unpack_bits(d_hard, B, g.n, g.nw, d_hard8, 0)) return rc;
    if (g.n) QB_CUDA(cudaMemcpy(hard_h, d_hard8, sB * g.n, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(converged_h, d_conv, sB, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(final_iter_h, d_fin, sB * 4, cudaMemcpyDeviceToHost));
    if (values_h && g.n) QB_CUDA(cudaMemcpy(values_h, d_post, sB * g.n * 8, cudaMemcpyDeviceToHost));
    return QB_OK;
}

int qb_syndrome_check_host(qb_decoder *dec, const int8_t *candidate_h, int32_t B, int8_t *syndrome_h)
{
    QB_REQUIRE(dec && candidate_h && syndrome_h, "NULL argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    if (int rc = dec->scratch.ensure(carve_size({sB * g.n, sB * g.nw * 4, sB * g.mw * 4, sB * g.m}))) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_c8 = cv.take<int8_t>(sB * g.n);
    uint32_t *d_c = cv.take<uint32_t>(sB * g.nw), *d_s = cv.take<uint32_t>(sB * g.mw);
    int8_t *d_s8 = cv.take<int8_t>(sB * g.m);
    if (g.n) QB_CUDA(cudaMemcpy(d_c8, candidate_h, sB * g.n, cudaMemcpyHostToDevice));
    if (int rc = launch_pack_bits(d_c8, B, g.n, d_c, g.nw, 0)) return rc;
    if (int rc = launch_syndrome_check(dec, d_c, B, d_s, 0)) return rc;
    if (int rc = launch_unpack_bits(d_s, B, g.m, g.mw, d_s8, 0)) return rc;
    if (g.m) QB_CUDA(cudaMemcpy(syndrome_h, d_s8, sB * g.m, cudaMemcpyDeviceToHost));
    return QB_OK;
}

int qb_osd0_batch(qb_decoder *dec, const uint32_t *syn_bits_d, uint32_t *hard_bits_d, const float *post_d,
                  const int32_t *fail_idx_d, int32_t F, const int32_t *n_fail_d, int32_t max_fail, void *stream)
{
    QB_REQUIRE(dec && syn_bits_d && hard_bits_d && post_d, "NULL argument");
    QB_CUDA(cudaSetDevice(dec->device));
    OsdLaunch a{};
    a.syn_bits = syn_bits_d; a.hard_bits = hard_bits_d; a.post = post_d; a.fail_idx = fail_idx_d;
    a.F = F >= 0 ? F : max_fail; a.n_fail_d = F >= 0 ? nullptr : n_fail_d;
    return launch_osd0(dec, a, static_cast<cudaStream_t>(stream));
}

int qb_osd0_host(qb_decoder *dec, const int8_t *syndrome_h, const int8_t *hard_h, const double *llr_h,
                 const int32_t *ordering_h, int32_t B, int64_t *solution_h, int32_t *rank_h, int32_t *pivots_h)
{
    QB_REQUIRE(dec && syndrome_h && hard_h && solution_h && (llr_h || ordering_h), "NULL argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    const int rcap = std::min(g.m, g.n);
    if (int rc = dec->scratch.ensure(carve_size({sB * g.m, sB * g.n, sB * g.mw * 4, sB * g.nw * 4, sB * g.n * 4, sB * g.n * 4,
                                                 sB * g.n * 8, sB * 4, sB * (size_t)std::max(1, rcap) * 4}))) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_syn8 = cv.take<int8_t>(sB * g.m), *d_hard8 = cv.take<int8_t>(sB * g.n);
    uint32_t *d_syn = cv.take<uint32_t>(sB * g.mw), *d_hard = cv.take<uint32_t>(sB * g.nw);
    float *d_post = cv.take<float>(sB * g.n);
    int32_t *d_ord = cv.take<int32_t>(sB * g.n);
    int64_t *d_sol = cv.take<int64_t>(sB * g.n);
    int32_t *d_rank = cv.take<int32_t>(sB);
    int32_t *d_piv = cv.take<int32_t>(sB * std::max(1, rcap));
    if (g.m) QB_CUDA(cudaMemcpy(d_syn8, syndrome_h, sB * g.m, cudaMemcpyHostToDevice));
    if (g.n) QB_CUDA(cudaMemcpy(d_hard8, hard_h, sB * g.n, cudaMemcpyHostToDevice));
    if (int rc = launch_pack_bits(d_syn8, B, g.m, d_syn, g.mw, 0)) return rc;
    if (int rc = launch_pack_bits(d_hard8, B, g.n, d_hard, g.nw, 0)) return rc;
    if (ordering_h) {
        for (size_t i = 0; i < sB * g.n; ++i) QB_REQUIRE(ordering_h[i] >= 0 && ordering_h[i] < g.n, "ordering entry out of range");
        if (g.n) QB_CUDA(cudaMemcpy(d_ord, ordering_h, sB * g.n * 4, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> pf(sB * g.n);
        for (size_t i = 0; i < pf.size(); ++i) pf[i] = (float)llr_h[i];
        if (g.n) QB_CUDA(cudaMemcpy(d_post, pf.data(), pf.size() * 4, cudaMemcpyHostToDevice));
    }
    OsdLaunch a{};
    a.syn_bits = d_syn; a.hard_bits = d_hard; a.post = d_post; a.ordering = ordering_h ? d_ord : nullptr;
    a.fail_idx = nullptr; a.F = B; a.rank_out = d_rank; a.pivots_out = d_piv;
    a.exact_rows = 1;        // arbitrary (possibly inconsistent) syndromes: reproduce the reference's pivot rows
    if (int rc = launch_osd0(dec, a, 0)) return rc;
    const size_t cnt = sB * g.n;
    if (cnt) {
        hard_to_i64_kernel<<<(unsigned)((cnt + 255) / 256), 256>>>(d_hard, B, g.n, g.nw, d_sol);
        QB_CUDA(cudaGetLastError());
        QB_CUDA(cudaMemcpy(solution_h, d_sol, cnt * 8, cudaMemcpyDeviceToHost));
    }
    if (rank_h) QB_CUDA(cudaMemcpy(rank_h, d_rank, sB * 4, cudaMemcpyDeviceToHost));
    if (pivots_h && rcap) QB_CUDA(cudaMemcpy(pivots_h, d_piv, sB * rcap * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaDeviceSynchronize());
    return QB_OK;
}

int qb_decoder_osd_stats(qb_decoder *dec, int32_t *out10_h)
{
    QB_REQUIRE(dec && out10_h, "NULL argument");
    QB_CUDA(cudaSetDevice(dec->device));
    return osd_free_stats(dec, out10_h);
}

int qb_osd0_pipeline_host(qb_decoder *dec, const int8_t *syndrome_h, const int8_t *hard_h, const float *post_h, int32_t B,
                          int8_t *solution_h, int32_t *osd_info_h)
{
    QB_REQUIRE(dec && syndrome_h && hard_h && post_h && solution_h, "NULL argument");
    if (B <= 0) return QB_OK;
    QB_CUDA(cudaSetDevice(dec->device));
    const GraphDev &g = dec->g;
    const size_t sB = (size_t)B;
    if (int rc = dec->scratch.ensure(carve_size({sB * g.m, sB * g.n, sB * g.mw * 4, sB * g.nw * 4, sB * g.n * 4, sB * 4, 4}))) return rc;
    Carver cv(dec->scratch.ptr);
    int8_t *d_syn8 = cv.take<int8_t>(sB * g.m), *d_hard8 = cv.take<int8_t>(sB * g.n);
    uint32_t *d_syn = cv.take<uint32_t>(sB * g.mw), *d_hard = cv.take<uint32_t>(sB * g.nw);
    float *d_post = cv.take<float>(sB * g.n);
    int32_t *d_info = cv.take<int32_t>(sB), *d_nfail = cv.take<int32_t>(1);
    if (g.m) QB_CUDA(cudaMemcpy(d_syn8, syndrome_h, sB * g.m, cudaMemcpyHostToDevice));
    if (g.n) QB_CUDA(cudaMemcpy(d_hard8, hard_h, sB * g.n, cudaMemcpyHostToDevice));
    if (g.n) QB_CUDA(cudaMemcpy(d_post, post_h, sB * g.n * 4, cudaMemcpyHostToDevice));
    QB_CUDA(cudaMemset(d_info, 0, sB * 4));
    QB_CUDA(cudaMemcpy(d_nfail, &B, 4, cudaMemcpyHostToDevice));
    if (int rc = launch_pack_bits(d_syn8, B, g.m, d_syn, g.mw, 0)) return rc;
    if (int rc = launch_pack_bits(d_hard8, B, g.n, d_hard, g.nw, 0)) return rc;
    OsdLaunch a{};                       // exactly what decode_batch() issues for one side, every shot in the queue
    a.syn_bits = d_syn; a.hard_bits = d_hard; a.post = d_post; a.fail_idx = nullptr; a.F = B; a.n_fail_d = d_nfail;
    a.rank_out = d_info; a.rank_tag = 2 << 16;
    if (int rc = launch_osd0(dec, a, 0)) return rc;
    if (int rc = launch_unpack_bits(d_hard, B, g.n, g.nw, d_hard8, 0)) return rc;
    if (g.n) QB_CUDA(cudaMemcpy(solution_h, d_hard8, sB * g.n, cudaMemcpyDeviceToHost));
    if (osd_info_h) QB_CUDA(cudaMemcpy(osd_info_h, d_info, sB * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaDeviceSynchronize());
    return QB_OK;
}

int qb_gf2_eliminate_host(int device, int64_t *A_h, int64_t *b_h, int32_t m, int32_t n, uint64_t *A_packed_h,
                          int64_t *pivot_rows_h, int64_t *pivot_cols_h, int32_t *num_pivots_h)
{
    QB_REQUIRE(A_h && b_h && pivot_rows_h && pivot_cols_h && num_pivots_h && m >= 0 && n >= 0, "bad argument");
    QB_CUDA(cudaSetDevice(device));
    *num_pivots_h = 0;
    if (m == 0 || n == 0) return QB_OK;
    const int nw64 = (n + 63) / 64, nw = nw64 * 2, bw = ceil_div(m, 32);
    std::vector<uint32_t> Ap((size_t)m * nw, 0u), bp(bw, 0u);
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c) if (A_h[(size_t)r * n + c] & 1) Ap[(size_t)r * nw + (c >> 5)] |= 1u << (c & 31);
        if (b_h[r] & 1) bp[r >> 5] |= 1u << (r & 31);
    }
    const int mn = std::min(m, n);
    // one allocation, released on every path
    struct DevBuf { void *p = nullptr; ~DevBuf() { if (p) cudaFree(p); } } buf;
    QB_CUDA(cudaMalloc(&buf.p, carve_size({Ap.size() * 4, bp.size() * 4, (size_t)mn * 4, (size_t)mn * 4, 4})));
    Carver cv(buf.p);
    uint32_t *dA = cv.take<uint32_t>(Ap.size()), *db = cv.take<uint32_t>(bp.size());
    int32_t *dpr = cv.take<int32_t>(mn), *dpc = cv.take<int32_t>(mn), *dnp = cv.take<int32_t>(1);
    QB_CUDA(cudaMemcpy(dA, Ap.data(), Ap.size() * 4, cudaMemcpyHostToDevice));
    QB_CUDA(cudaMemcpy(db, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    if (int rc = launch_gf2_dense(dA, db, m, n, nw, dpr, dpc, dnp, 0)) return rc;
    std::vector<int32_t> pr(mn), pc(mn);
    QB_CUDA(cudaMemcpy(Ap.data(), dA, Ap.size() * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(bp.data(), db, bp.size() * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(pr.data(), dpr, mn * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(pc.data(), dpc, mn * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(num_pivots_h, dnp, 4, cudaMemcpyDeviceToHost));
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c) A_h[(size_t)r * n + c] = (Ap[(size_t)r * nw + (c >> 5)] >> (c & 31)) & 1u;
        b_h[r] = (bp[r >> 5] >> (r & 31)) & 1u;
        if (A_packed_h)
            for (int w = 0; w < nw64; ++w)
                A_packed_h[(size_t)r * nw64 + w] = (uint64_t)Ap[(size_t)r * nw + 2 * w] | ((uint64_t)Ap[(size_t)r * nw + 2 * w + 1] << 32);
    }
    for (int i = 0; i < *num_pivots_h; ++i) { pivot_rows_h[i] = pr[i]; pivot_cols_h[i] = pc[i]; }
    return QB_OK;
}

// ---- sampler ------------------------------------------------------------------------------------
int qb_sampler_create(int device, int32_t L, const int32_t *loc_kind, const int32_t *loc_colZ, const int32_t *loc_colX,
                      int32_t mZ, int32_t nZ, const int32_t *colptrZ, const int32_t *rowsZ, const uint32_t *logmaskZ,
                      int32_t mX, int32_t nX, const int32_t *colptrX, const int32_t *rowsX, const uint32_t *logmaskX,
                      int32_t k, qb_sampler **out)
{
    QB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    QB_REQUIRE(L >= 0 && L < (1 << 24), "number of fault locations must be below 2^24");
    QB_REQUIRE(loc_kind && loc_colZ && loc_colX && colptrZ && colptrX && logmaskZ && logmaskX, "NULL table");
    QB_REQUIRE(k >= 0 && k <= 32, "at most 32 logical observables");
    for (int i = 0; i < L; ++i) {
        QB_REQUIRE(loc_kind[i] >= 0 && loc_kind[i] <= 3, "bad location kind");
        for (int v = 0; v < 4; ++v) {
            QB_REQUIRE(loc_colZ[i * 4 + v] >= -1 && loc_colZ[i * 4 + v] < nZ, "Z column out of range");
            QB_REQUIRE(loc_colX[i * 4 + v] >= -1 && loc_colX[i * 4 + v] < nX, "X column out of range");
        }
    }
    for (int p = 0; p < colptrZ[nZ]; ++p) QB_REQUIRE(rowsZ[p] >= 0 && rowsZ[p] < mZ, "Z row out of range");
    for (int p = 0; p < colptrX[nX]; ++p) QB_REQUIRE(rowsX[p] >= 0 && rowsX[p] < mX, "X row out of range");
    QB_CUDA(cudaSetDevice(device));
    qb_sampler *s = new qb_sampler();
    s->device = device; s->L = L; s->k = k; s->mZ = mZ; s->nZ = nZ; s->mX = mX; s->nX = nX;
    s->mwZ = std::max(1, ceil_div(mZ, 32)); s->mwX = std::max(1, ceil_div(mX, 32));
    cudaDeviceProp prop;
    QB_CUDA(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    std::vector<int8_t> kind(L);
    for (int i = 0; i < L; ++i) kind[i] = (int8_t)loc_kind[i];
    int rc = to_device(s->owned, kind, &s->d_kind);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(loc_colZ, loc_colZ + (size_t)L * 4), &s->d_colZ);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(loc_colX, loc_colX + (size_t)L * 4), &s->d_colX);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(colptrZ, colptrZ + nZ + 1), &s->d_cpZ);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(rowsZ, rowsZ + colptrZ[nZ]), &s->d_rowZ);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(colptrX, colptrX + nX + 1), &s->d_cpX);
    if (!rc) rc = to_device(s->owned, std::vector<int32_t>(rowsX, rowsX + colptrX[nX]), &s->d_rowX);
    if (!rc) rc = to_device(s->owned, std::vector<uint32_t>(logmaskZ, logmaskZ + nZ), &s->d_lmZ);
    if (!rc) rc = to_device(s->owned, std::vector<uint32_t>(logmaskX, logmaskX + nX), &s->d_lmX);
    if (rc) { qb_sampler_destroy(s); return rc; }
    *out = s;
    return QB_OK;
}

void qb_sampler_destroy(qb_sampler *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    for (void *p : s->owned) cudaFree(p);
    s->scratch.release();
    delete s;
}

int qb_syndrome_from_events(qb_sampler *s, const int32_t *ev_ptr_d, const uint32_t *events_d, int32_t B,
                            uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX, void *stream)
{
    QB_REQUIRE(s && ev_ptr_d && synZ && trueZ && synX && trueX, "NULL argument");
    QB_CUDA(cudaSetDevice(s->device));
    return launch_events_syndrome(s, ev_ptr_d, events_d, B, synZ, trueZ, synX, trueX, static_cast<cudaStream_t>(stream));
}

// host fault-event lists (CSR over shots): offsets start at 0 and never decrease, so that the kernel stays inside
// events[0 .. ev_ptr[B]); the kernel itself ignores events whose location is >= L
static int check_events(const qb_sampler *, const int32_t *ev_ptr_h, const uint32_t *events_h, int B)
{
    QB_REQUIRE(ev_ptr_h[0] == 0, "ev_ptr[0] must be 0");
    for (int b = 0; b < B; ++b) QB_REQUIRE(ev_ptr_h[b + 1] >= ev_ptr_h[b], "ev_ptr must be non-decreasing");
    QB_REQUIRE(ev_ptr_h[B] == 0 || events_h != nullptr, "events is NULL");
    return QB_OK;
}

int qb_syndrome_from_events_host(qb_sampler *s, const int32_t *ev_ptr_h, const uint32_t *events_h, int32_t B,
                                 int8_t *sparseZ_h, int8_t *trueZ_h, int8_t *sparseX_h, int8_t *trueX_h)
{
    QB_REQUIRE(s && ev_ptr_h && sparseZ_h && trueZ_h && sparseX_h && trueX_h, "NULL argument");
    if (B <= 0) return QB_OK;
    if (int rc = check_events(s, ev_ptr_h, events_h, B)) return rc;
    QB_CUDA(cudaSetDevice(s->device));
    const size_t sB = (size_t)B, nev = (size_t)ev_ptr_h[B];
    if (int rc = s->scratch.ensure(carve_size({(sB + 1) * 4, nev * 4, sB * s->mwZ * 4, sB * s->mwX * 4, sB * 4, sB * 4,
                                               sB * s->mZ, sB * s->mX}))) return rc;
    Carver cv(s->scratch.ptr);
    int32_t *d_ptr = cv.take<int32_t>(sB + 1);
    uint32_t *d_ev = cv.take<uint32_t>(nev), *d_sz = cv.take<uint32_t>(sB * s->mwZ), *d_sx = cv.take<uint32_t>(sB * s->mwX);
    uint32_t *d_tz = cv.take<uint32_t>(sB), *d_tx = cv.take<uint32_t>(sB);
    int8_t *d_z8 = cv.take<int8_t>(sB * s->mZ), *d_x8 = cv.take<int8_t>(sB * s->mX);
    QB_CUDA(cudaMemcpy(d_ptr, ev_ptr_h, (sB + 1) * 4, cudaMemcpyHostToDevice));
    if (nev) QB_CUDA(cudaMemcpy(d_ev, events_h, nev * 4, cudaMemcpyHostToDevice));
    if (int rc = launch_events_syndrome(s, d_ptr, d_ev, B, d_sz, d_tz, d_sx, d_tx, 0)) return rc;
    if (int rc = launch_unpack_bits(d_sz, B, s->mZ, s->mwZ, d_z8, 0)) return rc;
    if (int rc = launch_unpack_bits(d_sx, B, s->mX, s->mwX, d_x8, 0)) return rc;
    if (s->mZ) QB_CUDA(cudaMemcpy(sparseZ_h, d_z8, sB * s->mZ, cudaMemcpyDeviceToHost));
    if (s->mX) QB_CUDA(cudaMemcpy(sparseX_h, d_x8, sB * s->mX, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> tz(B), tx(B);
    QB_CUDA(cudaMemcpy(tz.data(), d_tz, sB * 4, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(tx.data(), d_tx, sB * 4, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < s->k; ++i) {
            trueZ_h[(size_t)b * s->k + i] = (int8_t)((tz[b] >> i) & 1u);
            trueX_h[(size_t)b * s->k + i] = (int8_t)((tx[b] >> i) & 1u);
        }
    return QB_OK;
}

int qb_sampler_geometric_table(double p, uint32_t *table_h, int32_t capacity)
{
    QB_REQUIRE(table_h != nullptr && p > 0.0 && p < 1.0, "bad argument");
    std::vector<uint32_t> T;
    geometric_table(p, T);
    QB_REQUIRE(capacity >= (int)T.size(), "table capacity too small (qb_sampler_geometric_table needs 1024 words)");
    memcpy(table_h, T.data(), sizeof(uint32_t) * T.size());
    return (int)T.size();
}

int qb_sample_syndromes(qb_sampler *s, uint64_t seed, uint64_t first_shot, int32_t B, double p,
                        uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX, int32_t *nfaults_d, void *stream)
{
    QB_REQUIRE(s && synZ && trueZ && synX && trueX, "NULL argument");
    QB_CUDA(cudaSetDevice(s->device));
    return launch_sample_syndrome(s, seed, first_shot, B, p, synZ, trueZ, synX, trueX, nfaults_d, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---- pipeline -------------------------------------------------------------------------------------
// Per-batch buffers.  A pipeline owns two of them: consecutive batches of one qb_pipeline_run call alternate, each on
// its own stream, so that the sampling and min-sum of batch i+1 fill the SMs the OSD-0 tail of batch i no longer uses.
struct qb_workspace {
    cudaStream_t st = nullptr;        // stream the batch is issued on (workspace 0: the pipeline's / the caller's stream)
    cudaStream_t own_st = nullptr;
    cudaStream_t side_st = nullptr;   // the X side's OSD-0 runs beside the Z side's and fills its tail
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_ms_done = nullptr, ev_osd_done = nullptr;   // min-sum / OSD of the batch in this workspace finished
    uint32_t *synZ = nullptr, *synX = nullptr, *trueZ = nullptr, *trueX = nullptr, *hardZ = nullptr, *hardX = nullptr;
    uint8_t *convZ = nullptr, *convX = nullptr, *flags = nullptr;
    int32_t *itZ = nullptr, *itX = nullptr, *failZ = nullptr, *failX = nullptr, *nfail = nullptr;   // nfail[2]
    int32_t *fwZ = nullptr, *fwX = nullptr, *sortZ = nullptr, *sortX = nullptr;   // failure weights / weight-sorted queues
    float *postZ = nullptr, *postX = nullptr;
    int32_t *osdinfoZ = nullptr, *osdinfoX = nullptr;   // per shot: pivots used | OSD path << 16 (detail mode only)
    bool ms_pending = false, osd_pending = false;        // the events above have been recorded in the current run
};

struct qb_pipeline {
    qb_sampler *s = nullptr;
    qb_decoder *dz = nullptr, *dx = nullptr;
    int max_batch = 0;
    cudaStream_t st = nullptr;        // stream of workspace 0 (own or the caller's)
    cudaStream_t own_st = nullptr;
    qb_workspace ws[2];
    int n_ws = 2;
    cudaEvent_t ev_begin = nullptr, ev_tail = nullptr;   // run start (second stream waits for it) / second stream drained
    int64_t *counts = nullptr;   // device [8]
    int32_t *ev_ptr = nullptr; uint32_t *events = nullptr; size_t ev_cap = 0;
    int32_t *ev_ptr_big = nullptr; size_t ev_ptr_cap = 0;      // event offsets of a whole qb_pipeline_run_events_host call
    int8_t *syn8 = nullptr;      // staging for decode_host
    int detail = 0;              // qb_pipeline_enable_detail
    int last_B = 0, last_ws = 0; // shots / workspace of the last batch (what qb_pipeline_last_batch_detail may read)
    std::vector<void *> owned;
    std::vector<cudaEvent_t> evs;
    qb_pipeline_stats stats{};
};

namespace qb {

enum { EV_START = 0, EV_SAMPLED, EV_MINSUM, EV_OSD, EV_END, EV_PER_BATCH };
constexpr int MAX_TIMED_BATCHES = 256;

template <class T> static int dalloc(qb_pipeline *p, T **ptr, size_t count)
{
    QB_CUDA(cudaMalloc(reinterpret_cast<void **>(ptr), sizeof(T) * std::max<size_t>(1, count)));
    p->owned.push_back(*ptr);
    return QB_OK;
}

static int check_cfg(const qb_decode_config *cfg)
{
    QB_REQUIRE(cfg != nullptr, "cfg is NULL");
    QB_REQUIRE(cfg->max_iter >= 0, "max_iter must be >= 0");
    QB_REQUIRE(cfg->precision == QB_PRECISION_F32 || cfg->precision == QB_PRECISION_HALF2, "unknown precision");
    if (int rc = check_alpha(cfg->alpha_mode, cfg->alpha_z, cfg->alpha_seq_z_h, cfg->alpha_len_z)) return rc;
    if (int rc = check_alpha(cfg->alpha_mode, cfg->alpha_x, cfg->alpha_seq_x_h, cfg->alpha_len_x)) return rc;
    return QB_OK;
}

static int prep_alpha(qb_pipeline *p, const qb_decode_config *cfg)
{
    std::vector<double> sz, sx;
    if (cfg->alpha_mode == QB_ALPHA_SEQUENCE) {
        sz.assign(cfg->alpha_seq_z_h, cfg->alpha_seq_z_h + cfg->alpha_len_z);
        sx.assign(cfg->alpha_seq_x_h, cfg->alpha_seq_x_h + cfg->alpha_len_x);
    }
    if (int rc = upload_alpha(p->dz, cfg->max_iter, cfg->alpha_mode, cfg->alpha_z, sz.data(), cfg->alpha_len_z, p->st)) return rc;
    return upload_alpha(p->dx, cfg->max_iter, cfg->alpha_mode, cfg->alpha_x, sx.data(), cfg->alpha_len_x, p->st);
}

// start of a run: counters cleared, alpha tables uploaded on workspace 0's stream; the second stream starts after that
static int begin_run(qb_pipeline *p, const qb_decode_config *cfg)
{
    p->stats = qb_pipeline_stats{};
    p->ws[0].st = p->st;
    for (int w = 0; w < 2; ++w) { p->ws[w].ms_pending = false; p->ws[w].osd_pending = false; }
    if (int rc = prep_alpha(p, cfg)) return rc;
    QB_CUDA(cudaMemsetAsync(p->counts, 0, 8 * sizeof(int64_t), p->st));
    QB_CUDA(cudaEventRecord(p->ev_begin, p->st));
    QB_CUDA(cudaStreamWaitEvent(p->ws[1].st, p->ev_begin, 0));
    return QB_OK;
}

// decode + logical check of one batch whose syndromes / true masks already sit in the workspace.  The decoders' own
// scratch (shot queue of the persistent min-sum kernel, OSD work areas) is shared by the two workspaces: min-sum of this
// batch waits for the other workspace's min-sum, OSD for its OSD -- which also is the order that keeps the SMs busy.
static int decode_batch(qb_pipeline *p, qb_workspace &W, qb_workspace &other, int B, const qb_decode_config *cfg, int batch_no)
{
    cudaStream_t st = W.st;
    const bool timed = batch_no < MAX_TIMED_BATCHES;
    cudaEvent_t *ev = timed ? &p->evs[(size_t)batch_no * EV_PER_BATCH] : nullptr;
    p->last_B = B; p->last_ws = (int)(&W - p->ws);
    QB_CUDA(cudaMemsetAsync(W.nfail, 0, 2 * sizeof(int32_t), st));
    if (p->detail) {
        QB_CUDA(cudaMemsetAsync(W.osdinfoZ, 0, (size_t)B * sizeof(int32_t), st));
        QB_CUDA(cudaMemsetAsync(W.osdinfoX, 0, (size_t)B * sizeof(int32_t), st));
    }
    if (other.ms_pending) QB_CUDA(cudaStreamWaitEvent(st, other.ev_ms_done, 0));
    for (int side = 0; side < 2; ++side) {
        qb_decoder *d = side ? p->dx : p->dz;
        MinsumLaunch a{};
        a.syn_bits = side ? W.synX : W.synZ; a.B = B; a.max_iter = cfg->max_iter; a.alpha_d = d->d_alpha;
        a.damping = 1.0f; a.clip = cfg->clip_llr; a.dense_variant = 0;
        a.hard_bits = side ? W.hardX : W.hardZ; a.converged = side ? W.convX : W.convZ;
        a.final_iter = side ? W.itX : W.itZ; a.post = side ? W.postX : W.postZ; a.post_failed_only = 1;
        a.fail_count = W.nfail + side; a.fail_idx = side ? W.failX : W.failZ; a.fail_wt = side ? W.fwX : W.fwZ;
        a.precision = cfg->precision;
        if (!cfg->use_osd && getenv("QLDPC_B200_NO_POST")) a.post = nullptr;     // measurement aid: min-sum without the posterior output
        if (int rc = launch_minsum(d, a, st)) return rc;
        p->stats.kernel_launches++;
    }
    QB_CUDA(cudaEventRecord(W.ev_ms_done, st)); W.ms_pending = true;
    if (timed) QB_CUDA(cudaEventRecord(ev[EV_MINSUM], st));
    if (cfg->use_osd) {
        if (other.osd_pending) QB_CUDA(cudaStreamWaitEvent(st, other.ev_osd_done, 0));
        // The two sides are independent (own decoders, own work areas): the X side is issued on a second stream, so
        // that its CTAs take the SMs the Z launch frees while its last, longest eliminations finish.
        const bool fork = p->dz != p->dx && W.side_st != nullptr;
        if (fork) { QB_CUDA(cudaEventRecord(W.ev_fork, st)); QB_CUDA(cudaStreamWaitEvent(W.side_st, W.ev_fork, 0)); }
        for (int side = 0; side < 2; ++side) {
            qb_decoder *d = side ? p->dx : p->dz;
            cudaStream_t ss = (fork && side) ? W.side_st : st;
            if (int rc = launch_sort_failures(side ? W.failX : W.failZ, side ? W.fwX : W.fwZ, W.nfail + side,
                                              side ? W.sortX : W.sortZ, ss)) return rc;
            p->stats.kernel_launches++;
            OsdLaunch a{};
            a.syn_bits = side ? W.synX : W.synZ; a.hard_bits = side ? W.hardX : W.hardZ;
            a.post = side ? W.postX : W.postZ; a.fail_idx = side ? W.sortX : W.sortZ;
            a.F = B; a.n_fail_d = W.nfail + side;
            if (p->detail) { a.rank_out = side ? W.osdinfoX : W.osdinfoZ; a.rank_tag = 2 << 16; }
            if (int rc = launch_osd0(d, a, ss)) return rc;
            p->stats.kernel_launches += osd_launches_per_call(d);
        }
        if (fork) { QB_CUDA(cudaEventRecord(W.ev_join, W.side_st)); QB_CUDA(cudaStreamWaitEvent(st, W.ev_join, 0)); }
        QB_CUDA(cudaEventRecord(W.ev_osd_done, st)); W.osd_pending = true;
    }
    if (timed) QB_CUDA(cudaEventRecord(ev[EV_OSD], st));
    if (int rc = launch_logical_check(p->dz, W.hardZ, W.trueZ, B, W.flags, 0, p->counts, 0, st)) return rc;
    if (int rc = launch_logical_check(p->dx, W.hardX, W.trueX, B, W.flags, 1, p->counts, 1, st)) return rc;
    batch_counts_kernel<<<std::max(1, std::min(ceil_div(B, 256), 256)), 256, 0, st>>>(
        W.flags, W.convZ, W.convX, W.itZ, W.itX, B, reinterpret_cast<unsigned long long *>(p->counts));
    QB_CUDA(cudaGetLastError());
    p->stats.kernel_launches += 3;
    if (timed) QB_CUDA(cudaEventRecord(ev[EV_END], st));
    return QB_OK;
}

static int finish_run(qb_pipeline *p, int n_batches, int64_t *counts_h)
{
    if (n_batches > 1) {      // the second stream's batches complete before the counters are read on workspace 0's stream
        QB_CUDA(cudaEventRecord(p->ev_tail, p->ws[1].st));
        QB_CUDA(cudaStreamWaitEvent(p->st, p->ev_tail, 0));
    }
    QB_CUDA(cudaMemcpyAsync(counts_h, p->counts, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, p->st));
    QB_CUDA(cudaStreamSynchronize(p->st));
    const int nb = std::min(n_batches, MAX_TIMED_BATCHES);
    float t_end = 0.f;
    for (int b = 0; b < nb; ++b) {
        cudaEvent_t *ev = &p->evs[(size_t)b * EV_PER_BATCH];
        float t;
        QB_CUDA(cudaEventElapsedTime(&t, ev[EV_START], ev[EV_SAMPLED])); p->stats.ms_sample += t;
        QB_CUDA(cudaEventElapsedTime(&t, ev[EV_SAMPLED], ev[EV_MINSUM])); p->stats.ms_minsum += t;
        QB_CUDA(cudaEventElapsedTime(&t, ev[EV_MINSUM], ev[EV_OSD])); p->stats.ms_osd += t;
        QB_CUDA(cudaEventElapsedTime(&t, ev[EV_OSD], ev[EV_END])); p->stats.ms_logical += t;
        QB_CUDA(cudaEventElapsedTime(&t, p->evs[EV_START], ev[EV_END])); t_end = std::max(t_end, t);
    }
    if (getenv("QLDPC_B200_TIMELINE"))
        for (int b = 0; b < nb; ++b) {
            cudaEvent_t *ev = &p->evs[(size_t)b * EV_PER_BATCH];
            float t[EV_PER_BATCH];
            for (int k = 0; k < EV_PER_BATCH; ++k) cudaEventElapsedTime(&t[k], p->evs[EV_START], ev[k]);
            fprintf(stderr, "batch %d ws %d: start %.2f sampled %.2f minsum %.2f osd %.2f end %.2f ms\n", b, b & 1, t[0], t[1], t[2], t[3], t[4]);
        }
    p->stats.ms_total = t_end;           // first batch's start to the last completion (stage times overlap across batches)
    p->stats.edge_messages = counts_h[6] * (int64_t)p->dz->g.nnz + counts_h[7] * (int64_t)p->dx->g.nnz;
    p->stats.osd_sides = counts_h[4] + counts_h[5];
    return QB_OK;
}

}  // namespace qb

extern "C" {

int qb_pipeline_create(qb_sampler *s, qb_decoder *decZ, qb_decoder *decX, int32_t max_batch, qb_pipeline **out)
{
    QB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    QB_REQUIRE(decZ && decX, "decoders are required");
    QB_REQUIRE(max_batch > 0, "max_batch must be positive");
    QB_REQUIRE(decZ->device == decX->device && (!s || s->device == decZ->device), "handles live on different devices");
    if (s) QB_REQUIRE(s->mZ == decZ->g.m && s->mX == decX->g.m, "sampler and decoder syndrome sizes differ");
    QB_CUDA(cudaSetDevice(decZ->device));
    qb_pipeline *p = new qb_pipeline();
    p->s = s; p->dz = decZ; p->dx = decX; p->max_batch = max_batch;
    const size_t B = (size_t)max_batch;
    const GraphDev &gz = decZ->g, &gx = decX->g;
    int rc = QB_OK;
    // OSD streams run at a higher priority than the streams that carry sampling and min-sum: when both have CTAs ready
    // (next batch's min-sum against this batch's OSD tail) the tail goes first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int w = 0; w < 2 && !rc; ++w) {
        qb_workspace &W = p->ws[w];
        if (cudaStreamCreateWithPriority(&W.own_st, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
            cudaStreamCreateWithPriority(&W.side_st, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&W.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&W.ev_join, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&W.ev_ms_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&W.ev_osd_done, cudaEventDisableTiming) != cudaSuccess) rc = QB_ERR_CUDA;
        W.st = W.own_st;
    }
    if (!rc && (cudaEventCreateWithFlags(&p->ev_begin, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_tail, cudaEventDisableTiming) != cudaSuccess)) rc = QB_ERR_CUDA;
    p->own_st = p->ws[0].own_st;
    p->st = p->own_st;
    // the second workspace costs a second set of posterior buffers (B x n floats per side): skipped for tiny pipelines
    // only when memory is short
    for (int w = 0; w < 2; ++w) {
        qb_workspace &W = p->ws[w];
#define AL(ptr, cnt) if (!rc) rc = dalloc(p, &W.ptr, cnt);
        AL(synZ, B * gz.mw) AL(synX, B * gx.mw) AL(trueZ, B) AL(trueX, B) AL(hardZ, B * gz.nw) AL(hardX, B * gx.nw)
        AL(convZ, B) AL(convX, B) AL(flags, B) AL(itZ, B) AL(itX, B) AL(failZ, B) AL(failX, B) AL(nfail, 2)
        AL(fwZ, B) AL(fwX, B) AL(sortZ, B) AL(sortX, B)
        AL(postZ, B * gz.n) AL(postX, B * gx.n) AL(osdinfoZ, B) AL(osdinfoX, B)
#undef AL
    }
    if (!rc) rc = dalloc(p, &p->counts, 8);
    if (!rc) rc = dalloc(p, &p->ev_ptr, B + 1);
    if (!rc) rc = dalloc(p, &p->syn8, B * std::max(gz.m, gx.m));
    if (!rc) {
        p->evs.resize((size_t)MAX_TIMED_BATCHES * EV_PER_BATCH);
        for (auto &e : p->evs) if (cudaEventCreate(&e) != cudaSuccess) { rc = QB_ERR_CUDA; break; }
    }
    if (rc) { if (rc == QB_ERR_CUDA) cuda_fail(cudaGetLastError(), "pipeline allocation", __FILE__, __LINE__); qb_pipeline_destroy(p); return rc; }
    *out = p;
    return QB_OK;
}

void qb_pipeline_destroy(qb_pipeline *p)
{
    if (!p) return;
    cudaSetDevice(p->dz->device);
    cudaDeviceSynchronize();
    for (void *q : p->owned) cudaFree(q);
    if (p->events) cudaFree(p->events);
    if (p->ev_ptr_big) cudaFree(p->ev_ptr_big);
    for (auto e : p->evs) if (e) cudaEventDestroy(e);
    for (int w = 0; w < 2; ++w) {
        qb_workspace &W = p->ws[w];
        for (cudaEvent_t e : {W.ev_fork, W.ev_join, W.ev_ms_done, W.ev_osd_done}) if (e) cudaEventDestroy(e);
        if (W.side_st) cudaStreamDestroy(W.side_st);
        if (W.own_st) cudaStreamDestroy(W.own_st);
    }
    if (p->ev_begin) cudaEventDestroy(p->ev_begin);
    if (p->ev_tail) cudaEventDestroy(p->ev_tail);
    delete p;
}

int qb_pipeline_set_stream(qb_pipeline *p, void *stream, int use_external)
{
    QB_REQUIRE(p != nullptr, "NULL argument");
    p->st = use_external ? static_cast<cudaStream_t>(stream) : p->own_st;
    p->ws[0].st = p->st;
    return QB_OK;
}

int qb_pipeline_run(qb_pipeline *p, uint64_t seed, uint64_t first_shot, int64_t n_shots, double error_rate,
                    const qb_decode_config *cfg, int64_t *counts_h, uint8_t *flags_h)
{
    QB_REQUIRE(p && counts_h, "NULL argument");
    QB_REQUIRE(p->s != nullptr, "pipeline was created without a sampler");
    QB_REQUIRE(n_shots >= 0, "n_shots must be >= 0");
    if (int rc = check_cfg(cfg)) return rc;
    QB_CUDA(cudaSetDevice(p->dz->device));
    if (int rc = begin_run(p, cfg)) return rc;
    int nb = 0;
    for (int64_t done = 0; done < n_shots; done += p->max_batch, ++nb) {
        const int B = (int)std::min<int64_t>(p->max_batch, n_shots - done);
        qb_workspace &W = p->ws[nb & 1], &O = p->ws[(nb & 1) ^ 1];
        if (nb < MAX_TIMED_BATCHES) QB_CUDA(cudaEventRecord(p->evs[(size_t)nb * EV_PER_BATCH + EV_START], W.st));
        if (int rc = launch_sample_syndrome(p->s, seed, first_shot + (uint64_t)done, B, error_rate, W.synZ, W.trueZ,
                                            W.synX, W.trueX, nullptr, W.st)) return rc;
        p->stats.kernel_launches++;
        if (nb < MAX_TIMED_BATCHES) QB_CUDA(cudaEventRecord(p->evs[(size_t)nb * EV_PER_BATCH + EV_SAMPLED], W.st));
        if (int rc = decode_batch(p, W, O, B, cfg, nb)) return rc;
        if (flags_h) QB_CUDA(cudaMemcpyAsync(flags_h + done, W.flags, (size_t)B, cudaMemcpyDeviceToHost, W.st));
    }
    return finish_run(p, nb, counts_h);
}

int qb_pipeline_run_events_host(qb_pipeline *p, const int32_t *ev_ptr_h, const uint32_t *events_h, int32_t B,
                                const qb_decode_config *cfg, int64_t *counts_h, uint8_t *flags_h,
                                uint8_t *converged_h, int32_t *final_iter_h)
{
    QB_REQUIRE(p && ev_ptr_h && counts_h, "NULL argument");
    QB_REQUIRE(p->s != nullptr, "pipeline was created without a sampler");
    QB_REQUIRE(B >= 0, "negative number of shots");
    QB_REQUIRE(B <= p->max_batch || (!converged_h && !final_iter_h), "per-shot detail is available for a single batch (B <= max_batch) only");
    if (int rc = check_cfg(cfg)) return rc;
    QB_CUDA(cudaSetDevice(p->dz->device));
    p->stats = qb_pipeline_stats{};
    if (B == 0) { memset(counts_h, 0, 8 * sizeof(int64_t)); return QB_OK; }
    if (int rc = check_events(p->s, ev_ptr_h, events_h, B)) return rc;
    const size_t nev = (size_t)ev_ptr_h[B];
    if (nev > p->ev_cap) {
        if (p->events) cudaFree(p->events);
        p->events = nullptr; p->ev_cap = 0;
        QB_CUDA(cudaMalloc(&p->events, (nev + nev / 2 + 1024) * 4));
        p->ev_cap = nev + nev / 2 + 1024;
    }
    if ((size_t)B + 1 > p->ev_ptr_cap) {
        if (p->ev_ptr_big) cudaFree(p->ev_ptr_big);
        p->ev_ptr_big = nullptr; p->ev_ptr_cap = 0;
        QB_CUDA(cudaMalloc(&p->ev_ptr_big, ((size_t)B + 1 + (size_t)B / 2) * 4));
        p->ev_ptr_cap = (size_t)B + 1 + (size_t)B / 2;
    }
    if (int rc = begin_run(p, cfg)) return rc;
    QB_CUDA(cudaEventRecord(p->evs[EV_START], p->st));
    // all fault events of the call go up in one copy; the batches (max_batch shots each) then alternate between the two
    // workspaces like in qb_pipeline_run, the event offsets stay absolute
    QB_CUDA(cudaMemcpyAsync(p->ev_ptr_big, ev_ptr_h, ((size_t)B + 1) * 4, cudaMemcpyHostToDevice, p->st));
    if (nev) QB_CUDA(cudaMemcpyAsync(p->events, events_h, nev * 4, cudaMemcpyHostToDevice, p->st));
    QB_CUDA(cudaEventRecord(p->ev_begin, p->st));
    QB_CUDA(cudaStreamWaitEvent(p->ws[1].st, p->ev_begin, 0));
    int nb = 0;
    for (int done = 0; done < B; done += p->max_batch, ++nb) {
        const int Bi = std::min(p->max_batch, B - done);
        qb_workspace &W = p->ws[nb & 1], &O = p->ws[(nb & 1) ^ 1];
        if (nb > 0 && nb < MAX_TIMED_BATCHES) QB_CUDA(cudaEventRecord(p->evs[(size_t)nb * EV_PER_BATCH + EV_START], W.st));
        if (int rc = launch_events_syndrome(p->s, p->ev_ptr_big + done, p->events, Bi, W.synZ, W.trueZ, W.synX, W.trueX, W.st)) return rc;
        p->stats.kernel_launches++;
        if (nb < MAX_TIMED_BATCHES) QB_CUDA(cudaEventRecord(p->evs[(size_t)nb * EV_PER_BATCH + EV_SAMPLED], W.st));
        if (int rc = decode_batch(p, W, O, Bi, cfg, nb)) return rc;
        if (flags_h) QB_CUDA(cudaMemcpyAsync(flags_h + done, W.flags, (size_t)Bi, cudaMemcpyDeviceToHost, W.st));
    }
    if (converged_h) {
        qb_workspace &W = p->ws[0];
        QB_CUDA(cudaMemcpyAsync(converged_h, W.convZ, (size_t)B, cudaMemcpyDeviceToHost, p->st));
        QB_CUDA(cudaMemcpyAsync(converged_h + B, W.convX, (size_t)B, cudaMemcpyDeviceToHost, p->st));
    }
    if (final_iter_h) {
        qb_workspace &W = p->ws[0];
        QB_CUDA(cudaMemcpyAsync(final_iter_h, W.itZ, (size_t)B * 4, cudaMemcpyDeviceToHost, p->st));
        QB_CUDA(cudaMemcpyAsync(final_iter_h + B, W.itX, (size_t)B * 4, cudaMemcpyDeviceToHost, p->st));
    }
    return finish_run(p, nb, counts_h);
}

int qb_pipeline_decode_host(qb_pipeline *p, const int8_t *sparseZ_h, const uint32_t *trueZ_h, const int8_t *sparseX_h,
                            const uint32_t *trueX_h, int32_t B, const qb_decode_config *cfg, int64_t *counts_h,
                            uint8_t *flags_h)
{
    QB_REQUIRE(p && sparseZ_h && trueZ_h && sparseX_h && trueX_h && counts_h, "NULL argument");
    QB_REQUIRE(B >= 0 && B <= p->max_batch, "batch exceeds the pipeline's max_batch");
    if (int rc = check_cfg(cfg)) return rc;
    QB_CUDA(cudaSetDevice(p->dz->device));
    p->stats = qb_pipeline_stats{};
    if (B == 0) { memset(counts_h, 0, 8 * sizeof(int64_t)); return QB_OK; }
    const GraphDev &gz = p->dz->g, &gx = p->dx->g;
    if (int rc = begin_run(p, cfg)) return rc;
    qb_workspace &W = p->ws[0];
    QB_CUDA(cudaEventRecord(p->evs[EV_START], p->st));
    QB_CUDA(cudaMemcpyAsync(p->syn8, sparseZ_h, (size_t)B * gz.m, cudaMemcpyHostToDevice, p->st));
    if (int rc = launch_pack_bits(p->syn8, B, gz.m, W.synZ, gz.mw, p->st)) return rc;
    QB_CUDA(cudaMemcpyAsync(p->syn8, sparseX_h, (size_t)B * gx.m, cudaMemcpyHostToDevice, p->st));
    if (int rc = launch_pack_bits(p->syn8, B, gx.m, W.synX, gx.mw, p->st)) return rc;
    QB_CUDA(cudaMemcpyAsync(W.trueZ, trueZ_h, (size_t)B * 4, cudaMemcpyHostToDevice, p->st));
    QB_CUDA(cudaMemcpyAsync(W.trueX, trueX_h, (size_t)B * 4, cudaMemcpyHostToDevice, p->st));
    p->stats.kernel_launches += 2;
    QB_CUDA(cudaEventRecord(p->evs[EV_SAMPLED], p->st));
    if (int rc = decode_batch(p, W, p->ws[1], B, cfg, 0)) return rc;
    if (flags_h) QB_CUDA(cudaMemcpyAsync(flags_h, W.flags, (size_t)B, cudaMemcpyDeviceToHost, p->st));
    return finish_run(p, 1, counts_h);
}

int qb_pipeline_enable_detail(qb_pipeline *p, int on)
{
    QB_REQUIRE(p != nullptr, "NULL argument");
    p->detail = on ? 1 : 0;
    return QB_OK;
}

int qb_pipeline_last_batch_detail(qb_pipeline *p, int32_t side, uint32_t *hard_bits_h, float *post_h, int32_t *osd_info_h)
{
    QB_REQUIRE(p != nullptr && (side == 0 || side == 1), "bad argument");
    QB_CUDA(cudaSetDevice(p->dz->device));
    QB_CUDA(cudaDeviceSynchronize());
    const size_t B = (size_t)p->last_B;
    if (B == 0) return QB_OK;
    const qb_workspace &W = p->ws[p->last_ws];
    const GraphDev &g = side ? p->dx->g : p->dz->g;
    if (hard_bits_h) QB_CUDA(cudaMemcpy(hard_bits_h, side ? W.hardX : W.hardZ, B * g.nw * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (post_h) QB_CUDA(cudaMemcpy(post_h, side ? W.postX : W.postZ, B * g.n * sizeof(float), cudaMemcpyDeviceToHost));
    if (osd_info_h) {
        QB_REQUIRE(p->detail, "per-shot OSD information is recorded only after qb_pipeline_enable_detail(p, 1)");
        QB_CUDA(cudaMemcpy(osd_info_h, side ? W.osdinfoX : W.osdinfoZ, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    return QB_OK;
}

int qb_pipeline_last_stats(qb_pipeline *p, qb_pipeline_stats *out)
{
    QB_REQUIRE(p && out, "NULL argument");
    *out = p->stats;
    return QB_OK;
}

}  // extern "C"
