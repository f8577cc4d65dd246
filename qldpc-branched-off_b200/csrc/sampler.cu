// K1 (Philox fault sampler) and K2 (bit-packed syndrome + true-logical accumulation).
//
// Reference semantics: run_trial_fast (src/noise/simulation.py:21-107) =
//   generate_noisy_circuit_jit (src/noise/kernels.py:176-353) + Pauli-frame propagation
//   (:14-172) + detector differencing (:357-380) + logical product (simulation.py:81,99).
// Propagation is GF(2)-linear, so the syndrome / logical flips of a shot are the XOR of the
// signatures of its faults; the signatures are exactly the columns of the decoding matrices
// (src/noise/builder.py:115-124).  One warp owns one shot; syndromes live in shared memory as
// bit-packed words and faults XOR their <= 6 rows in with shared-memory atomics.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qb {

// variant (component on q1 + 2*component on q2) per two-qubit outcome, reference order
// Xc,Yc,Zc,Xt,Yt,Zt,XX,YY,ZZ,XY,YX,YZ,ZY,XZ,ZX  (noise/kernels.py:283-342)
__constant__ uint8_t c_cnot_zvar[16] = {0, 1, 1, 0, 2, 2, 0, 3, 3, 2, 1, 3, 3, 2, 1, 1};
__constant__ uint8_t c_cnot_xvar[16] = {1, 1, 0, 2, 2, 0, 3, 3, 0, 3, 3, 1, 2, 1, 2, 2};

struct SamplerDev {
    int L, k, mwZ, mwX;
    const int8_t *kind;
    const int32_t *colZ, *colX;
    const int32_t *cpZ, *rowZ, *cpX, *rowX;
    const uint32_t *lmZ, *lmX;
};

__device__ __forceinline__ void apply_fault(const SamplerDev &s, int loc, int outcome, uint32_t *sZ, uint32_t *sX,
                                            uint32_t &tz, uint32_t &tx)
{
    const int kind = s.kind[loc];
    int zv, xv;
    if (kind == 0) { zv = 1; xv = 0; }                       // MeasX / PrepX: Z fault
    else if (kind == 1) { zv = 0; xv = 1; }                  // MeasZ / PrepZ: X fault
    else if (kind == 2) {                                    // IDLE: 0=X 1=Y else Z (noise/kernels.py:262-270)
        zv = (outcome != 0); xv = (outcome == 0 || outcome == 1);
    } else {                                                 // CNOT, outcome >= 14 -> ZX (the reference's else)
        const int o = outcome < 0 ? 14 : (outcome > 14 ? 14 : outcome);
        zv = c_cnot_zvar[o]; xv = c_cnot_xvar[o];
    }
    if (zv) {
        const int col = s.colZ[loc * 4 + zv];
        if (col >= 0) {
            for (int p = s.cpZ[col]; p < s.cpZ[col + 1]; ++p) { const int r = s.rowZ[p]; atomicXor(&sZ[r >> 5], 1u << (r & 31)); }
            tz ^= s.lmZ[col];
        }
    }
    if (xv) {
        const int col = s.colX[loc * 4 + xv];
        if (col >= 0) {
            for (int p = s.cpX[col]; p < s.cpX[col + 1]; ++p) { const int r = s.rowX[p]; atomicXor(&sX[r >> 5], 1u << (r & 31)); }
            tx ^= s.lmX[col];
        }
    }
}

__device__ __forceinline__ uint32_t warp_xor(uint32_t v)
{
    for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

constexpr int SAMP_WARPS = 8;
constexpr int GEO_LOG = 10, GEO_K = (1 << GEO_LOG) - 1;     // jump table entries 1 .. GEO_K (binary search of GEO_LOG steps)

// K2: explicit events
__global__ void __launch_bounds__(SAMP_WARPS * 32)
events_syndrome_kernel(SamplerDev s, const int32_t *ev_ptr, const uint32_t *events, int B,
                       uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sZ = sm + warp * (s.mwZ + s.mwX), *sX = sZ + s.mwZ;
    for (int shot = blockIdx.x * SAMP_WARPS + warp; shot < B; shot += gridDim.x * SAMP_WARPS) {
        for (int w = lane; w < s.mwZ + s.mwX; w += 32) sZ[w] = 0u;
        __syncwarp();
        uint32_t tz = 0u, tx = 0u;
        for (int e = ev_ptr[shot] + lane; e < ev_ptr[shot + 1]; e += 32) {
            const uint32_t ev = events[e];
            const int loc = ev & 0xFFFFFF;
            if (loc < s.L) apply_fault(s, loc, (int)(ev >> 24), sZ, sX, tz, tx);
        }
        __syncwarp();
        tz = warp_xor(tz); tx = warp_xor(tx);
        for (int w = lane; w < s.mwZ; w += 32) synZ[(size_t)shot * s.mwZ + w] = sZ[w];
        for (int w = lane; w < s.mwX; w += 32) synX[(size_t)shot * s.mwX + w] = sX[w];
        if (lane == 0) { trueZ[shot] = tz; trueX[shot] = tx; }
        __syncwarp();
    }
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// K1+K2 fused.  Faults are placed by sampling the GAPS between them instead of one Bernoulli draw per location (17 280
// draws for ~86 faults per gross-code shot): lane l of the shot's warp owns the locations [l*C, (l+1)*C), C = ceil(L/32), and
// walks them with geometric jumps, so that the 32 lanes apply their faults in lock step (~2.7 faults per lane at p = 0.005).
//   G ~ Geometric(p), P(G >= k) = (1-p)^k, exactly by inversion on a 32-bit word r against the table
//   T[k] = floor(2^32 (1-p)^k), k = 1..GEO_K:  G = #{k : r < T[k]};  r < T[GEO_K] means "at least GEO_K": the jump is extended
//   by another draw (memorylessness), so the distribution has no truncation error.  The table is built on the host
//   (qb_sampler_geometric_table returns the very words the kernel uses, for re-deriving a stream in the tests).
// Stream layout: Philox4x32-10, key = seed, counter = (shot_lo, shot_hi, lane + 32 * call, 2) for the call-th block of four
// words of a lane; a lane consumes its words in order: one per jump (more when extended), then, for a fault on an IDLE /
// CNOT location, one for the Pauli outcome floor(word * K / 2^32), K = 3 or 15 (src/noise/kernels.py:262-342).
__global__ void __launch_bounds__(SAMP_WARPS * 32)
sample_syndrome_kernel(SamplerDev s, uint64_t seed, uint64_t first_shot, int B, const uint32_t *geo,
                       uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX, int32_t *nfaults)
{
    extern __shared__ uint32_t sm[];
    uint32_t *T = sm;                                                  // [GEO_K + 1], T[0] unused
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sZ = sm + (GEO_K + 1) + warp * (s.mwZ + s.mwX), *sX = sZ + s.mwZ;
    for (int i = threadIdx.x; i <= GEO_K; i += blockDim.x) T[i] = geo ? geo[i] : 0u;
    __syncthreads();
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const int C = (s.L + 31) >> 5;
    const int lo = min(s.L, lane * C), hi = min(s.L, lo + C);
    for (int b = blockIdx.x * SAMP_WARPS + warp; b < B; b += gridDim.x * SAMP_WARPS) {
        const uint64_t shot = first_shot + (uint64_t)b;
        const uint32_t slo = (uint32_t)shot, shi = (uint32_t)(shot >> 32);
        for (int w = lane; w < s.mwZ + s.mwX; w += 32) sZ[w] = 0u;
        __syncwarp();
        uint32_t tz = 0u, tx = 0u;
        int nf = 0;
        uint4 buf = make_uint4(0u, 0u, 0u, 0u);
        int have = 0;
        uint32_t call = 0u;
        auto next_word = [&]() -> uint32_t {
            if (have == 0) { buf = philox4x32_10(make_uint4(slo, shi, (uint32_t)lane + 32u * call, 2u), key); ++call; have = 4; }
            const uint32_t r = buf.x;
            buf.x = buf.y; buf.y = buf.z; buf.z = buf.w; --have;
            return r;
        };
        int pos = lo;
        while (geo != nullptr) {
            // jump: number of fault-free locations before the next fault
            while (true) {
                const uint32_t r = next_word();
                int a = 0, c = GEO_K;                                  // largest k in [0, GEO_K] with r < T[k]  (T[0] = 2^32)
#pragma unroll
                for (int step = 0; step < GEO_LOG; ++step) { const int mid = (a + c + 1) >> 1; if (r < T[mid]) a = mid; else c = mid - 1; }
                pos += a;
                if (a < GEO_K || pos >= hi) break;
            }
            if (pos >= hi) break;
            const int loc = pos++;
            int outcome = 0;
            const int kind = s.kind[loc];
            if (kind >= 2) outcome = (int)__umulhi(next_word(), kind == 2 ? 3u : 15u);
            apply_fault(s, loc, outcome, sZ, sX, tz, tx);
            ++nf;
        }
        __syncwarp();
        tz = warp_xor(tz); tx = warp_xor(tx);
        for (int o = 16; o; o >>= 1) nf += __shfl_xor_sync(0xFFFFFFFFu, nf, o);
        for (int w = lane; w < s.mwZ; w += 32) synZ[(size_t)b * s.mwZ + w] = sZ[w];
        for (int w = lane; w < s.mwX; w += 32) synX[(size_t)b * s.mwX + w] = sX[w];
        if (lane == 0) { trueZ[b] = tz; trueX[b] = tx; if (nfaults) nfaults[b] = nf; }
        __syncwarp();
    }
}

static SamplerDev dev_view(const qb_sampler *s)
{
    SamplerDev d;
    d.L = s->L; d.k = s->k; d.mwZ = s->mwZ; d.mwX = s->mwX;
    d.kind = s->d_kind; d.colZ = s->d_colZ; d.colX = s->d_colX;
    d.cpZ = s->d_cpZ; d.rowZ = s->d_rowZ; d.cpX = s->d_cpX; d.rowX = s->d_rowX;
    d.lmZ = s->d_lmZ; d.lmX = s->d_lmX;
    return d;
}

int launch_events_syndrome(qb_sampler *s, const int32_t *ev_ptr_d, const uint32_t *events_d, int B,
                           uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    const int grid = std::max(1, std::min(ceil_div(B, SAMP_WARPS), s->sm_count * 8));
    const size_t smem = sizeof(uint32_t) * SAMP_WARPS * (s->mwZ + s->mwX);
    events_syndrome_kernel<<<grid, SAMP_WARPS * 32, smem, st>>>(dev_view(s), ev_ptr_d, events_d, B, synZ, trueZ, synX, trueX);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

// T[k] = floor(2^32 (1-p)^k) for k = 1 .. GEO_K (T[0] unused); p in (0, 1)
void geometric_table(double p, std::vector<uint32_t> &T)
{
    T.assign(GEO_K + 1, 0u);
    const double l1p = log1p(-p);
    for (int k = 1; k <= GEO_K; ++k) {
        const double v = floor(4294967296.0 * exp((double)k * l1p));
        T[k] = v >= 4294967295.0 ? 0xFFFFFFFFu : (v <= 0.0 ? 0u : (uint32_t)v);
    }
}
int geometric_table_size() { return GEO_K + 1; }

int launch_sample_syndrome(qb_sampler *s, uint64_t seed, uint64_t first_shot, int B, double p,
                           uint32_t *synZ, uint32_t *trueZ, uint32_t *synX, uint32_t *trueX,
                           int32_t *nfaults, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    QB_REQUIRE(p >= 0.0 && p < 1.0, "error_rate must be in [0, 1)");
    if (p > 0.0 && (s->geo_p != p || s->d_geo == nullptr)) {          // jump table of this error rate (rebuilt when p changes)
        geometric_table(p, s->h_geo);
        QB_CUDA(cudaStreamSynchronize(st));
        if (!s->d_geo) { QB_CUDA(cudaMalloc(reinterpret_cast<void **>(&s->d_geo), sizeof(uint32_t) * (GEO_K + 1))); s->owned.push_back(s->d_geo); }
        QB_CUDA(cudaMemcpy(s->d_geo, s->h_geo.data(), sizeof(uint32_t) * (GEO_K + 1), cudaMemcpyHostToDevice));
        s->geo_p = p;
    }
    const int grid = std::max(1, std::min(ceil_div(B, SAMP_WARPS), s->sm_count * 8));
    const size_t smem = sizeof(uint32_t) * ((GEO_K + 1) + SAMP_WARPS * (s->mwZ + s->mwX));
    sample_syndrome_kernel<<<grid, SAMP_WARPS * 32, smem, st>>>(dev_view(s), seed, first_shot, B, p > 0.0 ? s->d_geo : nullptr, synZ, trueZ,
                                                               synX, trueX, nfaults);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

// ---- small utility kernels ----------------------------------------------------------------------
__global__ void pack_bits_kernel(const int8_t *src, int B, int len, uint32_t *dst, int words)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * words) return;
    const int b = (int)(i / words), w = (int)(i % words);
    uint32_t v = 0u;
    const int8_t *row = src + (size_t)b * len;
    for (int t = 0; t < 32; ++t) { const int j = w * 32 + t; if (j < len && (row[j] & 1)) v |= 1u << t; }
    dst[i] = v;
}

__global__ void unpack_bits_kernel(const uint32_t *src, int B, int len, int words, int8_t *dst)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * len) return;
    const int b = (int)(i / len), j = (int)(i % len);
    dst[i] = (int8_t)((src[(size_t)b * words + (j >> 5)] >> (j & 31)) & 1u);
}

int launch_pack_bits(const int8_t *src, int B, int len, uint32_t *dst, int words, cudaStream_t st)
{
    const size_t total = (size_t)B * words;
    if (!total) return QB_OK;
    pack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, B, len, dst, words);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

int launch_unpack_bits(const uint32_t *src, int B, int len, int words, int8_t *dst, cudaStream_t st)
{
    const size_t total = (size_t)B * len;
    if (!total) return QB_OK;
    unpack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, B, len, words, dst);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

// K6: logical check  (HZ_logical @ det) % 2 != true  (engine.py:99-100, 119-120) + counters.
// One warp per shot; flags[shot] |= err << flag_bit; counts[count_slot] += number of errors.
__global__ void __launch_bounds__(256)
logical_check_kernel(GraphDev g, const uint32_t *hard_bits, const uint32_t *true_mask, int B, uint8_t *flags,
                     int flag_bit, unsigned long long *counts, int count_slot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __shared__ int s_err;
    if (threadIdx.x == 0) s_err = 0;
    __syncthreads();
    int my_err = 0;
    for (int shot = blockIdx.x * nwarps + warp; shot < B; shot += gridDim.x * nwarps) {
        uint32_t mask = 0u;
        for (int w = lane; w < g.nw; w += 32) {
            uint32_t bits = hard_bits[(size_t)shot * g.nw + w];
            while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; const int j = w * 32 + b; if (j < g.n) mask ^= g.logmask[j]; }
        }
        mask = warp_xor(mask);
        const int err = (mask != true_mask[shot]) ? 1 : 0;
        if (lane == 0) {
            if (flag_bit == 0) flags[shot] = (uint8_t)err;            // first side initialises the flag byte
            else flags[shot] = (uint8_t)(flags[shot] | (err << flag_bit));
            my_err += err;
        }
    }
    if (lane == 0 && my_err) atomicAdd(&s_err, my_err);
    __syncthreads();
    if (threadIdx.x == 0 && s_err) atomicAdd(&counts[count_slot], (unsigned long long)s_err);
}

int launch_logical_check(const qb_decoder *dec, const uint32_t *hard_bits, const uint32_t *true_mask, int B,
                         uint8_t *flags, int flag_bit, int64_t *counts, int count_slot, cudaStream_t st)
{
    if (B <= 0) return QB_OK;
    const int grid = std::max(1, std::min(ceil_div(B, 8), dec->sm_count * 8));
    logical_check_kernel<<<grid, 256, 0, st>>>(dec->g, hard_bits, true_mask, B, flags, flag_bit,
                                               reinterpret_cast<unsigned long long *>(counts), count_slot);
    QB_CUDA(cudaGetLastError());
    return QB_OK;
}

}  // namespace qb
