"""Kernel-level entry points with the reference's names and signatures
(reference ``src/decoding/kernels.py``); every one of them runs on the GPU through the C ABI.

reference function                      -> C entry point (include/qldpc_b200.h)
  minsum_decoder_full (:235-366)        -> qb_minsum_decode_host
  minsum_decoder_full_autoregressive    -> qb_minsum_decode_host (QB_ALPHA_SEQUENCE)
  minsum_core_sparse (:139-169)         -> qb_minsum_core_host
  gf2_elimination (:6-34)               -> qb_gf2_eliminate_host
  gf2_elimination_packed (:98-106)      -> qb_gf2_eliminate_host (+ uint64 packing of :36-46)
  syndrome_check (:223-231)             -> qb_syndrome_check_host
"""
import ctypes as C

import numpy as np

from .. import _lib


def _decoder(H_indices, H_indptr, n, prior):
    return _lib.cached_decoder_for(H_indptr, H_indices, n, prior)


def minsum_decoder_full(H_indices, H_indptr, syndrome, initialBelief, maxIter, use_dynamic_alpha, alpha_val,
                        damping, clip_llr):
    prior = np.asarray(initialBelief, dtype=np.float64)
    dec = _decoder(H_indices, H_indptr, len(prior), prior)
    mode = _lib.QB_ALPHA_DYNAMIC if use_dynamic_alpha else _lib.QB_ALPHA_FIXED
    hard, conv, values, fin = dec.minsum(np.asarray(syndrome, dtype=np.int8)[None, :], maxIter, mode, alpha=alpha_val,
                                         damping=damping, clip_llr=clip_llr)
    return hard[0], bool(conv[0]), values[0], int(fin[0])


def minsum_decoder_full_autoregressive(H_indices, H_indptr, syndrome, initialBelief, maxIter, alpha_seq, alpha_len,
                                       damping, clip_llr):
    prior = np.asarray(initialBelief, dtype=np.float64)
    dec = _decoder(H_indices, H_indptr, len(prior), prior)
    seq = np.asarray(alpha_seq, dtype=np.float64)[:alpha_len]
    hard, conv, values, fin = dec.minsum(np.asarray(syndrome, dtype=np.int8)[None, :], maxIter, _lib.QB_ALPHA_SEQUENCE,
                                         alpha_seq=seq, damping=damping, clip_llr=clip_llr)
    return hard[0], bool(conv[0]), values[0], int(fin[0])


def minsum_core_sparse(H_data, H_indices, H_indptr, Q_flat, syndrome_sign, alpha, m, n):
    dec = _decoder(H_indices, H_indptr, n, None)
    R, Rs = dec.minsum_core(np.asarray(Q_flat, dtype=np.float64)[None, :],
                            np.asarray(syndrome_sign, dtype=np.float64)[None, :], float(alpha))
    return R[0], Rs[0]


def syndrome_check(H_data, H_indices, H_indptr, candidate, m):
    n = len(candidate)
    dec = _decoder(H_indices, H_indptr, n, None)
    return dec.syndrome_check(np.asarray(candidate, dtype=np.int8)[None, :])[0]


def _pack_rows_uint64(A):
    """Rows packed little-endian into uint64 words, bit c of word c>>6 = column c (:36-46)."""
    m, n = A.shape
    by = np.packbits(np.ascontiguousarray(A, dtype=np.uint8), axis=1, bitorder="little")
    pad = (-by.shape[1]) % 8
    if pad:
        by = np.pad(by, ((0, 0), (0, pad)), mode="constant")
    return np.ascontiguousarray(by).view(np.uint64), n


def _eliminate(A64, b64, want_packed):
    lib = _lib.require_gpu()
    m, n = A64.shape
    mn = max(1, min(m, n))
    pr = np.zeros(mn, dtype=np.int64); pc = np.zeros(mn, dtype=np.int64)
    npv = C.c_int32(0)
    packed = np.zeros((m, max(1, (n + 63) // 64)), dtype=np.uint64) if want_packed else None
    _lib.check(lib.qb_gf2_eliminate_host(_lib.default_device(), _lib.ptr(A64), _lib.ptr(b64), m, n, _lib.ptr(packed),
                                         _lib.ptr(pr), _lib.ptr(pc), C.byref(npv)))
    return packed, pr[:npv.value], pc[:npv.value]


def gf2_elimination(A, b):
    """In-place GF(2) Gauss-Jordan; mutates ``A`` and ``b`` like the reference (:23-32)."""
    A64 = np.ascontiguousarray(A, dtype=np.int64)
    b64 = np.ascontiguousarray(b, dtype=np.int64)
    _, pr, pc = _eliminate(A64, b64, False)
    if A64 is not A:
        A[...] = A64
    if b64 is not b:
        b[...] = b64
    return A, b, pr, pc


def gf2_elimination_packed(A, b):
    """Packed variant: returns (A_packed uint64, b, pivot_rows, pivot_cols); ``b`` is reduced in place."""
    A64 = np.array(A, dtype=np.int64, order="C")
    b64 = np.ascontiguousarray(b, dtype=np.int64)
    packed, pr, pc = _eliminate(A64, b64, True)
    if b64 is not b:
        b[...] = b64
    return packed, b, pr, pc


def gf2_elimination_packed_core(A_packed, b, n):
    """Same sweep on an already packed matrix (:49-96); mutates ``A_packed`` and ``b``."""
    m, nw = A_packed.shape
    bits = np.unpackbits(np.ascontiguousarray(A_packed).view(np.uint8), axis=1, bitorder="little")[:, :n]
    A64 = bits.astype(np.int64)
    b64 = np.ascontiguousarray(b, dtype=np.int64)
    packed, pr, pc = _eliminate(A64, b64, True)
    A_packed[...] = packed[:, :nw]
    if b64 is not b:
        b[...] = b64
    return A_packed, b, pr, pc


def compute_metric(solution, llr_abs, syndrome_weight):
    """Host-side OSD metric (:196-204); only used by the order>0 search, which the simulated path never reaches."""
    metric = 1e10 + syndrome_weight * 1e8 if syndrome_weight > 0 else 0.0
    return metric + float(np.dot(np.asarray(solution, dtype=np.float64), llr_abs))


def recompute_solution(H_permuted, s_reduced, e_permuted, pivot_rows, pivot_cols):
    """Back-substitution of the order>0 search (:207-220), host side."""
    e_full = np.array(e_permuted, dtype=np.int64)
    for r, c in zip(pivot_rows, pivot_cols):
        row = np.asarray(H_permuted[r]) == 1
        contrib = int(np.bitwise_xor.reduce(e_full[row])) if row.any() else 0
        if row[c]:
            contrib ^= int(e_full[c])
        e_full[c] = int(s_reduced[r]) ^ contrib
    return e_full
