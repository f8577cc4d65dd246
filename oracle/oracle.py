"""Python face of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Same call signatures as the reference's entry points, implemented on
top of oracle/qldpc_oracle.c (float64, single thread).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

Pinning status: PINNED -- checked against the real reference (numba, imported from
/root/reference in the build container) by tests/golden/make_golden.py, whose outputs are the
committed fixtures tests/golden/*.npz that tests/test_oracle_golden.py replays everywhere.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_minsum_decode.restype = C.c_int
        _lib.orc_bp_decode.restype = C.c_int
        _lib.orc_gf2_elimination.restype = C.c_int
        _lib.orc_gf2_elimination_packed_core.restype = C.c_int
        _lib.orc_osd0.restype = C.c_int
        _lib.orc_decode_side.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _alpha_args(alpha, alpha_mode):
    """Validation and dispatch of sparse.py:18-39 / dense.py:19-33."""
    seq = np.zeros(1, dtype=np.float64)
    if alpha_mode is None:
        mode = 1 if alpha == 0 else 0
    elif alpha_mode == "dynamical":
        mode = 1
    elif alpha_mode == "alvarado":
        if alpha <= 0:
            raise ValueError("alpha must be > 0 when alpha_mode='alvarado'")
        mode = 0
    elif alpha_mode == "alvarado-autoregressive":
        seq = np.ascontiguousarray(alpha, dtype=np.float64)
        if seq.ndim != 1 or seq.size == 0:
            raise ValueError("alpha must be a non-empty 1D sequence for alvarado-autoregressive")
        mode = 2
    else:
        raise ValueError(f"Unsupported alpha_mode: {alpha_mode}")
    val = float(alpha) if mode != 2 else 0.0
    return mode, val, seq


def minsum_csr(H_indices, H_indptr, syndrome, prior, maxIter, mode, alpha_val, alpha_seq,
               damping=1.0, clip_llr=20.0, dense_variant=False):
    H_indices, H_indptr = _i32(H_indices), _i32(H_indptr)
    syndrome = np.ascontiguousarray(syndrome, dtype=np.int8)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    m, n = len(H_indptr) - 1, len(prior)
    hard = np.zeros(n, dtype=np.int8)
    values = np.zeros(n, dtype=np.float64)
    fin = C.c_int(0)
    conv = lib().orc_minsum_decode(_p(H_indices), _p(H_indptr), m, n, _p(syndrome), _p(prior),
                                   int(maxIter), int(mode), C.c_double(alpha_val), _p(alpha_seq),
                                   int(len(alpha_seq)), C.c_double(damping), C.c_double(clip_llr),
                                   int(dense_variant), _p(hard), _p(values), C.byref(fin))
    return hard, bool(conv), values, int(fin.value)


def performMinSum_Symmetric_Sparse(H_csr, syndrome, initialBelief, maxIter=100, alpha=1.0,
                                   alpha_mode="dynamical", damping=1.0, clip_llr=20.0):
    """Oracle for reference src/decoding/sparse.py:5-55."""
    mode, val, seq = _alpha_args(alpha, alpha_mode)
    return minsum_csr(H_csr.indices, H_csr.indptr, syndrome, initialBelief, maxIter, mode, val, seq,
                      damping, clip_llr)


def _dense_to_csr(H):
    H = np.asarray(H)
    mask = H != 0
    indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    indices = np.nonzero(mask)[1].astype(np.int32)
    return indices, indptr


def performMinSum_Symmetric(H, syndrome, initialBelief, maxIter=50, alpha=1.0, alpha_mode="dynamical",
                            damping=1.0, clip_llr=20.0, alpha_estimation=False):
    """Oracle for reference src/decoding/dense.py:5-73 (same arithmetic as the sparse decoder on
    the edges of H; the alpha_estimation early return of dense.py:54-56 is restated here)."""
    mode, val, seq = _alpha_args(alpha, alpha_mode)
    indices, indptr = _dense_to_csr(H)
    syndrome = np.asarray(syndrome, dtype=np.int8)
    prior = np.asarray(initialBelief, dtype=np.float64)
    m, n = np.asarray(H).shape
    if alpha_estimation:
        if maxIter < 1:
            return np.zeros(n, dtype=np.int8), False, None, -1
        a0 = 0.5 if mode == 1 else (float(seq[0]) if mode == 2 else val)   # alpha of iteration 0
        Q = np.ascontiguousarray(prior[indices])
        ss = np.ascontiguousarray(1.0 - 2.0 * syndrome.astype(np.float64))
        R = np.zeros(len(indices)); Rs = np.zeros(n)
        lib().orc_minsum_core_sparse(_p(indices), _p(indptr), m, n, _p(Q), _p(ss), C.c_double(a0), _p(R), _p(Rs))
        dense = np.zeros((m, n))
        rows = np.repeat(np.arange(m), np.diff(indptr))
        dense[rows, indices] = R
        return np.zeros(n, dtype=np.int8), False, dense / (a0 if a0 != 0 else 1.0), 0
    return minsum_csr(indices, indptr, syndrome, prior, maxIter, mode, val, seq, damping, clip_llr,
                      dense_variant=True)


def minsum_core_sparse(H_data, H_indices, H_indptr, Q_flat, syndrome_sign, alpha, m, n):
    """Oracle for reference src/decoding/kernels.py:139-169."""
    H_indices, H_indptr = _i32(H_indices), _i32(H_indptr)
    Q = np.ascontiguousarray(Q_flat, dtype=np.float64)
    ss = np.ascontiguousarray(syndrome_sign, dtype=np.float64)
    R = np.zeros(len(H_indices)); Rs = np.zeros(n)
    lib().orc_minsum_core_sparse(_p(H_indices), _p(H_indptr), int(m), int(n), _p(Q), _p(ss),
                                 C.c_double(alpha), _p(R), _p(Rs))
    return R, Rs


def performBeliefPropagationFast(H, syndrome, initialBelief, maxIter=50):
    """Oracle for reference src/decoding/dense.py:75-96."""
    indices, indptr = _dense_to_csr(H)
    syndrome = np.ascontiguousarray(syndrome, dtype=np.int8)
    prior = np.ascontiguousarray(initialBelief, dtype=np.float64)
    m, n = np.asarray(H).shape
    hard = np.zeros(n, dtype=np.int8); values = np.zeros(n); fin = C.c_int(0)
    conv = lib().orc_bp_decode(_p(indices), _p(indptr), m, n, _p(syndrome), _p(prior), int(maxIter),
                               _p(hard), _p(values), C.byref(fin))
    return hard, bool(conv), values, int(fin.value)


def gf2_elimination(A, b):
    """Oracle for reference src/decoding/kernels.py:6-34 (mutates A and b in place like it)."""
    assert A.dtype == np.int64 and b.dtype == np.int64 and A.flags.c_contiguous
    m, n = A.shape
    pr = np.zeros(min(m, n) + 1, dtype=np.int64); pc = np.zeros(min(m, n) + 1, dtype=np.int64)
    r = lib().orc_gf2_elimination(_p(A), _p(b), m, n, _p(pr), _p(pc))
    return A, b, pr[:r], pc[:r]


def pack_rows_uint64(A):
    """Restates src/decoding/kernels.py:36-46."""
    m, n = A.shape
    by = np.packbits(np.ascontiguousarray(A, dtype=np.uint8), axis=1, bitorder="little")
    pad = (-by.shape[1]) % 8
    if pad:
        by = np.pad(by, ((0, 0), (0, pad)))
    return np.ascontiguousarray(by).view(np.uint64), n


def gf2_elimination_packed(A, b):
    """Oracle for reference src/decoding/kernels.py:98-106."""
    Ap, n = pack_rows_uint64(A)
    Ap = np.ascontiguousarray(Ap)
    m, nw = Ap.shape
    assert b.dtype == np.int64
    pr = np.zeros(min(m, n) + 1, dtype=np.int64); pc = np.zeros(min(m, n) + 1, dtype=np.int64)
    r = lib().orc_gf2_elimination_packed_core(_p(Ap), _p(b), m, nw, n, _p(pr), _p(pc))
    return Ap, b, pr[:r], pc[:r]


def syndrome_check(H_data, H_indices, H_indptr, candidate, m):
    """Oracle for reference src/decoding/kernels.py:223-231."""
    out = np.zeros(m, dtype=np.int8)
    for i in range(m):
        s = 0
        for idx in range(H_indptr[i], H_indptr[i + 1]):
            s ^= int(candidate[H_indices[idx]])
        out[i] = s
    return out


def _csc(H):
    H = np.asarray(H)
    mask = (H != 0).T
    col_ptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    row_idx = np.nonzero(mask)[1].astype(np.int32)
    return col_ptr, row_idx


def osd0_csc(col_ptr, row_idx, m, n, syndrome, hard, ordering):
    syndrome = np.ascontiguousarray(syndrome, dtype=np.int8)
    hard = np.ascontiguousarray(hard, dtype=np.int8)
    ordering = np.ascontiguousarray(ordering, dtype=np.int64)
    sol = np.zeros(n, dtype=np.int64)
    piv = np.zeros(min(m, n) + 1, dtype=np.int64)
    r = lib().orc_osd0(_p(_i32(col_ptr)), _p(_i32(row_idx)), int(m), int(n), _p(syndrome), _p(hard),
                       _p(ordering), _p(sol), _p(piv))
    return sol, piv[:r]


def performOSD_enhanced(H, syndrome, llr, hard, order=0, max_combinations=None, ordering=None):
    """Oracle for reference src/decoding/osd.py:5-29 (OSD-0; the order>0 search of :31-77 is
    unreachable when the syndrome lies in the column space of H).
    ``ordering`` overrides ``np.argsort(|llr|)`` (same call as osd.py:12)."""
    H = np.asarray(H)
    m, n = H.shape
    if ordering is None:
        ordering = np.argsort(np.abs(llr))
    col_ptr, row_idx = _csc(H)
    sol, _ = osd0_csc(col_ptr, row_idx, m, n, syndrome, np.asarray(hard) & 1, ordering)
    osd_syn = (sol @ H.T) % 2
    if order != 0 and not np.all(osd_syn == np.asarray(syndrome)):
        raise NotImplementedError("oracle covers OSD-0 and consistent syndromes only")
    return sol


def run_trial_arrays(compiled, error_rate, Lx, Lz, random_vals, random_paulis, random_two_qubit):
    """Oracle for src/noise/simulation.py:21-107 given the three random arrays of :43-45."""
    c = compiled
    rv = np.ascontiguousarray(random_vals, dtype=np.float64)
    rp = np.ascontiguousarray(random_paulis, dtype=np.int32)
    r2 = np.ascontiguousarray(random_two_qubit, dtype=np.int32)
    Lx8 = np.ascontiguousarray(np.asarray(Lx) & 1, dtype=np.uint8)
    Lz8 = np.ascontiguousarray(np.asarray(Lz) & 1, dtype=np.uint8)
    k, nd = Lx8.shape
    mx, mz = c.num_meas_x, c.num_meas_z
    sz = np.zeros(mx, dtype=np.int8); sx = np.zeros(mz, dtype=np.int8)
    tz = np.zeros(k, dtype=np.int8); tx = np.zeros(k, dtype=np.int8)
    lib().orc_run_trial(_p(c.base_ops), _p(c.base_q1), _p(c.base_q2), len(c.base_ops),
                        _p(c.suffix_ops), _p(c.suffix_q1), _p(c.suffix_q2), len(c.suffix_ops),
                        c.total_qubits, C.c_double(error_rate), _p(rv), _p(rp), _p(r2),
                        _p(c.x_syn_positions), _p(c.x_syn_ptrs), c.num_x_checks,
                        _p(c.z_syn_positions), _p(c.z_syn_ptrs), c.num_z_checks,
                        _p(c.data_qubit_indices), nd, _p(Lx8), _p(Lz8), k,
                        _p(sz), _p(tz), _p(sx), _p(tx))
    return sz, tz, sx, tx


def run_trial_fast(compiled, error_rate, Lx, Lz):
    """Oracle for src/noise/simulation.py:21-107 including the legacy-RNG draws of :43-45."""
    n = compiled.num_error_locs
    rv = np.random.random(n)
    rp = np.random.randint(0, 3, n, dtype=np.int32)
    r2 = np.random.randint(0, 15, n, dtype=np.int32)
    return run_trial_arrays(compiled, error_rate, Lx, Lz, rv, rp, r2)


class SideGraph:
    """CSR/CSC/logical-row arrays of one decoding side for :func:`decode_side`."""

    def __init__(self, Hdec, H_logical, prior):
        Hdec = np.asarray(Hdec)
        self.m, self.n = Hdec.shape
        self.indices, self.indptr = _dense_to_csr(Hdec)
        self.col_ptr, self.row_idx = _csc(Hdec)
        self.lidx, self.lptr = _dense_to_csr(H_logical)
        self.k = np.asarray(H_logical).shape[0]
        self.prior = np.ascontiguousarray(prior, dtype=np.float64)


def decode_side(g, syndrome, true_l, maxIter, mode=1, alpha_val=1.0, alpha_seq=None):
    """BP + OSD-0 + logical comparison of one side (engine.py:83-100), stable-tie ordering.
    Returns (logical_error, converged, iterations)."""
    syndrome = np.ascontiguousarray(syndrome, dtype=np.int8)
    true_l = np.ascontiguousarray(true_l, dtype=np.int8)
    seq = np.zeros(1) if alpha_seq is None else np.ascontiguousarray(alpha_seq, dtype=np.float64)
    stats = np.zeros(2, dtype=np.int32)
    err = lib().orc_decode_side(_p(g.indices), _p(g.indptr), _p(g.col_ptr), _p(g.row_idx), g.m, g.n,
                                _p(g.lptr), _p(g.lidx), g.k, _p(syndrome), _p(true_l), _p(g.prior),
                                int(maxIter), int(mode), C.c_double(alpha_val), _p(seq), len(seq), _p(stats))
    return bool(err), bool(stats[0]), int(stats[1])


def llr_priors(channel_probs):
    """engine.py:210-212."""
    with np.errstate(divide="ignore", invalid="ignore"):
        cp = np.asarray(channel_probs, dtype=np.float64)
        return np.clip(np.nan_to_num(np.log((1 - cp) / cp)), -50, 50)


def alpha_messages(H_csr, syndromes, prior, alpha_prev, damping=1.0, clip_llr=20.0):
    """Oracle for the message collection of src/decoding/alpha.py:206-253 (autoregressive; with an empty
    ``alpha_prev`` it is the single-iteration case :127-137): advance with the given alphas, no convergence
    stop, then the unscaled check pass.  Returns R [B, nnz]."""
    indices, indptr = _i32(H_csr.indices), _i32(H_csr.indptr)
    m, n = H_csr.shape
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    out = []
    for syn in np.asarray(syndromes).reshape(-1, m):
        ss = np.ascontiguousarray(1.0 - 2.0 * syn.astype(np.float64))
        Q = np.ascontiguousarray(prior[indices])
        Qold = Q.copy()
        for a in alpha_prev:
            R, Rs = minsum_core_sparse(None, indices, indptr, Q, ss, float(a), m, n)
            q = (Rs + prior)[indices] - R
            q = np.where(np.isnan(q), 0.0, np.clip(q, -clip_llr, clip_llr))
            Q = np.clip(damping * q + (1.0 - damping) * Qold, -clip_llr, clip_llr)
            Qold = Q.copy()
        R, _ = minsum_core_sparse(None, indices, indptr, Q, ss, 1.0, m, n)
        out.append(R)
    return np.array(out)
