"""Decoding matrices and fault -> column tables (reference ``src/noise/builder.py:69-176``).

The reference simulates every single fault with the pure-Python Pauli-frame simulator, one
circuit per fault.  Propagation is GF(2)-linear, so here all faults of one side are propagated
*at once*: every qubit carries a bit-vector with one bit per fault (uint64 words) and each
gate is one vectorised XOR.  The fault order, the signature de-duplication order (first
appearance = column order, builder.py:115-124) and the probability accumulation order
(``sum(probs[i] for i in ...)``) are those of the reference, so HdecZ/HdecX/channel_probs come
out bit-identical to the reference's ``matrix_cache`` files.

Besides the reference's dict this module produces what the GPU sampler needs and the
reference never materialises: the map  (fault location, Pauli outcome) -> (Z column, X column).
"""
import numpy as np

from ..codes.bb_code import (OP_CNOT, OP_IDLE, OP_MEAS_X, OP_MEAS_Z, OP_PREP_X, OP_PREP_Z)

# location kinds
KIND_Z_ONLY, KIND_X_ONLY, KIND_IDLE, KIND_CNOT = 0, 1, 2, 3

# two-qubit outcome order of the reference sampler (noise/kernels.py:283-342):
_TWO_QUBIT = ("XI", "YI", "ZI", "IX", "IY", "IZ", "XX", "YY", "ZZ", "XY", "YX", "YZ", "ZY", "XZ", "ZX")
# variant index = (component on q1) + 2*(component on q2); 0 = no component on this side
CNOT_ZVAR = np.array([(p[0] in "YZ") + 2 * (p[1] in "YZ") for p in _TWO_QUBIT], dtype=np.int32)
CNOT_XVAR = np.array([(p[0] in "XY") + 2 * (p[1] in "XY") for p in _TWO_QUBIT], dtype=np.int32)
IDLE_ZVAR = np.array([0, 1, 1], dtype=np.int32)   # outcome 0=X, 1=Y, 2=Z (noise/kernels.py:262-270)
IDLE_XVAR = np.array([1, 1, 0], dtype=np.int32)


class SideTables:
    """One decoding side ('Z': Z faults seen by X checks, 'X': X faults seen by Z checks)."""

    def __init__(self):
        self.m = self.k = self.n_cols = 0
        self.col_ptr = self.col_rows = self.col_logmask = None
        self.fault_loc = self.fault_variant = self.fault_col = self.fault_weight_kind = None

    def dense(self, with_logical=False):
        rows = self.m + (self.k if with_logical else 0)
        H = np.zeros((rows, self.n_cols), dtype=np.int64)
        cols = np.repeat(np.arange(self.n_cols), np.diff(self.col_ptr))
        H[self.col_rows, cols] = 1
        if with_logical:
            for b in range(self.k):
                H[self.m + b, :] = (self.col_logmask >> np.uint64(b)) & np.uint64(1)
        return H

    def channel_probs(self, error_rate):
        """Per-column probability, accumulated in fault order like builder.py:123."""
        unit = np.array([error_rate, error_rate * 2 / 3, error_rate * 4 / 15])
        acc = np.zeros(self.n_cols, dtype=np.float64)
        np.add.at(acc, self.fault_col, unit[self.fault_weight_kind])
        return acc


class FaultTables:
    def __init__(self):
        self.L = 0
        self.loc_kind = None
        self.loc_colZ = None     # int32 [L,4]; [:,0] = -1
        self.loc_colX = None
        self.Z = SideTables()
        self.X = SideTables()


def _propagate_side(side, ops, q1, q2, n_base, total_qubits, syn_positions, syn_ptrs,
                    data_idx, Lmat):
    """Bit-parallel single-fault propagation for one side.

    ops/q1/q2 is base circuit followed by the noiseless suffix; faults live on the first
    ``n_base`` gates.  Returns a SideTables plus the per-location variant -> column map."""
    if side == "Z":
        meas_op, prep_op, k_single = OP_MEAS_X, OP_PREP_X, KIND_Z_ONLY
    else:
        meas_op, prep_op, k_single = OP_MEAS_Z, OP_PREP_Z, KIND_X_ONLY
    bops = ops[:n_base]
    # ---- enumerate faults in the reference's order (builder.py:88-106 / 131-149) -------------
    is_meas = bops == meas_op
    is_prep = bops == prep_op
    is_idle = bops == OP_IDLE
    is_cnot = bops == OP_CNOT
    per_gate = np.where(is_cnot, 3, (is_meas | is_prep | is_idle).astype(np.int64))
    F = int(per_gate.sum())
    first = np.concatenate([[0], np.cumsum(per_gate)[:-1]])
    fault_loc = np.repeat(np.arange(n_base), per_gate)
    fault_variant = np.ones(F, dtype=np.int32)
    cn = np.nonzero(is_cnot)[0]
    fault_variant[first[cn] + 1] = 2
    fault_variant[first[cn] + 2] = 3
    wk = np.zeros(F, dtype=np.int64)              # 0: p, 1: 2p/3, 2: 4p/15
    wk[np.repeat(is_idle, per_gate)] = 1
    wk[np.repeat(is_cnot, per_gate)] = 2
    before = np.repeat(is_meas, per_gate)          # Meas faults act before the gate

    W = (F + 63) // 64
    S = np.zeros((total_qubits, W), dtype=np.uint64)
    n_meas = int((ops == meas_op).sum())
    hist = np.zeros((n_meas, W), dtype=np.uint64)
    fword = (np.arange(F) >> 6)
    fbit = np.uint64(1) << (np.arange(F) & 63).astype(np.uint64)

    def inject(f0, f1, g):
        a, b = int(q1[g]), int(q2[g])
        for f in range(f0, f1):
            v = fault_variant[f]
            if v & 1:
                S[a, fword[f]] ^= fbit[f]
            if v & 2:
                S[b, fword[f]] ^= fbit[f]

    syn = 0
    for g in range(len(ops)):
        op = ops[g]
        if g < n_base and per_gate[g]:
            f0, f1 = int(first[g]), int(first[g] + per_gate[g])
            if before[f0]:
                inject(f0, f1, g)
        if op == OP_CNOT:
            if side == "Z":
                S[q1[g]] ^= S[q2[g]]       # Z: target -> control  (noise/kernels.py:57-59)
            else:
                S[q2[g]] ^= S[q1[g]]       # X: control -> target  (noise/kernels.py:138-140)
        elif op == prep_op:
            S[q1[g]] = 0
        elif op == meas_op:
            hist[syn] = S[q1[g]]
            syn += 1
        if g < n_base and per_gate[g]:
            if not before[f0]:
                inject(f0, f1, g)

    # ---- detectors = XOR of consecutive measurements of one check (noise/kernels.py:357-380) ---
    det = hist.copy()
    for c in range(len(syn_ptrs) - 1):
        pos = syn_positions[syn_ptrs[c]:syn_ptrs[c + 1]]
        if len(pos) > 1:
            det[pos[1:]] ^= hist[pos[:-1]]
    k = Lmat.shape[0]
    logi = np.zeros((k, W), dtype=np.uint64)
    for b in range(k):
        for j in np.nonzero(np.asarray(Lmat[b]) & 1)[0]:
            logi[b] ^= S[data_idx[j]]
    sig = np.vstack([det, logi])                 # (m+k) x W

    # ---- signature per fault, de-duplicated in first-appearance order ------------------------
    # The signatures are sparse (a fault flips a handful of detectors), so they are handled as row lists: the set bits
    # of sig in (row, fault) order, regrouped per fault (rows ascending), grouped by (length, two independent 64-bit
    # hashes of the row list) and then compared list against list with the group's first fault, which makes the grouping
    # exact.  (Transposing the dense bit matrix instead took 0.6 of the 1.1 s of the gross code's table build.)
    R = sig.shape[0]
    m = det.shape[0]
    wr, ww = np.nonzero(sig)                                            # non-zero 64-fault words, row-major
    wbits = np.unpackbits(np.ascontiguousarray(sig[wr, ww]).view(np.uint8).reshape(-1, 8), axis=1, bitorder="little")
    wi, wj = np.nonzero(wbits)
    rr_all, ff_all = wr[wi], ww[wi] * 64 + wj                           # rows ascending, faults ascending per row
    del wbits
    by_fault = np.argsort(ff_all, kind="stable")
    rows_f = rr_all[by_fault].astype(np.int64)                          # rows of fault 0, of fault 1, ...: ascending per fault
    cnt = np.bincount(ff_all, minlength=F).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    hrng = np.random.default_rng(0x51D0E5)
    h1 = hrng.integers(1, 1 << 63, size=R, dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    h2 = hrng.integers(1, 1 << 63, size=R, dtype=np.int64).astype(np.uint64)
    key = np.zeros((F, 3), dtype=np.uint64)
    key[:, 0] = cnt.astype(np.uint64)
    nz = np.nonzero(cnt)[0]
    if len(nz):
        with np.errstate(over="ignore"):
            key[nz, 1] = np.add.reduceat(h1[rows_f] * (rows_f.astype(np.uint64) + np.uint64(3)), ptr[nz])
        key[nz, 2] = np.bitwise_xor.reduceat(h2[rows_f], ptr[nz])
    kv = np.ascontiguousarray(key).view(np.dtype((np.void, 24))).ravel()
    _, first_idx, inverse = np.unique(kv, return_index=True, return_inverse=True)
    inverse = inverse.ravel()
    rep_of_fault = first_idx[inverse]                                   # first fault with the same key
    if len(rows_f):                                                     # exactness: every list equals its representative's
        shift = np.repeat(ptr[rep_of_fault] - ptr[:-1], cnt)
        if not np.array_equal(rows_f[np.arange(len(rows_f)) + shift], rows_f):
            raise RuntimeError("fault signature hash collision")       # (2^-128 per pair; never seen)
    order = np.argsort(first_idx, kind="stable")
    col_of_unique = np.empty(len(order), dtype=np.int64)
    col_of_unique[order] = np.arange(len(order))
    fault_col = col_of_unique[inverse]
    n_cols = len(order)
    rep = first_idx[order]                        # representative fault of each column
    # rows of the columns: detector rows -> CSC (row-sorted within each column), logical rows -> bit mask
    ccnt = cnt[rep]
    src = np.repeat(ptr[rep], ccnt) + (np.arange(int(ccnt.sum())) - np.repeat(np.concatenate([[0], np.cumsum(ccnt)[:-1]]), ccnt))
    crow = rows_f[src]
    ccol = np.repeat(np.arange(n_cols), ccnt)
    is_det = crow < m
    k = Lmat.shape[0]
    st = SideTables()
    st.m, st.k, st.n_cols = m, k, n_cols
    st.col_ptr = np.concatenate([[0], np.cumsum(np.bincount(ccol[is_det], minlength=n_cols))]).astype(np.int64)
    st.col_rows = crow[is_det].astype(np.int64)   # row-sorted within each column
    lm = np.zeros(n_cols, dtype=np.uint64)
    np.bitwise_or.at(lm, ccol[~is_det], np.uint64(1) << (crow[~is_det] - m).astype(np.uint64))
    st.col_logmask = lm
    st.fault_loc, st.fault_variant, st.fault_col, st.fault_weight_kind = fault_loc, fault_variant, fault_col, wk
    loc_col = np.full((n_base, 4), -1, dtype=np.int32)
    loc_col[fault_loc, fault_variant] = fault_col
    return st, loc_col


def build_fault_tables(compiled, Lx, Lz):
    """All per-code tables (independent of the physical error rate)."""
    ops = np.concatenate([compiled.base_ops, compiled.suffix_ops])
    q1 = np.concatenate([compiled.base_q1, compiled.suffix_q1])
    q2 = np.concatenate([compiled.base_q2, compiled.suffix_q2])
    nb = len(compiled.base_ops)
    ft = FaultTables()
    ft.L = nb
    bops = compiled.base_ops
    kind = np.full(nb, -1, dtype=np.int32)
    kind[(bops == OP_MEAS_X) | (bops == OP_PREP_X)] = KIND_Z_ONLY
    kind[(bops == OP_MEAS_Z) | (bops == OP_PREP_Z)] = KIND_X_ONLY
    kind[bops == OP_IDLE] = KIND_IDLE
    kind[bops == OP_CNOT] = KIND_CNOT
    if (kind < 0).any():
        raise ValueError("base circuit contains gates that are not fault locations")
    ft.loc_kind = kind
    ft.Z, ft.loc_colZ = _propagate_side("Z", ops, q1, q2, nb, compiled.total_qubits,
                                        compiled.x_syn_positions, compiled.x_syn_ptrs,
                                        compiled.data_qubit_indices, np.asarray(Lx))
    ft.X, ft.loc_colX = _propagate_side("X", ops, q1, q2, nb, compiled.total_qubits,
                                        compiled.z_syn_positions, compiled.z_syn_ptrs,
                                        compiled.data_qubit_indices, np.asarray(Lz))
    return ft


def fault_tables_for(compiled, Lx, Lz):
    """Memoised :func:`build_fault_tables` (keyed on the logical operators' bytes)."""
    key = (np.asarray(Lx).astype(np.uint8).tobytes(), np.asarray(Lz).astype(np.uint8).tobytes())
    ft = compiled._fault_tables.get(key)
    if ft is None:
        ft = build_fault_tables(compiled, Lx, Lz)
        compiled._fault_tables[key] = ft
    return ft


def matrices_from_tables(ft, error_rate, num_cycles):
    """The reference's matrices dict (builder.py:165-176), dense int64 like the cache files."""
    return {
        "HdecZ": ft.Z.dense(False), "HdecX": ft.X.dense(False),
        "channel_probsZ": ft.Z.channel_probs(error_rate),
        "channel_probsX": ft.X.channel_probs(error_rate),
        "HZ_full": ft.Z.dense(True), "HX_full": ft.X.dense(True),
        "first_logical_rowZ": ft.Z.m, "first_logical_rowX": ft.X.m,
        "num_cycles": num_cycles, "k": ft.Z.k,
    }


def build_decoding_matrices(circuit_builder, Lx, Lz, error_rate, verbose=True, num_workers=None):
    """Drop-in for reference ``build_decoding_matrices`` (builder.py:69-76); ``num_workers`` is
    accepted and ignored (no process pool is needed)."""
    from .compiled import CompiledCircuit
    if hasattr(circuit_builder, "cycle_ops"):
        compiled = CompiledCircuit.from_builder(circuit_builder)
    else:
        compiled = CompiledCircuit(circuit_builder.get_full_circuit(), circuit_builder.cycle * 2,
                                   circuit_builder.lin_order, circuit_builder.data_qubits,
                                   circuit_builder.Xchecks, circuit_builder.Zchecks)
    if verbose:
        print("Building Z/X decoding matrices (bit-parallel fault propagation)...")
    ft = fault_tables_for(compiled, Lx, Lz)
    return matrices_from_tables(ft, error_rate, circuit_builder.num_cycles)
