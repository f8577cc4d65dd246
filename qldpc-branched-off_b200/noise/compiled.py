"""Array form of the syndrome-extraction circuit (reference ``src/noise/compiled.py:10-173``).

``CompiledCircuit`` keeps the reference's constructor signature and attribute names so that
``run_trial_fast(compiled, p, Lx, Lz)`` is a drop-in; ``from_builder`` is the fast path that
skips the tuple lists and takes the arrays straight from :class:`BBCodeCircuit`.
"""
import numpy as np

from ..codes.bb_code import (OP_CNOT, OP_IDLE, OP_MEAS_X, OP_MEAS_Z, OP_PREP_X, OP_PREP_Z)

# same integers as the reference's src/noise/constants.py:8-51
GATE_TO_OPCODE = {
    "CNOT": 1, "PrepX": 2, "PrepZ": 3, "MeasX": 4, "MeasZ": 5, "IDLE": 6,
    "X": 10, "Y": 11, "Z": 12,
    "XX": 20, "XY": 21, "XZ": 22, "YX": 23, "YY": 24, "YZ": 25, "ZX": 26, "ZY": 27, "ZZ": 28,
}
_LOCATION_OPS = (OP_CNOT, OP_PREP_X, OP_PREP_Z, OP_MEAS_X, OP_MEAS_Z, OP_IDLE)


def circuit_to_arrays(circuit, lin_order):
    """Tuple circuit -> (ops, q1, q2) int32 arrays (reference compiled.py:10-41)."""
    n = len(circuit)
    ops = np.zeros(n, dtype=np.int32)
    q1 = np.full(n, -1, dtype=np.int32)
    q2 = np.full(n, -1, dtype=np.int32)
    for i, gate in enumerate(circuit):
        ops[i] = GATE_TO_OPCODE.get(gate[0], 0)
        if len(gate) >= 2 and gate[1] is not None:
            q1[i] = lin_order[gate[1]]
        if len(gate) >= 3 and gate[2] is not None:
            q2[i] = lin_order[gate[2]]
    return ops, q1, q2


def _syndrome_positions(ops, q1, meas_op, check_qubits):
    """CSR (positions, ptrs): for each check, the indices of its measurements in record order
    (reference ``build_syndrome_map_arrays`` compiled.py:72-103)."""
    meas_q = q1[ops == meas_op]
    slot = {int(q): i for i, q in enumerate(check_qubits)}
    buckets = [[] for _ in check_qubits]
    for pos, q in enumerate(meas_q):
        b = slot.get(int(q))
        if b is not None:
            buckets[b].append(pos)
    ptrs = np.zeros(len(check_qubits) + 1, dtype=np.int32)
    ptrs[1:] = np.cumsum([len(b) for b in buckets])
    flat = np.array([p for b in buckets for p in b], dtype=np.int32)
    return flat, ptrs


def count_error_locations(circuit):
    """Number of fault locations = every gate of the six circuit kinds (compiled.py:106-113)."""
    return sum(1 for g in circuit if g[0] in ("MeasX", "MeasZ", "PrepX", "PrepZ", "IDLE", "CNOT"))


class CompiledCircuit:
    def __init__(self, base_circuit, noiseless_suffix, lin_order, data_qubits, Xchecks, Zchecks):
        b = circuit_to_arrays(base_circuit, lin_order)
        s = circuit_to_arrays(noiseless_suffix, lin_order)
        self._finish(b, s, len(lin_order),
                     np.array([lin_order[q] for q in data_qubits], dtype=np.int32),
                     np.array([lin_order[q] for q in Xchecks], dtype=np.int32),
                     np.array([lin_order[q] for q in Zchecks], dtype=np.int32))
        self.lin_order = lin_order

    @classmethod
    def from_builder(cls, cb, suffix_cycles=2):
        """Fast constructor from a :class:`BBCodeCircuit` (same content as
        ``CompiledCircuit(cb.get_full_circuit(), cb.cycle*2, ...)``, reference engine.py:394-401)."""
        self = cls.__new__(cls)
        rep = lambda a, r: np.tile(a, r).astype(np.int32)
        base = tuple(rep(a, cb.num_cycles) for a in (cb.cycle_ops, cb.cycle_q1, cb.cycle_q2))
        suf = tuple(rep(a, suffix_cycles) for a in (cb.cycle_ops, cb.cycle_q1, cb.cycle_q2))
        n2 = cb.n2
        self._finish(base, suf, cb.total_qubits,
                     (n2 + np.arange(2 * n2)).astype(np.int32),
                     np.arange(n2, dtype=np.int32),
                     (3 * n2 + np.arange(n2)).astype(np.int32))
        self.lin_order = None
        return self

    def _finish(self, base, suffix, total_qubits, data_idx, x_idx, z_idx):
        self.total_qubits = int(total_qubits)
        self.base_ops, self.base_q1, self.base_q2 = base
        self.suffix_ops, self.suffix_q1, self.suffix_q2 = suffix
        full_ops = np.concatenate([self.base_ops, self.suffix_ops])
        full_q1 = np.concatenate([self.base_q1, self.suffix_q1])
        self.x_syn_positions, self.x_syn_ptrs = _syndrome_positions(full_ops, full_q1, OP_MEAS_X, x_idx)
        self.z_syn_positions, self.z_syn_ptrs = _syndrome_positions(full_ops, full_q1, OP_MEAS_Z, z_idx)
        self.x_check_indices = x_idx.copy()
        self.x_check_ptrs = np.arange(len(x_idx) + 1, dtype=np.int32)
        self.z_check_indices = z_idx.copy()
        self.z_check_ptrs = np.arange(len(z_idx) + 1, dtype=np.int32)
        self.data_qubit_indices = data_idx
        self.num_error_locs = int(np.isin(self.base_ops, _LOCATION_OPS).sum())
        self.max_circuit_size = len(self.base_ops) + len(self.suffix_ops) + self.num_error_locs
        self.num_meas_x = int((full_ops == OP_MEAS_X).sum())
        self.num_meas_z = int((full_ops == OP_MEAS_Z).sum())
        self.max_syndromes_x = self.num_meas_x + 100
        self.max_syndromes_z = self.num_meas_z + 100
        self.num_x_checks = len(x_idx)
        self.num_z_checks = len(z_idx)
        self._fault_tables = {}      # cache: id(Lx,Lz) -> FaultTables (see noise/builder.py)
