"""Bivariate-bicycle (BB) codes and their syndrome-extraction circuit.

Mirrors the interface of the reference ``src/codes/bb_code.py`` (``BBCodeCircuit``:4-197) and
the polynomial table of ``generate_codes.py:16-88``, but is built array-first: the circuit is
three int32 arrays (opcode, q1, q2) and the tuple list the reference exposes
(``cycle``, ``get_full_circuit()``) is a derived view.

Qubit linear order (reference ``bb_code.py:73-106``):
    [0,n2) X-check ancillas | [n2,2n2) left data | [2n2,3n2) right data | [3n2,4n2) Z-check ancillas
"""
import numpy as np

from ..utils.gf2 import css_logicals

# opcodes, same integer values as the reference (src/noise/constants.py:8-31)
OP_CNOT, OP_PREP_X, OP_PREP_Z, OP_MEAS_X, OP_MEAS_Z, OP_IDLE = 1, 2, 3, 4, 5, 6
_OP_NAME = {OP_CNOT: "CNOT", OP_PREP_X: "PrepX", OP_PREP_Z: "PrepZ",
            OP_MEAS_X: "MeasX", OP_MEAS_Z: "MeasZ", OP_IDLE: "IDLE"}

# A = x^a1 + y^a2 + y^a3 ; B = y^b1 + x^b2 + x^b3   (generate_codes.py:12-88)
BB_CODES = {
    "[[72, 12, 6]]": dict(ell=6, m=6, a_x_powers=[3], a_y_powers=[1, 2], b_y_powers=[3], b_x_powers=[1, 2], distance=6),
    "[[90, 8, 10]]": dict(ell=15, m=3, a_x_powers=[9], a_y_powers=[1, 2], b_y_powers=[0], b_x_powers=[2, 7], distance=10),
    "[[108, 8, 10]]": dict(ell=9, m=6, a_x_powers=[3], a_y_powers=[1, 2], b_y_powers=[3], b_x_powers=[1, 2], distance=10),
    "[[144, 12, 12]]": dict(ell=12, m=6, a_x_powers=[3], a_y_powers=[1, 2], b_y_powers=[3], b_x_powers=[1, 2], distance=12),
    "[[288, 12, 18]]": dict(ell=12, m=12, a_x_powers=[3], a_y_powers=[2, 7], b_y_powers=[3], b_x_powers=[1, 2], distance=18),
}

SCHEDULE_X = ("idle", 1, 4, 3, 5, 0, 2, "idle")   # bb_code.py:154
SCHEDULE_Z = (3, 5, 0, 1, 2, 4, "idle", "idle")   # bb_code.py:155


def _shift_targets(ell, m, kind, power):
    """Column index of the single 1 in each row of x^power (kind 'x') or y^power (kind 'y').

    x^p = kron(roll(I_ell, p, axis=1), I_m), y^p = kron(I_ell, roll(I_m, p, axis=1))
    (reference bb_code.py:50-63): row (i, j) has its 1 at ((i+p)%ell, j) resp. (i, (j+p)%m)."""
    i, j = np.divmod(np.arange(ell * m), m)
    if kind == "x":
        return ((i + power) % ell) * m + j
    return i * m + (j + power) % m


def bb_components(ell, m, a_x_powers, a_y_powers, b_y_powers, b_x_powers):
    """Per-component column maps: A_k[i] / B_k[i] = column of the 1 in row i (or -1)."""
    A = [_shift_targets(ell, m, "x", int(p)) for p in a_x_powers] + \
        [_shift_targets(ell, m, "y", int(p)) for p in a_y_powers]
    B = [_shift_targets(ell, m, "y", int(p)) for p in b_y_powers] + \
        [_shift_targets(ell, m, "x", int(p)) for p in b_x_powers]
    return A, B


def bb_parity_matrices(ell, m, a_x_powers, a_y_powers, b_y_powers, b_x_powers, **_):
    """Hx = [A | B], Hz = [B^T | A^T] (generate_codes.py:98-121)."""
    n2 = ell * m
    A_maps, B_maps = bb_components(ell, m, a_x_powers, a_y_powers, b_y_powers, b_x_powers)
    A = np.zeros((n2, n2), dtype=np.int64)
    B = np.zeros((n2, n2), dtype=np.int64)
    rows = np.arange(n2)
    for cm in A_maps:
        A[rows, cm] ^= 1
    for cm in B_maps:
        B[rows, cm] ^= 1
    return np.hstack([A, B]), np.hstack([B.T, A.T])


def make_bb_code(name):
    """Dict with the same keys as the reference ``codes/<name>.npz`` (generate_codes.py:154-168).

    Lx/Lz come from our own GF(2) nullspace computation (``qldpc`` is not available)."""
    spec = dict(BB_CODES[name])
    Hx, Hz = bb_parity_matrices(**spec)
    Lx, Lz = css_logicals(Hx, Hz)
    out = dict(Hx=Hx, Hz=Hz, Lx=Lx, Lz=Lz, distance=spec["distance"], ell=spec["ell"], m=spec["m"])
    for key in ("a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers"):
        out[key] = np.array(spec[key])
    return out


class BBCodeCircuit:
    """Syndrome-extraction circuit of a BB code (reference ``BBCodeCircuit``, bb_code.py:4-197).

    Public attributes kept from the reference: ``n``, ``n2``, ``num_cycles``, ``lin_order``,
    ``data_qubits``, ``Xchecks``, ``Zchecks``, ``total_qubits``, ``nbs``, ``cycle``,
    ``get_full_circuit()``, ``get_circuit_with_final_measurements()``.
    Array form (what the GPU tables are built from): ``cycle_ops``, ``cycle_q1``, ``cycle_q2``.
    """

    def __init__(self, Hx, Hz, num_cycles=12, ell=None, m=None, a_x_powers=None,
                 a_y_powers=None, b_y_powers=None, b_x_powers=None):
        self.Hx = np.asarray(Hx, dtype=int)
        self.Hz = np.asarray(Hz, dtype=int)
        self.num_cycles = int(num_cycles)
        self.m_checks, self.n = self.Hx.shape
        self.n2 = self.n // 2
        assert self.m_checks == self.n2, f"Expected square blocks: m={self.m_checks}, n2={self.n2}"
        self.ell = None if ell is None else int(ell)
        self.m_dim = None if m is None else int(m)
        self.a_x_powers = [] if a_x_powers is None else [int(p) for p in np.atleast_1d(a_x_powers)]
        self.a_y_powers = [] if a_y_powers is None else [int(p) for p in np.atleast_1d(a_y_powers)]
        self.b_y_powers = [] if b_y_powers is None else [int(p) for p in np.atleast_1d(b_y_powers)]
        self.b_x_powers = [] if b_x_powers is None else [int(p) for p in np.atleast_1d(b_x_powers)]
        self.has_component_params = self.ell is not None and self.m_dim is not None
        self.total_qubits = 4 * self.n2
        self._neighbor_table()
        self._cycle_arrays()

    # -- neighbours: nb[side][direction][check] = linear qubit index of the data qubit -------
    def _neighbor_table(self):
        n2 = self.n2
        nbX = np.zeros((6, n2), dtype=np.int64)
        nbZ = np.zeros((6, n2), dtype=np.int64)
        if self.has_component_params:
            A, B = bb_components(self.ell, self.m_dim, self.a_x_powers, self.a_y_powers,
                                 self.b_y_powers, self.b_x_powers)
            # missing components behave like zero matrices -> neighbour index 0 (bb_code.py:65-68,117-122)
            zero = np.zeros(n2, dtype=np.int64)
            inv = lambda cm: np.argsort(cm)          # transpose of a permutation matrix
            A3 = [A[k] if k < len(A) else None for k in range(3)]
            B3 = [B[k] if k < len(B) else None for k in range(3)]
            for k in range(3):
                nbX[k] = n2 + (A3[k] if A3[k] is not None else zero)          # X check -> left via A_k
                nbX[3 + k] = 2 * n2 + (B3[k] if B3[k] is not None else zero)  # X check -> right via B_k
                nbZ[k] = n2 + (inv(B3[k]) if B3[k] is not None else zero)     # Z check -> left via B_k^T
                nbZ[3 + k] = 2 * n2 + (inv(A3[k]) if A3[k] is not None else zero)
            self._nb_valid = np.ones((2, 6, n2), dtype=bool)
        else:
            # generic CSS fallback: first three left / right supports of each row (bb_code.py:138-160)
            self._nb_valid = np.zeros((2, 6, n2), dtype=bool)
            for s, H, nb in ((0, self.Hx, nbX), (1, self.Hz, nbZ)):
                for i in range(n2):
                    left = np.nonzero(H[i, :n2])[0][:3]
                    right = np.nonzero(H[i, n2:])[0][:3]
                    for j, idx in enumerate(left):
                        nb[j, i] = n2 + idx
                        self._nb_valid[s, j, i] = True
                    for j, idx in enumerate(right):
                        nb[3 + j, i] = 2 * n2 + idx
                        self._nb_valid[s, 3 + j, i] = True
        self.nbX, self.nbZ = nbX, nbZ

    def _cycle_arrays(self):
        n2 = self.n2
        xq = np.arange(n2)
        zq = 3 * n2 + np.arange(n2)
        data = n2 + np.arange(2 * n2)
        ops, q1, q2 = [], [], []

        def emit(op, a, b=None):
            a = np.asarray(a, dtype=np.int64)
            ops.append(np.full(a.shape, op, dtype=np.int64))
            q1.append(a)
            q2.append(np.full(a.shape, -1, dtype=np.int64) if b is None else np.asarray(b, dtype=np.int64))

        for t in range(8):
            busy = np.zeros(4 * n2, dtype=bool)
            if t == 0:
                emit(OP_PREP_X, xq)
            if SCHEDULE_X[t] != "idle":
                d = SCHEDULE_X[t]
                if not self._nb_valid[0, d].all():
                    raise KeyError((("Xcheck", int(np.nonzero(~self._nb_valid[0, d])[0][0])), d))
                emit(OP_CNOT, xq, self.nbX[d])
                busy[self.nbX[d]] = True
            if SCHEDULE_Z[t] != "idle":
                d = SCHEDULE_Z[t]
                if not self._nb_valid[1, d].all():
                    raise KeyError((("Zcheck", int(np.nonzero(~self._nb_valid[1, d])[0][0])), d))
                emit(OP_CNOT, self.nbZ[d], zq)
                busy[self.nbZ[d]] = True
            emit(OP_IDLE, data[~busy[data]])
            if t == 6:
                emit(OP_MEAS_Z, zq)
            if t == 7:
                emit(OP_MEAS_X, xq)
                emit(OP_PREP_Z, zq)
        self.cycle_ops = np.concatenate(ops).astype(np.int32)
        self.cycle_q1 = np.concatenate(q1).astype(np.int32)
        self.cycle_q2 = np.concatenate(q2).astype(np.int32)

    # -- reference-compatible tuple views ---------------------------------------------------
    def _node(self, q):
        n2 = self.n2
        q = int(q)
        if q < n2:
            return ("Xcheck", q)
        if q < 2 * n2:
            return ("data_left", q - n2)
        if q < 3 * n2:
            return ("data_right", q - 2 * n2)
        return ("Zcheck", q - 3 * n2)

    @property
    def Xchecks(self):
        return [("Xcheck", i) for i in range(self.n2)]

    @property
    def Zchecks(self):
        return [("Zcheck", i) for i in range(self.n2)]

    @property
    def data_qubits(self):
        return [("data_left", i) for i in range(self.n2)] + [("data_right", i) for i in range(self.n2)]

    @property
    def lin_order(self):
        return {self._node(q): q for q in range(self.total_qubits)}

    @property
    def nbs(self):
        out = {}
        for i in range(self.n2):
            for d in range(6):
                if self._nb_valid[0, d, i]:
                    out[(("Xcheck", i), d)] = self._node(self.nbX[d, i])
                if self._nb_valid[1, d, i]:
                    out[(("Zcheck", i), d)] = self._node(self.nbZ[d, i])
        return out

    @property
    def cycle(self):
        out = []
        for op, a, b in zip(self.cycle_ops, self.cycle_q1, self.cycle_q2):
            if op == OP_CNOT:
                out.append(("CNOT", self._node(a), self._node(b)))
            else:
                out.append((_OP_NAME[int(op)], self._node(a)))
        return out

    def get_full_circuit(self):
        return self.cycle * self.num_cycles

    def get_circuit_with_final_measurements(self):
        return self.cycle * self.num_cycles, self.cycle * 2
