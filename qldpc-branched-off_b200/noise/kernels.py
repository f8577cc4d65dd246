"""Array-interface twins of the reference's numba kernels in ``src/noise/kernels.py`` (same names, argument lists and
return values), for code that imports them directly -- e.g. the reference's own ``_warmup_jit`` (engine.py:38-65).

They run on the host in plain NumPy and are NOT on the hot path: a shot on the GPU path is the XOR of precomputed fault
signatures (kernel K2, ``csrc/sampler.cu``), which these functions are the ground truth for (``tests/test_host_logic.py``
checks them against the C oracle and, through it, against the real reference).  Pauli frames: the Z-type simulator
tracks the Z component of the error on every qubit, the X-type simulator the X component.
"""
import numpy as np

from .constants import (OP_CNOT, OP_IDLE, OP_MEAS_X, OP_MEAS_Z, OP_PREP_X, OP_PREP_Z, OP_X, OP_Y, OP_Z,
                        TWO_QUBIT_ERROR_OPCODES, TWO_QUBIT_ERROR_TARGET)

# which qubits of (q1, q2) a Pauli op flips in the Z frame / the X frame: bit 0 = q1, bit 1 = q2
_PAULI = "XYZ"
_FLIPS = {"Z": {OP_Z: 1, OP_Y: 1}, "X": {OP_X: 1, OP_Y: 1}}
for _i, _a in enumerate(_PAULI):
    for _j, _b in enumerate(_PAULI):
        _code = 20 + 3 * _i + _j
        _FLIPS["Z"][_code] = (1 if _a in "ZY" else 0) | (2 if _b in "ZY" else 0)
        _FLIPS["X"][_code] = (1 if _a in "XY" else 0) | (2 if _b in "XY" else 0)


def _propagate(frame, circuit_ops, circuit_q1, circuit_q2, total_qubits, max_syndromes):
    prep, meas = (OP_PREP_X, OP_MEAS_X) if frame == "Z" else (OP_PREP_Z, OP_MEAS_Z)
    flips = _FLIPS[frame]
    state = np.zeros(total_qubits, dtype=np.int8)
    history = np.zeros(max_syndromes, dtype=np.int8)
    syn_count = err_count = 0
    for op, a, b in zip(np.asarray(circuit_ops).tolist(), np.asarray(circuit_q1).tolist(), np.asarray(circuit_q2).tolist()):
        if op == OP_CNOT:
            if frame == "Z":
                state[a] ^= state[b]          # a Z on the target copies to the control
            else:
                state[b] ^= state[a]          # an X on the control copies to the target
        elif op == prep:
            state[a] = 0
        elif op == meas:
            if syn_count < max_syndromes:
                history[syn_count] = state[a]
            syn_count += 1
        elif op >= OP_X:
            f = flips.get(op, 0)          # only Paulis with a component in this frame count as errors
            err_count += f != 0
            if f & 1:
                state[a] ^= 1
            if f & 2:
                state[b] ^= 1
    return history, state, syn_count, err_count


def simulate_circuit_Z_jit(circuit_ops, circuit_q1, circuit_q2, total_qubits, x_check_indices, x_check_ptrs, max_syndromes):
    """src/noise/kernels.py:14-91 -> (syndrome_history int8[max_syndromes], state int8[total_qubits], syn_count, err_count)."""
    return _propagate("Z", circuit_ops, circuit_q1, circuit_q2, total_qubits, max_syndromes)


def simulate_circuit_X_jit(circuit_ops, circuit_q1, circuit_q2, total_qubits, z_check_indices, z_check_ptrs, max_syndromes):
    """src/noise/kernels.py:95-172."""
    return _propagate("X", circuit_ops, circuit_q1, circuit_q2, total_qubits, max_syndromes)


def generate_noisy_circuit_jit(base_ops, base_q1, base_q2, error_rate, random_vals, random_paulis, random_two_qubit,
                               out_ops, out_q1, out_q2):
    """src/noise/kernels.py:176-353: one Bernoulli(error_rate) draw per gate (random_vals[i]); measurement faults go in
    front of the gate, preparation faults behind it, an IDLE becomes its Pauli (or disappears), a CNOT is followed by
    one of the 15 two-qubit Paulis (an out-of-range draw counts as the last one, ZX).  Fills out_* and returns the length."""
    ops, q1, q2 = np.asarray(base_ops), np.asarray(base_q1), np.asarray(base_q2)
    n = len(ops)
    fired = np.asarray(random_vals)[:n] < error_rate
    rp = np.asarray(random_paulis)[:n]
    r2 = np.asarray(random_two_qubit)[:n]
    r2 = np.where((r2 < 0) | (r2 > 14), 14, r2)
    single = (OP_X, OP_Y, OP_Z)
    o = 0

    def emit(op, a, b=-1):
        nonlocal o
        out_ops[o], out_q1[o], out_q2[o] = op, a, b
        o += 1

    for i in range(n):
        op, a, b, hit = int(ops[i]), int(q1[i]), int(q2[i]), bool(fired[i])
        if op == OP_MEAS_X or op == OP_MEAS_Z:
            if hit:
                emit(OP_Z if op == OP_MEAS_X else OP_X, a)
            emit(op, a, b)
        elif op == OP_PREP_X or op == OP_PREP_Z:
            emit(op, a, b)
            if hit:
                emit(OP_Z if op == OP_PREP_X else OP_X, a)
        elif op == OP_IDLE:
            if hit:
                emit(single[int(rp[i]) % 3], a)
        elif op == OP_CNOT:
            emit(op, a, b)
            if hit:
                t = int(r2[i])
                where = int(TWO_QUBIT_ERROR_TARGET[t])
                emit(int(TWO_QUBIT_ERROR_OPCODES[t]), b if where == 1 else a, b if where == 2 else -1)
        else:
            emit(op, a, b)
    return o


def sparsify_syndrome_jit(syndrome_history, syn_count, check_positions, check_ptrs, num_checks):
    """src/noise/kernels.py:357-380: every measurement of a check XOR the RAW previous measurement of the same check."""
    raw = np.asarray(syndrome_history)[:syn_count]
    result = raw.copy()
    pos, ptr = np.asarray(check_positions), np.asarray(check_ptrs)
    for c in range(num_checks):
        p = pos[ptr[c]:ptr[c + 1]]
        cur, prev = p[1:], p[:-1]
        ok = (cur < syn_count) & (prev < syn_count)
        result[cur[ok]] ^= raw[prev[ok]]
    return result


def extract_data_state_jit(state, data_qubit_indices):
    """src/noise/kernels.py:384-393."""
    return np.asarray(state)[np.asarray(data_qubit_indices)].astype(np.int8)
