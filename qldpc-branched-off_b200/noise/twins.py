"""Host-side (CPU) twins of the Pauli-frame simulator with the reference's names and tuple-circuit interface
(``src/noise/simulation.py:114-229``, ``src/noise/model.py:4-58``).  The reference keeps these "for
compatibility/testing"; they are not on the hot path (shots use kernel K1/K2, the table builder uses the
bit-parallel propagation in ``builder.py``) and exist here only so that code importing them keeps working."""
import numpy as np

_Z_ON_Q1 = {"Z", "Y", "ZX", "YX"}
_Z_ON_Q2 = {"XZ", "XY"}
_Z_ON_BOTH = {"ZZ", "YY", "YZ", "ZY"}
_X_ON_Q1 = {"X", "Y", "XZ", "YZ"}
_X_ON_Q2 = {"ZX", "ZY"}
_X_ON_BOTH = {"XX", "YY", "XY", "YX"}
_TWO_QUBIT = ("X_", "Y_", "Z_", "_X", "_Y", "_Z", "XX", "YY", "ZZ", "XY", "YX", "YZ", "ZY", "XZ", "ZX")


def _simulate(circuit, lin_order, checks, side):
    prep, meas = ("PrepX", "MeasX") if side == "Z" else ("PrepZ", "MeasZ")
    on1, on2, both = (_Z_ON_Q1, _Z_ON_Q2, _Z_ON_BOTH) if side == "Z" else (_X_ON_Q1, _X_ON_Q2, _X_ON_BOTH)
    state = np.zeros(len(lin_order), dtype=int)
    history, syndrome_map, errors = [], {c: [] for c in checks}, 0
    for gate in circuit:
        kind = gate[0]
        if kind == "CNOT":
            c, t = lin_order[gate[1]], lin_order[gate[2]]
            if side == "Z":
                state[c] ^= state[t]
            else:
                state[t] ^= state[c]
        elif kind == prep:
            state[lin_order[gate[1]]] = 0
        elif kind == meas:
            syndrome_map[gate[1]].append(len(history))
            history.append(state[lin_order[gate[1]]])
        elif kind in on1:
            errors += 1; state[lin_order[gate[1]]] ^= 1
        elif kind in on2:
            errors += 1; state[lin_order[gate[2]]] ^= 1
        elif kind in both:
            errors += 1; state[lin_order[gate[1]]] ^= 1; state[lin_order[gate[2]]] ^= 1
    return np.array(history, dtype=int), state, syndrome_map, errors


def simulate_circuit_Z(circuit, lin_order, n, Xchecks):
    return _simulate(circuit, lin_order, Xchecks, "Z")


def simulate_circuit_X(circuit, lin_order, n, Zchecks):
    return _simulate(circuit, lin_order, Zchecks, "X")


def sparsify_syndrome(syndrome_history, syndrome_map, checks):
    out = np.array(syndrome_history, copy=True)
    for check in checks:
        pos = syndrome_map[check]
        for a, b in zip(pos[1:], pos[:-1]):
            out[a] = (out[a] + syndrome_history[b]) % 2
    return out


def extract_data_qubit_state(state, lin_order, data_qubits):
    return np.array([state[lin_order[q]] for q in data_qubits], dtype=int)


def generate_noisy_circuit(circuit, error_rate):
    """Tuple-circuit noise insertion with the reference's draw order (model.py:4-58)."""
    noisy = []
    for gate in circuit:
        kind = gate[0]
        if kind in ("MeasX", "MeasZ"):
            if np.random.random() < error_rate:
                noisy.append(("Z" if kind == "MeasX" else "X", gate[1]))
            noisy.append(gate)
        elif kind in ("PrepX", "PrepZ"):
            noisy.append(gate)
            if np.random.random() < error_rate:
                noisy.append(("Z" if kind == "PrepX" else "X", gate[1]))
        elif kind == "IDLE":
            if np.random.random() < error_rate:
                noisy.append((["X", "Y", "Z"][np.random.randint(3)], gate[1]))
        elif kind == "CNOT":
            noisy.append(gate)
            if np.random.random() < error_rate:
                code = _TWO_QUBIT[np.random.randint(15)]
                if code[1] == "_":
                    noisy.append((code[0], gate[1]))
                elif code[0] == "_":
                    noisy.append((code[1], gate[2]))
                else:
                    noisy.append((code, gate[1], gate[2]))
        else:
            noisy.append(gate)
    return noisy
