"""Integer gate codes of the array-form circuits (reference ``src/noise/constants.py:8-77``): same names, same
values, because compiled circuits built here must stay interchangeable with the reference's (``base_ops`` arrays
are compared value by value in the tests)."""
import numpy as np

from ..codes.bb_code import OP_CNOT, OP_IDLE, OP_MEAS_X, OP_MEAS_Z, OP_PREP_X, OP_PREP_Z

OP_X, OP_Y, OP_Z = 10, 11, 12
(OP_XX, OP_XY, OP_XZ, OP_YX, OP_YY, OP_YZ, OP_ZX, OP_ZY, OP_ZZ) = range(20, 29)

GATE_TO_OPCODE = {
    "CNOT": OP_CNOT, "PrepX": OP_PREP_X, "PrepZ": OP_PREP_Z, "MeasX": OP_MEAS_X, "MeasZ": OP_MEAS_Z, "IDLE": OP_IDLE,
    "X": OP_X, "Y": OP_Y, "Z": OP_Z,
    **{a + b: 20 + 3 * i + j for i, a in enumerate("XYZ") for j, b in enumerate("XYZ")},
}
OPCODE_TO_GATE = {v: k for k, v in GATE_TO_OPCODE.items()}

# outcome r of the uniform draw over the 15 two-qubit Paulis after a CNOT (src/noise/kernels.py:283-342):
# r = 0..2 X/Y/Z on the control, 3..5 X/Y/Z on the target, then XX YY ZZ XY YX YZ ZY XZ ZX
TWO_QUBIT_ERROR_OPCODES = np.array([OP_X, OP_Y, OP_Z, OP_X, OP_Y, OP_Z, OP_XX, OP_YY, OP_ZZ, OP_XY, OP_YX, OP_YZ, OP_ZY,
                                    OP_XZ, OP_ZX], dtype=np.int32)
TWO_QUBIT_ERROR_TARGET = np.array([0] * 3 + [1] * 3 + [2] * 9, dtype=np.int32)      # 0 control, 1 target, 2 both
