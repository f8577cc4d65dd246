"""Code files without the third-party ``qldpc`` package (reference ``generate_codes.py``).

The reference builds its five bivariate-bicycle codes with ``qldpc.codes.BBCode`` and takes ``get_logical_ops()`` from
it (generate_codes.py:9-80, :131-140); neither is available offline.  Here the parity checks come from the polynomial
exponents (``bb_parity_matrices``: hash-identical to the reference's shipped ``codes/*.npz``) and the logical operators
from a GF(2) nullspace computation (``css_logicals``): a different but equally valid basis -- Hz.Lx^T = 0, Hx.Lz^T = 0
and Lx.Lz^T = I, checked for all five codes in the tests.  Logical failure flags do not depend on the basis whenever
the decoder output reproduces the syndrome (every OSD-0 output does).

    python -m qldpc_b200.codes.generate [out_dir]

writes ``<out_dir>/<name>.npz`` with exactly the keys and dtypes of generate_codes.py:154-168.
"""
import os
import sys

import numpy as np

from .bb_code import BB_CODES, make_bb_code

KEYS = ("Hx", "Hz", "Lx", "Lz", "distance", "ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")


def code_arrays(name):
    """The arrays of ``codes/<name>.npz`` (generate_codes.py:154-168): Hx, Hz int64; Lx, Lz uint8; scalars and exponent lists int64."""
    c = make_bb_code(name)
    out = {
        "Hx": np.asarray(c["Hx"], dtype=np.int64), "Hz": np.asarray(c["Hz"], dtype=np.int64),
        "Lx": np.asarray(c["Lx"], dtype=np.uint8), "Lz": np.asarray(c["Lz"], dtype=np.uint8),
        "distance": np.int64(c["distance"]), "ell": np.int64(c["ell"]), "m": np.int64(c["m"]),
    }
    for key in ("a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers"):
        out[key] = np.asarray(c[key], dtype=np.int64)
    return out


def write_code_npz(name, out_dir="codes"):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{name}.npz")
    np.savez(path, **code_arrays(name))
    return path


def generate_all(out_dir="codes", names=None):
    return [write_code_npz(name, out_dir) for name in (names or BB_CODES)]


if __name__ == "__main__":
    for p in generate_all(sys.argv[1] if len(sys.argv) > 1 else "codes"):
        print("saved", p)
