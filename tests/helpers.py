"""Shared helpers for the tests: code/circuit/table construction with the reference's logicals."""
import functools
import json
import os

import numpy as np

import qldpc_b200  # noqa: F401  (import shim)
from qldpc_b200.codes.bb_code import BB_CODES, BBCodeCircuit, bb_parity_matrices
from qldpc_b200.noise.builder import fault_tables_for, matrices_from_tables
from qldpc_b200.noise.compiled import CompiledCircuit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = {"72": "[[72, 12, 6]]", "90": "[[90, 8, 10]]", "108": "[[108, 8, 10]]",
         "144": "[[144, 12, 12]]", "288": "[[288, 12, 18]]"}


def unpack(bits, n):
    return np.unpackbits(np.asarray(bits, dtype=np.uint8), bitorder="little", axis=-1)[..., :n]


@functools.lru_cache(maxsize=None)
def reference_logicals(tag):
    d = np.load(os.path.join(GOLDEN, "reference_logicals.npz"))
    return d[f"Lx_{tag}"], d[f"Lz_{tag}"]


@functools.lru_cache(maxsize=None)
def builder_hashes():
    with open(os.path.join(GOLDEN, "builder_hashes.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def code_setup(tag):
    """(spec, Hx, Hz, Lx, Lz, circuit builder, compiled circuit, fault tables) with the
    reference's own logical operators (so logical rows equal the reference's cache files)."""
    name = NAMES[tag]
    spec = BB_CODES[name]
    Hx, Hz = bb_parity_matrices(**spec)
    Lx, Lz = reference_logicals(tag)
    bb = {k: spec[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
    cb = BBCodeCircuit(Hx, Hz, num_cycles=spec["distance"], **bb)
    cc = CompiledCircuit.from_builder(cb)
    ft = fault_tables_for(cc, Lx, Lz)
    return dict(name=name, spec=spec, bb=bb, Hx=Hx, Hz=Hz, Lx=Lx, Lz=Lz, cb=cb, cc=cc, ft=ft)


@functools.lru_cache(maxsize=None)
def matrices(tag, p):
    s = code_setup(tag)
    return matrices_from_tables(s["ft"], p, s["spec"]["distance"])
