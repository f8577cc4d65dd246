"""Per-shot noise sampling + syndrome extraction on the GPU (reference ``src/noise/simulation.py``).

``run_trial_fast`` keeps the reference's signature, return tuple *and* its use of the legacy
``np.random`` stream (simulation.py:43-45), so with the same ``np.random.seed`` it returns exactly
the reference's arrays; the Pauli-frame propagation, detector differencing and logical product are
done by kernel K2 as an XOR of fault signatures (``qb_syndrome_from_events_host``).
"""
import numpy as np

from .. import _lib
from ..codes.bb_code import OP_CNOT, OP_IDLE
from .builder import fault_tables_for


def _sampler_for(compiled, Lx, Lz):
    ft = fault_tables_for(compiled, Lx, Lz)
    cache = compiled.__dict__.setdefault("_gpu_samplers", {})
    key = (id(ft), _lib.default_device())
    s = cache.get(key)
    if s is None:
        s = _lib.Sampler(ft)
        cache[key] = s
    return s


def events_from_random(compiled, error_rate, random_vals, random_paulis, random_two_qubit):
    """Fault events (location | outcome << 24) of one shot from the reference's three random arrays."""
    fired = np.nonzero(np.asarray(random_vals) < error_rate)[0]
    ops = compiled.base_ops[fired]
    outcome = np.where(ops == OP_CNOT, np.asarray(random_two_qubit)[fired],
                       np.where(ops == OP_IDLE, np.asarray(random_paulis)[fired], 0))
    return (fired.astype(np.uint32) | (outcome.astype(np.uint32) << np.uint32(24))).astype(np.uint32)


def run_trial_fast(compiled, error_rate, Lx, Lz):
    """Drop-in for reference ``run_trial_fast`` (simulation.py:21-107):
    returns (sparse_z int8[m], true_z int8[k], sparse_x int8[m], true_x int8[k])."""
    n_locs = compiled.num_error_locs
    random_vals = np.random.random(n_locs)
    random_paulis = np.random.randint(0, 3, n_locs, dtype=np.int32)
    random_two_qubit = np.random.randint(0, 15, n_locs, dtype=np.int32)
    ev = events_from_random(compiled, error_rate, random_vals, random_paulis, random_two_qubit)
    s = _sampler_for(compiled, Lx, Lz)
    sz, tz, sx, tx = s.syndromes_from_events(np.array([0, len(ev)], dtype=np.int32), ev)
    return sz[0], tz[0], sx[0], tx[0]


def run_trials_from_events(compiled, Lx, Lz, ev_ptr, events):
    """Batched K2: events of many shots (CSR) -> (sparse_z [B,m], true_z [B,k], sparse_x, true_x)."""
    return _sampler_for(compiled, Lx, Lz).syndromes_from_events(ev_ptr, events)
