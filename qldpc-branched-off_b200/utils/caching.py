"""Matrix cache: same key and .npz layout as the reference (src/utils/caching.py:6-42), so
existing ``matrix_cache/`` directories are read and written unchanged."""
import hashlib
import os

import numpy as np

_ARRAYS = ("HdecZ", "HdecX", "channel_probsZ", "channel_probsX", "HZ_full", "HX_full")
_SCALARS = ("first_logical_rowZ", "first_logical_rowX", "num_cycles", "k")


def compute_cache_key(Hx, Hz, Lx, Lz, num_cycles, error_rate) -> str:
    h = hashlib.sha256()
    for arr in (Hx, Hz, Lx, Lz):
        h.update(arr.tobytes())
    h.update(str(num_cycles).encode())
    h.update(f"{error_rate:.6f}".encode())
    return h.hexdigest()[:16]


def save_matrices(cache_dir, cache_key, matrices):
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"matrices_{cache_key}.npz")
    payload = {a: matrices[a] for a in _ARRAYS}
    payload.update({s: np.array([matrices[s]]) for s in _SCALARS})
    np.savez_compressed(path, **payload)
    return path


def load_matrices(cache_dir, cache_key):
    path = os.path.join(cache_dir, f"matrices_{cache_key}.npz")
    if not os.path.exists(path):
        return None
    try:
        with np.load(path) as data:
            out = {a: data[a] for a in _ARRAYS}
            out.update({s: int(data[s][0]) for s in _SCALARS})
        return out
    except Exception:
        return None
