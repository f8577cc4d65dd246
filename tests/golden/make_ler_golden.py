"""Per-shot logical-failure flags of the REAL reference, for pinning the logical error rate.

Run in the build container only (needs /root/reference and numba):
    python tests/golden/make_ler_golden.py [tag ...]          # tags: 72 90 108 144 288
Every shot is one unmodified call of the reference's own per-shot function
`src.simulation.engine._run_single_trial_fast(trial_idx, shared_data)` (engine.py:68-122) in a spawn pool, with
`shared_data` assembled exactly like `run_simulation` does (engine.py:204-212, :394-419: priors, sparse/dense switch,
compiled circuit, logical rows) from the reference's shipped `codes/*.npz` and `matrix_cache/*.npz`, alpha_mode
"dynamical", OSD order 0, base_seed 1234 (shot i draws from np.random.seed(1234 + i), engine.py:70).
Output: tests/golden/ler.npz with, per configuration `<tag>_<p*1e4>`, the packed z_err / x_err bits in shot order and
(maxIter, p, shots).  The fault events of shot i are a pure function of the seed (numpy's legacy MT19937 stream), so the
GPU tests replay them with numpy alone and compare flag by flag; nothing from this repo takes part in producing the file.
"""
import os
import sys
import time
import types

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/qldpc_golden_numba_cache")
os.environ.setdefault("NUMBA_NUM_THREADS", "1")
for _name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
REF = os.environ.get("QLDPC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np                                                     # noqa: E402
from multiprocessing import get_context                                # noqa: E402
from scipy.sparse import csr_matrix                                    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BASE_SEED = 1234
NAMES = {"72": "[[72, 12, 6]]", "90": "[[90, 8, 10]]", "108": "[[108, 8, 10]]", "144": "[[144, 12, 12]]",
         "288": "[[288, 12, 18]]"}
# (tag, p, maxIter, shots)
CONFIGS = [
    ("144", 0.005, 20, 4000),
    ("72", 0.004, 20, 20000),
    ("288", 0.006, 100, 240),
    ("90", 0.004, 20, 3000), ("90", 0.005, 20, 3000), ("90", 0.006, 20, 3000),
    ("108", 0.004, 20, 3000), ("108", 0.005, 20, 3000), ("108", 0.006, 20, 3000),
]

_shared = None


def _init(shared):
    global _shared
    import src.simulation.engine as eng
    _shared = shared
    eng._warmup_jit()


def _task(i):
    import src.simulation.engine as eng
    z, x, _ = eng._run_single_trial_fast(i, _shared)
    return bool(z), bool(x)


def shared_data_for(tag, p, max_iter):
    """engine.py:204-212 and :394-419, dynamical alpha, OSD-0."""
    from src.codes.bb_code import BBCodeCircuit
    from src.noise.compiled import CompiledCircuit
    from src.utils.caching import compute_cache_key, load_matrices
    d = np.load(os.path.join(REF, "codes", f"{NAMES[tag]}.npz"))
    Hx, Hz, Lx, Lz = d["Hx"], d["Hz"], d["Lx"], d["Lz"]
    bb = {k: d[k] for k in ["ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers"]}
    dist = int(d["distance"])
    cb = BBCodeCircuit(Hx, Hz, num_cycles=dist, **bb)
    M = load_matrices(os.path.join(REF, "matrix_cache"), compute_cache_key(Hx, Hz, Lx, Lz, dist, p))
    assert M is not None, (tag, p)
    with np.errstate(divide="ignore", invalid="ignore"):
        llrs_z = np.clip(np.nan_to_num(np.log((1 - M["channel_probsZ"]) / M["channel_probsZ"])), -50, 50)
        llrs_x = np.clip(np.nan_to_num(np.log((1 - M["channel_probsX"]) / M["channel_probsX"])), -50, 50)
    use_sparse = M["HdecZ"].shape[1] > 5000
    cc = CompiledCircuit(base_circuit=cb.get_full_circuit(), noiseless_suffix=cb.cycle * 2, lin_order=cb.lin_order,
                         data_qubits=cb.data_qubits, Xchecks=cb.Xchecks, Zchecks=cb.Zchecks)
    fz, fx, k = M["first_logical_rowZ"], M["first_logical_rowX"], Lx.shape[0]
    return {
        "error_rate": p, "Lx": Lx, "Lz": Lz,
        "HdecZ": np.asarray(M["HdecZ"], dtype=np.float64, order="C"),
        "HdecX": np.asarray(M["HdecX"], dtype=np.float64, order="C"), "llrs_z": llrs_z, "llrs_x": llrs_x,
        "HZ_logical": np.ascontiguousarray(M["HZ_full"][fz:fz + k]),
        "HX_logical": np.ascontiguousarray(M["HX_full"][fx:fx + k]),
        "alpha_mode": "dynamical", "alpha_z": 1.0, "alpha_x": 1.0, "maxIter": max_iter, "osd_order": 0,
        "base_seed": BASE_SEED, "use_sparse": use_sparse,
        "HdecZ_csr": csr_matrix(M["HdecZ"]) if use_sparse else None,
        "HdecX_csr": csr_matrix(M["HdecX"]) if use_sparse else None,
        "compiled_circuit": cc,
    }


def main():
    which = set(sys.argv[1:]) or set(NAMES)
    workers = int(os.environ.get("LER_WORKERS", "6"))
    path = os.path.join(HERE, "ler.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for tag, p, max_iter, shots in CONFIGS:
        if tag not in which:
            continue
        key = f"{tag}_{int(round(p * 1e4))}"
        t0 = time.time()
        shared = shared_data_for(tag, p, max_iter)
        nw = min(workers, 3) if tag == "288" else workers          # OSD copies the dense 2880 x 26k float matrix
        with get_context("spawn").Pool(nw, initializer=_init, initargs=(shared,)) as pool:
            flags = pool.map(_task, range(shots), chunksize=max(1, shots // (nw * 16)))
        z = np.array([f[0] for f in flags], dtype=np.uint8)
        x = np.array([f[1] for f in flags], dtype=np.uint8)
        out[key + "_z"] = np.packbits(z, bitorder="little")
        out[key + "_x"] = np.packbits(x, bitorder="little")
        out[key + "_meta"] = np.array([p, max_iter, shots, BASE_SEED], dtype=np.float64)
        print(f"{key}: shots {shots} maxIter {max_iter} errors z {z.sum()} x {x.sum()} total {(z | x).sum()} "
              f"LER {(z | x).mean():.4f}  ({time.time() - t0:.0f} s, {nw} workers)", flush=True)
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
