"""CPU tests of the host side: C-ABI exports, alpha-mode validation, sharding / early-stop logic,
fault-event construction, and the world_size-2 (gloo) reduction path."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import builder_hashes, code_setup
import qldpc_b200  # noqa: F401
from qldpc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qldpc_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTS)
    assert b"sm_100a" in _lib.load().qb_version()


def test_no_cpu_fallback_without_gpu():
    if _lib.load().qb_device_count() > 0:
        pytest.skip("GPU present")
    from scipy.sparse import identity
    from qldpc_b200.decoding.sparse import performMinSum_Symmetric_Sparse
    with pytest.raises(_lib.QbError):
        performMinSum_Symmetric_Sparse(identity(3, format="csr"), np.zeros(3, np.int8), np.ones(3))


def test_alpha_mode_validation_matches_reference():
    from qldpc_b200.decoding.sparse import _alpha_dispatch
    assert _alpha_dispatch(0, None)[0] == _lib.QB_ALPHA_DYNAMIC
    assert _alpha_dispatch(0.7, None)[:2] == (_lib.QB_ALPHA_FIXED, 0.7)
    assert _alpha_dispatch(1.0, "dynamical")[0] == _lib.QB_ALPHA_DYNAMIC
    with pytest.raises(ValueError):
        _alpha_dispatch(0.0, "alvarado")
    with pytest.raises(ValueError):
        _alpha_dispatch(1.0, "nope")
    with pytest.raises(ValueError):
        _alpha_dispatch(np.zeros((2, 2)), "alvarado-autoregressive")
    with pytest.raises(ValueError):
        _alpha_dispatch(np.zeros(0), "alvarado-autoregressive")
    mode, _, seq = _alpha_dispatch([0.3, 0.5], "alvarado-autoregressive")
    assert mode == _lib.QB_ALPHA_SEQUENCE and list(seq) == [0.3, 0.5]


def test_shard_and_early_stop_logic():
    from qldpc_b200.simulation.engine import early_stop_cut, shard_range
    for total in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    flags = np.array([0, 1, 0, 0, 3, 2, 0, 1], dtype=np.uint8)
    assert early_stop_cut(flags, 1) == 2
    assert early_stop_cut(flags, 3) == 6
    assert early_stop_cut(flags, 2, errors_before=1) == 2
    assert early_stop_cut(flags, 5) is None


def test_events_from_random_matches_golden():
    from qldpc_b200.noise.simulation import events_from_random
    g = np.load(os.path.join(ROOT, "tests", "golden", "shots_72.npz"))
    cc = code_setup("72")["cc"]
    p, L = float(g["p"]), cc.num_error_locs
    for i in range(4):
        np.random.seed(int(g["base_seed"]) + i)
        rv = np.random.random(L); rp = np.random.randint(0, 3, L, dtype=np.int32); r2 = np.random.randint(0, 15, L, dtype=np.int32)
        ev = events_from_random(cc, p, rv, rp, r2)
        lo, hi = g["ev_ptr"][i], g["ev_ptr"][i + 1]
        assert np.array_equal(ev & 0xFFFFFF, g["ev_loc"][lo:hi].astype(np.uint32))
        assert np.array_equal(ev >> 24, g["ev_outcome"][lo:hi].astype(np.uint32))


def test_own_logicals_are_valid():
    from qldpc_b200.codes.bb_code import make_bb_code
    c = make_bb_code("[[72, 12, 6]]")
    Hx, Hz, Lx, Lz = c["Hx"], c["Hz"], c["Lx"], c["Lz"]
    assert Lx.shape == (12, 72) and Lz.shape == (12, 72)
    assert not ((Hz @ Lx.T) % 2).any() and not ((Hx @ Lz.T) % 2).any()
    assert np.array_equal((Lx.astype(int) @ Lz.T.astype(int)) % 2, np.eye(12, dtype=int))


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
import qldpc_b200
from qldpc_b200.simulation import engine
rank, world = int(sys.argv[2]), 2
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3])
dist.init_process_group("gloo", rank=rank, world_size=world)
round_total = 11
lo, hi = engine.shard_range(round_total, rank, world)
flags = (np.arange(lo, hi) % 3 == 0).astype(np.uint8)
allf = engine._gather_flags(dist, world, flags, round_total)
assert np.array_equal(allf, (np.arange(round_total) % 3 == 0).astype(np.uint8)), allf
c = engine._reduce_counts(dist, np.array([rank + 1, 10, 0, hi - lo, 0, 0, 0, 0], dtype=np.int64))
assert list(c[:4]) == [3, 20, 0, round_total], c
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_install_as_src_aliases_reference_module_names():
    import importlib
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    try:
        qldpc_b200.install_as_src()
        eng = importlib.import_module("src.simulation.engine")
        assert hasattr(eng, "run_simulation")
        from src.decoding.sparse import performMinSum_Symmetric_Sparse  # noqa: F401
        from src.decoding.osd import performOSD_enhanced  # noqa: F401
        from src.noise.builder import build_decoding_matrices  # noqa: F401
        from src.utils.caching import compute_cache_key, load_matrices, save_matrices  # noqa: F401
        from src.codes.bb_code import BBCodeCircuit  # noqa: F401
        import inspect
        sig = inspect.signature(eng.run_simulation)
        for kw in ("Hx", "Hz", "Lx", "Lz", "error_rate", "num_trials", "num_cycles", "maxIter", "osd_order", "use_dynamic_alpha",
                   "alpha_mode", "alvarado_alpha", "precomputed_matrices", "num_workers", "base_seed", "use_jit",
                   "target_logical_errors", "max_trials", "scopt", "estimation_plot_dir"):
            assert kw in sig.parameters
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_host_twins_agree_with_table_builder():
    """The tuple-circuit twins (reference simulation.py:114-229) and the bit-parallel builder describe the same
    physics: a single inserted fault reproduces its decoding-matrix column and logical mask."""
    from qldpc_b200.noise import simulate_circuit_Z, simulate_circuit_X, sparsify_syndrome, extract_data_qubit_state
    s = code_setup("72"); cb, ft = s["cb"], s["ft"]
    base, suffix = cb.get_full_circuit(), cb.cycle * 2
    rng = np.random.default_rng(0)
    for side, sim, checks, L, tab in (("Z", simulate_circuit_Z, cb.Xchecks, s["Lx"], ft.Z), ("X", simulate_circuit_X, cb.Zchecks, s["Lz"], ft.X)):
        for f in rng.choice(len(tab.fault_loc), 12, replace=False):
            loc, var = int(tab.fault_loc[f]), int(tab.fault_variant[f])
            gate = base[loc]
            p = side
            err = (p, gate[1]) if var == 1 else ((p, gate[2]) if var == 2 else (p + p, gate[1], gate[2]))
            pos = loc if gate[0].startswith("Meas") else loc + 1
            circ = base[:pos] + [err] + base[pos:] + suffix
            hist, state, smap, _ = sim(circ, cb.lin_order, cb.n, checks)
            syn = sparsify_syndrome(hist, smap, checks)
            col = int(tab.fault_col[f])
            rows = tab.col_rows[tab.col_ptr[col]:tab.col_ptr[col + 1]]
            assert np.array_equal(np.nonzero(syn)[0], rows)
            logical = (np.asarray(L) @ extract_data_qubit_state(state, cb.lin_order, cb.data_qubits)) % 2
            mask = sum(int(b) << i for i, b in enumerate(logical))
            assert mask == int(tab.col_logmask[col])


def test_plotting_shim_returns_reference_r2():
    from qldpc_b200.utils.plotting import plot_alpha_comparison, plot_alpha_linearity, plot_simulation_results
    res = {"72": {0.005: {"logical_error_rate": 0.3, "alpha_values_z": [0.4, 0.5, 0.6], "alpha_values_x": [0.3, 0.6, 0.5]}}}
    r2 = plot_alpha_linearity(res, "/tmp/_qb_lin.png")
    assert abs(r2["72"][0.005]["z"] - 1.0) < 1e-12 and 0 < r2["72"][0.005]["x"] < 1
    plot_alpha_comparison(res, "/tmp/_qb_cmp.png")
    plot_simulation_results({"72": {0.004: {"logical_error_rate": 0.1}, 0.006: {"logical_error_rate": 0.4}}}, "/tmp/_qb_res.png")


def _layout_stats(H, prior, nwarps=32):
    import scipy.sparse as sp
    H = sp.csr_matrix(H); H.sort_indices()
    st = np.zeros(16, np.int64)
    ip, ix = H.indptr.astype(np.int32), H.indices.astype(np.int32)
    pr = np.ascontiguousarray(prior, dtype=np.float64)
    rc = _lib.load().qb_edge_layout_probe(H.shape[0], H.shape[1], _lib.ptr(ip), _lib.ptr(ix), _lib.ptr(pr), nwarps, _lib.ptr(st))
    assert rc == 0
    names = "ok n_rsl n_csl e_words idx_words groups wavefronts conflict_edges uniform_prior max_K max_cdeg consistent".split()
    return dict(zip(names, st.tolist()))


def test_edge_layout_is_consistent_and_conflict_free_for_bb_codes():
    """Host part of the per-edge min-sum kernel (csrc/edge_layout.cu): every Tanner edge owns exactly one shared-memory
    slot, and the slot colouring leaves (almost) no bank conflict in the variable-phase gathers."""
    from helpers import matrices
    for tag in ("72", "144"):
        M = matrices(tag, 0.005)
        for side in "ZX":
            H = np.asarray(M["Hdec" + side]) & 1
            probs = np.asarray(M["channel_probs" + side], dtype=np.float64)
            with np.errstate(all="ignore"):
                prior = np.clip(np.nan_to_num(np.log((1 - probs) / probs)), -50, 50)
            s = _layout_stats(H, prior)
            assert s["ok"] == 1 and s["consistent"] == 1 and s["uniform_prior"] == 1
            assert s["wavefronts"] <= 1.03 * s["groups"], s          # < 3 % extra wavefronts from bank conflicts
            assert (s["e_words"] + s["idx_words"]) * 4 < 220 * 1024   # fits one SM next to the small tables
            assert s["max_K"] == 9 and s["max_cdeg"] == 6


def test_edge_layout_generic_graphs():
    rng = np.random.default_rng(5)
    # random sparse graph with many distinct priors -> per-lane priors, still consistent
    H = (rng.random((40, 150)) < 0.08).astype(np.int8)
    s = _layout_stats(H, rng.normal(3.0, 1.0, 150))
    assert s["ok"] == 1 and s["consistent"] == 1 and s["uniform_prior"] == 0
    # degree-0 rows and columns, degree-1 rows, uniform prior
    H = np.zeros((6, 9), np.int8); H[0, :4] = 1; H[1, 2:7] = 1; H[2, 8] = 1; H[4, 0] = 1
    s = _layout_stats(H, np.full(9, 2.5))
    assert s["ok"] == 1 and s["consistent"] == 1 and s["uniform_prior"] == 1
    # too large for 16-bit slot indices: the decoder falls back to the compressed-state kernel
    H = (rng.random((300, 3000)) < 0.1).astype(np.int8)
    s = _layout_stats(H, np.full(3000, 1.0))
    assert s["ok"] == 0
    # non-finite prior: not usable
    assert _layout_stats(np.eye(3, dtype=np.int8), [1.0, np.inf, 1.0])["ok"] == 0


def test_noise_kernel_twins_match_oracle():
    """src.noise.kernels / .model / .constants aliases: the array-interface twins compose to the reference's
    run_trial_fast (simulation.py:21-107), checked against the C oracle (itself pinned to the real reference)."""
    import qldpc_b200
    from qldpc_b200.noise import kernels as K, constants as Cn, model
    from oracle import oracle as orc
    s = code_setup("72"); cc = s["cc"]; p = 0.02
    assert Cn.GATE_TO_OPCODE["CNOT"] == 1 and Cn.GATE_TO_OPCODE["ZX"] == 26 and Cn.OP_IDLE == 6 and callable(model.generate_noisy_circuit)
    src = qldpc_b200.install_as_src()
    import importlib
    assert importlib.import_module("src.noise.kernels") is K and importlib.import_module("src.noise.constants") is Cn
    assert importlib.import_module("src.noise.model") is model
    n = cc.num_error_locs
    rng = np.random.default_rng(4)
    for _ in range(3):
        rv = rng.random(n); rp = rng.integers(0, 3, n).astype(np.int32); r2 = rng.integers(0, 15, n).astype(np.int32)
        out = [np.empty(2 * n + 8, dtype=np.int32) for _ in range(3)]
        ln = K.generate_noisy_circuit_jit(cc.base_ops, cc.base_q1, cc.base_q2, p, rv, rp, r2, *out)
        ops = np.concatenate([out[0][:ln], cc.suffix_ops]); q1 = np.concatenate([out[1][:ln], cc.suffix_q1]); q2 = np.concatenate([out[2][:ln], cc.suffix_q2])
        hz, stz, nz, ez = K.simulate_circuit_Z_jit(ops, q1, q2, cc.total_qubits, None, None, cc.num_meas_x + 100)
        hx, stx, nx, ex = K.simulate_circuit_X_jit(ops, q1, q2, cc.total_qubits, None, None, cc.num_meas_z + 100)
        fired = int((rv < p).sum())
        assert nz == cc.num_meas_x and nx == cc.num_meas_z and 0 < ez <= fired and 0 < ex <= fired and ez + ex >= fired
        sz = K.sparsify_syndrome_jit(hz, nz, cc.x_syn_positions, cc.x_syn_ptrs, cc.num_x_checks)
        sx = K.sparsify_syndrome_jit(hx, nx, cc.z_syn_positions, cc.z_syn_ptrs, cc.num_z_checks)
        tz = (s["Lx"].astype(int) @ K.extract_data_state_jit(stz, cc.data_qubit_indices)) % 2
        tx = (s["Lz"].astype(int) @ K.extract_data_state_jit(stx, cc.data_qubit_indices)) % 2
        osz, otz, osx, otx = orc.run_trial_arrays(cc, p, s["Lx"], s["Lz"], rv, rp, r2)
        assert np.array_equal(sz, osz) and np.array_equal(sx, osx) and np.array_equal(tz, otz) and np.array_equal(tx, otx)


def test_code_generator_matches_reference_format_and_logicals_are_valid_for_all_codes(tmp_path):
    """generate_codes.py:154-168 without qldpc: same keys / dtypes / Hx, Hz as the reference's shipped files (hashes
    recorded from them by make_golden.py), and own logical operators that are valid for all five codes."""
    from qldpc_b200.codes.generate import KEYS, generate_all
    from qldpc_b200.utils.gf2 import rank
    import hashlib
    paths = generate_all(str(tmp_path))
    assert len(paths) == 5
    hashes = builder_hashes()
    for path in paths:
        name = os.path.basename(path)[:-4]
        d = np.load(path)
        assert tuple(d.files) == KEYS
        Hx, Hz, Lx, Lz = d["Hx"], d["Hz"], d["Lx"], d["Lz"]
        assert Hx.dtype == np.int64 and Lx.dtype == np.uint8 and d["distance"].shape == () and d["a_x_powers"].dtype == np.int64
        ref = hashes[name]
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        assert sha(Hx) == ref["Hx"] and sha(Hz) == ref["Hz"] and int(d["distance"]) == ref["distance"]
        k = ref["k"]
        assert Lx.shape == (k, Hx.shape[1]) and Lz.shape == (k, Hx.shape[1])
        Lxi, Lzi = Lx.astype(np.int64), Lz.astype(np.int64)
        assert not ((Hz @ Lxi.T) % 2).any() and not ((Hx @ Lzi.T) % 2).any(), name       # commute with the stabilisers
        assert np.array_equal((Lxi @ Lzi.T) % 2, np.eye(k, dtype=np.int64)), name          # symplectic pairs
        assert rank(np.vstack([Hx, Lxi])) == rank(Hx) + k and rank(np.vstack([Hz, Lzi])) == rank(Hz) + k, name


@pytest.mark.skipif(not os.path.exists("/root/reference/main.py"), reason="reference checkout not present")
def test_unmodified_reference_main_runs_up_to_the_gpu_boundary(tmp_path):
    """the reference's main.py, unmodified, under install_as_src(): every import of main.py:1-10 resolves,
    codes/*.npz from our generator load, compute_cache_key / load_matrices / build_decoding_matrices / save_matrices run,
    and run_simulation is entered with main.py's keyword arguments (main.py:83-88) -- in this container it must then stop
    with the library's "no CUDA device" error (there is no CPU fallback).  The GPU half of the same sequence is
    tests/test_gpu_parity.py::test_main_py_call_sequence_on_the_gpu."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import tools.run_reference_main as r\n"
        "r.run('/root/reference', %r)\n" % (ROOT, str(tmp_path)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CUDA device" in res.stderr and "QbError" in res.stderr, res.stderr[-2000:]
    assert "run_simulation" in res.stderr, "the failure must come from inside run_simulation"
    assert os.path.isdir(tmp_path / "matrix_cache") and len(os.listdir(tmp_path / "matrix_cache")) == 1, "the built matrices were cached"
    assert len(os.listdir(tmp_path / "codes")) == 5


def test_sampler_jump_table():
    """qb_sampler_geometric_table: T[k] = floor(2^32 (1-p)^k), strictly the inversion table of P(G >= k) = (1-p)^k."""
    for p in (0.0001, 0.005, 0.02, 0.3):
        T = _lib.geometric_table(p).astype(np.float64)
        k = np.arange(1, 1024)
        assert np.all(np.diff(T[1:]) <= 0)
        np.testing.assert_allclose(T[1:] / 2.0 ** 32, (1 - p) ** k, rtol=0, atol=2.0 ** -32 * 1.5)
    with pytest.raises(ValueError):
        _lib.geometric_table(0.0)


def test_circuit_tables_are_cached_per_code_and_cycle_count():
    """run_simulation builds the compiled circuit and the fault signature tables once per (code, cycles) and process: the
    reference's sweep over error rates (main.py:95-141) calls it once per error rate."""
    from qldpc_b200.codes.bb_code import make_bb_code
    from qldpc_b200.simulation import engine
    code = make_bb_code("[[72, 12, 6]]")
    bb = {k: code[k] for k in ("ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers")}
    engine._TABLE_CACHE.clear()
    a = engine._circuit_tables(code["Hx"], code["Hz"], code["Lx"], code["Lz"], 3, bb)
    b = engine._circuit_tables(code["Hx"].copy(), code["Hz"], code["Lx"], code["Lz"], 3, dict(bb))
    c = engine._circuit_tables(code["Hx"], code["Hz"], code["Lx"], code["Lz"], 2, bb)
    assert a[0] is b[0] and a[1] is b[1] and c[0] is not a[0] and len(engine._TABLE_CACHE) == 2
    Lz2 = code["Lz"].copy(); Lz2[0] ^= Lz2[1]
    d = engine._circuit_tables(code["Hx"], code["Hz"], code["Lx"], Lz2, 3, bb)
    assert d[1] is not a[1]
    for cyc in range(4, 9):
        engine._circuit_tables(code["Hx"], code["Hz"], code["Lx"], code["Lz"], cyc, bb)
    assert len(engine._TABLE_CACHE) == 4
