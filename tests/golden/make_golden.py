"""Generate the golden fixtures in tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference and numba):
    python tests/golden/make_golden.py
The reference is imported unmodified; matplotlib (absent here, imported by alpha.py/scopt.py/
plotting.py only) is stubbed.  Nothing from the repo's own package or oracle is used to produce
these files -- they are the pin the oracle (oracle/) and the CUDA path are tested against.
"""
import hashlib
import json
import os
import sys
import types

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/qldpc_golden_numba_cache")
for _name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(_name, types.ModuleType(_name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
REF = os.environ.get("QLDPC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np                                                     # noqa: E402
from scipy.sparse import csr_matrix                                    # noqa: E402
from src.codes.bb_code import BBCodeCircuit                            # noqa: E402
from src.decoding import kernels as rk                                 # noqa: E402
from src.decoding.dense import performBeliefPropagationFast, performMinSum_Symmetric   # noqa: E402
from src.decoding.osd import performOSD_enhanced                       # noqa: E402
from src.decoding.sparse import performMinSum_Symmetric_Sparse         # noqa: E402
from src.noise.compiled import CompiledCircuit                         # noqa: E402
from src.noise.simulation import run_trial_fast                        # noqa: E402
from src.utils.caching import compute_cache_key, load_matrices         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CODES = ["[[72, 12, 6]]", "[[90, 8, 10]]", "[[108, 8, 10]]", "[[144, 12, 12]]", "[[288, 12, 18]]"]
BASE_SEED = 1234


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def llrs(cp):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.clip(np.nan_to_num(np.log((1 - cp) / cp)), -50, 50)      # engine.py:210-212


def pack(a):
    return np.packbits(np.asarray(a, dtype=np.uint8) & 1, bitorder="little")


def load_code(name):
    d = np.load(os.path.join(REF, "codes", f"{name}.npz"))
    bb = {k: d[k] for k in ["ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers"]}
    return d, bb


def builder_fixture():
    """Hashes of the reference's shipped code files and matrix_cache arrays + its Lx/Lz."""
    out, logicals = {}, {}
    for name in CODES:
        d, _ = load_code(name)
        dist = int(d["distance"])
        ent = {"Hx": sha(d["Hx"].astype(np.int64)), "Hz": sha(d["Hz"].astype(np.int64)), "distance": dist,
               "k": int(d["Lx"].shape[0]), "cache": {}}
        tag = name.split(",")[0].strip("[ ")
        logicals[f"Lx_{tag}"] = np.asarray(d["Lx"], dtype=np.uint8)
        logicals[f"Lz_{tag}"] = np.asarray(d["Lz"], dtype=np.uint8)
        for p in (0.004, 0.005, 0.006):
            key = compute_cache_key(d["Hx"], d["Hz"], d["Lx"], d["Lz"], dist, p)
            M = load_matrices(os.path.join(REF, "matrix_cache"), key)
            if M is None:
                continue
            ent["cache"][f"{p:.6f}"] = {
                "key": key,
                "shapeZ": list(M["HdecZ"].shape), "shapeX": list(M["HdecX"].shape),
                **{a: sha(np.asarray(M[a])) for a in ("HdecZ", "HdecX", "channel_probsZ", "channel_probsX", "HZ_full", "HX_full")},
            }
        out[name] = ent
    with open(os.path.join(HERE, "builder_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "reference_logicals.npz"), **logicals)


def shots_fixture(name, p, n_shots, n_values, n_osd, max_iter=20):
    d, bb = load_code(name)
    dist = int(d["distance"])
    Lx, Lz = d["Lx"], d["Lz"]
    cb = BBCodeCircuit(d["Hx"], d["Hz"], num_cycles=dist, **bb)
    cc = CompiledCircuit(cb.get_full_circuit(), cb.cycle * 2, cb.lin_order, cb.data_qubits, cb.Xchecks, cb.Zchecks)
    M = load_matrices(os.path.join(REF, "matrix_cache"), compute_cache_key(d["Hx"], d["Hz"], Lx, Lz, dist, p))
    m = M["first_logical_rowZ"]
    k = Lx.shape[0]
    side = {
        "z": dict(csr=csr_matrix(M["HdecZ"]), dense=np.asarray(M["HdecZ"], dtype=np.float64),
                  llr=llrs(M["channel_probsZ"]), logical=np.ascontiguousarray(M["HZ_full"][m:m + k])),
        "x": dict(csr=csr_matrix(M["HdecX"]), dense=np.asarray(M["HdecX"], dtype=np.float64),
                  llr=llrs(M["channel_probsX"]), logical=np.ascontiguousarray(M["HX_full"][m:m + k])),
    }
    out = dict(p=p, n_shots=n_shots, base_seed=BASE_SEED, max_iter=max_iter, m=m, k=k)
    ev_ptr, ev_loc, ev_out = [0], [], []
    rec = {s: dict(syn=[], true=[], hard=[], conv=[], fin=[], values=[], err=[], osd_shot=[], osd_order=[], osd_sol=[]) for s in "zx"}
    L = cc.num_error_locs
    for i in range(n_shots):
        # the draws of simulation.py:43-45, replayed to record which locations fired
        np.random.seed(BASE_SEED + i)
        rv = np.random.random(L)
        rp = np.random.randint(0, 3, L, dtype=np.int32)
        r2 = np.random.randint(0, 15, L, dtype=np.int32)
        fired = np.nonzero(rv < p)[0]
        ops = cc.base_ops[fired]
        outcome = np.where(ops == 1, r2[fired], np.where(ops == 6, rp[fired], 0))
        ev_loc += fired.tolist(); ev_out += outcome.tolist(); ev_ptr.append(len(ev_loc))
        np.random.seed(BASE_SEED + i)
        sz, tz, sx, tx = run_trial_fast(cc, p, Lx, Lz)                        # engine.py:70-80
        for s, syn, tl in (("z", sz, tz), ("x", sx, tx)):
            S, r = side[s], rec[s]
            hard, conv, values, fin = performMinSum_Symmetric_Sparse(
                S["csr"], syn, S["llr"], maxIter=max_iter, alpha=1.0, alpha_mode="dynamical")
            r["syn"].append(pack(syn)); r["true"].append(pack(tl)); r["hard"].append(pack(hard))
            r["conv"].append(conv); r["fin"].append(fin)
            if i < n_values:
                r["values"].append(values.copy())
            det = hard
            if not conv:
                det = performOSD_enhanced(S["dense"], syn, values, hard, order=0)    # engine.py:96-97
                if len(r["osd_shot"]) < n_osd:
                    r["osd_shot"].append(i)
                    r["osd_order"].append(np.argsort(np.abs(values)).astype(np.int32))   # osd.py:11-12
                    r["osd_sol"].append(pack(det))
            dec = (S["logical"] @ det) % 2                                    # engine.py:99-100
            r["err"].append(not np.array_equal(dec, tl))
        print(f"  {name} shot {i}: conv z/x {rec['z']['conv'][-1]}/{rec['x']['conv'][-1]} err {rec['z']['err'][-1]}/{rec['x']['err'][-1]}", flush=True)
    out["ev_ptr"] = np.array(ev_ptr, dtype=np.int32)
    out["ev_loc"] = np.array(ev_loc, dtype=np.int32)
    out["ev_outcome"] = np.array(ev_out, dtype=np.int8)
    for s in "zx":
        r = rec[s]
        out[f"syn_{s}"] = np.array(r["syn"]); out[f"true_{s}"] = np.array(r["true"])
        out[f"hard_{s}"] = np.array(r["hard"]); out[f"conv_{s}"] = np.array(r["conv"])
        out[f"fin_{s}"] = np.array(r["fin"], dtype=np.int32)
        out[f"values_{s}"] = np.array(r["values"]); out[f"err_{s}"] = np.array(r["err"])
        out[f"osd_shot_{s}"] = np.array(r["osd_shot"], dtype=np.int32)
        out[f"osd_order_{s}"] = np.array(r["osd_order"], dtype=np.int32)
        out[f"osd_sol_{s}"] = np.array(r["osd_sol"])
    tag = name.split(",")[0].strip("[ ")
    np.savez_compressed(os.path.join(HERE, f"shots_{tag}.npz"), **out)


def small_kats():
    rng = np.random.default_rng(20260101)
    out = {}
    # GF(2) elimination, byte and packed (kernels.py:6-34, :36-106)
    for t, (m, n) in enumerate([(6, 10), (12, 12), (20, 70), (9, 130), (16, 5)]):
        A = rng.integers(0, 2, (m, n)).astype(np.int64)
        if t == 2:
            A[5] = A[3] ^ A[4]; A[:, 7] = 0
        b = rng.integers(0, 2, m).astype(np.int64)
        out[f"ge{t}_A"], out[f"ge{t}_b"] = A.copy(), b.copy()
        A1, b1, pr, pc = rk.gf2_elimination(A.copy(), b.copy())
        out[f"ge{t}_A_out"], out[f"ge{t}_b_out"], out[f"ge{t}_pr"], out[f"ge{t}_pc"] = A1, b1, pr, pc
        Ap, b2, pr2, pc2 = rk.gf2_elimination_packed(A.copy(), b.copy())
        out[f"ge{t}_Ap_out"], out[f"ge{t}_bp_out"], out[f"ge{t}_prp"], out[f"ge{t}_pcp"] = Ap, b2, pr2, pc2
    # small LDPC-like decoding problems for the dense entry points (dense.py:5-96) and OSD
    m, n = 18, 40
    H = np.zeros((m, n), dtype=np.int64)
    for j in range(n):
        H[rng.choice(m, size=rng.integers(1, 4), replace=False), j] = 1
    H[:, 11] = 0; H[4, :] = 0; H[7, :] = 0; H[7, 3] = 1          # empty column, empty row, degree-1 row
    out["H"] = H
    prior = np.log((1 - 0.06) / 0.06) * np.ones(n)
    prior[5] = -0.7; prior[11] = 0.0; prior[20] = 31.0
    out["prior"] = prior
    cases = []
    cfgs = [dict(alpha=1.0, alpha_mode="dynamical"), dict(alpha=0.8, alpha_mode="alvarado"),
            dict(alpha=np.array([0.4, 0.6, 0.9]), alpha_mode="alvarado-autoregressive"),
            dict(alpha=0.0, alpha_mode=None), dict(alpha=0.9, alpha_mode=None),
            dict(alpha=1.0, alpha_mode="dynamical", damping=0.7), dict(alpha=0.75, alpha_mode="alvarado", clip_llr=4.0, damping=0.5)]
    Hc = csr_matrix(H)
    for t in range(12):
        e = (rng.random(n) < 0.08).astype(np.int8)
        syn = ((H @ e) % 2).astype(np.int8)
        if t == 11:
            syn[4] ^= 1                                              # inconsistent syndrome (empty row)
        for c, cfg in enumerate(cfgs):
            for it in (1, 7):
                a = performMinSum_Symmetric(H, syn, prior, maxIter=it, **cfg)
                b = performMinSum_Symmetric_Sparse(Hc, syn, prior, maxIter=it, **cfg)
                cases.append((t, c, it))
                key = f"ms_{t}_{c}_{it}"
                out[key + "_syn"] = syn
                out[key + "_dense"] = np.concatenate([a[0].astype(np.float64), [float(a[1])], a[2], [float(a[3])]])
                out[key + "_sparse"] = np.concatenate([b[0].astype(np.float64), [float(b[1])], b[2], [float(b[3])]])
        bp = performBeliefPropagationFast(H, syn, prior, maxIter=9)     # bp_core un-jitted, see __main__
        out[f"bp_{t}"] = np.concatenate([bp[0].astype(np.float64), [float(bp[1])], bp[2], [float(bp[3])]])
        ae = performMinSum_Symmetric(H, syn, prior, maxIter=5, alpha=1.0, alpha_mode="dynamical", alpha_estimation=True)
        out[f"ae_{t}"] = ae[2]
        hard, conv, values, _ = performMinSum_Symmetric_Sparse(Hc, syn, prior, maxIter=3, alpha=1.0, alpha_mode="dynamical")
        if t != 11:
            sol = performOSD_enhanced(H.astype(np.float64), syn, values, hard, order=0)
            out[f"osd_{t}_order"] = np.argsort(np.abs(values)).astype(np.int32)
            out[f"osd_{t}_values"] = values; out[f"osd_{t}_hard"] = hard; out[f"osd_{t}_sol"] = sol
        Q = rng.normal(size=Hc.nnz) * 3
        R, Rs = rk.minsum_core_sparse(Hc.data.astype(np.float64), Hc.indices.astype(np.int32), Hc.indptr.astype(np.int32),
                                      Q, (1.0 - 2.0 * syn).astype(np.float64), 0.625, m, n)
        out[f"core_{t}_Q"], out[f"core_{t}_R"], out[f"core_{t}_Rs"] = Q, R, Rs
        out[f"sc_{t}"] = rk.syndrome_check(Hc.data.astype(np.float64), Hc.indices.astype(np.int32), Hc.indptr.astype(np.int32), e, m)
        out[f"e_{t}"] = e
    out["ms_cases"] = np.array(cases, dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "small_kats.npz"), **out)


def steane_fixture():
    """BASELINE config 1 restated as a code-capacity smoke."""
    d = np.load(os.path.join(REF, "codes", "steane.npz"))
    H = d["Hx"].astype(np.int64)
    Hc = csr_matrix(H)
    p, N = 0.005, 10000
    rng = np.random.default_rng(0)
    E = (rng.random((N, 7)) < p).astype(np.int8)
    prior = np.full(7, np.log((1 - p) / p))
    Lrow = np.array([1, 1, 1, 0, 0, 0, 0])
    nonconv = errs = 0
    fins = np.zeros(N, dtype=np.int8)
    for i in range(N):
        syn = ((H @ E[i]) % 2).astype(np.int8)
        hard, conv, values, fin = performMinSum_Symmetric_Sparse(Hc, syn, prior, maxIter=20, alpha=1.0, alpha_mode="dynamical")
        fins[i] = fin
        det = hard
        if not conv:
            nonconv += 1
            det = performOSD_enhanced(H.astype(np.float64), syn, values, hard, order=0)
        errs += int(((Lrow @ det) % 2) != ((Lrow @ E[i]) % 2))
    np.savez_compressed(os.path.join(HERE, "steane_smoke.npz"), H=H, p=p, N=N, seed=0, L=Lrow,
                        nonconverged=nonconv, logical_errors=errs, fins=fins)
    print("steane: nonconverged", nonconv, "logical errors", errs)


def alpha_fixture():
    """Alvarado alpha estimators (src/decoding/alpha.py) on the 72 code with seeded generators."""
    from src.decoding.alpha import estimate_alpha_alvarado, estimate_alpha_alvarado_autoregressive
    name, p = "[[72, 12, 6]]", 0.004
    d, bb = load_code(name)
    M = load_matrices(os.path.join(REF, "matrix_cache"), compute_cache_key(d["Hx"], d["Hz"], d["Lx"], d["Lz"], int(d["distance"]), p))
    out = {"p": p}
    for sd, H, cp in (("z", M["HdecZ"], M["channel_probsZ"]), ("x", M["HdecX"], M["channel_probsX"])):
        ll = llrs(cp)
        a, r2 = estimate_alpha_alvarado(H, p, trials=60, bins=50, rng=np.random.default_rng(5), llrs=ll)
        out[f"alv_{sd}"] = np.array([a, r2])
        av, rv = estimate_alpha_alvarado_autoregressive(H, p, maxIter=4, trials=40, bins=50, rng=np.random.default_rng(6), llrs=ll)
        out[f"auto_{sd}"] = np.array([av, rv])
        print("alpha", sd, a, r2, av, rv, flush=True)
    np.savez_compressed(os.path.join(HERE, "alpha_72.npz"), **out)


def scopt_fixture():
    """SCOPT beta estimator (src/decoding/scopt.py) on the 72 code with seeded generators (pure-Python edge loop inside:
    small trial counts)."""
    from src.decoding.scopt import estimate_scopt_beta
    name, p = "[[72, 12, 6]]", 0.004
    d, bb = load_code(name)
    M = load_matrices(os.path.join(REF, "matrix_cache"), compute_cache_key(d["Hx"], d["Hz"], d["Lx"], d["Lz"], int(d["distance"]), p))
    out = {"p": p, "trials": 150, "maxIter": 12, "bins": 30}
    for sd, H, cp in (("z", M["HdecZ"], M["channel_probsZ"]), ("x", M["HdecX"], M["channel_probsX"])):
        ll = llrs(cp)
        b, r2 = estimate_scopt_beta(H, p, trials=150, bins=30, alpha=1.0, alpha_mode="dynamical", maxIter=12,
                                    rng=np.random.default_rng(11), llrs=ll)
        out[f"dyn_{sd}"] = np.array([b, r2])
        b2, r22 = estimate_scopt_beta(H, p, trials=100, bins=30, alpha=0.8, alpha_mode="alvarado", maxIter=8,
                                      rng=np.random.default_rng(12), llrs=ll)
        out[f"alv_{sd}"] = np.array([b2, r22])
        print("scopt", sd, b, r2, b2, r22, flush=True)
    np.savez_compressed(os.path.join(HERE, "scopt_72.npz"), **out)


if __name__ == "__main__":
    # numba 0.65 cannot type np.clip on scalars inside bp_core (kernels.py:191), so the reference's
    # performBeliefPropagationFast does not compile here; run the same source un-jitted instead.
    import src.decoding.dense as _dense
    _dense.bp_core = rk.bp_core.py_func
    which = sys.argv[1:] or ["builder", "small", "steane", "72", "144", "alpha", "scopt"]
    if "alpha" in which:
        alpha_fixture()
    if "scopt" in which:
        scopt_fixture()
    if "builder" in which:
        builder_fixture()
    if "small" in which:
        small_kats()
    if "steane" in which:
        steane_fixture()
    if "72" in which:
        shots_fixture("[[72, 12, 6]]", 0.004, n_shots=48, n_values=6, n_osd=8)
    if "144" in which:
        shots_fixture("[[144, 12, 12]]", 0.005, n_shots=16, n_values=3, n_osd=6)
