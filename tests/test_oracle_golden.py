"""CPU tests: the oracle (oracle/) and the host-side table builder against the golden vectors
generated from the real reference (tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from helpers import GOLDEN, NAMES, builder_hashes, code_setup, matrices, unpack
from oracle import oracle as orc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["72", "90", "108", "144", "288"])
def test_builder_matches_reference_cache(tag):
    """Hx/Hz from the polynomials and HdecZ/X, channel_probs, HZ/HX_full of our bit-parallel
    builder hash-equal to the reference's codes/*.npz and matrix_cache/*.npz."""
    ent = builder_hashes()[NAMES[tag]]
    s = code_setup(tag)
    assert sha(s["Hx"].astype(np.int64)) == ent["Hx"] and sha(s["Hz"].astype(np.int64)) == ent["Hz"]
    assert ent["cache"], "no cache entries recorded"
    for pkey, c in ent["cache"].items():
        M = matrices(tag, float(pkey))
        for arr in ("HdecZ", "HdecX", "channel_probsZ", "channel_probsX", "HZ_full", "HX_full"):
            assert sha(np.asarray(M[arr])) == c[arr], (tag, pkey, arr)


def test_cache_key_matches_reference():
    from qldpc_b200.utils.caching import compute_cache_key
    for tag in ("72", "144"):
        s = code_setup(tag)
        ent = builder_hashes()[NAMES[tag]]
        for pkey, c in ent["cache"].items():
            key = compute_cache_key(s["Hx"], s["Hz"], s["Lx"], s["Lz"], s["spec"]["distance"], float(pkey))
            assert key == c["key"]


@pytest.mark.parametrize("tag", ["72", "144"])
def test_oracle_shots_against_reference(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz"))
    s = code_setup(tag)
    p, N, m, k = float(g["p"]), int(g["n_shots"]), int(g["m"]), int(g["k"])
    M = matrices(tag, p)
    cc = s["cc"]
    sides = {"z": (M["HdecZ"], M["HZ_full"][m:m + k], orc.llr_priors(M["channel_probsZ"])),
             "x": (M["HdecX"], M["HX_full"][m:m + k], orc.llr_priors(M["channel_probsX"]))}
    csr = {sd: csr_matrix(v[0]) for sd, v in sides.items()}
    osd_seen = {"z": 0, "x": 0}
    for i in range(N):
        np.random.seed(int(g["base_seed"]) + i)
        sz, tz, sx, tx = orc.run_trial_fast(cc, p, s["Lx"], s["Lz"])
        for sd, syn, tl in (("z", sz, tz), ("x", sx, tx)):
            assert np.array_equal(syn, unpack(g[f"syn_{sd}"][i], m))
            assert np.array_equal(tl, unpack(g[f"true_{sd}"][i], k))
            H, Hl, prior = sides[sd]
            n = H.shape[1]
            hard, conv, values, fin = orc.performMinSum_Symmetric_Sparse(
                csr[sd], syn, prior, maxIter=int(g["max_iter"]), alpha=1.0, alpha_mode="dynamical")
            assert np.array_equal(hard, unpack(g[f"hard_{sd}"][i], n))
            assert conv == bool(g[f"conv_{sd}"][i]) and fin == int(g[f"fin_{sd}"][i])
            if i < len(g[f"values_{sd}"]):
                ref = g[f"values_{sd}"][i]
                fin_mask = np.isfinite(ref)
                assert np.array_equal(np.isinf(ref), np.isinf(values)) and not np.isnan(values).any()
                assert np.array_equal(np.sign(ref[~fin_mask]), np.sign(values[~fin_mask]))
                # the reference is numba fastmath=True: identical up to re-association (<=1e-9)
                np.testing.assert_allclose(values[fin_mask], ref[fin_mask], rtol=0, atol=1e-8)
            det = hard
            if not conv:
                w = np.nonzero(g[f"osd_shot_{sd}"] == i)[0]
                if len(w):
                    order = g[f"osd_order_{sd}"][w[0]]
                    det = orc.performOSD_enhanced(H, syn, values, hard, order=0, ordering=order)
                    assert np.array_equal(det & 1, unpack(g[f"osd_sol_{sd}"][w[0]], n))
                    osd_seen[sd] += 1
                else:
                    det = orc.performOSD_enhanced(H, syn, values, hard, order=0)
            err = not np.array_equal((Hl @ det) % 2, tl)
            assert err == bool(g[f"err_{sd}"][i]), (i, sd)
    assert osd_seen["z"] + osd_seen["x"] > 0


def test_oracle_small_kats():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    for t in range(5):
        A, b = g[f"ge{t}_A"].copy(), g[f"ge{t}_b"].copy()
        A1, b1, pr, pc = orc.gf2_elimination(np.ascontiguousarray(A), b)
        assert np.array_equal(A1, g[f"ge{t}_A_out"]) and np.array_equal(b1, g[f"ge{t}_b_out"])
        assert np.array_equal(pr, g[f"ge{t}_pr"]) and np.array_equal(pc, g[f"ge{t}_pc"])
        Ap, b2, pr2, pc2 = orc.gf2_elimination_packed(g[f"ge{t}_A"].copy(), g[f"ge{t}_b"].copy())
        assert np.array_equal(Ap, g[f"ge{t}_Ap_out"]) and np.array_equal(b2, g[f"ge{t}_bp_out"])
        assert np.array_equal(pr2, g[f"ge{t}_prp"]) and np.array_equal(pc2, g[f"ge{t}_pcp"])
    H, prior = g["H"], g["prior"]
    m, n = H.shape
    Hc = csr_matrix(H)
    cfgs = [dict(alpha=1.0, alpha_mode="dynamical"), dict(alpha=0.8, alpha_mode="alvarado"),
            dict(alpha=np.array([0.4, 0.6, 0.9]), alpha_mode="alvarado-autoregressive"),
            dict(alpha=0.0, alpha_mode=None), dict(alpha=0.9, alpha_mode=None),
            dict(alpha=1.0, alpha_mode="dynamical", damping=0.7),
            dict(alpha=0.75, alpha_mode="alvarado", clip_llr=4.0, damping=0.5)]

    def check(res, ref):
        hard, conv, values, fin = res
        assert np.array_equal(hard, ref[:n].astype(np.int8))
        assert conv == bool(ref[n]) and fin == int(ref[-1])
        rv = ref[n + 1:n + 1 + n]
        assert np.array_equal(np.isinf(rv), np.isinf(values)) and np.array_equal(np.isnan(rv), np.isnan(values))
        f = np.isfinite(rv)
        np.testing.assert_allclose(values[f], rv[f], rtol=1e-10, atol=1e-9)

    for t, c, it in g["ms_cases"]:
        key = f"ms_{t}_{c}_{it}"
        syn = g[key + "_syn"]
        check(orc.performMinSum_Symmetric(H, syn, prior, maxIter=int(it), **cfgs[c]), g[key + "_dense"])
        check(orc.performMinSum_Symmetric_Sparse(Hc, syn, prior, maxIter=int(it), **cfgs[c]), g[key + "_sparse"])
    for t in range(12):
        syn = g[f"ms_{t}_0_1_syn"]
        check(orc.performBeliefPropagationFast(H, syn, prior, maxIter=9), g[f"bp_{t}"])
        ae = orc.performMinSum_Symmetric(H, syn, prior, maxIter=5, alpha=1.0, alpha_mode="dynamical", alpha_estimation=True)
        np.testing.assert_allclose(ae[2], g[f"ae_{t}"], rtol=0, atol=1e-12)
        R, Rs = orc.minsum_core_sparse(None, Hc.indices, Hc.indptr, g[f"core_{t}_Q"],
                                       (1.0 - 2.0 * syn).astype(np.float64), 0.625, m, n)
        np.testing.assert_allclose(R, g[f"core_{t}_R"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(Rs, g[f"core_{t}_Rs"], rtol=0, atol=1e-12)
        assert np.array_equal(orc.syndrome_check(None, Hc.indices, Hc.indptr, g[f"e_{t}"], m), g[f"sc_{t}"])
        if t != 11:
            sol = orc.performOSD_enhanced(H, syn, g[f"osd_{t}_values"], g[f"osd_{t}_hard"], order=0,
                                          ordering=g[f"osd_{t}_order"])
            assert np.array_equal(sol, g[f"osd_{t}_sol"])


def test_oracle_alpha_mode_errors():
    H = csr_matrix(np.eye(3))
    with pytest.raises(ValueError):
        orc.performMinSum_Symmetric_Sparse(H, np.zeros(3), np.ones(3), alpha_mode="bogus")
    with pytest.raises(ValueError):
        orc.performMinSum_Symmetric_Sparse(H, np.zeros(3), np.ones(3), alpha=0.0, alpha_mode="alvarado")
    with pytest.raises(ValueError):
        orc.performMinSum_Symmetric_Sparse(H, np.zeros(3), np.ones(3), alpha=np.zeros((2, 2)),
                                           alpha_mode="alvarado-autoregressive")


def test_oracle_steane_smoke():
    """BASELINE config 1 (restated as code-capacity smoke)."""
    g = np.load(os.path.join(GOLDEN, "steane_smoke.npz"))
    H = g["H"]; Hc = csr_matrix(H); p = float(g["p"]); N = int(g["N"])
    E = (np.random.default_rng(int(g["seed"])).random((N, 7)) < p).astype(np.int8)
    prior = np.full(7, np.log((1 - p) / p)); L = g["L"]
    errs = nonconv = 0
    for i in range(N):
        syn = ((H @ E[i]) % 2).astype(np.int8)
        hard, conv, values, fin = orc.performMinSum_Symmetric_Sparse(Hc, syn, prior, maxIter=20)
        assert fin == int(g["fins"][i])
        det = hard
        if not conv:
            nonconv += 1
            det = orc.performOSD_enhanced(H, syn, values, hard, order=0)
        errs += int(((L @ det) % 2) != ((L @ E[i]) % 2))
    assert nonconv == int(g["nonconverged"]) and errs == int(g["logical_errors"])


def test_oracle_alpha_messages_reproduce_reference_alphas():
    """The oracle's message collection (alpha.py:206-253 restated) + the host fit reproduce the alphas the real
    reference computed with the same seeded generators."""
    from qldpc_b200.decoding.alpha import _estimate_alpha_from_samples
    g = np.load(os.path.join(GOLDEN, "alpha_72.npz")); p = float(g["p"]); M = matrices("72", p)
    H = M["HdecZ"]; Hc = csr_matrix(H); prior = orc.llr_priors(M["channel_probsZ"])
    rng = np.random.default_rng(5)
    t0, t1 = [], []
    for _ in range(60):
        e = (rng.random(H.shape[1]) < p).astype(np.int8)
        syn = (Hc.dot(e) % 2).astype(np.int8)
        R = orc.alpha_messages(Hc, syn[None, :], prior, [])[0]
        bits = e[Hc.indices].astype(bool)
        t0.append(R[~bits]); t1.append(R[bits])
    a, r2 = _estimate_alpha_from_samples(np.concatenate(t0), np.concatenate(t1), bins=50)
    np.testing.assert_allclose([a, r2], g["alv_z"], rtol=1e-7, atol=1e-9)


def test_scopt_beta_host_logic_vs_reference_golden():
    """estimate_scopt_beta (reference scopt.py:8-176): sampling order, sample split, histogram and fit of the package's
    implementation, with the float64 oracle standing in for the GPU decoder, reproduce the real reference's (beta, r2)
    for the same seeded generators (tests/golden/scopt_72.npz, made by make_golden.py scopt)."""
    import qldpc_b200  # noqa: F401
    from qldpc_b200.decoding.scopt import estimate_scopt_beta
    from scipy.sparse import csr_matrix
    g = np.load(os.path.join(GOLDEN, "scopt_72.npz"))
    p = float(g["p"])
    M = matrices("72", p)
    for sd in "zx":
        H = csr_matrix(np.asarray(M["HdecZ" if sd == "z" else "HdecX"]) & 1)
        prior = orc.llr_priors(M["channel_probsZ" if sd == "z" else "channel_probsX"])

        def decoder(max_iter, **kw):
            def decode(syn):
                return np.stack([orc.performMinSum_Symmetric_Sparse(H, s_, prior, maxIter=max_iter, **kw)[2] for s_ in syn])
            return decode
        b, r2 = estimate_scopt_beta(H, p, trials=int(g["trials"]), bins=int(g["bins"]), alpha=1.0, alpha_mode="dynamical",
                                    maxIter=int(g["maxIter"]), rng=np.random.default_rng(11), llrs=prior,
                                    _decode=decoder(int(g["maxIter"]), alpha=1.0, alpha_mode="dynamical"))
        np.testing.assert_allclose([b, r2], g[f"dyn_{sd}"], rtol=1e-6)
        b, r2 = estimate_scopt_beta(H, p, trials=100, bins=30, alpha=0.8, alpha_mode="alvarado", maxIter=8,
                                    rng=np.random.default_rng(12), llrs=prior, _decode=decoder(8, alpha=0.8, alpha_mode="alvarado"))
        np.testing.assert_allclose([b, r2], g[f"alv_{sd}"], rtol=1e-6)
    with pytest.raises(ValueError):
        estimate_scopt_beta(H, 0.7, llrs=prior)
    with pytest.raises(ValueError):
        estimate_scopt_beta(H, p, alpha_mode="nope", llrs=prior)
    with pytest.raises(ValueError):
        estimate_scopt_beta(H, p, maxIter=0, llrs=prior)
