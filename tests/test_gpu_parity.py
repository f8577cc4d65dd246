"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libqldpc_b200.so and is compared with (a) the golden vectors produced by the real reference and
(b) the CPU oracle on the same seeded inputs."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from helpers import GOLDEN, code_setup, matrices, unpack
import qldpc_b200  # noqa: F401
from qldpc_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _events(g):
    return g["ev_ptr"], (g["ev_loc"].astype(np.uint32) | (g["ev_outcome"].astype(np.uint32) << 24)).astype(np.uint32)


def _decoder(tag, p, side):
    M = matrices(tag, p)
    H = M["HdecZ"] if side == "z" else M["HdecX"]
    prior = orc.llr_priors(M["channel_probsZ"] if side == "z" else M["channel_probsX"])
    Hc = csr_matrix(H)
    return _lib.Decoder(Hc.indptr, Hc.indices, H.shape[1], prior), Hc, H, prior


# ---- K2: syndromes / true logicals, bit-exact ------------------------------------------------------
@pytest.mark.parametrize("tag", ["72", "144"])
def test_k2_syndromes_bit_exact_vs_reference(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); s = code_setup(tag)
    m, k = int(g["m"]), int(g["k"])
    sz, tz, sx, tx = _lib.Sampler(s["ft"]).syndromes_from_events(*_events(g))
    assert np.array_equal(sz, unpack(g["syn_z"], m)) and np.array_equal(sx, unpack(g["syn_x"], m))
    assert np.array_equal(tz, unpack(g["true_z"], k)) and np.array_equal(tx, unpack(g["true_x"], k))


def test_run_trial_fast_is_a_drop_in():
    """Same np.random stream in, the reference's arrays out (simulation.py:21-107)."""
    from qldpc_b200.noise import run_trial_fast
    g = np.load(os.path.join(GOLDEN, "shots_72.npz")); s = code_setup("72")
    m, k = int(g["m"]), int(g["k"])
    for i in (0, 5, 17):
        np.random.seed(int(g["base_seed"]) + i)
        sz, tz, sx, tx = run_trial_fast(s["cc"], float(g["p"]), s["Lx"], s["Lz"])
        assert sz.dtype == np.int8 and tz.dtype == np.int8 and sz.shape == (m,) and tz.shape == (k,)
        assert np.array_equal(sz, unpack(g["syn_z"][i], m)) and np.array_equal(sx, unpack(g["syn_x"][i], m))
        assert np.array_equal(tz, unpack(g["true_z"][i], k)) and np.array_equal(tx, unpack(g["true_x"][i], k))


def test_k2_edge_cases():
    s = code_setup("72"); smp = _lib.Sampler(s["ft"])
    # empty shots, ragged event lists, and a repeated event (cancels by XOR)
    ev = np.array([5, 5, 100 | (7 << 24)], dtype=np.uint32)
    sz, tz, sx, tx = smp.syndromes_from_events(np.array([0, 0, 2, 3, 3], dtype=np.int32), ev)
    assert not sz[0].any() and not sx[0].any() and not sz[1].any() and not sx[1].any() and not sz[3].any()
    assert sz[2].any() or sx[2].any()
    # linearity: XOR of single-fault syndromes == multi-fault syndrome
    rng = np.random.default_rng(3)
    locs = rng.choice(smp.L, 40, replace=False).astype(np.uint32) | (rng.integers(0, 15, 40).astype(np.uint32) << 24)
    one = smp.syndromes_from_events(np.arange(41, dtype=np.int32), locs)
    allz = smp.syndromes_from_events(np.array([0, 40], dtype=np.int32), locs)
    for a, b in zip(one, allz):
        assert np.array_equal(np.bitwise_xor.reduce(a, axis=0), b[0])


# ---- K3: min-sum ------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["72", "144"])
def test_minsum_vs_reference_golden(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); p, m = float(g["p"]), int(g["m"])
    for sd in "zx":
        dec, Hc, H, prior = _decoder(tag, p, sd)
        n = H.shape[1]
        syn = unpack(g[f"syn_{sd}"], m).astype(np.int8)
        hard, conv, values, fin = dec.minsum(syn, int(g["max_iter"]), _lib.QB_ALPHA_DYNAMIC)
        assert np.array_equal(conv, g[f"conv_{sd}"].astype(bool))
        assert np.array_equal(fin, g[f"fin_{sd}"])
        gh = unpack(g[f"hard_{sd}"], n)
        cv = conv
        assert np.array_equal(hard[cv], gh[cv]), "converged shots must give the reference's correction"
        # non-converged shots: float32 vs float64 may flip a few near-zero posteriors
        assert (hard[~cv] != gh[~cv]).mean() < 1e-3 if (~cv).any() else True
        ref = g[f"values_{sd}"]
        if len(ref):
            mine = values[:len(ref)]
            assert np.array_equal(np.isinf(ref), np.isinf(mine)) and not np.isnan(mine).any()
            f = np.isfinite(ref)
            assert np.abs(ref[f] - mine[f]).max() < 0.25      # float32 messages, 20 chaotic iterations
        dec.close()


def test_minsum_agreement_rate_vs_oracle_72():
    """north_star bar: float32 min-sum agrees with the float64 reference on >= 99.99 % of shots
    (converged flag, iteration count and correction of converged shots), >= 1e4 sides."""
    s = code_setup("72"); p = 0.004
    smp = _lib.Sampler(s["ft"])
    B = 5000
    szb, _, sxb, _, _ = smp.sample(777, 0, B, p)
    total = agree = 0
    for sd, bits in (("z", szb), ("x", sxb)):
        dec, Hc, H, prior = _decoder("72", p, sd)
        m = H.shape[0]
        syn = unpack(bits.view(np.uint8), m).astype(np.int8)
        hard, conv, values, fin = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC, want_values=False)
        for i in range(B):
            oh, oc, ov, of = orc.performMinSum_Symmetric_Sparse(Hc, syn[i], prior, maxIter=20)
            same = (oc == conv[i]) and (of == fin[i]) and (not oc or np.array_equal(oh, hard[i]))
            agree += same; total += 1
        dec.close()
    assert total >= 10000 and agree / total >= 0.9999, (agree, total)


def test_minsum_modes_and_edge_cases():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    from qldpc_b200.decoding import dense as D, sparse as S
    H, prior = g["H"], g["prior"]; n = H.shape[1]; Hc = csr_matrix(H)
    cfgs = [dict(alpha=1.0, alpha_mode="dynamical"), dict(alpha=0.8, alpha_mode="alvarado"),
            dict(alpha=np.array([0.4, 0.6, 0.9]), alpha_mode="alvarado-autoregressive"),
            dict(alpha=0.0, alpha_mode=None), dict(alpha=0.9, alpha_mode=None),
            dict(alpha=1.0, alpha_mode="dynamical", damping=0.7),
            dict(alpha=0.75, alpha_mode="alvarado", clip_llr=4.0, damping=0.5)]
    for t, c, it in g["ms_cases"]:
        key = f"ms_{t}_{c}_{it}"; syn = g[key + "_syn"]
        for fn, Hin, ref in ((D.performMinSum_Symmetric, H, g[key + "_dense"]),
                             (S.performMinSum_Symmetric_Sparse, Hc, g[key + "_sparse"])):
            hard, conv, values, fin = fn(Hin, syn, prior, maxIter=int(it), **cfgs[c])
            assert hard.dtype == np.int8 and values.dtype == np.float64 and isinstance(conv, bool) and isinstance(fin, int)
            rv = ref[n + 1:2 * n + 1]
            assert np.array_equal(hard, ref[:n].astype(np.int8)) and conv == bool(ref[n]) and fin == int(ref[-1]), key
            assert np.array_equal(np.isinf(rv), np.isinf(values)) and not np.isnan(values).any()
            f = np.isfinite(rv)
            np.testing.assert_allclose(values[f], rv[f], rtol=1e-5, atol=1e-4)     # float32 tolerance
    for t in range(12):
        syn = g[f"ms_{t}_0_1_syn"]
        ae = D.performMinSum_Symmetric(H, syn, prior, maxIter=5, alpha_estimation=True)
        assert ae[1] is False and ae[3] == 0
        ref = g[f"ae_{t}"]
        assert np.array_equal(np.isinf(ref), np.isinf(ae[2]))
        f = np.isfinite(ref)
        np.testing.assert_allclose(ae[2][f], ref[f], rtol=0, atol=1e-12)
    # reference error behaviour
    with pytest.raises(ValueError):
        S.performMinSum_Symmetric_Sparse(Hc, g["ms_0_0_1_syn"], prior, alpha_mode="bogus")
    with pytest.raises(ValueError):
        D.performMinSum_Symmetric(H, g["ms_0_0_1_syn"], prior, alpha=-1.0, alpha_mode="alvarado")
    # ragged batches around the shots-per-CTA tile (1..9 syndromes) give the same per-shot results
    dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
    syns = np.array([g[f"ms_{t}_0_1_syn"] for t in range(9)], dtype=np.int8)
    full = dec.minsum(syns, 7, _lib.QB_ALPHA_DYNAMIC)
    for B in (1, 2, 3, 5):
        part = dec.minsum(syns[:B], 7, _lib.QB_ALPHA_DYNAMIC)
        assert np.array_equal(part[0], full[0][:B]) and np.array_equal(part[3], full[3][:B])
        assert np.array_equal(part[2], full[2][:B])
    # maxIter = 0: zero correction, not converged, final_iter -1 (kernels.py:267)
    z = dec.minsum(syns[:2], 0, _lib.QB_ALPHA_DYNAMIC)
    assert not z[0].any() and not z[1].any() and list(z[3]) == [-1, -1]
    dec.close()


def test_bp_core_and_syndrome_check_vs_reference():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    from qldpc_b200.decoding import dense as D, kernels as K
    H, prior = g["H"], g["prior"]; m, n = H.shape; Hc = csr_matrix(H)
    for t in range(12):
        syn = g[f"ms_{t}_0_1_syn"]
        hard, conv, values, fin = D.performBeliefPropagationFast(H, syn, prior, maxIter=9)
        ref = g[f"bp_{t}"]
        assert np.array_equal(hard, ref[:n].astype(np.int8)) and conv == bool(ref[n]) and fin == int(ref[-1])
        np.testing.assert_allclose(values, ref[n + 1:2 * n + 1], rtol=1e-5, atol=1e-5)   # priors are float32 on the device
        R, Rs = K.minsum_core_sparse(None, Hc.indices, Hc.indptr, g[f"core_{t}_Q"], 1.0 - 2.0 * syn, 0.625, m, n)
        for mine, ref2 in ((R, g[f"core_{t}_R"]), (Rs, g[f"core_{t}_Rs"])):
            assert np.array_equal(np.isinf(mine), np.isinf(ref2))
            f = np.isfinite(ref2)
            np.testing.assert_allclose(mine[f], ref2[f], rtol=0, atol=1e-12)
        assert np.array_equal(K.syndrome_check(None, Hc.indices, Hc.indptr, g[f"e_{t}"], m), g[f"sc_{t}"])


# ---- K5: GF(2) elimination / OSD-0, bit-exact ---------------------------------------------------------
def test_gf2_elimination_bit_exact_vs_reference():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    from qldpc_b200.decoding import kernels as K
    for t in range(5):
        A, b = g[f"ge{t}_A"].copy(), g[f"ge{t}_b"].copy()
        A1, b1, pr, pc = K.gf2_elimination(A, b)
        assert A1 is A and b1 is b, "reference mutates its inputs in place"
        assert np.array_equal(A, g[f"ge{t}_A_out"]) and np.array_equal(b, g[f"ge{t}_b_out"])
        assert np.array_equal(pr, g[f"ge{t}_pr"]) and np.array_equal(pc, g[f"ge{t}_pc"])
        A0 = g[f"ge{t}_A"].copy(); b0 = g[f"ge{t}_b"].copy()
        Ap, b2, pr2, pc2 = K.gf2_elimination_packed(A0, b0)
        assert np.array_equal(A0, g[f"ge{t}_A"]), "packed variant leaves A untouched"
        assert Ap.dtype == np.uint64 and np.array_equal(Ap, g[f"ge{t}_Ap_out"]) and np.array_equal(b2, g[f"ge{t}_bp_out"])
        assert np.array_equal(pr2, g[f"ge{t}_prp"]) and np.array_equal(pc2, g[f"ge{t}_pcp"])
    # larger random system against the oracle
    rng = np.random.default_rng(5)
    A = rng.integers(0, 2, (70, 300)).astype(np.int64); A[13] = A[2] ^ A[40]; b = rng.integers(0, 2, 70).astype(np.int64)
    Ao, bo, pro, pco = orc.gf2_elimination(A.copy(), b.copy())
    Ag, bg, prg, pcg = K.gf2_elimination(A.copy(), b.copy())
    assert np.array_equal(Ao, Ag) and np.array_equal(bo, bg) and np.array_equal(pro, prg) and np.array_equal(pco, pcg)


@pytest.mark.parametrize("tag", ["72", "144"])
def test_osd0_bit_exact_for_supplied_orderings(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); p, m = float(g["p"]), int(g["m"])
    for sd in "zx":
        dec, Hc, H, prior = _decoder(tag, p, sd)
        n = H.shape[1]
        ws = g[f"osd_shot_{sd}"]
        syn = unpack(g[f"syn_{sd}"], m).astype(np.int8)[ws]
        hard = unpack(g[f"hard_{sd}"], n)[ws]
        sol, rank = dec.osd0(syn, hard, ordering=g[f"osd_order_{sd}"])
        assert sol.dtype == np.int64
        assert np.array_equal(sol, unpack(g[f"osd_sol_{sd}"], n)), "OSD-0 must equal the reference for its own ordering"
        assert np.array_equal(dec.syndrome_check(sol.astype(np.int8)), syn)
        assert ((rank > 0) & (rank <= min(H.shape))).all(), "pivots needed: at least one, at most min(m, n)"
        dec.close()


def test_osd0_full_rank_sweep_and_stable_sort_vs_oracle():
    """Inconsistent syndromes disable the early stop (full sweep, reference pivot-row order) and
    random float32 reliabilities with many exact ties exercise the stable radix sort."""
    rng = np.random.default_rng(11)
    for (m, n, w) in ((40, 90, 3), (130, 400, 4), (33, 33, 2)):
        H = np.zeros((m, n), dtype=np.int64)
        for j in range(n):
            H[rng.choice(m, size=rng.integers(1, w + 1), replace=False), j] = 1
        H[:, 5] = 0; H[m // 2, :] = 0
        Hc = csr_matrix(H)
        dec = _lib.Decoder(Hc.indptr, Hc.indices, n, np.ones(n))
        B = 24
        llr = np.round(rng.normal(size=(B, n)) * 3, 1).astype(np.float32).astype(np.float64)    # many ties
        llr[:, 7] = np.inf; llr[0, :] = 1.0
        hard = (rng.random((B, n)) < 0.05).astype(np.int8)
        syn = rng.integers(0, 2, (B, m)).astype(np.int8)                      # generally inconsistent
        syn[B // 2:] = ((H @ (rng.random((n, B - B // 2)) < 0.1)) % 2).T       # consistent half
        sol, rank, piv = dec.osd0(syn, hard, llr=llr, want_pivots=True)
        col_ptr, row_idx = orc._csc(H)
        for i in range(B):
            order = np.argsort(np.abs(llr[i]), kind="stable")
            ref, rpiv = orc.osd0_csc(col_ptr, row_idx, m, n, syn[i], hard[i], order)
            assert np.array_equal(sol[i], ref), (m, n, i)
            assert np.array_equal(piv[i][:rank[i]], rpiv[:rank[i]])
        dec.close()


def test_perform_osd_enhanced_api():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    from qldpc_b200.decoding.osd import performOSD_enhanced
    H = g["H"]
    for t in range(11):
        syn = g[f"ms_{t}_0_1_syn"]
        sol = performOSD_enhanced(H.astype(np.float64), syn, g[f"osd_{t}_values"], g[f"osd_{t}_hard"], order=0,
                                  ordering=g[f"osd_{t}_order"])
        assert sol.dtype == np.int64 and np.array_equal(sol, g[f"osd_{t}_sol"])
        sol2 = performOSD_enhanced(H.astype(np.float64), syn, g[f"osd_{t}_values"], g[f"osd_{t}_hard"], order=2)
        assert np.array_equal((sol2 @ H.T) % 2, syn)


# ---- K1: Philox sampler ---------------------------------------------------------------------------------
def _philox4x32_10(c, k):
    c = [np.uint64(x) for x in c]; k = [np.uint64(x) for x in k]
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + np.uint64(0x9E3779B9)) & MASK, (k[1] + np.uint64(0xBB67AE85)) & MASK]
    return [int(x) for x in c]


def test_philox_known_answers_and_sampler_stream():
    assert _philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    s = code_setup("72"); smp = _lib.Sampler(s["ft"]); ft = s["ft"]
    seed, first, B, p = 0x1234567812345678, (1 << 33) + 5, 6, 0.02
    # host re-derivation of the documented stream (include/qldpc_b200.h, qb_sample_syndromes): lane l owns locations
    # [l*C, (l+1)*C); geometric jumps by inversion against the library's own table; outcome word for IDLE / CNOT faults
    T = _lib.geometric_table(p).astype(np.int64)
    K = len(T) - 1
    C = (smp.L + 31) // 32
    key = [seed & 0xFFFFFFFF, seed >> 32]
    ev_ptr, ev = [0], []
    for b in range(B):
        shot = first + b
        shot_events = []
        for lane in range(32):
            lo, hi = min(smp.L, lane * C), min(smp.L, lane * C + C)
            words, call = [], 0

            def next_word():
                nonlocal call
                if not words:
                    words.extend(_philox4x32_10([shot & 0xFFFFFFFF, shot >> 32, lane + 32 * call, 2], key)); call += 1
                return words.pop(0)

            pos = lo
            while True:
                while True:
                    r = next_word()
                    a = int(np.sum(r < T[1:]))                       # largest k with r < T[k] (T decreasing), 0 if none
                    pos += a
                    if a < K or pos >= hi:
                        break
                if pos >= hi:
                    break
                loc = pos; pos += 1
                kind = ft.loc_kind[loc]; out = 0
                if kind >= 2:
                    out = (next_word() * (3 if kind == 2 else 15)) >> 32
                shot_events.append(loc | (out << 24))
        ev += shot_events
        ev_ptr.append(len(ev))
    szb, tzb, sxb, txb, nf = smp.sample(seed, first, B, p)
    assert list(nf) == list(np.diff(ev_ptr))
    sz, tz, sx, tx = smp.syndromes_from_events(np.array(ev_ptr, np.int32), np.array(ev, np.uint32))
    m, k = smp.mZ, smp.k
    assert np.array_equal(unpack(szb.view(np.uint8), m), sz) and np.array_equal(unpack(sxb.view(np.uint8), smp.mX), sx)
    assert np.array_equal(unpack(tzb.view(np.uint8).reshape(B, 4), k), tz) and np.array_equal(unpack(txb.view(np.uint8).reshape(B, 4), k), tx)
    # counter-based: a sub-range reproduces the same shots
    sz2 = smp.sample(seed, first + 2, 3, p)[0]
    assert np.array_equal(sz2, szb[2:5])


def test_sampler_statistics_match_channel_probabilities():
    """Per-column firing frequencies of the GPU sampler vs channel_probs (builder.py:90-106)."""
    s = code_setup("72"); smp = _lib.Sampler(s["ft"]); p = 0.01; B = 200000
    szb, tzb, sxb, txb, nf = smp.sample(2024, 0, B, p)
    assert abs(nf.mean() - smp.L * p) < 5 * np.sqrt(smp.L * p * (1 - p) / B)
    # independent Bernoulli locations: the number of faults per shot is Binomial(L, p) -- its variance checks the gap
    # sampler's jumps (too regular or too bursty gaps would show here), its extremes the 32 lane chunks
    var = smp.L * p * (1 - p)
    assert abs(nf.var() - var) < 6 * var * np.sqrt(2.0 / B)
    assert nf.min() >= 0 and nf.max() < smp.L * p + 8 * np.sqrt(var)
    # detector marginals: P(bit) = (1 - prod(1 - 2 p_j)) / 2 over the columns touching the detector
    M = matrices("72", p)
    for bits, H, cp in ((szb, M["HdecZ"], M["channel_probsZ"]), (sxb, M["HdecX"], M["channel_probsX"])):
        m = H.shape[0]
        freq = unpack(bits.view(np.uint8), m).mean(axis=0)
        # channel_probs sums fault probabilities of merged faults; exact marginal uses each fault separately,
        # to first order identical: compare with tolerance of 5 sigma + second-order term
        expect = 0.5 * (1 - np.prod(np.where(H != 0, 1 - 2 * np.minimum(cp, 0.5)[None, :], 1.0), axis=1))
        sigma = np.sqrt(np.maximum(expect * (1 - expect), 1e-9) / B)
        assert (np.abs(freq - expect) < 5 * sigma + 0.02 * expect + 1e-4).all()


# ---- pipeline: end-to-end flags and LER -------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["72", "144"])
def test_pipeline_flags_vs_reference_golden(tag):
    from qldpc_b200.simulation.engine import ShotEngine
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); s = code_setup(tag); p = float(g["p"])
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], matrices(tag, p), max_batch=256)
    cfg = _lib.make_config(int(g["max_iter"]), _lib.QB_ALPHA_DYNAMIC)
    counts, flags, conv, fin = eng.pipeline.run_events(*_events(g), cfg, want_detail=True)
    assert np.array_equal(conv[0].astype(bool), g["conv_z"]) and np.array_equal(conv[1].astype(bool), g["conv_x"])
    assert np.array_equal(fin[0], g["fin_z"]) and np.array_equal(fin[1], g["fin_x"])
    ez, ex = (flags & 1) != 0, (flags & 2) != 0
    # converged shots are exact; OSD shots depend on the (unstable, float64) argsort of the reference, observed equal
    assert np.array_equal(ez[g["conv_z"]], g["err_z"][g["conv_z"]]) and np.array_equal(ex[g["conv_x"]], g["err_x"][g["conv_x"]])
    assert (ez == g["err_z"]).mean() >= 0.9 and (ex == g["err_x"]).mean() >= 0.9
    N = int(g["n_shots"])
    assert counts[3] == N and counts[0] == ez.sum() and counts[1] == ex.sum() and counts[2] == (ez | ex).sum()
    assert counts[4] == (~g["conv_z"]).sum() and counts[6] == (g["fin_z"] + 1).sum()
    # decode-only entry (host syndromes) gives the same flags
    m, k = int(g["m"]), int(g["k"])
    tz = (unpack(g["true_z"], k).astype(np.uint32) << np.arange(k, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
    tx = (unpack(g["true_x"], k).astype(np.uint32) << np.arange(k, dtype=np.uint32)).sum(axis=1).astype(np.uint32)
    c2, f2 = eng.pipeline.decode(unpack(g["syn_z"], m), tz, unpack(g["syn_x"], m), tx, cfg)
    assert np.array_equal(f2, flags) and np.array_equal(c2, counts)
    eng.close()


def test_pipeline_ler_within_reference_ci_72():
    """GPU LER (Philox sampler, 40k shots) inside the 99 % interval of the oracle's LER on 1500 shots
    driven by the reference's own RNG stream (np.random.seed(base_seed + i), engine.py:70)."""
    from qldpc_b200.simulation.engine import ShotEngine
    s = code_setup("72"); p = 0.004; M = matrices("72", p)
    m, k = M["first_logical_rowZ"], s["Lx"].shape[0]
    gz = orc.SideGraph(M["HdecZ"], M["HZ_full"][m:m + k], orc.llr_priors(M["channel_probsZ"]))
    gx = orc.SideGraph(M["HdecX"], M["HX_full"][m:m + k], orc.llr_priors(M["channel_probsX"]))
    N = 1500; errs = 0
    for i in range(N):
        np.random.seed(1234 + i)
        sz, tz, sx, tx = orc.run_trial_fast(s["cc"], p, s["Lx"], s["Lz"])
        ez = orc.decode_side(gz, sz, tz, 20)[0]; ex = orc.decode_side(gx, sx, tx, 20)[0]
        errs += int(ez or ex)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=8192)
    counts, _ = eng.pipeline.run(1234, 0, 40000, p, _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC))
    eng.close()
    ler_cpu, ler_gpu = errs / N, counts[2] / counts[3]
    half = 2.576 * np.sqrt(ler_cpu * (1 - ler_cpu) / N + ler_gpu * (1 - ler_gpu) / 40000)
    assert abs(ler_cpu - ler_gpu) < half, (ler_cpu, ler_gpu, half)


def test_pipeline_full_size_properties_gross():
    """BASELINE config 3 shape: every OSD output satisfies its syndrome, results are independent of the
    batch size, and flags are reproducible (counter-based RNG)."""
    from qldpc_b200.simulation.engine import ShotEngine
    s = code_setup("144"); p = 0.005; M = matrices("144", p)
    cfg = _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=4096)
    c1, f1 = eng.pipeline.run(42, 0, 6000, p, cfg, want_flags=True)
    eng.close()
    eng2 = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=1000)
    c2, f2 = eng2.pipeline.run(42, 0, 6000, p, cfg, want_flags=True)
    c3, f3 = eng2.pipeline.run(42, 3000, 3000, p, cfg, want_flags=True)
    eng2.close()
    assert np.array_equal(f1, f2) and np.array_equal(c1, c2) and np.array_equal(f3, f1[3000:])
    assert c1[3] == 6000 and 0.40 < c1[2] / c1[3] < 0.60          # reference LER ~0.5 at p = 0.005
    assert c1[4] > 0.8 * 6000                                       # ~95 % of sides do not converge in 20 iterations
    # OSD validity at full size through the batch API
    smp = _lib.Sampler(s["ft"])
    szb = smp.sample(42, 0, 512, p)[0]
    dec, Hc, H, prior = _decoder("144", p, "z")
    syn = unpack(szb.view(np.uint8), H.shape[0]).astype(np.int8)
    hard, conv, values, fin = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC)
    sol, rank = dec.osd0(syn[~conv], hard[~conv], llr=values[~conv])
    assert np.array_equal(dec.syndrome_check(sol.astype(np.int8)), syn[~conv])
    assert rank.max() <= 1008 and rank.mean() < 600, "early termination should need far fewer than rank(H) pivots"
    dec.close()


def test_run_simulation_api_and_early_stop():
    from qldpc_b200.simulation.engine import run_simulation
    s = code_setup("72"); p = 0.006
    M = matrices("72", p)
    res = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_cycles=6, maxIter=20, osd_order=2,
                         precomputed_matrices=M, alpha_mode="dynamical", num_workers=8, base_seed=7,
                         target_logical_errors=30, max_trials=5000, **s["bb"])
    assert set(res) == {"logical_error_rate", "z_logical_error_rate", "x_logical_error_rate", "num_trials", "logical_errors"}
    assert res["logical_errors"] == 30 and res["num_trials"] < 5000
    assert abs(res["logical_error_rate"] - 30 / res["num_trials"]) < 1e-12
    res2 = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_cycles=6, maxIter=20, precomputed_matrices=M,
                          alpha_mode="dynamical", base_seed=7, target_logical_errors=30, max_trials=5000, batch_size=300, **s["bb"])
    assert res2 == res, "early-stop cut must not depend on the batch size"
    res3 = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=2000, num_cycles=6, maxIter=20,
                          alpha_mode="alvarado", alvarado_alpha=(0.8, 0.8), base_seed=7, **s["bb"])
    assert res3["num_trials"] == 2000 and 0.2 < res3["logical_error_rate"] < 0.9
    with pytest.raises(ValueError):
        run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_cycles=6, alpha_mode="bogus", **s["bb"])


def test_steane_code_capacity_smoke():
    """BASELINE config 1 (restated): 1e4 iid-error shots on Hx of the Steane code."""
    g = np.load(os.path.join(GOLDEN, "steane_smoke.npz"))
    from qldpc_b200.decoding.sparse import performMinSum_Symmetric_Sparse_batch
    from qldpc_b200.decoding.osd import performOSD_enhanced
    H = g["H"]; Hc = csr_matrix(H); p = float(g["p"]); N = int(g["N"])
    E = (np.random.default_rng(int(g["seed"])).random((N, 7)) < p).astype(np.int8)
    syn = ((E @ H.T) % 2).astype(np.int8)
    prior = np.full(7, np.log((1 - p) / p))
    hard, conv, values, fin = performMinSum_Symmetric_Sparse_batch(Hc, syn, prior, maxIter=20)
    assert np.array_equal(fin, g["fins"].astype(np.int32))
    det = hard.astype(np.int64)
    for i in np.nonzero(~conv)[0]:
        det[i] = performOSD_enhanced(H.astype(np.float64), syn[i], values[i], hard[i], order=0)
    L = g["L"]
    errs = int((((det @ L) % 2) != ((E @ L) % 2)).sum())
    assert int((~conv).sum()) == int(g["nonconverged"]) and errs == int(g["logical_errors"])


# ---- BASELINE configs 4 and 5: other codes through the same pipeline ---------------------------------------
def _host_events(ft, B, p, seed):
    rng = np.random.default_rng(seed)
    fired = rng.random((B, ft.L)) < p
    sh, loc = np.nonzero(fired)
    kind = ft.loc_kind[loc]
    out = np.where(kind == 3, rng.integers(0, 15, len(loc)), np.where(kind == 2, rng.integers(0, 3, len(loc)), 0))
    ev_ptr = np.zeros(B + 1, dtype=np.int64); np.add.at(ev_ptr, sh + 1, 1)
    return np.cumsum(ev_ptr).astype(np.int32), (loc.astype(np.uint32) | (out.astype(np.uint32) << 24)).astype(np.uint32)


@pytest.mark.parametrize("tag,p,max_iter,B", [("90", 0.004, 20, 48), ("108", 0.006, 20, 48), ("288", 0.006, 100, 6)])
def test_other_codes_pipeline_vs_oracle(tag, p, max_iter, B):
    """Configs 4/5: [[90,8,10]], [[108,8,10]] (p sweep) and [[288,12,18]] at high max-iter + OSD: syndromes from
    host-sampled faults (K2), then per side converged / iterations exact and logical flags vs the oracle."""
    from qldpc_b200.simulation.engine import ShotEngine
    s = code_setup(tag); M = matrices(tag, p)
    ev_ptr, ev = _host_events(s["ft"], B, p, seed=int(tag))
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=64)
    cfg = _lib.make_config(max_iter, _lib.QB_ALPHA_DYNAMIC)
    counts, flags, conv, fin = eng.pipeline.run_events(ev_ptr, ev, cfg, want_detail=True)
    sz, tz, sx, tx = eng.sampler.syndromes_from_events(ev_ptr, ev)
    m, k = M["first_logical_rowZ"], s["Lx"].shape[0]
    gz = orc.SideGraph(M["HdecZ"], M["HZ_full"][m:m + k], orc.llr_priors(M["channel_probsZ"]))
    gx = orc.SideGraph(M["HdecX"], M["HX_full"][m:m + k], orc.llr_priors(M["channel_probsX"]))
    agree = tot = 0
    for i in range(B):
        for sd, g, syn, tl in ((0, gz, sz[i], tz[i]), (1, gx, sx[i], tx[i])):
            err, cv, its = orc.decode_side(g, syn, tl, max_iter)
            assert cv == bool(conv[sd][i]) and its == fin[sd][i] + 1, (tag, i, sd)
            mine = bool((flags[i] >> sd) & 1)
            if cv:
                assert mine == err
            agree += mine == err; tot += 1
    assert agree / tot >= 0.9
    assert counts[3] == B
    eng.close()


# ---- alpha estimation pre-pass ----------------------------------------------------------
def test_alpha_messages_vs_oracle_and_reference_alphas():
    """qb_alpha_messages_host against the oracle restatement of alpha.py:206-253, and the full estimators
    against the alphas the real reference produced with the same seeded generators (tests/golden/alpha_72.npz)."""
    from qldpc_b200.decoding.alpha import estimate_alpha_alvarado, estimate_alpha_alvarado_autoregressive
    g = np.load(os.path.join(GOLDEN, "alpha_72.npz")); p = float(g["p"]); M = matrices("72", p)
    for sd, H, cp in (("z", M["HdecZ"], M["channel_probsZ"]), ("x", M["HdecX"], M["channel_probsX"])):
        Hc = csr_matrix(H); prior = orc.llr_priors(cp)
        dec = _lib.cached_decoder(Hc.indptr, Hc.indices, H.shape[1], prior)
        rng = np.random.default_rng(1)
        errs = (rng.random((5, H.shape[1])) < 0.01).astype(np.int8)
        syn = ((errs @ H.T) % 2).astype(np.int8)
        for prev in ([], [0.5], [0.4, 0.7, 0.9]):
            mine = dec.alpha_messages(syn, prior, np.array(prev))
            ref = orc.alpha_messages(Hc, syn, prior, prev)
            assert np.array_equal(np.isinf(mine), np.isinf(ref))
            f = np.isfinite(ref)
            np.testing.assert_allclose(mine[f], ref[f], rtol=0, atol=1e-12)
        a, r2 = estimate_alpha_alvarado(H, p, trials=60, bins=50, rng=np.random.default_rng(5), llrs=prior)
        np.testing.assert_allclose([a, r2], g[f"alv_{sd}"], rtol=1e-6, atol=1e-8)
        av, rv = estimate_alpha_alvarado_autoregressive(H, p, maxIter=4, trials=40, bins=50, rng=np.random.default_rng(6), llrs=prior)
        np.testing.assert_allclose(av, g[f"auto_{sd}"][0], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(rv, g[f"auto_{sd}"][1], rtol=1e-6, atol=1e-8)


def test_run_simulation_alvarado_autoregressive_mode():
    """main.py's default alpha_mode (main.py:48) end to end: estimation pre-pass + decoding with the sequence."""
    from qldpc_b200.simulation.engine import run_simulation
    s = code_setup("72"); p = 0.005
    res = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=3000, num_cycles=6, maxIter=6, osd_order=2,
                         precomputed_matrices=matrices("72", p), alpha_mode="alvarado-autoregressive", base_seed=3,
                         alpha_estimation_trials=300, **s["bb"])
    assert res["num_trials"] == 3000 and 0.1 < res["logical_error_rate"] < 0.8
    assert len(res["alpha_values_z"]) == 6 and len(res["alpha_values_x"]) == 6 and len(res["alpha_r2_values_z"]) == 6
    assert all(0.1 < a < 1.5 for a in res["alpha_values_z"])
    res2 = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=1000, num_cycles=6, maxIter=6,
                          alpha_mode="alvarado", base_seed=3, alpha_estimation_trials=300, **s["bb"])
    assert 0.1 < res2["logical_error_rate"] < 0.8 and res2["alpha_r2_z"] is not None


# ---- per-edge min-sum kernel vs the compressed-state kernel and the oracle on awkward graphs ---------------------------
def test_edge_kernel_matches_compressed_state_kernel_and_oracle_on_random_graphs():
    """The per-edge kernel (minsum_edge.cu) and the compressed-state kernel (minsum.cu) implement the same float32
    recurrence: same hard decisions, convergence flags and iteration counts, posteriors equal up to the last bits
    (different but fixed summation trees are not involved: both add in row order).  Graphs cover the generic paths:
    column degree > 8, rows with more than 36 entries, empty rows and columns, degree-1 rows (+-inf messages), few
    and many distinct priors, partial slices, set_prior re-layout."""
    rng = np.random.default_rng(2024)
    cases = []
    H = (rng.random((60, 400)) < 0.06).astype(np.int8); H[5] = 0; H[:, 17] = 0
    cases.append((H, np.full(400, 2.0)))                                         # uniform prior, empty row / column
    cases.append((H, rng.choice([1.5, 2.5, -0.5, 4.0], 400)))                    # a few prior classes, one negative
    cases.append((H, rng.normal(2.5, 1.0, 400)))                                 # per-lane priors
    H2 = (rng.random((30, 120)) < 0.4).astype(np.int8)                           # row degree ~48 (> 9 chunks), column degree ~12
    cases.append((H2, np.full(120, 1.0)))
    H3 = np.zeros((40, 90), np.int8)
    for r in range(30):
        H3[r, rng.choice(80, 6, replace=False)] = 1
    for r in range(30, 40):
        H3[r, 80 + (r - 30)] = 1                                                  # degree-1 rows -> +-inf posteriors
    H3[3, 85] = 1; H3[4, 85] = 1
    cases.append((H3, np.full(90, 3.0)))
    # many low-degree rows and columns: more than two row slices per warp, several CTAs per SM, > 32 column slices for 8 warps
    m4, n4 = 2400, 9000
    H4 = np.zeros((m4, n4), np.int8)
    for j in range(n4):
        H4[rng.choice(m4, 2 + (j % 2), replace=False), j] = 1
    cases.append((H4, np.full(n4, 2.0)))
    for ci, (H, prior) in enumerate(cases):
        Hc = csr_matrix(H); m, n = H.shape
        B = 64
        e = (rng.random((B, n)) < (0.06 if n < 1000 else 0.004)).astype(np.int8)
        syn = (e.astype(np.int32) @ H.T.astype(np.int32) % 2).astype(np.int8)
        outs = []
        n_it = 140 if ci == 0 else 12          # > 128 iterations: the alpha schedule beyond the shared-memory copy
        for no_edge in ("", "1"):
            if no_edge:
                os.environ["QLDPC_B200_NO_EDGE"] = "1"
            else:
                os.environ.pop("QLDPC_B200_NO_EDGE", None)
            try:
                dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
                outs.append(dec.minsum(syn, n_it, _lib.QB_ALPHA_DYNAMIC))
                if not no_edge:      # prior change -> the layout is rebuilt
                    dec.set_prior(prior * 0.5)
                    alt = dec.minsum(syn, n_it, _lib.QB_ALPHA_DYNAMIC)
                    dec.set_prior(prior)
                    again = dec.minsum(syn, n_it, _lib.QB_ALPHA_DYNAMIC)
                    for a_, b_ in zip(again, outs[0]):
                        assert np.array_equal(a_, b_, equal_nan=True)
                    assert alt[0].shape == outs[0][0].shape
                dec.close()
            finally:
                os.environ.pop("QLDPC_B200_NO_EDGE", None)
        (h0, c0, v0, f0), (h1, c1, v1, f1) = outs
        assert np.array_equal(c0, c1) and np.array_equal(f0, f1) and np.array_equal(h0, h1)
        assert np.array_equal(np.isinf(v0), np.isinf(v1)) and np.array_equal(np.isnan(v0), np.isnan(v1))
        fin_ = np.isfinite(v0)
        np.testing.assert_allclose(v0[fin_], v1[fin_], rtol=1e-6, atol=1e-6)
        for i in range(0, B, 8):      # and the float64 oracle (hard decisions of converged shots, flags)
            oh, oc, ov, of = orc.performMinSum_Symmetric_Sparse(Hc, syn[i], prior, maxIter=n_it)
            assert oc == c0[i] and of == f0[i]
            if oc:
                assert np.array_equal(oh, h0[i])


# ---- per-edge min-sum on a thread-block cluster ([[288,12,18]]: the graph does not fit one SM) -----------------------------
@pytest.mark.parametrize("p,max_iter", [(0.001, 20), (0.003, 20), (0.006, 12)])
def test_cluster_kernel_288_matches_compressed_state_kernel_and_oracle(p, max_iter):
    """minsum_edge_cluster.cu against the compressed-state kernel (same float32 recurrence, both add in row order: equal
    hard decisions, flags, iteration counts; posteriors to the last bits) and against the float64 oracle on the sides
    both arithmetics converge on.  Low p: nearly every side converges (early exit, exact parity over the cluster)."""
    s = code_setup("288"); M = matrices("288", p)
    smp = _lib.Sampler(s["ft"])
    B = 192
    szb, _, sxb, _, _ = smp.sample(77, 0, B, p)
    smp.close()
    for sd, bits in (("Z", szb), ("X", sxb)):
        H = np.asarray(M["Hdec" + sd]) & 1; m, n = H.shape
        Hc = csr_matrix(H); prior = orc.llr_priors(M["channel_probs" + sd])
        syn = unpack(bits.view(np.uint8), m).astype(np.int8)
        outs = []
        for no_cluster in ("", "1"):
            if not no_cluster:
                os.environ["QLDPC_B200_CLUSTER"] = "1"       # opt-in: the compressed-state kernel is the (faster) default
            try:
                dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
                path = dec.minsum_path()
                assert (path == 0) if no_cluster else (path >= 2), path
                outs.append(dec.minsum(syn, max_iter, _lib.QB_ALPHA_DYNAMIC))
                if not no_cluster:       # a second call on the same handle (shot counter, cleared outputs)
                    again = dec.minsum(syn, max_iter, _lib.QB_ALPHA_DYNAMIC)
                    for a_, b_ in zip(again, outs[0]):
                        assert np.array_equal(a_, b_, equal_nan=True)
                dec.close()
            finally:
                os.environ.pop("QLDPC_B200_CLUSTER", None)
        (h0, c0, v0, f0), (h1, c1, v1, f1) = outs
        assert np.array_equal(c0, c1) and np.array_equal(f0, f1), (np.flatnonzero(c0 != c1), np.flatnonzero(f0 != f1))
        assert np.array_equal(h0, h1)
        np.testing.assert_allclose(v0, v1, rtol=1e-6, atol=1e-6)
        if p <= 0.001:
            assert c0.mean() > 0.5
        for i in range(0, B, 16):
            oh, oc, ov, of = orc.performMinSum_Symmetric_Sparse(Hc, syn[i], prior, maxIter=max_iter)
            assert oc == c0[i] and of == f0[i]
            if oc:
                assert np.array_equal(oh, h0[i])


# ---- SCOPT beta pre-pass (reference scopt.py) on the GPU decoder -----------------------------------------------------
def test_scopt_beta_vs_reference_golden_and_run_simulation():
    """The batched float32 GPU decoder behind estimate_scopt_beta gives the reference's beta for the same seeded generator
    up to the float32-vs-float64 perturbation of a histogram fit (golden: real reference, tests/golden/scopt_72.npz;
    the exact float64 host logic is checked on the CPU in test_oracle_golden.py)."""
    from qldpc_b200.decoding.scopt import estimate_scopt_beta
    g = np.load(os.path.join(GOLDEN, "scopt_72.npz"))
    p = float(g["p"])
    M = matrices("72", p)
    for sd in "zx":
        H = csr_matrix(np.asarray(M["HdecZ" if sd == "z" else "HdecX"]) & 1)
        prior = orc.llr_priors(M["channel_probsZ" if sd == "z" else "channel_probsX"])
        b, r2 = estimate_scopt_beta(H, p, trials=int(g["trials"]), bins=int(g["bins"]), alpha=1.0, alpha_mode="dynamical",
                                    maxIter=int(g["maxIter"]), rng=np.random.default_rng(11), llrs=prior)
        assert abs(b - g[f"dyn_{sd}"][0]) < 5e-3 and abs(r2 - g[f"dyn_{sd}"][1]) < 5e-2, (sd, b, r2, g[f"dyn_{sd}"])
        b, r2 = estimate_scopt_beta(H, p, trials=100, bins=30, alpha=0.8, alpha_mode="alvarado", maxIter=8,
                                    rng=np.random.default_rng(12), llrs=prior)
        assert abs(b - g[f"alv_{sd}"][0]) < 5e-3 and abs(r2 - g[f"alv_{sd}"][1]) < 5e-2, (sd, b, r2, g[f"alv_{sd}"])
    # run_simulation(scopt=True) reports the reference's extra result fields (engine.py:482-486)
    from qldpc_b200.simulation.engine import run_simulation
    s = code_setup("72")
    res = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], 0.004, num_trials=512, num_cycles=6, maxIter=10,
                         alpha_mode="dynamical", base_seed=7, scopt=True, **s["bb"])
    for k in ("beta_z", "beta_x", "beta_r2_z", "beta_r2_x", "logical_error_rate", "num_trials"):
        assert k in res
    assert res["num_trials"] == 512 and -1.0 < res["beta_z"] < 0.0 and -1.0 < res["beta_x"] < 0.0


# ---- main.py's call sequence on the GPU backend ---------------------------------------------------------
def test_main_py_call_sequence_on_the_gpu(tmp_path, monkeypatch):
    """The reference's main.py:56-151, statement by statement, through the ``src.*`` aliases installed by
    install_as_src() (the GPU box has no reference checkout, so the script itself cannot be executed there; the CPU test
    test_unmodified_reference_main_runs_up_to_the_gpu_boundary runs the real file up to the first device call, and
    tools/run_reference_main.py runs it end to end wherever both are present): code file from the generator, cache
    lookup / build / save, run_simulation with main.py's keyword arguments (alvarado-autoregressive, osd_order=2,
    target_logical_errors, max_trials), the three plotting calls and the results.npz layout of main.py:140-147."""
    import importlib
    qldpc_b200.install_as_src()
    BBCodeCircuit = importlib.import_module("src.codes.bb_code").BBCodeCircuit
    run_simulation = importlib.import_module("src.simulation.engine").run_simulation
    build_decoding_matrices = importlib.import_module("src.noise.builder").build_decoding_matrices
    plotting = importlib.import_module("src.utils.plotting")
    caching = importlib.import_module("src.utils.caching")
    from qldpc_b200.codes.generate import write_code_npz
    monkeypatch.chdir(tmp_path)
    write_code_npz("[[72, 12, 6]]", "codes")
    experiments = [{"code": "[[72, 12, 6]]", "name": "72", "physicalErrorRates": [0.006], "distance": 6}]
    target_logical_errors, max_trials, maxIter, osd_order, num_workers = 30, 20, 20, 2, 8      # main.py:41-45
    alpha_mode, scopt, cache_dir = "alvarado-autoregressive", False, "matrix_cache"
    output_dir = os.path.join("output", "run_test"); os.makedirs(output_dir)
    estimation_plot_dir = os.path.join(output_dir, "estimation_plots"); os.makedirs(estimation_plot_dir)
    results = {}
    for exp in experiments:
        data = np.load(f"codes/{exp['code']}.npz")
        Hx, Hz, Lx, Lz = data["Hx"], data["Hz"], data["Lx"], data["Lz"]
        bb_params = {k: data[k] for k in ["ell", "m", "a_x_powers", "a_y_powers", "b_y_powers", "b_x_powers"] if k in data}
        results[exp["name"]] = {}
        cb = BBCodeCircuit(Hx, Hz, num_cycles=exp["distance"], **bb_params)
        for p in exp["physicalErrorRates"]:
            key = caching.compute_cache_key(Hx, Hz, Lx, Lz, exp["distance"], p)
            matrices_ = caching.load_matrices(cache_dir, key)
            assert matrices_ is None
            matrices_ = build_decoding_matrices(cb, Lx, Lz, p, num_workers=num_workers)
            caching.save_matrices(cache_dir, key, matrices_)
            assert caching.load_matrices(cache_dir, key) is not None
            res = run_simulation(Hx, Hz, Lx, Lz, p, num_cycles=exp["distance"], maxIter=maxIter, osd_order=osd_order,
                                 precomputed_matrices=matrices_, alpha_mode=alpha_mode, num_workers=num_workers,
                                 target_logical_errors=target_logical_errors, max_trials=max_trials, scopt=scopt,
                                 estimation_plot_dir=estimation_plot_dir, alpha_estimation_trials=200, **bb_params)
            results[exp["name"]][p] = res
    res = results["72"][0.006]
    assert res["num_trials"] == max_trials and 0 <= res["logical_errors"] <= max_trials       # main.py's max_trials = 20
    assert len(res["alpha_values_z"]) == maxIter and len(res["alpha_r2_values_x"]) == maxIter
    plotting.plot_simulation_results(results, f"{output_dir}/simulation_results.png")
    plotting.plot_alpha_comparison(results, f"{output_dir}/alpha_comparison.png")
    alpha_r2_values = plotting.plot_alpha_linearity(results, f"{output_dir}/alpha_linearity.png")
    alpha_values = {c: {p: {"z": r.get("alpha_values_z"), "x": r.get("alpha_values_x")} for p, r in d.items()} for c, d in results.items()}
    np.savez(f"{output_dir}/results.npz", results=results, alpha_values=alpha_values, beta_values={}, alpha_r2_values=alpha_r2_values,
             estimation_r2_values={})
    back = np.load(f"{output_dir}/results.npz", allow_pickle=True)
    assert set(back.files) == {"results", "alpha_values", "beta_values", "alpha_r2_values", "estimation_r2_values"}
    assert back["results"].item()["72"][0.006]["num_trials"] == max_trials


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import helpers, qldpc_b200
from qldpc_b200.simulation.engine import run_simulation
rank = int(os.environ["RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
s = helpers.code_setup("72"); p = 0.006; M = helpers.matrices("72", p)
kw = dict(num_cycles=6, maxIter=20, precomputed_matrices=M, alpha_mode="dynamical", progress=False, **s["bb"])
a = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=6000, base_seed=11, batch_size=1024, **kw)
b = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, base_seed=11, target_logical_errors=40, max_trials=6000, batch_size=512, **kw)
c = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=2000, base_seed=None, batch_size=512, **kw)    # seed drawn on rank 0
out = [None] * dist.get_world_size()
dist.all_gather_object(out, (a, b, c))
if rank == 0:
    assert all(o == out[0] for o in out), out
    np.save(sys.argv[2], np.array([a["logical_errors"], a["num_trials"], b["logical_errors"], b["num_trials"]]))
dist.destroy_process_group()
"""


def test_run_simulation_two_ranks_nccl(tmp_path):
    """ADVICE (round 1): run_simulation under torchrun with NCCL -- collective tensors on each rank's own GPU, the seed
    broadcast from rank 0 when None, results identical on every rank and identical to a single-process run (shot
    streams are keyed by the global shot index), including the in-order early stop."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from qldpc_b200.simulation.engine import run_simulation
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"; script.write_text(_NCCL_WORKER)
    outfile = str(tmp_path / "out.npy")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script), root, outfile], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    two = np.load(outfile)
    s = code_setup("72"); p = 0.006; M = matrices("72", p)
    kw = dict(num_cycles=6, maxIter=20, precomputed_matrices=M, alpha_mode="dynamical", progress=False, **s["bb"])
    a = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, num_trials=6000, base_seed=11, batch_size=4096, **kw)
    b = run_simulation(s["Hx"], s["Hz"], s["Lx"], s["Lz"], p, base_seed=11, target_logical_errors=40, max_trials=6000, batch_size=700, **kw)
    assert [a["logical_errors"], a["num_trials"], b["logical_errors"], b["num_trials"]] == two.tolist()


def test_run_events_host_pipelines_several_batches():
    """qb_pipeline_run_events_host with more shots than max_batch: one upload, batches alternating between the two
    workspaces -- flags and counters equal those of batch-sized calls."""
    from qldpc_b200.simulation.engine import ShotEngine
    s = code_setup("72"); p = 0.006; M = matrices("72", p)
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=500)
    cfg = _lib.make_config(20, _lib.QB_ALPHA_DYNAMIC)
    ev_ptr, ev = _host_events(s["ft"], 1730, p, seed=5)              # 3 full batches + a ragged one
    c_all, f_all = eng.pipeline.run_events(ev_ptr, ev, cfg)
    tot = np.zeros(8, dtype=np.int64); parts = []
    for lo in range(0, 1730, 500):
        hi = min(1730, lo + 500)
        c, f = eng.pipeline.run_events(ev_ptr[lo:hi + 1] - ev_ptr[lo], ev[ev_ptr[lo]:ev_ptr[hi]], cfg)
        tot += c; parts.append(f)
    assert np.array_equal(np.concatenate(parts), f_all) and np.array_equal(tot, c_all) and c_all[3] == 1730
    with pytest.raises(ValueError):
        eng.pipeline.run_events(ev_ptr, ev, cfg, want_detail=True)    # per-shot detail: single batch only
    eng.close()
