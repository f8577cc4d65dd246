"""Exploratory GPU run: exercises each kernel against the oracle / golden data and prints diagnostics."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.sparse import csr_matrix
import qldpc_b200
from qldpc_b200 import _lib
from helpers import GOLDEN, code_setup, matrices, unpack
from oracle import oracle as orc

def section(name, fn):
    print(f"\n=== {name} ===", flush=True)
    t = time.time()
    try:
        fn()
        print(f"--- {name}: ok ({time.time()-t:.1f}s)", flush=True)
    except Exception:
        traceback.print_exc()
        print(f"--- {name}: FAILED", flush=True)

def events_of(g):
    return g["ev_ptr"], (g["ev_loc"].astype(np.uint32) | (g["ev_outcome"].astype(np.uint32) << 24)).astype(np.uint32)

def k2(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); s = code_setup(tag)
    smp = _lib.Sampler(s["ft"])
    ev_ptr, ev = events_of(g)
    sz, tz, sx, tx = smp.syndromes_from_events(ev_ptr, ev)
    m, k = int(g["m"]), int(g["k"])
    print("synZ equal", np.array_equal(sz, unpack(g["syn_z"], m)), "synX", np.array_equal(sx, unpack(g["syn_x"], m)),
          "trueZ", np.array_equal(tz, unpack(g["true_z"], k)), "trueX", np.array_equal(tx, unpack(g["true_x"], k)))

def minsum(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); p = float(g["p"]); M = matrices(tag, p)
    m = int(g["m"])
    for sd, H, cp in (("z", M["HdecZ"], M["channel_probsZ"]), ("x", M["HdecX"], M["channel_probsX"])):
        n = H.shape[1]; Hc = csr_matrix(H); prior = orc.llr_priors(cp)
        dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
        syn = unpack(g[f"syn_{sd}"], m).astype(np.int8)
        t = time.time(); hard, conv, values, fin = dec.minsum(syn, int(g["max_iter"]), _lib.QB_ALPHA_DYNAMIC); dt = time.time() - t
        gh = unpack(g[f"hard_{sd}"], n)
        print(sd, "hard equal shots", (hard == gh).all(axis=1).sum(), "/", len(gh), "conv equal", (conv == g[f"conv_{sd}"]).sum(),
              "fin equal", (fin == g[f"fin_{sd}"]).sum(), f"time {dt*1e3:.1f} ms")
        nv = len(g[f"values_{sd}"])
        if nv:
            ref = g[f"values_{sd}"]; mine = values[:nv]
            f = np.isfinite(ref)
            print("   values maxabs diff", np.abs(ref[f] - mine[f]).max(), "inf match", np.array_equal(np.isinf(ref), np.isinf(mine)))
        # OSD with golden ordering
        ws = g[f"osd_shot_{sd}"]
        if len(ws):
            sol, rank = dec.osd0(syn[ws], gh[ws], ordering=g[f"osd_order_{sd}"])
            gs = unpack(g[f"osd_sol_{sd}"], n)
            print("   OSD(golden order) equal", (sol == gs).all(axis=1).sum(), "/", len(ws), "pivots used", rank)
            sol2, rank2 = dec.osd0(syn[ws], gh[ws], llr=values[ws] if True else None)
            chk = dec.syndrome_check(sol2.astype(np.int8))
            print("   OSD(own sort) syndrome satisfied", (chk == syn[ws]).all(axis=1).sum(), "/", len(ws), "pivots", rank2)
        dec.close()

def small():
    g = np.load(os.path.join(GOLDEN, "small_kats.npz"))
    from qldpc_b200.decoding import kernels as K, dense as D, osd as O, sparse as S
    for t in range(5):
        A, b = g[f"ge{t}_A"].copy(), g[f"ge{t}_b"].copy()
        A1, b1, pr, pc = K.gf2_elimination(A, b)
        ok = np.array_equal(A1, g[f"ge{t}_A_out"]) and np.array_equal(b1, g[f"ge{t}_b_out"]) and np.array_equal(pr, g[f"ge{t}_pr"]) and np.array_equal(pc, g[f"ge{t}_pc"])
        Ap, b2, pr2, pc2 = K.gf2_elimination_packed(g[f"ge{t}_A"].copy(), g[f"ge{t}_b"].copy())
        ok2 = np.array_equal(Ap, g[f"ge{t}_Ap_out"]) and np.array_equal(b2, g[f"ge{t}_bp_out"]) and np.array_equal(pc2, g[f"ge{t}_pcp"])
        print("gf2", t, ok, ok2)
    H, prior = g["H"], g["prior"]; n = H.shape[1]
    cfgs = [dict(alpha=1.0, alpha_mode="dynamical"), dict(alpha=0.8, alpha_mode="alvarado"),
            dict(alpha=np.array([0.4, 0.6, 0.9]), alpha_mode="alvarado-autoregressive"),
            dict(alpha=0.0, alpha_mode=None), dict(alpha=0.9, alpha_mode=None),
            dict(alpha=1.0, alpha_mode="dynamical", damping=0.7), dict(alpha=0.75, alpha_mode="alvarado", clip_llr=4.0, damping=0.5)]
    bad = 0; worst = 0
    for t, c, it in g["ms_cases"]:
        key = f"ms_{t}_{c}_{it}"; syn = g[key + "_syn"]
        for fn, ref in ((D.performMinSum_Symmetric, g[key + "_dense"]), ):
            hard, conv, values, fin = fn(H, syn, prior, maxIter=int(it), **cfgs[c])
            rv = ref[n + 1:2 * n + 1]; f = np.isfinite(rv)
            d = np.abs(values[f] - rv[f]).max() if f.any() else 0
            worst = max(worst, d)
            if not (np.array_equal(hard, ref[:n].astype(np.int8)) and conv == bool(ref[n]) and fin == int(ref[-1]) and d < 1e-3 and np.array_equal(np.isinf(rv), np.isinf(values))):
                bad += 1; print("  mismatch", key, conv, bool(ref[n]), fin, ref[-1], d)
    print("dense minsum mismatches", bad, "worst value diff", worst)
    Hc = csr_matrix(H)
    for t in range(12):
        syn = g[f"ms_{t}_0_1_syn"]
        hard, conv, values, fin = D.performBeliefPropagationFast(H, syn, prior, maxIter=9)
        ref = g[f"bp_{t}"]; rv = ref[n + 1:2 * n + 1]
        okbp = np.array_equal(hard, ref[:n].astype(np.int8)) and conv == bool(ref[n]) and fin == int(ref[-1])
        R, Rs = K.minsum_core_sparse(None, Hc.indices, Hc.indptr, g[f"core_{t}_Q"], (1.0 - 2.0 * syn), 0.625, H.shape[0], n)
        sc = K.syndrome_check(None, Hc.indices, Hc.indptr, g[f"e_{t}"], H.shape[0])
        okosd = None
        if t != 11:
            sol = O.performOSD_enhanced(H, syn, g[f"osd_{t}_values"], g[f"osd_{t}_hard"], order=0, ordering=g[f"osd_{t}_order"])
            okosd = np.array_equal(sol, g[f"osd_{t}_sol"])
        print("case", t, "bp", okbp, np.abs(values - rv).max(), "core", np.abs(R - g[f"core_{t}_R"]).max(), np.abs(Rs - g[f"core_{t}_Rs"]).max(),
              "syncheck", np.array_equal(sc, g[f"sc_{t}"]), "osd", okosd)

def pipeline(tag):
    g = np.load(os.path.join(GOLDEN, f"shots_{tag}.npz")); s = code_setup(tag); p = float(g["p"]); M = matrices(tag, p)
    from qldpc_b200.simulation.engine import ShotEngine
    eng = ShotEngine(s["cc"], s["Lx"], s["Lz"], M, max_batch=4096)
    cfg = _lib.make_config(int(g["max_iter"]), _lib.QB_ALPHA_DYNAMIC)
    ev_ptr, ev = events_of(g)
    counts, flags, conv, fin = eng.pipeline.run_events(ev_ptr, ev, cfg, want_detail=True)
    gz, gx = g["err_z"], g["err_x"]
    print("counts", counts, "stats", eng.pipeline.stats())
    print("flags z equal", ((flags & 1) != 0) == gz, "\nflags x equal", (((flags >> 1) & 1) != 0) == gx)
    print("conv z eq", (conv[0].astype(bool) == g["conv_z"]).all(), "x", (conv[1].astype(bool) == g["conv_x"]).all(),
          "fin", (fin[0] == g["fin_z"]).all(), (fin[1] == g["fin_x"]).all())
    for n_shots in (4096, 16384):
        t = time.time(); counts, _ = eng.pipeline.run(1234, 0, n_shots, p, cfg); dt = time.time() - t
        st = eng.pipeline.stats()
        print(f"run {n_shots} shots: {dt*1e3:.1f} ms wall -> {n_shots/dt:.0f} shots/s; counts {counts}; LER {counts[2]/counts[3]:.4f}; stats {st}")
    eng.close()

def sampler_stats(tag):
    s = code_setup(tag); smp = _lib.Sampler(s["ft"])
    p = 0.005
    sz, tz, sx, tx, nf = smp.sample(99, 0, 20000, p)
    print("mean faults", nf.mean(), "expected", smp.L * p, "std", nf.std(), "expected", np.sqrt(smp.L * p * (1 - p)))
    print("mean syndrome weight Z", np.unpackbits(sz.view(np.uint8), axis=1).sum(axis=1).mean())

if __name__ == "__main__":
    print(_lib.load().qb_version(), "devices", _lib.load().qb_device_count())
    which = sys.argv[1:] or ["k2", "small", "minsum", "sampler", "pipeline"]
    if "k2" in which:
        section("K2 72", lambda: k2("72")); section("K2 144", lambda: k2("144"))
    if "small" in which: section("small KATs", small)
    if "minsum" in which:
        section("minsum+osd 72", lambda: minsum("72")); section("minsum+osd 144", lambda: minsum("144"))
    if "sampler" in which: section("sampler stats 144", lambda: sampler_stats("144"))
    if "pipeline" in which:
        section("pipeline 72", lambda: pipeline("72")); section("pipeline 144", lambda: pipeline("144"))
