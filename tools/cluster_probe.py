"""[[288,12,18]] min-sum: per-edge cluster kernel against the compressed-state kernel (time per batch, equality of outputs)."""
import os, sys, time
import numpy as np
from scipy.sparse import csr_matrix
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from helpers import code_setup, matrices, unpack
import qldpc_b200  # noqa: F401
from qldpc_b200 import _lib
from qldpc_b200.simulation.engine import llr_priors

tag = os.environ.get("QB_CODE", "288"); p = float(os.environ.get("QB_P", "0.006")); B = int(os.environ.get("QB_B", "4096"))
it = int(os.environ.get("QB_MAX_ITER", "100"))
s = code_setup(tag); M = matrices(tag, p)
smp = _lib.Sampler(s["ft"])
szb, _, sxb, _, _ = smp.sample(5, 0, B, p)
H = np.asarray(M["HdecZ"]) & 1; m, n = H.shape
Hc = csr_matrix(H); prior = llr_priors(M["channel_probsZ"])
syn = unpack(szb.view(np.uint8), m).astype(np.int8)
res = {}
for no in ("", "1"):
    if no: os.environ.pop("QLDPC_B200_CLUSTER", None)
    else: os.environ["QLDPC_B200_CLUSTER"] = "1"
    dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
    path = dec.minsum_path()
    dec.minsum(syn[:64], it, _lib.QB_ALPHA_DYNAMIC, want_values=False)
    t0 = time.perf_counter(); out = dec.minsum(syn, it, _lib.QB_ALPHA_DYNAMIC, want_values=False); t1 = time.perf_counter()
    res[no] = out
    print(f"path {path}: {B} sides, maxIter {it}: {1e3 * (t1 - t0):.1f} ms (host call incl. copies), converged {out[1].mean():.3f}", flush=True)
    dec.close()
a, b = res[""], res["1"]
print("equal hard/conv/iters:", np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]), np.array_equal(a[3], b[3]))
if hasattr(_lib.load(), "qb_debug_cluster_profile"):        # library built with QB_EXTRA_NVCC_FLAGS=-DQB_CLUSTER_PROFILE
    import ctypes as C
    out = np.zeros(256 * 32 * 8, np.uint64)
    # (profile of the LAST cluster launch in the process: rerun it)
    os.environ["QLDPC_B200_CLUSTER"] = "1"
    dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
    dec.minsum(syn, it, _lib.QB_ALPHA_DYNAMIC, want_values=False)
    print("rc", _lib.load().qb_debug_cluster_profile(out.ctypes.data_as(C.c_void_p)))
    o = out.reshape(256, 32, 8).astype(np.float64)
    o = o[o[:, 0, 7] > 0]
    per_it = o[:, :, :7] / o[:, :, 7:8]
    names = ["rows", "local sync", "columns", "wait A", "straddlers", "sync B", "check"]
    print("CTAs", o.shape[0], "cycles per iteration, mean over CTAs and warps / max over warps of the CTA mean")
    for k, nm in enumerate(names):
        print(f"  {nm:11s} {per_it[:, :, k].mean():8.0f} {per_it[:, :, k].mean(axis=0).max():8.0f}")
    print("  total", per_it.sum(axis=2).mean())
    for r in range(4):
        sel = per_it[r::4]
        print("  rank", r, " ".join(f"{sel[:, :, k].mean():7.0f}" for k in range(7)))
