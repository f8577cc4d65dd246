"""Min-sum agreement of the float32 kernels with the float64 recurrence on one configuration (GPU box only, test helper):
prints the rate and the kind of every disagreement."""
import os, sys
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
from scipy.sparse import csr_matrix
from oracle import oracle as orc      # test infrastructure (lives in tests/: the oracle is only ever the checker)

tag, p, B, max_iter = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 20
s = helpers.code_setup(tag); M = helpers.matrices(tag, p)
smp = _lib.Sampler(s["ft"])
szb, _, sxb, _, _ = smp.sample(4242, 0, B, p)
for sd, bits in (("Z", szb), ("X", sxb)):
    H = np.asarray(M["Hdec" + sd]) & 1; m, n = H.shape
    Hc = csr_matrix(H); prior = orc.llr_priors(M["channel_probs" + sd])
    dec = _lib.Decoder(Hc.indptr, Hc.indices, n, prior)
    syn = helpers.unpack(bits.view(np.uint8), m).astype(np.int8)
    hard, conv, vals, fin = dec.minsum(syn, max_iter, _lib.QB_ALPHA_DYNAMIC, want_values=False)
    def one(i):
        oh, oc, ov, of = orc.performMinSum_Symmetric_Sparse(Hc, syn[i], prior, maxIter=max_iter)
        same = (oc == conv[i]) and (of == fin[i]) and (not oc or np.array_equal(oh, hard[i]))
        return same, oc, of, ov
    with ThreadPoolExecutor(min(32, os.cpu_count())) as ex:
        res = list(ex.map(one, range(B)))
    bad = [i for i, r in enumerate(res) if not r[0]]
    print(tag, p, sd, "sides", B, "disagree", len(bad), "converged", int(conv.sum()))
    for i in bad[:8]:
        _, oc, of, ov = res[i]
        print("   shot", i, "gpu conv/fin", bool(conv[i]), int(fin[i]), "ref conv/fin", oc, of, "min |posterior| ref", float(np.min(np.abs(ov))), "n(|post|<1e-4)", int((np.abs(ov) < 1e-4).sum()))
