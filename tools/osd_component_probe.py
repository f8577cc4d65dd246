"""How much of the OSD-0 elimination is relevant to the residual syndrome?  For failed sides of the gross code: candidates
examined, pivots found, and how many of those candidates lie in a connected component (rows + examined columns) that
contains a residual-syndrome row."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

tag, p = "144", 0.005
s = helpers.code_setup(tag)
M = helpers.matrices(tag, p)
smp = _lib.Sampler(s["ft"])
B = 64
szb, _, sxb, _, _ = smp.sample(4321, 0, B, p)
H = np.asarray(M["HdecX"]) & 1
Hc = sp.csr_matrix(H); Hcsc = sp.csc_matrix(H)
probs = np.asarray(M["channel_probsX"], dtype=np.float64)
with np.errstate(all="ignore"):
    prior = np.clip(np.nan_to_num(np.log((1 - probs) / probs)), -50, 50)      # engine.py:210-212
dec = _lib.Decoder(Hc.indptr, Hc.indices, H.shape[1], prior)
m, n = H.shape
syn = helpers.unpack(sxb.view(np.uint8), m).astype(np.int8)
hard, conv, values, fin = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC)
idx = np.nonzero(~conv)[0]
sol, rank, piv = dec.osd0(syn[idx], hard[idx], llr=values[idx], want_pivots=True)
rows_out = []
for k, i in enumerate(idx):
    order = np.argsort(np.abs(values[i].astype(np.float32)), kind="stable")
    t = int(rank[k]); last = int(piv[k][t - 1]) if t > 0 else -1
    cols = order[:last + 1]
    resid = (syn[i] ^ (H @ hard[i] % 2)).astype(np.int8)
    # bipartite graph rows (0..m-1) + examined columns (m..m+len-1)
    sub = Hcsc[:, cols]
    nc = len(cols)
    A = sp.bmat([[None, sub], [sub.T, None]], format="csr") if nc else sp.csr_matrix((m, m))
    ncomp, lab = connected_components(A, directed=False)
    good = set(lab[np.nonzero(resid)[0]])
    rel_cols = [c for c in range(nc) if lab[m + c] in good]
    pivset = set(piv[k][:t].tolist())
    rel_piv = sum(1 for c in rel_cols if c in pivset)
    rows_out.append((int(resid.sum()), nc, t, len(rel_cols), rel_piv))
a = np.array(rows_out)
print("failed sides", len(a))
print("mean residual weight %.1f, candidates examined %.1f, pivots %.1f, candidates in residual components %.1f, pivots among them %.1f" % tuple(a.mean(axis=0)))
print("medians", np.median(a, axis=0))
print("ratio relevant/examined candidates: %.3f, relevant/all pivots: %.3f" % (a[:, 3].sum() / a[:, 1].sum(), a[:, 4].sum() / a[:, 2].sum()))
