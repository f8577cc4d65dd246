"""Per-side OSD-0 cost proxies for the gross code: residual-syndrome weight (the queue's sort key) against the pivots
found and the candidates examined (position of the last pivot in the reliability order)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
import scipy.sparse as sp

tag, p = "144", 0.005
s = helpers.code_setup(tag)
M = helpers.matrices(tag, p)
smp = _lib.Sampler(s["ft"])
B = 4096
szb, _, sxb, _, _ = smp.sample(4321, 0, B, p)
H = np.asarray(M["HdecX"]) & 1
Hc = sp.csr_matrix(H)
probs = np.asarray(M["channel_probsX"], dtype=np.float64)
with np.errstate(all="ignore"):
    prior = np.clip(np.nan_to_num(np.log((1 - probs) / probs)), -50, 50)      # engine.py:210-212
dec = _lib.Decoder(Hc.indptr, Hc.indices, H.shape[1], prior)
m, n = H.shape
syn = helpers.unpack(sxb.view(np.uint8), m).astype(np.int8)
hard, conv, values, fin = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC)
idx = np.nonzero(~conv)[0]
sol, rank, piv = dec.osd0(syn[idx], hard[idx], llr=values[idx], want_pivots=True)
resid = (syn[idx] ^ (Hc.dot(hard[idx].T.astype(np.int32)).T & 1)).sum(axis=1)
last = np.array([int(piv[k][int(rank[k]) - 1]) + 1 if rank[k] > 0 else 0 for k in range(len(idx))])
print("failed sides", len(idx))
for name, a in (("residual weight", resid), ("pivots", rank), ("candidates examined", last)):
    q = np.percentile(a, [50, 90, 99, 99.9, 100])
    print(f"{name:22s} mean {a.mean():7.1f}  p50/p90/p99/p99.9/max {q}")
print("corr(weight, pivots) %.3f  corr(weight, candidates) %.3f  corr(pivots, candidates) %.3f" % (
    np.corrcoef(resid, rank)[0, 1], np.corrcoef(resid, last)[0, 1], np.corrcoef(rank, last)[0, 1]))
cost = rank.astype(np.float64) ** 2 / 64 + last            # rough: stored-column updates grow with t^2
order = np.argsort(-resid, kind="stable")
print("cost share of the last 10 %% of the queue (sorted by weight): %.3f; heaviest 1 %% of sides hold %.3f of the cost" % (
    cost[order][int(0.9 * len(order)):].sum() / cost.sum(), np.sort(cost)[::-1][:max(1, len(cost) // 100)].sum() / cost.sum()))
pos = np.empty(len(order), int); pos[order] = np.arange(len(order))
heavy = np.argsort(-cost)[:10]
print("queue positions (0 = first) of the 10 costliest sides:", pos[heavy].tolist(), "of", len(order))
print("their pivots / candidates / weight:", rank[heavy].tolist(), last[heavy].tolist(), resid[heavy].tolist())
print("sides with > 640 candidates: %.4f, > 1024: %.4f" % ((last > 640).mean(), (last > 1024).mean()))
