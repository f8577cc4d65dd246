"""Per-warp phase timing of minsum_edge_kernel (needs a library built with QB_EXTRA_NVCC_FLAGS=-DQB_EDGE_PROFILE)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers
import qldpc_b200
from qldpc_b200 import _lib
import scipy.sparse as sp
tag, p, side = "144", 0.005, "X"
mats = helpers.matrices(tag, p)
H = sp.csr_matrix(np.asarray(mats["Hdec" + side]) & 1); H.sort_indices()
probs = np.asarray(mats["channel_probs" + side], dtype=np.float64)
with np.errstate(all="ignore"):
    prior = np.clip(np.nan_to_num(np.log((1 - probs) / probs)), -50, 50)
dec = _lib.Decoder(H.indptr, H.indices, H.shape[1], prior)
rng = np.random.default_rng(1)
B = 148 * 8
e = (rng.random((B, H.shape[1])) < probs[None, :] * 1.0).astype(np.int8)
syn = (H.dot(e.T).T & 1).astype(np.int8)
for _ in range(2):
    hard, conv, values, fin = dec.minsum(syn, 20, _lib.QB_ALPHA_DYNAMIC)
lib = _lib.load()
out = np.zeros(256 * 32 * 4, np.uint64)
print("rc", lib.qb_debug_edge_profile(out.ctypes.data_as(C.c_void_p)), "conv frac", conv.mean())
o = out.reshape(256, 32, 4)[:148].astype(np.float64)
per_it = o / (8 * 20)   # shots per CTA x iterations (roughly; converged shots stop early)
m = per_it.mean(axis=0)
print("per iteration cycles, mean over CTAs; columns: A work, wait1, B work, wait2")
for w in range(32):
    print(w, np.round(m[w]).astype(int))
print("mean over warps", np.round(m.mean(axis=0)).astype(int), "sum", int(m.mean(axis=0).sum()))
print("max A work", int(m[:, 0].max()), "max B work", int(m[:, 2].max()))
