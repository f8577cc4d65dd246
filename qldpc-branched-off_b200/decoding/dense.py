"""Dense-matrix entry points on the GPU (reference ``src/decoding/dense.py``)."""
import numpy as np

from .. import _lib
from .sparse import _alpha_dispatch


def _csr_of(H):
    """CSR pattern of the dense matrix, memoised per array object (the reference engine passes the same
    ``shared_data['HdecZ']`` on every call, engine.py:90-97)."""
    return _lib.dense_csr(H)


def performMinSum_Symmetric(H, syndrome, initialBelief, maxIter=50, alpha=1.0, alpha_mode="dynamical", damping=1.0,
                            clip_llr=20.0, alpha_estimation=False):
    """Reference ``performMinSum_Symmetric`` (dense.py:5-73), including the ``alpha_estimation``
    early return (:54-56: unscaled first-iteration check messages as a dense m x n array)."""
    mode, aval, seq = _alpha_dispatch(alpha, alpha_mode)
    H = np.asarray(H)
    m, n = H.shape
    prior = np.asarray(initialBelief, dtype=np.float64)
    syndrome = np.asarray(syndrome, dtype=np.int8)
    indptr, indices = _csr_of(H)
    dec = _lib.cached_decoder_for(indptr, indices, n, prior)
    if alpha_estimation:
        if maxIter < 1:
            return np.zeros(n, dtype=np.int8), False, None, -1
        a0 = 0.5 if mode == _lib.QB_ALPHA_DYNAMIC else (float(seq[0]) if seq is not None else aval)
        R, _ = dec.minsum_core(prior[indices][None, :], (1.0 - 2.0 * syndrome.astype(np.float64))[None, :], a0)
        dense = np.zeros((m, n))
        dense[np.repeat(np.arange(m), np.diff(indptr)), indices] = R[0]
        return np.zeros(n, dtype=np.int8), False, dense / (a0 if a0 != 0 else 1.0), 0
    hard, conv, values, fin = dec.minsum(syndrome[None, :], maxIter, mode, alpha=aval, alpha_seq=seq, damping=damping,
                                         clip_llr=clip_llr, dense_variant=True)
    return hard[0], bool(conv[0]), values[0], int(fin[0])


def performBeliefPropagationFast(H, syndrome, initialBelief, maxIter=50):
    """Reference ``performBeliefPropagationFast`` (dense.py:75-96): tanh/atanh sum-product."""
    H = np.asarray(H)
    prior = np.asarray(initialBelief, dtype=np.float64)
    indptr, indices = _csr_of(H)
    dec = _lib.cached_decoder_for(indptr, indices, H.shape[1], prior)
    hard, conv, values, fin = dec.bp(np.asarray(syndrome, dtype=np.int8)[None, :], maxIter)
    return hard[0], bool(conv[0]), values[0], int(fin[0])
