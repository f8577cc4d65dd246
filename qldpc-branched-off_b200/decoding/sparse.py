"""``performMinSum_Symmetric_Sparse`` on the GPU (reference ``src/decoding/sparse.py:5-55``)."""
import numpy as np

from .. import _lib


def _alpha_dispatch(alpha, alpha_mode):
    """alpha_mode validation of sparse.py:18-39 / dense.py:19-33 -> (QB mode, alpha, sequence)."""
    if alpha_mode is None:
        return (_lib.QB_ALPHA_DYNAMIC if alpha == 0 else _lib.QB_ALPHA_FIXED), float(alpha), None
    if alpha_mode == "dynamical":
        return _lib.QB_ALPHA_DYNAMIC, 1.0, None
    if alpha_mode == "alvarado":
        if alpha <= 0:
            raise ValueError("alpha must be > 0 when alpha_mode='alvarado'")
        return _lib.QB_ALPHA_FIXED, float(alpha), None
    if alpha_mode == "alvarado-autoregressive":
        seq = np.asarray(alpha, dtype=np.float64)
        if seq.ndim != 1 or seq.size == 0:
            raise ValueError("alpha must be a non-empty 1D sequence for alvarado-autoregressive")
        return _lib.QB_ALPHA_SEQUENCE, 0.0, seq
    raise ValueError(f"Unsupported alpha_mode: {alpha_mode}")


def performMinSum_Symmetric_Sparse(H_csr, syndrome, initialBelief, maxIter=100, alpha=1.0, alpha_mode="dynamical",
                                   damping=1.0, clip_llr=20.0):
    """Same signature and return tuple as the reference: (candidateError int8[n], converged bool,
    values float64[n], final_iter int).  Messages are float32 on the device."""
    mode, aval, seq = _alpha_dispatch(alpha, alpha_mode)
    prior = np.asarray(initialBelief, dtype=np.float64)
    dec = _lib.cached_decoder_for(H_csr.indptr, H_csr.indices, H_csr.shape[1], prior)
    hard, conv, values, fin = dec.minsum(np.asarray(syndrome, dtype=np.int8)[None, :], maxIter, mode, alpha=aval,
                                         alpha_seq=seq, damping=damping, clip_llr=clip_llr)
    return hard[0], bool(conv[0]), values[0], int(fin[0])


def performMinSum_Symmetric_Sparse_batch(H_csr, syndromes, initialBelief, maxIter=100, alpha=1.0,
                                         alpha_mode="dynamical", damping=1.0, clip_llr=20.0):
    """Batched form: syndromes int8 [B, m] -> (hard int8 [B, n], converged bool [B], values f64 [B, n], final_iter [B])."""
    mode, aval, seq = _alpha_dispatch(alpha, alpha_mode)
    prior = np.asarray(initialBelief, dtype=np.float64)
    dec = _lib.cached_decoder_for(H_csr.indptr, H_csr.indices, H_csr.shape[1], prior)
    return dec.minsum(syndromes, maxIter, mode, alpha=aval, alpha_seq=seq, damping=damping, clip_llr=clip_llr)
