"""SCOPT post-decoder scaling factor on the GPU (reference ``src/decoding/scopt.py:8-176``).

Same entry point, arguments and return value as the reference.  The error sampling (one draw of n uniforms per
trial, in the reference's order) and the histogram / least-squares fit are host NumPy/SciPy like the reference; the
decoder work -- ``trials`` full min-sum decodes, each stopped at convergence -- is one batched call of the CUDA
min-sum kernel per chunk (``qb_minsum_decode_host``, float32 posteriors; the reference iterates the same recurrence in
float64 with a pure-Python loop over the edges, scopt.py:88-124).
"""
import numpy as np
from scipy.optimize import curve_fit
from scipy.sparse import csr_matrix, isspmatrix_csr

from .. import _lib

_CHUNK = 4096     # trials per device batch


def _gpu_decode(H, prior, mode, alpha, alpha_seq, max_iter, damping, clip_llr):
    dec = _lib.cached_decoder(H.indptr, H.indices, H.shape[1], prior)

    def decode(syndromes):
        hard, conv, values, fin = dec.minsum(syndromes, max_iter, mode, alpha=alpha, alpha_seq=alpha_seq,
                                             damping=damping, clip_llr=clip_llr)
        return values
    return decode


def estimate_scopt_beta(code, error_rate, trials=10000, bins=50, alpha=1.0, alpha_mode="dynamical", maxIter=50,
                        damping=1.0, clip_llr=20.0, rng=None, plot_dir=None, plot_prefix=None, llrs=None, _decode=None):
    """beta = slope through the origin of log(f1/f0) over the final posterior, f0 / f1 the densities of the posteriors
    of the error-free / flipped bits (scopt.py:127-157).  Returns ``(beta, r2)``.  ``_decode`` (tests only) replaces
    the GPU decoder by another callable ``syndromes[int8 B x m] -> posteriors[float64 B x n]``."""
    if error_rate <= 0 or error_rate >= 0.5:
        raise ValueError("error_rate must be in (0, 0.5)")
    rng = np.random.default_rng() if rng is None else rng
    H = code if isspmatrix_csr(code) else csr_matrix(code)
    if maxIter <= 0:
        raise ValueError("maxIter must be > 0")
    if alpha_mode not in {"dynamical", "alvarado", "alvarado-autoregressive"}:
        raise ValueError(f"Unsupported alpha_mode: {alpha_mode}")
    alpha_seq = None
    if alpha_mode == "alvarado-autoregressive":
        alpha_seq = np.asarray(alpha, dtype=np.float64)
        if alpha_seq.ndim != 1 or alpha_seq.size == 0:
            raise ValueError("alpha must be a non-empty 1D sequence for alvarado-autoregressive")
    if llrs is None:
        raise TypeError("llrs (the decoder's initial beliefs) is required")       # the reference indexes None here
    m, n = H.shape
    prior = np.ascontiguousarray(llrs, dtype=np.float64)
    mode = {"dynamical": _lib.QB_ALPHA_DYNAMIC, "alvarado": _lib.QB_ALPHA_FIXED,
            "alvarado-autoregressive": _lib.QB_ALPHA_SEQUENCE}[alpha_mode]
    alpha_val = 1.0 if alpha_seq is not None else float(alpha)
    decode = _decode or _gpu_decode(H, prior, mode, alpha_val, alpha_seq, maxIter, damping, clip_llr)

    final_0, final_1 = [], []
    for lo in range(0, trials, _CHUNK):
        nb = min(_CHUNK, trials - lo)
        errors = np.empty((nb, n), dtype=np.int8)
        for t in range(nb):                                   # one draw of n uniforms per trial (scopt.py:81)
            errors[t] = rng.random(n) < error_rate
        syn = (H.dot(errors.T.astype(np.int32)) % 2).T.astype(np.int8)
        values = np.asarray(decode(syn), dtype=np.float64)
        bits = errors.astype(bool)
        for t in range(nb):                                   # the reference's per-trial concatenation order
            final_0.append(values[t][~bits[t]]); final_1.append(values[t][bits[t]])
    if not final_0 or not final_1:
        raise ValueError("Insufficient samples for beta estimation")
    final_0, final_1 = np.concatenate(final_0), np.concatenate(final_1)
    final_0, final_1 = final_0[np.isfinite(final_0)], final_1[np.isfinite(final_1)]
    if final_0.size == 0 or final_1.size == 0:
        raise ValueError("No finite samples for beta estimation")

    span = (min(final_0.min(), final_1.min()), max(final_0.max(), final_1.max()))
    h0, edges = np.histogram(final_0, bins=bins, range=span, density=True)
    h1, _ = np.histogram(final_1, bins=bins, range=span, density=True)
    centers = (edges[:-1] + edges[1:]) / 2.0
    ok = (h0 > 0) & (h1 > 0)
    y, x = np.log(h1[ok] / h0[ok]), centers[ok]
    (beta,), _ = curve_fit(lambda v, b: b * v, x, y)
    fit = beta * x
    ss_res, ss_tot = np.sum((y - fit) ** 2), np.sum((y - np.mean(y)) ** 2)
    r2 = 1.0 - (ss_res / ss_tot if ss_tot > 0 else np.nan)
    if plot_dir is not None:
        _plot(x, y, fit, r2, f"{plot_dir}/{plot_prefix or f'beta_p{error_rate:.6g}'}_beta_fit.png", error_rate)
    return beta, r2


def _plot(x, y, fit, r2, path, error_rate):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:      # plotting is optional; the estimate does not depend on it
        return
    plt.figure(figsize=(6, 4))
    plt.scatter(x, y, s=10, alpha=0.7, label="samples")
    plt.plot(x, fit, color="#64B791", label=f"fit (R^2={r2:.3f})")
    plt.xlabel("LLR"); plt.ylabel("log(f1/f0)"); plt.title(f"SCOPT beta fit (p={error_rate:.6g})")
    plt.grid(True, ls="-", alpha=0.4); plt.legend(); plt.tight_layout()
    plt.savefig(path, dpi=300); plt.close()
