"""Alvarado normalisation-factor estimation on the GPU (reference ``src/decoding/alpha.py:9-276``).

Same entry points and return values as the reference.  The error sampling (``rng.random(n) < p`` per
trial, in the reference's draw order) and the histogram / least-squares fit are host NumPy/SciPy like
the reference; all decoder work -- advancing every trial through the already estimated iterations and
producing the unscaled check messages -- runs batched in the CUDA kernel behind
``qb_alpha_messages_host`` (double precision, like the reference's pre-pass).
"""
import numpy as np
from scipy.optimize import curve_fit
from scipy.sparse import csr_matrix, isspmatrix_csr

from .. import _lib

_CHUNK = 256     # trials per device batch


def _estimate_alpha_from_samples(true_0, true_1, bins=50, plot_path=None, title=None):
    """Slope of log(f0/f1) over the message value, fitted through the origin (alpha.py:9-81)."""
    t0 = np.asarray(true_0, dtype=np.float64)
    t1 = np.asarray(true_1, dtype=np.float64)
    t0, t1 = t0[np.isfinite(t0)], t1[np.isfinite(t1)]
    if t0.size == 0 or t1.size == 0:
        raise ValueError("No finite samples for alpha estimation")
    span = (min(t0.min(), t1.min()), max(t0.max(), t1.max()))
    h0, edges = np.histogram(t0, bins=bins, range=span, density=True)
    h1, _ = np.histogram(t1, bins=bins, range=span, density=True)
    centers = 0.5 * (edges[:-1] + edges[1:])
    ok = (h0 > 0) & (h1 > 0)
    if not ok.any():
        raise ValueError("No overlapping histogram bins for alpha estimation")
    x, y = centers[ok], np.log(h0[ok] / h1[ok])
    (alpha,), _ = curve_fit(lambda lam, a: a * lam, x, y)
    fit = alpha * x
    ss_res, ss_tot = np.sum((y - fit) ** 2), np.sum((y - y.mean()) ** 2)
    r2 = 1.0 - (ss_res / ss_tot if ss_tot > 0 else np.nan)
    if plot_path is not None:
        _plot_fit(x, y, fit, r2, plot_path, title)
    return alpha, r2


def _plot_fit(x, y, fit, r2, path, title):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:      # plotting is optional; the estimate does not depend on it
        return
    plt.figure(figsize=(6, 4))
    plt.scatter(x, y, s=10, alpha=0.7, label="samples")
    plt.plot(x, fit, color="#DBA142", label=f"fit (R^2={r2:.3f})")
    plt.xlabel("Lambda"); plt.ylabel("log(f0/f1)"); plt.title(title or "Alpha estimation linear fit")
    plt.grid(True, ls="-", alpha=0.4); plt.legend(); plt.tight_layout()
    plt.savefig(path, dpi=300); plt.close()


def _setup(code, error_rate, llrs):
    if error_rate <= 0 or error_rate >= 0.5:
        raise ValueError("error_rate must be in (0, 0.5)")
    H = code if isspmatrix_csr(code) else csr_matrix(code)
    prior = np.ascontiguousarray(llrs, dtype=np.float64)
    dec = _lib.cached_decoder(H.indptr, H.indices, H.shape[1], prior)
    return H, prior, dec


def _collect(H, dec, prior, error_rate, trials, rng, alpha_prev, damping, clip_llr):
    """Messages of ``trials`` random error patterns split by the true bit of the edge's variable."""
    n = H.shape[1]
    cols = H.indices
    t0, t1 = [], []
    for lo in range(0, trials, _CHUNK):
        nb = min(_CHUNK, trials - lo)
        errors = np.empty((nb, n), dtype=np.int8)
        for t in range(nb):                                   # one draw of n uniforms per trial, like alpha.py:127 / :207
            errors[t] = rng.random(n) < error_rate
        syn = (H.dot(errors.T.astype(np.int32)) % 2).T.astype(np.int8)
        R = dec.alpha_messages(syn, prior, alpha_prev, damping, clip_llr)
        bits = errors[:, cols].astype(bool)
        for t in range(nb):                                   # keep the reference's per-trial concatenation order
            t0.append(R[t][~bits[t]]); t1.append(R[t][bits[t]])
    if not t0 or not t1:
        raise ValueError("Insufficient samples for alpha estimation")
    return np.concatenate(t0), np.concatenate(t1)


def estimate_alpha_alvarado(code, error_rate, trials=5000, bins=50, rng=None, plot_dir=None, plot_prefix=None, llrs=None):
    """Single Alvarado alpha from first-iteration message statistics (alpha.py:84-157)."""
    H, prior, dec = _setup(code, error_rate, llrs)
    rng = np.random.default_rng() if rng is None else rng
    t0, t1 = _collect(H, dec, prior, error_rate, trials, rng, np.zeros(0), 1.0, 20.0)
    plot_path = None
    if plot_dir is not None:
        plot_path = f"{plot_dir}/{plot_prefix or f'alvarado_p{error_rate:.6g}'}_alpha_fit.png"
    return _estimate_alpha_from_samples(t0, t1, bins=bins, plot_path=plot_path, title=f"Alvarado alpha fit (p={error_rate:.6g})")


def estimate_alpha_alvarado_autoregressive(code, error_rate, maxIter, trials=5000, bins=50, damping=1.0, clip_llr=20.0,
                                           rng=None, plot_dir=None, plot_prefix=None, llrs=None):
    """Per-iteration alpha sequence: alpha_k is fitted on the unscaled messages of iteration k after the
    decoder state was advanced with alpha_0..alpha_{k-1} (alpha.py:160-276)."""
    if maxIter <= 0:
        raise ValueError("maxIter must be > 0")
    H, prior, dec = _setup(code, error_rate, llrs)
    rng = np.random.default_rng() if rng is None else rng
    alphas, r2s = [], []
    for k in range(maxIter):
        t0, t1 = _collect(H, dec, prior, error_rate, trials, rng, np.asarray(alphas, dtype=np.float64), damping, clip_llr)
        plot_path = None
        if plot_dir is not None:
            plot_path = f"{plot_dir}/{plot_prefix or f'autoregressive_p{error_rate:.6g}'}_iter{k + 1}_alpha_fit.png"
        a, r2 = _estimate_alpha_from_samples(t0, t1, bins=bins, plot_path=plot_path,
                                             title=f"Autoregressive alpha fit (p={error_rate:.6g}, iter={k + 1})")
        alphas.append(float(a)); r2s.append(float(r2))
    return np.asarray(alphas, dtype=np.float64), np.asarray(r2s, dtype=np.float64)
