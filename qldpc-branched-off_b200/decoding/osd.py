"""``performOSD_enhanced`` on the GPU (reference ``src/decoding/osd.py:5-77``)."""
from itertools import combinations

import numpy as np

from .. import _lib
from .dense import _csr_of
from .kernels import compute_metric, gf2_elimination_packed, recompute_solution


def performOSD_enhanced(H, syndrome, llr, hard, order=0, max_combinations=None, ordering=None):
    """OSD decoder, same signature as the reference plus an optional explicit column ``ordering``.

    OSD-0 (osd.py:5-29) runs in the CUDA kernel.  For a syndrome inside the column space of H --
    always the case for simulated shots -- the reference returns the OSD-0 solution regardless of
    ``order`` (osd.py:27-29).  Only for an inconsistent syndrome with ``order > 0`` does it search
    weight-<=order flips (osd.py:31-77); that cold path is reproduced here on the host around the
    GPU Gauss-Jordan so the entry point stays a complete drop-in.
    The default ordering is the stable argsort of |llr| as float32 (ties by column index)."""
    H = np.asarray(H)
    m, n = H.shape
    syndrome = np.asarray(syndrome)
    hard = np.asarray(hard)
    indptr, indices = _csr_of(H)
    dec = _lib.cached_decoder_for(indptr, indices, n)          # OSD-0 never reads the priors: reuse the handle as is
    sol, _ = dec.osd0((syndrome.astype(np.int64) & 1).astype(np.int8)[None, :], hard[None, :],
                      llr=None if ordering is not None else np.asarray(llr, dtype=np.float64)[None, :],
                      ordering=None if ordering is None else np.asarray(ordering)[None, :])
    osd0_solution = sol[0]
    if order == 0:
        return osd0_solution
    osd0_syndrome = dec.syndrome_check(osd0_solution.astype(np.int8)[None, :])[0]
    if np.all(osd0_syndrome == syndrome):
        return osd0_solution
    return _higher_order_search(H, syndrome, np.asarray(llr, dtype=np.float64), hard, order, max_combinations,
                                ordering, osd0_solution, osd0_syndrome)


def _higher_order_search(H, syndrome, llr, hard, order, max_combinations, ordering, osd0_solution, osd0_syndrome):
    """Cold path of osd.py:31-77 (inconsistent syndrome, order > 0)."""
    m, n = H.shape
    llr_abs = np.abs(llr)
    if ordering is None:
        ordering = np.argsort(llr_abs.astype(np.float32), kind="stable")
    ordering = np.asarray(ordering, dtype=np.int64)
    Hp = (H[:, ordering] != 0).astype(np.int64)
    residual = ((syndrome + (hard @ H.T)) % 2).astype(np.int64)
    _, s_red, prow, pcol = gf2_elimination_packed(Hp.copy(), residual)
    e_perm = np.zeros(n, dtype=np.int64)
    e_perm[pcol] = s_red[prow]
    free = np.setdiff1d(np.arange(n), pcol)
    if free.size == 0:
        return osd0_solution
    free = free[np.argsort(llr_abs[ordering[free]])]
    tests = free[:min(free.size, order + 10)]
    best, best_metric = osd0_solution.copy(), compute_metric(osd0_solution.astype(np.float64), llr_abs,
                                                            int(np.sum(osd0_syndrome != syndrome)))
    found_valid, tested = False, 0
    for w in range(1, min(order + 1, len(tests) + 1)):
        for flips in combinations(tests, w):
            if max_combinations and tested >= max_combinations:
                return best
            trial = e_perm.copy()
            trial[list(flips)] ^= 1
            full = recompute_solution(Hp, s_red, trial, prow, pcol)
            corr = np.zeros(n, dtype=np.int64)
            corr[ordering] = full
            cand = (hard + corr) % 2
            csyn = (cand @ H.T) % 2
            valid = bool(np.all(csyn == syndrome))
            if valid:
                metric = compute_metric(cand.astype(np.float64), llr_abs, 0)
                if not found_valid or metric < best_metric:
                    best, best_metric, found_valid = cand.copy(), metric, True
            elif not found_valid:
                metric = compute_metric(cand.astype(np.float64), llr_abs, int(np.sum(csyn != syndrome)))
                if metric < best_metric:
                    best, best_metric = cand.copy(), metric
            tested += 1
    return best
