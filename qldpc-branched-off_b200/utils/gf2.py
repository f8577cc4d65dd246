"""Small host-side GF(2) linear algebra (setup time only; the per-shot GF(2) work is CUDA).

Used to derive logical operators of a CSS code when the reference's ``codes/*.npz`` files
(which were produced with the third-party ``qldpc`` package, reference
``generate_codes.py:131-140``) are not available.
"""
import numpy as np


def row_reduce(A):
    """Reduced row echelon form over GF(2). Returns (R, pivot_cols)."""
    A = (np.asarray(A) & 1).astype(np.uint8).copy()
    m, n = A.shape
    pivots = []
    r = 0
    for c in range(n):
        if r >= m:
            break
        rows = np.nonzero(A[r:, c])[0]
        if rows.size == 0:
            continue
        p = r + rows[0]
        if p != r:
            A[[r, p]] = A[[p, r]]
        mask = A[:, c].astype(bool)
        mask[r] = False
        A[mask] ^= A[r]
        pivots.append(c)
        r += 1
    return A[:r], pivots


def rank(A):
    return len(row_reduce(A)[1])


def nullspace(A):
    """Basis (rows) of {x : A x = 0 mod 2}."""
    A = np.asarray(A)
    n = A.shape[1]
    R, piv = row_reduce(A)
    free = [c for c in range(n) if c not in set(piv)]
    basis = np.zeros((len(free), n), dtype=np.uint8)
    for t, f in enumerate(free):
        basis[t, f] = 1
        for r, pc in enumerate(piv):
            if R[r, f]:
                basis[t, pc] = 1
    return basis


def _complement_basis(space, sub):
    """Rows of ``space`` that extend rowspace(``sub``) to rowspace(``space`` + ``sub``)."""
    cur = np.asarray(sub, dtype=np.uint8).copy()
    r = rank(cur) if cur.size else 0
    picked = []
    for v in space:
        trial = np.vstack([cur, v[None, :]])
        rt = rank(trial)
        if rt > r:
            cur, r = trial, rt
            picked.append(v)
    return np.array(picked, dtype=np.uint8).reshape(len(picked), space.shape[1])


def css_logicals(Hx, Hz):
    """Logical operators (Lx, Lz), k x n each, for the CSS code (Hx, Hz).

    Lx spans ker(Hz)/rowspace(Hx), Lz spans ker(Hx)/rowspace(Hz).  The basis is *not* the
    one ``qldpc`` picks; logical-failure flags are basis independent.
    Lz is rotated so that Lx @ Lz.T = I (symplectic pairing)."""
    Hx = np.asarray(Hx) & 1
    Hz = np.asarray(Hz) & 1
    Lx = _complement_basis(nullspace(Hz), Hx)
    Lz = _complement_basis(nullspace(Hx), Hz)
    k = Lx.shape[0]
    assert Lz.shape[0] == k
    # pair them: solve (Lx Lz^T) M = I  ->  Lz' = M^T Lz
    G = (Lx.astype(np.int64) @ Lz.T.astype(np.int64)) & 1
    aug = np.hstack([G, np.eye(k, dtype=np.int64)]).astype(np.uint8)
    R, piv = row_reduce(aug)
    assert piv[:k] == list(range(k)), "logical pairing matrix is singular"
    Ginv = R[:, k:]
    Lz = ((Ginv.T.astype(np.int64) @ Lz.astype(np.int64)) & 1).astype(np.uint8)
    return Lx.astype(np.uint8), Lz
