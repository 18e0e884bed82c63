"""n-space solvers: everything an aggregator does *after* the Gram pass.

With ``n <= ~128`` nodes and ``d`` in the 1e7..1e9 range, every distance-based
robust aggregator factors into

    pass 1   G = X X^T                      (one read of n*d, (n,n) result)
    solve    W = solve(G)                   (this module; O(poly(n)), no access to X)
    pass 2   Y = W X                        (one read of n*d)

because every quantity the algorithms need -- pairwise distances, norms,
distances to an iterate that stays in the span of the rows -- is a function of
G alone.  The functions here are the host (NumPy fp64) solvers:
they define the semantics, run the CPU path, and are the oracle for the
single-CTA CUDA solvers in ``csrc/nspace.cu``.

Behavioural parity targets (semantics only):
  krum / multi-krum  reference aggregators/geometric_wise/krum.py:177-194
  monna              reference aggregators/geometric_wise/monna.py:70-82
  cge                reference aggregators/norm_wise/comparative_gradient_elimination.py:68-79
  weiszfeld          reference aggregators/geometric_wise/geometric_median.py:79-104
  centered clipping  reference aggregators/norm_wise/center_clipping.py:131-156
  caf                reference aggregators/norm_wise/caf.py:133-184
  mda                reference aggregators/geometric_wise/minimum_diameter_average.py:328-386
  smea               reference aggregators/geometric_wise/smea.py:63-88
  nnm / arc / clip   reference pre_aggregators/nnm.py:82-97, arc.py:36-51, clipping.py:113-117
"""
from __future__ import annotations

import math
from itertools import combinations, islice
from typing import Optional, Sequence, Tuple

import numpy as np


def sqdist(G: np.ndarray) -> np.ndarray:
    """Pairwise squared distances from a Gram matrix (clamped at 0, NaN -> +inf, zero diagonal)."""
    G = np.asarray(G, dtype=np.float64)
    g = np.diag(G)
    with np.errstate(invalid="ignore", over="ignore"):
        D = g[:, None] + g[None, :] - 2.0 * G
    D = np.where(np.isnan(D), np.inf, D)
    D = np.maximum(D, 0.0)
    np.fill_diagonal(D, 0.0)
    return D


def _stable_smallest(values: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k smallest values, ties broken by lower index."""
    order = np.argsort(values, kind="stable")
    return order[:k]


# ----------------------------------------------------------------------- selection
def krum_scores(G: np.ndarray, f: int) -> np.ndarray:
    """Krum score of every row: sum of its ``n - f - 1`` smallest squared distances to the other rows."""
    D = sqdist(G)
    n = D.shape[0]
    S = np.sort(D, axis=1)
    return S[:, 1: n - f].sum(axis=1)


def krum_weights(G: np.ndarray, f: int, q: int) -> np.ndarray:
    """Multi-Krum weights: ``1/q`` on the ``q`` best-scored rows (ties to the lower index), 0 elsewhere."""
    n = G.shape[0]
    if not (0 <= f < n - 1):
        raise ValueError(f"f must satisfy 0 <= f < n-1 (got n={n}, f={f})")
    if not (1 <= q <= n - f):
        raise ValueError(f"q must satisfy 1 <= q <= n - f (got n={n}, f={f}, q={q})")
    sel = _stable_smallest(krum_scores(G, f), q)
    w = np.zeros(n)
    w[sel] = 1.0 / q
    return w


def monna_weights(G: np.ndarray, f: int, reference_index: int = 0) -> np.ndarray:
    """MoNNA weights: ``1/(n-f)`` on the ``n - f`` rows nearest to row ``reference_index`` (itself included)."""
    n = G.shape[0]
    if not (0 <= 2 * f < n):
        raise ValueError(f"2f must be < n (got n={n}, f={f})")
    if not (0 <= reference_index < n):
        raise ValueError("reference_index out of range")
    D = sqdist(G)[reference_index].copy()
    D[reference_index] = -1.0  # the trusted vector always comes first
    sel = _stable_smallest(D, n - f)
    w = np.zeros(n)
    w[sel] = 1.0 / (n - f)
    return w


def cge_weights(G: np.ndarray, f: int) -> np.ndarray:
    """CGE weights: ``1/(n-f)`` on the ``n - f`` rows with the smallest norms (NaN norms count as infinite)."""
    n = G.shape[0]
    if not (0 <= f < n):
        raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={f})")
    norms = np.sqrt(np.maximum(np.diag(np.asarray(G, dtype=np.float64)), 0.0))
    norms = np.where(np.isnan(norms), np.inf, norms)
    sel = _stable_smallest(norms, n - f)
    w = np.zeros(n)
    w[sel] = 1.0 / (n - f)
    return w


# ------------------------------------------------------------------- Gram-space iterates
def _dist_to_iterate(G: np.ndarray, a: np.ndarray) -> np.ndarray:
    """||x_i - z|| for z = sum_j a_j x_j, from the Gram matrix."""
    Ga = G @ a
    q = float(a @ Ga)
    with np.errstate(invalid="ignore"):
        d2 = np.diag(G) - 2.0 * Ga + q
    d2 = np.where(np.isnan(d2), np.inf, d2)
    return np.sqrt(np.maximum(d2, 0.0))


def weiszfeld_coeffs(G: np.ndarray, n_real: int, a0: np.ndarray, *, tol: float, max_iter: int,
                     eps: float) -> Tuple[np.ndarray, int]:
    """Weiszfeld iterations in the span of the rows.

    ``G`` is the Gram of the ``n_real`` data rows followed by optional auxiliary
    rows (e.g. the coordinate-wise median used as the starting point); ``a0``
    are the start coefficients over all rows.  Returns (coefficients, iterations).
    """
    G = np.asarray(G, dtype=np.float64)
    a = np.asarray(a0, dtype=np.float64).copy()
    it = 0
    for it in range(1, max_iter + 1):
        dist = np.maximum(_dist_to_iterate(G, a)[:n_real], eps)
        w = 1.0 / dist
        a_new = np.zeros_like(a)
        a_new[:n_real] = w / w.sum()
        delta = a_new - a
        with np.errstate(invalid="ignore"):
            step2 = float(delta @ (G @ delta))
        a = a_new
        if not (math.sqrt(max(step2, 0.0)) > tol):
            break
    return a, it


def centered_clip_coeffs(G: np.ndarray, n_real: int, a0: np.ndarray, *, c_tau: float, M: int,
                         eps: float) -> np.ndarray:
    """``v <- v + (1/n) sum_i min(1, c_tau/||x_i - v||)(x_i - v)`` for M rounds, in coefficients."""
    G = np.asarray(G, dtype=np.float64)
    a = np.asarray(a0, dtype=np.float64).copy()
    for _ in range(M):
        dist = np.maximum(_dist_to_iterate(G, a)[:n_real], eps)
        alpha = np.minimum(1.0, c_tau / dist)
        s = alpha.sum()
        a = (1.0 - s / n_real) * a
        a[:n_real] += alpha / n_real
    return a


def caf_coeffs(G: np.ndarray, n_real: int, f: int, *, power_iters: int) -> np.ndarray:
    """Covariance-bound agnostic filter in Gram space.

    ``G`` is the Gram of the n data rows followed by ONE auxiliary row r (the
    random start direction of the power iteration), so ``t = X r = G[:n, n]``
    and ``|r|^2 = G[n, n]``.  Returns the weights of the selected mean ``mu``.
    """
    G = np.asarray(G, dtype=np.float64)
    n = n_real
    if 2 * f >= n:
        raise ValueError(f"Cannot tolerate 2f >= n (got n={n}, f={f}).")
    Gx = G[:n, :n]
    t = G[:n, n]
    rnorm = math.sqrt(max(G[n, n], 0.0))
    c = np.ones(n)
    total = float(n)
    best = c / total
    best_lam = math.inf
    ones = np.ones(n)
    while total > n - 2 * f:
        p = c / total                                   # mu = sum p_i x_i
        # Gram of the centred rows y_i = x_i - mu
        Gp = Gx @ p
        Gy = Gx - Gp[:, None] - Gp[None, :] + float(p @ Gp)
        mu_r = float(p @ t)
        proj = (t - mu_r * ones) / (rnorm if rnorm > 0 else 1.0)   # Y v0
        b: Optional[np.ndarray] = None                  # vec = sum b_i y_i (None: vec = v0)
        for _ in range(power_iters):
            nb = c * proj
            nn2 = float(nb @ (Gy @ nb))
            nn = math.sqrt(max(nn2, 0.0))
            if nn <= 1e-12:
                break
            b = nb / nn
            proj = Gy @ b
        lam = float(np.sum(c * proj ** 2) / max(1e-12, c.sum()))
        if lam < best_lam:
            best_lam = lam
            best = p.copy()
        tau = proj ** 2
        tau_max = float(tau.max())
        if tau_max <= 1e-12:
            break
        c = np.clip(c * (1.0 - tau / tau_max), 0.0, None)
        total = float(c.sum())
        if total <= 0:
            break
    return best


# ---------------------------------------------------------------------- subset search
def mda_subset(D: np.ndarray, m: int) -> Tuple[int, ...]:
    """Lexicographically first m-subset minimising the maximum pairwise entry of ``D``."""
    n = D.shape[0]
    if not (1 <= m <= n):
        raise ValueError("subset size out of range")
    if m == 1:
        return (0,)
    iu = np.triu_indices(n, k=1)
    cand = np.unique(D[iu])

    def adjacency(t: float):
        adj = []
        for i in range(n):
            bits = 0
            row = D[i]
            for j in range(n):
                if j != i and row[j] <= t:
                    bits |= 1 << j
            adj.append(bits)
        return adj

    def first_clique(adj) -> Optional[Tuple[int, ...]]:
        full = (1 << n) - 1
        chosen: list = []

        def dfs(cand_mask: int) -> bool:
            need = m - len(chosen)
            if need == 0:
                return True
            while cand_mask and bin(cand_mask).count("1") >= need:
                v = (cand_mask & -cand_mask).bit_length() - 1
                cand_mask &= ~(1 << v)
                nxt = cand_mask & adj[v]          # only indices > v remain in cand_mask
                if bin(nxt).count("1") >= need - 1:
                    chosen.append(v)
                    if dfs(nxt):
                        return True
                    chosen.pop()
            return False

        return tuple(chosen) if dfs(full) else None

    lo, hi = 0, len(cand) - 1
    best = None
    while lo <= hi:
        mid = (lo + hi) // 2
        got = first_clique(adjacency(float(cand[mid])))
        if got is not None:
            best = got
            hi = mid - 1
        else:
            lo = mid + 1
    if best is None:  # non-finite distances everywhere: fall back to the first m rows
        best = tuple(range(m))
    return best


def mda_weights(G: np.ndarray, f: int) -> np.ndarray:
    """MDA weights: ``1/(n-f)`` on the lexicographically first ``(n-f)``-subset of minimum diameter."""
    n = G.shape[0]
    if not (0 <= f < n):
        raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={f})")
    sel = mda_subset(sqdist(G), n - f)
    w = np.zeros(n)
    w[list(sel)] = 1.0 / (n - f)
    return w


def smea_subset(G: np.ndarray, m: int, batch: int = 32768) -> Tuple[int, ...]:
    """First (in ``itertools.combinations`` order) m-subset whose centred covariance has the
    smallest top eigenvalue; eigenvalues come from the m x m centred Gram blocks."""
    G = np.asarray(G, dtype=np.float64)
    n = G.shape[0]
    if m <= 1:
        return tuple(range(m))
    H = np.eye(m) - np.full((m, m), 1.0 / m)
    best_val, best_combo = math.inf, None
    it = combinations(range(n), m)
    while True:
        chunk = list(islice(it, batch))
        if not chunk:
            break
        idx = np.asarray(chunk, dtype=np.int64)                       # (B, m)
        sub = G[idx[:, :, None], idx[:, None, :]]                     # (B, m, m)
        cen = H @ sub @ H
        cen = 0.5 * (cen + np.swapaxes(cen, 1, 2))
        try:
            top = np.linalg.eigvalsh(cen)[:, -1]
        except np.linalg.LinAlgError:
            top = np.full(len(chunk), np.inf)
        top = np.where(np.isnan(top), np.inf, np.maximum(top, 0.0) / m)
        k = int(np.argmin(top))
        if top[k] < best_val:
            best_val, best_combo = float(top[k]), chunk[k]
    if best_combo is None:
        best_combo = tuple(range(m))
    return tuple(best_combo)


def smea_weights(G: np.ndarray, f: int) -> np.ndarray:
    """SMEA weights: ``1/(n-f)`` on the ``(n-f)``-subset whose covariance has the smallest top eigenvalue."""
    n = G.shape[0]
    if not (0 <= 2 * f < n):
        raise ValueError(f"2f must be < n (got n={n}, f={f})")
    sel = smea_subset(G, n - f)
    w = np.zeros(n)
    w[list(sel)] = 1.0 / (n - f)
    return w


# ------------------------------------------------------------------- pre-aggregators
def clip_scales(G: np.ndarray, threshold: float) -> np.ndarray:
    """Per-row factors of static clipping: ``min(1, threshold / max(norm, 1e-12))`` from the Gram diagonal."""
    norms = np.sqrt(np.maximum(np.diag(np.asarray(G, dtype=np.float64)), 0.0))
    return np.minimum(1.0, threshold / np.maximum(norms, 1e-12))


def arc_scales(G: np.ndarray, f: int) -> np.ndarray:
    """Per-row factors of adaptive robust clipping (the ``floor(2f(n-f)/n)`` largest norms are clipped to the next one)."""
    G = np.asarray(G, dtype=np.float64)
    n = G.shape[0]
    norms = np.sqrt(np.maximum(np.diag(G), 0.0))
    nb = int(math.floor(2.0 * f / n * (n - f))) if n > 0 else 0
    nb = min(max(nb, 0), n - 1)
    if nb == 0:
        return np.ones(n)
    tau = np.sort(norms)[n - nb - 1]
    return np.minimum(1.0, tau / np.maximum(norms, 1e-12))


def nnm_matrix(G: np.ndarray, f: int) -> np.ndarray:
    """(n, n) mixing matrix: row i averages the k = n - f nearest rows of x_i (self included)."""
    n = G.shape[0]
    if not (0 <= f < n):
        raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={f})")
    D = sqdist(G).copy()
    k = n - f
    np.fill_diagonal(D, -1.0)                       # self always belongs to its own neighbourhood
    return _k_smallest_mask(D, k) * (1.0 / k)


def _k_smallest_mask(D: np.ndarray, k: int) -> np.ndarray:
    """0/1 matrix marking, per row, the k smallest entries with ties broken by lower index -- what a stable
    argsort selects, from one O(n) partition per row instead of a full sort."""
    n = D.shape[1]
    if k >= n:
        return np.ones_like(D)
    thr = np.partition(D, k - 1, axis=1)[:, k - 1: k]          # the k-th smallest value of each row
    upto = D <= thr
    if np.all(np.count_nonzero(upto, axis=1) == k):             # no tie at the threshold (the usual case)
        return upto.astype(np.float64)
    below = D < thr
    at = D == thr
    room = k - below.sum(axis=1, keepdims=True)                 # how many entries equal to it still fit
    keep_at = at & (np.cumsum(at, axis=1) <= room)
    return (below | keep_at).astype(np.float64)


def bucket_matrix(n: int, bucket_size: int, perm: Sequence[int]) -> np.ndarray:
    """(ceil(n/s), n) block-averaging matrix over the permuted rows."""
    if bucket_size <= 0:
        raise ValueError("bucket_size must be > 0")
    perm = list(perm)
    if sorted(perm) != list(range(n)):
        raise ValueError("perm must be a permutation of range(n)")
    nb = (n + bucket_size - 1) // bucket_size
    W = np.zeros((nb, n))
    for b in range(nb):
        members = perm[b * bucket_size: (b + 1) * bucket_size]
        W[b, members] = 1.0 / len(members)
    return W


__all__ = ["sqdist", "krum_scores", "krum_weights", "monna_weights", "cge_weights",
           "weiszfeld_coeffs", "centered_clip_coeffs", "caf_coeffs", "mda_subset", "mda_weights",
           "smea_subset", "smea_weights", "clip_scales", "arc_scales", "nnm_matrix",
           "bucket_matrix"]
