"""Independent fp64 CPU oracle for the multi-GPU checks.

Nothing here imports byzpy_b200: every aggregate is written out with plain torch tensor algebra on
the host in float64, following the reference's direct paths (reference
aggregators/coordinate_wise/median.py:102-106, trimmed_mean.py:110-115, mean_of_medians.py:71-81,
geometric_wise/krum.py:177-194, geometric_median.py:79-104, norm_wise/center_clipping.py:131-156,
pre_aggregators/bucketing.py:101-120, attacks/little.py:113-131).  The multi-GPU checks compare the
fused kernels against THIS, not against the repo's own CUDA operators.
"""
import math

import torch


def _stack(rows):
    return torch.stack([r.detach().double().cpu().reshape(-1) for r in rows])


def median(rows):
    return _stack(rows).median(dim=0).values           # lower median for even n (torch semantics)


def trimmed_mean(rows, f):
    X = _stack(rows).sort(dim=0).values
    n = X.shape[0]
    return X[f:n - f].mean(dim=0)


def meamed(rows, f):
    X = _stack(rows)
    n = X.shape[0]
    m = X.median(dim=0).values
    idx = (X - m).abs().argsort(dim=0, stable=True)[: n - f]
    return X.gather(0, idx).mean(dim=0)


def multikrum(rows, f, q):
    X = _stack(rows)
    n = X.shape[0]
    D = torch.cdist(X, X) ** 2
    D.fill_diagonal_(float("inf"))
    scores = D.sort(dim=1).values[:, : n - f - 1].sum(dim=1)
    pick = scores.argsort(stable=True)[:q]
    return X[pick].mean(dim=0)


def geometric_median(rows, tol=1e-6, max_iter=256, eps=1e-12, init="median"):
    X = _stack(rows)
    z = X.median(dim=0).values if init == "median" else X.mean(dim=0)
    for _ in range(max_iter):
        w = 1.0 / (X - z).norm(dim=1).clamp_min(eps)
        z_new = (w[:, None] * X).sum(dim=0) / w.sum()
        if (z_new - z).norm().item() <= tol:
            z = z_new
            break
        z = z_new
    return z


def centered_clipping(rows, c_tau, M=10, eps=1e-12, init="mean"):
    X = _stack(rows)
    v = {"mean": X.mean(dim=0), "median": X.median(dim=0).values, "zero": torch.zeros_like(X[0])}[init]
    for _ in range(M):
        diff = X - v
        scale = (c_tau / diff.norm(dim=1).clamp_min(eps)).clamp(max=1.0)
        v = v + (scale[:, None] * diff).mean(dim=0)
    return v


def bucketing(rows, bucket_size, perm):
    X = _stack(rows)[list(perm)]
    return [X[s:s + bucket_size].mean(dim=0) for s in range(0, X.shape[0], bucket_size)]


def _ndtri(p):
    # inverse normal CDF by bisection on erf (independent of the repo's Acklam approximation)
    lo, hi = -10.0, 10.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if 0.5 * (1.0 + math.erf(mid / math.sqrt(2.0))) < p:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def little(honest_rows, f, N=None):
    X = _stack(honest_rows)
    N = X.shape[0] + f if N is None else N
    s = max(1, N // 2 + 1 - f)
    z = _ndtri((N - s) / N)
    return X.mean(dim=0) + z * X.var(dim=0, unbiased=False).sqrt()


def empire(honest_rows, scale=-1.0):
    return scale * _stack(honest_rows).mean(dim=0)


def aggregate(name, rows):
    """The named configurations of tests/multi_gpu/check_*.py."""
    if name == "median":
        return median(rows)
    if name == "trmean":
        return trimmed_mean(rows, 2)
    if name == "meamed":
        return meamed(rows, 2)
    if name == "multikrum":
        return multikrum(rows, 2, 4)
    if name == "gm":
        return geometric_median(rows, tol=1e-7)
    if name == "gm_mean":
        return geometric_median(rows, init="mean")
    if name == "cclip":
        return centered_clipping(rows, 0.5, M=8)
    if name == "bucket_krum":
        return multikrum(bucketing(rows, 2, [3, 0, 6, 1, 7, 2, 5, 4]), 1, 2)
    raise KeyError(name)
