"""Plain PyTorch implementations of every kernel in :mod:`byzpy_b200.ops`.

They serve two purposes: (1) the CPU execution path of the operator library
(tests, CPU actor pools), and (2) the fp32/fp64 oracle that GPU numerics tests
compare the hand-written kernels against.  Semantics follow the kernels
exactly, including NaN canonicalisation (NaN -> +inf before sorting).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

MODE_MEDIAN, MODE_TRMEAN, MODE_MEAMED, MODE_MEAN = 0, 1, 2, 3


def _stack(rows: Sequence[torch.Tensor], scales=None) -> torch.Tensor:
    base = rows[0]
    dtype = base.dtype if base.dtype.is_floating_point else torch.float32
    if all(r.dim() == 1 and r.dtype == dtype and r.device == base.device for r in rows):
        X = torch.stack(list(rows), dim=0)          # common case: nothing to normalise
    else:
        X = torch.stack([r.reshape(-1).to(device=base.device, dtype=dtype) for r in rows], dim=0)
    if scales is not None and len(scales):
        s = torch.as_tensor(list(scales), dtype=X.dtype, device=X.device)
        X = X * s[:, None]
    return X


def _canon(X: torch.Tensor) -> torch.Tensor:
    return torch.nan_to_num(X, nan=float("inf"), posinf=float("inf"), neginf=float("-inf"))


def cw_select(rows, mode: int, f: int = 0, *, scales=None,
              virtual: Optional[Tuple[int, int, float, float]] = None) -> torch.Tensor:
    X = _canon(_stack(rows, scales))
    if virtual is not None and virtual[0] > 0:
        nv, nh, a, b = virtual
        H = X[:nh]
        mean = H.mean(dim=0)
        std = ((H - mean) ** 2).mean(dim=0).sqrt()
        v = _canon(a * mean + b * std)
        X = torch.cat([X, v.unsqueeze(0).expand(nv, -1)], dim=0)
    n = X.shape[0]
    if mode == MODE_MEAN:
        return X.mean(dim=0)
    if mode == MODE_MEDIAN:
        # NaN was canonicalised to +inf above, so a k-th value selection (no full sort) has exactly
        # the kernel's lower-median semantics; kthvalue beats sort / median on CPU for n = 8..64
        return X.kthvalue((n - 1) // 2 + 1, dim=0).values
    S, _ = torch.sort(X, dim=0)
    mid = (n - 1) // 2
    if mode == MODE_TRMEAN:
        return S[f:n - f].mean(dim=0)
    if mode == MODE_MEAMED:
        m = S[mid]
        k = n - f
        # the k values closest to the median form a contiguous window of the sorted order
        left = (m.unsqueeze(0) - S[: n - k]) > (S[k:] - m.unsqueeze(0)) if n - k > 0 else None
        if left is None:
            return S.mean(dim=0)
        l = left.sum(dim=0)  # (d,)
        idx = l.unsqueeze(0) + torch.arange(k, device=X.device).unsqueeze(1)
        return torch.gather(S, 0, idx).mean(dim=0)
    raise ValueError(f"unknown mode {mode}")


def gram(rows, *, scales=None, want64: bool = False) -> torch.Tensor:
    X = _stack(rows, scales)
    if want64:
        X = X.double()
        return X @ X.T
    return (X.double() @ X.double().T).to(X.dtype)


def weighted_sum(rows, W: torch.Tensor, *, scales=None) -> torch.Tensor:
    X = _stack(rows, scales)
    W = W.to(device=X.device, dtype=X.dtype)
    if bool(torch.isfinite(X).all()):
        return W @ X
    # rows with zero weight are excluded outright (0 * inf must not poison the sum)
    out = torch.zeros((W.shape[0], X.shape[1]), dtype=X.dtype, device=X.device)
    for r in range(W.shape[0]):
        nz = (W[r] != 0).nonzero().flatten()
        if nz.numel():
            out[r] = (W[r, nz, None] * X[nz]).sum(dim=0)
    return out


def colstat(rows, a: float, b: float, *, scales=None) -> torch.Tensor:
    X = _stack(rows, scales)
    mean = X.mean(dim=0)
    out = a * mean
    if b != 0.0:
        out = out + b * ((X - mean) ** 2).mean(dim=0).sqrt()
    return out


def sgd_step(grad: torch.Tensor, *, params, moms=None, lr: float, momentum: float = 0.0,
             weight_decay: float = 0.0) -> None:
    with torch.no_grad():
        for r, p in enumerate(params):
            g = grad.reshape(p.shape).to(p.dtype)
            if weight_decay:
                g = g + weight_decay * p
            if moms:
                m = moms[r]
                m.mul_(momentum).add_(g)
                g = m
            p.add_(g, alpha=-lr)
