"""Plain PyTorch implementations of every kernel in :mod:`byzpy_b200.ops`.

The math is that of the reference operators (median.py:102-106, trimmed_mean.py:110-115,
mean_of_medians.py:71-81, krum.py:41-44, little.py:113-131, ...) expressed on a stacked (n, d) matrix.
They serve two purposes: (1) the CPU execution path of the operator library
(tests, CPU actor pools), and (2) the fp32/fp64 oracle that GPU numerics tests
compare the hand-written kernels against.  Semantics follow the kernels
exactly, including NaN canonicalisation (NaN -> +inf before sorting).
"""
from __future__ import annotations

import weakref
from typing import Optional, Sequence, Tuple

import torch

MODE_MEDIAN, MODE_TRMEAN, MODE_MEAMED, MODE_MEAN = 0, 1, 2, 3


# The (n, d) stack built by gram() is reused by the weighted_sum() that follows on the same,
# unmodified rows (every Gram-family operator does exactly that): one entry, consumed on use.
_STACK_CACHE: dict = {}


def _stack_key(rows, scales):
    return (tuple((id(r), r._version) for r in rows), None if scales is None else tuple(float(v) for v in scales))


def _stack(rows: Sequence[torch.Tensor], scales=None, *, remember: bool = False, reuse: bool = False) -> torch.Tensor:
    if remember or reuse:
        key = _stack_key(rows, scales)
        if reuse:
            hit = _STACK_CACHE.pop("entry", None)
            # ids are only unique among LIVE objects: the weak references prove that the remembered rows are
            # these very tensors and not freed ones whose ids (and version counters) were reused
            if hit is not None and hit[0] == key and all(ref() is r for ref, r in zip(hit[2], rows)):
                return hit[1]
    X = _stack_impl(rows, scales)
    if remember and X.numel() <= (1 << 26):
        _STACK_CACHE["entry"] = (key, X, tuple(weakref.ref(r) for r in rows))
    return X


def _stack_impl(rows: Sequence[torch.Tensor], scales=None) -> torch.Tensor:
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
    if not bool(torch.isnan(X).any()):
        return X                                             # the common case: skip a full pass
    return torch.nan_to_num(X, nan=float("inf"), posinf=float("inf"), neginf=float("-inf"))


def cw_select(rows, mode: int, f: int = 0, *, scales=None,
              virtual: Optional[Tuple[int, int, float, float]] = None) -> torch.Tensor:
    """NaN counts as +inf (the kernels canonicalise on load).  ATen's sort / k-th value already order NaN as
    the largest value, so instead of scanning the whole (n, d) matrix for NaN up front the clean path runs
    first and only its (d,)-sized result is inspected."""
    X = _stack(rows, scales)
    if virtual is not None and virtual[0] > 0:
        X = _canon(X)
        nv, nh, a, b = virtual
        H = X[:nh]
        mean = H.mean(dim=0)
        std = ((H - mean) ** 2).mean(dim=0).sqrt()
        v = _canon(a * mean + b * std)
        X = torch.cat([X, v.unsqueeze(0).expand(nv, -1)], dim=0)
    n = X.shape[0]
    if mode == MODE_MEAN:
        out = X.mean(dim=0)
        return _canon(X).mean(dim=0) if bool(torch.isnan(out).any()) else out
    if mode == MODE_MEDIAN:
        # a k-th value selection (no full sort) has exactly the kernel's lower-median semantics;
        # kthvalue beats sort / median on CPU for n = 8..64, median wins from n ~ 32 (measured)
        k = (n - 1) // 2 + 1
        if n >= 32:
            out = X.median(dim=0).values                     # propagates NaN from anywhere in the column
            bad = torch.isnan(out)
            if bool(bad.any()):
                cols = bad.nonzero().flatten()
                out[cols] = _canon(X[:, cols]).kthvalue(k, dim=0).values
            return out
        out = X.kthvalue(k, dim=0).values                    # NaN ranks last: NaN out == +inf canonically
        return torch.where(torch.isnan(out), torch.full_like(out, float("inf")), out)
    S, _ = torch.sort(X, dim=0)                              # NaN ranks last here too
    if bool(torch.isnan(S[-1]).any()):
        S = torch.nan_to_num(S, nan=float("inf"), posinf=float("inf"), neginf=float("-inf"))
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


_GRAM_CHUNK = 8192


def gram(rows, *, scales=None, want64: bool = False, diag_only: bool = False) -> torch.Tensor:
    """``G = X X^T``.  fp32 inputs: fp32 GEMMs over column chunks, accumulated in fp64 (split-K, like
    the CUDA kernels) -- several times faster than a straight fp64 GEMM on CPU and far more accurate
    than one fp32 GEMM.  ``diag_only``: only the squared row norms are needed (Clipping / ARC / CGE),
    O(n d) instead of O(n^2 d); off-diagonal entries are zero."""
    if diag_only:
        # per-row norms, no (n, d) stack: O(n d) reads and nothing else
        sc = list(scales) if scales is not None and len(scales) else None
        sq = []
        for i, r in enumerate(rows):
            r = r.reshape(-1)
            # BLAS dot: ~10x faster than an fp64-accumulating norm and 1e-7 accurate; squares that overflow
            # fp32 (a 1e20-sized attack vector) or non-finite entries take the fp64 norm instead
            v = float(torch.dot(r, r)) if r.dtype in (torch.float32, torch.float64) else float("inf")
            if not (v < float("inf")):
                v = float(torch.linalg.vector_norm(r, ord=2, dtype=torch.float64)) ** 2
            sq.append(v * (float(sc[i]) ** 2 if sc is not None else 1.0))
        base = rows[0]
        out_dtype = torch.float64 if want64 else (base.dtype if base.dtype.is_floating_point else torch.float32)
        return torch.diag(torch.tensor(sq, dtype=torch.float64, device=base.device)).to(out_dtype)
    X = _stack(rows, scales, remember=True)
    out_dtype = torch.float64 if want64 else X.dtype
    if X.dtype == torch.float64:
        return (X @ X.T).to(out_dtype)
    d = X.shape[1]
    if d <= _GRAM_CHUNK:
        G = (X @ X.T).double() if X.dtype == torch.float32 else X.double() @ X.double().T
        return G.to(out_dtype)
    G = torch.zeros((X.shape[0], X.shape[0]), dtype=torch.float64, device=X.device)
    Xf = X.float() if X.dtype != torch.float32 else X
    for s0 in range(0, d, _GRAM_CHUNK):
        c = Xf[:, s0:s0 + _GRAM_CHUNK]
        G += (c @ c.T).double()
    return G.to(out_dtype)


def weighted_sum(rows, W: torch.Tensor, *, scales=None) -> torch.Tensor:
    """Plain PyTorch ``Y = W X`` over the rows (with optional per-row scales): the CPU path and the oracle of the CUDA kernels."""
    if W.dim() == 1 or W.shape[0] == 1:
        # one output row: accumulate the non-zero rows directly (no (n, d) stack, and rows with zero
        # weight -- possibly holding inf -- are never touched)
        w = W.reshape(-1).tolist()
        base = rows[0]
        dtype = base.dtype if base.dtype.is_floating_point else torch.float32
        sc = list(scales) if scales is not None and len(scales) else None
        out = torch.zeros(base.numel(), dtype=dtype, device=base.device)
        for j, wj in enumerate(w):
            if wj != 0.0:
                out.add_(rows[j].reshape(-1).to(device=base.device, dtype=dtype),
                         alpha=wj * (float(sc[j]) if sc is not None else 1.0))
        return out.unsqueeze(0)
    m, n = W.shape
    nzmask = W != 0
    per_row = int(nzmask.sum(dim=1).max().item()) if m else 0
    if per_row * 4 <= n and not (m == n and per_row <= 1):
        # sparse rows (bucket means, small neighbourhoods): accumulate straight from the rows
        base = rows[0]
        dtype = base.dtype if base.dtype.is_floating_point else torch.float32
        sc = list(scales) if scales is not None and len(scales) else None
        out = torch.zeros((m, base.numel()), dtype=dtype, device=base.device)
        Wl = W.tolist()
        for r in range(m):
            for j in nzmask[r].nonzero().flatten().tolist():
                out[r].add_(rows[j].reshape(-1).to(device=base.device, dtype=dtype),
                            alpha=Wl[r][j] * (float(sc[j]) if sc is not None else 1.0))
        return out
    X = _stack(rows, scales, reuse=True)
    W = W.to(device=X.device, dtype=X.dtype)
    if _all_finite(X):
        if m == n and per_row <= 1 and bool((nzmask == torch.eye(n, dtype=torch.bool, device=W.device)).all()):
            return X * torch.diagonal(W)[:, None]            # diagonal map (Clipping / ARC): n d, not n^2 d
        if per_row * 4 <= n:
            # sparse rows (bucket means, small neighbourhoods): gather + weighted sum per output row
            out = torch.empty((m, X.shape[1]), dtype=X.dtype, device=X.device)
            for r in range(m):
                idx = nzmask[r].nonzero().flatten()
                if idx.numel() == 0:
                    out[r].zero_()
                else:
                    out[r] = W[r, idx] @ X[idx]
            return out
        return W @ X
    # rows with zero weight are excluded outright (0 * inf must not poison the sum)
    out = torch.zeros((W.shape[0], X.shape[1]), dtype=X.dtype, device=X.device)
    for r in range(W.shape[0]):
        nz = (W[r] != 0).nonzero().flatten()
        if nz.numel():
            out[r] = (W[r, nz, None] * X[nz]).sum(dim=0)
    return out


def _all_finite(X: torch.Tensor) -> bool:
    """A finite total proves every entry finite (any inf / NaN entry makes the sum non-finite) in one
    reduction; a non-finite total may just be overflow, so only then pay for the exact element-wise test."""
    return bool(torch.isfinite(X.sum())) or bool(torch.isfinite(X).all())


def colstat(rows, a: float, b: float, *, scales=None) -> torch.Tensor:
    """Plain PyTorch ``a * mean + b * std`` per coordinate (population std) over the rows."""
    X = _stack(rows, scales)
    mean = X.mean(dim=0)
    out = a * mean
    if b != 0.0:
        out = out + b * ((X - mean) ** 2).mean(dim=0).sqrt()
    return out


def sgd_step(grad: torch.Tensor, *, params, moms=None, lr: float, momentum: float = 0.0,
             weight_decay: float = 0.0) -> None:
    """Plain PyTorch SGD(+momentum, weight decay) update of flat parameter / momentum buffers with the flat gradient ``grad``."""
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
