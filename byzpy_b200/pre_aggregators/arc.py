"""Adaptive robust clipping (ARC): clip every vector to the norm of the
``n - floor(2f(n-f)/n)``-th smallest one (reference pre_aggregators/arc.py:36-161)."""
from __future__ import annotations

import numpy as np

from ..ops import nspace
from .base import LinearPreAggregator


class ARC(LinearPreAggregator):
    """Adaptive Robust Clipping: clip to a radius chosen from the data instead of a fixed threshold.

    The vectors are ranked by norm; the ``floor(2 f (n - f) / n)`` largest ones are clipped to the norm of the largest
    remaining vector.  With ``f = 0`` nothing is clipped.

    Parameters
    ----------
    f : int, default 0
        Expected number of Byzantine vectors; ``0 <= f <= n``.
    chunk_size : int, default 32
        Vectors per subtask on an actor pool.

    Notes
    -----
    Diagonal row map like :class:`Clipping`; the threshold is found from the Gram diagonal on the device
    (``csrc/nspace_maps.cu``), so a fused round that uses ARC needs no host round trip and stays graph-capturable.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.pre_aggregators import ARC
    >>> xs = [torch.tensor([1.0, 0.0]), torch.tensor([0.0, 2.0]), torch.tensor([3.0, 0.0]), torch.tensor([0.0, 40.0])]
    >>> [round(x.norm().item(), 4) for x in ARC(f=1).pre_aggregate(xs)]
    [1.0, 2.0, 3.0, 3.0]
    """

    name = "pre-agg/arc"
    gram_diag_only = True       # only the row norms are used: the torch fallback skips the n^2 d work
    diagonal_map = True

    def __init__(self, f: int = 0, *, chunk_size: int = 32) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)

    def _validate(self, n: int) -> None:
        if self.f > n:
            raise ValueError(f"f must be <= number of vectors (got f={self.f}, n={n})")

    def row_map(self, G, n):
        return np.diag(nspace.arc_scales(G, self.f))


    def row_map_device(self, G, n):
        from ..ops import nspace_cuda

        return nspace_cuda.arc_matrix(G, self.f)

__all__ = ["ARC"]
