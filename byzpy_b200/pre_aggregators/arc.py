"""Adaptive robust clipping (ARC): clip every vector to the norm of the
``n - floor(2f(n-f)/n)``-th smallest one (reference pre_aggregators/arc.py:36-161)."""
from __future__ import annotations

import numpy as np

from ..ops import nspace
from .base import LinearPreAggregator


class ARC(LinearPreAggregator):
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
