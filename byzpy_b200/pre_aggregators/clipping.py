"""Static norm clipping: ``x_i <- x_i * min(1, tau / max(||x_i||, 1e-12))``
(reference pre_aggregators/clipping.py:35-130)."""
from __future__ import annotations

import numpy as np

from ..ops import nspace
from .base import LinearPreAggregator


class Clipping(LinearPreAggregator):
    """Static norm clipping: scale every vector whose Euclidean norm exceeds ``threshold`` back onto that radius.

    Parameters
    ----------
    threshold : float, default 2.0
        Maximum norm after clipping (``>= 0``).
    chunk_size : int, default 32
        Vectors per subtask on an actor pool.

    Notes
    -----
    A diagonal row map: only the row norms are needed (one norm pass), then ``n`` scaled copies -- or no copy at
    all when an aggregator follows, which then folds the scales into its own pass over the data (the fused
    coordinate-wise kernels take per-row scales; Gram-family aggregators rescale the Gram matrix).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.pre_aggregators import Clipping
    >>> Clipping(threshold=1.0).pre_aggregate([torch.tensor([3.0, 4.0]), torch.tensor([0.3, 0.4])])
    [tensor([0.6000, 0.8000]), tensor([0.3000, 0.4000])]
    """

    name = "pre-agg/clipping"
    gram_diag_only = True       # only the row norms are used: the torch fallback skips the n^2 d work
    diagonal_map = True

    def __init__(self, threshold: float = 2.0, *, chunk_size: int = 32) -> None:
        if threshold < 0:
            raise ValueError("threshold must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.threshold = float(threshold)
        self.chunk_size = int(chunk_size)

    def row_map(self, G, n):
        return np.diag(nspace.clip_scales(G, self.threshold))


    def row_map_device(self, G, n):
        from ..ops import nspace_cuda

        return nspace_cuda.clip_matrix(G, self.threshold)

__all__ = ["Clipping"]
