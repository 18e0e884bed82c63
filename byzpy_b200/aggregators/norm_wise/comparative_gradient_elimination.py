"""CGE: drop the f largest-norm gradients, average the rest (reference
aggregators/norm_wise/comparative_gradient_elimination.py:28-154).  Only diag(G) is needed."""
from __future__ import annotations

import numpy as np

from ...ops import nspace
from ..base import GramAggregator


class ComparativeGradientElimination(GramAggregator):
    """CGE: drop the ``f`` gradients with the largest norms and average the rest.

    Parameters
    ----------
    f : int
        Number of gradients eliminated; ``0 <= f < n``.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool (partial squared norms, then partial sums).

    Notes
    -----
    Only the diagonal of the Gram matrix is used: a row-norm pass and a weighted-sum pass.  Ties in the norm keep the
    lower index.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.norm_wise import ComparativeGradientElimination
    >>> grads = [torch.tensor([1.0, 0.0]), torch.tensor([0.0, 1.0]), torch.tensor([30.0, 30.0])]
    >>> ComparativeGradientElimination(f=1).aggregate(grads)
    tensor([0.5000, 0.5000])
    """

    name = "comparative-gradient-elimination"
    gram_diag_only = True       # only the row norms are used: the torch fallback skips the n^2 d work
    device_solve = True

    def __init__(self, f: int, *, chunk_size: int = 8192) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)

    def _validate(self, n: int) -> None:
        if not (0 <= self.f < n):
            raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={self.f})")

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        return nspace.cge_weights(G, self.f)

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.cge_weights(G, n, self.f)


__all__ = ["ComparativeGradientElimination"]
