"""Minimum Diameter Averaging: mean of the (n-f)-subset with the smallest diameter
(max pairwise squared distance); ties resolve to the lexicographically first subset, as the
reference's seeded DFS does (reference
aggregators/geometric_wise/minimum_diameter_average.py:328-386).  The search runs on the
(n, n) distance matrix only (threshold bisection + bitset clique search)."""
from __future__ import annotations

import numpy as np

from ...ops import nspace
from ..base import GramAggregator


class MinimumDiameterAveraging(GramAggregator):
    """Minimum Diameter Averaging: mean of the ``n - f`` gradients that fit in the smallest ball-like set.

    Among all subsets of size ``n - f`` the one with the smallest diameter (largest pairwise distance) is averaged;
    between subsets of equal diameter the lexicographically first wins.

    Parameters
    ----------
    f : int
        Number of gradients left out; ``0 <= f < n``.
    chunk_size : int, default 256
        Subsets scored per subtask on an actor pool.

    Notes
    -----
    The search runs on the ``n x n`` distance matrix only.  Instead of enumerating ``C(n, f)`` subsets, the diameter is
    found by bisection over the sorted distances with a bitset clique search ("is there a set of ``n - f`` points whose
    pairwise distances are all below t?"), which is what makes ``n = 30, f = 10`` take milliseconds; small instances
    are searched exhaustively on the device (``csrc/nspace.cu``) so the round stays capturable in a CUDA graph.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import MinimumDiameterAveraging
    >>> grads = [torch.tensor([0.0]), torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([40.0])]
    >>> MinimumDiameterAveraging(f=1).aggregate(grads)
    tensor([1.])
    """

    name = "minimum-diameter-averaging"
    shift_invariant = True       # distances only

    def __init__(self, f: int, *, chunk_size: int = 256) -> None:
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
        return nspace.mda_weights(G, self.f)

    # exhaustive subset search on the device (csrc/nspace.cu) when C(n, n-f) is small enough;
    # returns None otherwise and the host search above takes over
    device_solve = True

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.subset_weights(G, n, n - self.f, "mda")

    def _device_solve_feasible(self, n: int) -> bool:
        from ...ops import nspace_cuda

        return nspace_cuda.subset_search_feasible(n, n - self.f)


__all__ = ["MinimumDiameterAveraging"]
