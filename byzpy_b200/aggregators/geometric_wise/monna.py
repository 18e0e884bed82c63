"""MoNNA: mean of the n-f vectors nearest to a trusted reference vector
(reference aggregators/geometric_wise/monna.py:36-178)."""
from __future__ import annotations

import numpy as np

from ...ops import nspace
from ..base import GramAggregator


class MoNNA(GramAggregator):
    """MoNNA: mean of the ``n - f`` gradients nearest to a trusted reference gradient.

    Parameters
    ----------
    f : int
        Number of gradients left out; ``0 <= 2 f < n``.
    reference_index : int, default 0
        Position of the trusted gradient in the input list (a node aggregating its neighbours passes its own vector
        first).
    chunk_size : int, default 32
        Rows per subtask on an actor pool.

    Notes
    -----
    Needs one row of the distance matrix, read off the Gram matrix; ties go to the lower index.  Because the
    parameter server's device path keeps gradients in rank order, ``reference_index`` names the same worker every
    round (the reference collects in completion order).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import MoNNA
    >>> grads = [torch.tensor([1.0]), torch.tensor([1.2]), torch.tensor([0.9]), torch.tensor([-30.0])]
    >>> MoNNA(f=1).aggregate(grads)
    tensor([1.0333])
    """

    name = "monna"
    shift_invariant = True       # distances only
    device_solve = True

    def __init__(self, f: int, *, reference_index: int = 0, chunk_size: int = 32) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if reference_index < 0:
            raise ValueError("reference_index must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.reference_index = int(reference_index)
        self.chunk_size = int(chunk_size)

    def _validate(self, n: int) -> None:
        if not (0 <= 2 * self.f < n):
            raise ValueError(f"2f must be < n (got n={n}, f={self.f})")
        if self.reference_index >= n:
            raise ValueError(f"reference_index {self.reference_index} out of range for n={n}")

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        return nspace.monna_weights(G, self.f, self.reference_index)

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.monna_weights(G, n, self.f, self.reference_index)


__all__ = ["MoNNA"]
