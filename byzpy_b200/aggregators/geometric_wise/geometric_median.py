"""Geometric median via Weiszfeld iterations, run entirely in Gram space.

The reference makes up to ``max_iter`` (256) full passes over the (n, d) matrix with a host
sync per iteration (reference aggregators/geometric_wise/geometric_median.py:79-104).  The
iterate never leaves the affine span of the rows (plus the start point), so here ONE Gram pass
feeds an O(n^2)-per-iteration solve on the coefficients, and one weighted-sum pass emits the
result: 2 reads of the data instead of up to 256.  Same update rule, ``eps`` clamp and
``||z_t+1 - z_t|| <= tol`` stopping rule.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ... import ops
from ...ops import nspace
from ..base import GramAggregator


class GeometricMedian(GramAggregator):
    """Geometric median: the point minimising the sum of Euclidean distances to the gradients (Weiszfeld's algorithm).

    Parameters
    ----------
    tol : float, default 1e-6
        Stop when two successive iterates are closer than this (Euclidean norm).
    max_iter : int, default 256
        Upper bound on Weiszfeld iterations.
    eps : float, default 1e-12
        Lower clamp on a distance before it is inverted (an iterate that lands on an input).
    init : {"median", "mean"}, default "median"
        Start point: coordinate-wise (lower) median or arithmetic mean of the inputs.
    chunk_size : int, default 32
        Rows per subtask; on an actor pool every iteration is one barriered round of subtasks.

    Attributes
    ----------
    last_iterations : int
        Iterations the most recent call took.

    Notes
    -----
    Every iterate is an affine combination of the inputs and the start point, so the iteration is carried out on the
    coefficient vector using only the Gram matrix of ``[X; start]``: one Gram pass, an ``O(n^2)`` update per
    iteration on one CTA, one weighted-sum pass.  The gradients are read twice however many iterations it takes (the
    reference reads them once per iteration and synchronises with the host each time).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import GeometricMedian
    >>> pts = [torch.tensor([0.0, 0.0]), torch.tensor([2.0, 0.0]), torch.tensor([1.0, 1.0]), torch.tensor([1.0, 50.0])]
    >>> out = GeometricMedian().aggregate(pts)
    >>> bool((out - torch.tensor([1.0, 1.0])).norm() < 1e-3)
    True
    """

    name = "geometric-median"
    shift_invariant = True       # distances only
    supports_barriered_subtasks = True
    device_solve = True

    def _fused_aux(self):
        return ("median",) if self.init == "median" else ()

    def __init__(self, *, tol: float = 1e-6, max_iter: int = 256, eps: float = 1e-12,
                 init: str = "median", chunk_size: int = 32) -> None:
        if tol <= 0:
            raise ValueError("tol must be > 0")
        if max_iter < 0:
            raise ValueError("max_iter must be >= 0")
        if eps <= 0:
            raise ValueError("eps must be > 0")
        if init not in ("median", "mean"):
            raise ValueError("init must be 'median' or 'mean'")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.tol, self.max_iter, self.eps, self.init = float(tol), int(max_iter), float(eps), init
        self.chunk_size = int(chunk_size)
        self.last_iterations = 0

    def _aux_rows(self, rows: List[torch.Tensor]) -> List[torch.Tensor]:
        if self.init == "median":
            return [ops.cw_median(rows)]
        return []

    def _start(self, n: int) -> np.ndarray:
        if self.init == "median":
            a0 = np.zeros(n + 1)
            a0[n] = 1.0
            return a0
        return np.full(n, 1.0 / n)

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        a, iters = nspace.weiszfeld_coeffs(G, n, self._start(n), tol=self.tol,
                                           max_iter=self.max_iter, eps=self.eps)
        self.last_iterations = iters
        return a

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.weiszfeld_coeffs(G, n, self._start(n), tol=self.tol,
                                            max_iter=self.max_iter, eps=self.eps)


__all__ = ["GeometricMedian"]
