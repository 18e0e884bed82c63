"""CAF -- covariance-bound agnostic filter (reference aggregators/norm_wise/caf.py:133-184).

Iteratively down-weights points with a large projection on the dominant eigen-direction of the
weighted covariance (power iteration started from the fixed ``default_rng(0)`` direction r, as
the reference does) and returns the weighted mean with the smallest eigenvalue seen.  The
power iteration lives in the span of the centred rows, so after ONE Gram pass over
``[X; r]`` the whole filter loop is O(n^2) arithmetic on the host."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from ...ops import nspace
from ..base import GramAggregator

_DIR_CACHE: Dict[Tuple[int, str], torch.Tensor] = {}


def _start_direction(d: int, like: torch.Tensor) -> torch.Tensor:
    key = (d, str(like.device))
    r = _DIR_CACHE.get(key)
    if r is None:
        vec = np.random.default_rng(0).normal(size=d).astype(np.float32, copy=False)
        r = torch.from_numpy(vec).to(like.device)
        if len(_DIR_CACHE) > 4:
            _DIR_CACHE.clear()
        _DIR_CACHE[key] = r
    return r


class CAF(GramAggregator):
    """Covariance-bound Agnostic Filter: an iteratively re-weighted mean that suppresses the dominant outlier direction.

    Each round computes the weighted mean and covariance, finds the covariance's leading eigenvector by power
    iteration, and shrinks the weight of every gradient in proportion to its squared projection on it; rounds stop
    once the total removed weight reaches ``2 f``.  The weighted mean with the smallest leading eigenvalue seen is
    returned.

    Parameters
    ----------
    f : int
        Upper bound on the number of Byzantine inputs; ``2 f < n``.
    chunk_size : int, default 256
        Rows per subtask on an actor pool.
    power_iters : int, default 3
        Power-iteration steps per round (started from a fixed pseudo-random direction, so results are reproducible).

    Notes
    -----
    The power iteration stays in the span of the centred gradients and the start direction: after one Gram pass over
    ``[X; r]`` the whole filter is ``O(n^2)`` arithmetic per step, done by one single-CTA kernel in fp64
    (``csrc/nspace_maps.cu``), followed by one weighted-sum pass.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.norm_wise import CAF
    >>> torch.manual_seed(0)
    <torch._C.Generator object at ...>
    >>> honest = [torch.ones(8) + 0.01 * torch.randn(8) for _ in range(6)]
    >>> out = CAF(f=1).aggregate(honest + [torch.full((8,), 50.0)])
    >>> bool((out - 1.0).abs().max() < 0.5)
    True
    """

    name = "caf"

    def __init__(self, f: int, *, chunk_size: int = 256, power_iters: int = 3) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        if power_iters < 0:
            raise ValueError("power_iters must be >= 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)
        self.power_iters = int(power_iters)

    def _validate(self, n: int) -> None:
        if 2 * self.f >= n:
            raise ValueError(f"Cannot tolerate 2f >= n (got n={n}, f={self.f}).")

    def _aux_rows(self, rows: List[torch.Tensor]) -> List[torch.Tensor]:
        r = _start_direction(rows[0].numel(), rows[0])
        return [r.to(rows[0].dtype)]

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        w = np.zeros(n + 1)
        w[:n] = nspace.caf_coeffs(G, n, self.f, power_iters=self.power_iters)
        return w

    # ---- sm_100a path: the filter loop is one single-CTA kernel on the device Gram (csrc/nspace_maps.cu)
    device_solve = True

    def _solve_device(self, G: torch.Tensor, n: int):
        from ...ops import nspace_cuda

        return nspace_cuda.caf_coeffs(G, n, self.f, power_iters=self.power_iters)

    def _device_solve_feasible(self, n: int) -> bool:
        return n <= 127

    def _fused_aux(self):
        # the fused round appends the fixed start direction as a constant auxiliary row
        return (("const", _start_direction),)


__all__ = ["CAF"]
