"""Coordinate-wise median (lower median for even n, i.e. ``torch.median`` semantics;
reference aggregators/coordinate_wise/median.py:28-178).  On CUDA this is one launch
of the register selection-network kernel (``csrc/cw_select.cu``)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class CoordinateWiseMedian(CoordinateWiseAggregator):
    """Median of every coordinate over the submitted gradients.

    For an even number of inputs the lower of the two middle values is returned (``torch.median`` semantics), so the
    output of every coordinate is one of the submitted values.  Tolerates up to ``(n - 1) // 2`` arbitrary inputs per
    coordinate.

    Parameters
    ----------
    chunk_size : int, default 8192
        Coordinates per subtask when the operator runs on an :class:`~byzpy_b200.engine.graph.pool.ActorPool`
        (adapted to the pool size, see ``aggregators/_chunking.py``).  The direct call ignores it.

    Notes
    -----
    CUDA inputs: one launch of the register selection-network kernel (``csrc/cw_select.cu``), NaN sorts last.
    Inside a fused parameter-server round the same network runs on the coordinate shard each GPU owns, reading the
    workers' rows from peer memory (``csrc/fused_ps.cu``).  Non-finite inputs are allowed; ``+inf`` / ``-inf`` order as
    numbers.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> grads = [torch.tensor([1.0, 10.0]), torch.tensor([2.0, -5.0]), torch.tensor([100.0, 0.0])]
    >>> CoordinateWiseMedian().aggregate(grads)
    tensor([2., 0.])
    """

    name = "coordinate-wise-median"
    _mode = ops.MODE_MEDIAN

    def __init__(self, *, chunk_size: int = 8192) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.chunk_size = int(chunk_size)


__all__ = ["CoordinateWiseMedian"]
