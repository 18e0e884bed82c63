"""Coordinate-wise median (lower median for even n, i.e. ``torch.median`` semantics;
reference aggregators/coordinate_wise/median.py:28-178).  On CUDA this is one launch
of the register selection-network kernel (``csrc/cw_select.cu``)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class CoordinateWiseMedian(CoordinateWiseAggregator):
    name = "coordinate-wise-median"
    _mode = ops.MODE_MEDIAN

    def __init__(self, *, chunk_size: int = 8192) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.chunk_size = int(chunk_size)


__all__ = ["CoordinateWiseMedian"]
