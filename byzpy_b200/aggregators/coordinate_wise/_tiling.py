"""``flatten_gradients``: inputs (tensors / ndarrays / shm handles / handle dicts) ->
``(feature_shape, (n, d) float ndarray)``; contract of reference
aggregators/coordinate_wise/_tiling.py:18-38."""
from __future__ import annotations

from typing import Any, Sequence, Tuple

import numpy as np

from ...engine.storage.shared_store import materialize


def flatten_gradients(gradients: Sequence[Any]) -> Tuple[Tuple[int, ...], np.ndarray]:
    """``(feature shape, (n, d) float array)`` of a gradient list (tensors, arrays or shared-memory handles)."""
    arrays = [materialize(g).detach().cpu().numpy() for g in gradients]
    stacked = np.stack(arrays, axis=0)
    return tuple(stacked.shape[1:]), stacked.reshape(stacked.shape[0], -1)


__all__ = ["flatten_gradients"]
