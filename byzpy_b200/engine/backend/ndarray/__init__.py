"""Minimal array-backend protocol (reference engine/backend/ndarray/base.py:8-27,
torch.py:10-71): the 16 array primitives the generic operator code may use, for torch and numpy.
The hot paths do not go through this layer -- they call :mod:`byzpy_b200.ops` directly."""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
# NOT ``import torch``: this package has a submodule called ``torch`` (the reference's import path
# ``engine.backend.ndarray.torch``), and importing a submodule rebinds the attribute of the same name on the
# package -- i.e. this module's global -- to the SUBMODULE, after which every ``torch.stack`` below would fail
import torch as _torch


class _TorchBackend:
    name = "torch"

    def asarray(self, x: Any, like: Any = None):
        if isinstance(like, _torch.Tensor):
            return _torch.as_tensor(x, dtype=like.dtype, device=like.device)
        return _torch.as_tensor(x)

    def stack(self, xs: Sequence[Any], axis: int = 0):
        return _torch.stack(list(xs), dim=axis)

    def median(self, x, axis: int = 0):
        return _torch.median(x, dim=axis).values

    def mean(self, x, axis=None):
        return x.mean() if axis is None else x.mean(dim=axis)

    def sum(self, x, axis=None):
        return x.sum() if axis is None else x.sum(dim=axis)

    def sort(self, x, axis: int = 0):
        return _torch.sort(x, dim=axis).values

    def argsort(self, x, axis: int = -1):
        return _torch.argsort(x, dim=axis)

    def sqrt(self, x):
        return _torch.sqrt(x)

    def maximum(self, a, b):
        return _torch.maximum(a, b)

    def minimum(self, a, b):
        return _torch.minimum(a, b)

    def abs(self, x):
        return _torch.abs(x)

    def reshape(self, x, shape):
        return x.reshape(shape)

    def copy(self, x):
        return x.clone()

    def matmul(self, a, b):
        return a @ b

    def index_select(self, x, axis: int, indices):
        return _torch.index_select(x, axis, _torch.as_tensor(indices, device=x.device))

    def max(self, x, axis=None):
        return x.max() if axis is None else _torch.amax(x, dim=axis)

    def take_along_axis(self, a, indices, axis: int = 0):
        return _torch.take_along_dim(a, _torch.as_tensor(indices, device=a.device), dim=axis)


class _NumpyBackend:
    name = "numpy"

    def asarray(self, x: Any, like: Any = None):
        return np.asarray(x, dtype=getattr(like, "dtype", None))

    def stack(self, xs, axis: int = 0):
        return np.stack(list(xs), axis=axis)

    def median(self, x, axis: int = 0):
        n = x.shape[axis]
        return np.take(np.sort(x, axis=axis), (n - 1) // 2, axis=axis)  # lower median, like torch

    def mean(self, x, axis=None):
        return np.mean(x, axis=axis)

    def sum(self, x, axis=None):
        return np.sum(x, axis=axis)

    def sort(self, x, axis: int = 0):
        return np.sort(x, axis=axis)

    def argsort(self, x, axis: int = -1):
        return np.argsort(x, axis=axis, kind="stable")

    def sqrt(self, x):
        return np.sqrt(x)

    def maximum(self, a, b):
        return np.maximum(a, b)

    def minimum(self, a, b):
        return np.minimum(a, b)

    def abs(self, x):
        return np.abs(x)

    def reshape(self, x, shape):
        return np.reshape(x, shape)

    def copy(self, x):
        return np.array(x, copy=True)

    def matmul(self, a, b):
        return a @ b

    def index_select(self, x, axis: int, indices):
        return np.take(x, np.asarray(indices), axis=axis)

    def max(self, x, axis=None):
        return np.max(x, axis=axis)

    def take_along_axis(self, a, indices, axis: int = 0):
        return np.take_along_axis(a, np.asarray(indices), axis=axis)


_BACKENDS = {"torch": _TorchBackend, "numpy": _NumpyBackend}


def get_array_backend(name: str = "torch"):
    """Array backend by name: ``"torch"`` or ``"numpy"`` (the small array API operators written against the backend
    protocol use; ``configs.backend.set_backend`` selects the process default).
    """
    try:
        return _BACKENDS[name]()
    except KeyError:
        raise ValueError(f"unknown backend {name!r}; choose 'torch' or 'numpy'") from None


__all__ = ["get_array_backend", "_TorchBackend", "_NumpyBackend"]
