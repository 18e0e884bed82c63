"""Torch implementation of the array-backend protocol (reference engine/backend/ndarray/torch.py:10-71).
Import path kept for code written against the reference; the class lives in the package."""
from . import _TorchBackend

__all__ = ["_TorchBackend"]
