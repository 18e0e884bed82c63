"""NumPy implementation of the array-backend protocol (the reference's ``set_backend("numpy")`` is dead
code; here it is honoured)."""
from . import _NumpyBackend

__all__ = ["_NumpyBackend"]
