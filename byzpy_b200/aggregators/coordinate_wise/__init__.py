"""Coordinate-wise robust aggregators: one streaming pass with a register selection network.

Names are resolved from the table below (name -> defining submodule) so that the package namespace
and ``__all__`` cannot drift apart."""
from importlib import import_module as _import_module

_WHERE = {
    "MeanOfMedians": "mean_of_medians",
    "CoordinateWiseMedian": "median",
    "CoordinateWiseTrimmedMean": "trimmed_mean",
}

for _name, _module in _WHERE.items():
    globals()[_name] = getattr(_import_module(f"{__name__}.{_module}"), _name)

__all__ = list(_WHERE)
