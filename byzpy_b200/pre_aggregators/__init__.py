"""Pre-aggregation operators: linear maps X' = W X with a data-dependent W (counterpart of ``byzpy.pre_aggregators``).

Names are resolved from the table below (name -> defining submodule) so that the package namespace
and ``__all__`` cannot drift apart."""
from importlib import import_module as _import_module

_WHERE = {
    "PreAggregator": "base",
    "Bucketing": "bucketing",
    "NearestNeighborMixing": "nnm",
    "Clipping": "clipping",
    "ARC": "arc",
}

for _name, _module in _WHERE.items():
    globals()[_name] = getattr(_import_module(f"{__name__}.{_module}"), _name)

__all__ = list(_WHERE)
