"""Pre-aggregation operators (reference package ``byzpy.pre_aggregators``)."""
from .arc import ARC
from .base import PreAggregator
from .bucketing import Bucketing
from .clipping import Clipping
from .nnm import NearestNeighborMixing

__all__ = ["PreAggregator", "Bucketing", "NearestNeighborMixing", "Clipping", "ARC"]
