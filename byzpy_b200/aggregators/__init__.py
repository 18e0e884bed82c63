"""Robust gradient aggregators (reference package ``byzpy.aggregators``)."""
from .base import Aggregator

__all__ = ["Aggregator"]
