from . import metrics
from .checkpoint import load_checkpoint, save_checkpoint
from .tracing import Tracer, cuda_time_ms, nvtx_range
from .training import train_with_progress

__all__ = ["train_with_progress", "save_checkpoint", "load_checkpoint", "Tracer", "nvtx_range", "cuda_time_ms", "metrics"]
