from .training import train_with_progress

__all__ = ["train_with_progress"]
