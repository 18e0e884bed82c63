"""byzpy_b200 -- a Blackwell (B200, sm_100a) native Byzantine-robust distributed training
framework with the capabilities and public API surface of Byzpy/byzpy.

Top-level exports mirror the reference package (reference python/byzpy/__init__.py:1-4)."""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: keeps `import byzpy_b200` cheap and free of circular imports
    if name in ("run_operator", "OperatorExecutor"):
        from .engine.graph import executor

        return getattr(executor, name)
    raise AttributeError(f"module 'byzpy_b200' has no attribute {name!r}")


__all__ = ["run_operator", "OperatorExecutor", "__version__"]

# Library logging convention: everything logs under "byzpy_b200" and stays silent unless the application configures
# logging (failure detection, dropped peers and refused frames are reported at WARNING).
import logging as _logging  # noqa: E402

_logging.getLogger("byzpy_b200").addHandler(_logging.NullHandler())

