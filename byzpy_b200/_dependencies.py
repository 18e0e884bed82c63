"""Runtime dependency report (the reference's ``_dependencies.py`` selects cupy/ucxx extras at
install time via ``BYZPY_FORCE_GPU`` / ``BYZPY_FORCE_CPU``, reference _dependencies.py:34-77).
This framework has no optional GPU wheels -- its kernels are built in-tree -- so the same
environment variables only steer :func:`preferred_device`."""
from __future__ import annotations

import os


def _flag(name: str) -> bool:
    return os.environ.get(name, "").strip().lower() in ("1", "true", "yes", "on")


def preferred_device() -> str:
    """``"cuda"`` when a CUDA device is usable, else ``"cpu"``; ``BYZPY_FORCE_CPU`` / ``BYZPY_FORCE_GPU`` override (the latter
    raises when there is no device).
    """
    if _flag("BYZPY_FORCE_CPU"):
        return "cpu"
    try:
        import torch

        if torch.cuda.is_available():
            return "cuda"
    except Exception:
        pass
    if _flag("BYZPY_FORCE_GPU"):
        raise RuntimeError("BYZPY_FORCE_GPU is set but no CUDA device is available")
    return "cpu"


def base_requirements():
    """Run-time requirements of the package."""
    return ["torch>=2.6", "numpy>=1.24", "cloudpickle>=2.2", "tqdm>=4.65", "pybind11>=2.11"]


# The reference's packaging hooks (reference _dependencies.py:80-97).  There: cupy / ucxx wheels are added when a
# GPU is detected.  Here the GPU path has no extra wheels (the kernels are compiled in-tree by nvcc, CUDA tensors
# travel as CUDA-IPC handles), so the GPU list is empty and the full list does not depend on the machine.
def get_dependencies():
    """Requirements to install on this machine (the same everywhere: no per-platform GPU wheels)."""
    return list(base_requirements())


def get_gpu_optional_dependencies():
    """Extra wheels of the GPU path: none, the kernels are compiled in-tree."""
    return []


def get_dev_optional_dependencies():
    """Requirements of the test-suite."""
    return ["pytest>=8", "pytest-timeout", "hypothesis"]


__all__ = ["preferred_device", "base_requirements", "get_dependencies", "get_gpu_optional_dependencies",
           "get_dev_optional_dependencies"]
