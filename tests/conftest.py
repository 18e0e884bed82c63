import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")
    config.addinivalue_line("markers", "real_actor_backends: use real actor backends in pool tests")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if item.get_closest_marker("gpu") is not None:
            item.add_marker(skip)


def run_async(coro):
    import asyncio

    return asyncio.run(coro)


@pytest.fixture
def byzpy_alias():
    """``import byzpy...`` resolves to THIS package for the duration of the test, whatever earlier tests left in
    ``sys.modules`` (the parity tests import the real reference from ``baseline/_ref`` under the same name): the
    real modules are set aside, the alias is installed, and both are put back afterwards."""
    from byzpy_b200 import compat

    saved = {k: v for k, v in sys.modules.items() if k == "byzpy" or k.startswith("byzpy.")}
    for k in saved:
        del sys.modules[k]
    had_alias = compat._finder is not None
    compat.install_alias()
    try:
        yield compat
    finally:
        if not had_alias:
            compat.uninstall_alias()
        for k in [k for k in sys.modules if k == "byzpy" or k.startswith("byzpy.")]:
            del sys.modules[k]
        sys.modules.update(saved)
