"""Drop-in import alias: make ``import byzpy...`` resolve to this framework.

    import byzpy_b200.compat; byzpy_b200.compat.install_alias()
    from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian      # -> byzpy_b200
    from byzpy.engine.graph.pool import ActorPool, ActorPoolConfig

The module tree mirrors the reference, so every public import path of
Byzpy/byzpy maps 1:1.  The alias is opt-in (it would shadow a real ``byzpy`` installation).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

_PREFIX = "byzpy"
_TARGET = "byzpy_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name: str) -> None:
        self.real_name = real_name

    def create_module(self, spec):
        module = importlib.import_module(self.real_name)
        self._real_spec = getattr(module, "__spec__", None)
        return module

    def exec_module(self, module):  # already executed under its real name
        # importlib has just stamped the ALIAS spec on the (shared) module object; with it, ``__spec__.parent``
        # (``byzpy...``) no longer matches ``__package__`` (``byzpy_b200...``) and every later relative import
        # inside the module warns (an error from Python 3.15 on).  The module keeps its own identity.
        if self._real_spec is not None:
            module.__spec__ = self._real_spec
        return None


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _PREFIX and not fullname.startswith(_PREFIX + "."):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real),
                                               is_package=real_spec.submodule_search_locations is not None)


_finder = None


def install_alias() -> None:
    """Make ``import byzpy`` (and every ``byzpy.x.y``) resolve to ``byzpy_b200`` (``byzpy_b200.x.y``) from now on, through a
    meta-path finder; idempotent.  Both names then refer to the SAME module objects.
    """
    global _finder
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)


def uninstall_alias() -> None:
    """Remove the finder and the ``byzpy*`` aliases it created from ``sys.modules``."""
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in [m for m in sys.modules if m == _PREFIX or m.startswith(_PREFIX + ".")]:
        if getattr(sys.modules[name], "__name__", "").startswith(_TARGET):
            del sys.modules[name]


__all__ = ["install_alias", "uninstall_alias"]
