"""The examples in the docstrings of the operator library are executed (doctest), and every public class of the
operator library has a docstring with a Parameters section when its constructor takes parameters."""
import doctest
import importlib
import inspect
import pkgutil

import pytest

import byzpy_b200
import byzpy_b200.aggregators
import byzpy_b200.attacks
import byzpy_b200.pre_aggregators

PACKAGES = [byzpy_b200.aggregators, byzpy_b200.pre_aggregators, byzpy_b200.attacks]


def _modules(packages):
    out = []
    for pkg in packages:
        for info in pkgutil.walk_packages(pkg.__path__, pkg.__name__ + "."):
            leaf = info.name.rsplit(".", 1)[-1]
            if not leaf.startswith("_") and ".tests" not in info.name:
                out.append(info.name)
    return sorted(out)


MODULES = _modules(PACKAGES)                       # the operator library: documentation is mandatory
ALL_MODULES = _modules([byzpy_b200])               # everything: whatever examples exist must run


@pytest.mark.parametrize("name", ALL_MODULES)
def test_docstring_examples_run(name):
    mod = importlib.import_module(name)
    res = doctest.testmod(mod, optionflags=doctest.ELLIPSIS | doctest.NORMALIZE_WHITESPACE)
    assert res.failed == 0, f"{name}: {res.failed} of {res.attempted} docstring examples failed"


def test_every_operator_class_is_documented():
    from byzpy_b200.engine.graph.operator import Operator

    undocumented, without_example = [], []
    for name in MODULES:
        mod = importlib.import_module(name)
        for cls_name, cls in vars(mod).items():
            if not inspect.isclass(cls) or cls.__module__ != name or cls_name.startswith("_"):
                continue
            if not issubclass(cls, Operator) or inspect.isabstract(cls) and cls_name not in ("Attack", "PreAggregator",
                                                                                              "Aggregator"):
                continue
            doc = cls.__dict__.get("__doc__") or ""
            if len(doc) < 80:
                undocumented.append(f"{name}.{cls_name}")
                continue
            params = [p for p in inspect.signature(cls.__init__).parameters if p not in ("self", "args", "kwargs")]
            if params and not inspect.isabstract(cls) and "Parameters" not in doc:
                undocumented.append(f"{name}.{cls_name} (no Parameters section)")
            if ">>>" not in doc and not inspect.isabstract(cls):
                without_example.append(f"{name}.{cls_name}")
    assert not undocumented, undocumented
    assert not without_example, without_example


def test_every_public_definition_has_a_docstring():
    """Every public top-level class and function of the package carries a docstring (what autodoc renders in
    docs/source/api_reference.md)."""
    import ast
    import os

    root = os.path.dirname(byzpy_b200.__file__)
    missing = []
    for base, dirs, files in os.walk(root):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "csrc")]
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(base, f)
            with open(path) as fh:
                tree = ast.parse(fh.read())
            for node in tree.body:
                if isinstance(node, (ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)) \
                        and not node.name.startswith("_") and not ast.get_docstring(node):
                    missing.append(f"{os.path.relpath(path, root)}:{node.lineno} {node.name}")
    assert not missing, missing
