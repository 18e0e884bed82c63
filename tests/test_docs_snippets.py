"""Python code blocks of the user-facing documentation are executed: the architecture overview and the two
self-contained blocks of the getting-started page (operators, custom aggregator)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = os.path.join(ROOT, "docs", "source")
_BLOCK = re.compile(r"```python\n(.*?)```", re.S)


def _blocks(name):
    with open(os.path.join(DOCS, name)) as f:
        return _BLOCK.findall(f.read())


def test_overview_blocks_run(capsys):
    blocks = _blocks("overview.md")
    assert blocks
    for src in blocks:
        exec(compile(src, "overview.md", "exec"), {"__name__": "__docs__"})
    assert "'median': (10000,)" in capsys.readouterr().out


def test_getting_started_operator_and_extension_blocks_run():
    blocks = _blocks("getting_started.md")
    ops_block = next(b for b in blocks if "EmpireAttack(scale=-1.0)" in b)
    ext_block = next(b for b in blocks if "class ClippedMean" in b)
    ns = {"__name__": "__docs__"}
    exec(compile(ops_block.replace("1_000_000", "10_000"), "getting_started.md", "exec"), ns)
    assert ns["g1"].shape == ns["g2"].shape == (10_000,)
    exec(compile(ext_block, "getting_started.md", "exec"), ns)
    import torch

    out = ns["ClippedMean"](tau=1.0).aggregate([torch.tensor([3.0, 4.0]), torch.tensor([0.0, 0.5])])
    assert torch.allclose(out, torch.tensor([0.3, 0.65]))


def test_installation_alias_block_runs(byzpy_alias):
    import importlib

    assert importlib.import_module("byzpy.aggregators").__name__ in ("byzpy_b200.aggregators", "byzpy.aggregators")


def _doc_files():
    out = [os.path.join("docs", "source", f) for f in sorted(os.listdir(DOCS)) if f.endswith(".md")]
    out += [f for f in ("README.md", "DESIGN.md", "ROUND2.md", "COMPONENTS.md", "CHANGELOG.md", "CONTRIBUTING.md", "benchmarks/README.md",
                        "docs/BUILDING.md", "docs/README.md") if os.path.exists(os.path.join(ROOT, f))]
    for sub in ("profiles", "examples"):
        for base, _, files in os.walk(os.path.join(ROOT, sub)):
            out += [os.path.relpath(os.path.join(base, f), ROOT) for f in sorted(files) if f.endswith(".md")]
    return out


@pytest.mark.parametrize("name", _doc_files())
def test_paths_mentioned_in_the_docs_exist(name):
    """Backticked repository paths (``examples/...py``, ``tests/...py``, ``scripts/...``, ``csrc/...``) point at files
    that exist."""
    with open(os.path.join(ROOT, name)) as f:
        text = f.read()
    missing = []
    for path in set(re.findall(r"`((?:examples|tests|scripts|benchmarks|profiles|docs|bench|byzpy_b200)/[A-Za-z0-9_./-]+\.(?:py|sh|md|yaml|json|cu|cuh|h|cpp|csv|sass))`", text)):
        if not os.path.exists(os.path.join(ROOT, path)):
            missing.append(path)
    for path in set(re.findall(r"`(csrc/[A-Za-z0-9_./-]+\.(?:cu|cuh|h|cpp))`", text)):
        if not os.path.exists(os.path.join(ROOT, "byzpy_b200", path)):
            missing.append(path)
    assert not missing, f"{name}: {sorted(missing)}"
