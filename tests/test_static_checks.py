"""Static checks that stand in for the device box: the GPU-only Python paths cannot run here, but a NameError
or a native call whose argument list drifted from the pybind11 signature can be found without a GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script)], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_no_undefined_names_anywhere():
    r = _run("lint_names.py")
    assert r.returncode == 0, r.stdout[-2000:]


def test_no_unused_imports():
    r = _run("lint_imports.py")
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout[-2000:]


def test_native_call_sites_match_the_pybind_signatures():
    pytest.importorskip("byzpy_b200._C")
    r = _run("lint_ext_calls.py")
    assert r.returncode == 0, r.stdout[-2000:]
    assert int(r.stdout.strip().splitlines()[-1].split()[0]) > 50      # the walker really found the call sites


def test_api_reference_lists_every_public_module():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_api_reference.py"), "--check"],
                       capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]


def test_every_documented_module_imports():
    """Each ``automodule`` target of docs/source/api_reference.md is importable on a CPU-only box (autodoc imports
    them one by one; a module that needs a GPU at import time would break the docs build)."""
    import importlib
    import re

    text = open(os.path.join(ROOT, "docs", "source", "api_reference.md")).read()
    mods = re.findall(r"automodule:: (\S+)", text)
    assert len(mods) > 100
    bad = []
    for m in mods:
        try:
            importlib.import_module(m)
        except Exception as exc:  # noqa: BLE001
            bad.append((m, repr(exc)))
    assert not bad, bad
