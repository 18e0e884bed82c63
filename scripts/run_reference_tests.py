"""Run the reference's OWN test-suite against this package through the import alias.

The reference ships its tests inside the package (``byzpy/**/tests/test_*.py``).  This script stages
them, unmodified, in a scratch directory (nothing is copied into this repository), makes ``import byzpy``
resolve to ``byzpy_b200`` in the test process and in every child interpreter (``sitecustomize``), provides
a minimal stand-in for ``pytest-asyncio`` when that plugin is not installed, and runs pytest.

    python scripts/run_reference_tests.py --reference /path/to/byzpy/python/byzpy [-k expr] [--keep]

It is a compatibility probe, not part of the regular suite: the expected residue is listed in
docs/source/testing.md (UCX-only tests).  ``BYZPY_P2P_CONTEXT=process`` is set for the run: one process per P2P
node is the reference's default and its suite asserts it; here it is opt-in.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

SITECUSTOMIZE = '''import sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, {stage!r})
try:
    import byzpy_b200.compat as _compat
    _compat.install_alias()
except Exception as exc:  # pragma: no cover
    print("byzpy alias could not be installed:", exc)
'''

ASYNCIO_SHIM = '''"""Minimal pytest-asyncio stand-in: one event loop shared by async fixtures and async tests."""
import asyncio, functools, inspect
import pytest

LOOP = asyncio.new_event_loop()
asyncio.set_event_loop(LOOP)


def fixture(*dargs, **dkw):
    def deco(fn):
        if inspect.isasyncgenfunction(fn):
            @functools.wraps(fn)
            def wrapper(*a, **k):
                gen = fn(*a, **k)
                yield LOOP.run_until_complete(gen.__anext__())
                try:
                    LOOP.run_until_complete(gen.__anext__())
                except StopAsyncIteration:
                    pass
            return pytest.fixture(**dkw)(wrapper)
        if inspect.iscoroutinefunction(fn):
            @functools.wraps(fn)
            def wrapper(*a, **k):
                return LOOP.run_until_complete(fn(*a, **k))
            return pytest.fixture(**dkw)(wrapper)
        return pytest.fixture(**dkw)(fn)
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return deco(dargs[0])
    return deco
'''

CONFTEST = '''import inspect
import pytest

try:
    import pytest_asyncio
    HAVE_PLUGIN = hasattr(pytest_asyncio, "plugin")
except ImportError:  # pragma: no cover
    pytest_asyncio, HAVE_PLUGIN = None, False


def pytest_configure(config):
    config.addinivalue_line("markers", "asyncio: coroutine test")
    config.addinivalue_line("markers", "real_actor_backends: run pool tests against real backends")


if not HAVE_PLUGIN:
    @pytest.hookimpl(tryfirst=True)
    def pytest_pyfunc_call(pyfuncitem):
        fn = pyfuncitem.obj
        if inspect.iscoroutinefunction(fn):
            kwargs = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
            pytest_asyncio.LOOP.run_until_complete(fn(**kwargs))
            return True
'''


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--reference", required=True, help="path of the reference's package directory (…/python/byzpy)")
    ap.add_argument("-k", default=None, help="pytest -k expression")
    ap.add_argument("--timeout", type=int, default=120, help="per-test timeout (needs pytest-timeout)")
    ap.add_argument("--keep", action="store_true", help="keep the staging directory")
    a, extra = ap.parse_known_args()
    ref = Path(a.reference).resolve()
    tests = sorted(p for p in ref.rglob("test_*.py") if "tests" in p.parts)
    if not tests:
        print(f"no tests found under {ref}", file=sys.stderr)
        return 2
    stage = Path(tempfile.mkdtemp(prefix="byzpy_reftests_"))
    shim = stage / "_site"
    shim.mkdir()
    try:
        for p in tests:                       # flat, collision-free names: engine_graph__test_pool.py
            rel = p.relative_to(ref).as_posix().replace("/tests/", "__").replace("/", "_")
            shutil.copy(p, stage / rel)
        (shim / "sitecustomize.py").write_text(SITECUSTOMIZE.format(repo=str(REPO), stage=str(stage)))
        (stage / "conftest.py").write_text(CONFTEST)
        try:
            import pytest_asyncio  # noqa: F401
        except ImportError:
            (stage / "pytest_asyncio.py").write_text(ASYNCIO_SHIM)
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([str(shim), env.get("PYTHONPATH", "")]).rstrip(os.pathsep)
        # the suite asserts the reference's defaults; the one default that differs here is switchable
        env.setdefault("BYZPY_P2P_CONTEXT", "process")
        cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-W", "ignore",
               *[str(f) for f in sorted(stage.glob("*test_*.py"))]]
        try:
            import pytest_timeout  # noqa: F401

            cmd += ["--timeout", str(a.timeout)]
        except ImportError:
            pass
        if a.k:
            cmd += ["-k", a.k]
        print(f"[reference tests] {len(tests)} files staged in {stage}", flush=True)
        return subprocess.call(cmd + extra, cwd=stage, env=env)
    finally:
        if not a.keep:
            shutil.rmtree(stage, ignore_errors=True)


if __name__ == "__main__":
    raise SystemExit(main())
