#!/usr/bin/env python
"""Mini undefined-name check (pyflakes stand-in; no network to install one).

For every ``*.py`` file: collect every name bound ANYWHERE in the file (imports, defs, classes, assignment /
loop / with / except / comprehension / walrus targets, arguments, ``global`` / ``nonlocal``) plus the builtins,
and report names that are read but never bound.  Coarser than real scope analysis -- it cannot see a name used
outside the scope that binds it -- but it catches the failure that matters for code paths the CPU suite cannot
execute (GPU-only branches): a typo or a missing import that would raise ``NameError`` on the device box.

    python scripts/lint_names.py [paths...]        # default: the package, tests, bench / benchmarks / examples
"""
from __future__ import annotations

import ast
import builtins
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ["byzpy_b200", "tests", "bench", "benchmarks", "examples", "scripts", "bench.py", "__graft_entry__.py"]
SKIP_DIRS = {"__pycache__", "_ref", "build", ".git"}


def _targets(node, out):
    if isinstance(node, ast.Name):
        out.add(node.id)
    elif isinstance(node, (ast.Tuple, ast.List)):
        for e in node.elts:
            _targets(e, out)
    elif isinstance(node, ast.Starred):
        _targets(node.value, out)


def check(path: str):
    src = open(path, encoding="utf-8").read()
    try:
        tree = ast.parse(src, path)
    except SyntaxError as exc:
        return [(exc.lineno or 0, f"syntax error: {exc.msg}")]
    bound = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__package__", "__class__"}
    star = False
    for n in ast.walk(tree):
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                if a.name == "*":
                    star = True
                bound.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(n.name)
            if not isinstance(n, ast.ClassDef):
                a = n.args
                for arg in a.posonlyargs + a.args + a.kwonlyargs + [x for x in (a.vararg, a.kwarg) if x]:
                    bound.add(arg.arg)
        elif isinstance(n, ast.Lambda):
            a = n.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + [x for x in (a.vararg, a.kwarg) if x]:
                bound.add(arg.arg)
        elif isinstance(n, (ast.Assign,)):
            for t in n.targets:
                _targets(t, bound)
        elif isinstance(n, (ast.AugAssign, ast.AnnAssign)):
            _targets(n.target, bound)
        elif isinstance(n, (ast.For, ast.AsyncFor, ast.comprehension)):
            _targets(n.target, bound)
        elif isinstance(n, (ast.With, ast.AsyncWith)):
            for it in n.items:
                if it.optional_vars is not None:
                    _targets(it.optional_vars, bound)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, ast.NamedExpr):
            _targets(n.target, bound)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
        elif isinstance(n, ast.MatchAs) and n.name:
            bound.add(n.name)
    if star:
        return []
    bad = []
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound:
            bad.append((n.lineno, f"undefined name {n.id!r}"))
    return bad


def main(argv):
    paths = argv or [os.path.join(ROOT, p) for p in DEFAULT]
    files = []
    for p in paths:
        if os.path.isfile(p):
            files.append(p)
            continue
        for dp, dn, fn in os.walk(p):
            dn[:] = [d for d in dn if d not in SKIP_DIRS]
            files += [os.path.join(dp, f) for f in fn if f.endswith(".py")]
    n = 0
    for f in sorted(files):
        for line, msg in check(f):
            print(f"{os.path.relpath(f, ROOT)}:{line}: {msg}")
            n += 1
    print("clean" if n == 0 else f"{n} finding(s)")
    return 1 if n else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
