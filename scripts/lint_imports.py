"""Tiny stand-in for pyflakes (no linters in the image): reports imported names a module never uses
and names listed in ``__all__`` that the module does not define.

    python scripts/lint_imports.py [paths ...]        # default: the package, bench.py, examples, benchmarks
"""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(path):
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src, path)
    imported = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                imported[(a.asname or a.name).split(".")[0]] = node.lineno
        elif isinstance(node, ast.ImportFrom):
            if node.module == "__future__":
                continue
            for a in node.names:
                if a.name != "*":
                    imported[a.asname or a.name] = node.lineno
    used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)}
    strings = {n.value for n in ast.walk(tree) if isinstance(n, ast.Constant) and isinstance(n.value, str)}
    defined = set(imported) | {n.name for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef))}
    defined |= {t.id for n in ast.walk(tree) if isinstance(n, (ast.Assign, ast.AnnAssign, ast.AugAssign))
                for t in ast.walk(n.targets[0] if isinstance(n, ast.Assign) else n.target) if isinstance(t, ast.Name)}
    problems = []
    is_init = os.path.basename(path) == "__init__.py"
    for name, line in sorted(imported.items(), key=lambda kv: kv[1]):
        if name not in used and name not in strings and not is_init and "noqa" not in src.splitlines()[line - 1]:
            problems.append(f"{path}:{line}: unused import {name}")
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            if isinstance(node.value, (ast.List, ast.Tuple)):
                for elt in node.value.elts:
                    if isinstance(elt, ast.Constant) and elt.value not in defined and "globals()" not in src \
                            and "def __getattr__" not in src:
                        problems.append(f"{path}:{node.lineno}: __all__ names undefined {elt.value!r}")
    return problems


def main():
    targets = sys.argv[1:] or [os.path.join(ROOT, p) for p in ("byzpy_b200", "bench.py", "__graft_entry__.py", "examples",
                                                              "benchmarks", "bench", "scripts", "tests")]
    out = []
    for t in targets:
        if os.path.isfile(t):
            out += check(t)
            continue
        for dirpath, _, files in os.walk(t):
            for f in sorted(files):
                if f.endswith(".py"):
                    out += check(os.path.join(dirpath, f))
    print("\n".join(os.path.relpath(p, ROOT) if p.startswith(ROOT) else p for p in out) or "clean")
    return 1 if out else 0


if __name__ == "__main__":
    sys.exit(main())
