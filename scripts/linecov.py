"""Line coverage of ``byzpy_b200`` under pytest without third-party tooling (uses ``sys.monitoring``,
Python >= 3.12).  Child processes are not traced.

    python scripts/linecov.py [pytest args ...]        # default: tests -q -m "not gpu"
"""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "byzpy_b200") + os.sep
sys.path.insert(0, ROOT)

hit = {}
mon = sys.monitoring
TOOL = mon.COVERAGE_ID


def on_line(code, line):
    fn = code.co_filename
    if fn.startswith(PKG):
        hit.setdefault(fn, set()).add(line)
    return mon.DISABLE          # one report per code location is enough


def executable_lines(path):
    with open(path) as fh:
        tree = ast.parse(fh.read())
    lines = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.stmt) and not isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            if isinstance(node, ast.Expr) and isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
                continue         # docstring
            lines.add(node.lineno)
    return lines


def main():
    import pytest

    mon.use_tool_id(TOOL, "linecov")
    mon.register_callback(TOOL, mon.events.LINE, on_line)
    mon.set_events(TOOL, mon.events.LINE)
    args = sys.argv[1:] or ["tests", "-q", "-m", "not gpu", "-p", "no:warnings"]
    rc = pytest.main(args)
    mon.set_events(TOOL, 0)
    rows, tot_e, tot_h = [], 0, 0
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            ex = executable_lines(path)
            got = hit.get(path, set()) & ex
            tot_e += len(ex)
            tot_h += len(got)
            if ex:
                rows.append((len(got) / len(ex), len(ex) - len(got), os.path.relpath(path, ROOT), sorted(ex - got)))
    rows.sort()
    for frac, miss, path, missing in rows:
        print(f"{100 * frac:5.1f}%  miss {miss:4d}  {path}")
    print(f"TOTAL {100 * tot_h / max(1, tot_e):.1f}%  ({tot_h}/{tot_e} statements)")
    for suffix in filter(None, os.environ.get("LINECOV_DETAIL", "").split(",")):   # e.g. "engine/node/context.py,ops/reference.py"
        for frac, miss, path, missing in rows:
            if path.endswith(suffix):
                print(path, missing)
    return rc


if __name__ == "__main__":
    sys.exit(main())
