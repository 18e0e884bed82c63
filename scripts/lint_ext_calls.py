#!/usr/bin/env python
"""Check every Python call into the native extension against the pybind11 signature.

The kernels' Python callers live mostly in code the CPU suite cannot execute (device rounds, symmetric heaps), and
their argument lists are long and positional.  This walks the package / tests / benches, finds calls of the form
``<something ending in ext | _C | _ext>.NAME(...)`` where ``NAME`` is a function of ``byzpy_b200._C``, and
compares the number of positional arguments and the keyword names with the signature pybind11 records in the
function's docstring.  Calls with ``*args`` / ``**kwargs`` are skipped.

    python scripts/lint_ext_calls.py
"""
from __future__ import annotations

import ast
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RECEIVERS = {"ext", "_C", "_ext", "C"}


def signatures():
    from byzpy_b200 import _C

    sigs = {}
    for name in dir(_C):
        fn = getattr(_C, name)
        doc = getattr(fn, "__doc__", None)
        if not callable(fn) or not doc:
            continue
        overloads = []
        for line in doc.splitlines():
            m = re.match(rf"\s*(?:\d+\.\s*)?{re.escape(name)}\((.*)\)\s*->", line)
            if not m:
                continue
            params, depth, cur = [], 0, ""
            for ch in m.group(1):
                if ch in "[(":
                    depth += 1
                elif ch in "])":
                    depth -= 1
                if ch == "," and depth == 0:
                    params.append(cur.strip())
                    cur = ""
                else:
                    cur += ch
            if cur.strip():
                params.append(cur.strip())
            names = [p.split(":")[0].strip() for p in params]
            required = sum(1 for p in params if "=" not in p.split(":", 1)[-1])
            overloads.append((names, required))
        if overloads:
            sigs[name] = overloads
    return sigs


def receiver_name(node):
    if isinstance(node, ast.Name):
        return node.id
    if isinstance(node, ast.Attribute):
        return node.attr
    return None


def main():
    sigs = signatures()
    findings = 0
    checked = 0
    for top in ("byzpy_b200", "tests", "bench", "benchmarks", "examples", "bench.py", "__graft_entry__.py"):
        p = os.path.join(ROOT, top)
        files = [p] if os.path.isfile(p) else [os.path.join(dp, f) for dp, dn, fn in os.walk(p) for f in fn
                                               if f.endswith(".py") and "_ref" not in dp]
        for f in sorted(files):
            tree = ast.parse(open(f, encoding="utf-8").read(), f)
            for n in ast.walk(tree):
                if not (isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)):
                    continue
                name = n.func.attr
                if name not in sigs or receiver_name(n.func.value) not in RECEIVERS:
                    continue
                if any(isinstance(a, ast.Starred) for a in n.args) or any(k.arg is None for k in n.keywords):
                    continue
                checked += 1
                npos, kws = len(n.args), [k.arg for k in n.keywords]
                ok = False
                why = ""
                for names, required in sigs[name]:
                    if npos > len(names):
                        why = f"{npos} positional arguments, signature takes {len(names)}"
                        continue
                    unknown = [k for k in kws if k not in names[npos:]]
                    if unknown:
                        why = f"unknown / duplicate keyword(s) {unknown}"
                        continue
                    supplied = npos + len(kws)
                    missing = [nm for i, nm in enumerate(names[:required]) if i >= npos and nm not in kws]
                    if missing:
                        why = f"missing required argument(s) {missing} ({supplied} supplied, {required} required)"
                        continue
                    ok = True
                    break
                if not ok:
                    findings += 1
                    print(f"{os.path.relpath(f, ROOT)}:{n.lineno}: {name}(): {why}")
    print(f"{checked} extension calls checked; " + ("clean" if not findings else f"{findings} finding(s)"))
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
