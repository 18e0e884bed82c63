"""Command line interface (reference cli.py:14-168; console script ``byzpy-b200``).

    byzpy-b200 version
    byzpy-b200 doctor [--format json]          python / torch+CUDA / kernel extension / NVLink peers
    byzpy-b200 list {aggregators,attacks,pre-aggregators} [--format json] [--short]
    byzpy-b200 build [--force]                  compile the sm_100a kernel library in-tree
    byzpy-b200 bench -- <bench.py args>         run the headline benchmark
"""
from __future__ import annotations

import argparse
import importlib
import inspect
import json
import os
import pkgutil
import platform
import subprocess
import sys
from typing import List, Optional, Sequence

from . import __version__


def _load_subclasses(package: str, base: type) -> List[str]:
    """Fully qualified names (``package.module.Class``, as the reference's CLI prints them) of the concrete
    ``base`` subclasses defined under ``package`` (test packages and private modules skipped)."""
    pkg = importlib.import_module(package)
    found = set()
    for info in pkgutil.walk_packages(pkg.__path__, prefix=pkg.__name__ + "."):
        if ".tests" in info.name or info.name.rsplit(".", 1)[-1].startswith("_"):
            continue
        try:
            module = importlib.import_module(info.name)
        except Exception:
            continue
        for _, obj in inspect.getmembers(module, inspect.isclass):
            if issubclass(obj, base) and obj is not base and not inspect.isabstract(obj) \
                    and obj.__module__.startswith(package) \
                    and not obj.__module__.endswith(".base"):
                found.add(f"{obj.__module__}.{obj.__name__}")
    return sorted(found)


def _doctor() -> dict:
    report = {
        "python_version": platform.python_version(),
        "platform": platform.platform(),
        "byzpy_b200": __version__,
        "torch": {"available": False, "cuda": False},
        "kernels": {"available": False},
        "nvlink": {},
    }
    try:
        import torch

        t = report["torch"]
        t["available"] = True
        t["version"] = torch.__version__
        t["cuda"] = bool(torch.cuda.is_available())
        if t["cuda"]:
            t["cuda_device_count"] = torch.cuda.device_count()
            t["devices"] = [torch.cuda.get_device_name(i) for i in range(torch.cuda.device_count())]
            t["capability"] = list(torch.cuda.get_device_capability(0))
    except Exception as exc:  # pragma: no cover
        report["torch"]["error"] = str(exc)
    try:
        from . import ops

        ext = ops._load_ext()
        k = report["kernels"]
        k["available"] = ext is not None
        if ext is not None:
            k["arch"] = ext.ARCH
            k["path"] = getattr(ext, "__file__", None)
            k["entry_points"] = sorted(n for n in dir(ext) if not n.startswith("_") and n.islower())
            k["host_kernels"] = [n for n in ("host_cw_select", "host_colstat") if hasattr(ext, n)]
            if report["torch"].get("cuda"):
                n = ext.device_count()
                report["nvlink"]["peer_access"] = [[bool(i == j or ext.can_access_peer(i, j))
                                                    for j in range(n)] for i in range(n)]
            # which symmetric heap DeviceRound / DeviceP2PRound will use on this box (parallel/symmetric.py)
            if hasattr(ext, "vmm_support"):
                try:
                    sup = dict(ext.vmm_support(0))
                except Exception as exc:  # noqa: BLE001
                    sup = {"vmm": False, "posix_fd": False, "multicast": False, "reason": repr(exc)}
                heap = "vmm" if (sup.get("vmm") and sup.get("posix_fd")) else "ipc"
                report["symmetric_heap"] = {"heap": heap, "nvls_multicast": bool(sup.get("multicast")) and heap == "vmm",
                                            "tma_tensor_maps": sup.get("reason", "") == "",
                                            "driver": sup.get("reason") or "ok"}
        else:
            k["error"] = repr(ops._C_err)
    except Exception as exc:  # pragma: no cover
        report["kernels"]["error"] = str(exc)
    # process-wide switches that change which code path runs (all optional; shown with their effective value)
    import os

    report["switches"] = {
        "BYZPY_SYMM": os.environ.get("BYZPY_SYMM", "auto (vmm where supported, else ipc)"),
        "BYZPY_CW_IMPL": os.environ.get("BYZPY_CW_IMPL", "auto (direct / staged; 'tiled' is opt-in)"),
        "BYZPY_FUSED_MAPCW": os.environ.get("BYZPY_FUSED_MAPCW", "0 (pre-aggregator -> coordinate-wise fused round on fused=True only)"),
        "BYZPY_GRAM_TMA": os.environ.get("BYZPY_GRAM_TMA", "1"),
        "BYZPY_GRAM_CENTER": os.environ.get("BYZPY_GRAM_CENTER", "off"),
        "BYZPY_POOL_DISPATCH": os.environ.get("BYZPY_POOL_DISPATCH", "adaptive"),
        "BYZPY_OVERLAP_GRID": os.environ.get("BYZPY_OVERLAP_GRID", "auto (a quarter of the SMs)"),
        "BYZPY_BN_CLUSTER": os.environ.get("BYZPY_BN_CLUSTER", "1"),
        "BYZPY_INTRAOP_GOVERNOR": os.environ.get("BYZPY_INTRAOP_GOVERNOR", "1"),
    }
    return report


def _cmd_version(_: argparse.Namespace) -> int:
    print(__version__)
    return 0


def _cmd_doctor(args: argparse.Namespace) -> int:
    data = _doctor()
    if args.format == "json":
        print(json.dumps(data, indent=2, sort_keys=True))
        return 0
    for key, value in data.items():
        if isinstance(value, dict):
            print(f"{key}:")
            for k, v in value.items():
                print(f"  - {k}: {v}")
        else:
            print(f"{key}: {value}")
    return 0


def _cmd_list(args: argparse.Namespace) -> int:
    if args.component == "aggregators":
        from .aggregators.base import Aggregator as base

        items = _load_subclasses("byzpy_b200.aggregators", base)
    elif args.component == "attacks":
        from .attacks.base import Attack as base

        items = _load_subclasses("byzpy_b200.attacks", base)
    else:
        from .pre_aggregators.base import PreAggregator as base

        items = _load_subclasses("byzpy_b200.pre_aggregators", base)
    from . import compat

    if compat._finder is not None:
        # the ``byzpy`` import alias is active (``python -m byzpy.cli``, or a program that installed it): print the
        # names the way that program can import them
        items = sorted("byzpy" + name[len("byzpy_b200"):] for name in items)
    if args.short:
        items = sorted({name.rsplit(".", 1)[-1] for name in items})
    if args.format == "json":
        print(json.dumps({"component": args.component, "items": items}, indent=2))
    else:
        if not items:
            print(f"No {args.component} found.")
        for name in items:
            print(name)
    return 0


def _cmd_build(args: argparse.Namespace) -> int:
    from . import _build

    print(_build.build(force=args.force, verbose=args.verbose))
    return 0


def _cmd_bench(args: argparse.Namespace) -> int:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "bench.py")
    if not os.path.exists(script):
        print("bench.py not found next to the package", file=sys.stderr)
        return 2
    return subprocess.call([sys.executable, script, *args.rest])


def build_parser() -> argparse.ArgumentParser:
    """The ``byzpy-b200`` argument parser: ``version``, ``doctor``, ``list``, ``build``, ``bench`` (the first three as in the
    reference's CLI, reference cli.py).
    """
    parser = argparse.ArgumentParser(prog="byzpy-b200",
                                     description="Utilities for inspecting byzpy_b200 installations.")
    sub = parser.add_subparsers(dest="command", required=True)
    sub.add_parser("version", help="Print the installed version.").set_defaults(func=_cmd_version)
    doc = sub.add_parser("doctor", help="Diagnose local dependencies.")
    doc.add_argument("--format", choices=("human", "json"), default="human")
    doc.set_defaults(func=_cmd_doctor)
    lst = sub.add_parser("list", help="List built-in components.")
    lst.add_argument("component", choices=("aggregators", "attacks", "pre-aggregators"))
    lst.add_argument("--format", choices=("human", "json"), default="human")
    lst.add_argument("--short", action="store_true", help="class names only instead of package.module.Class")
    lst.set_defaults(func=_cmd_list)
    bld = sub.add_parser("build", help="Compile the sm_100a kernel library in-tree.")
    bld.add_argument("--force", action="store_true")
    bld.add_argument("--verbose", action="store_true")
    bld.set_defaults(func=_cmd_build)
    bch = sub.add_parser("bench", help="Run the headline benchmark (arguments after --).")
    bch.add_argument("rest", nargs=argparse.REMAINDER)
    bch.set_defaults(func=_cmd_bench)
    return parser


def main(argv: Optional[Sequence[str]] = None) -> int:
    """Entry point of the ``byzpy-b200`` console script; returns the exit status (``argv`` defaults to ``sys.argv[1:]``)."""
    args = build_parser().parse_args(argv)
    if getattr(args, "rest", None) and args.rest and args.rest[0] == "--":
        args.rest = args.rest[1:]
    return int(args.func(args))


if __name__ == "__main__":  # pragma: no cover
    raise SystemExit(main())
