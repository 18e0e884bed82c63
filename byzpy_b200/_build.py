"""In-tree build of the sm_100a kernel library ``byzpy_b200/_C*.so``.

Every translation unit under ``byzpy_b200/csrc`` is compiled by ``nvcc`` with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles on a box
with no GPU) and linked into one pybind11 extension that lives inside the
package directory, so it travels with a repo snapshot and is visible to the
"which .so did the process load" audit.

Usage: ``python -m byzpy_b200._build [--force] [--verbose]``
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR / "csrc" / "build"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def ext_path() -> Path:
    """Where the built kernel library lives: ``byzpy_b200/_C<EXT_SUFFIX>`` inside the source tree."""
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return PKG_DIR / f"_C{suffix}"


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build byzpy_b200._C")
    return cand


def sources() -> list[Path]:
    """Every ``.cu`` / ``.cpp`` file under ``csrc/``, sorted."""
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _headers() -> list[Path]:
    return sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile (if stale) and return the path of the extension module."""
    import pybind11

    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    py_inc = sysconfig.get_paths()["include"]
    common = [
        nvcc,
        *ARCH_FLAGS,
        "-lineinfo",
        "-O3",
        "-std=c++17",
        "--expt-relaxed-constexpr",
        "-Xcompiler",
        "-fPIC,-fvisibility=hidden",
        f"-I{CSRC}",
        f"-I{py_inc}",
        f"-I{pybind11.get_include()}",
    ]
    hdrs = _headers()
    jobs = []
    objs = []
    for src in sources():
        obj = OBJ_DIR / (src.name + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *hdrs]):
            cmd = [*common, "-x", "cu", "-c", str(src), "-o", str(obj)]
            jobs.append((src, cmd))

    def _run(job):
        src, cmd = job
        if verbose:
            print("[byzpy_b200 build]", " ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_run, jobs))
    out = ext_path()
    if force or jobs or _stale(out, objs):
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(out), *map(str, objs), "-lcudart", "-ldl"]
        if verbose:
            print("[byzpy_b200 build]", " ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(p)
