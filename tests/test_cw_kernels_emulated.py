"""The coordinate-wise CUDA kernels executed on the HOST by a warp-lockstep emulator (tests/native/cuda_host_emu.h):
every lane a coroutine, __syncwarp() a real barrier, cp.async performed either at the covering wait_group (latest
legal moment: exposes reads that are not ordered behind the copy) or at issue (earliest: exposes slots refilled under
their readers), shared memory poisoned.  The GPU-validated direct and staged kernels must reproduce (that validates
the emulator); the warp-tiled kernel, which no GPU has run yet, gets its functional test here.  A mutant of the tiled
kernel with one barrier removed must FAIL, so the check is not vacuous."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "byzpy_b200", "csrc")
NATIVE = os.path.join(ROOT, "tests", "native")
FLAGS = ["-std=c++17", "-O1", "-DBZ_HOST_EMU", "-Wno-unknown-pragmas", "-Wno-attributes"]


def _gxx():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    return gxx


def _build(inc, out):
    res = subprocess.run([_gxx(), *FLAGS, "-I", inc, "-I", NATIVE, os.path.join(NATIVE, "cw_kernels_emu.cpp"), "-o", out],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]


def test_kernels_reproduce_under_both_cp_async_schedules(tmp_path):
    exe = str(tmp_path / "cw_kernels_emu")
    _build(CSRC, exe)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]
    lines = res.stdout.strip().splitlines()
    assert sum("warp-tiled kernel" in ln and " 0 failures, 0 emulator errors" in ln for ln in lines) == 2, res.stdout
    assert sum("self-check" in ln and " 0 failures" in ln for ln in lines) == 2, res.stdout
    total = lines[-1]
    assert total.startswith("total:") and int(total.split()[1]) > 250_000 and " 0 failures, 0 emulator errors" in total


@pytest.mark.parametrize("name,old,new,mode", [
    ("refill-under-readers", "    __syncwarp();                                                // all reads done before the slot is refilled\n", "",
     "eager"),
    ("read-before-the-copies-landed", "    __syncwarp();                                                // every lane's copies of this slot have landed\n",
     "", "deferred"),
])
def test_a_kernel_with_a_barrier_removed_is_caught(tmp_path, name, old, new, mode):
    inc = tmp_path / "csrc"
    inc.mkdir()
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")) or f == "cw_select.cu":
            shutil.copy(os.path.join(CSRC, f), inc / f)
    src = (inc / "cw_select.cu").read_text()
    assert src.count(old) == 1, name
    (inc / "cw_select.cu").write_text(src.replace(old, new))
    exe = str(tmp_path / "mutant")
    _build(str(inc), exe)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert res.returncode != 0
    line = next(ln for ln in res.stdout.splitlines() if ln.startswith(f"[{mode} cp.async] warp-tiled kernel"))
    assert " 0 failures" not in line, res.stdout
    # the untouched kernels still pass in the mutant build
    assert all(" 0 failures" in ln for ln in res.stdout.splitlines() if "self-check" in ln), res.stdout
