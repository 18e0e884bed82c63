"""Runs ``tests/gpu_late/late_cases.py`` -- GPU tests of OPT-IN code written after the round's GPU budget was spent,
never executed on a B200 -- in a child interpreter with a hard time limit, and reports its outcome counts.

Why a child: the cases are ``xfail(strict=False)``, but an xfail mark cannot contain a sticky CUDA error (every later
test of the process would fail) or a kernel that never returns (``pytest-timeout`` cannot interrupt a thread blocked in
``cudaStreamSynchronize``).  In a child both are bounded: this wrapper kills the process group after ``LIMIT_S`` and the
tier of the measured kernels (everything that sorts before this file) keeps its verdict.  The child's full report goes
to ``gpurun_out/late_gpu_cases.txt``; the summary line is repeated as a warning so that it shows in ``pytest -q``."""
import os
import signal
import subprocess
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIMIT_S = 600


@pytest.mark.gpu
@pytest.mark.timeout(LIMIT_S + 120)
def test_late_opt_in_gpu_cases_run_in_a_child_interpreter():
    cmd = [sys.executable, "-m", "pytest", os.path.join("tests", "gpu_late", "late_cases.py"), "-q", "-m", "gpu", "-rA",
           "-p", "no:cacheprovider"]
    proc = subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True, env=dict(os.environ, PYTHONPATH=ROOT))
    try:
        out, _ = proc.communicate(timeout=LIMIT_S)
        note = f"exit code {proc.returncode}"
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, _ = proc.communicate()
        note = f"killed after {LIMIT_S} s"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "late_gpu_cases.txt"), "w") as fh:
        fh.write(out or "")
    lines = [ln for ln in (out or "").splitlines() if ln.strip()]
    summary = lines[-1] if lines else "(no output)"
    warnings.warn(f"late opt-in GPU cases ({note}): {summary}")
    print("\n".join(lines[-80:]))
    if proc.returncode not in (0, 1) or "killed" in note:
        pytest.xfail(f"the late opt-in cases did not run to completion ({note}): {summary}")
