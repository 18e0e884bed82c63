"""Every benchmark / example script that the reference also ships accepts the reference's command-line flags.

The reference's documentation is written as command lines (benchmarks/README.md, examples/*/README.md); a user who
switches must be able to paste them.  For each script that exists under the same relative path in both trees, the
option strings of the reference's ``add_argument`` calls are read from its source and looked up in the ``--help`` of the
script here.  Skipped when the reference tree is not present.
"""
import glob
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BYZPY_REFERENCE_ROOT", "/root/reference")

_FLAG = re.compile(r'add_argument\(\s*"(--[A-Za-z0-9-]+)"')


def _pairs():
    out = []
    for sub in ("benchmarks", "examples"):
        for ref_path in sorted(glob.glob(os.path.join(REF, sub, "**", "*.py"), recursive=True)):
            rel = os.path.relpath(ref_path, REF)
            if not os.path.exists(os.path.join(ROOT, rel)):
                continue
            with open(ref_path) as f:
                flags = sorted(set(_FLAG.findall(f.read())))
            if flags:
                out.append((rel, flags))
    return out


PAIRS = _pairs()


@pytest.mark.skipif(not PAIRS, reason="reference tree not available")
@pytest.mark.parametrize("rel,flags", PAIRS, ids=[p[0] for p in PAIRS])
def test_script_accepts_the_reference_flags(rel, flags):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, rel, "--help"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-1500:]
    missing = [f for f in flags if not re.search(re.escape(f) + r"\b", res.stdout)]
    assert not missing, f"{rel} does not accept {missing}"


def test_reference_documented_invocations_parse():
    """Command lines quoted in the reference's benchmarks/README.md, shrunk to toy sizes, run to completion."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    small = ["--pool-workers", "2", "--pool-backend", "thread", "--warmup", "1", "--repeat", "1"]
    cases = [
        ["benchmarks/pytorch/multikrum_actor_pool.py", "--num-grads", "20", "--grad-dim", "512", "--f", "4", "--q", "5",
         "--chunk-size", "8"] + small,
        ["benchmarks/pytorch/geometric_median_actor_pool.py", "--num-grads", "12", "--grad-dim", "512", "--tol", "1e-5",
         "--max-iter", "16", "--init", "mean", "--chunk-size", "4"] + small,
        ["benchmarks/pytorch/bucketing_actor_pool.py", "--num-vectors", "32", "--dim", "512", "--bucket-size", "4",
         "--feature-chunk", "128"] + small,
        ["benchmarks/pytorch/clipping_preagg.py", "--num-vectors", "16", "--dim", "512", "--threshold", "1.5",
         "--chunk-size", "4"] + small,
        ["benchmarks/pytorch/little_actor_pool.py", "--num-grads", "16", "--grad-dim", "512", "--f", "3", "--N", "16"]
        + small,
        ["benchmarks/byzfl/centered_clipping_compare.py", "--num-grads", "12", "--grad-dim", "512", "--tau", "0.3",
         "--iters", "3", "--warmup", "1", "--repeat", "1", "--seed", "3"],
        ["benchmarks/pytorch/actor_pool_python.py", "--tasks", "32", "--inner-iters", "200", "--chunk-size", "8",
         "--pool-workers", "2", "--pool-backend", "thread", "--warmup", "0", "--repeat", "1", "--seed", "1"],
    ]
    for argv in cases:
        res = subprocess.run([sys.executable] + argv, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, (argv, res.stderr[-1500:])
        assert res.stdout.strip().splitlines()[-1].startswith("{"), (argv, res.stdout[-500:])
