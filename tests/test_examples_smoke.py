"""The shipped examples keep running (CPU, tiny round counts, one subprocess each)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    (["examples/ps/thread/mnist.py", "--rounds", "2"], "final:"),
    (["examples/p2p/thread/mnist.py", "--rounds", "2"], "node0 test loss"),
    (["examples/p2p/decentralized_process_mnist.py", "--rounds", "2"], "node0 test loss"),
    (["examples/p2p/decentralized_autonomous_mnist.py", "--rounds", "1"], "'rounds': 1"),
    (["examples/ps/decentralized_demo.py"], "aggregate = [1.0, 2.0, 3.0]"),
    (["examples/p2p/decentralized_demo.py"], "round 5:"),
    (["examples/distributed/mnist.py", "--local", "--rounds", "2"], "|aggregate|"),
    (["benchmarks/config1_cpu_plumbing.py", "--repeat", "3"], '"pool_x4_ms"'),
]


@pytest.mark.parametrize("argv,expect", CASES, ids=[c[0][0] for c in CASES])
def test_example_runs(argv, expect):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable] + argv, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert expect in res.stdout, res.stdout[-2000:]
