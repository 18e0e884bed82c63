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
    (["benchmarks/pytorch/sign_flip_actor_pool.py", "--num-grads", "8", "--grad-dim", "4096", "--pool-workers", "2",
      "--repeat", "1"], '"op": "sign-flip"'),
    (["benchmarks/pytorch/mimic_actor_pool.py", "--num-grads", "8", "--grad-dim", "4096", "--pool-workers", "2",
      "--repeat", "1"], '"op": "mimic"'),
    (["benchmarks/pytorch/bucketing_actor_pool.py", "--num-grads", "32", "--grad-dim", "4096", "--pool-workers", "2",
      "--repeat", "1"], '"op": "bucketing"'),
    (["benchmarks/byzfl/ipm_attack_compare.py", "--num-grads", "8", "--grad-dim", "4096"], '"byzpy_b200_ms"'),
    (["benchmarks/byzfl/parameter_server_multikrum_compare.py", "--rounds", "1", "--honest", "4", "--byzantine", "1"],
     '"ms_per_round"'),
    (["benchmarks/scheduler/pipeline_benchmark.py", "--branches", "2", "--num-grads", "8", "--grad-dim", "2000",
      "--repeat", "1"], '"parallel_scheduler_ms"'),
]


@pytest.mark.parametrize("argv,expect", CASES, ids=[c[0][0] for c in CASES])
def test_example_runs(argv, expect):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable] + argv, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert expect in res.stdout, res.stdout[-2000:]


def test_distributed_rehearsal_script_runs_two_servers_and_a_driver():
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="",
               ROUNDS="1", PORT_A="29311", PORT_B="29312")
    res = subprocess.run(["bash", "examples/distributed/test_local.sh"], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "distributed parameter-server rehearsal finished" in res.stdout


def test_example_actor_servers_parse_their_arguments():
    for script in ("examples/p2p/remote/server.py", "examples/p2p/heterogeneous/server.py",
                   "examples/ps/heterogenous/server.py"):
        res = subprocess.run([sys.executable, script, "--help"], cwd=ROOT, capture_output=True, text=True, timeout=120,
                             env=dict(os.environ, PYTHONPATH=ROOT))
        assert res.returncode == 0 and "--gpu-direct" in res.stdout, res.stderr[-500:]
