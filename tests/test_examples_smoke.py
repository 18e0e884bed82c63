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
    (["examples/distributed/mnist.py", "--local", "--rounds", "2", "--num-honest", "4", "--num-byz", "1", "--f", "1",
      "--eval-interval", "1"], "[round 0002] test loss="),
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


# ------------------------------------------------------------------ remote-TCP examples (several processes each)
def _env(**extra):
    return dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="",
                OMP_NUM_THREADS="2", **extra)


def _free_ports(k):
    import socket

    socks = [socket.socket() for _ in range(k)]
    for s in socks:
        s.bind(("127.0.0.1", 0))
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _finish(procs, timeout):
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
            out += "\n<<killed after timeout>>"
        outs.append(out)
    return outs


def test_remote_tcp_parameter_server_survives_a_leaving_worker_and_rejects_a_bad_mac(tmp_path):
    """examples/ps/remote_tcp/ps_node.py: 1 server + 4 workers (one Byzantine); worker w2 drops its connection
    after round 2 and the training finishes with the remaining three; a client signing with the wrong secret
    never gets in."""
    import yaml

    (port,) = _free_ports(1)
    cfg = {"server": {"host": "127.0.0.1", "port": port}, "rounds": 4, "round_timeout": 30, "lr": 0.05,
           "aggregator": {"name": "trimmed_mean", "f": 1},
           "workers": [{"id": "w0", "role": "honest"}, {"id": "w1", "role": "honest"},
                       {"id": "w2", "role": "honest", "leave_after": 2}, {"id": "w3", "role": "byzantine"}]}
    path = tmp_path / "nodes.yaml"
    path.write_text(yaml.safe_dump(cfg))
    script = os.path.join("examples", "ps", "remote_tcp", "ps_node.py")
    env = _env(BYZPY_HMAC_SECRET="rehearsal-secret")
    popen = lambda args, e=env: subprocess.Popen([sys.executable, script] + args + ["--config", str(path)], cwd=ROOT,  # noqa: E731
                                                 env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    server = popen(["server"])
    intruder = popen(["worker", "--id", "w0"], _env(BYZPY_HMAC_SECRET="wrong-secret"))
    intruder_out = _finish([intruder], 120)[0]
    workers = [popen(["worker", "--id", w]) for w in ("w0", "w1", "w2", "w3")]
    outs = _finish(workers + [server], 240)
    srv_out = outs[-1]
    assert server.returncode == 0, srv_out[-2000:]
    assert "rejected connection: bad HMAC" in srv_out, srv_out[-2000:]
    assert intruder.returncode != 0, intruder_out[-500:]          # the server closed the socket on it
    assert "[round 2] 4 gradients" in srv_out and "worker w2 left" in srv_out, srv_out[-2000:]
    assert "[round 4] 3 gradients" in srv_out, srv_out[-2000:]
    assert "[w2] leaving before round 3" in outs[2] and all("finished" in o for o in outs[:2] + outs[3:4]), outs


def _p2p_cfg(tmp_path, rounds=2):
    import yaml

    ports = _free_ports(5)
    cfg = {"server": {"host": "127.0.0.1", "port": ports[4]}, "topology": "complete", "rounds": rounds,
           "nodes": [{"id": str(i), "host": "127.0.0.1", "port": ports[i], "role": "honest" if i < 3 else "byzantine"}
                     for i in range(4)]}
    path = tmp_path / "nodes.yaml"
    path.write_text(yaml.safe_dump(cfg))
    return str(path)


def test_remote_tcp_mesh_example_four_processes(tmp_path):
    """examples/p2p/remote_tcp/mesh_client.py: every node is a TCP server plus clients to all peers."""
    cfg = _p2p_cfg(tmp_path)
    procs = [subprocess.Popen([sys.executable, "examples/p2p/remote_tcp/mesh_client.py", "--config", cfg, "--node-id", str(i)],
                              cwd=ROOT, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for i in range(4)]
    outs = _finish(procs, 240)
    assert all(p.returncode == 0 for p in procs), [o[-800:] for o in outs]
    finals = [ln for o in outs[:3] for ln in o.splitlines() if "round 2:" in ln]
    assert len(finals) == 3 and all("3 neighbour vectors" in ln for ln in finals), outs
    assert len({ln.split("|theta| = ")[1] for ln in finals}) == 1        # the honest nodes agree after aggregation
    assert "round 2: attacked with 3 honest vectors" in outs[3], outs[3][-800:]


def test_remote_tcp_hub_example_server_and_four_clients(tmp_path):
    """examples/p2p/remote_tcp/server.py + client.py: hub-and-spoke relay through a RemoteNodeServer."""
    cfg = _p2p_cfg(tmp_path)
    hub = subprocess.Popen([sys.executable, "examples/p2p/remote_tcp/server.py", "--config", cfg], cwd=ROOT, env=_env(),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        import time

        time.sleep(2.0)
        procs = [subprocess.Popen([sys.executable, "examples/p2p/remote_tcp/client.py", "--config", cfg, "--node-id", str(i)],
                                  cwd=ROOT, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for i in range(4)]
        outs = _finish(procs, 240)
    finally:
        hub.terminate()
        hub_out = _finish([hub], 30)[0]
    assert all(p.returncode == 0 for p in procs), [o[-800:] for o in outs] + [hub_out[-800:]]
    finals = [ln for o in outs[:3] for ln in o.splitlines() if "round 2:" in ln]
    assert len(finals) == 3 and all("3 neighbour vectors" in ln for ln in finals), outs
    assert "hub listening" in hub_out


def test_remote_tcp_hub_example_reference_command_line():
    """The hub example driven the way the reference's README does it: no node list, the server given --host/--port,
    every client told the shape of the network (--server-host --total-nodes --honest-nodes --node-type ...)."""
    import time

    (port,) = _free_ports(1)
    hub = subprocess.Popen([sys.executable, "examples/p2p/remote_tcp/server.py", "--host", "127.0.0.1", "--port", str(port)],
                           cwd=ROOT, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        time.sleep(2.0)
        common = ["--server-host", "127.0.0.1", "--server-port", str(port), "--total-nodes", "3", "--honest-nodes", "2",
                  "--rounds", "2", "--batch-size", "32", "--lr", "0.05", "--seed", "1"]
        argv = [["--node-id", "0", "--node-type", "honest", "--data-shard", "0"],
                ["--node-id", "1", "--node-type", "honest", "--data-shard", "1"],
                ["--node-id", "2", "--node-type", "byzantine", "--byz-scale", "-2.0"]]
        procs = [subprocess.Popen([sys.executable, "examples/p2p/remote_tcp/client.py"] + common + extra, cwd=ROOT,
                                  env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for extra in argv]
        outs = _finish(procs, 240)
    finally:
        hub.terminate()
        hub_out = _finish([hub], 30)[0]
    assert all(p.returncode == 0 for p in procs), [o[-800:] for o in outs] + [hub_out[-800:]]
    finals = [ln for o in outs[:2] for ln in o.splitlines() if "round 2:" in ln]
    assert len(finals) == 2 and all("2 neighbour vectors" in ln for ln in finals), outs
    assert "round 2: attacked with 2 honest vectors" in outs[2], outs[2][-800:]


def test_remote_tcp_parameter_server_reference_command_line(tmp_path):
    """ps_node.py with the reference's spelling: --role, --worker-id as a position, --worker-type / --rounds
    overriding the node list."""
    import yaml

    (port,) = _free_ports(1)
    cfg = {"server": {"host": "127.0.0.1", "port": port}, "rounds": 9, "round_timeout": 30, "lr": 0.05,
           "aggregator": {"name": "trimmed_mean", "f": 1},
           "workers": [{"id": f"w{i}", "role": "honest"} for i in range(4)]}
    path = tmp_path / "nodes.yaml"
    path.write_text(yaml.safe_dump(cfg))
    script = os.path.join("examples", "ps", "remote_tcp", "ps_node.py")
    popen = lambda args: subprocess.Popen([sys.executable, script, "--config", str(path), "--rounds", "2"] + args,  # noqa: E731
                                          cwd=ROOT, env=_env(BYZPY_HMAC_SECRET="s"), stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)
    server = popen(["--role", "server"])
    workers = [popen(["--role", "worker", "--worker-id", str(i)] + (["--worker-type", "byzantine"] if i == 3 else []))
               for i in range(4)]
    outs = _finish(workers + [server], 240)
    assert server.returncode == 0 and all(w.returncode == 0 for w in workers), [o[-800:] for o in outs]
    assert "[round 2] 4 gradients" in outs[-1] and "[round 3]" not in outs[-1], outs[-1][-1500:]
    assert all(f"[w{i}] finished" in outs[i] for i in range(4)), outs[:4]
