"""SymmetricBuffer's handle-exchange protocol on CPU ranks (gloo, world 3) with a fake driver: the VMM heap with
its NVLS multicast alias, and the team-wide fallbacks when a driver call fails on SOME rank -- every rank must
end up on the same heap, nobody may be left waiting in a collective, nothing may leak."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "multi_cpu", "symm_fallback_worker.py")


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fail, multicast="none", world=3):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), WORKER, json.dumps(fail), multicast]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert lines, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    return json.loads(lines[-1][7:])


def _uniform(res, key):
    vals = {json.dumps(r.get(key)) for r in res}
    assert len(vals) == 1, (key, res)
    return res[0].get(key)


def test_vmm_heap_with_multicast_alias():
    res = _run({})
    assert _uniform(res, "kind") == "vmm" and _uniform(res, "mc") is True
    assert all(r["ptrs_ok"] and r["distinct"] and r["leaked_maps"] == 0 and r["leaked_handles"] == 0 for r in res), res


@pytest.mark.parametrize("call,ranks", [("mc_create", [0]), ("mc_add_device", [2]), ("mc_bind", [1]),
                                        ("vmm_map_mc", [0, 2]), ("no_mc_support", [1])])
def test_multicast_failure_on_some_rank_drops_multicast_everywhere(call, ranks):
    res = _run({call: ranks})
    assert _uniform(res, "kind") == "vmm" and _uniform(res, "mc") is False, res
    assert all(r["ptrs_ok"] and r["leaked_maps"] == 0 and r["leaked_handles"] == 0 for r in res), res


@pytest.mark.parametrize("call,ranks", [("vmm_alloc", [1]), ("vmm_export_fd", [2]), ("vmm_import_fd", [0]),
                                        ("vmm_map", [1, 2])])
def test_vmm_failure_on_some_rank_moves_the_team_to_the_ipc_heap(call, ranks):
    res = _run({call: ranks})
    assert _uniform(res, "kind") == "ipc" and _uniform(res, "mc") is False, res
    assert all(r["ptrs_ok"] and r["distinct"] and r["leaked_maps"] == 0 and r["leaked_handles"] == 0 for r in res), res


def test_explicit_multicast_request_fails_loudly_on_every_rank():
    res = _run({"mc_bind": [2]}, multicast="true")
    assert all("error" in r and "multicast=True" in r["error"] for r in res), res


def test_multicast_can_be_declined():
    res = _run({}, multicast="false")
    assert _uniform(res, "kind") == "vmm" and _uniform(res, "mc") is False


@pytest.mark.parametrize("world", [1, 2])
def test_other_team_sizes(world):
    """world 1 (the single-GPU bench: no exchange, no multicast) and world 2."""
    res = _run({}, world=world)
    assert _uniform(res, "kind") == "vmm" and _uniform(res, "mc") is (world > 1)
    assert all(r["ptrs_ok"] and r["distinct"] and r["leaked_maps"] == 0 and r["leaked_handles"] == 0 for r in res), res


def test_single_rank_allocation_failure_falls_back_to_the_ipc_heap():
    res = _run({"vmm_alloc": [0]}, world=1)
    assert res[0].get("kind") == "ipc" and res[0]["ptrs_ok"] and res[0]["leaked_maps"] == 0, res
