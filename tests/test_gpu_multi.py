"""Multi-GPU tests (skipped unless >= 2 CUDA devices): launch torchrun subprocesses."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("agg", ["median", "trmean"])
def test_fused_round_two_ranks(agg):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "multi_gpu", "check_fused_round.py"), "--agg", agg]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_FUSED_ROUND PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("agg,topology", [("trmean", "complete"), ("gm", "complete"), ("trmean", "ring")])
def test_p2p_round_two_ranks(agg, topology):
    """Device gossip round across ranks: remote vectors staged over NVLink once per round, robust
    aggregation per peer, compared with an NCCL all_gather + plain-tensor reference."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29534",
           os.path.join(ROOT, "tests", "multi_gpu", "check_p2p_round.py"), "--agg", agg, "--topology", topology]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_P2P_ROUND PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("workers,agg", [(7, "multikrum"), (5, "median")])
def test_fused_round_uneven_worker_counts(workers, agg):
    """RowLayout.spread: worker counts that do not divide the world size (4 + 3, 3 + 2 replicas)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29535",
           os.path.join(ROOT, "tests", "multi_gpu", "check_fused_round.py"), "--agg", agg, "--workers", str(workers)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_FUSED_ROUND PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_fused_round_rank_without_replica():
    """One honest replica + two virtual Little rows on two ranks: rank 1 hosts no replica and only takes
    part in the aggregation of its coordinate shard."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29536",
           os.path.join(ROOT, "tests", "multi_gpu", "check_fused_round.py"), "--agg", "median", "--workers", "3",
           "--attack", "little"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_FUSED_ROUND PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("agg,graph", [("median", 1), ("trmean", 0)])
def test_fused_round_bucketed_overlap_two_ranks(agg, graph):
    """The round as a sequence of bucket launches enqueued from inside backward, with per-bucket
    sequence numbers in the cross-GPU flag words, against the independent fp64 oracle."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29537",
           os.path.join(ROOT, "tests", "multi_gpu", "check_fused_round.py"), "--agg", agg, "--buckets", "3",
           "--graph", str(graph), "--steps", "5"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_FUSED_ROUND PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    assert "overlapped=True" in res.stdout and "buckets=1 " not in res.stdout, res.stdout[-1500:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_fused_round_survives_a_silent_rank():
    """Fault injection: a rank stops taking part, the survivor's kernels time out and name
    it, ParameterServer.recover() drops its rows and training continues on the remaining ones."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29539",
           os.path.join(ROOT, "tests", "multi_gpu", "check_fault.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "MULTI_GPU_FAULT PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
