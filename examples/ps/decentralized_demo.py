"""Legacy building block demo: a prototype parameter server over blocking ``NodeRunner`` processes
(``ParameterServerRunner``): three worker processes compute gradients on ``step``, the server process
aggregates its inbox with a coordinate-wise median (counterpart of the reference's
examples/ps/decentralized_demo.py).

    python examples/ps/decentralized_demo.py [--transport local|tcp]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from byzpy_b200.engine.parameter_server.runner import ParameterServerRunner  # noqa: E402
from byzpy_b200.engine.transport.local import LocalTransport  # noqa: E402
from byzpy_b200.engine.transport.tcp import TcpTransport  # noqa: E402


def g_honest_a():
    return torch.tensor([1.0, 2.0, 3.0])


def g_honest_b():
    return torch.tensor([1.2, 1.8, 3.1])


def g_byzantine():
    return torch.tensor([-100.0, 100.0, -100.0])


def median(grads):
    return torch.stack(list(grads)).median(dim=0).values


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", choices=["local", "tcp"], default="local",
                    help="how the node processes exchange messages: in-process queues or loopback sockets")
    a = ap.parse_args()
    transport = LocalTransport() if a.transport == "local" else TcpTransport()
    runner = ParameterServerRunner([g_honest_a, g_honest_b, g_byzantine], aggregator=median, transport=transport)
    runner.start()
    try:
        for r in range(3):
            print(f"round {r}: aggregate = {runner.run_round().tolist()}")
    finally:
        runner.stop()
        if hasattr(transport, "close"):
            transport.close()
