"""Legacy building block demo: gossip averaging over a ring of blocking ``NodeRunner`` processes
driven by ``NodeCluster`` (counterpart of the reference's examples/p2p/decentralized_demo.py, which is
stale against its own API at the surveyed commit; this one runs).

Each node holds a scalar; on every ``step`` it averages the values in its inbox with its own; the
launcher forwards every node's value to its ring neighbours between steps.

    python examples/p2p/decentralized_demo.py [--transport local|tcp]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from byzpy_b200.engine.node_cluster import NodeCluster  # noqa: E402
from byzpy_b200.engine.transport.local import LocalTransport  # noqa: E402
from byzpy_b200.engine.transport.tcp import TcpTransport  # noqa: E402


def step(state: dict) -> dict:
    inbox = state.get("inbox") or []
    if inbox:
        state["value"] = (state["value"] + sum(inbox)) / (1 + len(inbox))
        state["inbox"] = []
    return state


def on_msg(state: dict, msg) -> dict:
    state.setdefault("inbox", []).append(float(msg))
    return state


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", choices=["local", "tcp"], default="local",
                    help="how the node processes exchange messages: in-process queues or loopback sockets")
    a = ap.parse_args()
    transport = LocalTransport() if a.transport == "local" else TcpTransport()
    n = 4
    cluster = NodeCluster(transport=transport)
    for i in range(n):
        cluster.add_node(f"n{i}", step, on_msg, init_state={"value": float(10 * i)})
    cluster.start_all()
    try:
        for r in range(6):
            values = [cluster.state(f"n{i}")["value"] for i in range(n)]
            print(f"round {r}: " + "  ".join(f"{v:7.3f}" for v in values))
            for i in range(n):
                for j in ((i - 1) % n, (i + 1) % n):
                    cluster.send(f"n{j}", values[i])
            for i in range(n):
                cluster._nodes[f"n{i}"].step()
    finally:
        cluster.stop_all()
        if hasattr(transport, "close"):
            transport.close()
