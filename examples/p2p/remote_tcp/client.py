"""One node of the hub-and-spoke remote P2P example (``RemoteContext`` -> ``RemoteNodeServer``).

    python examples/p2p/remote_tcp/server.py &
    for i in 0 1 2 3; do python examples/p2p/remote_tcp/client.py --node-id $i & done; wait

or without a node list, every client told the shape of the network (the reference's command line,
examples/p2p/remote_tcp/README.md:40-52):

    python examples/p2p/remote_tcp/server.py --host 0.0.0.0 --port 8888
    python examples/p2p/remote_tcp/client.py --server-host localhost --server-port 8888 \
        --node-id 0 --node-type honest --total-nodes 3 --honest-nodes 2 --data-shard 0
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import add_training_flags, apply_training_flags, gossip, load_config, make_node  # noqa: E402

from byzpy_b200.engine.node.context import RemoteContext  # noqa: E402


async def main(cfg, node_id):
    entry = next(e for e in cfg["nodes"] if str(e["id"]) == node_id)
    node = make_node(cfg, node_id, RemoteContext(cfg["server"]["host"], int(cfg["server"]["port"])))
    await node.start()
    try:
        await gossip(node, cfg, entry.get("role", "honest"), settle=2.0)
    finally:
        await node.shutdown()


def config_from_flags(a) -> dict:
    """The node list the flags describe: ids 0..total-1, the first ``--honest-nodes`` of them honest."""
    if a.total_nodes is None or a.honest_nodes is None:
        raise SystemExit("--server-host needs --total-nodes and --honest-nodes (or use --config)")
    return {"server": {"host": a.server_host, "port": a.server_port}, "topology": "complete", "rounds": 50,
            "nodes": [{"id": str(i), "role": "honest" if i < a.honest_nodes else "byzantine"}
                      for i in range(a.total_nodes)]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.yaml"))
    ap.add_argument("--node-id", required=True)
    ap.add_argument("--server-host", default=None, help="hub address; given: the network comes from the flags below")
    ap.add_argument("--server-port", type=int, default=8888)
    ap.add_argument("--total-nodes", type=int, default=None)
    ap.add_argument("--honest-nodes", type=int, default=None)
    ap.add_argument("--data-shard", type=int, default=None, help="which honest data shard this node trains on")
    add_training_flags(ap)
    a = ap.parse_args()
    cfg = config_from_flags(a) if a.server_host is not None else load_config(a.config)
    asyncio.run(main(apply_training_flags(cfg, a, str(a.node_id)), str(a.node_id)))
