"""One node of the hub-and-spoke remote P2P example (``RemoteContext`` -> ``RemoteNodeServer``).

    python examples/p2p/remote_tcp/server.py &
    for i in 0 1 2 3; do python examples/p2p/remote_tcp/client.py --node-id $i & done; wait
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import gossip, load_config, make_node  # noqa: E402

from byzpy_b200.engine.node.context import RemoteContext  # noqa: E402


async def main(cfg, node_id):
    entry = next(e for e in cfg["nodes"] if str(e["id"]) == node_id)
    node = make_node(cfg, node_id, RemoteContext(cfg["server"]["host"], int(cfg["server"]["port"])))
    await node.start()
    try:
        await gossip(node, cfg, entry.get("role", "honest"), settle=2.0)
    finally:
        await node.shutdown()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.yaml"))
    ap.add_argument("--node-id", required=True)
    a = ap.parse_args()
    asyncio.run(main(load_config(a.config), str(a.node_id)))
