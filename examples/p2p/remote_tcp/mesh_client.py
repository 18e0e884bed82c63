"""One node of the full-mesh remote P2P example: every node is a TCP server plus clients to all
peers (``MeshRemoteContext``; dead peers are re-dialled, sends fall back outbound -> inbound).

    for i in 0 1 2 3; do python examples/p2p/remote_tcp/mesh_client.py --node-id $i & done; wait
    python examples/p2p/remote_tcp/mesh_client.py --config nodes.yaml --node-id 0 --node-type honest --rounds 50 --lr 0.05
    # on one NVSwitch box add --gpu-direct to ship CUDA tensors as CUDA-IPC handles
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import add_training_flags, apply_training_flags, gossip, load_config, make_node  # noqa: E402

from byzpy_b200.engine.node.context import MeshRemoteContext  # noqa: E402


async def main(cfg, node_id, gpu_direct):
    entry = next(e for e in cfg["nodes"] if str(e["id"]) == node_id)
    peers = {str(e["id"]): (e["host"], int(e["port"])) for e in cfg["nodes"] if str(e["id"]) != node_id}
    ctx = MeshRemoteContext(entry["host"], int(entry["port"]), peers, reconnect_interval=0.5,
                            gpu_direct=gpu_direct)
    node = make_node(cfg, node_id, ctx)
    await node.start()
    try:
        await gossip(node, cfg, entry.get("role", "honest"), settle=3.0)
    finally:
        await asyncio.sleep(0.5)
        await node.shutdown()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.json"))
    ap.add_argument("--node-id", required=True)
    ap.add_argument("--gpu-direct", action="store_true")
    add_training_flags(ap)
    a = ap.parse_args()
    asyncio.run(main(apply_training_flags(load_config(a.config), a, str(a.node_id)), str(a.node_id), a.gpu_direct))
