"""Peer process of tests/test_gpu_control_plane.py::test_mesh_contexts_move_cuda_tensors...:
node "B" of a two-node mesh with GPU-direct payloads; doubles every vector it receives and sends it back.

    python tests/multi_gpu/mesh_peer.py <my_port> <peer_port>
"""
import asyncio
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from byzpy_b200.engine.node.context import MeshRemoteContext  # noqa: E402


class _Node:
    node_id = "B"


async def main(my_port: int, peer_port: int) -> None:
    torch.cuda.init()
    ctx = MeshRemoteContext("127.0.0.1", my_port, {"A": ("127.0.0.1", peer_port)}, gpu_direct=True)
    await ctx.start(_Node())
    keep = []
    async for msg in ctx.receive_messages():
        if msg.get("type") != "vec":
            continue
        v = msg["payload"]["vector"]
        out = v * 2
        torch.cuda.synchronize()
        keep.append(out)                      # the reply's memory must outlive the send
        await ctx.send_message("A", "echo", {"vector": out, "was_cuda": bool(v.is_cuda), "pid": os.getpid()})


if __name__ == "__main__":
    asyncio.run(main(int(sys.argv[1]), int(sys.argv[2])))
