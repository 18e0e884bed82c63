"""Hub for the hub-and-spoke remote P2P example: a ``RemoteNodeServer`` that relays messages between
the ``RemoteContext`` clients (examples/p2p/remote_tcp/client.py).

    python examples/p2p/remote_tcp/server.py --config examples/p2p/remote_tcp/nodes_example.yaml
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import load_config  # noqa: E402

from byzpy_b200.engine.node.remote_server import RemoteNodeServer  # noqa: E402


async def main(cfg):
    srv = RemoteNodeServer(cfg["server"]["host"], int(cfg["server"]["port"]))
    await srv.start()
    print(f"hub listening on {cfg['server']['host']}:{srv.port}", flush=True)
    await srv.serve()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.yaml"))
    asyncio.run(main(load_config(ap.parse_args().config)))
