"""Hub for the hub-and-spoke remote P2P example: a ``RemoteNodeServer`` that relays messages between
the ``RemoteContext`` clients (examples/p2p/remote_tcp/client.py).

    python examples/p2p/remote_tcp/server.py --config examples/p2p/remote_tcp/nodes_example.yaml
    python examples/p2p/remote_tcp/server.py --host 0.0.0.0 --port 8888        # no node list needed
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
    ap.add_argument("--host", default=None, help="bind address (default: the node list's server.host)")
    ap.add_argument("--port", type=int, default=None, help="bind port (default: the node list's server.port)")
    a = ap.parse_args()
    cfg = {"server": {"host": a.host, "port": a.port}} if None not in (a.host, a.port) else load_config(a.config)
    if a.host is not None:
        cfg["server"]["host"] = a.host
    if a.port is not None:
        cfg["server"]["port"] = a.port
    asyncio.run(main(cfg))
