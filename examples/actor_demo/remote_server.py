"""Start a TCP actor server (use ``tcp://host:port`` as an actor/pool backend against it).

    python examples/actor_demo/remote_server.py --host 0.0.0.0 --port 29000 [--gpu-direct]
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from byzpy_b200.engine.actor.backends.gpu import start_ucx_actor_server  # noqa: E402
from byzpy_b200.engine.actor.backends.remote import start_actor_server  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=29000)
    ap.add_argument("--gpu-direct", action="store_true", help="ucx:// scheme: CUDA-IPC tensor payloads")
    a = ap.parse_args()
    asyncio.run((start_ucx_actor_server if a.gpu_direct else start_actor_server)(a.host, a.port))
