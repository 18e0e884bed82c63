"""Actor server for the heterogeneous parameter-server example (examples/ps/heterogenous/mnist.py): hosts node actors that a driver on another machine (or
another terminal) places with the ``tcp://host:port`` backend spec (``ucx://host:port`` with
``--gpu-direct``: CUDA tensors travel as CUDA-IPC handles instead of pickled host copies).

    python examples/ps/heterogenous/server.py --host 0.0.0.0 --port 29000
    python examples/ps/heterogenous/mnist.py --servers 127.0.0.1:29000
"""
import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.engine.actor.backends.gpu import start_ucx_actor_server  # noqa: E402
from byzpy_b200.engine.actor.backends.remote import start_actor_server  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=29000)
    ap.add_argument("--gpu-direct", action="store_true", help="serve the ucx:// scheme (CUDA-IPC tensor payloads)")
    a = ap.parse_args()
    serve = start_ucx_actor_server if a.gpu_direct else start_actor_server
    print(f"[actor-server] {'ucx' if a.gpu_direct else 'tcp'}://{a.host}:{a.port}", flush=True)
    asyncio.run(serve(a.host, a.port))


if __name__ == "__main__":
    main()
