"""Standalone TCP parameter server / honest worker / Byzantine worker (no actors, no scheduler):
the minimal multi-host deployment, one process per role, HMAC-authenticated frames.

Frame = 4-byte big-endian length | 32-byte HMAC-SHA256(secret, body) | body (pickle).  The secret
comes from ``BYZPY_HMAC_SECRET`` (same variable as the reference example); frames with a bad MAC are
dropped before unpickling.  Protocol (server <-> worker):

    worker -> {"type": "hello", "id": ...}                      server -> {"type": "init_model", "state_dict": {k: cpu}}
    server -> {"type": "round", "round": r}                     worker -> {"type": "gradient", "round": r, "vector": g}
    server -> {"type": "update", "round": r, "vector": agg}     (workers apply it with their own optimizer)
    server -> {"type": "done"}

A round is skipped when fewer than the aggregator's minimum of gradients arrive within
``round_timeout`` seconds (counterpart of the reference's examples/ps/remote_tcp/ps_node.py).

    python examples/ps/remote_tcp/ps_node.py server &
    for w in w0 w1 w2 w3; do python examples/ps/remote_tcp/ps_node.py worker --id $w & done; wait
"""
from __future__ import annotations

import argparse
import asyncio
import hashlib
import hmac
import os
import pickle
import struct
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.attacks import EmpireAttack  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.parallel.arena import flatten_grads, write_vector_to_grads_  # noqa: E402
from byzpy_b200.utils.data import batch_source, evaluate, mnist_like, shard_indices  # noqa: E402

SECRET = os.environ.get("BYZPY_HMAC_SECRET", "change-me").encode()


# ------------------------------------------------------------------------------ framing
async def send_frame(writer: asyncio.StreamWriter, obj) -> None:
    body = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    mac = hmac.new(SECRET, body, hashlib.sha256).digest()
    writer.write(struct.pack(">I", len(body)) + mac + body)
    await writer.drain()


async def recv_frame(reader: asyncio.StreamReader):
    (n,) = struct.unpack(">I", await reader.readexactly(4))
    mac = await reader.readexactly(32)
    body = await reader.readexactly(n)
    if not hmac.compare_digest(mac, hmac.new(SECRET, body, hashlib.sha256).digest()):
        raise PermissionError("bad HMAC: frame dropped")
    return pickle.loads(body)


def load_config(path):
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


def build_aggregator(spec):
    name = spec.get("name", "median")
    if name == "median":
        return CoordinateWiseMedian()
    if name == "trimmed_mean":
        return CoordinateWiseTrimmedMean(f=int(spec.get("f", 1)))
    if name == "multi_krum":
        return MultiKrum(f=int(spec.get("f", 1)), q=int(spec.get("q", 2)))
    raise ValueError(f"unknown aggregator {name!r}")


# ------------------------------------------------------------------------------ server
async def run_server(cfg):
    agg = build_aggregator(cfg.get("aggregator", {}))
    expected = {str(w["id"]) for w in cfg["workers"]}
    torch.manual_seed(0)
    model = SmallCNN()
    conns = {}
    all_in = asyncio.Event()

    async def on_conn(reader, writer):
        try:
            hello = await recv_frame(reader)
        except Exception as exc:  # noqa: BLE001
            print("rejected connection:", exc, flush=True)
            writer.close()
            return
        wid = str(hello["id"])
        conns[wid] = (reader, writer)
        await send_frame(writer, {"type": "init_model",
                                  "state_dict": {k: v.cpu() for k, v in model.state_dict().items()}})
        if expected <= set(conns):
            all_in.set()

    srv = await asyncio.start_server(on_conn, cfg["server"]["host"], int(cfg["server"]["port"]))
    print(f"parameter server on {cfg['server']['host']}:{cfg['server']['port']}, waiting for {len(expected)} workers",
          flush=True)
    await all_in.wait()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    xt, yt = mnist_like(2000, train=False)
    timeout = float(cfg.get("round_timeout", 60))
    for r in range(1, int(cfg.get("rounds", 5)) + 1):
        for _, w in conns.values():
            await send_frame(w, {"type": "round", "round": r})

        async def one(wid):
            msg = await recv_frame(conns[wid][0])
            return msg["vector"] if msg.get("round") == r else None

        done, pending = await asyncio.wait([asyncio.ensure_future(one(w)) for w in conns], timeout=timeout)
        for p in pending:
            p.cancel()
        grads = [d.result() for d in done if not d.exception() and d.result() is not None]
        try:
            vec = agg.aggregate(grads)
        except ValueError as exc:       # not enough gradients for this aggregator
            print(f"[round {r}] skipped ({len(grads)} gradients: {exc})", flush=True)
            continue
        write_vector_to_grads_(model, vec)
        opt.step()
        for _, w in conns.values():
            await send_frame(w, {"type": "update", "round": r, "vector": vec})
        loss, acc = evaluate(model, xt, yt, torch.device("cpu"))
        print(f"[round {r}] {len(grads)} gradients  test loss={loss:.4f} acc={acc:.4f}", flush=True)
    for _, w in conns.values():
        await send_frame(w, {"type": "done"})
        w.close()
    srv.close()


# ------------------------------------------------------------------------------ workers
async def run_worker(cfg, wid):
    entry = next(w for w in cfg["workers"] if str(w["id"]) == wid)
    honest_ids = [str(w["id"]) for w in cfg["workers"] if w.get("role", "honest") == "honest"]
    for attempt in range(50):
        try:
            reader, writer = await asyncio.open_connection(cfg["server"]["host"], int(cfg["server"]["port"]))
            break
        except OSError:
            await asyncio.sleep(0.2)
    else:
        raise ConnectionError("parameter server not reachable")
    await send_frame(writer, {"type": "hello", "id": wid})
    init = await recv_frame(reader)
    model = SmallCNN()
    model.load_state_dict(init["state_dict"], strict=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    lossf = torch.nn.CrossEntropyLoss()
    byz = entry.get("role", "honest") != "honest"
    if byz:
        me = 0
        attack = EmpireAttack(scale=-1.0)
    else:
        me = honest_ids.index(wid)
    x, y = mnist_like(6000)
    idx = torch.as_tensor(shard_indices(6000, max(1, len(honest_ids)))[me])
    nxt = batch_source(x[idx], y[idx], 64, seed=me)
    while True:
        msg = await recv_frame(reader)
        if msg["type"] == "done":
            break
        if msg["type"] == "round":
            xb, yb = nxt()
            model.zero_grad(set_to_none=True)
            lossf(model(xb), yb).backward()
            g = flatten_grads(model)
            if byz:     # a non-omniscient Empire: scale its own honest gradient
                g = torch.as_tensor(attack.apply(honest_grads=[g]))
            await send_frame(writer, {"type": "gradient", "round": msg["round"], "vector": g})
        elif msg["type"] == "update":
            write_vector_to_grads_(model, msg["vector"])
            opt.step()
    writer.close()
    print(f"[{wid}] finished", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("role", choices=["server", "worker"])
    ap.add_argument("--id", default="w0")
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.yaml"))
    a = ap.parse_args()
    cfg = load_config(a.config)
    asyncio.run(run_server(cfg) if a.role == "server" else run_worker(cfg, a.id))
